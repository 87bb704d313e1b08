// Microbenchmark behind DESIGN.md "K_B fill": what is the ceiling for streaming ~6 GB of fp32 rows of 1036 bytes
// (259 floats: rows are only 4-byte aligned) out of the SMs, and which store path gets closest?
//   (a) st.global.cs.f32, lane <-> consecutive floats of one row (what fill_rows_kernel does)
//   (b) st.global.cs.v4.f32 on a 16-byte aligned stream (upper bound for LSU stores)
//   (c) shared-memory staging + cp.async.bulk.global.shared (TMA store) of 16-byte aligned spans
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bench store_paths.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROW = 259;

__global__ void __launch_bounds__(256) rows_scalar(float *out, size_t rows) {
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (size_t)gridDim.x * 8;
    for (size_t r = warp; r < rows; r += nw) {
        float *dst = out + r * ROW;
#pragma unroll
        for (int j = 0; j < 8; ++j) __stcs(dst + 3 + j * 32 + lane, (float)j);
        if (lane < 3) dst[lane] = 1.0f;
    }
}

__global__ void __launch_bounds__(256) stream_v4(float4 *out, size_t n4) {
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = i0; i < n4; i += stride) __stcs(out + i, make_float4(1, 2, 3, 4));
}

// each warp owns consecutive spans of SPAN bytes of the output stream: fill a smem buffer, then one bulk store
template <int SPAN>
__global__ void __launch_bounds__(256) stream_tma(float *out, size_t bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *buf = reinterpret_cast<float *>(smem + (size_t)w * SPAN);
    const size_t warp = (size_t)blockIdx.x * 8 + w, nw = (size_t)gridDim.x * 8;
    const size_t spans = bytes / SPAN;
    for (size_t s = warp; s < spans; s += nw) {
        // wait until the previous bulk store has finished READING the buffer
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll 4
        for (int i = lane; i < SPAN / 4; i += 32) buf[i] = (float)i;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char *>(out) + s * SPAN),
                         "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"(SPAN)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const size_t rows = 5760000;
    const size_t bytes = rows * ROW * 4;     // 5.97 GB
    float *out;
    cudaMalloc(&out, bytes + 4096);
    const int blocks = 148 * 8;
    float ms = timeit([&] { rows_scalar<<<blocks, 256>>>(out, rows); }, 10);
    printf("(a) scalar 4-byte row stores   %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    ms = timeit([&] { stream_v4<<<blocks, 256>>>(reinterpret_cast<float4 *>(out), bytes / 16); }, 10);
    printf("(b) aligned float4 stream      %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    cudaFuncSetAttribute(stream_tma<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096);
    ms = timeit([&] { stream_tma<4096><<<148 * 6, 256, 8 * 4096>>>(out, bytes); }, 10);
    printf("(c) smem + TMA bulk store 4 KB %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    cudaFuncSetAttribute(stream_tma<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
    ms = timeit([&] { stream_tma<8192><<<148 * 3, 256, 8 * 8192>>>(out, bytes); }, 10);
    printf("(c) smem + TMA bulk store 8 KB %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
