// Microbenchmark behind DESIGN.md "K_A": how fast can a B200 gather random, row-aligned chunks of a
// 983 MB channels-last feature stack (1 KB rows), as a function of the chunk size?  It bounds what the
// Stage A gather can reach and tells whether splitting the 256 channels into L2-resident passes
// (smaller gathers, fewer DRAM re-reads) can pay off.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_granularity.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

// one group of (bytes/16) lanes per gathered chunk; `inflight` chunks issued before the adds
template <int LANES, int INFLIGHT>
__global__ void gather_kernel(const float4 *__restrict__ buf, const uint32_t *__restrict__ idx, size_t n_idx,
                              int row_f4, float4 *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int grp = lane / LANES, lig = lane % LANES;
    constexpr int GPW = 32 / LANES;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t base = warp * GPW * INFLIGHT; base + GPW * INFLIGHT <= n_idx; base += nwarps * GPW * INFLIGHT) {
        float4 v[INFLIGHT];
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) {
            const uint32_t row = idx[base + u * GPW + grp];
            v[u] = __ldg(buf + (size_t)row * row_f4 + lig);
        }
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int LANES, int INFLIGHT>
float run(const float4 *buf, const uint32_t *idx, size_t n_idx, float4 *out, int blocks, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    gather_kernel<LANES, INFLIGHT><<<blocks, 256>>>(buf, idx, n_idx, 64, out);
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) gather_kernel<LANES, INFLIGHT><<<blocks, 256>>>(buf, idx, n_idx, 64, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv) {
    const size_t rows = 960000;             // 50 views x 120 x 160 pixels, 1 KB each = 983 MB
    const size_t n_idx = 1u << 21;          // ~2 M gathers per pass, like cfg 2's valid (voxel, view) pairs
    float4 *buf, *out; uint32_t *idx;
    cudaMalloc(&buf, rows * 1024);
    cudaMemset(buf, 0, rows * 1024);
    cudaMalloc(&out, 148 * 32 * 256 * sizeof(float4));
    std::vector<uint32_t> h(n_idx);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n_idx; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s % rows); }
    cudaMalloc(&idx, n_idx * 4);
    cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    const int blocks = 148 * 8;
    printf("chunk_bytes inflight ms GB/s(requested)\n");
#define RUN(L, I) { float ms = run<L, I>(buf, idx, n_idx, out, blocks, 20); \
        printf("%5d %2d %.4f %.0f\n", L * 16, I, ms, (double)n_idx * L * 16 / ms / 1e6); }
    // 1 KB rows need 64 lanes: two passes of 32 lanes on adjacent halves == 512 B chunks at 2x the count;
    // emulate 1 KB by letting LANES=32 read both halves (INFLIGHT doubles as the second half).
    RUN(4, 4) RUN(4, 8) RUN(8, 4) RUN(8, 8) RUN(16, 4) RUN(16, 8) RUN(32, 2) RUN(32, 4) RUN(32, 8) RUN(32, 16)
    return 0;
}
