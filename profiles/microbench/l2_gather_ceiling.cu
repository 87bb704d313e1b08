// Microbenchmark behind DESIGN.md "K_A": what does a B200 deliver for Stage A's access pattern — 1 KB rows fetched
// by TMA bulk copies into per-warp shared-memory buffers (BATCH rows per mbarrier phase, STAGES phases in flight per
// warp; the real kernel is 6 x 1: a 6 KB buffer per warp, 4 CTAs of 8 warps per SM) — as a function of the size of the set the rows are drawn from?  A set that fits the L2
// gives the L2→SM ceiling of the pattern, a 983 MB set (cfg 2's maps) its DRAM ceiling with no reuse at all, and
// the sets in between the blend Stage A lives in (its L2 hit rate is 31 %).  `consume` adds the shared-memory
// reads and the adds of the real kernel; `project` adds a stand-in for its projection arithmetic per row.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_ceiling l2_gather_ceiling.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, int n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n));
}
__device__ __forceinline__ void bar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W%=;\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

constexpr int kRow = 1024, kWarps = 8;

template <int kBatch, int kStages>
__global__ void __launch_bounds__(256)
gather_kernel(const char *__restrict__ base, uint32_t set_rows, int batches, int consume, int project,
              float4 *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[kWarps * kStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *buf = smem + (size_t)warp * kStages * kBatch * kRow;
    if (lane == 0)
        for (int s = 0; s < kStages; ++s) bar_init(smem_u32(&bars[warp * kStages + s]), 1);
    __syncwarp();
    const uint32_t gwarp = blockIdx.x * kWarps + warp;
    float4 acc = make_float4(0, 0, 0, 0);
    auto issue = [&](int b) {
        const int s = b % kStages;
        const uint32_t bar = smem_u32(&bars[warp * kStages + s]);
        if (lane == 0) bar_expect(bar, kBatch * kRow);
        __syncwarp();
        if (lane < kBatch) {
            uint32_t h = mix(gwarp * 0x9e3779b9u + (uint32_t)b * 8191u + lane);
            if (project) {          // ~40 dependent flops per row, like one voxel x view projection
                float x = __uint_as_float((h & 0x007fffffu) | 0x3f800000u), y = x;
                for (int k = 0; k < 10; ++k) { y = __fmaf_rn(y, 1.0001f, x); x = __fmaf_rn(x, 0.9999f, y); y = __fdiv_rn(y, x + 2.f); }
                h ^= __float_as_uint(y) & 1u;
            }
            const uint32_t row = (uint32_t)(((uint64_t)h * set_rows) >> 32);
            bulk_g2s(smem_u32(buf + ((size_t)s * kBatch + lane) * kRow), base + (size_t)row * kRow, kRow, bar);
        }
    };
    for (int b = 0; b < kStages - 1 && b < batches; ++b) issue(b);
    for (int b = 0; b < batches; ++b) {
        if (b + kStages - 1 < batches) issue(b + kStages - 1);
        const int s = b % kStages;
        bar_wait(smem_u32(&bars[warp * kStages + s]), (b / kStages) & 1);
        if (consume) {
            const float4 *rows = reinterpret_cast<const float4 *>(buf + (size_t)s * kBatch * kRow);
#pragma unroll
            for (int r = 0; r < kBatch; ++r) {
                const float4 a = rows[r * 64 + lane], c = rows[r * 64 + 32 + lane];
                acc.x += a.x + c.x; acc.y += a.y + c.y; acc.z += a.z + c.z; acc.w += a.w + c.w;
            }
        }
        __syncwarp();
    }
    if (consume) out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int kBatch, int kStages>
void sweep(const char *buf, float4 *out, int ctas_per_sm, bool brief = false) {
    const int blocks = 148 * ctas_per_sm;
    const int smem = kWarps * kStages * kBatch * kRow;
    cudaFuncSetAttribute(gather_kernel<kBatch, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, gather_kernel<kBatch, kStages>, 256, smem);
    const int batches = (int)(2.5e6 / ((double)blocks * kWarps * kBatch) + 0.5);   // ~2.5 M rows per launch, like cfg 2
    const double bytes = (double)blocks * kWarps * batches * kBatch * kRow;
    printf("# batch %d x stages %d: %d CTAs x 256 threads, %d resident per SM, %d B of shared memory each, %.2f GB per launch\n",
           kBatch, kStages, blocks, resident, smem, bytes / 1e9);
    printf("%10s %8s %8s %10s %10s\n", "set_MB", "consume", "project", "ms", "TB/s");
    const double sets_mb[] = {16, 32, 64, 96, 128, 192, 256, 384, 512, 983.04};
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mode = 0; mode < 3; ++mode) {
        const int consume = mode >= 1, project = mode >= 2;
        for (double mb : sets_mb) {
            if (brief && (mode == 1 || (mb != 64 && mb != 384 && mb < 900))) continue;
            const uint32_t set_rows = (uint32_t)(mb * 1e6 / kRow);
            for (int w = 0; w < 2; ++w)
                gather_kernel<kBatch, kStages><<<blocks, 256, smem>>>(buf, set_rows, batches, consume, project, out);
            const int reps = 10;
            cudaEventRecord(a);
            for (int r = 0; r < reps; ++r)
                gather_kernel<kBatch, kStages><<<blocks, 256, smem>>>(buf, set_rows, batches, consume, project, out);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            ms /= reps;
            printf("%10.1f %8d %8d %10.4f %10.2f\n", mb, consume, project, ms, bytes / ms / 1e9);
        }
    }
}

int main(int argc, char **argv) {
    const size_t max_bytes = 983040000;       // 50 x 120 x 160 rows of 1 KB
    char *buf; float4 *out;
    cudaMalloc(&buf, max_bytes);
    cudaMemset(buf, 0, max_bytes);
    cudaMalloc(&out, (size_t)148 * 4 * 256 * sizeof(float4));
    if (argc > 1) {               // batch-size sensitivity: the real kernel's last batch of a voxel is short
        sweep<1, 1>(buf, out, 4, true);
        sweep<2, 1>(buf, out, 4, true);
        sweep<3, 1>(buf, out, 4, true);
        sweep<4, 1>(buf, out, 4, true);
        sweep<6, 1>(buf, out, 4, true);
        sweep<3, 2>(buf, out, 4, true);
        sweep<2, 3>(buf, out, 4, true);
        return cudaDeviceSynchronize() != cudaSuccess;
    }
    sweep<6, 1>(buf, out, 4);     // the real kernel's shape
    sweep<6, 2>(buf, out, 2);     // twice the buffer, half the CTAs
    sweep<8, 3>(buf, out, 1);     // one fat CTA per SM: 192 KB in flight per SM at all times
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
