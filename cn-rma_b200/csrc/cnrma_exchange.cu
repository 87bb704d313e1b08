// Voxel-sharded Stage A across GPUs, the exchange step (the reference itself only has scene data
// parallelism, dist_train.sh:8).  Every rank owns one box of the volume and lifts ALL views into it, so no partial
// sums cross the wire and the fp32 sums keep the reference's view order (rm.py:243); what a rank needs from the views
// its peers hold are the feature rows its own voxels project to -- 20-35 % of them for the boxes and cameras at hand.
//
//   mark_rows_kernel   thread <-> voxel of the box, views in a loop with the camera matrix warp-uniform: the projection
//                      arithmetic of the gather kernels (rm.py:47-58, bit-exact), each hit sets one bit of a
//                      (view, pixel) bitmap.
//   pull_rows_kernel   warps walk the bitmap 32 pixels at a time; every marked row is fetched from the peer-mapped
//                      source with one TMA bulk copy global -> shared (the read crosses NVLink) and pushed to the same
//                      offset of the local staging copy with one bulk copy shared -> global, two stages per warp so
//                      that the reads of one batch overlap the writes of the previous one.  No registers hold data in
//                      flight; a few CTAs keep megabytes of NVLink reads outstanding and leave the rest of the GPU to
//                      the gather kernel that runs beside it.
#include "cnrma_internal.cuh"

namespace cnrma {

// ---- mark ----------------------------------------------------------------------------------------------------

struct MarkParams {
    GridDev g;          // the box (extents + first voxel) inside the grid
    int V, H, W, words; // words = ceil(H*W / 32) per view
    float stride;
    const float *proj;
    int64_t proj_stride;
    uint32_t *bitmap;
};

constexpr int kMarkThreads = 256;
constexpr int kMarkViewsPerPass = 64;   // camera matrices staged per pass (3 KB of shared memory)

__global__ void __launch_bounds__(kMarkThreads) mark_rows_kernel(const MarkParams p) {
    __shared__ __align__(16) float sP[kMarkViewsPerPass * 12];
    const int nvox = p.g.nx * p.g.ny * p.g.nz;
    const int vox = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = vox < nvox;
    const int v = active ? vox : 0;
    const int vz = v % p.g.nz, vxy = v / p.g.nz;
    const int vy = vxy % p.g.ny, vx = vxy / p.g.ny;
    const float wx = world_coord(vx + p.g.x0, p.g.vs, p.g.ox);
    const float wy = world_coord(vy + p.g.y0, p.g.vs, p.g.oy);
    const float wz = world_coord(vz + p.g.z0, p.g.vs, p.g.oz);
    for (int v0 = 0; v0 < p.V; v0 += kMarkViewsPerPass) {
        const int nv = min(kMarkViewsPerPass, p.V - v0);
        __syncthreads();
        for (int i = threadIdx.x; i < 12 * nv; i += blockDim.x) {
            const int vv = i / 12, k = i % 12;
            float val = __ldg(p.proj + (int64_t)(v0 + vv) * p.proj_stride + k);
            if (k < 8) val = __fdiv_rn(val, p.stride);   // rows 0-1 / stride (rm.py:238-239)
            sP[i] = val;
        }
        __syncthreads();
        if (!active) continue;
        for (int vv = 0; vv < nv; ++vv) {
            int px, py;
            if (!project_voxel(sP + 12 * vv, 1, wx, wy, wz, p.H, p.W, px, py)) continue;
            const int pix = py * p.W + px;
            uint32_t *word = p.bitmap + (int64_t)(v0 + vv) * p.words + (pix >> 5);
            const uint32_t bit = 1u << (pix & 31);
            if (!(*reinterpret_cast<volatile uint32_t *>(word) & bit)) atomicOr(word, bit);   // mostly already set
        }
    }
}

cudaError_t run_mark_rows(const GridDev &box, const float *proj, int64_t proj_stride, int V, float stride, int H, int W,
                          uint32_t *bitmap, cudaStream_t stream) {
    MarkParams p;
    p.g = box;
    p.V = V; p.H = H; p.W = W;
    p.words = (H * W + 31) / 32;
    p.stride = stride;
    p.proj = proj;
    p.proj_stride = proj_stride;
    p.bitmap = bitmap;
    const int nvox = box.nx * box.ny * box.nz;
    mark_rows_kernel<<<(nvox + kMarkThreads - 1) / kMarkThreads, kMarkThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---- pull ----------------------------------------------------------------------------------------------------

struct PullParams {
    const uint32_t *bitmap;
    int views, words, pixels, row_bytes, rows_per_stage;
    const unsigned char *src;
    int64_t src_vs;
    unsigned char *dst;
    int64_t dst_vs;
};

constexpr int kPullWarps = 4;
constexpr int kPullStageBytes = 8192;                                  // per warp and stage; two stages
constexpr int kPullSmem = kPullWarps * 2 * kPullStageBytes + 64;       // + 8 mbarriers

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kPullWarps * 32) pull_rows_kernel(const PullParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *sBar = reinterpret_cast<uint64_t *>(smem_raw + kPullWarps * 2 * kPullStageBytes);   // [kPullWarps][2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kPullWarps * 2) mbar_init(smem_u32(&sBar[threadIdx.x]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t stage0 = smem_u32(smem_raw + (size_t)warp * 2 * kPullStageBytes);
    const uint32_t bar0 = smem_u32(&sBar[warp * 2]);
    uint32_t parity[2] = {0u, 0u};
    // the batch whose rows are in flight into the other stage: this lane's destination (nullptr: no row)
    unsigned char *prev_dst = nullptr;
    bool pending = false;
    int s = 0;

    const int64_t total_words = (int64_t)p.views * p.words;
    const int64_t nwarps = (int64_t)gridDim.x * kPullWarps;
    for (int64_t w = (int64_t)blockIdx.x * kPullWarps + warp; w < total_words; w += nwarps) {
        uint32_t bits = __ldg(p.bitmap + w);
        const int view = (int)(w / p.words);
        const int pix0 = (int)(w % p.words) * 32;
        if (pix0 + 32 > p.pixels) bits &= (1u << (p.pixels - pix0)) - 1u;   // the last word of a view may be partial
        const unsigned char *vsrc = p.src + (int64_t)view * p.src_vs + (int64_t)pix0 * p.row_bytes;
        unsigned char *vdst = p.dst + (int64_t)view * p.dst_vs + (int64_t)pix0 * p.row_bytes;
        while (bits) {
            const int n = min(__popc(bits), p.rows_per_stage);
            const int my = (lane < n) ? (int)__fns(bits, 0, lane + 1) : -1;   // position of this lane's set bit
            // slot `lane` of stage s was last read by this lane's own store of the batch before the previous one, which
            // sits in the group committed one batch ago (the shared-memory read of a store is short and local; the
            // NVLink read of the previous batch stays in flight across this wait)
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            const uint32_t bar = bar0 + 8u * s;
            if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(n * p.row_bytes));
            __syncwarp();
            const uint32_t slot = stage0 + (uint32_t)(s * kPullStageBytes + lane * p.row_bytes);
            if (my >= 0) bulk_g2s(slot, vsrc + (int64_t)my * p.row_bytes, (uint32_t)p.row_bytes, bar);
            if (pending) {   // the previous batch: wait for its bytes, push them to the local copy
                const int q = s ^ 1;
                mbar_wait(bar0 + 8u * q, parity[q]);
                parity[q] ^= 1u;
                if (prev_dst != nullptr)
                    bulk_s2g(prev_dst, stage0 + (uint32_t)(q * kPullStageBytes + lane * p.row_bytes), (uint32_t)p.row_bytes);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // one group per batch and lane, possibly empty
            prev_dst = (my >= 0) ? vdst + (int64_t)my * p.row_bytes : nullptr;
            pending = true;
            s ^= 1;
            // drop the bits taken by this batch (the n lowest set bits)
            const int last = __shfl_sync(0xffffffffu, my, n - 1);
            bits = (last >= 31) ? 0u : (bits & ~((2u << last) - 1u));
        }
    }
    if (pending) {
        const int q = s ^ 1;
        mbar_wait(bar0 + 8u * q, parity[q]);
        if (prev_dst != nullptr)
            bulk_s2g(prev_dst, stage0 + (uint32_t)(q * kPullStageBytes + lane * p.row_bytes), (uint32_t)p.row_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the stores are complete before the CTA retires
}

int pull_default_ctas() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms / 2;   // 74 CTAs x 4 warps x 8 KB per stage = 2.4 MB of reads in flight (NVLink: ~0.77 TB/s x ~2.5 us)
}

cudaError_t run_pull_rows(const uint32_t *bitmap, int views, int H, int W, int row_bytes, const void *src, int64_t src_vs,
                          void *dst, int64_t dst_vs, int ctas, cudaStream_t stream) {
    PullParams p;
    p.bitmap = bitmap;
    p.views = views;
    p.pixels = H * W;
    p.words = (H * W + 31) / 32;
    p.row_bytes = row_bytes;
    p.rows_per_stage = kPullStageBytes / row_bytes;
    if (p.rows_per_stage > 32) p.rows_per_stage = 32;
    p.src = static_cast<const unsigned char *>(src);
    p.src_vs = src_vs;
    p.dst = static_cast<unsigned char *>(dst);
    p.dst_vs = dst_vs;
    static thread_local int configured_dev = -1;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (configured_dev != dev) {
        err = cudaFuncSetAttribute(pull_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPullSmem);
        if (err != cudaSuccess) return err;
        configured_dev = dev;
    }
    if (ctas <= 0) ctas = pull_default_ctas();
    const int64_t total_words = (int64_t)views * p.words;
    const int64_t needed = (total_words + kPullWarps - 1) / kPullWarps;
    if (needed < ctas) ctas = (int)needed;
    if (ctas < 1) return cudaSuccess;
    pull_rows_kernel<<<ctas, kPullWarps * 32, kPullSmem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cnrma
