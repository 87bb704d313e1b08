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

constexpr int kMarkMaxParts = 16;

struct MarkParams {
    GridDev g;          // the box (extents + first voxel) inside the grid
    int V, H, W, words; // words = ceil(H*W / 32) per view
    float stride;
    const float *proj;
    int64_t proj_stride;
    uint32_t *bitmap;
    // the box may be cut into x-ranges ("parts", blockIdx.y), each with its own bitmap: part k covers x in
    // [part_x0[k], part_x0[k+1]) of the box and marks bitmap + k * part_stride
    int parts;
    int part_x0[kMarkMaxParts + 1];
    int64_t part_stride;   // uint32 words between the bitmaps of consecutive parts
};

constexpr int kMarkThreads = 256;
constexpr int kMarkVoxelsPerThread = 8;   // 2048 voxels per CTA: their hits of one view fall into a few dozen bitmap words
constexpr int kMarkViewsPerPass = 64;     // camera matrices staged per pass (3 KB of shared memory)
constexpr int kMarkViewsPerCta = 16;      // blockIdx.z walks chunks of views: small boxes (fewer CTAs than SMs along x)
                                          // would otherwise serialise ~2 us per view in every CTA

// Hits are collected per view in a shared-memory copy of that view's bitmap (shared-memory atomics, no contention
// across CTAs) and only its non-zero words go to global memory: ~50x fewer global atomics than one per hit, which is
// what bounded the first version of this kernel (0.25 ms for 41 M voxel-view pairs; profiles/r02_multi_gpu.md).
__global__ void __launch_bounds__(kMarkThreads) mark_rows_kernel(const MarkParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sP = reinterpret_cast<float *>(smem_raw);                                       // [kMarkViewsPerPass][12]
    uint32_t *sBits = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float) * 12 * kMarkViewsPerPass);   // [2][words]
    const int px0 = p.part_x0[blockIdx.y];
    const int nvox = (p.part_x0[blockIdx.y + 1] - px0) * p.g.ny * p.g.nz;
    const int base = blockIdx.x * (kMarkThreads * kMarkVoxelsPerThread);
    if (base >= nvox) return;   // the grid is sized for the largest part
    uint32_t *const bitmap = p.bitmap + (int64_t)blockIdx.y * p.part_stride;
    const float fW = (float)p.W - 0.5f, fH = (float)p.H - 0.5f;
    float wx[kMarkVoxelsPerThread], wy[kMarkVoxelsPerThread], wz[kMarkVoxelsPerThread];
    bool active[kMarkVoxelsPerThread];
#pragma unroll
    for (int j = 0; j < kMarkVoxelsPerThread; ++j) {
        const int vox = base + j * kMarkThreads + threadIdx.x;
        active[j] = vox < nvox;
        const int v = active[j] ? vox : 0;
        const int vz = v % p.g.nz, vxy = v / p.g.nz;
        const int vy = vxy % p.g.ny, vx = px0 + vxy / p.g.ny;
        wx[j] = world_coord(vx + p.g.x0, p.g.vs, p.g.ox);
        wy[j] = world_coord(vy + p.g.y0, p.g.vs, p.g.oy);
        wz[j] = world_coord(vz + p.g.z0, p.g.vs, p.g.oz);
    }
    for (int i = threadIdx.x; i < 2 * p.words; i += blockDim.x) sBits[i] = 0u;
    const int vbeg = blockIdx.z * kMarkViewsPerCta, vend = min(p.V, vbeg + kMarkViewsPerCta);
    for (int v0 = vbeg; v0 < vend; v0 += kMarkViewsPerPass) {
        const int nv = min(kMarkViewsPerPass, vend - v0);
        __syncthreads();
        for (int i = threadIdx.x; i < 12 * nv; i += blockDim.x) {
            const int vv = i / 12, k = i % 12;
            float val = __ldg(p.proj + (int64_t)(v0 + vv) * p.proj_stride + k);
            if (k < 8) val = __fdiv_rn(val, p.stride);   // rows 0-1 / stride (rm.py:238-239)
            sP[i] = val;
        }
        __syncthreads();
        for (int vv = 0; vv < nv; ++vv) {
            uint32_t *bits = sBits + (vv & 1) * p.words;
#pragma unroll
            const float4 a = *reinterpret_cast<const float4 *>(sP + 12 * vv);       // warp-uniform: broadcast reads
            const float4 b = *reinterpret_cast<const float4 *>(sP + 12 * vv + 4);
            const float4 c = *reinterpret_cast<const float4 *>(sP + 12 * vv + 8);
#pragma unroll
            for (int j = 0; j < kMarkVoxelsPerThread; ++j) {
                const float cx = row_dot4(a.x, a.y, a.z, a.w, wx[j], wy[j], wz[j], 1.0f);
                const float cy = row_dot4(b.x, b.y, b.z, b.w, wx[j], wy[j], wz[j], 1.0f);
                const float cz = row_dot4(c.x, c.y, c.z, c.w, wx[j], wy[j], wz[j], 1.0f);
                // cheap superset of the frustum test on the un-divided coordinates (as in the list gather kernel):
                // three quarters of the pairs fail it and skip the division / rounding / exact test
                const float slack = 1.0e-3f * cz;
                const bool maybe = active[j] && (cz > 0.0f) && (cx + 0.5f * cz >= -slack) && (fW * cz - cx >= -slack) &&
                                   (cy + 0.5f * cz >= -slack) && (fH * cz - cy >= -slack);
                if (!maybe) continue;
                float rx, ry;
                rounded_pixel(cx, cy, cz, rx, ry);
                if (in_frustum(rx, ry, cz, p.H, p.W)) {
                    const int pix = (int)ry * p.W + (int)rx;
                    atomicOr(bits + (pix >> 5), 1u << (pix & 31));
                }
            }
            __syncthreads();
            // flush this view's words while the next view collects into the other buffer; the buffer is clean again
            // before it is reused two views later (one barrier in between)
            uint32_t *gbits = bitmap + (int64_t)(v0 + vv) * p.words;
            for (int i = threadIdx.x; i < p.words; i += blockDim.x) {
                const uint32_t w = bits[i];
                if (w) {
                    bits[i] = 0u;
                    if ((gbits[i] & w) != w) atomicOr(gbits + i, w);
                }
            }
        }
    }
}

cudaError_t run_mark_rows(const GridDev &box, const float *proj, int64_t proj_stride, int V, float stride, int H, int W,
                          uint32_t *bitmap, int parts, int64_t part_stride, cudaStream_t stream) {
    MarkParams p;
    p.g = box;
    if (parts < 1) parts = 1;
    if (parts > kMarkMaxParts) parts = kMarkMaxParts;
    if (parts > box.nx) parts = box.nx;
    p.parts = parts;
    int widest = 0;
    for (int k = 0; k <= parts; ++k) {
        p.part_x0[k] = (int)((int64_t)box.nx * k / parts);   // the cuts of distributed.x_chunks
        if (k > 0 && p.part_x0[k] - p.part_x0[k - 1] > widest) widest = p.part_x0[k] - p.part_x0[k - 1];
    }
    p.part_stride = part_stride;
    p.V = V; p.H = H; p.W = W;
    p.words = (H * W + 31) / 32;
    p.stride = stride;
    p.proj = proj;
    p.proj_stride = proj_stride;
    p.bitmap = bitmap;
    const int nvox = widest * box.ny * box.nz;
    const size_t smem = sizeof(float) * 12 * kMarkViewsPerPass + sizeof(uint32_t) * 2 * (size_t)p.words;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;   // images beyond ~800 k pixels: not a feature-map size
    static thread_local int configured_dev = -1;
    static thread_local size_t configured_smem = 0;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (smem > 48 * 1024 && (configured_dev != dev || configured_smem < smem)) {
        err = cudaFuncSetAttribute(mark_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        configured_dev = dev;
        configured_smem = smem;
    }
    const int per_cta = kMarkThreads * kMarkVoxelsPerThread;
    mark_rows_kernel<<<dim3((nvox + per_cta - 1) / per_cta, parts, (V + kMarkViewsPerCta - 1) / kMarkViewsPerCta),
                       kMarkThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

// ---- pull ----------------------------------------------------------------------------------------------------

struct PullParams {
    const uint32_t *bitmap;
    uint32_t *done;   // optional: rows already present in dst (skipped), updated with the rows pulled by this launch
    int views, words, pixels, row_bytes, rows_per_stage;
    unsigned char *dst;
    int64_t dst_vs;
    unsigned int *work;   // optional: zeroed counter; warps then claim bitmap words one at a time (balances sparse bitmaps)
    int64_t first_word;   // the walk over the bitmap starts here and wraps around: ranks that pull from the same owners
                          // start at different owners, so that no owner serves every reader at once
    const void *src[kMaxViewsPerLaunch];   // per view: its map in the owner's memory (peer-mapped)
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

__global__ void __launch_bounds__(kPullWarps * 32) pull_rows_kernel(const __grid_constant__ PullParams p) {
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
    int64_t w = (int64_t)blockIdx.x * kPullWarps + warp;
    for (;;) {
        if (p.work != nullptr) {   // dynamic: one word per claim (up to 32 rows)
            unsigned int claimed = 0;
            if (lane == 0) claimed = atomicAdd(p.work, 1u);
            w = __shfl_sync(0xffffffffu, claimed, 0);
        }
        if (w >= total_words) break;
        int64_t wi = w + p.first_word;
        if (wi >= total_words) wi -= total_words;
        uint32_t bits = __ldg(p.bitmap + wi);
        if (p.done != nullptr) {   // every word is handled by exactly one warp: a plain read-modify-write
            const uint32_t have = p.done[wi];
            __syncwarp();
            if (lane == 0 && (bits & ~have)) p.done[wi] = have | bits;
            bits &= ~have;
        }
        const int view = (int)(wi / p.words);
        const int pix0 = (int)(wi % p.words) * 32;
        if (pix0 + 32 > p.pixels) bits &= (1u << (p.pixels - pix0)) - 1u;   // the last word of a view may be partial
        const unsigned char *vsrc = static_cast<const unsigned char *>(p.src[view]) + (int64_t)pix0 * p.row_bytes;
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
        w += nwarps;   // static striding when there is no work counter
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

// The same copy through the load/store units instead of the TMA: 16-byte loads of the peer-mapped rows into registers,
// eight in flight per lane, then streaming stores.  Used when the puller runs BESIDE the gather kernel, whose every row
// gather is a bulk copy: on shared SMs the puller's bulk copies queue behind thousands of the gather's (measured:
// the pulls took twice as long as alone, profiles/r02_multi_gpu.md), while the LSU path is all but idle there.
constexpr int kPullLsuWarps = 8;
constexpr int kPullLsuVecs = 8;   // 16-byte vectors in flight per lane: 4 KB per warp

__global__ void __launch_bounds__(kPullLsuWarps * 32) pull_rows_lsu_kernel(const __grid_constant__ PullParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = p.row_bytes >> 4;                       // 16-byte vectors per row
    const int64_t total_words = (int64_t)p.views * p.words;
    const int64_t nwarps = (int64_t)gridDim.x * kPullLsuWarps;
    int64_t w = (int64_t)blockIdx.x * kPullLsuWarps + warp;
    for (;;) {
        if (p.work != nullptr) {
            unsigned int claimed = 0;
            if (lane == 0) claimed = atomicAdd(p.work, 1u);
            w = __shfl_sync(0xffffffffu, claimed, 0);
        }
        if (w >= total_words) break;
        int64_t wi = w + p.first_word;
        if (wi >= total_words) wi -= total_words;
        uint32_t bits = __ldg(p.bitmap + wi);
        if (p.done != nullptr) {
            const uint32_t have = p.done[wi];
            __syncwarp();
            if (lane == 0 && (bits & ~have)) p.done[wi] = have | bits;
            bits &= ~have;
        }
        const int view = (int)(wi / p.words);
        const int pix0 = (int)(wi % p.words) * 32;
        if (pix0 + 32 > p.pixels) bits &= (1u << (p.pixels - pix0)) - 1u;
        const unsigned char *vsrc = static_cast<const unsigned char *>(p.src[view]) + (int64_t)pix0 * p.row_bytes;
        unsigned char *vdst = p.dst + (int64_t)view * p.dst_vs + (int64_t)pix0 * p.row_bytes;
        // the word's rows as one list of 16-byte vectors: vector j of the list is vector (j % nvec) of the (j / nvec)-th
        // marked row; the warp moves 32 * kPullLsuVecs of them per round
        const int total = __popc(bits) * nvec;
        for (int j0 = 0; j0 < total; j0 += 32 * kPullLsuVecs) {
            uint4 val[kPullLsuVecs];
            int64_t off[kPullLsuVecs];
#pragma unroll
            for (int u = 0; u < kPullLsuVecs; ++u) {
                const int j = j0 + u * 32 + lane;
                off[u] = -1;
                if (j < total) {
                    const int r = j / nvec;
                    const int row = (int)__fns(bits, 0, r + 1);
                    off[u] = (int64_t)row * p.row_bytes + (int64_t)(j - r * nvec) * 16;
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(val[u].x), "=r"(val[u].y), "=r"(val[u].z), "=r"(val[u].w)
                                 : "l"(vsrc + off[u]));
                }
            }
#pragma unroll
            for (int u = 0; u < kPullLsuVecs; ++u)
                if (off[u] >= 0) __stcs(reinterpret_cast<uint4 *>(vdst + off[u]), val[u]);
        }
        w += nwarps;
    }
}

int pull_default_ctas() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return sms / 2;   // 74 CTAs x 4 warps x 8 KB per stage = 2.4 MB of reads in flight (NVLink: ~0.77 TB/s x ~2.5 us)
}

cudaError_t run_pull_rows(const uint32_t *bitmap, uint32_t *done, int views, int H, int W, int row_bytes,
                          const void *const *src_views_host, void *dst, int64_t dst_vs, int ctas, unsigned int *work,
                          int first_view, int use_lsu, cudaStream_t stream) {
    PullParams p;
    p.bitmap = bitmap;
    p.done = done;
    p.work = work;
    for (int v = 0; v < views; ++v) p.src[v] = src_views_host[v];
    p.views = views;
    p.pixels = H * W;
    p.words = (H * W + 31) / 32;
    p.first_word = (int64_t)first_view * p.words;
    p.row_bytes = row_bytes;
    p.rows_per_stage = kPullStageBytes / row_bytes;
    if (p.rows_per_stage > 32) p.rows_per_stage = 32;
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
    if (use_lsu) {
        pull_rows_lsu_kernel<<<ctas, kPullLsuWarps * 32, 0, stream>>>(p);
        return cudaGetLastError();
    }
    pull_rows_kernel<<<ctas, kPullWarps * 32, kPullSmem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cnrma
