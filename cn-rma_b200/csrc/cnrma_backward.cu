// Backward of the two feature gathers: what makes the drop-in usable inside the
// reference's forward_train (rm.py:409-451), where the gradient flows through both lifts into the 2D features.
//
//   Stage A   volume[vox,:] = (1/count) * sum_{views that see vox} feat[view, py, px, :]          (rm.py:61-64, :243, :251)
//             d feat[view, py, px, :] += d volume[vox,:] / count                                  -- a scatter-add
//   Stage B   rows[m, 3 + :] = feat[view, v, u, :] * w_m / mean(w)   (weights carry no gradient: rm.py:705 no_grad)
//             d feat[view, v, u, :] = sum_{rows of that ray} d rows[m, 3 + :] * w_m / mean(w)     -- a per-ray sum
//
// Stage A: same decomposition as the forward kernel (persistent warps, lanes <-> views, bit-exact projection); the
// voxel's gradient row is written to shared memory once and added into every visible view's pixel row with ONE
// bulk asynchronous reduction per view (TMA, cp.reduce.async.bulk.global.shared::cta.add.f32) -- fp32 atomics
// performed by the memory system, like the index_put_(accumulate=True) autograd runs for the reference.
// Stage B: one CTA per 256 rays like the fill kernel; every ray owns its pixel, so there are no atomics at all.
#include <cstdlib>

#include "cnrma_internal.cuh"

namespace cnrma {

struct AggBwdParams {
    GridDev g;
    int V, C, H, W, nvox;
    int64_t stride_y, stride_x;   // of the gradient feature maps, elements
    float stride;
    const float *proj;
    int64_t proj_stride;
    const float *grad_volume;
    int64_t vsv, vsc;
    const int32_t *count;
    uint32_t flags;               // CNRMA_AGG_MEAN: the forward divided by the count
    int chunk_floats;             // floats of a row handled per pass (<= 256)
    SweepOrder sweep;             // same traversal order as the forward kernel
    float *views[kMaxViewsPerLaunch];
};

__device__ __forceinline__ void bulk_reduce_add_f32(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src),
                 "r"(bytes)
                 : "memory");
}

__global__ void __launch_bounds__(kAggThreads, 4) aggregate_views_backward_kernel(const __grid_constant__ AggBwdParams p) {
    constexpr int kWarps = kAggThreads / kWarp;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Vpad = p.V | 1;
    float **sView = reinterpret_cast<float **>(smem_raw);                                   // [V]
    float *sP = reinterpret_cast<float *>(smem_raw + sizeof(void *) * p.V);                 // [12][Vpad]
    const size_t head = (sizeof(void *) * p.V + sizeof(float) * 12 * Vpad + 127) & ~(size_t)127;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *rowbuf = reinterpret_cast<float *>(smem_raw + head) + (size_t)warp * p.chunk_floats;
    const uint32_t row_addr = smem_u32(rowbuf);

    for (int i = threadIdx.x; i < 12 * p.V; i += blockDim.x) {
        const int v = i / 12, k = i % 12;
        float val = __ldg(p.proj + (int64_t)v * p.proj_stride + k);
        if (k < 8) val = __fdiv_rn(val, p.stride);
        sP[k * Vpad + v] = val;
    }
    for (int i = threadIdx.x; i < p.V; i += blockDim.x) sView[i] = p.views[i];
    __syncthreads();

    const int warps_total = gridDim.x * kWarps;
    const int chunks = (p.C + p.chunk_floats - 1) / p.chunk_floats;
    for (int it = blockIdx.x * kWarps + warp; it < p.nvox; it += warps_total) {
        int vx, vy, vz;
        sweep_voxel(p.sweep, it, vx, vy, vz);
        const int vox = (vx * p.g.ny + vy) * p.g.nz + vz;
        const int cnt = __ldg(p.count + vox);
        if (cnt == 0) continue;   // no view sees the voxel: its gradient goes nowhere
        const float wx = world_coord(vx, p.g.vs, p.g.ox);
        const float wy = world_coord(vy, p.g.vs, p.g.oy);
        const float wz = world_coord(vz, p.g.vs, p.g.oz);
        const float n = (float)cnt, y = __frcp_rn(n);
        for (int ch = 0; ch < chunks; ++ch) {
            const int c0 = ch * p.chunk_floats;
            const int nf = min(p.chunk_floats, p.C - c0);
            // every lane waits for the bulk reductions it issued from this buffer, then the warp refills it
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            for (int c = lane; c < nf; c += 32) {
                float gval = __ldg(p.grad_volume + (int64_t)vox * p.vsv + (int64_t)(c0 + c) * p.vsc);
                if (p.flags & CNRMA_AGG_MEAN) gval = div_by_count(gval, n, y);   // d(sum/count) = d / count
                rowbuf[c] = gval;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            for (int v0 = 0; v0 < p.V; v0 += 32) {
                const int view = v0 + lane;
                if (view < p.V) {
                    int px, py;
                    if (project_voxel(sP + view, Vpad, wx, wy, wz, p.H, p.W, px, py)) {
                        float *dst = sView[view] + py * p.stride_y + px * p.stride_x + c0;
                        bulk_reduce_add_f32(dst, row_addr, (uint32_t)(nf * sizeof(float)));
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- Stage A backward, short rows --------------------------------------------------------------------------
// Rows below 512 bytes (the reference's 32-channel maps: 128 bytes): one bulk reduction per (voxel, view) is mostly
// overhead (reference grid: 3.6 ms).  Organised like the forward list kernel (cnrma_stage_a_list.cu): phase 1
// lane <-> voxel projects a batch of 32 voxels through the views and appends the row offsets of the views that see each
// voxel to per-voxel lists in shared memory; phase 2 lane group <-> voxel loads the voxel's gradient (16 bytes per
// lane, divided by the count once) and adds it into every listed pixel row with one vector reduction per lane
// (red.global.add.v4.f32).  Needs equally spaced gradient maps (one [V,...] tensor) and a power-of-two number of
// 16-byte vectors per row; everything else takes the bulk-reduction kernel above.
struct AggBwdListParams {
    GridDev g;
    int V, H, W, nvox;
    float stride;
    const float *proj;
    int64_t proj_stride;
    const float *grad_volume;
    int64_t vsv, vsc;
    const int32_t *count;
    uint32_t flags;
    int nb, lcap;
    SweepOrder sweep;
    unsigned char *view0;                       // gradient maps: view v at view0 + v * view_stride16 * 16
    uint32_t view_stride16, stride_y16, stride_x16;
};

__device__ __forceinline__ void red_add_v4(float *dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int kBwdListThreads = 256;

template <int G>
__global__ void __launch_bounds__(kBwdListThreads) aggregate_views_backward_list_kernel(const AggBwdListParams p) {
    constexpr int kWarps = kBwdListThreads / kWarp;
    constexpr int kVPW = kWarp / G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sP = reinterpret_cast<float *>(smem_raw);                                  // [V][12]
    uint32_t *sList = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float) * 12 * p.V);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *my_lists = sList + (size_t)warp * p.nb * p.lcap;
    for (int i = threadIdx.x; i < 12 * p.V; i += blockDim.x) {
        const int v = i / 12, k = i % 12;
        float val = __ldg(p.proj + (int64_t)v * p.proj_stride + k);
        if (k < 8) val = __fdiv_rn(val, p.stride);
        sP[i] = val;
    }
    __syncthreads();
    const int grp = lane / G, lig = lane % G;
    const int units = (p.nvox + p.nb - 1) / p.nb;
    const int warps_total = gridDim.x * kWarps;
    const float fW = (float)p.W - 0.5f, fH = (float)p.H - 0.5f;
    for (int u = blockIdx.x * kWarps + warp; u < units; u += warps_total) {
        const int it = u * p.nb + lane;
        const bool active = lane < p.nb && it < p.nvox;
        int vx, vy, vz;
        sweep_voxel(p.sweep, active ? it : 0, vx, vy, vz);
        const int vox = (vx * p.g.ny + vy) * p.g.nz + vz;
        const int total = active ? __ldg(p.count + vox) : 0;      // the forward's view count (0: nothing to do)
        const float wx = world_coord(vx, p.g.vs, p.g.ox);
        const float wy = world_coord(vy, p.g.vs, p.g.oy);
        const float wz = world_coord(vz, p.g.vs, p.g.oz);
        uint32_t *lst = my_lists + lane * p.lcap;
        int cnt = 0;
        if (__any_sync(0xffffffffu, total > 0)) {
            for (int v = 0; v < p.V; ++v) {
                const float4 a = *reinterpret_cast<const float4 *>(sP + 12 * v);
                const float4 b = *reinterpret_cast<const float4 *>(sP + 12 * v + 4);
                const float4 c = *reinterpret_cast<const float4 *>(sP + 12 * v + 8);
                const float cx = row_dot4(a.x, a.y, a.z, a.w, wx, wy, wz, 1.0f);
                const float cy = row_dot4(b.x, b.y, b.z, b.w, wx, wy, wz, 1.0f);
                const float cz = row_dot4(c.x, c.y, c.z, c.w, wx, wy, wz, 1.0f);
                const float slack = 1.0e-3f * cz;   // conservative pre-test, as in the forward list kernel
                const bool maybe = (total > 0) && (cz > 0.0f) && (cx + 0.5f * cz >= -slack) && (fW * cz - cx >= -slack) &&
                                   (cy + 0.5f * cz >= -slack) && (fH * cz - cy >= -slack);
                if (!__any_sync(0xffffffffu, maybe)) continue;
                float rx, ry;
                rounded_pixel(cx, cy, cz, rx, ry);
                if (total > 0 && in_frustum(rx, ry, cz, p.H, p.W))
                    lst[cnt++] = (uint32_t)v * p.view_stride16 + (uint32_t)(int)ry * p.stride_y16 + (uint32_t)(int)rx * p.stride_x16;
            }
        }
        __syncwarp();
        for (int b0 = 0; b0 < p.nb; b0 += kVPW) {
            const int j = b0 + grp;
            const int jcnt = __shfl_sync(0xffffffffu, cnt, j & 31);
            const int jvox = __shfl_sync(0xffffffffu, vox, j & 31);
            const int jtot = __shfl_sync(0xffffffffu, total, j & 31);
            if (j >= p.nb || jcnt == 0) continue;
            const float *src = p.grad_volume + (int64_t)jvox * p.vsv + (int64_t)(lig * 4) * p.vsc;
            float g0 = __ldg(src), g1 = __ldg(src + p.vsc), g2 = __ldg(src + 2 * p.vsc), g3 = __ldg(src + 3 * p.vsc);
            if (p.flags & CNRMA_AGG_MEAN) {   // d(sum / count) = d / count
                const float n = (float)jtot, y = __frcp_rn(n);
                g0 = div_by_count(g0, n, y);
                g1 = div_by_count(g1, n, y);
                g2 = div_by_count(g2, n, y);
                g3 = div_by_count(g3, n, y);
            }
            const uint32_t *jl = my_lists + j * p.lcap;
            for (int k = 0; k < jcnt; ++k)
                red_add_v4(reinterpret_cast<float *>(p.view0 + (int64_t)jl[k] * 16 + lig * 16), g0, g1, g2, g3);
        }
        __syncwarp();
    }
}

// Returns cudaErrorNotSupported when the shape does not fit the list kernel (the caller then uses the bulk kernel).
static cudaError_t run_aggregate_views_backward_list(const GridDev &g, const cnrma_features &gf, const float *proj,
                                                     int64_t proj_stride, float stride, uint32_t flags,
                                                     const float *grad_volume, int64_t vsv, int64_t vsc,
                                                     const int32_t *count, cudaStream_t stream) {
    const int nvec = gf.channels / 4;
    const int nv = gf.views;
    if (gf.channels % 4 != 0 || nvec > 32 || (nvec & (nvec - 1)) != 0 || nv < 1 || nv > kListViewsMax) return cudaErrorNotSupported;
    if ((gf.stride_y * 4) % 16 != 0 || (gf.stride_x * 4) % 16 != 0) return cudaErrorNotSupported;
    const intptr_t base = reinterpret_cast<intptr_t>(gf.view_ptrs_host[0]);
    const intptr_t step = nv > 1 ? reinterpret_cast<intptr_t>(gf.view_ptrs_host[1]) - base : 0;
    if (step < 0 || step % 16 != 0) return cudaErrorNotSupported;
    for (int i = 2; i < nv; ++i)
        if (reinterpret_cast<intptr_t>(gf.view_ptrs_host[i]) - base != (intptr_t)i * step) return cudaErrorNotSupported;
    const int64_t span = (int64_t)(nv - 1) * step + (int64_t)gf.height * gf.stride_y * 4 + (int64_t)gf.width * gf.stride_x * 4;
    if (span / 16 >= ((int64_t)1 << 32)) return cudaErrorNotSupported;
    AggBwdListParams p;
    p.g = g;
    p.V = nv; p.H = gf.height; p.W = gf.width;
    p.nvox = g.nx * g.ny * g.nz;
    p.stride = stride;
    p.proj = proj;
    p.proj_stride = proj_stride;
    p.grad_volume = grad_volume;
    p.vsv = vsv; p.vsc = vsc;
    p.count = count;
    p.flags = flags;
    p.lcap = nv | 1;
    int nb = 32;
    while (nb > 1 && (size_t)nb * p.lcap * 4 > 8192) nb >>= 1;
    if (nb < 32 / nvec) nb = 32 / nvec;
    p.nb = nb;
    p.sweep = make_sweep(g.nx, g.ny, g.nz, sweep_thickness(g.ny, g.nz, nv, gf.channels * 4));
    p.view0 = static_cast<unsigned char *>(const_cast<void *>(gf.view_ptrs_host[0]));
    p.view_stride16 = (uint32_t)(step / 16);
    p.stride_y16 = (uint32_t)(gf.stride_y * 4 / 16);
    p.stride_x16 = (uint32_t)(gf.stride_x * 4 / 16);
    const size_t smem = sizeof(float) * 12 * nv + sizeof(uint32_t) * (size_t)(kBwdListThreads / kWarp) * p.nb * p.lcap;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int units = (p.nvox + p.nb - 1) / p.nb;
    const int needed = (units + (kBwdListThreads / kWarp) - 1) / (kBwdListThreads / kWarp);
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int per_sm = 0;
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBwdListThreads, smem);
        if (err != cudaSuccess) return err;
        const int ctas = sms * (per_sm > 0 ? per_sm : 1);
        kernel<<<needed < ctas ? needed : ctas, kBwdListThreads, smem, stream>>>(p);
        return cudaGetLastError();
    };
    switch (nvec) {
        case 1: return go(aggregate_views_backward_list_kernel<1>);
        case 2: return go(aggregate_views_backward_list_kernel<2>);
        case 4: return go(aggregate_views_backward_list_kernel<4>);
        case 8: return go(aggregate_views_backward_list_kernel<8>);
        case 16: return go(aggregate_views_backward_list_kernel<16>);
        default: return go(aggregate_views_backward_list_kernel<32>);
    }
}

cudaError_t run_aggregate_views_backward(const GridDev &g, const cnrma_features &gf, const float *proj,
                                         int64_t proj_stride, float stride, uint32_t flags, const float *grad_volume,
                                         int64_t vsv, int64_t vsc, const int32_t *count, cudaStream_t stream) {
    // short rows: list kernel with vector reductions (CNRMA_AGG_BWD_KERNEL=bulk|list overrides)
    bool use_list = gf.channels * 4 < 512;
    if (tuning().agg_bwd_kernel >= 0) use_list = tuning().agg_bwd_kernel == 1;   // CNRMA_AGG_BWD_KERNEL
    if (use_list) {
        const cudaError_t e = run_aggregate_views_backward_list(g, gf, proj, proj_stride, stride, flags, grad_volume, vsv, vsc,
                                                                count, stream);
        if (e != cudaErrorNotSupported) return e;
    }
    AggBwdParams p;
    p.g = g;
    p.C = gf.channels; p.H = gf.height; p.W = gf.width;
    p.nvox = g.nx * g.ny * g.nz;
    p.stride_y = gf.stride_y; p.stride_x = gf.stride_x;
    p.stride = stride;
    p.proj_stride = proj_stride;
    p.grad_volume = grad_volume;
    p.vsv = vsv; p.vsc = vsc;
    p.count = count;
    p.flags = flags;
    p.chunk_floats = gf.channels < 256 ? gf.channels : 256;
    p.sweep = make_sweep(g.nx, g.ny, g.nz, sweep_thickness(g.ny, g.nz, gf.views, gf.channels * 4));
    static thread_local int ctas = 0, ctas_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int v0 = 0; v0 < gf.views; v0 += kMaxViewsPerLaunch) {
        const int nv = (gf.views - v0 < kMaxViewsPerLaunch) ? (gf.views - v0) : kMaxViewsPerLaunch;
        p.V = nv;
        p.proj = proj + (int64_t)v0 * proj_stride;
        for (int i = 0; i < nv; ++i) p.views[i] = static_cast<float *>(const_cast<void *>(gf.view_ptrs_host[v0 + i]));
        const int Vpad = nv | 1;
        const size_t head = (sizeof(void *) * nv + sizeof(float) * 12 * Vpad + 127) & ~(size_t)127;
        const size_t smem = head + (size_t)(kAggThreads / kWarp) * p.chunk_floats * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(aggregate_views_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ctas_dev != dev) {
            int sms = 0, per_sm = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, aggregate_views_backward_kernel, kAggThreads, smem);
            ctas = sms * (per_sm > 0 ? per_sm : 1);
            ctas_dev = dev;
        }
        const int needed = (p.nvox + (kAggThreads / kWarp) - 1) / (kAggThreads / kWarp);
        aggregate_views_backward_kernel<<<needed < ctas ? needed : ctas, kAggThreads, smem, stream>>>(p);
        const cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

// ---- Stage B ------------------------------------------------------------------------------------------------
struct FillBwdParams {
    int V, C, H, W;
    int64_t stride_y, stride_x;   // of the gradient feature maps
    int normalize;
    int64_t rays;
    const int32_t *counts;
    const int64_t *blk_off;
    const float *rec_w;
    const float *mean;
    const float *grad_rows;
    int64_t row_stride;
    int view_base;
    float *views[kMaxViewsPerLaunch];
};

__global__ void __launch_bounds__(kRayThreads) fill_rows_backward_kernel(const __grid_constant__ FillBwdParams p) {
    __shared__ int s_warp_rows[kRayThreads / kWarp];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hw = p.H * p.W;
    const int64_t ray0 = (int64_t)p.view_base * hw + (int64_t)blockIdx.x * kRayThreads;
    const int64_t my_ray = ray0 + threadIdx.x;
    const int64_t ray_end = (int64_t)(p.view_base + p.V) * hw;
    const bool mine = my_ray < ray_end && my_ray < p.rays;
    const int my_cnt = mine ? p.counts[my_ray] : 0;
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nn = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nn;
    }
    if (lane == 31) s_warp_rows[warp] = incl;
    __syncthreads();
    int64_t base = p.blk_off[ray0 / kRayThreads];
    for (int i = 0; i < warp; ++i) base += s_warp_rows[i];
    const int64_t my_off = base + (incl - my_cnt);
    const float mean = p.normalize ? __ldg(p.mean) : 1.0f;
    const int col0 = p.normalize ? 3 : 4;

    for (int r = 0; r < 32; ++r) {
        const int64_t ray = ray0 + warp * 32 + r;
        if (!(ray < ray_end && ray < p.rays)) break;
        const int cnt = __shfl_sync(0xffffffffu, my_cnt, r);
        const int64_t off = __shfl_sync(0xffffffffu, my_off, r);
        const int view = (int)(ray / hw);
        const int pix = (int)(ray % hw);
        float *dst = p.views[view - p.view_base] + (int64_t)(pix / p.W) * p.stride_y + (int64_t)(pix % p.W) * p.stride_x;
        for (int cbase = 0; cbase < p.C; cbase += 32 * 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
            for (int k0 = 0; k0 < cnt; k0 += 32) {
                const int nk = min(32, cnt - k0);
                float wk = 0.0f;
                if (lane < nk) {
                    const float w = __ldg(p.rec_w + (int64_t)(k0 + lane) * p.rays + ray);
                    wk = p.normalize ? __fdiv_rn(w, mean) : 1.0f;
                }
                for (int k = 0; k < nk; ++k) {
                    const float wn = __shfl_sync(0xffffffffu, wk, k);
                    const float *src = p.grad_rows + (off + k0 + k) * p.row_stride + col0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = cbase + j * 32 + lane;
                        if (c < p.C) acc[j] = __fmaf_rn(__ldg(src + c), wn, acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = cbase + j * 32 + lane;
                if (c < p.C) dst[c] = acc[j];   // every pixel row is written exactly once (zeros where nothing was kept)
            }
        }
    }
}

cudaError_t run_fill_backward(const cnrma_features &gf, const void *workspace, const RmaWorkspace &ws, int normalize,
                              const float *mean, const float *grad_rows, int64_t row_stride, cudaStream_t stream) {
    const unsigned char *base = static_cast<const unsigned char *>(workspace);
    FillBwdParams p;
    p.C = gf.channels; p.H = gf.height; p.W = gf.width;
    p.stride_y = gf.stride_y; p.stride_x = gf.stride_x;
    p.normalize = normalize;
    p.rays = ws.rays;
    p.counts = reinterpret_cast<const int32_t *>(base + ws.off_counts);
    p.blk_off = reinterpret_cast<const int64_t *>(base + ws.off_blk_off);
    p.rec_w = reinterpret_cast<const float *>(base + ws.off_rec_w);
    p.mean = mean;
    p.grad_rows = grad_rows;
    p.row_stride = row_stride;
    const int hw = gf.height * gf.width;
    int group = gf.views;
    if (gf.views > kMaxViewsPerLaunch) {
        group = kMaxViewsPerLaunch;
        while (group > 0 && ((int64_t)group * hw) % kRayThreads != 0) --group;
        if (group == 0) return cudaErrorInvalidValue;
    }
    for (int v0 = 0; v0 < gf.views; v0 += group) {
        const int nv = (gf.views - v0 < group) ? (gf.views - v0) : group;
        p.view_base = v0;
        p.V = nv;
        for (int i = 0; i < nv; ++i) p.views[i] = static_cast<float *>(const_cast<void *>(gf.view_ptrs_host[v0 + i]));
        const int64_t rays = (int64_t)nv * hw;
        fill_rows_backward_kernel<<<(unsigned)((rays + kRayThreads - 1) / kRayThreads), kRayThreads, 0, stream>>>(p);
        const cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

}  // namespace cnrma
