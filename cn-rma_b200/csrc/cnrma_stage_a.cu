// Stage A: dense back-projection -- voxel->pixel projection, frustum mask, nearest gather and the
// multi-view sum / mean, fused into one pass (reference: rm.py:21-69, :220-257).
//
// Work decomposition of the gather kernel (DESIGN.md "K_A"):
//   * a group of G lanes owns one voxel and a `chunk` of G*VPL 16-byte vectors of its channels; a warp
//     carries 32/G voxels.  For C = 256 fp32: G = 32, VPL = 2, one voxel per warp, two LDG.128 per lane
//     per visible view -- every gather is one fully coalesced 1 KB row of the channels-last map.
//   * the lanes of a group project their voxel through G views at a time (lane <-> view), a ballot turns
//     the frustum tests into a bit mask, and the set bits are walked in ascending view order, so the fp32
//     sum is formed in exactly the order of the reference's `self.volume + volume` loop.
//   * up to kUnroll gathers are issued before the first add to keep several 16-byte loads in flight
//     per lane.
//   * camera matrices (pre-divided by the backbone stride) and per-view base pointers sit in shared
//     memory, matrices as structure-of-arrays so that lane <-> view reads are conflict free.
//   * blockIdx.y walks channel chunks (pass-major CTA order) when C exceeds one chunk.
#include "cnrma_internal.cuh"

namespace cnrma {

struct AggParams {
    GridDev g;
    int V, C, H, W;
    int nvox;
    int64_t stride_y, stride_x;   // elements
    float stride;                 // backbone2d_stride
    const float *proj;            // [V] 3x4, view stride proj_stride
    int64_t proj_stride;
    float *volume;
    int64_t vsv, vsc;             // volume strides (voxel, channel)
    int32_t *count;
    uint8_t *valid;
    uint32_t flags;
    int vec_store;                // volume is channels-last and 16-byte aligned
    int chunk_base;               // first channel chunk of this launch (added to blockIdx.y)
    int write_count;              // this launch owns count/valid (see run_aggregate)
    const void *views[kMaxViewsPerLaunch];
};


template <int G, int VPL, typename T>
__global__ void __launch_bounds__(kAggThreads) aggregate_views_kernel(const __grid_constant__ AggParams p) {
    using V16 = Vec16<T>;
    constexpr int E = V16::kElems;             // channels per 16-byte vector
    constexpr int kVoxPerWarp = kWarp / G;
    constexpr unsigned kGroupMask = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
    // gathers issued back to back before the first add: bounded so the staging registers stay <= 32 floats
    constexpr int kUnroll = (32 / (VPL * E)) > 0 ? (32 / (VPL * E)) : 1;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Vpad = p.V | 1;                  // odd stride: conflict-free lane<->view reads
    const T **sView = reinterpret_cast<const T **>(smem_raw);                        // [V]
    float *sP = reinterpret_cast<float *>(smem_raw + sizeof(void *) * p.V);          // [12][Vpad]

    for (int i = threadIdx.x; i < 12 * p.V; i += blockDim.x) {
        const int v = i / 12, k = i % 12;
        float val = __ldg(p.proj + (int64_t)v * p.proj_stride + k);
        if (k < 8) val = __fdiv_rn(val, p.stride);   // rows 0-1 / stride (rm.py:238-239)
        sP[k * Vpad + v] = val;
    }
    for (int i = threadIdx.x; i < p.V; i += blockDim.x) sView[i] = static_cast<const T *>(p.views[i]);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp = lane / G;
    const int lig = lane % G;
    const int vox = (blockIdx.x * (kAggThreads / kWarp) + warp) * kVoxPerWarp + grp;
    const bool vox_ok = vox < p.nvox;
    const int c0 = ((p.chunk_base + blockIdx.y) * (G * VPL) + lig) * E;   // first channel of this lane

    // voxel order of datasets/tsdf.py:24-29: flat = (x*ny + y)*nz + z
    const int vz = vox % p.g.nz;
    const int vxy = vox / p.g.nz;
    const int vy = vxy % p.g.ny;
    const int vx = vxy / p.g.ny;
    const float wx = world_coord(vx, p.g.vs, p.g.ox);
    const float wy = world_coord(vy, p.g.vs, p.g.oy);
    const float wz = world_coord(vz, p.g.vs, p.g.oz);

    float acc[VPL][E];
    int cnt = 0;
    if ((p.flags & CNRMA_AGG_ACCUMULATE) && vox_ok) {
        cnt = (p.flags & CNRMA_AGG_COUNT_F32) ? (int)reinterpret_cast<const float *>(p.count)[vox] : p.count[vox];
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int e = 0; e < E; ++e)
                acc[k][e] = p.volume[(int64_t)vox * p.vsv + (int64_t)(c0 + k * G * E + e) * p.vsc];
    } else {
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int e = 0; e < E; ++e) acc[k][e] = 0.0f;
    }

    for (int v0 = 0; v0 < p.V; v0 += G) {
        const int view = v0 + lig;
        int off = -1;
        if (view < p.V && vox_ok) {
            int px, py;
            if (project_voxel(sP + view, Vpad, wx, wy, wz, p.H, p.W, px, py))
                off = (int)(py * p.stride_y + px * p.stride_x);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, off >= 0);
        unsigned bits = (ballot >> (grp * G)) & kGroupMask;
        cnt += __popc(bits);
        while (__any_sync(0xffffffffu, bits != 0u)) {
            const T *src[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int j = bits ? (__ffs(bits) - 1) : 0;
                const int o = __shfl_sync(0xffffffffu, off, grp * G + j);
                src[u] = bits ? (sView[v0 + j] + o + c0) : nullptr;
                bits &= bits - 1u;
            }
            V16 val[kUnroll][VPL];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (src[u] != nullptr) {
#pragma unroll
                    for (int k = 0; k < VPL; ++k) val[u][k] = V16::load(src[u] + k * G * E);
                }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (src[u] != nullptr) {
#pragma unroll
                    for (int k = 0; k < VPL; ++k)
#pragma unroll
                        for (int e = 0; e < E; ++e) acc[k][e] = __fadd_rn(acc[k][e], val[u][k].v[e]);
                }
        }
    }

    if (!vox_ok) return;
    if (p.flags & CNRMA_AGG_MEAN) {
        const float n = (float)cnt;   // rm.py:251: fp32 sum / int64 count -> true division by float(count)
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int e = 0; e < E; ++e) acc[k][e] = (cnt > 0) ? __fdiv_rn(acc[k][e], n) : 0.0f;
    }
    if (p.vec_store) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            float *dst = p.volume + (int64_t)vox * p.vsv + (c0 + k * G * E);
#pragma unroll
            for (int e = 0; e < E; e += 4)
                __stcs(reinterpret_cast<float4 *>(dst + e),
                       make_float4(acc[k][e], acc[k][e + 1], acc[k][e + 2], acc[k][e + 3]));
        }
    } else {
#pragma unroll
        for (int k = 0; k < VPL; ++k)
#pragma unroll
            for (int e = 0; e < E; ++e)
                p.volume[(int64_t)vox * p.vsv + (int64_t)(c0 + k * G * E + e) * p.vsc] = acc[k][e];
    }
    if (lig == 0 && blockIdx.y == 0 && p.write_count) {
        if (p.flags & CNRMA_AGG_COUNT_F32) reinterpret_cast<float *>(p.count)[vox] = (float)cnt;
        else p.count[vox] = cnt;
        if (p.valid != nullptr) p.valid[vox] = (uint8_t)(cnt > 0);
    }
}

// ---- launch ------------------------------------------------------------------------------------

template <int G, int VPL, typename T>
static cudaError_t launch_agg(const AggParams &p, int chunks, cudaStream_t stream) {
    constexpr int kVoxPerCta = (kAggThreads / kWarp) * (kWarp / G);
    const dim3 grid((p.nvox + kVoxPerCta - 1) / kVoxPerCta, chunks);
    const int Vpad = p.V | 1;
    const size_t smem = sizeof(void *) * p.V + sizeof(float) * 12 * Vpad;
    aggregate_views_kernel<G, VPL, T><<<grid, kAggThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int G, typename T>
static cudaError_t launch_agg_vpl(const AggParams &p, int vpl, int chunks, cudaStream_t stream) {
    switch (vpl) {
        case 1: return launch_agg<G, 1, T>(p, chunks, stream);
        case 2: return launch_agg<G, 2, T>(p, chunks, stream);
        case 3: return launch_agg<G, 3, T>(p, chunks, stream);
        case 4: return launch_agg<G, 4, T>(p, chunks, stream);
    }
    return cudaErrorInvalidValue;
}

template <typename T>
static cudaError_t launch_agg_g(const AggParams &p, int g, int vpl, int chunks, cudaStream_t stream) {
    switch (g) {
        case 1: return launch_agg_vpl<1, T>(p, vpl, chunks, stream);
        case 2: return launch_agg_vpl<2, T>(p, vpl, chunks, stream);
        case 4: return launch_agg_vpl<4, T>(p, vpl, chunks, stream);
        case 8: return launch_agg_vpl<8, T>(p, vpl, chunks, stream);
        case 16: return launch_agg_vpl<16, T>(p, vpl, chunks, stream);
        case 32: return launch_agg_vpl<32, T>(p, vpl, chunks, stream);
    }
    return cudaErrorInvalidValue;
}

// Chooses lanes-per-voxel G, vectors-per-lane VPL and the number of channel chunks for `nvec` 16-byte
// vectors per feature row.  `max_chunk_vecs` (0 = no limit) caps G*VPL, which trades redundant
// projection work for a smaller per-pass L2 working set.
static void plan_agg(int nvec, int max_chunk_vecs, int &g, int &vpl, int &chunks) {
    g = 1;
    while (g < 32 && (nvec % (g * 2)) == 0) g *= 2;
    if (max_chunk_vecs > 0)
        while (g > 1 && g > max_chunk_vecs) g /= 2;
    const int q = nvec / g;
    vpl = 1;
    for (int d = 4; d >= 1; --d)
        if (q % d == 0 && (max_chunk_vecs <= 0 || g * d <= max_chunk_vecs || d == 1)) {
            vpl = d;
            break;
        }
    chunks = q / vpl;
}

static cudaError_t run_aggregate(const AggParams &p_in, int dtype, int max_chunk_vecs, cudaStream_t stream) {
    const int e = (dtype == CNRMA_BF16) ? 8 : 4;
    int g, vpl, chunks;
    plan_agg(p_in.C / e, max_chunk_vecs, g, vpl, chunks);
    AggParams p = p_in;
    p.chunk_base = 0;
    p.write_count = 1;
    if (!(p.flags & CNRMA_AGG_ACCUMULATE) || chunks == 1) {
        if (dtype == CNRMA_BF16) return launch_agg_g<__nv_bfloat16>(p, g, vpl, chunks, stream);
        return launch_agg_g<float>(p, g, vpl, chunks, stream);
    }
    // Accumulating launches read the previous count; with several channel chunks in one grid the chunk
    // that rewrites it would race with the others, so the chunks go out one launch at a time (stream
    // ordered) and only the last one stores the new count.
    for (int c = 0; c < chunks; ++c) {
        p.chunk_base = c;
        p.write_count = (c == chunks - 1);
        const cudaError_t err = (dtype == CNRMA_BF16) ? launch_agg_g<__nv_bfloat16>(p, g, vpl, 1, stream)
                                                       : launch_agg_g<float>(p, g, vpl, 1, stream);
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

// Views [v0, v0 + nv) of `f` in one launch (nv <= kMaxViewsPerLaunch).
cudaError_t run_aggregate_views(const GridDev &g, const cnrma_features &f, int v0, int nv, const float *proj,
                                int64_t proj_stride, float stride, uint32_t flags, float *volume, int64_t vsv,
                                int64_t vsc, int32_t *count, uint8_t *valid, int max_chunk_vecs, cudaStream_t stream) {
    AggParams p;
    p.g = g;
    p.V = nv;
    p.C = f.channels;
    p.H = f.height;
    p.W = f.width;
    p.nvox = g.nx * g.ny * g.nz;
    p.stride_y = f.stride_y;
    p.stride_x = f.stride_x;
    p.stride = stride;
    p.proj = proj;
    p.proj_stride = proj_stride;
    p.volume = volume;
    p.vsv = vsv;
    p.vsc = vsc;
    p.count = count;
    p.valid = valid;
    p.flags = flags;
    p.vec_store = (vsc == 1) && (vsv % 4 == 0) && (reinterpret_cast<uintptr_t>(volume) % 16 == 0);
    p.chunk_base = 0;
    p.write_count = 1;
    for (int i = 0; i < nv; ++i) p.views[i] = f.view_ptrs_host[v0 + i];
    return run_aggregate(p, f.dtype, max_chunk_vecs, stream);
}

// ---- per-view indices and masks (parity surface of rm.py:47-58) ---------------------------------

__global__ void __launch_bounds__(256) project_views_kernel(GridDev g, const float *__restrict__ proj,
                                                            int64_t proj_stride, int V, float stride, int H, int W,
                                                            int nvox, int32_t *__restrict__ px_out,
                                                            int32_t *__restrict__ py_out,
                                                            uint8_t *__restrict__ valid_out) {
    __shared__ float sP[12];
    const int view = blockIdx.y;
    if (threadIdx.x < 12) {
        float val = __ldg(proj + (int64_t)view * proj_stride + threadIdx.x);
        if (threadIdx.x < 8) val = __fdiv_rn(val, stride);
        sP[threadIdx.x] = val;
    }
    __syncthreads();
    const int vox = blockIdx.x * blockDim.x + threadIdx.x;
    if (vox >= nvox) return;
    const int vz = vox % g.nz;
    const int vxy = vox / g.nz;
    const int vy = vxy % g.ny;
    const int vx = vxy / g.ny;
    int px, py;
    const bool ok = project_voxel(sP, 1, world_coord(vx, g.vs, g.ox), world_coord(vy, g.vs, g.oy),
                                  world_coord(vz, g.vs, g.oz), H, W, px, py);
    const int64_t o = (int64_t)view * nvox + vox;
    if (px_out) px_out[o] = px;
    if (py_out) py_out[o] = py;
    if (valid_out) valid_out[o] = (uint8_t)ok;
}

cudaError_t run_project_views(const GridDev &g, const float *proj, int64_t proj_stride, int V, float stride, int H,
                              int W, int32_t *px, int32_t *py, uint8_t *valid, cudaStream_t stream) {
    const int nvox = g.nx * g.ny * g.nz;
    const dim3 grid((nvox + 255) / 256, V);
    project_views_kernel<<<grid, 256, 0, stream>>>(g, proj, proj_stride, V, stride, H, W, nvox, px, py, valid);
    return cudaGetLastError();
}

// ---- NCHW -> channels-last ------------------------------------------------------------------------
// 32 channels x 32 pixels per tile through shared memory: reads coalesced along the pixel axis of the
// source planes, writes coalesced along the channel axis of the destination rows.
template <typename T>
__global__ void __launch_bounds__(256) to_channels_last_kernel(const T *__restrict__ src, int C, int H, int W,
                                                               int64_t sc, int64_t sy, int64_t sx,
                                                               T *__restrict__ dst) {
    __shared__ T tile[32][33];
    const int hw = H * W;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int c = c0 + ty + i, pix = p0 + tx;
        if (c < C && pix < hw) tile[ty + i][tx] = src[(int64_t)c * sc + (int64_t)(pix / W) * sy + (int64_t)(pix % W) * sx];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
        const int pix = p0 + ty + i, c = c0 + tx;
        if (c < C && pix < hw) dst[(int64_t)pix * C + c] = tile[tx][ty + i];
    }
}

cudaError_t run_to_channels_last(const void *src, int dtype, int C, int H, int W, int64_t sc, int64_t sy, int64_t sx,
                                 void *dst, cudaStream_t stream) {
    const dim3 grid((H * W + 31) / 32, (C + 31) / 32);
    if (dtype == CNRMA_BF16)
        to_channels_last_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16 *>(src), C, H, W,
                                                                        sc, sy, sx, static_cast<__nv_bfloat16 *>(dst));
    else
        to_channels_last_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(src), C, H, W, sc, sy, sx,
                                                                static_cast<float *>(dst));
    return cudaGetLastError();
}

}  // namespace cnrma
