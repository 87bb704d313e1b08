// Stage A with bilinear sampling -- an OPT-IN extra, not part of the parity path.  BASELINE.json's north_star names a
// "bilinear feature gather"; the reference itself samples nearest (rm.py:52-53 `.round()`), and parity with the
// reference wins, so nearest is the default everywhere.  This kernel offers the smoother variant with the SAME
// validity mask and view count as the nearest path (so `count` / `valid` stay bit-identical to the reference's):
//   (fx, fy) = (cx/cz, cy/cz);  x0 = floor(fx), ax = fx - x0 (same for y);  neighbours clamped to the image border
//   value = (1-ax)(1-ay) f[y0][x0] + ax(1-ay) f[y0][x0+1] + (1-ax) ay f[y0+1][x0] + ax ay f[y0+1][x0+1]
// i.e. torch.nn.functional.grid_sample(mode='bilinear', padding_mode='border', align_corners=True) at the projected
// position, summed over the visible views in view order and divided by their count.
// One warp per voxel, lanes across the channel vectors, views walked sequentially; simple rather than tuned.
#include "cnrma_internal.cuh"

namespace cnrma {

struct BilinearParams {
    GridDev g;
    int V, C, H, W, nvox;
    int64_t stride_y, stride_x;
    float stride;
    const float *proj;
    int64_t proj_stride;
    float *volume;      // [nvox, C] channels-last
    int32_t *count;
    uint8_t *valid;
    uint32_t flags;
    const void *views[kMaxViewsPerLaunch];
};

template <typename T>
__global__ void __launch_bounds__(256) aggregate_views_bilinear_kernel(const __grid_constant__ BilinearParams p) {
    using V16 = Vec16<T>;
    constexpr int E = V16::kElems;
    extern __shared__ __align__(16) float sP[];   // [V][12]
    for (int i = threadIdx.x; i < 12 * p.V; i += blockDim.x) {
        float val = __ldg(p.proj + (int64_t)(i / 12) * p.proj_stride + (i % 12));
        if (i % 12 < 8) val = __fdiv_rn(val, p.stride);
        sP[i] = val;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = p.C / E;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    for (int vox = blockIdx.x * (blockDim.x >> 5) + warp; vox < p.nvox; vox += warps_total) {
        const int vz = vox % p.g.nz, vxy = vox / p.g.nz, vy = vxy % p.g.ny, vx = vxy / p.g.ny;
        const float wx = world_coord(vx, p.g.vs, p.g.ox), wy = world_coord(vy, p.g.vs, p.g.oy), wz = world_coord(vz, p.g.vs, p.g.oz);
        int cnt = 0;
        for (int j0 = 0; j0 < nvec; j0 += 32) {       // channel vectors in groups of 32 lanes
            const int j = j0 + lane;
            float acc[E];
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = 0.0f;
            int n = 0;
            for (int vb = 0; vb < p.V; vb += 32) {     // lane <-> view projection, then walk the visible ones
                const int vl = vb + lane;
                float fx = 0.0f, fy = 0.0f;
                bool ok = false;
                if (vl < p.V) {
                    const float *P = sP + 12 * vl;
                    const float cx = row_dot4(P[0], P[1], P[2], P[3], wx, wy, wz, 1.0f);
                    const float cy = row_dot4(P[4], P[5], P[6], P[7], wx, wy, wz, 1.0f);
                    const float cz = row_dot4(P[8], P[9], P[10], P[11], wx, wy, wz, 1.0f);
                    fx = __fdiv_rn(cx, cz);
                    fy = __fdiv_rn(cy, cz);
                    ok = in_frustum(rintf(fx), rintf(fy), cz, p.H, p.W);   // the reference's mask (rm.py:58)
                }
                unsigned hit = __ballot_sync(0xffffffffu, ok);
                n += __popc(hit);
                while (hit) {                          // ascending view order, like the reference's accumulation
                    const int src = __ffs(hit) - 1;
                    hit &= hit - 1;
                    const float sx = __shfl_sync(0xffffffffu, fx, src), sy = __shfl_sync(0xffffffffu, fy, src);
                    const float x0f = floorf(sx), y0f = floorf(sy);
                    const float ax = sx - x0f, ay = sy - y0f;
                    const int x0 = min(max((int)x0f, 0), p.W - 1), x1 = min(max((int)x0f + 1, 0), p.W - 1);
                    const int y0 = min(max((int)y0f, 0), p.H - 1), y1 = min(max((int)y0f + 1, 0), p.H - 1);
                    if (j < nvec) {
                        const T *base = static_cast<const T *>(p.views[vb + src]) + j * E;
                        const V16 f00 = V16::load(base + y0 * p.stride_y + x0 * p.stride_x);
                        const V16 f10 = V16::load(base + y0 * p.stride_y + x1 * p.stride_x);
                        const V16 f01 = V16::load(base + y1 * p.stride_y + x0 * p.stride_x);
                        const V16 f11 = V16::load(base + y1 * p.stride_y + x1 * p.stride_x);
                        const float w00 = (1.0f - ax) * (1.0f - ay), w10 = ax * (1.0f - ay), w01 = (1.0f - ax) * ay, w11 = ax * ay;
#pragma unroll
                        for (int e = 0; e < E; ++e)
                            acc[e] += w00 * f00.v[e] + w10 * f10.v[e] + w01 * f01.v[e] + w11 * f11.v[e];
                    }
                }
            }
            cnt = n;
            if (j < nvec) {
                float *dst = p.volume + (int64_t)vox * p.C + j * E;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    float r = acc[e];
                    if (p.flags & CNRMA_AGG_MEAN) r = (n > 0) ? r / (float)n : 0.0f;
                    dst[e] = r;
                }
            }
        }
        if (lane == 0) {
            p.count[vox] = cnt;
            if (p.valid) p.valid[vox] = (uint8_t)(cnt > 0);
        }
    }
}

cudaError_t run_aggregate_bilinear(const GridDev &g, const cnrma_features &f, const float *proj, int64_t proj_stride,
                                   float stride, uint32_t flags, float *volume, int32_t *count, uint8_t *valid,
                                   cudaStream_t stream) {
    BilinearParams p;
    p.g = g;
    p.V = f.views; p.C = f.channels; p.H = f.height; p.W = f.width;
    p.nvox = g.nx * g.ny * g.nz;
    p.stride_y = f.stride_y; p.stride_x = f.stride_x;
    p.stride = stride;
    p.proj = proj;
    p.proj_stride = proj_stride;
    p.volume = volume; p.count = count; p.valid = valid;
    p.flags = flags;
    for (int i = 0; i < f.views; ++i) p.views[i] = f.view_ptrs_host[i];
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int needed = (p.nvox + 7) / 8;
    const int blocks = needed < sms * 8 ? needed : sms * 8;
    const size_t smem = sizeof(float) * 12 * f.views;
    if (f.dtype == CNRMA_BF16)
        aggregate_views_bilinear_kernel<__nv_bfloat16><<<blocks, 256, smem, stream>>>(p);
    else
        aggregate_views_bilinear_kernel<float><<<blocks, 256, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace cnrma
