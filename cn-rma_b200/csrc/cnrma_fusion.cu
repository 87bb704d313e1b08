// GT TSDF fusion: TSDFFusion.integrate of data_prepare/scannet/tsdf.py:402-451 ("fz.py")
// -- the offline step that turns posed depth maps into the ground-truth TSDF.  It projects every voxel with the very
// idiom of backproject (fz.py:413-420 == rm.py:51-58), so the Stage A front end is reused as is.
//
// The reference runs ~30 full-volume torch kernels per frame.  Here one thread owns one voxel, keeps its state
// (tsdf, weight, colour sums, label) in registers and walks a batch of frames in order -- the per-voxel update
// sequence is exactly the reference's, so the volumes are bit-identical -- reading one depth sample per visible
// frame.  Camera matrices are warp-uniform (shared memory broadcast).
#include "cnrma_internal.cuh"

namespace cnrma {

constexpr int kFusionFramesPerLaunch = 128;

struct FusionParams {
    GridDev g;
    int frames, H, W, nvox;
    float trunc_margin;
    const float *proj;
    int64_t proj_stride;
    float *tsdf, *weight, *color;
    int64_t *label;
    const float *depth[kFusionFramesPerLaunch];
    const float *color_img[kFusionFramesPerLaunch];
    const int64_t *label_img[kFusionFramesPerLaunch];
};

__global__ void __launch_bounds__(256) tsdf_integrate_kernel(const __grid_constant__ FusionParams p) {
    extern __shared__ __align__(16) float sP[];   // [frames][12]
    for (int i = threadIdx.x; i < 12 * p.frames; i += blockDim.x)
        sP[i] = __ldg(p.proj + (int64_t)(i / 12) * p.proj_stride + (i % 12));
    __syncthreads();
    const int vox = blockIdx.x * blockDim.x + threadIdx.x;
    if (vox >= p.nvox) return;
    const int vz = vox % p.g.nz, vxy = vox / p.g.nz, vy = vxy % p.g.ny, vx = vxy / p.g.ny;
    const float wx = world_coord(vx, p.g.vs, p.g.ox);   // coords.float() * voxel_size + origin (fz.py:376)
    const float wy = world_coord(vy, p.g.vs, p.g.oy);
    const float wz = world_coord(vz, p.g.vs, p.g.oz);
    float tsdf = p.tsdf[vox], weight = p.weight[vox];
    float col[3] = {0.0f, 0.0f, 0.0f};
    const bool has_color = p.color != nullptr;
    if (has_color) {
#pragma unroll
        for (int c = 0; c < 3; ++c) col[c] = p.color[(int64_t)c * p.nvox + vox];
    }
    int64_t label = p.label ? p.label[vox] : -1;
    const int plane = p.H * p.W;
    for (int f = 0; f < p.frames; ++f) {
        const float4 a = *reinterpret_cast<const float4 *>(sP + 12 * f);
        const float4 b = *reinterpret_cast<const float4 *>(sP + 12 * f + 4);
        const float4 c = *reinterpret_cast<const float4 *>(sP + 12 * f + 8);
        const float cx = row_dot4(a.x, a.y, a.z, a.w, wx, wy, wz, 1.0f);
        const float cy = row_dot4(b.x, b.y, b.z, b.w, wx, wy, wz, 1.0f);
        const float cz = row_dot4(c.x, c.y, c.z, c.w, wx, wy, wz, 1.0f);
        float rx, ry;
        rounded_pixel(cx, cy, cz, rx, ry);
        if (!in_frustum(rx, ry, cz, p.H, p.W)) continue;                       // fz.py:419-420
        const int pix = (int)ry * p.W + (int)rx;
        const float d = __ldg(p.depth[f] + pix);
        if (!(d > 0.0f)) continue;                                             // fz.py:422-423
        float dist = __fdiv_rn(__fsub_rn(cz, d), p.trunc_margin);              // fz.py:426-427
        dist = (dist < -1.0f) ? -1.0f : dist;                                  // clamp(min=-1)
        if (!(dist < 1.0f)) continue;                                          // fz.py:430-433
        const bool first = (weight == 0.0f);                                   // fz.py:436
        if (first) tsdf = dist;
        if (dist > -1.0f) {                                                    // near surface, fz.py:440-445
            if (!first) tsdf = __fadd_rn(tsdf, dist);
            weight = __fadd_rn(weight, 1.0f);
            if (has_color && p.color_img[f] != nullptr) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) col[ch] = __fadd_rn(col[ch], __ldg(p.color_img[f] + ch * plane + pix));
            }
            if (p.label && p.label_img[f] != nullptr) label = __ldg(p.label_img[f] + pix);   // newest label wins
        }
    }
    p.tsdf[vox] = tsdf;
    p.weight[vox] = weight;
    if (has_color) {
#pragma unroll
        for (int c = 0; c < 3; ++c) p.color[(int64_t)c * p.nvox + vox] = col[c];
    }
    if (p.label) p.label[vox] = label;
}

cudaError_t run_tsdf_integrate(const GridDev &g, const float *proj, int64_t proj_stride, int frames,
                               const float *const *depth_host, const float *const *color_host,
                               const int64_t *const *label_host, int H, int W, float trunc_margin, float *tsdf,
                               float *weight, float *color, int64_t *label, cudaStream_t stream) {
    FusionParams p;
    p.g = g;
    p.H = H; p.W = W;
    p.nvox = g.nx * g.ny * g.nz;
    p.trunc_margin = trunc_margin;
    p.proj_stride = proj_stride;
    p.tsdf = tsdf; p.weight = weight; p.color = color; p.label = label;
    for (int f0 = 0; f0 < frames; f0 += kFusionFramesPerLaunch) {
        const int nf = (frames - f0 < kFusionFramesPerLaunch) ? (frames - f0) : kFusionFramesPerLaunch;
        p.frames = nf;
        p.proj = proj + (int64_t)f0 * proj_stride;
        for (int i = 0; i < nf; ++i) {
            p.depth[i] = depth_host[f0 + i];
            p.color_img[i] = color_host ? color_host[f0 + i] : nullptr;
            p.label_img[i] = label_host ? label_host[f0 + i] : nullptr;
        }
        tsdf_integrate_kernel<<<(p.nvox + 255) / 256, 256, sizeof(float) * 12 * nf, stream>>>(p);
        const cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

}  // namespace cnrma
