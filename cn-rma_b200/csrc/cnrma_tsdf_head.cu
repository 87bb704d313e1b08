// TSDF head hand-over: one scale of AtlasTSDFHead.forward,
// projects/mvsdetection/models/atlas_head.py:38-52 ("ah.py") -- the step that produces `scene_tsdf_004`, the volume
// Stage B marches through.
//
//   tsdf = tanh(conv1x1x1(x)) * label_smoothing                                  ah.py:40
//   prev = nearest x2 upsample of the previous (coarser) scale's tsdf            ah.py:44-46
//   surface = |prev| < sparse_threshold                                          ah.py:47
//   tsdf[~surface] = sign(prev[~surface]) * .999                                 ah.py:48
//
// The reference spends a cudnn convolution over the whole volume plus ~10 full-volume elementwise / index kernels
// (one with a nonzero() synchronisation) per scale.  Here: one pass, one thread per 4 consecutive voxels of a
// z-column, channel planes streamed with 16-byte loads -- and a voxel group whose coarse parents are both
// non-surface never touches x at all: its output does not depend on it.  On real scenes most of the fine volume is
// in that state, so the pass reads a fraction of the compulsory bytes of the dense formulation.  HBM-bound.
//
// The channel dot product is a floating-point reduction whose order cudnn does not specify: this row's parity
// is a tolerance (1e-5) with the surface mask compared outside a band around the threshold (tests/test_tsdf_head.py).
#include "cnrma_internal.cuh"

namespace cnrma {

constexpr int kHeadMaxChannels = 1024;
constexpr int kHeadThreads = 256;
constexpr int kHeadChunk = 32;          // channels per backward pass (accumulators held in registers)
constexpr int kHeadMaxBlocks = 1024;    // backward: rows of the per-block partial sums

struct HeadParams {
    const void *x;
    int64_t stride_c, stride_v;
    int C, nx, ny, nz;
    const float *weight, *prev;
    float ls, thr;
    float *tsdf;
    uint8_t *mask;
};

__device__ __forceinline__ float load_as_float(const float *p) { return __ldg(p); }
__device__ __forceinline__ float load_as_float(const __nv_bfloat16 *p) {
    return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p)) << 16);
}

template <typename T, int VEC>
struct VoxelVec;
template <>
struct VoxelVec<float, 4> {
    __device__ __forceinline__ static void load(const float *p, float (&v)[4]) {
        const float4 t = __ldcs(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <>
struct VoxelVec<__nv_bfloat16, 4> {
    __device__ __forceinline__ static void load(const __nv_bfloat16 *p, float (&v)[4]) {
        const uint2 t = __ldcs(reinterpret_cast<const uint2 *>(p));
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
};
template <typename T>
struct VoxelVec<T, 1> {
    __device__ __forceinline__ static void load(const T *p, float (&v)[1]) { v[0] = load_as_float(p); }
};

// The coarse parents of VEC consecutive voxels of one z-column starting at flat index v0 (z0 % VEC == 0).
template <int VEC>
__device__ __forceinline__ void parents(const HeadParams &p, int64_t v0, float (&pv)[VEC]) {
    const int z0 = (int)(v0 % p.nz);
    const int64_t xy = v0 / p.nz;
    const int y = (int)(xy % p.ny), x = (int)(xy / p.ny);
    const float *col = p.prev + ((int64_t)(x >> 1) * (p.ny >> 1) + (y >> 1)) * (p.nz >> 1);
#pragma unroll
    for (int i = 0; i < VEC; ++i) pv[i] = __ldg(col + ((z0 + i) >> 1));
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kHeadThreads) tsdf_head_scale_kernel(const HeadParams p) {
    __shared__ float sw[kHeadMaxChannels];
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) sw[c] = __ldg(p.weight + c);
    __syncthreads();
    const int64_t nvox = (int64_t)p.nx * p.ny * p.nz;
    const int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (v0 >= nvox) return;
    float pv[VEC];
    bool surf[VEC], any = true;
    if (p.prev) {
        parents<VEC>(p, v0, pv);
        any = false;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            surf[i] = fabsf(pv[i]) < p.thr;
            any |= surf[i];
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) surf[i] = true;
    }
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.0f;
    if (any) {
        const T *px = static_cast<const T *>(p.x) + v0 * p.stride_v;
#pragma unroll 8
        for (int c = 0; c < p.C; ++c) {
            float xv[VEC];
            VoxelVec<T, VEC>::load(px + (int64_t)c * p.stride_c, xv);
            const float w = sw[c];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = __fmaf_rn(w, xv[i], acc[i]);
        }
    }
    float out[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        if (surf[i]) out[i] = __fmul_rn(tanhf(acc[i]), p.ls);
        else out[i] = __fmul_rn((float)((pv[i] > 0.0f) - (pv[i] < 0.0f)), 0.999f);
    }
    if (VEC == 4) {
        __stcs(reinterpret_cast<float4 *>(p.tsdf + v0), make_float4(out[0], out[1], out[2], out[3]));
        if (p.mask) {
            const uint32_t m = (uint32_t)surf[0] | ((uint32_t)surf[1] << 8) | ((uint32_t)surf[2] << 16) | ((uint32_t)surf[3] << 24);
            *reinterpret_cast<uint32_t *>(p.mask + v0) = m;
        }
    } else {
        p.tsdf[v0] = out[0];
        if (p.mask) p.mask[v0] = (uint8_t)surf[0];
    }
}

cudaError_t run_tsdf_head_scale(const void *x, int dtype, int C, int nx, int ny, int nz, int64_t stride_c,
                                int64_t stride_v, const float *weight, const float *prev, float ls, float thr,
                                float *tsdf, uint8_t *mask, cudaStream_t stream) {
    HeadParams p{x, stride_c, stride_v, C, nx, ny, nz, weight, prev, ls, thr, tsdf, mask};
    const int64_t nvox = (int64_t)nx * ny * nz;
    const int esize = dtype == CNRMA_BF16 ? 2 : 4;
    const bool vec = stride_v == 1 && nz % 4 == 0 && (stride_c * esize) % (4 * esize) == 0 &&
                     reinterpret_cast<uintptr_t>(x) % (4 * esize) == 0 && reinterpret_cast<uintptr_t>(tsdf) % 16 == 0 &&
                     (!mask || reinterpret_cast<uintptr_t>(mask) % 4 == 0);
    const int64_t threads = vec ? nvox / 4 : nvox;
    const unsigned blocks = (unsigned)((threads + kHeadThreads - 1) / kHeadThreads);
    if (dtype == CNRMA_BF16) {
        if (vec) tsdf_head_scale_kernel<__nv_bfloat16, 4><<<blocks, kHeadThreads, 0, stream>>>(p);
        else tsdf_head_scale_kernel<__nv_bfloat16, 1><<<blocks, kHeadThreads, 0, stream>>>(p);
    } else {
        if (vec) tsdf_head_scale_kernel<float, 4><<<blocks, kHeadThreads, 0, stream>>>(p);
        else tsdf_head_scale_kernel<float, 1><<<blocks, kHeadThreads, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

// ---- backward --------------------------------------------------------------------------------------------------
// autograd of ah.py:40-48 for one scale.  With g = dL/dtsdf and t = tanh(pre) = tsdf / label_smoothing:
//   gp = surface ? g * label_smoothing * (1 - t^2) : 0        (the index_put of ah.py:48 cuts the gradient)
//   dL/dx[c, v] = gp[v] * w[c]                                 dL/dw[c] = sum_v gp[v] * x[c, v]
// No gradient reaches the previous scale (sign() and the mask are piecewise constant).
// Persistent blocks walk the voxels; blockIdx.y selects a chunk of 32 channels whose weight-gradient accumulators
// live in registers; per-block partial sums go to the workspace and a second kernel adds them in block order in
// double -- deterministic, no atomics.
struct HeadBackwardParams {
    HeadParams f;
    const float *grad_tsdf;
    float *grad_x;        // same strides as x, or null
    float *partial;       // [gridDim.x, C] or null (no weight gradient wanted)
};

template <int VEC>
__global__ void __launch_bounds__(kHeadThreads) tsdf_head_backward_kernel(const HeadBackwardParams p) {
    __shared__ float sw[kHeadChunk];
    __shared__ float sred[kHeadThreads / 32][kHeadChunk];
    const int c0 = blockIdx.y * kHeadChunk;
    const int nc = min(kHeadChunk, p.f.C - c0);
    if (threadIdx.x < kHeadChunk) sw[threadIdx.x] = threadIdx.x < nc ? __ldg(p.f.weight + c0 + threadIdx.x) : 0.0f;
    __syncthreads();
    const int64_t nvox = (int64_t)p.f.nx * p.f.ny * p.f.nz;
    float acc[kHeadChunk];
#pragma unroll
    for (int c = 0; c < kHeadChunk; ++c) acc[c] = 0.0f;
    const float *x = static_cast<const float *>(p.f.x);
    for (int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC; v < nvox;
         v += (int64_t)gridDim.x * blockDim.x * VEC) {
        float gp[VEC];
        bool any = false;
        {
            float pv[VEC], out[VEC], g[VEC];
            if (p.f.prev) parents<VEC>(p.f, v, pv);
            VoxelVec<float, VEC>::load(p.f.tsdf + v, out);
            VoxelVec<float, VEC>::load(p.grad_tsdf + v, g);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const bool surf = !p.f.prev || fabsf(pv[i]) < p.f.thr;
                const float t = __fdiv_rn(out[i], p.f.ls);
                gp[i] = surf ? g[i] * p.f.ls * (1.0f - t * t) : 0.0f;
                any |= surf;
            }
        }
        const int64_t base = v * p.f.stride_v + (int64_t)c0 * p.f.stride_c;
        if (p.grad_x) {
#pragma unroll
            for (int c = 0; c < kHeadChunk; ++c)
                if (c < nc) {
                    float *dst = p.grad_x + base + (int64_t)c * p.f.stride_c;
                    if (VEC == 4) __stcs(reinterpret_cast<float4 *>(dst), make_float4(gp[0] * sw[c], gp[1] * sw[c], gp[2] * sw[c], gp[3] * sw[c]));
                    else __stcs(dst, gp[0] * sw[c]);
                }
        }
        if (p.partial && any) {
#pragma unroll
            for (int c = 0; c < kHeadChunk; ++c)
                if (c < nc) {
                    float xv[VEC];
                    VoxelVec<float, VEC>::load(x + base + (int64_t)c * p.f.stride_c, xv);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[c] = __fmaf_rn(gp[i], xv[i], acc[c]);
                }
        }
    }
    if (!p.partial) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < kHeadChunk; ++c) {
        float s = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sred[warp][c] = s;
    }
    __syncthreads();
    if (threadIdx.x < nc) {
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < kHeadThreads / 32; ++w) s += sred[w][threadIdx.x];
        p.partial[(int64_t)blockIdx.x * p.f.C + c0 + threadIdx.x] = s;
    }
}

// One warp per channel: lanes stride over the per-block partial sums in double, fixed shuffle tree -> deterministic.
__global__ void tsdf_head_weight_grad_kernel(const float *partial, int blocks, int C, float *grad_weight) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < blocks; b += 32) s += (double)partial[(int64_t)b * C + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) grad_weight[c] = (float)s;
}

size_t tsdf_head_workspace_bytes(int C) { return (size_t)kHeadMaxBlocks * (size_t)C * sizeof(float); }

cudaError_t run_tsdf_head_backward(const float *x, int C, int nx, int ny, int nz, int64_t stride_c, int64_t stride_v,
                                   const float *weight, const float *prev, const float *tsdf, const float *grad_tsdf,
                                   float ls, float thr, float *grad_x, float *grad_weight, void *workspace,
                                   cudaStream_t stream) {
    HeadBackwardParams p;
    p.f = HeadParams{x, stride_c, stride_v, C, nx, ny, nz, weight, prev, ls, thr, const_cast<float *>(tsdf), nullptr};
    p.grad_tsdf = grad_tsdf;
    p.grad_x = grad_x;
    p.partial = grad_weight ? static_cast<float *>(workspace) : nullptr;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t nvox = (int64_t)nx * ny * nz;
    const int chunks = (C + kHeadChunk - 1) / kHeadChunk;
    const bool vec = stride_v == 1 && nz % 4 == 0 && stride_c % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(tsdf) % 16 == 0 && reinterpret_cast<uintptr_t>(grad_tsdf) % 16 == 0 &&
                     (!grad_x || reinterpret_cast<uintptr_t>(grad_x) % 16 == 0);
    const int64_t threads = vec ? nvox / 4 : nvox;
    int64_t bx = (threads + kHeadThreads - 1) / kHeadThreads;
    const int64_t cap = std::min<int64_t>(kHeadMaxBlocks, std::max<int64_t>(1, (int64_t)sms * 2 / chunks));
    if (bx > cap) bx = cap;
    if (vec) tsdf_head_backward_kernel<4><<<dim3((unsigned)bx, (unsigned)chunks), kHeadThreads, 0, stream>>>(p);
    else tsdf_head_backward_kernel<1><<<dim3((unsigned)bx, (unsigned)chunks), kHeadThreads, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !grad_weight) return e;
    tsdf_head_weight_grad_kernel<<<(C + 3) / 4, 128, 0, stream>>>(p.partial, (int)bx, C, grad_weight);
    return cudaGetLastError();
}

}  // namespace cnrma
