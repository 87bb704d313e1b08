// Internal launch interface between cnrma_abi.cu and the kernel translation units.
#pragma once

#include <algorithm>

#include "cnrma_common.cuh"

namespace cnrma {

// Scratch layout of the ray-march path (byte offsets into the caller's workspace).
struct RmaWorkspace {
    int64_t rays;     // V*H*W
    int64_t blocks;   // ceil(rays / kRayThreads)
    int cap;          // (step, weight) records per ray
    size_t off_counts, off_blk_rows, off_blk_wsum, off_blk_off, off_rec_w, off_rec_i, off_dist, off_sigmoid, off_cell, total;
};

// cnrma_stage_a.cu
cudaError_t run_project_views(const GridDev &g, const float *proj, int64_t proj_stride, int V, float stride, int H,
                              int W, int32_t *px, int32_t *py, uint8_t *valid, cudaStream_t stream);
cudaError_t run_aggregate_views(const GridDev &g, const cnrma_features &f, int v0, int nv, const float *proj,
                                int64_t proj_stride, float stride, uint32_t flags, float *volume, int64_t vsv,
                                int64_t vsc, int32_t *count, uint8_t *valid, int max_chunk_bytes, cudaStream_t stream,
                                const OutputRoute *route = nullptr, int reserve_ctas = 0);
bool list_kernel_supports(int V, int H, int W);
bool long_list_supports(int channels, int dtype);
cudaError_t run_aggregate_list(const GridDev &g, const cnrma_features &f, int v0, int nv, const float *proj,
                               int64_t proj_stride, float stride, uint32_t flags, float *volume, int64_t vsv, int64_t vsc,
                               int32_t *count, uint8_t *valid, cudaStream_t stream, const OutputRoute *route = nullptr,
                               int reserve_ctas = 0);
cudaError_t run_aggregate_bilinear(const GridDev &g, const cnrma_features &f, const float *proj, int64_t proj_stride,
                                   float stride, uint32_t flags, float *volume, int32_t *count, uint8_t *valid,
                                   cudaStream_t stream);
cudaError_t run_finalize_routed(const float *recv, int n_src, int slab, int row_floats, int rows, int C, int mean,
                                float *volume, int32_t *count, uint8_t *valid, cudaStream_t stream);
cudaError_t run_selftest_count_division(int max_n, unsigned long long *mismatches, cudaStream_t stream);
cudaError_t run_to_channels_last(const void *const *views_host, int views, int dtype, int C, int H, int W, int64_t sc,
                                 int64_t sy, int64_t sx, void *dst, cudaStream_t stream);

// cnrma_exchange.cu
cudaError_t run_mark_rows(const GridDev &box, const float *proj, int64_t proj_stride, int V, float stride, int H, int W,
                          uint32_t *bitmap, int parts, int64_t part_stride, cudaStream_t stream);
cudaError_t run_pull_rows(const uint32_t *bitmap, uint32_t *done, int views, int H, int W, int row_bytes,
                          const void *const *src_views_host, void *dst, int64_t dst_vs, int ctas, unsigned int *work,
                          int first_view, int use_lsu, cudaStream_t stream);
int pull_default_ctas();

// cnrma_stage_b.cu
RmaWorkspace rma_workspace(int views, int height, int width, int grids, int mode, float threshold, int depth_points,
                           int64_t nvox = 0);
cudaError_t run_march(const GridDev &g, const float *pinv, int V, int H, int W, const float *tsdf, int N, float t_one,
                      int mode, float thr, int depth_points, void *workspace, const RmaWorkspace &ws,
                      cnrma_rma_result *result, cudaStream_t stream);
cudaError_t run_fill(const GridDev &g, const float *pinv, const cnrma_features &f, float t_one, int mode,
                     const void *workspace, const RmaWorkspace &ws, int normalize, const float *mean, float *rows,
                     int64_t row_stride, int64_t capacity, float *wsum, float *wtot, const uint8_t *sel_mask,
                     const int32_t *sel_prefix, const float *sel_off_host, int64_t sel_rows, cudaStream_t stream);
cudaError_t run_ray_parameters(const float *pinv, int V, int H, int W, float *o, float *d, cudaStream_t stream);
cudaError_t run_expand(int64_t rays, int N, const void *workspace, const RmaWorkspace &ws, float *weights,
                       uint8_t *keep, cudaStream_t stream);

// cnrma_handoff.cu
size_t handoff_workspace_bytes(int64_t M);
cudaError_t run_mask_prefix(const uint8_t *mask, int64_t M, void *workspace, int32_t *prefix, int64_t *total,
                            cudaStream_t stream);
cudaError_t run_select_rows(const float *rows, int64_t row_stride, int cols, int64_t M, const uint8_t *mask,
                            const int32_t *prefix, const float *offset3_host, float *out, int64_t out_stride,
                            int64_t capacity, cudaStream_t stream);

size_t quantize_workspace_bytes(int64_t n);
cudaError_t run_quantize_mark(const float *rows, int64_t stride, int64_t n, float vs, void *workspace, uint8_t *keep,
                              cudaStream_t stream);
cudaError_t run_quantize_compact(const float *rows, int64_t stride, int cols, int64_t n, float vs, const uint8_t *keep,
                                 const int32_t *prefix, float *out, int64_t out_stride, int32_t *cells, int64_t capacity,
                                 cudaStream_t stream);

size_t sample_workspace_bytes();
cudaError_t run_sample_mask(int64_t n, int64_t k, unsigned long long seed, void *workspace, uint8_t *mask,
                            cudaStream_t stream, const long long *n_dev = nullptr);

constexpr uint32_t kAggBilinearInternal = 0x80000000u;   // library-internal flag bit: run_aggregate_views in bilinear mode
constexpr int kListViewsMax = 96;     // the list kernel's per-voxel lists stop paying beyond this many views per launch
constexpr int kListViewsBatch = 63;   // batch size used to split longer view lists (32 voxels x 63 entries fit a warp's 8 KB)

// cnrma_tsdf_head.cu
cudaError_t run_tsdf_head_scale(const void *x, int dtype, int C, int nx, int ny, int nz, int64_t stride_c,
                                int64_t stride_v, const float *weight, const float *prev, float ls, float thr,
                                float *tsdf, uint8_t *mask, cudaStream_t stream);
size_t tsdf_head_workspace_bytes(int C);
cudaError_t run_tsdf_head_backward(const float *x, int C, int nx, int ny, int nz, int64_t stride_c, int64_t stride_v,
                                   const float *weight, const float *prev, const float *tsdf, const float *grad_tsdf,
                                   float ls, float thr, float *grad_x, float *grad_weight, void *workspace,
                                   cudaStream_t stream);
// cnrma_fusion.cu
cudaError_t run_tsdf_integrate(const GridDev &g, const float *proj, int64_t proj_stride, int frames,
                               const float *const *depth_host, const float *const *color_host,
                               const int64_t *const *label_host, int H, int W, float trunc_margin, float *tsdf,
                               float *weight, float *color, int64_t *label, cudaStream_t stream);

// cnrma_backward.cu
cudaError_t run_aggregate_views_backward(const GridDev &g, const cnrma_features &gf, const float *proj,
                                         int64_t proj_stride, float stride, uint32_t flags, const float *grad_volume,
                                         int64_t vsv, int64_t vsc, const int32_t *count, cudaStream_t stream);
cudaError_t run_fill_backward(const cnrma_features &gf, const void *workspace, const RmaWorkspace &ws, int normalize,
                              const float *mean, const float *grad_rows, int64_t row_stride, cudaStream_t stream);

}  // namespace cnrma
