// extern "C" surface of libcnrma_b200.so (include/cnrma_b200.h): argument validation and launches.
// No torch types, no allocation, no global state beyond a thread-local last-CUDA-error slot.
#include <cmath>
#include <cstdlib>

#include "cnrma_internal.cuh"

namespace cnrma {

static thread_local int g_last_cuda_error = 0;

static Tuning read_tuning() {
    Tuning t;
    auto num = [](const char *name) { const char *e = std::getenv(name); return e ? std::atoi(e) : 0; };
    if (const char *e = std::getenv("CNRMA_AGG_KERNEL")) t.agg_kernel = (e[0] == 'l') ? 1 : 0;
    t.agg_slab = num("CNRMA_AGG_SLAB");
    t.agg_tile = num("CNRMA_AGG_TILE");
    if (const char *e = std::getenv("CNRMA_AGG_CULL")) t.agg_cull = std::atoi(e);
    if (const char *e = std::getenv("CNRMA_AGG_PIPE")) t.agg_pipe = std::atoi(e);
    t.agg_chunk_bytes = num("CNRMA_AGG_CHUNK_BYTES");
    t.agg_warp_buffer = num("CNRMA_AGG_WARP_BUFFER");
    t.agg_list_views = num("CNRMA_AGG_LIST_VIEWS");
    if (const char *e = std::getenv("CNRMA_AGG_BWD_KERNEL")) t.agg_bwd_kernel = (e[0] == 'l') ? 1 : 0;
    t.bilinear_simple = std::getenv("CNRMA_BILINEAR_SIMPLE") != nullptr;
    t.march_unfused = std::getenv("CNRMA_MARCH_UNFUSED_PREPASS") != nullptr;
    if (const char *e = std::getenv("CNRMA_MARCH_JUMP")) t.march_jump = std::atoi(e);
    if (const char *e = std::getenv("CNRMA_FILL_KERNEL")) t.fill_kernel = (e[0] == 'p') ? 1 : 0;
    t.fill_stage_half = num("CNRMA_FILL_STAGE_HALF");
    if (const char *e = std::getenv("CNRMA_FILL_SELECT_KERNEL")) t.fill_select_scalar = (e[0] == 's');
    return t;
}

static Tuning &tuning_slot() {
    static Tuning t = read_tuning();   // thread-safe one-time initialisation
    return t;
}

const Tuning &tuning() { return tuning_slot(); }

static int fail_cuda(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return CNRMA_ERR_CUDA;
}

static bool grid_ok(const cnrma_grid *g) {
    if (!g || g->nx <= 0 || g->ny <= 0 || g->nz <= 0) return false;
    if (!(g->voxel_size > 0.0f)) return false;
    return (int64_t)g->nx * g->ny * g->nz < (int64_t)1 << 31;
}

// The Stage A sweep decodes voxel coordinates with fast_divmod (cnrma_common.cuh), exact for quotients below 2^22:
// the largest quotient is an (x, y) column index, so planes of 4 M columns and more are refused instead of mis-decoded.
static bool sweep_ok(int nx, int ny) { return (int64_t)nx * ny < ((int64_t)1 << 22); }

static int features_ok(const cnrma_features *f, bool need_channels_last, bool allow_empty = false) {
    if (f && allow_empty && f->views == 0 && f->channels > 0 && (f->dtype == CNRMA_F32 || f->dtype == CNRMA_BF16))
        return f->channels % ((f->dtype == CNRMA_BF16) ? 8 : 4) == 0 ? CNRMA_OK : CNRMA_ERR_LAYOUT;
    if (!f || !f->view_ptrs_host || f->views <= 0 || f->channels <= 0 || f->height <= 0 || f->width <= 0)
        return CNRMA_ERR_ARG;
    if (f->dtype != CNRMA_F32 && f->dtype != CNRMA_BF16) return CNRMA_ERR_ARG;
    for (int v = 0; v < f->views; ++v)
        if (!f->view_ptrs_host[v]) return CNRMA_ERR_ARG;
    if ((int64_t)f->height * f->stride_y + (int64_t)f->width * f->stride_x >= (int64_t)1 << 31) return CNRMA_ERR_UNSUPPORTED;
    if (need_channels_last) {
        const int e = (f->dtype == CNRMA_BF16) ? 8 : 4;
        if (f->stride_c != 1 || f->channels % e != 0 || f->stride_x % e != 0 || f->stride_y % e != 0)
            return CNRMA_ERR_LAYOUT;
        for (int v = 0; v < f->views; ++v)
            if (reinterpret_cast<uintptr_t>(f->view_ptrs_host[v]) % 16 != 0) return CNRMA_ERR_LAYOUT;
    }
    return CNRMA_OK;
}

static int device_ok() {
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail_cuda(e);
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return fail_cuda(e);
    return major == 10 ? CNRMA_OK : CNRMA_ERR_DEVICE;
}

}  // namespace cnrma

using namespace cnrma;

extern "C" {

int cnrma_abi_version(void) { return CNRMA_ABI_VERSION; }

const char *cnrma_status_string(int status) {
    switch (status) {
        case CNRMA_OK: return "ok";
        case CNRMA_ERR_ARG: return "invalid argument";
        case CNRMA_ERR_LAYOUT: return "feature maps must be channels-last and 16-byte aligned (see cnrma_to_channels_last)";
        case CNRMA_ERR_CAPACITY: return "output buffer or workspace too small";
        case CNRMA_ERR_DEVICE: return "CUDA device is not compute capability 10.x (B200)";
        case CNRMA_ERR_CUDA: return "CUDA runtime error (see cnrma_last_cuda_error)";
        case CNRMA_ERR_UNSUPPORTED: return "unsupported shape";
    }
    return "unknown status";
}

int cnrma_last_cuda_error(void) { return g_last_cuda_error; }

int cnrma_check_device(void) { return device_ok(); }

void cnrma_reload_tuning(void) { tuning_slot() = read_tuning(); }

int cnrma_project_views(const cnrma_grid *grid, const float *projections, int64_t proj_view_stride, int views,
                        float stride, int height, int width, int32_t *px, int32_t *py, uint8_t *valid, void *stream) {
    if (!grid_ok(grid) || !projections || views <= 0 || height <= 0 || width <= 0 || !(stride > 0.0f))
        return CNRMA_ERR_ARG;
    if (views > 65535) return CNRMA_ERR_UNSUPPORTED;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_project_views(to_dev(*grid), projections, proj_view_stride, views, stride, height, width,
                                            px, py, valid, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

static bool box_ok(const cnrma_grid *g, const cnrma_box *b) {
    if (!b) return false;
    const int n[3] = {g->nx, g->ny, g->nz};
    for (int a = 0; a < 3; ++a)
        if (b->lo[a] < 0 || b->dim[a] <= 0 || (int64_t)b->lo[a] + b->dim[a] > n[a]) return false;
    return true;
}

static int aggregate_views_impl(const cnrma_grid *grid, const cnrma_box *box, const cnrma_features *features,
                                const float *projections, int64_t proj_view_stride, float stride, uint32_t flags,
                                float *volume, int64_t vol_stride_voxel, int64_t vol_stride_channel, int32_t *count,
                                uint8_t *valid, int reserve_ctas, void *stream) {
    if (!grid_ok(grid) || !features || !volume || !count || !(stride > 0.0f)) return CNRMA_ERR_ARG;
    if (box && !box_ok(grid, box)) return CNRMA_ERR_ARG;
    if (!(box ? sweep_ok(box->dim[0], box->dim[1]) : sweep_ok(grid->nx, grid->ny))) return CNRMA_ERR_UNSUPPORTED;
    if (!projections && features->views > 0) return CNRMA_ERR_ARG;
    if (flags & ~(CNRMA_AGG_ACCUMULATE | CNRMA_AGG_MEAN | CNRMA_AGG_COUNT_F32)) return CNRMA_ERR_ARG;
    const int fs = features_ok(features, true, (flags & CNRMA_AGG_ACCUMULATE) != 0);
    if (fs != CNRMA_OK) return fs;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    // Tuning knob (DESIGN.md "K_A"): cap on the bytes of a feature row gathered per channel pass (<= 1024).
    const int max_chunk_vecs = tuning().agg_chunk_bytes;
    const GridDev g = box ? to_dev_box(*grid, *box) : to_dev(*grid);
    // Views go out in batches that accumulate into the volume in view order (the fp32 chain of the reference's
    // `self.volume + volume` loop is unchanged).  Short rows with many views are split into batches the list kernel
    // can serve (DESIGN.md "K_A'"): the extra read-modify-write of the volume costs far less than the per-row
    // overhead of the TMA kernel on rows below 512 bytes.
    int batch = kMaxViewsPerLaunch;
    const int row_bytes = features->channels * (features->dtype == CNRMA_BF16 ? 2 : 4);
    if (row_bytes < 512 && features->views > kListViewsMax && !long_list_supports(features->channels, features->dtype)) {
        const int per = tuning().agg_list_views > 0 ? tuning().agg_list_views : kListViewsBatch;   // tuning aid
        if (per >= 1 && per <= kListViewsMax) {
            const int nbatch = (features->views + per - 1) / per;
            batch = (features->views + nbatch - 1) / nbatch;
        }
    }
    for (int v0 = 0; v0 < features->views || v0 == 0; v0 += batch) {
        const int nv = (features->views - v0 < batch) ? (features->views - v0) : batch;
        uint32_t fl = flags & (CNRMA_AGG_ACCUMULATE | CNRMA_AGG_COUNT_F32);
        if (v0 > 0) fl |= CNRMA_AGG_ACCUMULATE;
        if ((flags & CNRMA_AGG_MEAN) && v0 + nv == features->views) fl |= CNRMA_AGG_MEAN;
        const cudaError_t e = run_aggregate_views(g, *features, v0, nv, projections + (int64_t)v0 * proj_view_stride,
                                                  proj_view_stride, stride, fl, volume, vol_stride_voxel,
                                                  vol_stride_channel, count, valid, max_chunk_vecs,
                                                  static_cast<cudaStream_t>(stream), nullptr, reserve_ctas);
        if (e != cudaSuccess) return fail_cuda(e);
    }
    return CNRMA_OK;
}

int cnrma_aggregate_views(const cnrma_grid *grid, const cnrma_features *features, const float *projections,
                          int64_t proj_view_stride, float stride, uint32_t flags, float *volume,
                          int64_t vol_stride_voxel, int64_t vol_stride_channel, int32_t *count, uint8_t *valid,
                          void *stream) {
    return aggregate_views_impl(grid, nullptr, features, projections, proj_view_stride, stride, flags, volume,
                                vol_stride_voxel, vol_stride_channel, count, valid, 0, stream);
}

int cnrma_aggregate_views_box(const cnrma_grid *grid, const cnrma_box *box, const cnrma_features *features,
                              const float *projections, int64_t proj_view_stride, float stride, uint32_t flags,
                              float *volume, int64_t vol_stride_voxel, int64_t vol_stride_channel, int32_t *count,
                              uint8_t *valid, int reserve_ctas, void *stream) {
    if (!box || reserve_ctas < 0) return CNRMA_ERR_ARG;
    return aggregate_views_impl(grid, box, features, projections, proj_view_stride, stride, flags, volume,
                                vol_stride_voxel, vol_stride_channel, count, valid, reserve_ctas, stream);
}

int cnrma_mark_rows(const cnrma_grid *grid, const cnrma_box *box, const float *projections, int64_t proj_view_stride,
                    int views, float stride, int height, int width, uint32_t *bitmap, int parts, int64_t part_stride,
                    void *stream) {
    if (!grid_ok(grid) || !box_ok(grid, box) || !projections || !bitmap || views <= 0 || height <= 0 || width <= 0 ||
        !(stride > 0.0f) || parts < 1 || parts > 16 || parts > box->dim[0] || (parts > 1 && part_stride <= 0))
        return CNRMA_ERR_ARG;
    if ((int64_t)height * width >= ((int64_t)1 << 30)) return CNRMA_ERR_UNSUPPORTED;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_mark_rows(to_dev_box(*grid, *box), projections, proj_view_stride, views, stride, height,
                                        width, bitmap, parts, part_stride, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_pull_rows(const uint32_t *bitmap, uint32_t *done, int views, int height, int width, int row_bytes,
                    const void *const *src_view_ptrs_host, void *dst, int64_t dst_view_stride, int ctas, uint32_t *work,
                    int first_view, void *stream) {
    if (!bitmap || !src_view_ptrs_host || !dst || views <= 0 || height <= 0 || width <= 0 || ctas < 0 || first_view < 0 ||
        first_view >= views)
        return CNRMA_ERR_ARG;
    if (views > kMaxViewsPerLaunch) return CNRMA_ERR_UNSUPPORTED;
    if (row_bytes < 16 || row_bytes % 16 != 0 || row_bytes > 8192) return CNRMA_ERR_LAYOUT;
    if (reinterpret_cast<uintptr_t>(dst) % 16 != 0 || dst_view_stride % 16 != 0) return CNRMA_ERR_LAYOUT;
    for (int v = 0; v < views; ++v)
        if (!src_view_ptrs_host[v] || reinterpret_cast<uintptr_t>(src_view_ptrs_host[v]) % 16 != 0) return CNRMA_ERR_LAYOUT;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_pull_rows(bitmap, done, views, height, width, row_bytes, src_view_ptrs_host, dst,
                                        dst_view_stride, ctas & 0xFFFF, work, first_view, (ctas >> 16) & 1,
                                        static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_pull_default_ctas(void) { return pull_default_ctas(); }

int cnrma_aggregate_views_bilinear(const cnrma_grid *grid, const cnrma_features *features, const float *projections,
                                   int64_t proj_view_stride, float stride, uint32_t flags, float *volume, int32_t *count,
                                   uint8_t *valid, void *stream) {
    if (!grid_ok(grid) || !features || !projections || !volume || !count || !(stride > 0.0f)) return CNRMA_ERR_ARG;
    if (flags & ~CNRMA_AGG_MEAN) return CNRMA_ERR_ARG;
    const int fs = features_ok(features, true);
    if (fs != CNRMA_OK) return fs;
    if (features->views > kMaxViewsPerLaunch || !sweep_ok(grid->nx, grid->ny)) return CNRMA_ERR_UNSUPPORTED;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    cudaError_t e;
    if (tuning().bilinear_simple)   // the first, simple kernel (kept as a second opinion for the tests)
        e = run_aggregate_bilinear(to_dev(*grid), *features, projections, proj_view_stride, stride, flags, volume, count,
                                   valid, static_cast<cudaStream_t>(stream));
    else
        e = run_aggregate_views(to_dev(*grid), *features, 0, features->views, projections, proj_view_stride, stride,
                                (flags & CNRMA_AGG_MEAN) | kAggBilinearInternal, volume, features->channels, 1, count, valid,
                                0, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_aggregate_views_routed(const cnrma_grid *grid, const cnrma_features *features, const float *projections,
                                 int64_t proj_view_stride, float stride, int n_owners, int slab_voxels, int row_floats,
                                 float *const *owner_rows_host, void *stream) {
    if (!grid_ok(grid) || !features || !projections || !owner_rows_host || !(stride > 0.0f)) return CNRMA_ERR_ARG;
    if (n_owners < 1 || n_owners > kMaxOwners || slab_voxels < 1) return CNRMA_ERR_ARG;
    const int fs = features_ok(features, true);
    if (fs != CNRMA_OK) return fs;
    if (features->views < 1 || features->views > kMaxViewsPerLaunch || !sweep_ok(grid->nx, grid->ny))
        return CNRMA_ERR_UNSUPPORTED;
    const int64_t nvox = (int64_t)grid->nx * grid->ny * grid->nz;
    if ((int64_t)n_owners * slab_voxels < nvox || row_floats < features->channels + 1 || row_floats % 4 != 0)
        return CNRMA_ERR_ARG;
    OutputRoute route;
    route.n_owners = n_owners;
    route.slab = slab_voxels;
    route.row_floats = row_floats;
    route.inv_slab = 1.0f / (float)slab_voxels;
    for (int o = 0; o < n_owners; ++o) {
        if (!owner_rows_host[o] || reinterpret_cast<uintptr_t>(owner_rows_host[o]) % 16 != 0) return CNRMA_ERR_LAYOUT;
        route.owner_base[o] = owner_rows_host[o];
    }
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    // volume / count pointers are unused in routed mode (the route carries the destinations)
    const cudaError_t e = run_aggregate_views(to_dev(*grid), *features, 0, features->views, projections, proj_view_stride,
                                              stride, 0, route.owner_base[0], row_floats, 1, nullptr, nullptr, 0,
                                              static_cast<cudaStream_t>(stream), &route);
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_finalize_routed(const float *recv, int n_src, int slab_voxels, int row_floats, int rows, int channels, int mean,
                          float *volume, int32_t *count, uint8_t *valid, void *stream) {
    if (!recv || !volume || !count || n_src < 1 || slab_voxels < 1 || rows < 0 || rows > slab_voxels || channels < 4 ||
        channels % 4 != 0 || row_floats < channels + 1 || row_floats % 4 != 0)
        return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_finalize_routed(recv, n_src, slab_voxels, row_floats, rows, channels, mean, volume, count, valid,
                                              static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_selftest_count_division(int max_n, uint64_t *mismatches, void *stream) {
    if (!mismatches || max_n < 1 || max_n > 65535) return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_selftest_count_division(max_n, reinterpret_cast<unsigned long long *>(mismatches),
                                                      static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_to_channels_last(const cnrma_features *src, void *dst, void *stream) {
    const int fs = features_ok(src, false);
    if (fs != CNRMA_OK) return fs;
    if (!dst) return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_to_channels_last(src->view_ptrs_host, src->views, src->dtype, src->channels, src->height,
                                               src->width, src->stride_c, src->stride_y, src->stride_x, dst,
                                               static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

float cnrma_t_one(const cnrma_grid *grid, double voxel_size, int grids) {
    // rm.py:710-711: sqrt(X**2 + Y**2 + Z**2) * voxel_size / N in python doubles; `arange * t_one` then
    // multiplies by the value rounded to float.
    const double x = grid->nx, y = grid->ny, z = grid->nz;
    return (float)(std::sqrt(x * x + y * y + z * z) * voxel_size / (double)grids);
}

int cnrma_ray_parameters(const float *pinv, int views, int height, int width, float *o, float *d, void *stream) {
    if (!pinv || !o || !d || views <= 0 || height <= 0 || width <= 0) return CNRMA_ERR_ARG;
    const int dv = device_ok();
    if (dv != CNRMA_OK) return dv;
    const cudaError_t e = run_ray_parameters(pinv, views, height, width, o, d, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_rma_workspace_bytes(const cnrma_grid *grid, int views, int height, int width, int grids, int mode,
                              float threshold, int depth_points, size_t *bytes) {
    if (!grid_ok(grid) || !bytes || views <= 0 || height <= 0 || width <= 0 || grids <= 0) return CNRMA_ERR_ARG;
    if (mode != CNRMA_MARCH_NEUS && mode != CNRMA_MARCH_DEPTH) return CNRMA_ERR_ARG;
    if (mode == CNRMA_MARCH_DEPTH && depth_points < 0) return CNRMA_ERR_ARG;
    *bytes = rma_workspace(views, height, width, grids, mode, threshold, depth_points,
                           (int64_t)grid->nx * grid->ny * grid->nz).total;
    return CNRMA_OK;
}

int cnrma_rma_march(const cnrma_grid *grid, const float *pinv, int views, int height, int width, const float *tsdf,
                    int grids, float t_one, int mode, float threshold, int depth_points, void *workspace,
                    size_t workspace_bytes, cnrma_rma_result *result, void *stream) {
    if (!grid_ok(grid) || !pinv || !tsdf || !workspace || !result) return CNRMA_ERR_ARG;
    size_t need = 0;
    const int s = cnrma_rma_workspace_bytes(grid, views, height, width, grids, mode, threshold, depth_points, &need);
    if (s != CNRMA_OK) return s;
    if (workspace_bytes < need) return CNRMA_ERR_CAPACITY;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return CNRMA_ERR_LAYOUT;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const RmaWorkspace ws = rma_workspace(views, height, width, grids, mode, threshold, depth_points,
                                          (int64_t)grid->nx * grid->ny * grid->nz);
    if (ws.blocks >= ((int64_t)1 << 31)) return CNRMA_ERR_UNSUPPORTED;
    const cudaError_t e = run_march(to_dev(*grid), pinv, views, height, width, tsdf, grids, t_one, mode, threshold,
                                    depth_points, workspace, ws, result, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

// fill / scatter share one launcher; the workspace layout is re-derived from the same capacity-defining
// arguments (grids, mode, threshold, depth_points) the march call was given.
static int fill_common(const cnrma_grid *grid, const float *pinv, const cnrma_features *features, int grids,
                       float t_one, int mode, float threshold, int depth_points, const void *workspace, int normalize,
                       const float *mean, float *rows, int64_t row_stride, int64_t capacity, float *wsum, float *wtot,
                       const uint8_t *sel_mask, const int32_t *sel_prefix, const float *sel_off_host, int64_t sel_rows,
                       void *stream) {
    if (!grid_ok(grid) || !pinv || !workspace) return CNRMA_ERR_ARG;
    const int fs = features_ok(features, false);
    if (fs != CNRMA_OK) return fs;
    if (features->stride_c != 1) return CNRMA_ERR_LAYOUT;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const RmaWorkspace ws = rma_workspace(features->views, features->height, features->width, grids, mode, threshold,
                                          depth_points);
    const cudaError_t e = run_fill(to_dev(*grid), pinv, *features, t_one, mode, workspace, ws, normalize, mean, rows,
                                   row_stride, capacity, wsum, wtot, sel_mask, sel_prefix, sel_off_host, sel_rows,
                                   static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_rma_fill(const cnrma_grid *grid, const float *pinv, const cnrma_features *features, int grids, float t_one,
                   int mode, float threshold, int depth_points, const void *workspace, const cnrma_rma_result *result,
                   int64_t rows_host, int normalize, const float *mean, float *rows, int64_t row_stride,
                   int64_t capacity, void *stream) {
    if (!rows || !result || capacity < 0) return CNRMA_ERR_ARG;
    if (!features) return CNRMA_ERR_ARG;
    const int cols = features->channels + (normalize ? 3 : 4);
    if (row_stride < cols) return CNRMA_ERR_ARG;
    if (rows_host >= 0 && capacity < rows_host) return CNRMA_ERR_CAPACITY;
    if (rows_host == 0 || capacity == 0) return CNRMA_OK;
    const float *mean_ptr = mean ? mean : &result->mean;
    return fill_common(grid, pinv, features, grids, t_one, mode, threshold, depth_points, workspace, normalize,
                       mean_ptr, rows, row_stride, capacity, nullptr, nullptr, nullptr, nullptr, nullptr, 0, stream);
}

int cnrma_rma_fill_selected(const cnrma_grid *grid, const float *pinv, const cnrma_features *features, int grids,
                            float t_one, int mode, float threshold, int depth_points, const void *workspace,
                            const cnrma_rma_result *result, int normalize, const float *mean, const uint8_t *mask,
                            const int32_t *prefix, int64_t mask_rows, const float *offset_host, float *rows,
                            int64_t row_stride, int64_t capacity, void *stream) {
    if (!rows || !result || !mask || !prefix || !offset_host || capacity < 0 || mask_rows < 0 || !features)
        return CNRMA_ERR_ARG;
    const int cols = features->channels + (normalize ? 3 : 4);
    if (row_stride < cols) return CNRMA_ERR_ARG;
    if (capacity == 0) return CNRMA_OK;
    const float *mean_ptr = mean ? mean : &result->mean;
    return fill_common(grid, pinv, features, grids, t_one, mode, threshold, depth_points, workspace, normalize,
                       mean_ptr, rows, row_stride, capacity, nullptr, nullptr, mask, prefix, offset_host, mask_rows, stream);
}

int cnrma_handoff_workspace_bytes(int64_t rows, size_t *bytes) {
    if (!bytes || rows < 0) return CNRMA_ERR_ARG;
    *bytes = handoff_workspace_bytes(rows);
    return CNRMA_OK;
}

int cnrma_mask_prefix(const uint8_t *mask, int64_t rows, void *workspace, size_t workspace_bytes, int32_t *prefix,
                      int64_t *kept, void *stream) {
    if (!mask || !workspace || !prefix || !kept || rows < 0) return CNRMA_ERR_ARG;
    if (rows >= ((int64_t)1 << 31)) return CNRMA_ERR_UNSUPPORTED;
    if (workspace_bytes < handoff_workspace_bytes(rows)) return CNRMA_ERR_CAPACITY;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_mask_prefix(mask, rows, workspace, prefix, kept, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_select_rows(const float *rows, int64_t row_stride, int cols, int64_t n_rows, const uint8_t *mask,
                      const int32_t *prefix, const float *offset_host, float *out, int64_t out_stride,
                      int64_t capacity, void *stream) {
    if (!rows || !mask || !prefix || !offset_host || !out || cols < 3 || row_stride < cols || out_stride < cols ||
        n_rows < 0 || capacity < 0)
        return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_select_rows(rows, row_stride, cols, n_rows, mask, prefix, offset_host, out, out_stride,
                                          capacity, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_rma_scatter(const cnrma_grid *grid, const float *pinv, const cnrma_features *features, int grids,
                      float t_one, int mode, float threshold, int depth_points, const void *workspace, float *wsum,
                      float *wtot, void *stream) {
    if (!wsum || !wtot) return CNRMA_ERR_ARG;
    return fill_common(grid, pinv, features, grids, t_one, mode, threshold, depth_points, workspace, 0, nullptr,
                       nullptr, 0, 0, wsum, wtot, nullptr, nullptr, nullptr, 0, stream);
}

int cnrma_aggregate_views_backward(const cnrma_grid *grid, const cnrma_features *grad_features, const float *projections,
                                   int64_t proj_view_stride, float stride, uint32_t flags, const float *grad_volume,
                                   int64_t vol_stride_voxel, int64_t vol_stride_channel, const int32_t *count,
                                   void *stream) {
    if (!grid_ok(grid) || !projections || !grad_volume || !count || !(stride > 0.0f)) return CNRMA_ERR_ARG;
    if (flags & ~CNRMA_AGG_MEAN) return CNRMA_ERR_ARG;
    const int fs = features_ok(grad_features, true);
    if (fs != CNRMA_OK) return fs;
    if (grad_features->dtype != CNRMA_F32 || !sweep_ok(grid->nx, grid->ny)) return CNRMA_ERR_UNSUPPORTED;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_aggregate_views_backward(to_dev(*grid), *grad_features, projections, proj_view_stride, stride,
                                                       flags, grad_volume, vol_stride_voxel, vol_stride_channel, count,
                                                       static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_rma_fill_backward(const cnrma_grid *grid, const cnrma_features *grad_features, int grids, int mode,
                            float threshold, int depth_points, const void *workspace, const cnrma_rma_result *result,
                            int normalize, const float *mean, const float *grad_rows, int64_t row_stride, void *stream) {
    if (!grid_ok(grid) || !workspace || !result || !grad_rows) return CNRMA_ERR_ARG;
    const int fs = features_ok(grad_features, false);
    if (fs != CNRMA_OK) return fs;
    if (grad_features->dtype != CNRMA_F32 || grad_features->stride_c != 1) return CNRMA_ERR_LAYOUT;
    if (row_stride < grad_features->channels + (normalize ? 3 : 4)) return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const RmaWorkspace ws = rma_workspace(grad_features->views, grad_features->height, grad_features->width, grids, mode,
                                          threshold, depth_points, (int64_t)grid->nx * grid->ny * grid->nz);
    const cudaError_t e = run_fill_backward(*grad_features, workspace, ws, normalize, mean ? mean : &result->mean,
                                            grad_rows, row_stride, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_sample_workspace_bytes(size_t *bytes) {
    if (!bytes) return CNRMA_ERR_ARG;
    *bytes = sample_workspace_bytes();
    return CNRMA_OK;
}

int cnrma_sample_mask(int64_t rows, int64_t keep, uint64_t seed, void *workspace, size_t workspace_bytes, uint8_t *mask,
                      void *stream) {
    if (!workspace || !mask || rows < 0 || keep < 0) return CNRMA_ERR_ARG;
    if (rows >= ((int64_t)1 << 27)) return CNRMA_ERR_UNSUPPORTED;   // the boundary bin (rows / 65536 keys on average) must fit its 8192-key buffer
    if (workspace_bytes < sample_workspace_bytes()) return CNRMA_ERR_CAPACITY;
    if (reinterpret_cast<uintptr_t>(workspace) % 8 != 0) return CNRMA_ERR_LAYOUT;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_sample_mask(rows, keep, seed, workspace, mask, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_sample_mask_for_result(const cnrma_rma_result *result, int64_t capacity, int64_t keep, uint64_t seed,
                                 void *workspace, size_t workspace_bytes, uint8_t *mask, void *stream) {
    if (!result || !workspace || !mask || capacity < 0 || keep < 0) return CNRMA_ERR_ARG;
    if (capacity >= ((int64_t)1 << 27)) return CNRMA_ERR_UNSUPPORTED;
    if (workspace_bytes < sample_workspace_bytes()) return CNRMA_ERR_CAPACITY;
    if (reinterpret_cast<uintptr_t>(workspace) % 8 != 0) return CNRMA_ERR_LAYOUT;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    static_assert(sizeof(long long) == sizeof(int64_t), "rows is read as long long");
    const cudaError_t e = run_sample_mask(capacity, keep, seed, workspace, mask, static_cast<cudaStream_t>(stream),
                                          reinterpret_cast<const long long *>(&result->rows));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_quantize_workspace_bytes(int64_t rows, size_t *bytes) {
    if (!bytes || rows < 0) return CNRMA_ERR_ARG;
    if (rows >= ((int64_t)1 << 30)) return CNRMA_ERR_UNSUPPORTED;
    *bytes = quantize_workspace_bytes(rows);
    return CNRMA_OK;
}

int cnrma_quantize_mark(const float *rows, int64_t row_stride, int64_t n_rows, float voxel_size, void *workspace,
                        size_t workspace_bytes, uint8_t *keep, void *stream) {
    if (!rows || !workspace || !keep || row_stride < 3 || n_rows < 0 || !(voxel_size > 0.0f)) return CNRMA_ERR_ARG;
    if (n_rows >= ((int64_t)1 << 30)) return CNRMA_ERR_UNSUPPORTED;
    if (workspace_bytes < quantize_workspace_bytes(n_rows)) return CNRMA_ERR_CAPACITY;
    if (reinterpret_cast<uintptr_t>(workspace) % 8 != 0) return CNRMA_ERR_LAYOUT;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_quantize_mark(rows, row_stride, n_rows, voxel_size, workspace, keep,
                                            static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_quantize_compact(const float *rows, int64_t row_stride, int cols, int64_t n_rows, float voxel_size,
                           const uint8_t *keep, const int32_t *prefix, float *out, int64_t out_stride, int32_t *cells,
                           int64_t capacity, void *stream) {
    if (!rows || !keep || !prefix || !out || cols < 3 || row_stride < cols || out_stride < cols || n_rows < 0 ||
        capacity < 0 || !(voxel_size > 0.0f))
        return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_quantize_compact(rows, row_stride, cols, n_rows, voxel_size, keep, prefix, out, out_stride,
                                               cells, capacity, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_tsdf_integrate(const cnrma_grid *grid, const float *projections, int64_t proj_frame_stride, int frames,
                         const float *const *depth_ptrs_host, const float *const *color_ptrs_host,
                         const int64_t *const *label_ptrs_host, int height, int width, float trunc_margin, float *tsdf,
                         float *weight, float *color, int64_t *label, void *stream) {
    if (!grid_ok(grid) || !projections || !depth_ptrs_host || !tsdf || !weight || frames <= 0 || height <= 0 ||
        width <= 0 || !(trunc_margin > 0.0f))
        return CNRMA_ERR_ARG;
    for (int f = 0; f < frames; ++f)
        if (!depth_ptrs_host[f]) return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_tsdf_integrate(to_dev(*grid), projections, proj_frame_stride, frames, depth_ptrs_host,
                                             color_ptrs_host, label_ptrs_host, height, width, trunc_margin, tsdf, weight,
                                             color, label, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

static bool head_dims_ok(int channels, int nx, int ny, int nz, int64_t stride_c, int64_t stride_v, const float *prev) {
    if (channels <= 0 || channels > 1024 || nx <= 0 || ny <= 0 || nz <= 0 || stride_c <= 0 || stride_v <= 0) return false;
    if ((int64_t)nx * ny * nz >= ((int64_t)1 << 31)) return false;
    if (prev && ((nx | ny | nz) & 1)) return false;   // F.interpolate(scale_factor=2) doubles every extent (ah.py:45)
    return true;
}

int cnrma_tsdf_head_scale(const void *x, int dtype, int channels, int nx, int ny, int nz, int64_t stride_c,
                          int64_t stride_v, const float *weight, const float *prev, float label_smoothing,
                          float sparse_threshold, float *tsdf, uint8_t *mask, void *stream) {
    if (!x || !weight || !tsdf || (dtype != CNRMA_F32 && dtype != CNRMA_BF16)) return CNRMA_ERR_ARG;
    if (!head_dims_ok(channels, nx, ny, nz, stride_c, stride_v, prev)) return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_tsdf_head_scale(x, dtype, channels, nx, ny, nz, stride_c, stride_v, weight, prev,
                                              label_smoothing, sparse_threshold, tsdf, mask,
                                              static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

size_t cnrma_tsdf_head_workspace_bytes(int channels) { return channels > 0 ? tsdf_head_workspace_bytes(channels) : 0; }

int cnrma_tsdf_head_scale_backward(const float *x, int channels, int nx, int ny, int nz, int64_t stride_c,
                                   int64_t stride_v, const float *weight, const float *prev, const float *tsdf,
                                   const float *grad_tsdf, float label_smoothing, float sparse_threshold, float *grad_x,
                                   float *grad_weight, void *workspace, size_t workspace_bytes, void *stream) {
    if (!x || !weight || !tsdf || !grad_tsdf || (!grad_x && !grad_weight)) return CNRMA_ERR_ARG;
    if (!head_dims_ok(channels, nx, ny, nz, stride_c, stride_v, prev) || !(label_smoothing != 0.0f)) return CNRMA_ERR_ARG;
    if (grad_weight && (!workspace || workspace_bytes < tsdf_head_workspace_bytes(channels))) return CNRMA_ERR_CAPACITY;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const cudaError_t e = run_tsdf_head_backward(x, channels, nx, ny, nz, stride_c, stride_v, weight, prev, tsdf,
                                                 grad_tsdf, label_smoothing, sparse_threshold, grad_x, grad_weight,
                                                 workspace, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

int cnrma_rma_expand(int views, int height, int width, int grids, float threshold, const void *workspace,
                     float *weights, uint8_t *keep, void *stream) {
    if (!workspace || views <= 0 || height <= 0 || width <= 0 || grids <= 0) return CNRMA_ERR_ARG;
    const int d = device_ok();
    if (d != CNRMA_OK) return d;
    const RmaWorkspace ws = rma_workspace(views, height, width, grids, CNRMA_MARCH_NEUS, threshold, 0);
    const cudaError_t e = run_expand(ws.rays, grids, workspace, ws, weights, keep, static_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? CNRMA_OK : fail_cuda(e);
}

}  // extern "C"
