/*
 * cnrma_b200.h -- C ABI of libcnrma_b200.so: CN-RMA's ray-marching aggregation (the 2D -> 3D feature
 * lift) as hand-written sm_100a CUDA kernels.
 *
 * The reference (SerCharles/CN-RMA) is pure PyTorch and has no FFI; the interface each entry point
 * replaces is therefore a Python function of projects/mvsdetection/models/ray_marching.py ("rm.py").
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - The caller owns all memory (PyTorch in practice); the library never allocates device memory and
 *     keeps no global state.  Scratch space is sized by the *_workspace_bytes queries.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises except
 *     where stated.  Safe to call concurrently from several host threads on different streams.
 *   - Return value: CNRMA_OK (0) or a negative cnrma_status; CUDA runtime errors are returned as
 *     CNRMA_ERR_CUDA and the cudaError_t is available from cnrma_last_cuda_error().
 *   - There is no CPU path.  A device that is not compute capability 10.x gets CNRMA_ERR_DEVICE.
 *   - Voxel order is the reference's: flat = (x*ny + y)*nz + z (datasets/tsdf.py:24-29).
 */
#ifndef CNRMA_B200_H
#define CNRMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNRMA_ABI_VERSION 2 /* 2: box / exchange entry points, mask_rows of cnrma_rma_fill_selected, cnrma_sample_mask_for_result */

typedef enum cnrma_status {
    CNRMA_OK = 0,
    CNRMA_ERR_ARG = -1,       /* null pointer, non-positive dimension, unknown enum value */
    CNRMA_ERR_LAYOUT = -2,    /* feature maps not channels-last / not 16-byte aligned for the vector path */
    CNRMA_ERR_CAPACITY = -3,  /* caller-provided buffer or workspace too small */
    CNRMA_ERR_DEVICE = -4,    /* current device is not sm_100-class */
    CNRMA_ERR_CUDA = -5,      /* a CUDA runtime call failed: see cnrma_last_cuda_error() */
    CNRMA_ERR_UNSUPPORTED = -6
} cnrma_status;

typedef enum cnrma_dtype { CNRMA_F32 = 0, CNRMA_BF16 = 1 } cnrma_dtype;

/* The voxel volume: RayMarching.voxel_dim / voxel_size / origin (rm.py:166-184). */
typedef struct cnrma_grid {
    int32_t nx, ny, nz;
    float voxel_size;
    float origin[3];
} cnrma_grid;

/* A box of the volume: voxels [lo, lo + dim) on every axis (cnrma_aggregate_views_box, cnrma_mark_rows). */
typedef struct cnrma_box {
    int32_t lo[3];
    int32_t dim[3];
} cnrma_box;

/* A stack of per-view feature maps (the `features` argument of rm.py:21 / :220 / :260 / :687).
 * view_ptrs_host is a HOST array of `views` DEVICE pointers, one per view's [C,H,W] map, so views that
 * live in separate tensors need no concatenation.  Strides are in elements.  The gather kernels
 * require channels-last maps: stride_c == 1, stride_x % 4 == 0 (f32) or % 8 == 0 (bf16), 16-byte
 * aligned bases; use cnrma_to_channels_last() first for NCHW inputs. */
typedef struct cnrma_features {
    int32_t views, channels, height, width;
    int32_t dtype; /* cnrma_dtype */
    int64_t stride_c, stride_y, stride_x;
    const void *const *view_ptrs_host;
} cnrma_features;

/* Flags of cnrma_aggregate_views. */
#define CNRMA_AGG_ACCUMULATE 1u /* add to the sums/counts already in volume/count (rm.py:243-244) */
#define CNRMA_AGG_MEAN 2u       /* finish with volume/count, 0 where count==0 (rm.py:247-257) */
#define CNRMA_AGG_COUNT_F32 4u  /* `count` holds float32 (exact below 2^24): lets sums and counts share one
                                   fp32 all-reduce buffer when views are sharded across GPUs */

/* Ray-march flavours (RayMarching.ray_marching_type, rm.py:188-194). */
typedef enum cnrma_march_mode { CNRMA_MARCH_NEUS = 0, CNRMA_MARCH_DEPTH = 1 } cnrma_march_mode;

/* Result block written by cnrma_rma_march (device memory, 32 bytes). */
typedef struct cnrma_rma_result {
    int64_t rows;     /* M: kept samples over all views */
    double weight_sum; /* sum of their weights */
    float mean;       /* (float)(weight_sum / rows), the divisor of rm.py:303 */
    int32_t overflow; /* rays whose kept-sample count exceeded the per-ray record capacity (must be 0) */
} cnrma_rma_result;

int cnrma_abi_version(void);
const char *cnrma_status_string(int status);
int cnrma_last_cuda_error(void);
/* CNRMA_OK when the current CUDA device can run the kernels (compute capability 10.x). */
int cnrma_check_device(void);
/* The tuning / test knobs (CNRMA_* environment variables) are read once, at the first call that needs one;
 * this re-reads them (for tests that flip a knob between calls; not for use while other threads are launching). */
void cnrma_reload_tuning(void);

/* ---------------------------------------------------------------------------------------------
 * Stage A -- dense back-projection (voxel driven)
 * ------------------------------------------------------------------------------------------- */

/* Replaces the index/mask part of backproject(), rm.py:47-58, for V views at once.
 *   projections  [V] 3x4 row-major matrices, UN-scaled; element stride between views = proj_view_stride
 *   stride       RayMarching.backbone2d_stride: rows 0-1 are divided by it first (rm.py:238-239)
 *   px, py       int32 [V, nvox]: round(x/z), round(y/z); 0 where !valid
 *   valid        uint8 [V, nvox]: px>=0 & py>=0 & px<W & py<H & pz>0
 * Any of px / py / valid may be NULL.  Bit-exact with the reference on CPU. */
int cnrma_project_views(const cnrma_grid *grid, const float *projections, int64_t proj_view_stride, int views,
                        float stride, int height, int width, int32_t *px, int32_t *py, uint8_t *valid,
                        void *stream);

/* Replaces  V x aggregate_2d_features (rm.py:220-244; backproject rm.py:21-69 inside it)  and, with
 * CNRMA_AGG_MEAN, clear_3d_features (rm.py:247-257) -- one fused pass: project, mask, nearest gather,
 * sum over views in view order (fp32, so bit-exact with the reference's running sum), divide.
 *   volume  f32, element (voxel, c) at volume[voxel*vol_stride_voxel + c*vol_stride_channel]
 *           ([nvox,C] channels-last: (C,1);  the reference's NCDHW: (1,nvox))
 *   count   int32 [nvox] number of views that see the voxel (self.valid before clear_3d_features)
 *   valid   uint8 [nvox] count > 0 (self.valid after clear_3d_features); may be NULL
 * With V = 1 and flags = 0 this is backproject() itself.  features->views may be 0 together with
 * CNRMA_AGG_ACCUMULATE | CNRMA_AGG_MEAN: a finalise-only pass over sums and counts produced elsewhere
 * (e.g. after an all-reduce of view-sharded partial results). */
int cnrma_aggregate_views(const cnrma_grid *grid, const cnrma_features *features, const float *projections,
                          int64_t proj_view_stride, float stride, uint32_t flags, float *volume,
                          int64_t vol_stride_voxel, int64_t vol_stride_channel, int32_t *count, uint8_t *valid,
                          void *stream);

/* cnrma_aggregate_views restricted to a box of the grid -- the unit of work of the voxel-sharded multi-GPU mode
 * (every rank owns one box and lifts ALL views into it) and of chunk-pipelined collectives.
 * World coordinates are formed from the global voxel indices, so every voxel gets the bits the full-grid call gives
 * it.  volume / count / valid are the BOX's buffers: voxel (x, y, z) of the box at ((x*dim[1] + y)*dim[2] + z).
 *   reserve_ctas  number of this kernel's CTA slots to leave free for a kernel that runs beside it (the row puller
 *                 below); 0 = take the whole GPU. */
int cnrma_aggregate_views_box(const cnrma_grid *grid, const cnrma_box *box, const cnrma_features *features,
                              const float *projections, int64_t proj_view_stride, float stride, uint32_t flags,
                              float *volume, int64_t vol_stride_voxel, int64_t vol_stride_channel, int32_t *count,
                              uint8_t *valid, int reserve_ctas, void *stream);

/* Voxel-sharded Stage A, the exchange step: a rank that owns `box` needs, from the views other ranks hold, only the
 * feature rows its voxels project to.
 *   cnrma_mark_rows  sets bit (pixel % 32) of bitmap[view * words + pixel / 32], words = ceil(H*W / 32), for every
 *                    (view, pixel) some voxel of the box gathers (same projection arithmetic as the gather kernels,
 *                    so the set is exact).  The caller zeroes `bitmap` first; bits are only ever set.
 *                    parts > 1 (<= 16, <= dim[0]): the box is cut into `parts` x-ranges at dim[0]*k/parts and range k
 *                    marks its own bitmap at bitmap + k*part_stride (uint32 units) -- one launch for a box that is
 *                    served part by part.
 *   cnrma_pull_rows  copies the marked rows of `views` channels-last maps from their owners (src_view_ptrs_host[v]: HOST
 *                    array of DEVICE pointers, PEER-MAPPED in the multi-GPU use: the reads cross NVLink) to the same
 *                    offsets of `dst` (local staging, view v at dst + v*dst_view_stride), as TMA bulk copies global ->
 *                    shared -> global in a two-stage pipeline per warp.  row_bytes = C * sizeof(element), a multiple
 *                    of 16, <= 8192; up to 512 views per call.
 *                    `work` (may be NULL): a zeroed device counter; warps then claim bitmap words one at a time, which
 *                    balances sparse bitmaps (without it the words are dealt out round-robin).
 *                    `first_view`: the walk over the views starts there and wraps around -- ranks that pull from the
 *                    same owners start at different ones, so that no owner's NVLink egress serves every reader at once.
 *                    `done` (same shape as bitmap, may be NULL): rows whose bit is set there are already in dst and
 *                    are skipped; the rows this launch pulls are added to it -- lets a box be served part by part
 *                    (the gather of one part runs while the next part's rows arrive) without pulling a row twice.
 *                    ctas = 0 picks the default grid (cnrma_pull_default_ctas); OR CNRMA_PULL_LSU into it to move the
 *                    rows with 16-byte loads / stores through registers instead of bulk copies -- the better path
 *                    when the puller shares the SMs with the gather kernel, whose own bulk copies fill the TMA queues. */
#define CNRMA_PULL_LSU 0x10000
int cnrma_mark_rows(const cnrma_grid *grid, const cnrma_box *box, const float *projections, int64_t proj_view_stride,
                    int views, float stride, int height, int width, uint32_t *bitmap, int parts, int64_t part_stride,
                    void *stream);
int cnrma_pull_rows(const uint32_t *bitmap, uint32_t *done, int views, int height, int width, int row_bytes,
                    const void *const *src_view_ptrs_host, void *dst, int64_t dst_view_stride, int ctas, uint32_t *work,
                    int first_view, void *stream);
int cnrma_pull_default_ctas(void);

/* OPT-IN extra, not a replacement of any reference function: Stage A with bilinear instead of nearest sampling
 * (BASELINE.json's north_star names a bilinear gather; the reference samples nearest and parity with it wins, so
 * every other entry point samples nearest).  Same validity mask and view count as cnrma_aggregate_views; value =
 * grid_sample(bilinear, border, align_corners=True) at (cx/cz, cy/cz), summed in view order (/ count with
 * CNRMA_AGG_MEAN).  volume f32 [nvox, C] channels-last contiguous; up to 512 views.  Runs in the same TMA gather
 * machinery as cnrma_aggregate_views (four bulk copies per visible view). */
int cnrma_aggregate_views_bilinear(const cnrma_grid *grid, const cnrma_features *features, const float *projections,
                                   int64_t proj_view_stride, float stride, uint32_t flags, float *volume, int32_t *count,
                                   uint8_t *valid, void *stream);

/* View-sharded Stage A over peer memory ("by view subset"): instead of an all-reduce of full-size partial
 * volumes, the voxels are split into n_owners contiguous ranges of slab_voxels voxels, and the kernel stores voxel v's
 * un-normalised sums (C floats) and view count (one float at [C]) into row v % slab_voxels of owner v / slab_voxels:
 *   owner_rows_host[o]  HOST array of DEVICE pointers: THIS source's section [slab_voxels][row_floats] of owner o's
 *                       buffer -- a peer-mapped pointer when o is another GPU (the stores then cross NVLink by
 *                       themselves, overlapped with the gather); row_floats >= C + 1, multiple of 4, 16-byte aligned rows
 * After all sources are done (a device barrier between the ranks, the caller's business), every owner runs
 * cnrma_finalize_routed over its buffer recv [n_src][slab_voxels][row_floats]: partial sums and counts are added in
 * source order and divided (mean != 0) like rm.py:251; volume f32 [rows, C], count int32 [rows], valid uint8 or NULL.
 * Indices, masks and counts are exact; the sums are regrouped by source, hence 1e-5. */
int cnrma_aggregate_views_routed(const cnrma_grid *grid, const cnrma_features *features, const float *projections,
                                 int64_t proj_view_stride, float stride, int n_owners, int slab_voxels, int row_floats,
                                 float *const *owner_rows_host, void *stream);
int cnrma_finalize_routed(const float *recv, int n_src, int slab_voxels, int row_floats, int rows, int channels, int mean,
                          float *volume, int32_t *count, uint8_t *valid, void *stream);

/* Proof obligation of the fused mean (rm.py:251 `volume / valid`): the kernel divides by the view count with
 * one correctly rounded reciprocal and a Markstein correction instead of a generic division.  This entry
 * point checks that shortcut against IEEE division for every count in [1, max_n] and every fp32 significand
 * (both signs, three binades) and writes the number of mismatches -- which must be 0 -- to *mismatches
 * (device uint64). */
int cnrma_selftest_count_division(int max_n, uint64_t *mismatches, void *stream);

/* Layout helper for reference-layout (NCHW) feature maps: src [V][C,H,W] with the given strides ->
 * dst channels-last [V,H,W,C] contiguous, same dtype.  One read + one write of the features. */
int cnrma_to_channels_last(const cnrma_features *src, void *dst, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Stage B -- ray marching aggregation (ray driven)
 * ------------------------------------------------------------------------------------------- */

/* t_one of rm.py:710-711 evaluated like the reference (python doubles, then one rounding to float). */
float cnrma_t_one(const cnrma_grid *grid, double voxel_size, int grids);

/* Replaces get_ray_parameter (rm.py:71-111) as tensors: o, d f32 [V,3,H*W] from pinv [V] 4x4 (see
 * cnrma_rma_march for how pinv is obtained).  The march / fill kernels recompute the same values inline. */
int cnrma_ray_parameters(const float *pinv, int views, int height, int width, float *o, float *d, void *stream);

/* Bytes of scratch cnrma_rma_march / _fill / _scatter need for `views` x H x W rays. */
int cnrma_rma_workspace_bytes(const cnrma_grid *grid, int views, int height, int width, int grids, int mode,
                              float threshold, int depth_points, size_t *bytes);

/* Replaces get_ray_parameter (rm.py:71-111; the 4x4 inverses are an input, see below) and the march /
 * weight / selection part of ray_projection_neus (rm.py:710-767) or ray_projection_depth
 * (rm.py:826-911) for all V views (batch 1, like the reference, rm.py:707).
 *   pinv       [V] 4x4 row-major inverse of [P_scaled; 0 0 0 1] (rm.py:96-102).  The host computes it with
 *              the same LAPACK call as the reference (torch.inverse) so ray parameters stay bit-exact.
 *   tsdf       f32 [nx,ny,nz]
 *   grids      N, samples per ray (rm.py:687 default 300);  t_one from cnrma_t_one()
 *   threshold  weight_threshold / neus_threshold (NEUS);  depth_points = select_grids (DEPTH)
 *   result     device cnrma_rma_result: rows, weight sum, mean.  Read result->rows back (the one sync
 *              of the path; the reference syncs once per view at rm.py:781-782) to size the output.
 * Sample order is the reference's: ascending (view, v, u, step). */
int cnrma_rma_march(const cnrma_grid *grid, const float *pinv, int views, int height, int width,
                    const float *tsdf, int grids, float t_one, int mode, float threshold, int depth_points,
                    void *workspace, size_t workspace_bytes, cnrma_rma_result *result, void *stream);

/* Replaces the compaction + feature gather of rm.py:769-807 and, with normalize != 0, the concat +
 * weight normalisation of aggregate_2d_features_ray_marching (rm.py:289-307).
 *   rows        f32 [M, row_stride]; columns  normalize ? [x,y,z, feat*w/mean] : [x,y,z,w, feat]
 *   mean        device float used as the divisor when normalize != 0; NULL = result->mean of the march
 *               (a view-sharded caller passes the all-reduced mean here)
 *   capacity    rows the buffer can hold; rows beyond it are not written
 *   rows_host   M as read back from result->rows: CNRMA_ERR_CAPACITY if capacity < rows_host.  Pass -1 to launch
 *               speculatively BEFORE reading M (the kernel takes its offsets from the workspace, not from M): the
 *               caller then compares result->rows with capacity afterwards and repeats the call if it did not fit
 * grids / mode / threshold / depth_points must repeat the march call's values (they fix the workspace
 * layout).  Features need stride_c == 1; no alignment requirement. */
int cnrma_rma_fill(const cnrma_grid *grid, const float *pinv, const cnrma_features *features, int grids,
                   float t_one, int mode, float threshold, int depth_points, const void *workspace,
                   const cnrma_rma_result *result, int64_t rows_host, int normalize, const float *mean,
                   float *rows, int64_t row_stride, int64_t capacity, void *stream);

/* Derived dense operator (north_star's output form; SURVEY.md section 8a "dense_rma"): instead of
 * emitting rows, adds w*feat into wsum[voxel, :] and w into wtot[voxel] for the voxel each kept sample
 * rounds to (rm.py:730).  wsum f32 [nvox, C] channels-last, wtot f32 [nvox]; both are accumulated into
 * (zero them first; partial results from view shards add).  fp32 atomics: order-dependent rounding. */
int cnrma_rma_scatter(const cnrma_grid *grid, const float *pinv, const cnrma_features *features, int grids,
                      float t_one, int mode, float threshold, int depth_points, const void *workspace, float *wsum,
                      float *wtot, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Backward of the two feature gathers (what autograd does for rm.py:61-64 / :243 / :251 and rm.py:799 / :304 when
 * the detector trains, rm.py:409-451).  Gradients are fp32, channels-last, described like feature maps.
 * ------------------------------------------------------------------------------------------- */

/* d features of cnrma_aggregate_views: grad_features[view][py][px][:] += grad_volume[voxel][:] (/ count[voxel]
 * with CNRMA_AGG_MEAN) for every view that sees the voxel.  grad_features must be zeroed by the caller (the adds are
 * fp32 reductions performed by the memory system; order-dependent rounding like index_put_(accumulate=True)).
 * `count` is the forward's int32 view count. */
int cnrma_aggregate_views_backward(const cnrma_grid *grid, const cnrma_features *grad_features, const float *projections,
                                   int64_t proj_view_stride, float stride, uint32_t flags, const float *grad_volume,
                                   int64_t vol_stride_voxel, int64_t vol_stride_channel, const int32_t *count,
                                   void *stream);

/* d features of cnrma_rma_fill: grad_features[view][v][u][:] = sum over the rows of that ray of
 * grad_rows[row][3 + :] * w/mean (normalize) or grad_rows[row][4 + :] (un-normalised rows).  Weights and positions
 * carry no gradient (rm.py:705).  Every pixel row is written (zeros where the ray kept nothing); needs the
 * workspace / result of the forward march. */
int cnrma_rma_fill_backward(const cnrma_grid *grid, const cnrma_features *grad_features, int grids, int mode,
                            float threshold, int depth_points, const void *workspace, const cnrma_rma_result *result,
                            int normalize, const float *mean, const float *grad_rows, int64_t row_stride, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Point-cloud hand-off (RayMarching.switch_pointcloud, rm.py:339-407): `coord + offset` and the sub-sampling mask of
 * sample_points (datasets/pipelines/fcaf3d_transforms.py:283-296, max_points) applied to whole rows in one ordered
 * compaction instead of one torch.masked_select per column (rm.py:380-402).  The mask itself is an input: the
 * reference draws it with numpy's RNG on the host, and only the same mask gives the same rows.
 * ------------------------------------------------------------------------------------------- */

/* Bytes of scratch for cnrma_mask_prefix over `rows` rows. */
int cnrma_handoff_workspace_bytes(int64_t rows, size_t *bytes);

/* prefix[i] = number of kept rows before row i (int32 [rows]); *kept (device int64) = total kept. */
int cnrma_mask_prefix(const uint8_t *mask, int64_t rows, void *workspace, size_t workspace_bytes, int32_t *prefix,
                      int64_t *kept, void *stream);

/* out[prefix[i], :] = rows[i, :] for every kept row i, with offset_host[0..2] added to columns 0..2.
 * rows f32 [n_rows, row_stride] (first `cols` columns used), out f32 [capacity, out_stride]. */
int cnrma_select_rows(const float *rows, int64_t row_stride, int cols, int64_t n_rows, const uint8_t *mask,
                      const int32_t *prefix, const float *offset_host, float *out, int64_t out_stride,
                      int64_t capacity, void *stream);

/* On-device alternative to sample_points' numpy draw (an option, never the parity path: it is equivalent in
 * distribution, not the same random stream): mask[i] = 1 for exactly min(keep, rows) rows chosen uniformly at random
 * (keyed by `seed`, deterministic), 0 elsewhere.  Removes the ~0.2 s host-side np.random.choice per scene. */
int cnrma_sample_workspace_bytes(size_t *bytes);
int cnrma_sample_mask(int64_t rows, int64_t keep, uint64_t seed, void *workspace, size_t workspace_bytes, uint8_t *mask,
                      void *stream);
/* The same draw with the row count taken from the march's result block IN DEVICE MEMORY (result->rows, clamped to
 * `capacity`, the length of `mask`), so that the hand-off can be queued behind the march without reading M back first:
 * mask[i] = 1 for exactly min(keep, rows) of the first `rows` entries, 0 for every other i < capacity.  For equal
 * (rows, keep, seed) the mask equals cnrma_sample_mask's. */
int cnrma_sample_mask_for_result(const cnrma_rma_result *result, int64_t capacity, int64_t keep, uint64_t seed,
                                 void *workspace, size_t workspace_bytes, uint8_t *mask, void *stream);

/* The last step of the hand-off (rm.py:330-332): `coords / voxel_size_fcaf3d` goes into MinkowskiEngine 0.5.4
 * (ME.utils.batch_sparse_collate + ME.SparseTensor; third-party, not part of the reference tree), whose collate step
 * truncates the quotient to int32 and whose sparse tensor keeps one row per occupied cell.  These two calls do that on
 * the device, deterministically: the FIRST row of every cell (in row order) survives, survivors stay in row order.
 *   cnrma_quantize_mark     keep[i] = 1 iff row i is the first row of its cell trunc(coord / voxel_size); rows whose cell
 *                           index leaves +-2^20 on an axis (or is NaN) are dropped.  workspace: cnrma_quantize_workspace_bytes.
 *   cnrma_quantize_compact  with prefix = exclusive prefix sum of keep (cnrma_mask_prefix): surviving rows copied whole to
 *                           out[prefix[i]], their cells to cells[prefix[i]] (int32 [K,3]; may be NULL). */
int cnrma_quantize_workspace_bytes(int64_t rows, size_t *bytes);
int cnrma_quantize_mark(const float *rows, int64_t row_stride, int64_t n_rows, float voxel_size, void *workspace,
                        size_t workspace_bytes, uint8_t *keep, void *stream);
int cnrma_quantize_compact(const float *rows, int64_t row_stride, int cols, int64_t n_rows, float voxel_size,
                           const uint8_t *keep, const int32_t *prefix, float *out, int64_t out_stride, int32_t *cells,
                           int64_t capacity, void *stream);

/* cnrma_rma_fill fused with the hand-off: only the kept rows are produced (at out row prefix[row], offset added),
 * i.e. aggregate_2d_features_ray_marching + switch_pointcloud in one pass.  mask / prefix index the M rows of the
 * march in their (view, v, u, step) order and hold `mask_rows` entries: rows beyond them are dropped (0: they cover
 * every row), which lets the launch be queued with buffers sized from a guess of M before M has been read back. */
int cnrma_rma_fill_selected(const cnrma_grid *grid, const float *pinv, const cnrma_features *features, int grids,
                            float t_one, int mode, float threshold, int depth_points, const void *workspace,
                            const cnrma_rma_result *result, int normalize, const float *mean, const uint8_t *mask,
                            const int32_t *prefix, int64_t mask_rows, const float *offset_host, float *rows,
                            int64_t row_stride, int64_t capacity, void *stream);

/* ---------------------------------------------------------------------------------------------
 * GT TSDF fusion (offline data preparation): TSDFFusion.integrate, data_prepare/scannet/tsdf.py:402-451 (same code in
 * data_prepare/arkit/tsdf.py), for a batch of frames in one pass over the volume; frames are applied in order, so
 * the volumes are bit-identical to calling the reference once per frame.
 *   projections   [frames] 3x4 row-major (full resolution: no stride scaling), element stride proj_frame_stride
 *   depth / color / label pointer tables: HOST arrays of DEVICE pointers: depth f32 [H,W] (0 = no reading),
 *                 color f32 [3,H,W] (table or entries may be NULL), label i64 [H,W] (table or entries may be NULL)
 *   trunc_margin  voxel_size * trunc_ratio (tsdf.py:373)
 *   tsdf, weight  f32 [nvox] state (tsdf starts at 1, weight at 0: tsdf.py:379-380); color f32 [3,nvox] or NULL;
 *                 label i64 [nvox] or NULL (starts at -1)
 * ------------------------------------------------------------------------------------------- */
int cnrma_tsdf_integrate(const cnrma_grid *grid, const float *projections, int64_t proj_frame_stride, int frames,
                         const float *const *depth_ptrs_host, const float *const *color_ptrs_host,
                         const int64_t *const *label_ptrs_host, int height, int width, float trunc_margin, float *tsdf,
                         float *weight, float *color, int64_t *label, void *stream);

/* ---------------------------------------------------------------------------------------------
 * TSDF head hand-over: one scale of AtlasTSDFHead.forward, projects/mvsdetection/models/atlas_head.py:38-52 -- the
 * producer of `scene_tsdf_004`, the volume cnrma_rma_march walks.
 *   tsdf = tanh(sum_c weight[c] * x[c, v]) * label_smoothing; with prev (the coarser scale's output, extents
 *   nx/2 x ny/2 x nz/2, nearest x2 upsampling): where !(|prev| < sparse_threshold): tsdf = sign(prev) * .999
 *   x       f32 or bf16, element strides stride_c (channels) / stride_v (flat voxel index (x*ny+y)*nz+z):
 *           NCDHW -> (nvox, 1); channels-last -> (1, C).  Voxels whose coarse parent is not surface are not read.
 *   weight  f32 [C] device (the Conv3d(C,1,1,bias=False) kernel, ah.py:29);  prev NULL for the first scale
 *   tsdf    f32 [nvox] out;  mask uint8 [nvox] out or NULL (= |prev| < thr, ah.py:47; all ones without prev)
 * Floating-point row (the reference's cudnn reduction order is unspecified): parity 1e-5, mask exact outside a
 * 1e-5 band around the threshold.
 * cnrma_tsdf_head_scale_backward: autograd of the above for dL/dtsdf = grad_tsdf (f32 x only): grad_x (x's strides)
 * and/or grad_weight [C] may be NULL; workspace (cnrma_tsdf_head_workspace_bytes) is needed for grad_weight.
 * ------------------------------------------------------------------------------------------- */
int cnrma_tsdf_head_scale(const void *x, int dtype, int channels, int nx, int ny, int nz, int64_t stride_c,
                          int64_t stride_v, const float *weight, const float *prev, float label_smoothing,
                          float sparse_threshold, float *tsdf, uint8_t *mask, void *stream);
size_t cnrma_tsdf_head_workspace_bytes(int channels);
int cnrma_tsdf_head_scale_backward(const float *x, int channels, int nx, int ny, int nz, int64_t stride_c,
                                   int64_t stride_v, const float *weight, const float *prev, const float *tsdf,
                                   const float *grad_tsdf, float label_smoothing, float sparse_threshold, float *grad_x,
                                   float *grad_weight, void *workspace, size_t workspace_bytes, void *stream);

/* Dense per-sample view of the march records for parity tests: weights f32 [V*H*W*N] (0 where not kept,
 * i.e. rm.py:767 `weights * valid_final`) and keep uint8 [V*H*W*N].  NEUS mode only. */
int cnrma_rma_expand(int views, int height, int width, int grids, float threshold, const void *workspace,
                     float *weights, uint8_t *keep, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CNRMA_B200_H */
