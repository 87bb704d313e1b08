// Stage B: ray-marching aggregation -- per-pixel rays marched through the TSDF, NeuS opacity /
// transmittance weights (or the depth-crossing variant), ordered compaction, feature copy and weight
// normalisation (reference: rm.py:71-111, :260-307, :687-956).
//
// Three launches per scene (DESIGN.md "K_B"):
//   march  one thread per ray; writes the ray's kept samples as (step, weight) records, its count, and a
//          per-CTA row total and weight sum.  Exact early exits: transmittance below the threshold, and
//          leaving the (convex) grid after having been inside it.
//   scan   one CTA: exclusive scan of the per-CTA totals -> row offsets, M, sum(w), mean(w).
//   fill   one CTA per 256 rays: per-ray offsets by warp scan, then each warp streams its rays' rows --
//          the pixel's feature vector is read once into registers and written once per kept sample,
//          scaled by w/mean, as coalesced row segments.
#include <cstdlib>

#include "cnrma_internal.cuh"

namespace cnrma {

// ---- workspace layout ------------------------------------------------------------------------------
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int rma_record_capacity(int grids, int mode, float threshold, int depth_points) {
    if (mode == CNRMA_MARCH_DEPTH) return depth_points > 0 ? 2 * depth_points : 1;
    // kept NeuS weights are >= threshold and sum to at most 1 (+ rounding), so a ray keeps at most 1/threshold
    if (threshold > 0.0f) {
        const double k = 1.0 / (double)threshold + 2.0;
        return k < (double)grids ? (int)k : grids;
    }
    return grids;
}

RmaWorkspace rma_workspace(int views, int height, int width, int grids, int mode, float threshold, int depth_points,
                           int64_t nvox) {
    RmaWorkspace w;
    w.rays = (int64_t)views * height * width;
    w.blocks = (w.rays + kRayThreads - 1) / kRayThreads;
    w.cap = rma_record_capacity(grids, mode, threshold, depth_points);
    size_t o = 0;
    w.off_counts = o;   o = align256(o + sizeof(int32_t) * (size_t)w.rays);
    w.off_blk_rows = o; o = align256(o + sizeof(int32_t) * (size_t)w.blocks);
    w.off_blk_wsum = o; o = align256(o + sizeof(double) * (size_t)w.blocks);
    w.off_blk_off = o;  o = align256(o + sizeof(int64_t) * (size_t)w.blocks);
    w.off_rec_w = o;    o = align256(o + sizeof(float) * (size_t)w.rays * (size_t)w.cap);
    w.off_rec_i = o;    o = align256(o + sizeof(float) * (size_t)w.rays * (size_t)w.cap);
    w.off_dist = o;     o = align256(o + 2 * (size_t)nvox);   // two uint8 distance fields (ping-pong)
    w.off_sigmoid = o;  o = align256(o + sizeof(float) * (size_t)nvox);
    w.off_cell = o;     o = align256(o + sizeof(uint2) * (size_t)nvox);   // {bits of s, clearance} per voxel
    w.total = o;
    return w;
}

struct MarchParams {
    GridDev g;
    const float *pinv;   // [V][16]
    const float *tsdf;
    int V, H, W, N;
    float t_one, thr;
    int depth_points;
    int cap;
    int64_t rays;
    int32_t *counts;
    int32_t *blk_rows;
    double *blk_wsum;
    float *rec_w, *rec_i;
    cnrma_rma_result *result;
    const uint8_t *dist;     // [nvox] Chebyshev distance (capped) to the nearest voxel of another sigmoid value
    const float *sig;        // [nvox] sigmoid(-tsdf), written by tsdf_prepare_kernel
    const uint2 *cell;       // [nvox] {bits of sig, dist}: what a skipping march reads per sample, in one 8-byte load
};

__device__ __forceinline__ float sigmoid_neg(float tv) {
    // torch.sigmoid(-tsdf) = 1 / (1 + exp(tsdf)) (rm.py:757); the correctly rounded reciprocal is the IEEE quotient
    return __frcp_rn(__fadd_rn(1.0f, expf(tv)));
}

// Per-ray constants of the position -> voxel id mapping.
struct VoxelMap {
    float inv_vs;      // RN(1 / voxel_size)
    bool zero_origin;  // origin == (0,0,0): `places - origin` is the identity and is skipped
};

// round((p - origin) / voxel_size) per coordinate (rm.py:730: sub, true division, half-to-even).
// The quotient is only needed rounded to an integer, so the IEEE division is replaced by a multiplication with
// the rounded reciprocal whenever that provably cannot change the result: q_fast = RN(rel * RN(1/vs)) is
// within |q| * 2^-23 of the exact quotient, hence rint(q_fast) == rint(RN(rel / vs)) unless q_fast lies within
// |q| * 2^-22 of a half-integer -- in which case (about one sample in 10^4) the division is done for real.
//
// One sample of a ray (rm.py:729-733): position -> rounded voxel id; returns the flat id or -1 outside the grid
// (ix, iy, iz are the integer coordinates when inside).  Callers guarantee finite o, d (non-finite rays keep
// nothing in the reference either: their ids convert to INT64_MIN and fail the bounds test).
// `frac` (optional) receives the sample's offset from the centre of its voxel, in voxels, per axis (|frac| <= 0.5 up to
// the 2^-22 relative error of q).
__device__ __forceinline__ int sample_voxel(const GridDev &g, const VoxelMap &m, const float o[3], const float d[3],
                                            float t, int &ix, int &iy, int &iz, float *frac = nullptr) {
    const float org[3] = {g.ox, g.oy, g.oz};
    float rel[3], r[3], q[3];
    bool ambiguous = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float p = __fadd_rn(o[a], __fmul_rn(d[a], t));
        rel[a] = m.zero_origin ? p : __fsub_rn(p, org[a]);
        q[a] = __fmul_rn(rel[a], m.inv_vs);
        r[a] = rintf(q[a]);
        const float safe = __fmaf_rn(fabsf(q[a]), -2.384185791015625e-07f, 0.5f);   // 0.5 - |q| * 2^-22
        ambiguous = ambiguous || !(fabsf(__fsub_rn(q[a], r[a])) < safe);             // also true for huge q
    }
    if (ambiguous) {
#pragma unroll
        for (int a = 0; a < 3; ++a) r[a] = rintf(__fdiv_rn(rel[a], g.vs));
    }
    if (frac != nullptr) {
#pragma unroll
        for (int a = 0; a < 3; ++a) frac[a] = __fsub_rn(q[a], r[a]);
    }
    // integer-valued floats; __float2int_rn saturates, so far-away samples simply fail the unsigned range test
    ix = __float2int_rn(r[0]);
    iy = __float2int_rn(r[1]);
    iz = __float2int_rn(r[2]);
    const bool inb = ((unsigned)ix < (unsigned)g.nx) && ((unsigned)iy < (unsigned)g.ny) && ((unsigned)iz < (unsigned)g.nz);
    return inb ? ((ix * g.ny + iy) * g.nz + iz) : -1;
}

__device__ __forceinline__ int sample_voxel(const GridDev &g, const VoxelMap &m, const float o[3], const float d[3],
                                            float t) {
    int ix, iy, iz;
    return sample_voxel(g, m, o, d, t, ix, iy, iz);
}

__device__ __forceinline__ bool ray_is_finite(const float o[3], const float d[3]) {
    return isfinite(o[0] + o[1] + o[2]) && isfinite(d[0] + d[1] + d[2]);
}

// ---- empty-space skipping ----------------------------------------------------------------------------------
// The TSDF the reference marches through is piecewise constant over large regions (free space and unobserved
// space are set to +-0.999 by the coarse-to-fine head, atlas_head.py:47).  Consecutive samples that read the
// same value have alpha == 0 exactly: they change neither the transmittance nor (for thr > 0) the kept set.
// A pre-pass tabulates s = sigmoid(-tsdf) per voxel and a capped Chebyshev distance field D: D(c) = k means every
// voxel within L-infinity distance k of c holds the same s as c (voxels next to a different value, and the grid
// border, have D = 0).  A sample that rounds to voxel c sits within 0.5 voxel of its centre, so the next
// n = floor((k - margin) / max_a |step_a|) samples stay within k + 0.5 - margin of the centre, round to voxels
// inside that cube and read the same s: the march jumps over them.  ("step" is the per-sample advance in voxel
// units; margin = 0.05 voxel dwarfs the ~1e-5 voxel fp32 error of o + d*t.)  Where a sample advances well under a
// voxel the march uses the sharper per-axis form: sample i + j sits at frac_a + j * step_a from the centre of the
// voxel sample i rounded to (frac = its offset, |frac| <= 0.5), and reads the same s while
// |frac_a + j * step_a| < k + 0.5 - margin on every axis -- which also covers k = 0: the further samples inside
// one and the same voxel.  The march reads s and k with one 8-byte load from a combined table.
constexpr int kDistCap = 15;
constexpr float kSkipMargin = 0.05f;

__global__ void __launch_bounds__(256) tsdf_sigmoid_kernel(const float *__restrict__ tsdf, int nvox,
                                                           float *__restrict__ sig) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nvox) sig[v] = sigmoid_neg(__ldg(tsdf + v));
}

// Pass 1 (along z, the contiguous axis): boundary test + 1-D distance.  A voxel is a boundary voxel when it lies on
// the grid border or any of its 26 neighbours holds a different s.
__device__ __forceinline__ bool is_boundary(const GridDev &g, const float *__restrict__ sig, int x, int y, int z) {
    if (x == 0 || y == 0 || z == 0 || x == g.nx - 1 || y == g.ny - 1 || z == g.nz - 1) return true;
    const uint32_t ref = __float_as_uint(__ldg(sig + (x * g.ny + y) * g.nz + z));
    bool differs = false;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz)
                differs = differs || (__float_as_uint(__ldg(sig + ((x + dx) * g.ny + (y + dy)) * g.nz + (z + dz))) != ref);
    return differs;
}

__global__ void __launch_bounds__(256) dist_boundary_kernel(GridDev g, const float *__restrict__ sig,
                                                            uint8_t *__restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.nx * g.ny * g.nz) return;
    const int z = v % g.nz, xy = v / g.nz, y = xy % g.ny, x = xy / g.ny;
    out[v] = is_boundary(g, sig, x, y, z) ? 0 : kDistCap;
}

// Separable L-infinity distance transform: out(v) = min over |d| <= cap along `axis` of max(|d|, in(v + d)).
// The last pass also writes the march's combined table when `cell` is given.
__global__ void __launch_bounds__(256) dist_pass_kernel(GridDev g, int axis, const uint8_t *__restrict__ in,
                                                        uint8_t *__restrict__ out, const float *__restrict__ sig = nullptr,
                                                        uint2 *__restrict__ cell = nullptr) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.nx * g.ny * g.nz) return;
    const int z = v % g.nz, xy = v / g.nz, y = xy % g.ny, x = xy / g.ny;
    const int pos = (axis == 0) ? x : ((axis == 1) ? y : z);
    const int n = (axis == 0) ? g.nx : ((axis == 1) ? g.ny : g.nz);
    const int stride = (axis == 0) ? g.ny * g.nz : ((axis == 1) ? g.nz : 1);
    int best = in[v];
    for (int d = 1; d < best; ++d) {   // a neighbour at distance d can only help while d < best
        if (pos - d >= 0) best = min(best, max(d, (int)in[v - d * stride]));
        if (pos + d < n) best = min(best, max(d, (int)in[v + d * stride]));
    }
    out[v] = (uint8_t)best;
    if (cell != nullptr) cell[v] = make_uint2(__float_as_uint(sig[v]), (unsigned)best);
}

// Fused pre-pass for grids whose (y, z) plane fits shared memory: one CTA per x index tabulates s for its plane,
// runs the boundary test (on the raw TSDF bits: equal bits give equal s; unequal bits with equal s only make the
// field more conservative) and the z and y passes of the distance transform in shared memory.  With the x pass that
// follows, the pre-pass is two launches instead of five (the march of cfg 2 is ~280 us; every launch costs 5-7 us).
constexpr int kSlabThreads = 1024;
__global__ void __launch_bounds__(kSlabThreads) tsdf_prepare_slab_kernel(GridDev g, const float *__restrict__ tsdf,
                                                                         float *__restrict__ sig, uint8_t *__restrict__ out,
                                                                         cnrma_rma_result *result) {
    extern __shared__ uint8_t slab[];
    const int plane = g.ny * g.nz;
    uint8_t *A = slab, *B = slab + plane;
    const int x = blockIdx.x;
    if (x == 0 && threadIdx.x == 0) {   // the march that follows adds to `overflow`; the scan kernel writes the rest
        result->rows = 0;
        result->weight_sum = 0.0;
        result->mean = 0.0f;
        result->overflow = 0;
    }
    const float *base = tsdf + (int64_t)x * plane;
    for (int i = threadIdx.x; i < plane; i += kSlabThreads) {
        const int y = i / g.nz, z = i - y * g.nz;
        const float tv = __ldg(base + i);
        sig[(int64_t)x * plane + i] = sigmoid_neg(tv);
        if (out != nullptr) {
            bool boundary = (x == 0 || y == 0 || z == 0 || x == g.nx - 1 || y == g.ny - 1 || z == g.nz - 1);
            if (!boundary) {
                const uint32_t ref = __float_as_uint(tv);
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dz = -1; dz <= 1; ++dz)
                            boundary = boundary || (__float_as_uint(__ldg(base + dx * plane + (y + dy) * g.nz + (z + dz))) != ref);
            }
            A[i] = boundary ? 0 : kDistCap;
        }
    }
    if (out == nullptr) return;
    __syncthreads();
    for (int i = threadIdx.x; i < plane; i += kSlabThreads) {   // along z
        const int y = i / g.nz, z = i - y * g.nz;
        int best = A[i];
        for (int d = 1; d < best; ++d) {
            if (z - d >= 0) best = min(best, max(d, (int)A[i - d]));
            if (z + d < g.nz) best = min(best, max(d, (int)A[i + d]));
        }
        B[i] = (uint8_t)best;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plane; i += kSlabThreads) {   // along y
        const int y = i / g.nz;
        int best = B[i];
        for (int d = 1; d < best; ++d) {
            if (y - d >= 0) best = min(best, max(d, (int)B[i - d * g.nz]));
            if (y + d < g.ny) best = min(best, max(d, (int)B[i + d * g.nz]));
        }
        out[(int64_t)x * plane + i] = (uint8_t)best;
    }
}

__device__ __forceinline__ VoxelMap make_voxel_map(const GridDev &g) {
    VoxelMap m;
    m.inv_vs = __frcp_rn(g.vs);
    m.zero_origin = (g.ox == 0.0f) && (g.oy == 0.0f) && (g.oz == 0.0f);
    return m;
}


// Per-CTA totals: kept rows (the fill kernel works on the same blocks of kRayThreads consecutive rays) and the sum
// of the kept weights (double, fixed shape -> deterministic).
// (Mapping CTAs to 32x8-pixel tiles with 8x4-pixel warps instead of 256 consecutive rays was measured: no gain,
// 336 -> 350 us on cfg 2 plus an extra block-count kernel, so rays stay in flat order.  Per-warp totals without the CTA
// barrier -- ncu shows 16 % of the march's warp samples parked there, waiting for the block's longest ray -- were
// measured too: the CTA keeps its slot until its last warp is done either way, and the single-CTA scan then walks eight
// times as many partials: march phase 0.334 -> 0.375 ms on cfg 2, 0.504 -> 0.550 ms on the reference test grid.)
template <int BLOCK>
__device__ __forceinline__ void block_totals(int kept, double wsum, int32_t *blk_rows, double *blk_wsum) {
    __shared__ int s_rows[BLOCK / kWarp];
    __shared__ double s_w[BLOCK / kWarp];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kept += __shfl_down_sync(0xffffffffu, kept, o);
        wsum += __shfl_down_sync(0xffffffffu, wsum, o);
    }
    if (lane == 0) {
        s_rows[warp] = kept;
        s_w[warp] = wsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int r = 0;
        double w = 0.0;
        for (int i = 0; i < BLOCK / kWarp; ++i) {
            r += s_rows[i];
            w += s_w[i];
        }
        blk_rows[blockIdx.x] = r;
        blk_wsum[blockIdx.x] = w;
    }
}

// NeuS march (rm.py:710-767).  Streaming form of
//   s_i = sigmoid(-tsdf_i); alpha_i = max((s_i - s_{i+1}) / s_i, 0) with s_N := s_{N-1};
//   T_i = prod_{j<i} (1 - alpha_j) (sequential product, like torch.cumprod on CPU); w_i = T_i * alpha_i;
//   keep_i = in-bounds_i & (w_i >= thr).
// Consecutive samples that read the same TSDF value (same voxel, or neighbouring voxels of equal value such as
// free space) have alpha == 0 exactly: the transmittance is unchanged and, for thr > 0, nothing is kept, so the
// exp / divisions are only evaluated where the TSDF value changes.
// SKIP: thr > 0 and the clearance table exists (the normal case); one 8-byte load per sample brings s and the clearance.
// FINE: the jump also uses where the sample sits inside its voxel (see the loop) -- three more multiply-adds per
// sample that pay when a sample advances well under a voxel (cfg 2: 0.39 voxel per sample, march 0.316 -> 0.267 ms) and
// cost a little when it advances more than one (the reference's 0.04 m grid: 1.25 voxels per sample, 0.465 -> 0.472 ms);
// the launcher picks.  Registers capped at 40 (6 CTAs per SM): the FINE form would take 47, measured equal or slower.
template <bool SKIP, bool FINE>
__global__ void __launch_bounds__(kRayThreads, 6) march_neus_kernel(const __grid_constant__ MarchParams p) {
    const int64_t ray = (int64_t)blockIdx.x * kRayThreads + threadIdx.x;
    int kept = 0;
    double wsum = 0.0;
    if (ray < p.rays) {
        const int hw = p.H * p.W;
        const int view = (int)(ray / hw);
        const int pix = (int)(ray % hw);
        float o[3], d[3];
        ray_of_pixel(p.pinv + 16 * view, pix % p.W, pix / p.W, o, d);
        const VoxelMap vm = make_voxel_map(p.g);
        const bool keep_zero = SKIP ? false : !(p.thr > 0.0f);   // thr <= 0 (or NaN): zero weights pass `w >= thr` too

        // FINE: samples per voxel of travel along each axis, signed: 1 / (d_a * t_one / vs).  (An axis the ray does not
        // move along gets a huge finite value: it never limits the jump.)  Otherwise one figure for the fastest axis.
        float inv_step[3];
        if (FINE) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float step = d[a] * p.t_one / p.g.vs;
                inv_step[a] = copysignf(1.0f / fmaxf(fabsf(step), 1e-6f), step);
            }
        } else {
            const float step_max = fmaxf(fmaxf(fabsf(d[0]), fabsf(d[1])), fabsf(d[2])) * p.t_one / p.g.vs;
            inv_step[0] = inv_step[1] = inv_step[2] = (step_max > 0.0f) ? 1.0f / step_max : 0.0f;
        }
        // non-finite rays keep nothing (see sample_voxel).  (The bound is p.N itself, a kernel constant: a per-ray bound
        // was being re-derived from spilled values at every sample under the 40-register cap.)
        const int n_steps = p.N;
        int overflow = 0;
        if (ray_is_finite(o, d)) {
            float T = 1.0f;
            float s_cur = 0.0f;   // sigmoid(-tsdf) of the current sample (set at i == 0)
            const float s_out = sigmoid_neg(1.0f);   // samples outside the grid read tsdf = 1.0 (rm.py:744)
            int vox_cur = -1;
            bool entered = false;
            for (int i = 0; i <= n_steps; ++i) {
                int vox_next = -1;
                int skip_to = -1;
                float s_next = s_cur;   // i == N: last sample repeated (rm.py:758); same voxel -> same value
                if (i < p.N) {
                    // the clearance is re-read at every sample: keeping it in a register while the voxel is unchanged was
                    // measured and is slower (cfg 2 march phase 0.350 -> 0.383 ms, cfg 1 0.244 -> 0.264 ms: the load hits L1 and
                    // the extra live register / select costs more than it saves)
                    if (SKIP && FINE) {
                        int ix, iy, iz;
                        float frac[3];
                        vox_next = sample_voxel(p.g, vm, o, d, __fmul_rn((float)i, p.t_one), ix, iy, iz, frac);
                        s_next = s_out;
                        if (vox_next >= 0) {
                            const uint2 c = __ldg(p.cell + vox_next);
                            s_next = __uint_as_float(c.x);
                            // Clearance k: every voxel within k of this one (L-infinity) holds the same s.  Sample i + j sits at
                            // frac_a + j * step_a from this voxel's centre; while that stays below k + 0.5 - margin on every axis
                            // it rounds to a voxel inside that cube (for k = 0: to this very voxel) and reads the same s.
                            const float room = (float)c.y + (0.5f - kSkipMargin);
                            float n = __fmaf_rn(-frac[0], inv_step[0], room * fabsf(inv_step[0]));
                            n = fminf(n, __fmaf_rn(-frac[1], inv_step[1], room * fabsf(inv_step[1])));
                            n = fminf(n, __fmaf_rn(-frac[2], inv_step[2], room * fabsf(inv_step[2])));
                            skip_to = i + (int)n;   // n < 1 (or negative): no jump
                        }
                    } else if (SKIP) {
                        vox_next = sample_voxel(p.g, vm, o, d, __fmul_rn((float)i, p.t_one));
                        s_next = s_out;
                        if (vox_next >= 0) {
                            const uint2 c = __ldg(p.cell + vox_next);
                            s_next = __uint_as_float(c.x);
                            // wherever the sample sits in its voxel (at most 0.5 from the centre), the next
                            // floor((k - margin) / step) samples stay inside the cube of clearance k
                            if (c.y > 0) skip_to = i + (int)(((float)c.y - kSkipMargin) * inv_step[0]);
                        }
                    } else {
                        vox_next = sample_voxel(p.g, vm, o, d, __fmul_rn((float)i, p.t_one));
                        if (i == 0 || vox_next != vox_cur) s_next = (vox_next >= 0) ? __ldg(p.sig + vox_next) : s_out;
                    }
                }
                if (i > 0 && (s_next != s_cur || keep_zero)) {
                    float a = __fdiv_rn(__fsub_rn(s_cur, s_next), s_cur);
                    a = (a < 0.0f) ? 0.0f : a;   // clamp(min=0), NaN-propagating like torch
                    const float w = __fmul_rn(T, a);
                    if (vox_cur >= 0 && w >= p.thr) {
                        if (kept < p.cap) {
                            p.rec_w[(int64_t)kept * p.rays + ray] = w;
                            p.rec_i[(int64_t)kept * p.rays + ray] = (float)(i - 1);
                            wsum += (double)w;
                            ++kept;
                        } else {
                            overflow = 1;
                        }
                    }
                    T = __fmul_rn(T, __fsub_rn(1.0f, a));
                    // exact early exit: every later weight is <= T < thr
                    if ((SKIP || p.thr > 0.0f) && T < p.thr) break;
                }
                // exact early exit: the grid is convex and the rounded sample ids are monotone along the ray, so once
                // left it is never re-entered, and samples outside are never kept
                if (entered && vox_next < 0) break;
                entered = entered || (vox_next >= 0);
                s_cur = s_next;
                vox_cur = vox_next;
                // samples i+1 .. skip_to all round to voxels that hold s_cur, so alpha == 0 and nothing
                // changes; resume with sample skip_to + 1 (never past the repeated last sample)
                if (skip_to > i) i = min(skip_to, p.N - 1);
            }
        }
        p.counts[ray] = kept;
        if (overflow) atomicAdd(&p.result->overflow, 1);
    }
    block_totals<kRayThreads>(kept, wsum, p.blk_rows, p.blk_wsum);
}

// Depth-crossing march (rm.py:826-911): first step i with tsdf_i * tsdf_{i+1} <= 0, then either the 2k
// neighbouring steps with triangular weights (k = depth_points > 0) or the half-step point (k == 0).
__global__ void __launch_bounds__(kRayThreads) march_depth_kernel(const __grid_constant__ MarchParams p) {
    const int64_t ray = (int64_t)blockIdx.x * kRayThreads + threadIdx.x;
    int kept = 0;
    double wsum = 0.0;
    if (ray < p.rays) {
        const int hw = p.H * p.W;
        const int view = (int)(ray / hw);
        const int pix = (int)(ray % hw);
        float o[3], d[3];
        ray_of_pixel(p.pinv + 16 * view, pix % p.W, pix / p.W, o, d);
        const VoxelMap vm = make_voxel_map(p.g);
        int best = -1;
        float tv_cur = 1.0f;
        int vox_cur = -1;
        bool entered = false;
        const int n_steps = ray_is_finite(o, d) ? p.N : 0;   // non-finite rays: every sample reads 1.0, no crossing
        for (int i = 0; i < n_steps; ++i) {
            const int vox = sample_voxel(p.g, vm, o, d, __fmul_rn((float)i, p.t_one));
            const float tv = (i > 0 && vox == vox_cur) ? tv_cur : ((vox >= 0) ? __ldg(p.tsdf + vox) : 1.0f);
            if (i > 0 && __fmul_rn(tv_cur, tv) <= 0.0f) {
                best = i - 1;
                break;
            }
            // exact early exit: the grid is convex and the rounded sample ids are monotone along the ray, so a ray that
            // has left it reads 1.0 from here on: 1.0 * 1.0 > 0, no crossing can follow (the step out was tested above)
            if (entered && vox < 0) break;
            entered = entered || (vox >= 0);
            tv_cur = tv;
            vox_cur = vox;
        }
        if (best >= 0) {
            const int k = p.depth_points;
            if (k > 0) {
                for (int j = 0; j < 2 * k; ++j) {
                    const int idx = best + j - k + 1;
                    const int tri = (j < k) ? (j + 1) : (2 * k - j);
                    const float w = __fdiv_rn((float)tri, (float)k);   // multi_weight.float() / select_grids
                    if (idx < 0 || idx >= p.N || !(w > 0.0f)) continue;
                    p.rec_w[(int64_t)kept * p.rays + ray] = w;
                    p.rec_i[(int64_t)kept * p.rays + ray] = (float)idx;
                    wsum += (double)w;
                    ++kept;
                }
            } else {
                p.rec_w[ray] = 1.0f;
                p.rec_i[ray] = __fadd_rn((float)best, 0.5f);
                wsum = 1.0;
                kept = 1;
            }
        }
        p.counts[ray] = kept;
    }
    block_totals<kRayThreads>(kept, wsum, p.blk_rows, p.blk_wsum);
}

// Exclusive scan of the per-CTA row totals; fixed-shape reduction of the weight sums (deterministic).
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) scan_blocks_kernel(const int32_t *__restrict__ blk_rows,
                                                                   const double *__restrict__ blk_wsum,
                                                                   int64_t *__restrict__ blk_off, int64_t blocks,
                                                                   cnrma_rma_result *result) {
    __shared__ int64_t s_rows[kScanThreads];
    __shared__ double s_w[kScanThreads];
    const int t = threadIdx.x;
    const int64_t per = (blocks + kScanThreads - 1) / kScanThreads;
    const int64_t lo = (int64_t)t * per;
    const int64_t hi = (lo + per < blocks) ? lo + per : blocks;
    int64_t rows = 0;
    double w = 0.0;
    for (int64_t i = lo; i < hi; ++i) {
        rows += blk_rows[i];
        w += blk_wsum[i];
    }
    s_rows[t] = rows;
    s_w[t] = w;
    __syncthreads();
    // Hillis-Steele inclusive scan of the strip totals (rows) and a pairwise tree for the weights
    for (int o = 1; o < kScanThreads; o <<= 1) {
        const int64_t add = (t >= o) ? s_rows[t - o] : 0;
        __syncthreads();
        s_rows[t] += add;
        __syncthreads();
    }
    for (int o = kScanThreads / 2; o > 0; o >>= 1) {
        if (t < o) s_w[t] += s_w[t + o];
        __syncthreads();
    }
    int64_t run = s_rows[t] - rows;   // exclusive prefix of this strip
    for (int64_t i = lo; i < hi; ++i) {
        blk_off[i] = run;
        run += blk_rows[i];
    }
    if (t == kScanThreads - 1) {
        result->rows = s_rows[t];
        result->weight_sum = s_w[0];
        result->mean = (s_rows[t] > 0) ? (float)(s_w[0] / (double)s_rows[t]) : 0.0f;
    }
}

// ---- fill ------------------------------------------------------------------------------------------
struct FillParams {
    GridDev g;
    const float *pinv;
    int V, C, H, W;
    int64_t stride_y, stride_x;
    float t_one;
    int mode;          // cnrma_march_mode: selects how a record's step becomes a position
    int normalize;     // 1: [x,y,z, feat*w/mean]   0: [x,y,z,w, feat]
    int64_t rays;
    const int32_t *counts;
    const int64_t *blk_off;
    const float *rec_w, *rec_i;
    const float *mean;
    float *rows;
    int64_t row_stride;
    int64_t capacity;     // rows the output buffer holds: rows beyond it are dropped (speculative launches)
    const uint8_t *sel_mask;     // optional hand-off selection: only rows with sel_mask[row] != 0 are written,
    const int32_t *sel_prefix;   // at row index sel_prefix[row], with sel_off added to x, y, z (rm.py:365, :380-402)
    int64_t sel_rows;            // entries of sel_mask / sel_prefix: rows beyond them are dropped (a mask sized from a guess of M)
    float sel_off[3];
    float *wsum, *wtot;   // scatter variant
    const void *views[kMaxViewsPerLaunch];
    int view_base;        // first view of this launch
    int stage_half;       // packed kernel: bytes per staging half (two per warp)
};

__device__ __forceinline__ void record_position(const FillParams &p, const float o[3], const float d[3], float fi,
                                                float pos[3]) {
    if (p.mode == CNRMA_MARCH_NEUS) {
        const float t = __fmul_rn(fi, p.t_one);   // places = o + d * (i * t_one), rm.py:715,729
#pragma unroll
        for (int r = 0; r < 3; ++r) pos[r] = __fadd_rn(o[r], __fmul_rn(d[r], t));
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r)   // selected_places = o + (d * index) * t_one, rm.py:901,910
            pos[r] = __fadd_rn(o[r], __fmul_rn(__fmul_rn(d[r], fi), p.t_one));
    }
}

template <typename T>
__device__ __forceinline__ float load_feat(const T *p);
template <>
__device__ __forceinline__ float load_feat<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_feat<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return __uint_as_float(((uint32_t)__ldg(reinterpret_cast<const unsigned short *>(p))) << 16);
}

constexpr int kFillRegs = 8;   // channels per lane per pass in the fill kernel (8*32 = 256 channels)

// SCATTER = false: write rows.  SCATTER = true: atomically add w*feat into wsum[voxel] and w into wtot[voxel].
template <typename T, bool SCATTER>
__global__ void __launch_bounds__(kRayThreads) fill_rows_kernel(const __grid_constant__ FillParams p) {
    __shared__ int s_warp_rows[kRayThreads / kWarp];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hw = p.H * p.W;
    // this launch covers views [view_base, view_base + V): rays are numbered globally
    const int64_t ray0 = (int64_t)p.view_base * hw + (int64_t)blockIdx.x * kRayThreads;
    const int64_t my_ray = ray0 + threadIdx.x;
    const int64_t ray_end = (int64_t)(p.view_base + p.V) * hw;
    const int my_cnt = (my_ray < ray_end && my_ray < p.rays) ? p.counts[my_ray] : 0;
    // exclusive offsets: warp scan + per-warp totals
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp_rows[warp] = incl;
    __syncthreads();
    int64_t base = p.blk_off[ray0 / kRayThreads];
    for (int i = 0; i < warp; ++i) base += s_warp_rows[i];
    const int64_t my_off = base + (incl - my_cnt);
    const float mean = (p.normalize && !SCATTER) ? __ldg(p.mean) : 1.0f;
    const int col0 = SCATTER ? 0 : (p.normalize ? 3 : 4);

    // lane <-> ray: the ray of my pixel and the address of its feature vector, shuffled to the warp when the ray's turn
    // comes (evaluating them warp-wide per ray costs 32x the instructions; with a hand-off selection most rays keep
    // little or nothing and that setup was the bulk of the kernel)
    float mo[3] = {0.0f, 0.0f, 0.0f}, md[3] = {0.0f, 0.0f, 0.0f};
    const T *my_feat = nullptr;
    if (my_cnt > 0) {
        const int view = (int)(my_ray / hw);
        const int pix = (int)(my_ray % hw);
        const int u = pix % p.W, v = pix / p.W;
        ray_of_pixel(p.pinv + 16 * view, u, v, mo, md);
        my_feat = static_cast<const T *>(p.views[view - p.view_base]) + (int64_t)v * p.stride_y + (int64_t)u * p.stride_x;
    }

    for (int r = 0; r < 32; ++r) {
        const int cnt = __shfl_sync(0xffffffffu, my_cnt, r);
        if (cnt == 0) continue;
        const int64_t off = __shfl_sync(0xffffffffu, my_off, r);
        const int64_t ray = ray0 + warp * 32 + r;
        float o[3], d[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            o[a] = __shfl_sync(0xffffffffu, mo[a], r);
            d[a] = __shfl_sync(0xffffffffu, md[a], r);
        }
        const T *feat = reinterpret_cast<const T *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(my_feat), r));
        // records in chunks of 32: lane k prepares record k0 + k (position, scaled weight, target voxel) and writes
        // the row's leading columns; the feature columns are then streamed row by row with the weight broadcast
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int nk = min(32, cnt - k0);
            float wk = 0.0f;       // weight that multiplies the features of this lane's record
            int64_t vox = -1;      // scatter target
            int64_t dst_row = -1;  // output row of this lane's record (-1: not written)
            if (lane < nk) {
                const int64_t row_id = off + k0 + lane;
                if (!SCATTER) {
                    dst_row = row_id;
                    if (p.sel_mask != nullptr)
                        dst_row = (row_id < p.sel_rows && __ldg(p.sel_mask + row_id)) ? (int64_t)__ldg(p.sel_prefix + row_id) : -1;
                    if (dst_row >= p.capacity) dst_row = -1;
                }
                if (SCATTER || dst_row >= 0) {
                    const float w = __ldg(p.rec_w + (int64_t)(k0 + lane) * p.rays + ray);
                    const float fi = __ldg(p.rec_i + (int64_t)(k0 + lane) * p.rays + ray);
                    float pos[3];
                    record_position(p, o, d, fi, pos);
                    if (!SCATTER) {
                        float *row = p.rows + dst_row * p.row_stride;
                        if (p.sel_mask != nullptr) {   // coord + offsets[b]: one more rounding (rm.py:365)
                            pos[0] = __fadd_rn(pos[0], p.sel_off[0]);
                            pos[1] = __fadd_rn(pos[1], p.sel_off[1]);
                            pos[2] = __fadd_rn(pos[2], p.sel_off[2]);
                        }
                        row[0] = pos[0];
                        row[1] = pos[1];
                        row[2] = pos[2];
                        if (p.normalize) {
                            wk = __fdiv_rn(w, mean);   // weights / mean(weights), rm.py:303
                        } else {
                            row[3] = w;
                            wk = 1.0f;
                        }
                    } else {
                        const float qx = rintf(__fdiv_rn(__fsub_rn(pos[0], p.g.ox), p.g.vs));
                        const float qy = rintf(__fdiv_rn(__fsub_rn(pos[1], p.g.oy), p.g.vs));
                        const float qz = rintf(__fdiv_rn(__fsub_rn(pos[2], p.g.oz), p.g.vs));
                        const bool inb = (qx >= 0.0f) && (qx < (float)p.g.nx) && (qy >= 0.0f) && (qy < (float)p.g.ny) &&
                                         (qz >= 0.0f) && (qz < (float)p.g.nz);
                        if (inb) {
                            vox = ((int64_t)(int)qx * p.g.ny + (int)qy) * p.g.nz + (int)qz;
                            atomicAdd(p.wtot + vox, w);
                        }
                        wk = w;
                    }
                }
            }
            if (!SCATTER && !__any_sync(0xffffffffu, dst_row >= 0)) continue;   // nothing of this ray is kept
            for (int cbase = 0; cbase < p.C; cbase += 32 * kFillRegs) {
                float f[kFillRegs];
#pragma unroll
                for (int j = 0; j < kFillRegs; ++j) {
                    const int c = cbase + j * 32 + lane;
                    f[j] = (c < p.C) ? load_feat<T>(feat + c) : 0.0f;
                }
                for (int k = 0; k < nk; ++k) {
                    const float wn = __shfl_sync(0xffffffffu, wk, k);
                    if (!SCATTER) {
                        const int64_t drow = __shfl_sync(0xffffffffu, dst_row, k);
                        if (drow < 0) continue;
                        float *dst = p.rows + drow * p.row_stride + col0;
#pragma unroll
                        for (int j = 0; j < kFillRegs; ++j) {
                            const int c = cbase + j * 32 + lane;
                            if (c < p.C) __stcs(dst + c, p.normalize ? __fmul_rn(f[j], wn) : f[j]);
                        }
                    } else {
                        const int64_t tv = __shfl_sync(0xffffffffu, vox, k);
                        if (tv < 0) continue;
#pragma unroll
                        for (int j = 0; j < kFillRegs; ++j) {
                            const int c = cbase + j * 32 + lane;
                            if (c < p.C) atomicAdd(p.wsum + tv * p.C + c, __fmul_rn(wn, f[j]));
                        }
                    }
                }
            }
        }
    }
}

// ---- fill, TMA-store variant ---------------------------------------------------------------------------
// Rows are 4*(3+C) bytes -- 1036 for C = 256 -- so consecutive rows are only 4-byte aligned and plain stores
// reach ~4.8 TB/s (profiles/microbench/store_paths.cu).  The rows of one ray are contiguous in the output, so
// each warp assembles a batch of them in a shared-memory staging buffer that mirrors the global byte range
// (same offset modulo 16), pushes the 16-byte aligned interior with ONE bulk asynchronous store (TMA,
// cp.async.bulk.global.shared) and writes the few unaligned floats at either end itself: ~6.3 TB/s.
constexpr int kStageBytes = 8192;   // per-warp staging buffer; 3 CTAs x 8 warps x 8 KB = 192 KB per SM

template <typename T, int NJ>   // NJ = ceil(C / 32) register groups per lane (1, 2, 4 or 8)
__global__ void __launch_bounds__(kRayThreads) fill_rows_tma_kernel(const __grid_constant__ FillParams p) {
    extern __shared__ __align__(128) unsigned char stage_raw[];
    __shared__ int s_warp_rows[kRayThreads / kWarp];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *stage = reinterpret_cast<float *>(stage_raw + (size_t)warp * kStageBytes);
    const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(stage);
    const int hw = p.H * p.W;
    const int64_t ray0 = (int64_t)p.view_base * hw + (int64_t)blockIdx.x * kRayThreads;
    const int64_t my_ray = ray0 + threadIdx.x;
    const int64_t ray_end = (int64_t)(p.view_base + p.V) * hw;
    const int my_cnt = (my_ray < ray_end && my_ray < p.rays) ? p.counts[my_ray] : 0;
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp_rows[warp] = incl;
    __syncthreads();
    int64_t base = p.blk_off[ray0 / kRayThreads];
    for (int i = 0; i < warp; ++i) base += s_warp_rows[i];
    const int64_t my_off = base + (incl - my_cnt);
    const float mean = p.normalize ? __ldg(p.mean) : 1.0f;
    const int col0 = p.normalize ? 3 : 4;
    const int cols = p.C + col0;
    const int row_bytes = cols * 4;
    const int rows_per_batch = (kStageBytes - 16) / row_bytes;

    for (int r = 0; r < 32; ++r) {
        const int cnt = __shfl_sync(0xffffffffu, my_cnt, r);
        if (cnt == 0) continue;
        const int64_t off = __shfl_sync(0xffffffffu, my_off, r);
        const int64_t ray = ray0 + warp * 32 + r;
        const int view = (int)(ray / hw);
        const int pix = (int)(ray % hw);
        const int u = pix % p.W, v = pix / p.W;
        float o[3], d[3];
        ray_of_pixel(p.pinv + 16 * view, u, v, o, d);
        const T *feat = static_cast<const T *>(p.views[view - p.view_base]) + (int64_t)v * p.stride_y + (int64_t)u * p.stride_x;
        float f[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int c = j * 32 + lane;
            f[j] = (c < p.C) ? load_feat<T>(feat + c) : 0.0f;
        }
        for (int k0 = 0; k0 < cnt; k0 += 32) {
            const int nk = (int)min((int64_t)min(32, cnt - k0), max((int64_t)0, p.capacity - (off + k0)));   // rows that fit
            float wk = 0.0f, wraw = 0.0f, pos[3] = {0.0f, 0.0f, 0.0f};
            if (lane < nk) {
                wraw = __ldg(p.rec_w + (int64_t)(k0 + lane) * p.rays + ray);
                const float fi = __ldg(p.rec_i + (int64_t)(k0 + lane) * p.rays + ray);
                record_position(p, o, d, fi, pos);
                wk = p.normalize ? __fdiv_rn(wraw, mean) : 1.0f;   // weights / mean(weights), rm.py:303
            }
            for (int kb = 0; kb < nk; kb += rows_per_batch) {
                const int nb = min(rows_per_batch, nk - kb);
                unsigned char *gbase = reinterpret_cast<unsigned char *>(p.rows) + (off + k0 + kb) * (int64_t)row_bytes;
                const int h = (int)(reinterpret_cast<uintptr_t>(gbase) & 15);   // same offset modulo 16 in the buffer
                const int hw4 = h >> 2;
                // the previous bulk store must have finished reading the buffer
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                for (int k = 0; k < nb; ++k) {
                    const float wn = __shfl_sync(0xffffffffu, wk, kb + k);
                    float *row = stage + hw4 + k * cols + col0;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int c = j * 32 + lane;
                        if (c < p.C) row[c] = p.normalize ? __fmul_rn(f[j], wn) : f[j];
                    }
                }
                if (lane >= kb && lane < kb + nb) {
                    float *row = stage + hw4 + (lane - kb) * cols;
                    row[0] = pos[0];
                    row[1] = pos[1];
                    row[2] = pos[2];
                    if (!p.normalize) row[3] = wraw;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                const int total = nb * row_bytes;
                const int a0 = (h + 15) & ~15;            // first 16-byte aligned buffer offset inside the data
                const int a1 = (h + total) & ~15;         // end of the aligned interior
                if (lane == 0 && a1 > a0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gbase - h + a0),
                                 "r"(stage_addr + (uint32_t)a0), "r"((uint32_t)(a1 - a0))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                // unaligned floats in front of / behind the interior (at most 3 each)
                const int head = (min(a0, h + total) - h) >> 2;
                const int tail_lo = max(a1, a0);
                const int tail = (h + total > tail_lo) ? ((h + total - tail_lo) >> 2) : 0;
                if (lane < head) reinterpret_cast<float *>(gbase)[lane] = stage[hw4 + lane];
                if (lane < tail) reinterpret_cast<float *>(gbase - h + tail_lo)[lane] = stage[(tail_lo >> 2) + lane];
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Dense view of the records for parity tests: weights * valid_final and valid_final (rm.py:765-767).
__global__ void __launch_bounds__(256) expand_records_kernel(int64_t rays, int N, const int32_t *__restrict__ counts,
                                                             const float *__restrict__ rec_w,
                                                             const float *__restrict__ rec_i,
                                                             float *__restrict__ weights, uint8_t *__restrict__ keep) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= rays) return;
    const int cnt = counts[ray];
    for (int k = 0; k < cnt; ++k) {
        const int i = (int)rec_i[(int64_t)k * rays + ray];
        if (weights) weights[ray * N + i] = rec_w[(int64_t)k * rays + ray];
        if (keep) keep[ray * N + i] = 1;
    }
}

// get_ray_parameter as tensors (rm.py:71-111): o, d [V,3,H*W].
__global__ void __launch_bounds__(256) ray_parameters_kernel(const float *__restrict__ pinv, int V, int H, int W,
                                                             float *__restrict__ o_out, float *__restrict__ d_out) {
    const int hw = H * W;
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= (int64_t)V * hw) return;
    const int view = (int)(ray / hw), pix = (int)(ray % hw);
    float o[3], d[3];
    ray_of_pixel(pinv + 16 * view, pix % W, pix / W, o, d);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        o_out[((int64_t)view * 3 + r) * hw + pix] = o[r];
        d_out[((int64_t)view * 3 + r) * hw + pix] = d[r];
    }
}

cudaError_t run_ray_parameters(const float *pinv, int V, int H, int W, float *o, float *d, cudaStream_t stream) {
    const int64_t rays = (int64_t)V * H * W;
    ray_parameters_kernel<<<(unsigned)((rays + 255) / 256), 256, 0, stream>>>(pinv, V, H, W, o, d);
    return cudaGetLastError();
}

// ---- host-side launchers -----------------------------------------------------------------------------

cudaError_t run_march(const GridDev &g, const float *pinv, int V, int H, int W, const float *tsdf, int N, float t_one,
                      int mode, float thr, int depth_points, void *workspace, const RmaWorkspace &ws,
                      cnrma_rma_result *result, cudaStream_t stream) {
    unsigned char *base = static_cast<unsigned char *>(workspace);
    MarchParams p;
    p.g = g;
    p.pinv = pinv;
    p.tsdf = tsdf;
    p.V = V; p.H = H; p.W = W; p.N = N;
    p.t_one = t_one;
    p.thr = thr;
    p.depth_points = depth_points;
    p.cap = ws.cap;
    p.rays = ws.rays;
    p.counts = reinterpret_cast<int32_t *>(base + ws.off_counts);
    p.blk_rows = reinterpret_cast<int32_t *>(base + ws.off_blk_rows);
    p.blk_wsum = reinterpret_cast<double *>(base + ws.off_blk_wsum);
    p.rec_w = reinterpret_cast<float *>(base + ws.off_rec_w);
    p.rec_i = reinterpret_cast<float *>(base + ws.off_rec_i);
    p.result = result;
    p.dist = nullptr;
    p.sig = nullptr;
    p.cell = nullptr;
    cudaError_t err = cudaSuccess;
    const int nvox = g.nx * g.ny * g.nz;
    const size_t slab_bytes = 2 * (size_t)g.ny * g.nz;
    const bool slab = mode == CNRMA_MARCH_NEUS && slab_bytes <= 96 * 1024 && !tuning().march_unfused;
    if (!slab) {
        err = cudaMemsetAsync(result, 0, sizeof(cnrma_rma_result), stream);
        if (err != cudaSuccess) return err;
    }
    if (mode == CNRMA_MARCH_NEUS) {
        const unsigned vb = (unsigned)((nvox + 255) / 256);
        float *sig = reinterpret_cast<float *>(base + ws.off_sigmoid);
        uint8_t *d0 = reinterpret_cast<uint8_t *>(base + ws.off_dist);
        uint8_t *d1 = d0 + nvox;
        uint2 *cell = reinterpret_cast<uint2 *>(base + ws.off_cell);
        p.sig = sig;
        if (slab) {
            static thread_local int attr_dev = -1;
            int dev = 0;
            cudaGetDevice(&dev);
            if (attr_dev != dev) {
                err = cudaFuncSetAttribute(tsdf_prepare_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
                if (err != cudaSuccess) return err;
                attr_dev = dev;
            }
            tsdf_prepare_slab_kernel<<<(unsigned)g.nx, kSlabThreads, slab_bytes, stream>>>(g, tsdf, sig, thr > 0.0f ? d0 : nullptr, result);
            if (thr > 0.0f) {
                dist_pass_kernel<<<vb, 256, 0, stream>>>(g, 0, d0, d1, sig, cell);
                p.dist = d1;
                p.cell = cell;
            }
        } else {
            tsdf_sigmoid_kernel<<<vb, 256, 0, stream>>>(tsdf, nvox, sig);
            if (thr > 0.0f) {
                dist_boundary_kernel<<<vb, 256, 0, stream>>>(g, sig, d0);
                dist_pass_kernel<<<vb, 256, 0, stream>>>(g, 2, d0, d1);
                dist_pass_kernel<<<vb, 256, 0, stream>>>(g, 1, d1, d0);
                dist_pass_kernel<<<vb, 256, 0, stream>>>(g, 0, d0, d1, sig, cell);
                p.dist = d1;
                p.cell = cell;
            }
        }
        err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    if (mode == CNRMA_MARCH_DEPTH)
        march_depth_kernel<<<(unsigned)ws.blocks, kRayThreads, 0, stream>>>(p);
    else if (p.cell != nullptr) {
        // voxels per sample along a unit direction; see the kernel's FINE note for the two measured points
        const bool fine = tuning().march_jump >= 0 ? tuning().march_jump != 0 : t_one <= 0.75f * g.vs;
        if (fine)
            march_neus_kernel<true, true><<<(unsigned)ws.blocks, kRayThreads, 0, stream>>>(p);
        else
            march_neus_kernel<true, false><<<(unsigned)ws.blocks, kRayThreads, 0, stream>>>(p);
    } else
        march_neus_kernel<false, false><<<(unsigned)ws.blocks, kRayThreads, 0, stream>>>(p);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    scan_blocks_kernel<<<1, kScanThreads, 0, stream>>>(p.blk_rows, p.blk_wsum,
                                                       reinterpret_cast<int64_t *>(base + ws.off_blk_off), ws.blocks,
                                                       result);
    return cudaGetLastError();
}

// ---- fill, short rows ------------------------------------------------------------------------------------
// The reference's own maps have 32 channels: a row is 140 bytes and a ray keeps ~6 of them, so in the kernel above the
// per-ray work (the ray of the pixel -- computed by all 32 lanes for the same ray --, one small bulk store and its
// head / tail fix-ups per ray) outweighs the rows themselves (ref test grid: 21 % of the HBM peak).  Here the per-ray
// quantities are computed lane <-> ray once per warp and shuffled, and the rows of consecutive rays -- contiguous in
// the output -- share staging buffers: two 4 KB halves per warp, filled across rays and pushed with one bulk store
// when full, so the store of one half overlaps the assembly of the other.  Same arithmetic, same bits.
constexpr int kHalfStageBytes = kStageBytes / 2;
constexpr int kPackedRowBytesMax = 600;   // rows up to this size take the packed kernel (measured: see DESIGN.md K_B3)

// SELECT: the hand-off's row selection (rm.py:380-402) fused in: only rows with sel_mask != 0 are produced, at output
// row sel_prefix[row] (an exclusive prefix sum of the mask, so kept rows stay contiguous and in order), with sel_off
// added to the coordinates (rm.py:365).
template <typename T, int NJ, bool SELECT>
__global__ void __launch_bounds__(kRayThreads) fill_rows_packed_kernel(const __grid_constant__ FillParams p) {
    extern __shared__ __align__(128) unsigned char stage_raw[];
    __shared__ int s_warp_rows[kRayThreads / kWarp];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half_bytes = p.stage_half;   // two halves per warp
    unsigned char *stage_base = stage_raw + (size_t)warp * 2 * half_bytes;
    const int hw = p.H * p.W;
    const int64_t ray0 = (int64_t)p.view_base * hw + (int64_t)blockIdx.x * kRayThreads;
    const int64_t my_ray = ray0 + threadIdx.x;
    const int64_t ray_end = (int64_t)(p.view_base + p.V) * hw;
    const int my_cnt = (my_ray < ray_end && my_ray < p.rays) ? p.counts[my_ray] : 0;
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp_rows[warp] = incl;
    __syncthreads();
    int64_t base = p.blk_off[ray0 / kRayThreads];
    for (int i = 0; i < warp; ++i) base += s_warp_rows[i];
    const int64_t my_off = base + (incl - my_cnt);
    // rows beyond the output capacity are dropped (speculative launches): clamp every ray's count once
    int my_keep = (int)min((int64_t)my_cnt, max((int64_t)0, p.capacity - my_off));
    int64_t my_out0 = my_off;                                           // output row of my ray's first produced row
    if (SELECT) {
        my_keep = 0;
        for (int k = 0; k < my_cnt; ++k) my_keep += (my_off + k < p.sel_rows && __ldg(p.sel_mask + my_off + k)) ? 1 : 0;
        my_out0 = (my_cnt > 0 && my_off < p.sel_rows) ? (int64_t)__ldg(p.sel_prefix + my_off) : 0;
    }
    const unsigned with_rows = __ballot_sync(0xffffffffu, my_cnt > 0);
    const int first_lane = with_rows ? (__ffs(with_rows) - 1) : 0;
    const int64_t warp_off = __shfl_sync(0xffffffffu, my_out0, first_lane);   // the warp's rows are contiguous from here
    unsigned char *const warp_gbase = reinterpret_cast<unsigned char *>(p.rows) + warp_off * (int64_t)((p.C + (p.normalize ? 3 : 4)) * 4);
    int rows_done = 0;                                                  // rows of this warp already staged
    const float mean = p.normalize ? __ldg(p.mean) : 1.0f;
    const int col0 = p.normalize ? 3 : 4;
    const int cols = p.C + col0;
    const int row_bytes = cols * 4;
    const int rows_per_stage = (half_bytes - 16) / row_bytes;

    // lane <-> ray: the ray of my pixel and the address of its feature vector
    float mo[3] = {0.0f, 0.0f, 0.0f}, md[3] = {0.0f, 0.0f, 0.0f};
    const T *my_feat = nullptr;
    if (my_keep > 0) {
        const int view = (int)(my_ray / hw);
        const int pix = (int)(my_ray % hw);
        const int u = pix % p.W, v = pix / p.W;
        ray_of_pixel(p.pinv + 16 * view, u, v, mo, md);
        my_feat = static_cast<const T *>(p.views[view - p.view_base]) + (int64_t)v * p.stride_y + (int64_t)u * p.stride_x;
    }

    int sbuf = 0, srows = 0, h = 0;        // warp-uniform staging state: half in use, rows in it, byte offset mod 16
    unsigned char *gbase = nullptr;        // global address of the first row in the stage
    auto flush = [&]() {
        float *stage = reinterpret_cast<float *>(stage_base + sbuf * half_bytes);
        const uint32_t stage_addr = (uint32_t)__cvta_generic_to_shared(stage);
        const int hw4 = h >> 2;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        const int total = srows * row_bytes;
        const int a0 = (h + 15) & ~15;
        const int a1 = (h + total) & ~15;
        if (lane == 0) {
            if (a1 > a0)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gbase - h + a0),
                             "r"(stage_addr + (uint32_t)a0), "r"((uint32_t)(a1 - a0))
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");      // one group per flush, possibly empty
        }
        const int head = (min(a0, h + total) - h) >> 2;
        const int tail_lo = max(a1, a0);
        const int tail = (h + total > tail_lo) ? ((h + total - tail_lo) >> 2) : 0;
        if (lane < head) reinterpret_cast<float *>(gbase)[lane] = stage[hw4 + lane];
        if (lane < tail) reinterpret_cast<float *>(gbase - h + tail_lo)[lane] = stage[(tail_lo >> 2) + lane];
        sbuf ^= 1;
        srows = 0;
        // the other half was pushed one flush ago: its bulk store must have finished reading before it is refilled
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
    };

    for (int r = 0; r < 32; ++r) {
        const int cnt = __shfl_sync(0xffffffffu, my_keep, r);
        if (cnt == 0) continue;
        const int64_t ray = ray0 + warp * 32 + r;
        float o[3], d[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            o[a] = __shfl_sync(0xffffffffu, mo[a], r);
            d[a] = __shfl_sync(0xffffffffu, md[a], r);
        }
        const T *feat = reinterpret_cast<const T *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(my_feat), r));
        float f[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int c = j * 32 + lane;
            f[j] = (c < p.C) ? load_feat<T>(feat + c) : 0.0f;
        }
        const int cnt_all = SELECT ? __shfl_sync(0xffffffffu, my_cnt, r) : cnt;      // records of the ray (cnt: rows produced)
        const int64_t row0 = SELECT ? __shfl_sync(0xffffffffu, my_off, r) : 0;          // un-selected index of its first row
        for (int k0 = 0; k0 < cnt_all; k0 += 32) {
            const int nk = min(32, cnt_all - k0);
            float wk = 0.0f, wraw = 0.0f, pos[3] = {0.0f, 0.0f, 0.0f};
            bool sel = lane < nk;
            if (SELECT && sel) sel = (row0 + k0 + lane < p.sel_rows) && __ldg(p.sel_mask + row0 + k0 + lane) != 0;
            if (sel) {
                wraw = __ldg(p.rec_w + (int64_t)(k0 + lane) * p.rays + ray);
                const float fi = __ldg(p.rec_i + (int64_t)(k0 + lane) * p.rays + ray);
                record_position(p, o, d, fi, pos);
                if (SELECT) {   // coord + offsets[b]: one more rounding (rm.py:365)
                    pos[0] = __fadd_rn(pos[0], p.sel_off[0]);
                    pos[1] = __fadd_rn(pos[1], p.sel_off[1]);
                    pos[2] = __fadd_rn(pos[2], p.sel_off[2]);
                }
                wk = p.normalize ? __fdiv_rn(wraw, mean) : 1.0f;   // weights / mean(weights), rm.py:303
            }
            if (SELECT) {   // row by row over the selected records
                unsigned bits = __ballot_sync(0xffffffffu, sel);
                while (bits) {
                    const int kk = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if (srows == 0) {
                        gbase = warp_gbase + (int64_t)rows_done * row_bytes;
                        h = (int)(reinterpret_cast<uintptr_t>(gbase) & 15);
                    }
                    float *row = reinterpret_cast<float *>(stage_base + sbuf * half_bytes) + (h >> 2) + srows * cols;
                    const float wn = __shfl_sync(0xffffffffu, wk, kk);
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int c = j * 32 + lane;
                        if (c < p.C) row[col0 + c] = p.normalize ? __fmul_rn(f[j], wn) : f[j];
                    }
                    if (lane == kk) {
                        row[0] = pos[0];
                        row[1] = pos[1];
                        row[2] = pos[2];
                        if (!p.normalize) row[3] = wraw;
                    }
                    ++srows;
                    ++rows_done;
                    if (srows == rows_per_stage) flush();
                }
                continue;
            }
            int k = 0;
            while (k < nk) {
                if (srows == 0) {   // rays of a warp own consecutive output rows, so a stage simply continues where the last ended
                    gbase = warp_gbase + (int64_t)rows_done * row_bytes;
                    h = (int)(reinterpret_cast<uintptr_t>(gbase) & 15);
                }
                float *stage = reinterpret_cast<float *>(stage_base + sbuf * half_bytes);
                const int hw4 = h >> 2;
                const int nb = min(rows_per_stage - srows, nk - k);
                for (int kk = 0; kk < nb; ++kk) {
                    const float wn = __shfl_sync(0xffffffffu, wk, k + kk);
                    float *row = stage + hw4 + (srows + kk) * cols + col0;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int c = j * 32 + lane;
                        if (c < p.C) row[c] = p.normalize ? __fmul_rn(f[j], wn) : f[j];
                    }
                }
                if (lane >= k && lane < k + nb) {
                    float *row = stage + hw4 + (srows + lane - k) * cols;
                    row[0] = pos[0];
                    row[1] = pos[1];
                    row[2] = pos[2];
                    if (!p.normalize) row[3] = wraw;
                }
                srows += nb;
                rows_done += nb;
                k += nb;
                if (srows == rows_per_stage) flush();
            }
        }
    }
    if (srows > 0) flush();
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <bool SCATTER>
static cudaError_t launch_fill(FillParams &p, int dtype, const void *const *view_ptrs_host, int V_total,
                               cudaStream_t stream) {
    const int hw = p.H * p.W;
    // One launch normally.  Beyond kMaxViewsPerLaunch views (pointer table in kernel params) the views go
    // out in groups whose ray count is a whole number of CTAs, so that the per-CTA row offsets written by
    // the march kernel line up with this kernel's CTAs.
    int group = V_total;
    if (V_total > kMaxViewsPerLaunch) {
        group = kMaxViewsPerLaunch;
        while (group > 0 && ((int64_t)group * hw) % kRayThreads != 0) --group;
        if (group == 0) return cudaErrorInvalidValue;
    }
    for (int v0 = 0; v0 < V_total; v0 += group) {
        const int nv = (V_total - v0 < group) ? (V_total - v0) : group;
        p.view_base = v0;
        p.V = nv;
        for (int i = 0; i < nv; ++i) p.views[i] = view_ptrs_host[v0 + i];
        const int64_t rays = (int64_t)nv * hw;
        const unsigned blocks = (unsigned)((rays + kRayThreads - 1) / kRayThreads);
        // packed rows (row_stride == columns) of up to 256 channels take the TMA-store kernel
        const int cols = p.C + (p.normalize ? 3 : 4);
        const bool tma = !SCATTER && p.sel_mask == nullptr && p.row_stride == cols && p.C <= 32 * kFillRegs &&
                         cols * 4 <= kStageBytes - 16 &&
                         reinterpret_cast<uintptr_t>(p.rows) % 4 == 0;
        size_t smem = (size_t)(kRayThreads / kWarp) * kStageBytes;
        static thread_local int attr_dev[24] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                                -1, -1, -1, -1, -1, -1, -1, -1};   // dynamic-smem opt-in, once per device
        int dev = 0;
        cudaError_t attr_err = cudaGetDevice(&dev);
        if (attr_err != cudaSuccess) return attr_err;
        const int nj = (p.C + 31) / 32;
        auto go = [&](auto kernel, int slot) {
            if (attr_dev[slot] != dev) {
                attr_err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)((kRayThreads / kWarp) * kStageBytes));
                if (attr_err != cudaSuccess) return;   // reported after the dispatch below; nothing is launched
                attr_dev[slot] = dev;
            }
            kernel<<<blocks, kRayThreads, smem, stream>>>(p);
        };
        // hand-off selection fused in: the packed kernel's SELECT form (CNRMA_FILL_SELECT_KERNEL=scalar: the plain kernel)
        bool select_packed = !SCATTER && p.sel_mask != nullptr && p.row_stride == cols && p.C <= 32 * kFillRegs &&
                             cols * 4 <= kHalfStageBytes - 16 && reinterpret_cast<uintptr_t>(p.rows) % 4 == 0;
        if (tuning().fill_select_scalar) select_packed = false;   // CNRMA_FILL_SELECT_KERNEL=scalar
        if (select_packed) {
            p.stage_half = (cols * 4 <= 2048 - 16) ? 2048 : kHalfStageBytes;
            smem = (size_t)(kRayThreads / kWarp) * 2 * p.stage_half;
            if (dtype == CNRMA_BF16) {
                if (nj <= 1) go(fill_rows_packed_kernel<__nv_bfloat16, 1, true>, 16);
                else if (nj <= 2) go(fill_rows_packed_kernel<__nv_bfloat16, 2, true>, 17);
                else if (nj <= 4) go(fill_rows_packed_kernel<__nv_bfloat16, 4, true>, 18);
                else go(fill_rows_packed_kernel<__nv_bfloat16, 8, true>, 19);
            } else {
                if (nj <= 1) go(fill_rows_packed_kernel<float, 1, true>, 20);
                else if (nj <= 2) go(fill_rows_packed_kernel<float, 2, true>, 21);
                else if (nj <= 4) go(fill_rows_packed_kernel<float, 4, true>, 22);
                else go(fill_rows_packed_kernel<float, 8, true>, 23);
            }
        } else if (tma) {
            // short rows: per-ray work dominates -> the packed kernel (CNRMA_FILL_KERNEL=tma|packed overrides)
            bool packed = cols * 4 <= kPackedRowBytesMax;
            if (tuning().fill_kernel >= 0) packed = (tuning().fill_kernel == 1) && cols * 4 <= kHalfStageBytes - 16;   // CNRMA_FILL_KERNEL
            if (packed) {
                // small staging halves for short rows: more resident CTAs (the kernel is latency-bound there)
                // (measured: 2 KB halves beat 4 KB ones on 140-, 268- and 524-byte rows alike; 1 KB ones lose)
                p.stage_half = (cols * 4 <= 2048 - 16) ? 2048 : kHalfStageBytes;
                if (tuning().fill_stage_half > 0) p.stage_half = tuning().fill_stage_half;   // CNRMA_FILL_STAGE_HALF
                smem = (size_t)(kRayThreads / kWarp) * 2 * p.stage_half;
                if (dtype == CNRMA_BF16) {
                    if (nj <= 1) go(fill_rows_packed_kernel<__nv_bfloat16, 1, false>, 8);
                    else if (nj <= 2) go(fill_rows_packed_kernel<__nv_bfloat16, 2, false>, 9);
                    else if (nj <= 4) go(fill_rows_packed_kernel<__nv_bfloat16, 4, false>, 10);
                    else go(fill_rows_packed_kernel<__nv_bfloat16, 8, false>, 11);
                } else {
                    if (nj <= 1) go(fill_rows_packed_kernel<float, 1, false>, 12);
                    else if (nj <= 2) go(fill_rows_packed_kernel<float, 2, false>, 13);
                    else if (nj <= 4) go(fill_rows_packed_kernel<float, 4, false>, 14);
                    else go(fill_rows_packed_kernel<float, 8, false>, 15);
                }
            } else if (dtype == CNRMA_BF16) {
                if (nj <= 1) go(fill_rows_tma_kernel<__nv_bfloat16, 1>, 0);
                else if (nj <= 2) go(fill_rows_tma_kernel<__nv_bfloat16, 2>, 1);
                else if (nj <= 4) go(fill_rows_tma_kernel<__nv_bfloat16, 4>, 2);
                else go(fill_rows_tma_kernel<__nv_bfloat16, 8>, 3);
            } else {
                if (nj <= 1) go(fill_rows_tma_kernel<float, 1>, 4);
                else if (nj <= 2) go(fill_rows_tma_kernel<float, 2>, 5);
                else if (nj <= 4) go(fill_rows_tma_kernel<float, 4>, 6);
                else go(fill_rows_tma_kernel<float, 8>, 7);
            }
        } else if (dtype == CNRMA_BF16) {
            fill_rows_kernel<__nv_bfloat16, SCATTER><<<blocks, kRayThreads, 0, stream>>>(p);
        } else {
            fill_rows_kernel<float, SCATTER><<<blocks, kRayThreads, 0, stream>>>(p);
        }
        if (attr_err != cudaSuccess) return attr_err;   // the shared-memory opt-in of a fill kernel failed
        const cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

cudaError_t run_fill(const GridDev &g, const float *pinv, const cnrma_features &f, float t_one, int mode,
                     const void *workspace, const RmaWorkspace &ws, int normalize, const float *mean, float *rows,
                     int64_t row_stride, int64_t capacity, float *wsum, float *wtot, const uint8_t *sel_mask,
                     const int32_t *sel_prefix, const float *sel_off_host, int64_t sel_rows, cudaStream_t stream) {
    const unsigned char *base = static_cast<const unsigned char *>(workspace);
    FillParams p;
    p.g = g;
    p.pinv = pinv;
    p.V = f.views; p.C = f.channels; p.H = f.height; p.W = f.width;
    p.stride_y = f.stride_y; p.stride_x = f.stride_x;
    p.t_one = t_one;
    p.mode = mode;
    p.normalize = normalize;
    p.rays = ws.rays;
    p.counts = reinterpret_cast<const int32_t *>(base + ws.off_counts);
    p.blk_off = reinterpret_cast<const int64_t *>(base + ws.off_blk_off);
    p.rec_w = reinterpret_cast<const float *>(base + ws.off_rec_w);
    p.rec_i = reinterpret_cast<const float *>(base + ws.off_rec_i);
    p.mean = mean;
    p.rows = rows;
    p.row_stride = row_stride;
    p.capacity = capacity;
    p.sel_mask = sel_mask;
    p.sel_prefix = sel_prefix;
    p.sel_rows = sel_rows > 0 ? sel_rows : INT64_MAX;
    for (int a = 0; a < 3; ++a) p.sel_off[a] = sel_off_host ? sel_off_host[a] : 0.0f;
    p.wsum = wsum;
    p.wtot = wtot;
    p.view_base = 0;
    if (wsum != nullptr) return launch_fill<true>(p, f.dtype, f.view_ptrs_host, f.views, stream);
    return launch_fill<false>(p, f.dtype, f.view_ptrs_host, f.views, stream);
}

cudaError_t run_expand(int64_t rays, int N, const void *workspace, const RmaWorkspace &ws, float *weights,
                       uint8_t *keep, cudaStream_t stream) {
    const unsigned char *base = static_cast<const unsigned char *>(workspace);
    cudaError_t err = cudaSuccess;
    if (weights) err = cudaMemsetAsync(weights, 0, sizeof(float) * (size_t)rays * N, stream);
    if (err == cudaSuccess && keep) err = cudaMemsetAsync(keep, 0, (size_t)rays * N, stream);
    if (err != cudaSuccess) return err;
    expand_records_kernel<<<(unsigned)((rays + 255) / 256), 256, 0, stream>>>(
        rays, N, reinterpret_cast<const int32_t *>(base + ws.off_counts),
        reinterpret_cast<const float *>(base + ws.off_rec_w), reinterpret_cast<const float *>(base + ws.off_rec_i),
        weights, keep);
    return cudaGetLastError();
}

}  // namespace cnrma
