// Shared device helpers for libcnrma_b200 (sm_100a only; no CPU path).
//
// Numerics contract (DESIGN.md "Numerics"): the reference's fp32 torch ops are reproduced one rounding
// at a time with explicit round-to-nearest intrinsics, so results do not depend on nvcc's contraction
// choices; the library is additionally compiled with --fmad=false.
#pragma once

#include <cmath>
#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cnrma_b200.h"

namespace cnrma {

constexpr int kWarp = 32;
constexpr int kAggThreads = 256;            // 8 warps per CTA
constexpr int kMaxViewsPerLaunch = 512;     // per-view device pointers travel in kernel params (4 KB of the 32 KB CUDA >= 12.1 allows)
constexpr int kRayThreads = 256;            // rays per CTA in the march / fill kernels

// Tuning / test knobs (environment variables, DESIGN.md "Tuning / test knobs"), read ONCE per process -- the launch path
// never calls getenv -- and again only on cnrma_reload_tuning() (the tests flip knobs between calls).
struct Tuning {
    int agg_kernel = -1;        // CNRMA_AGG_KERNEL: -1 automatic, 0 tma, 1 list
    int agg_slab = 0;           // CNRMA_AGG_SLAB: slab thickness of the Stage A sweep, 0 automatic
    int agg_tile = 0;           // CNRMA_AGG_TILE: (x, y) tile edge inside a slab, 0 none
    int agg_cull = -1;          // CNRMA_AGG_CULL: -1 automatic, 0 voxel units, 1 column units with view culling, 2 CTA columns
    int agg_pipe = -1;          // CNRMA_AGG_PIPE: 1 selects the software-pipelined form of the TMA kernel (opt-in)
    int agg_chunk_bytes = 0;    // CNRMA_AGG_CHUNK_BYTES
    int agg_warp_buffer = 0;    // CNRMA_AGG_WARP_BUFFER
    int agg_list_views = 0;     // CNRMA_AGG_LIST_VIEWS
    int agg_bwd_kernel = -1;    // CNRMA_AGG_BWD_KERNEL: -1 automatic, 0 bulk, 1 list
    int bilinear_simple = 0;    // CNRMA_BILINEAR_SIMPLE
    int march_unfused = 0;      // CNRMA_MARCH_UNFUSED_PREPASS
    int march_jump = -1;        // CNRMA_MARCH_JUMP: -1 automatic, 0 clearance only, 1 clearance + position inside the voxel
    int fill_kernel = -1;       // CNRMA_FILL_KERNEL: -1 automatic, 0 tma, 1 packed
    int fill_stage_half = 0;    // CNRMA_FILL_STAGE_HALF
    int fill_select_scalar = 0; // CNRMA_FILL_SELECT_KERNEL=scalar
};
const Tuning &tuning();

struct GridDev {
    int nx, ny, nz;
    float vs;
    float ox, oy, oz;
    int x0, y0, z0;   // Stage A forward on a box of the grid: (nx, ny, nz) are the box's extents, (x0, y0, z0) its first voxel
};

static inline GridDev to_dev(const cnrma_grid &g) {
    return GridDev{g.nx, g.ny, g.nz, g.voxel_size, g.origin[0], g.origin[1], g.origin[2], 0, 0, 0};
}

// The sub-volume [lo, lo + dim) of `g`: world coordinates are formed from the GLOBAL voxel index (float(i) * vs + origin,
// two roundings, rm.py:48), so a box reproduces the bits of the full-grid pass; outputs are indexed inside the box.
static inline GridDev to_dev_box(const cnrma_grid &g, const cnrma_box &b) {
    return GridDev{b.dim[0], b.dim[1], b.dim[2], g.voxel_size, g.origin[0], g.origin[1], g.origin[2], b.lo[0], b.lo[1], b.lo[2]};
}

// torch.bmm([3|4]x4 @ 4xN) in fp32 == this FMA chain in k order (rm.py:51, :105-107).
__device__ __forceinline__ float row_dot4(float p0, float p1, float p2, float p3, float x, float y, float z,
                                          float w) {
    float acc = __fmul_rn(p0, x);
    acc = __fmaf_rn(p1, y, acc);
    acc = __fmaf_rn(p2, z, acc);
    acc = __fmaf_rn(p3, w, acc);
    return acc;
}

// round(cx / cz) and round(cy / cz) (rm.py:52-53: IEEE division, half-to-even).  Only the rounded integers are
// needed, so the two divisions are replaced by one correctly rounded reciprocal and two multiplications whenever that
// provably cannot change the result: q = RN(c * RN(1/cz)) is within |q| * 2^-23 of the exact quotient, hence
// rint(q) == rint(RN(c / cz)) unless q lies within |q| * 2^-21 of a half-integer -- then (and for NaN / inf / huge
// quotients) the divisions are done for real.
__device__ __forceinline__ void rounded_pixel(float cx, float cy, float cz, float &rx, float &ry) {
    const float r = __frcp_rn(cz);
    const float qx = __fmul_rn(cx, r), qy = __fmul_rn(cy, r);
    rx = rintf(qx);
    ry = rintf(qy);
    const bool safe = (fabsf(__fsub_rn(qx, rx)) < __fmaf_rn(fabsf(qx), -4.76837158203125e-07f, 0.5f)) &&
                      (fabsf(__fsub_rn(qy, ry)) < __fmaf_rn(fabsf(qy), -4.76837158203125e-07f, 0.5f));
    if (!safe) {
        rx = rintf(__fdiv_rn(cx, cz));
        ry = rintf(__fdiv_rn(cy, cz));
    }
}

// Frustum test of rm.py:58 on the rounded pixel coordinates.  The reference converts to int64 and compares; comparing
// the integer-valued floats is equivalent for every finite value, and NaN / +-inf (which x86 turns into INT64_MIN)
// fail the test either way.
__device__ __forceinline__ bool in_frustum(float rx, float ry, float cz, int H, int W) {
    return (rx >= 0.0f) && (ry >= 0.0f) && (rx < (float)W) && (ry < (float)H) && (cz > 0.0f);
}

// One voxel through one stride-scaled projection (rm.py:48-58).  P points at 12 floats with element
// stride `ps` (shared memory, structure-of-arrays over views).  Returns true when the voxel is inside
// the view frustum; px/py are the rounded pixel coordinates.
__device__ __forceinline__ bool project_voxel(const float *P, int ps, float wx, float wy, float wz, int H, int W,
                                              int &px, int &py) {
    const float cx = row_dot4(P[0 * ps], P[1 * ps], P[2 * ps], P[3 * ps], wx, wy, wz, 1.0f);
    const float cy = row_dot4(P[4 * ps], P[5 * ps], P[6 * ps], P[7 * ps], wx, wy, wz, 1.0f);
    const float cz = row_dot4(P[8 * ps], P[9 * ps], P[10 * ps], P[11 * ps], wx, wy, wz, 1.0f);
    float rx, ry;
    rounded_pixel(cx, cy, cz, rx, ry);
    const bool ok = in_frustum(rx, ry, cz, H, W);
    px = ok ? (int)rx : 0;
    py = ok ? (int)ry : 0;
    return ok;
}

// world = float(index) * voxel_size + origin: two roundings (rm.py:48).
__device__ __forceinline__ float world_coord(int i, float vs, float o) { return __fadd_rn(__fmul_rn((float)i, vs), o); }

// get_ray_parameter for one pixel (rm.py:71-111) given Pinv (16 floats, row major).
__device__ __forceinline__ void ray_of_pixel(const float *__restrict__ Pinv, int u, int v, float o[3], float d[3]) {
    const float fu = (float)u, fv = (float)v;
    const float zu = __fmul_rn(fu, 0.0f), zv = __fmul_rn(fv, 0.0f);
    float raw[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float p0 = Pinv[4 * r + 0], p1 = Pinv[4 * r + 1], p2 = Pinv[4 * r + 2], p3 = Pinv[4 * r + 3];
        o[r] = row_dot4(p0, p1, p2, p3, zu, zv, 0.0f, 1.0f);
        raw[r] = __fsub_rn(row_dot4(p0, p1, p2, p3, fu, fv, 1.0f, 1.0f), o[r]);
    }
    // F.normalize: three separately rounded squares summed left to right, sqrt, clamp_min(1e-12), divide
    float ss = __fadd_rn(__fmul_rn(raw[0], raw[0]), __fmul_rn(raw[1], raw[1]));
    ss = __fadd_rn(ss, __fmul_rn(raw[2], raw[2]));
    float nrm = __fsqrt_rn(ss);
    nrm = (nrm < 1e-12f) ? 1e-12f : nrm;
#pragma unroll
    for (int r = 0; r < 3; ++r) d[r] = __fdiv_rn(raw[r], nrm);
}

// ---- 16-byte feature vectors ---------------------------------------------------------------------
template <typename T>
struct Vec16;

template <>
struct Vec16<float> {
    static constexpr int kElems = 4;
    float v[4];
    __device__ __forceinline__ static Vec16 load(const float *p) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        Vec16 r;
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
        return r;
    }
    __device__ __forceinline__ static Vec16 load_shared(const void *p) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        Vec16 r;
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
        return r;
    }
    // the 16 bytes as loaded; widened later, so that several loads can be in flight before the first conversion
    __device__ __forceinline__ static uint4 load_raw(const float *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    __device__ __forceinline__ static Vec16 widen(const uint4 t) {
        Vec16 r;
        r.v[0] = __uint_as_float(t.x); r.v[1] = __uint_as_float(t.y); r.v[2] = __uint_as_float(t.z); r.v[3] = __uint_as_float(t.w);
        return r;
    }
};

template <>
struct Vec16<__nv_bfloat16> {
    static constexpr int kElems = 8;
    float v[8];
    __device__ __forceinline__ static Vec16 load(const __nv_bfloat16 *p) {
        return widen(__ldg(reinterpret_cast<const uint4 *>(p)));
    }
    __device__ __forceinline__ static Vec16 load_shared(const void *p) { return widen(*reinterpret_cast<const uint4 *>(p)); }
    __device__ __forceinline__ static uint4 load_raw(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    __device__ __forceinline__ static Vec16 widen(const uint4 t) {
        Vec16 r;
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r.v[2 * i + 0] = __uint_as_float(w[i] << 16);          // bf16 -> f32 is exact
            r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
        return r;
    }
};

// ---- mbarrier / bulk-copy primitives (PTX) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy global -> this CTA's shared memory, completion (bytes) signalled on `bar`.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// n / d and n % d for 0 <= n < 2^31, 0 < d, quotient < 2^22: float estimate, exact integer fix-up (the estimate is
// within 1 of the true quotient), ~8 instructions instead of the ~40 of a generic 32-bit division.
__device__ __forceinline__ void fast_divmod(int n, int d, float inv_d, int &q, int &r) {
    q = (int)((float)n * inv_d);
    r = n - q * d;
    if (r < 0) { r += d; --q; }
    else if (r >= d) { r -= d; ++q; }
}

// Traversal order of the Stage A kernels (forward and backward): the volume is swept in slabs of T z-slices, inside a
// slab x is the slow axis, then y, then the slab's z.  Voxels that share a pixel row lie along one camera ray, 8+
// voxels apart and a few slices up or down; what the resident warps touch at any time is a window of a few thousand
// voxels (L2 size / bytes gathered per voxel), and a window of ~T x-rows by all y by T slices keeps more of those
// pairs together than one of whole z-slices (T = 1) or whole columns (T = nz).  LRU simulation and measurement in
// DESIGN.md "K_A".  T is chosen on the host: sweep_thickness().
struct SweepOrder {
    int ny, T, Tlast, nfull, full;   // full = nx*ny*T voxels per full slab, nfull full slabs, then one of Tlast slices
    float inv_full, inv_T, inv_Tlast, inv_ny;
    // optional (x, y) tiling inside a slab (tile > 0, both extents multiples of it): tiles of tile x tile columns are
    // swept one after the other, x-major, each column over the slab's z -- a more compact window in x and y
    int tile, tiles_y;
    float inv_tile, inv_tile2, inv_tiles_y;
};

inline SweepOrder make_sweep(int nx, int ny, int nz, int T) {
    SweepOrder s;
    s.ny = ny;
    s.T = T < 1 ? 1 : (T > nz ? nz : T);
    s.nfull = nz / s.T;
    s.Tlast = nz - s.nfull * s.T;
    if (s.Tlast == 0) s.Tlast = s.T;   // unused then; keeps the reciprocal finite
    s.full = nx * ny * s.T;
    s.inv_full = 1.0f / (float)s.full;
    s.inv_T = 1.0f / (float)s.T;
    s.inv_Tlast = 1.0f / (float)s.Tlast;
    s.inv_ny = 1.0f / (float)ny;
    s.tile = 0;
    s.tiles_y = 1;
    s.inv_tile = s.inv_tile2 = s.inv_tiles_y = 1.0f;
    {   // tuning aid (CNRMA_AGG_TILE)
        const int t = tuning().agg_tile;
        if (t > 1 && nx % t == 0 && ny % t == 0) {
            s.tile = t;
            s.tiles_y = ny / t;
            s.inv_tile = 1.0f / (float)t;
            s.inv_tile2 = 1.0f / (float)(t * t);
            s.inv_tiles_y = 1.0f / (float)s.tiles_y;
        }
    }
    return s;
}

// Slab thickness: the window holds about L2_window / (visible views x row bytes) voxels; make it as deep in z as it is
// long in x (it always spans y): T = sqrt(window_voxels / ny).  A quarter of the views see a voxel in the scenes at hand.
inline int sweep_thickness(int ny, int nz, int views, int row_bytes) {
    if (tuning().agg_slab > 0) return tuning().agg_slab;   // tuning aid (CNRMA_AGG_SLAB)
    const double per_voxel = 0.25 * (views > 0 ? views : 1) * (row_bytes > 0 ? row_bytes : 1);
    const double window_voxels = 64.0 * 1024 * 1024 / per_voxel;
    int T = (int)(std::sqrt(window_voxels / (ny > 0 ? ny : 1)) + 0.5);
    return T < 1 ? 1 : (T > nz ? nz : T);
}

__device__ __forceinline__ void sweep_voxel(const SweepOrder &s, int it, int &vx, int &vy, int &vz) {
    int slab, rem, xy, zi;
    fast_divmod(it, s.full, s.inv_full, slab, rem);
    const bool last = slab >= s.nfull;
    fast_divmod(rem, last ? s.Tlast : s.T, last ? s.inv_Tlast : s.inv_T, xy, zi);
    vz = slab * s.T + zi;
    if (s.tile > 0) {
        int t, in_tile, tx, ty, lx, ly;
        fast_divmod(xy, s.tile * s.tile, s.inv_tile2, t, in_tile);
        fast_divmod(t, s.tiles_y, s.inv_tiles_y, tx, ty);
        fast_divmod(in_tile, s.tile, s.inv_tile, lx, ly);
        vx = tx * s.tile + lx;
        vy = ty * s.tile + ly;
        return;
    }
    fast_divmod(xy, s.ny, s.inv_ny, vx, vy);
}

// Output routing of the view-sharded Stage A over peer memory (cnrma_aggregate_views_routed): the voxels are split
// into `n_owners` contiguous ranges of `slab` voxels; voxel v belongs to owner v / slab, and its un-normalised sums
// (C floats) and view count (one float, at [C]) go to row v % slab of that owner's buffer -- a peer-mapped pointer
// when the owner is another GPU, so the kernel's stores cross NVLink themselves.  n_owners == 0: not routed.
constexpr int kMaxOwners = 8;
struct OutputRoute {
    int n_owners, slab, row_floats;
    float inv_slab;
    float *owner_base[kMaxOwners];   // this source's section of each owner's buffer: [slab][row_floats]
};

__device__ __forceinline__ float *route_row(const OutputRoute &r, int vox) {
    int o, lv;
    fast_divmod(vox, r.slab, r.inv_slab, o, lv);
    return r.owner_base[o] + (int64_t)lv * r.row_floats;
}

// IEEE-correct a / n for a small positive integer n, given y = RN(1/n): q = RN(a*y); r = a - n*q (exact, FMA);
// q' = RN(q + r*y) (Markstein).  Values whose residual could leave the normal range take the generic path.
__device__ __forceinline__ float div_by_count(float a, float n, float y) {
    const float mag = fabsf(a);
    if (!(mag > 1e-30f && mag < 1e30f)) return __fdiv_rn(a, n);
    const float q = __fmul_rn(a, y);
    const float r = __fmaf_rn(-n, q, a);
    return __fmaf_rn(r, y, q);
}

}  // namespace cnrma
