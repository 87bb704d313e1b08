/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the CN-RMA ray-marching aggregation path.
 *
 * A plain-C scalar restatement of the algorithm in the reference's
 *   projects/mvsdetection/models/ray_marching.py   (cited below as rm.py:LINE)
 *   projects/mvsdetection/datasets/tsdf.py         (tsdf.py:LINE)
 * used ONLY as the checker in tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  It is never imported, linked or called by
 * the product package (cn-rma_b200/), which has no CPU path at all.
 *
 * Parity status: the reference ships no tests, golden vectors or fixtures for this
 * path, so this oracle is pinned against outputs of the
 * reference itself: oracle/make_golden.py imports the unmodified reference in the
 * build container and commits input/output vectors under tests/golden/;
 * tests/test_oracle_golden.py replays them against this file (Stage A indices,
 * masks, sums: bit-exact; Stage B weights: <= 1e-5 relative, because the
 * reference's torch.sigmoid uses a vectorised exp that no libm reproduces bitwise).
 *
 * Numerics contract restated from the reference's fp32 torch ops (measured
 * against torch 2.11 CPU in the build container, see DESIGN.md "Numerics"):
 *   - torch.bmm of a [3|4 x 4] matrix with a [4 x N] matrix == sequential FMA chain in
 *     k order:  acc = p0*x; acc = fma(p1,y,acc); acc = fma(p2,z,acc); acc = fma(p3,w,acc)
 *   - every other elementwise torch op is one correctly rounded fp32 operation
 *     (no contraction across ops): compile with -ffp-contract=off.
 *   - Tensor.round() is round-half-to-even (rintf).
 *   - Tensor / python_scalar on CPU is a true IEEE division by (float)scalar.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

ORACLE_API int cnrma_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORACLE_API void cnrma_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* float -> int64 the way x86 cvttss2si does it for the reference's `.type(torch.long)`
 * (rm.py:52-53, :730): out-of-range and NaN give INT64_MIN.  Only used for values
 * that the validity mask discards anyway. */
static inline int64_t f2l(float r) {
    if (!(r >= -9.2233720368547758e18f && r < 9.2233720368547758e18f)) return INT64_MIN;
    return (int64_t)r;
}

/* rm.py:238-239 / :275-276 -- rows 0 and 1 of the 3x4 projection divided by the stride. */
ORACLE_API void cnrma_oracle_scale_projection(const float *p_in, float stride, float *p_out) {
    for (int i = 0; i < 12; ++i) p_out[i] = (i < 8) ? p_in[i] / stride : p_in[i];
}

/* One voxel through one (already stride-scaled) projection.
 * tsdf.py:24-29  voxel order: flat = (x*ny + y)*nz + z
 * rm.py:48       world = float(coord) * voxel_size + origin   (two roundings)
 * rm.py:51       camera = bmm(P, [world;1])                   (FMA chain, k order)
 * rm.py:52-54    px = round(cx/cz), py = round(cy/cz), pz = cz
 * rm.py:58       valid = px>=0 & py>=0 & px<W & py<H & pz>0
 */
static inline int project_voxel(const float *P, float voxel_size, const float *origin, int x, int y, int z,
                                int H, int W, int64_t *px_out, int64_t *py_out) {
    float wx = (float)x * voxel_size + origin[0];
    float wy = (float)y * voxel_size + origin[1];
    float wz = (float)z * voxel_size + origin[2];
    float cam[3];
    for (int r = 0; r < 3; ++r) {
        float acc = P[4 * r + 0] * wx;
        acc = fmaf(P[4 * r + 1], wy, acc);
        acc = fmaf(P[4 * r + 2], wz, acc);
        acc = fmaf(P[4 * r + 3], 1.0f, acc);
        cam[r] = acc;
    }
    int64_t px = f2l(rintf(cam[0] / cam[2]));
    int64_t py = f2l(rintf(cam[1] / cam[2]));
    *px_out = px;
    *py_out = py;
    return (px >= 0) & (py >= 0) & (px < W) & (py < H) & (cam[2] > 0.0f);
}

/* rm.py:47-58 for one view: per-voxel pixel indices and frustum mask.
 * P is the *scaled* 3x4 projection. px/py int64 [nvox], valid uint8 [nvox]. */
ORACLE_API void cnrma_oracle_project(int nx, int ny, int nz, float voxel_size, const float *origin, const float *P,
                                     int H, int W, int64_t *px, int64_t *py, uint8_t *valid) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z) {
                size_t i = ((size_t)x * ny + y) * nz + z;
                valid[i] = (uint8_t)project_voxel(P, voxel_size, origin, x, y, z, H, W, &px[i], &py[i]);
            }
}

/* Stage A, all views of one batch element:
 *   rm.py:21-69 backproject (nearest gather, zero where invalid)
 *   rm.py:220-244 aggregate_2d_features (volume += v; valid += mask, in view order)
 *   rm.py:247-257 clear_3d_features (volume / count, 0 where count == 0) when `mean` != 0
 * projections: [V,3,4] UN-scaled; features: [V,C,H,W] (NCHW); volume: [C,nvox]; count: int64 [nvox].
 *
 * Loop order follows the reference's data flow (per view: indices once, then every channel plane gathered
 * through them), which is also the cache-friendly order for NCHW maps: phase 1 builds, per view, the list of
 * (voxel, pixel offset) pairs inside the frustum; phase 2 walks channels in parallel and, for each channel,
 * the views in order -- so every volume element still receives its additions in view order (bit-exact with
 * the reference's running sum; the zeros it adds for invisible views do not change an fp32 sum that starts
 * at +0). */
ORACLE_API void cnrma_oracle_aggregate_views(int V, int C, int H, int W, int nx, int ny, int nz, float voxel_size,
                                             const float *origin, float stride, const float *projections,
                                             const float *features, int mean, float *volume, int64_t *count) {
    size_t nvox = (size_t)nx * ny * nz;
    size_t plane = (size_t)H * W;
    float *P = (float *)malloc(sizeof(float) * 12 * (size_t)(V > 0 ? V : 1));
    int32_t **vox_list = (int32_t **)calloc((size_t)(V > 0 ? V : 1), sizeof(int32_t *));
    int32_t **off_list = (int32_t **)calloc((size_t)(V > 0 ? V : 1), sizeof(int32_t *));
    size_t *n_list = (size_t *)calloc((size_t)(V > 0 ? V : 1), sizeof(size_t));
    for (int v = 0; v < V; ++v) cnrma_oracle_scale_projection(projections + 12 * v, stride, P + 12 * v);
#pragma omp parallel for schedule(dynamic, 1)
    for (int v = 0; v < V; ++v) {
        int32_t *vl = (int32_t *)malloc(sizeof(int32_t) * nvox);
        int32_t *ol = (int32_t *)malloc(sizeof(int32_t) * nvox);
        size_t n = 0;
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y)
                for (int z = 0; z < nz; ++z) {
                    int64_t px, py;
                    if (!project_voxel(P + 12 * v, voxel_size, origin, x, y, z, H, W, &px, &py)) continue;
                    vl[n] = (int32_t)(((size_t)x * ny + y) * nz + z);
                    ol[n] = (int32_t)(py * W + px);
                    ++n;
                }
        vox_list[v] = vl;
        off_list[v] = ol;
        n_list[v] = n;
    }
    for (size_t i = 0; i < nvox; ++i) count[i] = 0;
    for (int v = 0; v < V; ++v)
        for (size_t k = 0; k < n_list[v]; ++k) count[vox_list[v][k]] += 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < C; ++c) {
        float *acc = volume + (size_t)c * nvox;
        for (size_t i = 0; i < nvox; ++i) acc[i] = 0.0f;
        for (int v = 0; v < V; ++v) {
            const float *f = features + ((size_t)v * C + c) * plane;
            const int32_t *vl = vox_list[v], *ol = off_list[v];
            for (size_t k = 0; k < n_list[v]; ++k) acc[vl[k]] = acc[vl[k]] + f[ol[k]];
        }
        if (mean)
            for (size_t i = 0; i < nvox; ++i) acc[i] = (count[i] > 0) ? acc[i] / (float)count[i] : 0.0f;
    }
    for (int v = 0; v < V; ++v) {
        free(vox_list[v]);
        free(off_list[v]);
    }
    free(vox_list);
    free(off_list);
    free(n_list);
    free(P);
}

/* rm.py:71-111 get_ray_parameter for one pixel, given Pinv = inverse([P;0 0 0 1]) (4x4, row major).
 * o = (Pinv @ (0,0,0,1))[:3]; d = normalize((Pinv @ (u,v,1,1))[:3] - o), F.normalize eps = 1e-12.
 * The bmm rows are FMA chains in k order, so o reduces to column 3 of Pinv. */
static inline void ray_of_pixel(const float *Pinv, int u, int v, float *o, float *d) {
    float fu = (float)u, fv = (float)v;
    float raw[3];
    for (int r = 0; r < 3; ++r) {
        float zacc = Pinv[4 * r + 0] * (fu * 0.0f);
        zacc = fmaf(Pinv[4 * r + 1], fv * 0.0f, zacc);
        zacc = fmaf(Pinv[4 * r + 2], 0.0f, zacc);
        zacc = fmaf(Pinv[4 * r + 3], 1.0f, zacc);
        o[r] = zacc;
        float acc = Pinv[4 * r + 0] * fu;
        acc = fmaf(Pinv[4 * r + 1], fv, acc);
        acc = fmaf(Pinv[4 * r + 2], 1.0f, acc);
        acc = fmaf(Pinv[4 * r + 3], 1.0f, acc);
        raw[r] = acc - zacc;
    }
    /* F.normalize -> vector_norm(p=2, dim=1): three separately rounded squares summed left to right
     * (measured: the FMA form mismatches torch on ~1.6 % of rays, this form on none). */
    float ss = raw[0] * raw[0] + raw[1] * raw[1];
    ss = ss + raw[2] * raw[2];
    float nrm = sqrtf(ss);
    if (nrm < 1e-12f) nrm = 1e-12f;
    for (int r = 0; r < 3; ++r) d[r] = raw[r] / nrm;
}

ORACLE_API void cnrma_oracle_rays(const float *Pinv, int H, int W, float *o, float *d) {
    /* o: [3][H*W], d: [3][H*W] like the reference's B x 3 x (H*W) tensors */
    size_t hw = (size_t)H * W;
    for (int v = 0; v < H; ++v)
        for (int u = 0; u < W; ++u) {
            float oo[3], dd[3];
            ray_of_pixel(Pinv, u, v, oo, dd);
            for (int r = 0; r < 3; ++r) {
                o[r * hw + (size_t)v * W + u] = oo[r];
                d[r * hw + (size_t)v * W + u] = dd[r];
            }
        }
}

/* rm.py:710-715: t_one = sqrt(X^2+Y^2+Z^2) * voxel_size / N evaluated in python doubles, then
 * `arange(N, float32) * t_one` multiplies by (float)t_one. voxel_size here is the python double. */
ORACLE_API float cnrma_oracle_t_one(int X, int Y, int Z, double voxel_size, int N) {
    double t_max = sqrt((double)X * X + (double)Y * Y + (double)Z * Z) * voxel_size;
    return (float)(t_max / (double)N);
}

/* March one ray (rm.py:729-767). Writes per-step weight (0 where not kept) into w[N] if w != NULL, the
 * kept flag into keep[N] if keep != NULL, the sample positions into place[3*N] if place != NULL, and
 * the raw (unmasked) weight into wraw[N] if wraw != NULL. Returns the number of kept samples.
 *   places = o + d*t                         (mul, then add: two roundings)      rm.py:729
 *   id = round((places - origin)/voxel_size) (sub, true division, rint)          rm.py:730
 *   in-bounds mask, out-of-bounds tsdf := 1.0                                    rm.py:731-744
 *   s = sigmoid(-tsdf); alpha = max((s - s_next)/s, 0), last alpha 0             rm.py:757-759
 *   T = exclusive cumprod(1 - alpha); w = T*alpha                                rm.py:760-763
 *   keep = in-bounds & (w >= thr)                                                rm.py:765-767
 */
static int march_ray(const float *o, const float *d, float t_one, int N, const float *origin, float voxel_size,
                     int X, int Y, int Z, const float *tsdf, float thr, float *w, uint8_t *keep, float *place,
                     float *wraw) {
    int kept = 0;
    float T = 1.0f;
    float s_cur = 0.0f;
    int inb_cur = 0;
    float p_cur[3] = {0, 0, 0};
    for (int i = 0; i <= N; ++i) {
        /* sample i (i == N is the repeated last sample, rm.py:758) */
        float s_next;
        int inb_next = 0;
        float p_next[3] = {0, 0, 0};
        if (i < N) {
            float t = (float)i * t_one;
            int64_t id[3];
            for (int r = 0; r < 3; ++r) {
                float dt = d[r] * t;
                p_next[r] = o[r] + dt;
                float rel = p_next[r] - origin[r];
                id[r] = f2l(rintf(rel / voxel_size));
            }
            inb_next = (id[0] >= 0) & (id[0] < X) & (id[1] >= 0) & (id[1] < Y) & (id[2] >= 0) & (id[2] < Z);
            float tv = inb_next ? tsdf[((size_t)id[0] * Y + (size_t)id[1]) * Z + (size_t)id[2]] : 1.0f;
            s_next = 1.0f / (1.0f + expf(tv)); /* sigmoid(-tv) */
        } else {
            s_next = s_cur;
        }
        if (i > 0) {
            int j = i - 1;
            float a = (s_cur - s_next) / s_cur;
            if (a < 0.0f) a = 0.0f; /* clamp(min=0); NaN propagates like torch.clamp */
            float wj = T * a;
            int k = inb_cur && (wj >= thr);
            if (wraw) wraw[j] = wj;
            if (w) w[j] = k ? wj : wj * 0.0f; /* weights * valid_final, rm.py:767 */
            if (keep) keep[j] = (uint8_t)k;
            if (place) {
                place[3 * j + 0] = p_cur[0];
                place[3 * j + 1] = p_cur[1];
                place[3 * j + 2] = p_cur[2];
            }
            kept += k;
            T = T * (1.0f - a);
        }
        s_cur = s_next;
        inb_cur = inb_next;
        p_cur[0] = p_next[0];
        p_cur[1] = p_next[1];
        p_cur[2] = p_next[2];
    }
    return kept;
}

/* Dense per-sample outputs of one view, for parity checks of rm.py:729-767:
 *   w [H*W*N] masked weights, keep [H*W*N], places [H*W*N*3] (sample-major xyz), wraw unmasked. */
ORACLE_API void cnrma_oracle_neus_dense(const float *Pinv, int H, int W, int N, float t_one, const float *origin,
                                        float voxel_size, int X, int Y, int Z, const float *tsdf, float thr,
                                        float *w, uint8_t *keep, float *places, float *wraw) {
#pragma omp parallel for schedule(static)
    for (int pix = 0; pix < H * W; ++pix) {
        float o[3], d[3];
        ray_of_pixel(Pinv, pix % W, pix / W, o, d);
        size_t base = (size_t)pix * N;
        march_ray(o, d, t_one, N, origin, voxel_size, X, Y, Z, tsdf, thr, w ? w + base : NULL,
                  keep ? keep + base : NULL, places ? places + 3 * base : NULL, wraw ? wraw + base : NULL);
    }
}

/* Per-ray kept-sample counts of one view (first pass of the ordered compaction, rm.py:781). */
ORACLE_API int64_t cnrma_oracle_neus_count(const float *Pinv, int H, int W, int N, float t_one, const float *origin,
                                           float voxel_size, int X, int Y, int Z, const float *tsdf, float thr,
                                           int32_t *per_ray) {
    int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int pix = 0; pix < H * W; ++pix) {
        float o[3], d[3];
        ray_of_pixel(Pinv, pix % W, pix / W, o, d);
        int k = march_ray(o, d, t_one, N, origin, voxel_size, X, Y, Z, tsdf, thr, NULL, NULL, NULL, NULL);
        per_ray[pix] = k;
        total += k;
    }
    return total;
}

/* rm.py:769-807: rows [x, y, z, w, features[:, v, u]] of one view in ascending (v, u, step) order.
 * features: [C,H,W] of this view. rows: [M, 4+C] with M = sum(per_ray). per_ray from _neus_count. */
ORACLE_API void cnrma_oracle_neus_rows(const float *Pinv, int H, int W, int N, float t_one, const float *origin,
                                       float voxel_size, int X, int Y, int Z, const float *tsdf, float thr,
                                       int C, const float *features, const int32_t *per_ray, float *rows) {
    size_t hw = (size_t)H * W;
    int64_t *offset = (int64_t *)malloc(sizeof(int64_t) * (hw + 1));
    offset[0] = 0;
    for (size_t i = 0; i < hw; ++i) offset[i + 1] = offset[i] + per_ray[i];
#pragma omp parallel
    {
        float *w = (float *)malloc(sizeof(float) * (size_t)N);
        uint8_t *keep = (uint8_t *)malloc((size_t)N);
        float *place = (float *)malloc(sizeof(float) * 3 * (size_t)N);
#pragma omp for schedule(static)
        for (int pix = 0; pix < H * W; ++pix) {
            if (per_ray[pix] == 0) continue;
            float o[3], d[3];
            ray_of_pixel(Pinv, pix % W, pix / W, o, d);
            march_ray(o, d, t_one, N, origin, voxel_size, X, Y, Z, tsdf, thr, w, keep, place, NULL);
            float *row = rows + (size_t)offset[pix] * (size_t)(4 + C);
            for (int i = 0; i < N; ++i) {
                if (!keep[i]) continue;
                row[0] = place[3 * i + 0];
                row[1] = place[3 * i + 1];
                row[2] = place[3 * i + 2];
                row[3] = w[i];
                for (int c = 0; c < C; ++c) row[4 + c] = features[(size_t)c * hw + pix];
                row += 4 + C;
            }
        }
        free(w);
        free(keep);
        free(place);
    }
    free(offset);
}

/* rm.py:298-307: w /= mean(w); feat *= w; output rows [x,y,z, feat*w].  rows_in [M,4+C] -> rows_out [M,3+C].
 * The mean is accumulated in double (torch's cascade sum is accurate to ~1 ulp of the true mean; a
 * sequential fp32 sum over millions of rows is not). Returns the mean used. */
ORACLE_API float cnrma_oracle_normalize_rows(int64_t M, int C, const float *rows_in, float *rows_out) {
    double sum = 0.0;
    for (int64_t m = 0; m < M; ++m) sum += (double)rows_in[(size_t)m * (4 + C) + 3];
    float mean = (float)(sum / (double)M);
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        const float *r = rows_in + (size_t)m * (4 + C);
        float *q = rows_out + (size_t)m * (3 + C);
        float wn = r[3] / mean;
        q[0] = r[0];
        q[1] = r[1];
        q[2] = r[2];
        for (int c = 0; c < C; ++c) q[3 + c] = r[4 + c] * wn;
    }
    return mean;
}

/* rm.py:809-956 ray_projection_depth for one view: the first step where tsdf_i * tsdf_{i+1} <= 0
 * (rm.py:870-875), then 2k neighbours with triangular weights (k = select_grids > 0, rm.py:878-901) or the
 * half-step point itself (k == 0, rm.py:902-911).  Output rows [x,y,z,w,feat] in (v,u,j) order; returns M.
 * Pass rows == NULL to count only. */
ORACLE_API int64_t cnrma_oracle_depth_rows(const float *Pinv, int H, int W, int N, float t_one, const float *origin,
                                           float voxel_size, int X, int Y, int Z, const float *tsdf,
                                           int select_grids, int C, const float *features, float *rows) {
    size_t hw = (size_t)H * W;
    int64_t M = 0;
    float *tv = (float *)malloc(sizeof(float) * (size_t)N);
    for (int pix = 0; pix < H * W; ++pix) {
        float o[3], d[3];
        ray_of_pixel(Pinv, pix % W, pix / W, o, d);
        for (int i = 0; i < N; ++i) {
            float t = (float)i * t_one;
            int64_t id[3];
            for (int r = 0; r < 3; ++r) {
                float p = o[r] + d[r] * t;
                id[r] = f2l(rintf((p - origin[r]) / voxel_size));
            }
            int inb = (id[0] >= 0) & (id[0] < X) & (id[1] >= 0) & (id[1] < Y) & (id[2] >= 0) & (id[2] < Z);
            tv[i] = inb ? tsdf[((size_t)id[0] * Y + (size_t)id[1]) * Z + (size_t)id[2]] : 1.0f;
        }
        int best = -1;
        for (int i = 0; i + 1 < N; ++i)
            if (tv[i] * tv[i + 1] <= 0.0f) { best = i; break; }
        if (best < 0) continue; /* best_mask false -> weight 0 -> not selected */
        if (select_grids > 0) {
            int NUM = 2 * select_grids;
            for (int j = 0; j < NUM; ++j) {
                int idx = best + j - select_grids + 1;
                int tri = (j < select_grids) ? (j + 1) : (NUM - j);
                float wj = 1.0f * ((float)tri / (float)select_grids);
                if (idx < 0 || idx >= N) continue; /* selected_mask zeroes the weight -> not > 0 */
                if (!(wj > 0.0f)) continue;
                if (rows) {
                    float *row = rows + (size_t)M * (size_t)(4 + C);
                    /* selected_places = o + d * idx * t_one : (d * float(idx)) * t_one, rm.py:901 */
                    for (int r = 0; r < 3; ++r) row[r] = o[r] + (d[r] * (float)idx) * t_one;
                    row[3] = wj;
                    for (int c = 0; c < C; ++c) row[4 + c] = features[(size_t)c * hw + pix];
                }
                ++M;
            }
        } else {
            if (rows) {
                float *row = rows + (size_t)M * (size_t)(4 + C);
                float fi = (float)best + 0.5f;
                for (int r = 0; r < 3; ++r) row[r] = o[r] + (d[r] * fi) * t_one;
                row[3] = 1.0f;
                for (int c = 0; c < C; ++c) row[4 + c] = features[(size_t)c * hw + pix];
            }
            ++M;
        }
    }
    free(tv);
    return M;
}

/* Derived operator (SURVEY.md section 8a "dense_rma"): scatter the un-normalised rows of
 * ray_projection_neus into voxels by the voxel id the reference already computes (rm.py:730):
 *   wsum[c, vox] += w * feat[c];  wtot[vox] += w.   rows: [M,4+C].  Accumulated in double so the
 * oracle is order-independent; outputs fp32 [C,nvox] and [nvox]. */
ORACLE_API void cnrma_oracle_scatter_rows(int64_t M, int C, const float *rows, const float *origin,
                                          float voxel_size, int X, int Y, int Z, double *wsum, double *wtot) {
    size_t nvox = (size_t)X * Y * Z;
    for (int64_t m = 0; m < M; ++m) {
        const float *r = rows + (size_t)m * (4 + C);
        int64_t id[3];
        for (int k = 0; k < 3; ++k) id[k] = f2l(rintf((r[k] - origin[k]) / voxel_size));
        if (id[0] < 0 || id[0] >= X || id[1] < 0 || id[1] >= Y || id[2] < 0 || id[2] >= Z) continue;
        size_t vox = ((size_t)id[0] * Y + (size_t)id[1]) * Z + (size_t)id[2];
        wtot[vox] += (double)r[3];
        for (int c = 0; c < C; ++c) wsum[(size_t)c * nvox + vox] += (double)r[3] * (double)r[4 + c];
    }
}

/* GT TSDF fusion, one depth frame: data_prepare/scannet/tsdf.py:402-451
 * TSDFFusion.integrate, restated per voxel ("fz.py" below = that file).
 *   fz.py:413-416  camera = P @ [world;1]; px, py = round(x/z), round(y/z); pz = z     (same idiom as rm.py:51-54)
 *   fz.py:419-423  valid = frustum & depth[py,px] > 0
 *   fz.py:426-427  dist = clamp((pz - depth) / trunc_margin, min=-1)
 *   fz.py:430-433  valid &= dist < 1
 *   fz.py:436-437  weight == 0: tsdf = dist
 *   fz.py:440-445  near = dist > -1; weight != 0 & near: tsdf += dist; near: weight += 1
 *   fz.py:447-450  near: color += color_img[:, py, px]; label = label_img[py, px]
 * P is the frame's 3x4 projection (full resolution, not stride-scaled). color / label images and volumes may be NULL. */
ORACLE_API void cnrma_oracle_tsdf_integrate(int nx, int ny, int nz, float voxel_size, const float *origin,
                                            const float *P, int H, int W, const float *depth, const float *color_img,
                                            const int64_t *label_img, float trunc_margin, float *tsdf, float *weight,
                                            float *color, int64_t *label) {
    size_t nvox = (size_t)nx * ny * nz;
    size_t plane = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int x = 0; x < nx; ++x)
        for (int y = 0; y < ny; ++y)
            for (int z = 0; z < nz; ++z) {
                size_t i = ((size_t)x * ny + y) * nz + z;
                float wx = (float)x * voxel_size + origin[0];
                float wy = (float)y * voxel_size + origin[1];
                float wz = (float)z * voxel_size + origin[2];
                float cam[3];
                for (int r = 0; r < 3; ++r) {
                    float acc = P[4 * r + 0] * wx;
                    acc = fmaf(P[4 * r + 1], wy, acc);
                    acc = fmaf(P[4 * r + 2], wz, acc);
                    acc = fmaf(P[4 * r + 3], 1.0f, acc);
                    cam[r] = acc;
                }
                int64_t px = f2l(rintf(cam[0] / cam[2]));
                int64_t py = f2l(rintf(cam[1] / cam[2]));
                if (!((px >= 0) & (py >= 0) & (px < W) & (py < H) & (cam[2] > 0.0f))) continue;
                float d = depth[(size_t)py * W + (size_t)px];
                if (!(d > 0.0f)) continue;
                float dist = (cam[2] - d) / trunc_margin;
                if (dist < -1.0f) dist = -1.0f;          /* clamp(min=-1) */
                if (!(dist < 1.0f)) continue;
                int first = (weight[i] == 0.0f);
                if (first) tsdf[i] = dist;
                if (dist > -1.0f) {
                    if (!first) tsdf[i] = tsdf[i] + dist;
                    weight[i] = weight[i] + 1.0f;
                    if (color && color_img)
                        for (int c = 0; c < 3; ++c) color[(size_t)c * nvox + i] += color_img[(size_t)c * plane + (size_t)py * W + (size_t)px];
                    if (label && label_img) label[i] = label_img[(size_t)py * W + (size_t)px];
                }
            }
}

/* ---- TSDF head, one scale ------------------------------------------------------
 * models/atlas_head.py:38-52 for one decoder: tsdf = tanh(conv1x1x1(x)) * label_smoothing; with a previous scale,
 * prev = nearest x2 upsample; where !(|prev| < sparse_threshold): tsdf = sign(prev) * .999.
 * x planar [C, nvox] (NCDHW), weight [C], prev [nx/2, ny/2, nz/2] or NULL.  The channel dot product is accumulated
 * in double: torch's conv kernel sums in an unspecified order, so parity on this row is a tolerance (see the test),
 * with the mask compared outside a band around the threshold.  mask (optional) = |prev| < thr as 0/1. */
ORACLE_API void cnrma_oracle_tsdf_head_scale(int C, int nx, int ny, int nz, const float *x, const float *weight,
                                             const float *prev, float label_smoothing, float sparse_threshold,
                                             float *tsdf, uint8_t *mask) {
    size_t nvox = (size_t)nx * ny * nz;
    int py = ny / 2, pz = nz / 2;
#pragma omp parallel for schedule(static)
    for (int ix = 0; ix < nx; ++ix)
        for (int iy = 0; iy < ny; ++iy)
            for (int iz = 0; iz < nz; ++iz) {
                size_t i = ((size_t)ix * ny + iy) * nz + iz;
                double acc = 0.0;
                for (int c = 0; c < C; ++c) acc += (double)weight[c] * (double)x[(size_t)c * nvox + i];
                float t = (float)tanh((double)(float)acc) * label_smoothing;
                uint8_t m = 1;
                if (prev) {
                    float p = prev[((size_t)(ix / 2) * py + (iy / 2)) * pz + (iz / 2)];
                    m = fabsf(p) < sparse_threshold;
                    if (!m) t = (float)((p > 0.0f) - (p < 0.0f)) * 0.999f;
                }
                tsdf[i] = t;
                if (mask) mask[i] = m;
            }
}
