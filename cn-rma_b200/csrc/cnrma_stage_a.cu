// Stage A: dense back-projection -- voxel->pixel projection, frustum mask, nearest gather and the
// multi-view sum / mean, fused into one pass (reference: rm.py:21-69, :220-257).
//
// Design of the gather kernel (DESIGN.md "K_A"):
//   * persistent warps; one warp owns one voxel at a time and walks the views 32 at a time, lane <-> view:
//     every lane projects the voxel through its view (explicit FMA chain, IEEE division, rint -- bit-exact
//     with the reference), a ballot turns the frustum tests into a bit mask.
//   * each lane whose view sees the voxel issues ONE bulk asynchronous copy (TMA, cp.async.bulk
//     global -> shared, completion on the warp's mbarrier) of the pixel's channels-last feature row --
//     1 KB for 256 fp32 channels -- into the warp's row buffer, at the slot given by its rank among the
//     visible views.  No registers are tied up by loads in flight and a 1 KB gather costs one instruction
//     instead of 64 LDG.128; with 32 resident warps per SM up to ~190 KB of gathers are in flight per SM.
//   * when the buffer is full (or the views are exhausted) the warp waits on the mbarrier and adds the rows
//     from shared memory in slot order == view order, so the fp32 sum is formed exactly like the
//     reference's `self.volume + volume` loop (conflict-free LDS.128, lane <-> 4 channels).
//   * camera matrices (pre-divided by the backbone stride) and per-view base pointers sit in shared
//     memory, matrices as structure-of-arrays so that lane <-> view reads are conflict free.
//   * the count-mean divides by a small integer: one correctly rounded reciprocal per voxel and a
//     Markstein correction step per channel give the IEEE quotient (verified exhaustively by
//     cnrma_selftest_count_division) at a quarter of the instructions of eight generic divisions.
//   * blockIdx.y walks channel chunks (pass-major CTA order) when a feature row exceeds kMaxChunkBytes.
#include <cstdlib>

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
    int chunk_bytes;              // bytes of a feature row handled per pass (multiple of 16, <= kMaxChunkBytes)
    int rows_cap;                 // row slots per warp buffer
    SweepOrder sweep;             // traversal order of the voxels
    int bilinear;                 // opt-in variant: four neighbouring rows per visible view
    int reserve_ctas;             // host only: CTA slots left free for a kernel that runs beside this one
    int cull;                     // work unit = a column of the slab with per-column view culling (else one voxel)
    OutputRoute route;            // view-sharded output over peer memory (n_owners == 0: plain output)
    const void *views[kMaxViewsPerLaunch];
};

constexpr int kMaxChunkBytes = 1024;     // 64 16-byte vectors: two per lane

// bytes in front of the row buffers: mbarriers, per-view pointers, camera matrices [12][Vpad], cull coefficients
// [5][Vpad], per-warp candidate-view lists (uint16 [Vpad + 1] each)
__host__ __device__ inline size_t agg_head_bytes(int V) {
    const int Vpad = V | 1;
    const size_t warps = kAggThreads / kWarp;
    const size_t rounds = (V + 31) / 32;   // CTA-cooperative culling: two buffers of per-round survivor lists + counts
    return (8 * warps + sizeof(void *) * V + sizeof(float) * 17 * Vpad + sizeof(uint16_t) * warps * (Vpad + 1) +
            2 * rounds * (32 * sizeof(uint16_t) + sizeof(int)) + 8 + 127) & ~(size_t)127;
}
// head of the pipelined kernel's shared memory: mbarriers, per-view pointers, camera matrices, per-warp offset buffers
__host__ __device__ inline size_t pipe_head_bytes(int V) {
    const int Vpad = V | 1;
    const size_t warps = kAggThreads / kWarp;
    return (8 * warps + sizeof(void *) * V + sizeof(float) * 12 * Vpad + sizeof(int32_t) * warps * 2 * 2 * 32 + 127) & ~(size_t)127;
}

constexpr int kWarpBufferBytes = 6144;   // per-warp row buffer: 4 CTAs x 8 warps x 6 KB = 192 KB per SM
constexpr int kAggCtasPerSm = 4;

// BILINEAR (opt-in variant, cnrma_aggregate_views_bilinear): every visible view contributes its four neighbouring pixel
// rows -- four bulk copies into four consecutive slots, the interpolation weights beside them -- instead of the one
// nearest row; the validity mask and the count are the nearest path's.
template <int VPL, typename T, bool BILINEAR>
__global__ void __launch_bounds__(kAggThreads, BILINEAR ? 2 : kAggCtasPerSm)
aggregate_views_kernel(const __grid_constant__ AggParams p) {
    using V16 = Vec16<T>;
    constexpr int E = V16::kElems;   // channels per 16-byte vector
    constexpr int kWarps = kAggThreads / kWarp;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Vpad = p.V | 1;   // odd stride: conflict-free lane<->view reads
    uint64_t *sBar = reinterpret_cast<uint64_t *>(smem_raw);                                   // [kWarps]
    const unsigned char **sView = reinterpret_cast<const unsigned char **>(smem_raw + 8 * kWarps);   // [V]
    float *sP = reinterpret_cast<float *>(smem_raw + 8 * kWarps + sizeof(void *) * p.V);       // [12][Vpad]
    float *sCull = sP + 12 * Vpad;                                                             // [5][Vpad]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint16_t *sCand = reinterpret_cast<uint16_t *>(sCull + 5 * Vpad) + (size_t)warp * (Vpad + 1);   // [kWarps][Vpad + 1]
    const int cull_rounds = (p.V + 31) >> 5;
    uint16_t *sCandR = reinterpret_cast<uint16_t *>(sCull + 5 * Vpad) + (size_t)kWarps * (Vpad + 1);   // [2][rounds][32]
    int *sCnt = reinterpret_cast<int *>((reinterpret_cast<uintptr_t>(sCandR + 2 * cull_rounds * 32) + 3) & ~(uintptr_t)3);   // [2][rounds]
    const size_t head = agg_head_bytes(p.V);
    unsigned char *rowbuf = smem_raw + head + (size_t)warp * p.rows_cap * p.chunk_bytes;
    float *wbuf = reinterpret_cast<float *>(smem_raw + head + (size_t)kWarps * p.rows_cap * p.chunk_bytes) + warp * p.rows_cap;   // BILINEAR: weight per slot

    for (int i = threadIdx.x; i < 12 * p.V; i += blockDim.x) {
        const int v = i / 12, k = i % 12;
        float val = __ldg(p.proj + (int64_t)v * p.proj_stride + k);
        if (k < 8) val = __fdiv_rn(val, p.stride);   // rows 0-1 / stride (rm.py:238-239)
        sP[k * Vpad + v] = val;
    }
    const int chunk = p.chunk_base + blockIdx.y;
    for (int i = threadIdx.x; i < p.V; i += blockDim.x)
        sView[i] = static_cast<const unsigned char *>(p.views[i]) + (size_t)chunk * p.chunk_bytes;
    if (threadIdx.x < kWarps) mbar_init(smem_u32(&sBar[threadIdx.x]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    // View culling per voxel column (see the main loop): how fast the five frustum forms of rm.py:58 can grow along z,
    // per view, in world units -- |d/dz| of  cx + cz/2,  (W - 1/2) cz - cx,  cy + cz/2,  (H - 1/2) cz - cy,  cz.
    const float fW = (float)p.W - 0.5f, fH = (float)p.H - 0.5f;
    for (int v = threadIdx.x; v < p.V; v += blockDim.x) {
        const float px = sP[2 * Vpad + v], py = sP[6 * Vpad + v], pz = sP[10 * Vpad + v];
        sCull[0 * Vpad + v] = fabsf(px + 0.5f * pz);
        sCull[1 * Vpad + v] = fabsf(fW * pz - px);
        sCull[2 * Vpad + v] = fabsf(py + 0.5f * pz);
        sCull[3 * Vpad + v] = fabsf(fH * pz - py);
        sCull[4 * Vpad + v] = fabsf(pz);
    }
    __syncthreads();

    const uint32_t bar = smem_u32(&sBar[warp]);
    const uint32_t buf0 = smem_u32(rowbuf);
    const int nvec = p.chunk_bytes >> 4;
    const int c0 = chunk * (p.chunk_bytes / (int)sizeof(T));   // first channel of this pass
    const int esz = (int)sizeof(T);
    uint32_t parity = 0;

    // The voxels are visited in the order of SweepOrder (cnrma_common.cuh: slabs of T z-slices, x slow, then y, then the
    // slab's z); all resident warps sweep the volume together, so repeated gathers of a pixel's row fall inside the L2
    // residency window.  Two granularities (p.cull, chosen on the host):
    //   voxel units   a warp takes one voxel at a time and projects it through every view (lane <-> view, 32 at a time).
    //   column units  a warp takes one (x, y) column of a slab -- up to T voxels that differ in z only -- and first culls
    //                 the views: a conservative form of the frustum test of rm.py:58 on the UN-divided camera coordinates
    //                 of the column's centre, widened by how far each form can move over the column (sCull); a view
    //                 that fails it sees none of the column's voxels.  The exact projection then runs over the surviving
    //                 views only (in view order): for 50 ring cameras ~20 survive, one lane round instead of two.
    //                 Fewer instructions per voxel, but the voxels in flight span T times more columns: it pays on fine
    //                 grids, where a pixel's row is gathered ~20 times and the L2 window is lost anyway (cfg 4: 2.08 ->
    //                 1.89 ms), not where the window just fits (cfg 2: 0.286 -> 0.313 ms); profiles/r02_kernels.md.
    //   CTA columns   slabs of 8 slices; a CTA takes one column and its eight warps one voxel each, after culling the
    //                 views TOGETHER (each warp tests 32 views, one CTA barrier): the instruction saving of column units
    //                 with the voxels in flight as close together as with voxel units.
    const int warps_total = gridDim.x * kWarps;
    const int columns = p.g.nx * p.g.ny;
    const int nslabs = p.sweep.nfull + (p.sweep.nfull * p.sweep.T < p.g.nz ? 1 : 0);
    const float inv_columns = 1.0f / (float)columns;
    const int total_columns = columns * nslabs;
    // column units: every warp takes the same number of whole columns, round-robin; the columns left over after the last
    // full round are dealt out voxel by voxel (one-voxel units), so that no warp runs a whole column longer than the others
    const int full_columns = (p.cull == 1) ? (total_columns / warps_total) * warps_total : 0;
    const int total_units = full_columns + (total_columns - full_columns) * p.sweep.T;
    // CTA columns: the CTA walks the columns (slabs of kWarps slices), warp w takes the column's w-th voxel
    const bool cta_mode = p.cull == 2;
    const int u_first = cta_mode ? (int)blockIdx.x : (int)blockIdx.x * kWarps + warp;
    const int u_step = cta_mode ? (int)gridDim.x : warps_total;
    const int u_total = cta_mode ? total_columns : total_units;
    int round_robin = 0;   // CTA columns: which of the two shared survivor buffers this column uses
    for (int u = u_first; u < u_total; u += u_step, round_robin ^= 1) {
        int col = u, zfirst = 0, zcount = p.sweep.T;
        if (!cta_mode && u >= full_columns) {
            fast_divmod(u - full_columns, p.sweep.T, p.sweep.inv_T, col, zfirst);
            col += full_columns;
            zcount = 1;
        }
        int slab, xy, vx, vy;
        fast_divmod(col, columns, inv_columns, slab, xy);
        fast_divmod(xy, p.g.ny, p.sweep.inv_ny, vx, vy);
        int z0 = slab * p.sweep.T + zfirst;
        int nzu = min(zcount, p.g.nz - z0);     // <= 0: a one-voxel unit beyond the last (thinner) slab
        if (!cta_mode && nzu <= 0) continue;
        const float wx = world_coord(vx + p.g.x0, p.g.vs, p.g.ox);
        const float wy = world_coord(vy + p.g.y0, p.g.vs, p.g.oy);
        const float rho = 0.5f * (float)(nzu - 1) * p.g.vs;                          // half the column's length
        const float wzc = ((float)(z0 + p.g.z0) + 0.5f * (float)(nzu - 1)) * p.g.vs + p.g.oz;   // its centre (cull only)
        // lane <-> view: may this view see any voxel of the column?  (conservative; see above)
        auto column_may_be_seen = [&](int view) -> bool {
            if (view >= p.V) return false;
            const float *P = sP + view;
            const float cx = row_dot4(P[0 * Vpad], P[1 * Vpad], P[2 * Vpad], P[3 * Vpad], wx, wy, wzc, 1.0f);
            const float cy = row_dot4(P[4 * Vpad], P[5 * Vpad], P[6 * Vpad], P[7 * Vpad], wx, wy, wzc, 1.0f);
            const float cz = row_dot4(P[8 * Vpad], P[9 * Vpad], P[10 * Vpad], P[11 * Vpad], wx, wy, wzc, 1.0f);
            const float *Q = sCull + view;
            const float slack = 1.0e-3f * (fabsf(cz) + rho * Q[4 * Vpad]);      // rounding of the fp32 chains: ~1e-6
            return (cz + rho * Q[4 * Vpad] >= -slack) && (cx + 0.5f * cz + rho * Q[0 * Vpad] >= -slack) &&
                   (fW * cz - cx + rho * Q[1 * Vpad] >= -slack) && (cy + 0.5f * cz + rho * Q[2 * Vpad] >= -slack) &&
                   (fH * cz - cy + rho * Q[3 * Vpad] >= -slack);
        };
        int ncand = p.cull ? 0 : p.V;
        if (cta_mode) {
            // the warps share the cull: warp w tests views 32 w .. 32 w + 31 (+ 256 ...) and leaves the survivors of its
            // round in shared memory; after ONE CTA barrier every warp strings the rounds together into its own list.
            // Two buffers: a warp that is ahead writes the next column's survivors while a slower one still reads this
            // column's -- the barrier of the next column is what separates columns two apart.
            uint16_t *cand_r = sCandR + round_robin * cull_rounds * 32;
            int *cnt_r = sCnt + round_robin * cull_rounds;
            for (int r = warp; r < cull_rounds; r += kWarps) {
                const bool maybe = column_may_be_seen(r * 32 + lane);
                const unsigned mb = __ballot_sync(0xffffffffu, maybe);
                if (maybe) cand_r[r * 32 + __popc(mb & ((1u << lane) - 1u))] = (uint16_t)(r * 32 + lane);
                if (lane == 0) cnt_r[r] = __popc(mb);
            }
            __syncthreads();
            for (int r = 0; r < cull_rounds; ++r) {
                const int c = cnt_r[r];
                if (lane < c) sCand[ncand + lane] = cand_r[r * 32 + lane];
                ncand += c;
            }
            z0 += warp;                            // this warp's voxel of the column
            nzu = (warp < nzu) ? 1 : 0;
        } else if (p.cull) {
            for (int v0 = 0; v0 < p.V; v0 += 32) {
                const bool maybe = column_may_be_seen(v0 + lane);
                const unsigned mb = __ballot_sync(0xffffffffu, maybe);
                if (maybe) sCand[ncand + __popc(mb & ((1u << lane) - 1u))] = (uint16_t)(v0 + lane);
                ncand += __popc(mb);
            }
        }
        __syncwarp();
      for (int zi = 0; zi < nzu; ++zi) {
        const int vz = z0 + zi;
        // voxel order of datasets/tsdf.py:24-29: flat = (x*ny + y)*nz + z
        const int vox = (vx * p.g.ny + vy) * p.g.nz + vz;
        const float wz = world_coord(vz + p.g.z0, p.g.vs, p.g.oz);

        float acc[VPL][E];
        int cnt = 0;
        if (p.flags & CNRMA_AGG_ACCUMULATE) {
            cnt = (p.flags & CNRMA_AGG_COUNT_F32) ? (int)reinterpret_cast<const float *>(p.count)[vox] : p.count[vox];
#pragma unroll
            for (int k = 0; k < VPL; ++k)
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int c = c0 + (lane + 32 * k) * E + e;
                    acc[k][e] = (lane + 32 * k < nvec) ? p.volume[(int64_t)vox * p.vsv + (int64_t)c * p.vsc] : 0.0f;
                }
        } else {
#pragma unroll
            for (int k = 0; k < VPL; ++k)
#pragma unroll
                for (int e = 0; e < E; ++e) acc[k][e] = 0.0f;
        }

        int filled = 0;   // rows issued into the buffer and not yet added (warp-uniform)
        // adds the buffered rows in slot order (== view order) once their bytes have landed
        auto drain = [&]() {
            if (lane == 0) mbar_arrive(bar);
            mbar_wait(bar, parity);
            parity ^= 1u;
            const unsigned char *row = rowbuf + (lane << 4);
            if (BILINEAR) {
                __syncwarp();   // the weights were written by the issuing lanes
                for (int r = 0; r < filled; r += 4, row += 4 * p.chunk_bytes) {
                    const float w00 = wbuf[r], w10 = wbuf[r + 1], w01 = wbuf[r + 2], w11 = wbuf[r + 3];
#pragma unroll
                    for (int k = 0; k < VPL; ++k) {
                        if (lane + 32 * k < nvec) {
                            const V16 f00 = V16::load_shared(row + (k << 9));
                            const V16 f10 = V16::load_shared(row + p.chunk_bytes + (k << 9));
                            const V16 f01 = V16::load_shared(row + 2 * p.chunk_bytes + (k << 9));
                            const V16 f11 = V16::load_shared(row + 3 * p.chunk_bytes + (k << 9));
#pragma unroll
                            for (int e = 0; e < E; ++e)
                                acc[k][e] += w00 * f00.v[e] + w10 * f10.v[e] + w01 * f01.v[e] + w11 * f11.v[e];
                        }
                    }
                }
            } else {
                for (int r = 0; r < filled; ++r, row += p.chunk_bytes) {
#pragma unroll
                    for (int k = 0; k < VPL; ++k) {
                        if (lane + 32 * k < nvec) {
                            const V16 val = V16::load_shared(row + (k << 9));
#pragma unroll
                            for (int e = 0; e < E; ++e) acc[k][e] = __fadd_rn(acc[k][e], val.v[e]);
                        }
                    }
                }
            }
            filled = 0;
            __syncwarp();   // all lanes are done reading before the slots are overwritten
        };

        for (int v0 = 0; v0 < ncand; v0 += 32) {
            const int view = (v0 + lane < ncand) ? (p.cull ? (int)sCand[v0 + lane] : v0 + lane) : -1;   // ascending: slot order == view order
            int64_t off = -1;
            int64_t off4[4] = {0, 0, 0, 0};
            float w4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (view >= 0) {
                if (BILINEAR) {
                    const float *P = sP + view;
                    const float cx = row_dot4(P[0 * Vpad], P[1 * Vpad], P[2 * Vpad], P[3 * Vpad], wx, wy, wz, 1.0f);
                    const float cy = row_dot4(P[4 * Vpad], P[5 * Vpad], P[6 * Vpad], P[7 * Vpad], wx, wy, wz, 1.0f);
                    const float cz = row_dot4(P[8 * Vpad], P[9 * Vpad], P[10 * Vpad], P[11 * Vpad], wx, wy, wz, 1.0f);
                    const float fx = __fdiv_rn(cx, cz), fy = __fdiv_rn(cy, cz);
                    if (in_frustum(rintf(fx), rintf(fy), cz, p.H, p.W)) {   // the reference's mask (rm.py:58)
                        const float x0f = floorf(fx), y0f = floorf(fy);
                        const float ax = fx - x0f, ay = fy - y0f;
                        const int x0 = min(max((int)x0f, 0), p.W - 1), x1 = min(max((int)x0f + 1, 0), p.W - 1);
                        const int y0 = min(max((int)y0f, 0), p.H - 1), y1 = min(max((int)y0f + 1, 0), p.H - 1);
                        off = 0;
                        off4[0] = (y0 * p.stride_y + x0 * p.stride_x) * esz;
                        off4[1] = (y0 * p.stride_y + x1 * p.stride_x) * esz;
                        off4[2] = (y1 * p.stride_y + x0 * p.stride_x) * esz;
                        off4[3] = (y1 * p.stride_y + x1 * p.stride_x) * esz;
                        w4[0] = (1.0f - ax) * (1.0f - ay);
                        w4[1] = ax * (1.0f - ay);
                        w4[2] = (1.0f - ax) * ay;
                        w4[3] = ax * ay;
                    }
                } else {
                    int px, py;
                    if (project_voxel(sP + view, Vpad, wx, wy, wz, p.H, p.W, px, py))
                        off = (py * p.stride_y + px * p.stride_x) * esz;
                }
            }
            const unsigned bits = __ballot_sync(0xffffffffu, off >= 0);
            const int n = __popc(bits);
            const int rank = __popc(bits & ((1u << lane) - 1u));
            cnt += n;
            constexpr int kSlots = BILINEAR ? 4 : 1;   // buffer slots per visible view
            int done = 0;   // visible views of this round already issued
            while (done < n) {
                if (filled + kSlots > p.rows_cap) drain();
                const int take = min((p.rows_cap - filled) / kSlots, n - done);
                if (lane == 0) mbar_expect_tx(bar, (uint32_t)(take * kSlots * p.chunk_bytes));
                __syncwarp();
                if (off >= 0 && rank >= done && rank < done + take) {
                    const int slot = filled + (rank - done) * kSlots;
                    if (BILINEAR) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            bulk_g2s(buf0 + (uint32_t)((slot + q) * p.chunk_bytes), sView[view] + off4[q], (uint32_t)p.chunk_bytes, bar);
                            wbuf[slot + q] = w4[q];
                        }
                    } else {
                        bulk_g2s(buf0 + (uint32_t)(slot * p.chunk_bytes), sView[view] + off, (uint32_t)p.chunk_bytes, bar);
                    }
                }
                filled += take * kSlots;
                done += take;
            }
        }
        if (filled > 0) drain();

        if (p.flags & CNRMA_AGG_MEAN) {
            // rm.py:251: fp32 sum / int64 count -> IEEE division by float(count); 0 where count == 0
            const float n = (float)cnt;
            const float y = __frcp_rn(n);
#pragma unroll
            for (int k = 0; k < VPL; ++k)
#pragma unroll
                for (int e = 0; e < E; ++e) acc[k][e] = (cnt > 0) ? div_by_count(acc[k][e], n, y) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int j = lane + 32 * k;
            if (j < nvec) {
                const int c = c0 + j * E;
                if (p.vec_store) {
                    float *dst = (p.route.n_owners > 0 ? route_row(p.route, vox) : p.volume + (int64_t)vox * p.vsv) + c;
#pragma unroll
                    for (int e = 0; e < E; e += 4)
                        __stcs(reinterpret_cast<float4 *>(dst + e),
                               make_float4(acc[k][e], acc[k][e + 1], acc[k][e + 2], acc[k][e + 3]));
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) p.volume[(int64_t)vox * p.vsv + (int64_t)(c + e) * p.vsc] = acc[k][e];
                }
            }
        }
        if (lane == 0 && blockIdx.y == 0 && p.write_count) {
            if (p.route.n_owners > 0) route_row(p.route, vox)[p.C] = (float)cnt;
            else if (p.flags & CNRMA_AGG_COUNT_F32) reinterpret_cast<float *>(p.count)[vox] = (float)cnt;
            else p.count[vox] = cnt;
            if (p.valid != nullptr) p.valid[vox] = (uint8_t)(cnt > 0);
        }
      }            // voxels of the column
        __syncwarp();   // the candidate list is rewritten by the next column
    }
}

// ---- the same kernel, software-pipelined ------------------------------------------------------------
// Voxel units, nearest sampling, up to 64 views.  ncu on the kernel above (cfg 2): half of the warp samples are waiting --
// on the barrier of the bulk copies (30 %) and on fixed-latency dependencies (23 %) -- while the issue slots are 71 % busy:
// a warp's own instruction work (project the next voxel: ~300 of its ~1000 instructions) and its own memory wait happen
// one after the other.  Here a warp projects voxel k+1 right after it has issued the first batch of voxel k's gathers,
// (An experiment that is kept as an option, CNRMA_AGG_PIPE=1, because it did NOT pay: see run_aggregate.)
// i.e. while those rows are in flight; the per-lane pixel offsets of the projected voxel wait in shared memory (two
// parities x two lane rounds x 32 lanes of int32 per warp), so no registers are tied up.  Same arithmetic, same slot
// order == view order, same bits.
constexpr int kPipeRounds = 2;   // lane rounds (views / 32) the offsets buffer holds

template <int VPL, typename T>
__global__ void __launch_bounds__(kAggThreads, kAggCtasPerSm) aggregate_views_pipe_kernel(const __grid_constant__ AggParams p) {
    using V16 = Vec16<T>;
    constexpr int E = V16::kElems;
    constexpr int kWarps = kAggThreads / kWarp;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int Vpad = p.V | 1;
    uint64_t *sBar = reinterpret_cast<uint64_t *>(smem_raw);                                   // [kWarps]
    const unsigned char **sView = reinterpret_cast<const unsigned char **>(smem_raw + 8 * kWarps);   // [V]
    float *sP = reinterpret_cast<float *>(smem_raw + 8 * kWarps + sizeof(void *) * p.V);       // [12][Vpad]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    int32_t *sOff = reinterpret_cast<int32_t *>(sP + 12 * Vpad) + warp * (2 * kPipeRounds * 32);   // [kWarps][2][rounds][32]
    const size_t head = pipe_head_bytes(p.V);
    unsigned char *rowbuf = smem_raw + head + (size_t)warp * p.rows_cap * p.chunk_bytes;

    for (int i = threadIdx.x; i < 12 * p.V; i += blockDim.x) {
        const int v = i / 12, k = i % 12;
        float val = __ldg(p.proj + (int64_t)v * p.proj_stride + k);
        if (k < 8) val = __fdiv_rn(val, p.stride);   // rows 0-1 / stride (rm.py:238-239)
        sP[k * Vpad + v] = val;
    }
    const int chunk = p.chunk_base + blockIdx.y;
    for (int i = threadIdx.x; i < p.V; i += blockDim.x)
        sView[i] = static_cast<const unsigned char *>(p.views[i]) + (size_t)chunk * p.chunk_bytes;
    if (threadIdx.x < kWarps) mbar_init(smem_u32(&sBar[threadIdx.x]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t bar = smem_u32(&sBar[warp]);
    const uint32_t buf0 = smem_u32(rowbuf);
    const int nvec = p.chunk_bytes >> 4;
    const int c0 = chunk * (p.chunk_bytes / (int)sizeof(T));
    const int esz = (int)sizeof(T);
    const int rounds = (p.V + 31) >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t parity = 0;

    // projects sweep position `it` through every view: the lanes' byte offsets (or -1) go to sOff[slot], the ballots
    // of the (up to two) lane rounds come back; returns the flat voxel id
    auto project = [&](int it, int slot, unsigned &b0, unsigned &b1) -> int {
        int vx, vy, vz;
        sweep_voxel(p.sweep, it, vx, vy, vz);
        const float wx = world_coord(vx + p.g.x0, p.g.vs, p.g.ox);
        const float wy = world_coord(vy + p.g.y0, p.g.vs, p.g.oy);
        const float wz = world_coord(vz + p.g.z0, p.g.vs, p.g.oz);
        b0 = b1 = 0u;
        for (int r = 0; r < rounds; ++r) {
            const int view = r * 32 + lane;
            int off = -1;
            int px, py;
            if (view < p.V && project_voxel(sP + view, Vpad, wx, wy, wz, p.H, p.W, px, py))
                off = (int)((py * p.stride_y + px * p.stride_x) * esz);
            sOff[(slot * kPipeRounds + r) * 32 + lane] = off;
            const unsigned bits = __ballot_sync(0xffffffffu, off >= 0);
            if (r == 0) b0 = bits; else b1 = bits;
        }
        return (vx * p.g.ny + vy) * p.g.nz + vz;   // voxel order of datasets/tsdf.py:24-29
    };

    const int warps_total = gridDim.x * kWarps;
    int it = blockIdx.x * kWarps + warp;
    unsigned b0 = 0u, b1 = 0u;
    int vox = 0, slot = 0;
    if (it < p.nvox) vox = project(it, slot, b0, b1);
    while (it < p.nvox) {
        float acc[VPL][E];
        int cnt = __popc(b0) + __popc(b1);
        if (p.flags & CNRMA_AGG_ACCUMULATE) {
            cnt += (p.flags & CNRMA_AGG_COUNT_F32) ? (int)reinterpret_cast<const float *>(p.count)[vox] : p.count[vox];
#pragma unroll
            for (int k = 0; k < VPL; ++k)
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int c = c0 + (lane + 32 * k) * E + e;
                    acc[k][e] = (lane + 32 * k < nvec) ? p.volume[(int64_t)vox * p.vsv + (int64_t)c * p.vsc] : 0.0f;
                }
        } else {
#pragma unroll
            for (int k = 0; k < VPL; ++k)
#pragma unroll
                for (int e = 0; e < E; ++e) acc[k][e] = 0.0f;
        }

        // issues the next rows of the current voxel (at most rows_cap, in view order); returns how many
        int r_cur = 0, done = 0;   // lane round being issued, its visible views already issued
        auto issue_batch = [&]() -> int {
            int filled = 0;
            while (r_cur < rounds && filled < p.rows_cap) {
                const unsigned bits = r_cur == 0 ? b0 : b1;
                const int n = __popc(bits);
                if (done >= n) {
                    ++r_cur;
                    done = 0;
                    continue;
                }
                const int take = min(p.rows_cap - filled, n - done);
                if (lane == 0) mbar_expect_tx(bar, (uint32_t)(take * p.chunk_bytes));
                __syncwarp();
                const int off = sOff[(slot * kPipeRounds + r_cur) * 32 + lane];
                const int rank = __popc(bits & lt);
                if (off >= 0 && rank >= done && rank < done + take)
                    bulk_g2s(buf0 + (uint32_t)((filled + rank - done) * p.chunk_bytes), sView[r_cur * 32 + lane] + off,
                             (uint32_t)p.chunk_bytes, bar);
                filled += take;
                done += take;
            }
            return filled;
        };
        auto drain = [&](int filled) {
            if (lane == 0) mbar_arrive(bar);
            mbar_wait(bar, parity);
            parity ^= 1u;
            const unsigned char *row = rowbuf + (lane << 4);
            for (int r = 0; r < filled; ++r, row += p.chunk_bytes) {
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    if (lane + 32 * k < nvec) {
                        const V16 val = V16::load_shared(row + (k << 9));
#pragma unroll
                        for (int e = 0; e < E; ++e) acc[k][e] = __fadd_rn(acc[k][e], val.v[e]);
                    }
                }
            }
            __syncwarp();   // all lanes are done reading before the slots are overwritten
        };

        int filled = issue_batch();
        // while the first rows are in flight: the next voxel's projection
        const int it_next = it + warps_total;
        unsigned nb0 = 0u, nb1 = 0u;
        int vox_next = 0;
        if (it_next < p.nvox) vox_next = project(it_next, slot ^ 1, nb0, nb1);
        while (filled > 0) {
            drain(filled);
            filled = issue_batch();
        }

        if (p.flags & CNRMA_AGG_MEAN) {
            // rm.py:251: fp32 sum / int64 count -> IEEE division by float(count); 0 where count == 0
            const float n = (float)cnt;
            const float y = __frcp_rn(n);
#pragma unroll
            for (int k = 0; k < VPL; ++k)
#pragma unroll
                for (int e = 0; e < E; ++e) acc[k][e] = (cnt > 0) ? div_by_count(acc[k][e], n, y) : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int j = lane + 32 * k;
            if (j < nvec) {
                const int c = c0 + j * E;
                if (p.vec_store) {
                    float *dst = p.volume + (int64_t)vox * p.vsv + c;
#pragma unroll
                    for (int e = 0; e < E; e += 4)
                        __stcs(reinterpret_cast<float4 *>(dst + e),
                               make_float4(acc[k][e], acc[k][e + 1], acc[k][e + 2], acc[k][e + 3]));
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) p.volume[(int64_t)vox * p.vsv + (int64_t)(c + e) * p.vsc] = acc[k][e];
                }
            }
        }
        if (lane == 0 && blockIdx.y == 0 && p.write_count) {
            if (p.flags & CNRMA_AGG_COUNT_F32) reinterpret_cast<float *>(p.count)[vox] = (float)cnt;
            else p.count[vox] = cnt;
            if (p.valid != nullptr) p.valid[vox] = (uint8_t)(cnt > 0);
        }
        it = it_next;
        vox = vox_next;
        b0 = nb0;
        b1 = nb1;
        slot ^= 1;
    }
}

// ---- launch ------------------------------------------------------------------------------------

// Largest divisor of `row_bytes` that is a multiple of 16 and <= kMaxChunkBytes.
static int plan_chunk_bytes(int row_bytes, int max_chunk_bytes) {
    const int units = row_bytes / 16;
    int best = 1;
    for (int d = 1; d <= units; ++d)
        if (units % d == 0 && d * 16 <= max_chunk_bytes) best = d;
    return best * 16;
}

template <int VPL, typename T, bool BILINEAR = false>
static cudaError_t launch_agg(const AggParams &p, int chunks, cudaStream_t stream) {
    const size_t head = agg_head_bytes(p.V);
    const size_t smem = head + (size_t)(kAggThreads / kWarp) * p.rows_cap * (p.chunk_bytes + (BILINEAR ? sizeof(float) : 0));
    auto kernel = aggregate_views_kernel<VPL, T, BILINEAR>;
    // launch configuration cached per (kernel instantiation, device, smem size): the attribute / occupancy queries
    // cost more than the launch itself
    struct Cached { int dev = -1; size_t smem = 0; int ctas = 0; };
    static thread_local Cached cache;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (cache.dev != dev || cache.smem != smem) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kAggThreads, smem);
        if (err != cudaSuccess) return err;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        cache.dev = dev;
        cache.smem = smem;
        cache.ctas = sms * per_sm;
    }
    int persistent = cache.ctas - p.reserve_ctas;
    if (persistent < 1) persistent = 1;
    const int needed = (p.nvox + (kAggThreads / kWarp) - 1) / (kAggThreads / kWarp);   // at least one voxel per warp
    const dim3 grid(needed < persistent ? needed : persistent, chunks);
    kernel<<<grid, kAggThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int VPL, typename T>
static cudaError_t launch_pipe(const AggParams &p, int chunks, cudaStream_t stream) {
    const size_t smem = pipe_head_bytes(p.V) + (size_t)(kAggThreads / kWarp) * p.rows_cap * p.chunk_bytes;
    auto kernel = aggregate_views_pipe_kernel<VPL, T>;
    struct Cached { int dev = -1; size_t smem = 0; int ctas = 0; };
    static thread_local Cached cache;
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    if (cache.dev != dev || cache.smem != smem) {
        err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kAggThreads, smem);
        if (err != cudaSuccess) return err;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        cache.dev = dev;
        cache.smem = smem;
        cache.ctas = sms * per_sm;
    }
    int persistent = cache.ctas - p.reserve_ctas;
    if (persistent < 1) persistent = 1;
    const int needed = (p.nvox + (kAggThreads / kWarp) - 1) / (kAggThreads / kWarp);
    const dim3 grid(needed < persistent ? needed : persistent, chunks);
    kernel<<<grid, kAggThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

static cudaError_t run_aggregate(const AggParams &p_in, int dtype, int max_chunk_bytes, cudaStream_t stream) {
    const int esz = (dtype == CNRMA_BF16) ? 2 : 4;
    AggParams p = p_in;
    if (max_chunk_bytes <= 0 || max_chunk_bytes > kMaxChunkBytes) max_chunk_bytes = kMaxChunkBytes;
    p.chunk_bytes = plan_chunk_bytes(p.C * esz, max_chunk_bytes);
    const int chunks = (p.C * esz) / p.chunk_bytes;
    int warp_buffer = p.bilinear ? 2 * kWarpBufferBytes : kWarpBufferBytes;   // bilinear: 2 CTAs per SM, 12 KB per warp
    if (tuning().agg_warp_buffer > 0) warp_buffer = tuning().agg_warp_buffer;   // tuning aid (CNRMA_AGG_WARP_BUFFER)
    p.rows_cap = warp_buffer / p.chunk_bytes;
    if (p.rows_cap < 1) p.rows_cap = 1;
    if (p.rows_cap > 32) p.rows_cap = 32;
    if (p.bilinear) p.rows_cap = p.rows_cap < 4 ? 4 : (p.rows_cap & ~3);
    const int vpl = (p.chunk_bytes / 16 + 31) / 32;
    // voxel units, nearest sampling, up to 64 views, plain output: the software-pipelined form, opt-in (CNRMA_AGG_PIPE=1):
    // measured on cfg 2 at 0.304 ms against 0.286 ms for the plain kernel (0.284 with 8 KB warp buffers), cfg 4 2.22 vs 2.14
    const bool pipe = !p.bilinear && p.cull == 0 && p.V >= 1 && p.V <= 32 * kPipeRounds && p.route.n_owners == 0 &&
                      ((int64_t)p.H * p.stride_y + (int64_t)p.W * p.stride_x) * esz < ((int64_t)1 << 31) &&   // int32 byte offsets
                      tuning().agg_pipe == 1;
    auto launch = [&](const AggParams &q, int nchunks) -> cudaError_t {
        if (pipe) {
            if (dtype == CNRMA_BF16)
                return vpl == 1 ? launch_pipe<1, __nv_bfloat16>(q, nchunks, stream) : launch_pipe<2, __nv_bfloat16>(q, nchunks, stream);
            return vpl == 1 ? launch_pipe<1, float>(q, nchunks, stream) : launch_pipe<2, float>(q, nchunks, stream);
        }
        if (q.bilinear) {
            if (dtype == CNRMA_BF16)
                return vpl == 1 ? launch_agg<1, __nv_bfloat16, true>(q, nchunks, stream) : launch_agg<2, __nv_bfloat16, true>(q, nchunks, stream);
            return vpl == 1 ? launch_agg<1, float, true>(q, nchunks, stream) : launch_agg<2, float, true>(q, nchunks, stream);
        }
        if (dtype == CNRMA_BF16)
            return vpl == 1 ? launch_agg<1, __nv_bfloat16>(q, nchunks, stream) : launch_agg<2, __nv_bfloat16>(q, nchunks, stream);
        return vpl == 1 ? launch_agg<1, float>(q, nchunks, stream) : launch_agg<2, float>(q, nchunks, stream);
    };
    p.chunk_base = 0;
    p.write_count = 1;
    if (!(p.flags & CNRMA_AGG_ACCUMULATE) || chunks == 1) return launch(p, chunks);
    // Accumulating launches read the previous count; with several channel chunks in one grid the chunk
    // that rewrites it would race with the others, so the chunks go out one launch at a time (stream
    // ordered) and only the last one stores the new count.
    for (int c = 0; c < chunks; ++c) {
        p.chunk_base = c;
        p.write_count = (c == chunks - 1);
        const cudaError_t err = launch(p, 1);
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

// Views [v0, v0 + nv) of `f` in one launch (nv <= kMaxViewsPerLaunch).
cudaError_t run_aggregate_views(const GridDev &g, const cnrma_features &f, int v0, int nv, const float *proj,
                                int64_t proj_stride, float stride, uint32_t flags, float *volume, int64_t vsv,
                                int64_t vsc, int32_t *count, uint8_t *valid, int max_chunk_bytes, cudaStream_t stream,
                                const OutputRoute *route, int reserve_ctas) {
    // Kernel choice (DESIGN.md "K_A"): rows of 512 bytes and more go through the TMA kernel below (DRAM-bound, deep
    // register-free gather queue); shorter rows -- and the finalise-only pass -- through the list kernel
    // (cnrma_stage_a_list.cu), whose lane <-> voxel projection needs a third of the instructions.
    const int row_bytes = f.channels * ((f.dtype == CNRMA_BF16) ? 2 : 4);
    // (beyond ~96 views the per-voxel lists no longer fit 32 voxels per warp and the list kernel loses its edge)
    bool use_list = ((row_bytes < 512 && (nv <= kListViewsMax || (route == nullptr && long_list_supports(f.channels, f.dtype)))) ||
                     nv == 0) && !(flags & kAggBilinearInternal);
    if (tuning().agg_kernel >= 0) use_list = (tuning().agg_kernel == 1) && !(flags & kAggBilinearInternal);   // CNRMA_AGG_KERNEL
    if (use_list && list_kernel_supports(nv, f.height, f.width))
        return run_aggregate_list(g, f, v0, nv, proj, proj_stride, stride, flags, volume, vsv, vsc, count, valid, stream, route,
                                  reserve_ctas);
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
    p.route.n_owners = 0;
    if (route != nullptr) {
        p.route = *route;
        p.vec_store = 1;   // routed rows are 16-byte aligned by construction (cnrma_aggregate_views_routed)
    }
    p.chunk_base = 0;
    p.write_count = 1;
    for (int i = 0; i < nv; ++i) p.views[i] = f.view_ptrs_host[v0 + i];
    p.chunk_bytes = 0;
    p.rows_cap = 0;
    p.bilinear = (flags & kAggBilinearInternal) ? 1 : 0;
    p.reserve_ctas = reserve_ctas > 0 ? reserve_ctas : 0;
    p.flags = flags & ~kAggBilinearInternal;
    p.sweep = make_sweep(g.nx, g.ny, g.nz, sweep_thickness(g.ny, g.nz, nv, row_bytes));
    // column units + view culling where every pixel row is gathered many times over (see the kernel): about a quarter of
    // the views see a voxel, so a row is gathered ~nvox / (4 H W) times
    p.cull = (double)p.nvox >= 32.0 * f.height * f.width;
    if (tuning().agg_cull >= 0) p.cull = tuning().agg_cull;   // CNRMA_AGG_CULL = 0 | 1 | 2
    if (p.cull == 2) p.sweep = make_sweep(g.nx, g.ny, g.nz, kAggThreads / kWarp);   // CTA columns: one slice per warp
    return run_aggregate(p, f.dtype, max_chunk_bytes, stream);
}

// ---- exhaustive check of div_by_count ---------------------------------------------------------------
// For every count n in [1, max_n] and every fp32 significand (one full binade, both signs, plus two far
// binades: scaling by a power of two is exact, so one binade stands for all normal values inside the guarded
// range), compares div_by_count against __fdiv_rn.  Returns the number of mismatches (must be 0).
__global__ void __launch_bounds__(256) selftest_count_division_kernel(int max_n, unsigned long long *mismatches) {
    const uint32_t sig = blockIdx.x * blockDim.x + threadIdx.x;   // 2^23 significands
    if (sig >= (1u << 23)) return;
    const int n_lo = blockIdx.y * 16 + 1;
    unsigned long long bad = 0;
    for (int ni = n_lo; ni < n_lo + 16 && ni <= max_n; ++ni) {
        const float n = (float)ni;
        const float y = __frcp_rn(n);
        const uint32_t exps[3] = {127u, 127u - 40u, 127u + 40u};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float a = __uint_as_float((exps[k] << 23) | sig);
            bad += (__float_as_uint(div_by_count(a, n, y)) != __float_as_uint(__fdiv_rn(a, n)));
            bad += (__float_as_uint(div_by_count(-a, n, y)) != __float_as_uint(__fdiv_rn(-a, n)));
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t run_selftest_count_division(int max_n, unsigned long long *mismatches, cudaStream_t stream) {
    cudaError_t err = cudaMemsetAsync(mismatches, 0, sizeof(unsigned long long), stream);
    if (err != cudaSuccess) return err;
    const dim3 grid((1u << 23) / 256, (max_n + 15) / 16);
    selftest_count_division_kernel<<<grid, 256, 0, stream>>>(max_n, mismatches);
    return cudaGetLastError();
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
    const bool ok = project_voxel(sP, 1, world_coord(vx + g.x0, g.vs, g.ox), world_coord(vy + g.y0, g.vs, g.oy),
                                  world_coord(vz + g.z0, g.vs, g.oz), H, W, px, py);
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
// Owner-side pass of the view-sharded Stage A: adds the partial sums and counts that the `n_src` sources stored for
// this owner's voxels (source order, so the result does not depend on timing) and divides by the total count.
// recv [n_src][slab][row_floats] (sums at [0, C), count at [C]); volume [rows, C]; count int32 [rows]; valid uint8.
__global__ void __launch_bounds__(256) finalize_routed_kernel(const float *__restrict__ recv, int n_src, int slab,
                                                              int row_floats, int rows, int C, int mean,
                                                              float *__restrict__ volume, int32_t *__restrict__ count,
                                                              uint8_t *__restrict__ valid) {
    const int nvec = C / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * nvec) return;
    const int row = (int)(i / nvec), j = (int)(i % nvec);
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    float n = 0.0f;
    for (int s = 0; s < n_src; ++s) {
        const float *r = recv + ((int64_t)s * slab + row) * row_floats;
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(r + 4 * j));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        n += __ldg(r + C);
    }
    if (mean) {
        const bool seen = n > 0.0f;
        acc.x = seen ? __fdiv_rn(acc.x, n) : 0.0f;
        acc.y = seen ? __fdiv_rn(acc.y, n) : 0.0f;
        acc.z = seen ? __fdiv_rn(acc.z, n) : 0.0f;
        acc.w = seen ? __fdiv_rn(acc.w, n) : 0.0f;
    }
    __stcs(reinterpret_cast<float4 *>(volume + (int64_t)row * C + 4 * j), acc);
    if (j == 0) {
        count[row] = (int32_t)n;
        if (valid != nullptr) valid[row] = (uint8_t)(n > 0.0f);
    }
}

cudaError_t run_finalize_routed(const float *recv, int n_src, int slab, int row_floats, int rows, int C, int mean,
                                float *volume, int32_t *count, uint8_t *valid, cudaStream_t stream) {
    const int64_t threads = (int64_t)rows * (C / 4);
    if (threads == 0) return cudaSuccess;
    finalize_routed_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(recv, n_src, slab, row_floats, rows, C, mean,
                                                                             volume, count, valid);
    return cudaGetLastError();
}

// Reference-layout (NCHW) maps -> channels-last rows, all views of a stack in one launch: tiles of 32 channels x 64
// pixels through shared memory, reads coalesced along the pixel axis of the source planes (eight loads in flight per
// thread), writes coalesced along the channel axis of the destination rows.  Contiguous planes (the usual case) are
// addressed linearly; other strides go through the division by W.
struct ChannelsLastParams {
    int C, H, W;
    int64_t sc, sy, sx;
    int linear;                 // sx == 1 && sy == W: pixel offset == pixel index
    unsigned char *dst;
    size_t per_view_bytes;
    const void *views[kMaxViewsPerLaunch];
};

template <typename T>
__global__ void __launch_bounds__(256) to_channels_last_kernel(const __grid_constant__ ChannelsLastParams p) {
    __shared__ T tile[32][65];
    const int hw = p.H * p.W;
    const int p0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const T *__restrict__ src = static_cast<const T *>(p.views[blockIdx.z]);
    T *__restrict__ dst = reinterpret_cast<T *>(p.dst + (size_t)blockIdx.z * p.per_view_bytes);
    T v[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = c0 + ty + 8 * i, pix = p0 + tx + 32 * h;
            v[i][h] = T(0.0f);
            if (c < p.C && pix < hw) {
                const int64_t off = p.linear ? (int64_t)pix : ((int64_t)(pix / p.W) * p.sy + (int64_t)(pix % p.W) * p.sx);
                v[i][h] = src[(int64_t)c * p.sc + off];
            }
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) tile[ty + 8 * i][tx + 32 * h] = v[i][h];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int pix = p0 + ty + 8 * i, c = c0 + tx;
        if (c < p.C && pix < hw) dst[(int64_t)pix * p.C + c] = tile[tx][ty + 8 * i];
    }
}

// src: `views` maps with common strides (host array of device pointers) -> dst [views, H, W, C] contiguous.
cudaError_t run_to_channels_last(const void *const *views_host, int views, int dtype, int C, int H, int W, int64_t sc,
                                 int64_t sy, int64_t sx, void *dst, cudaStream_t stream) {
    ChannelsLastParams p;
    p.C = C; p.H = H; p.W = W;
    p.sc = sc; p.sy = sy; p.sx = sx;
    p.linear = (sx == 1 && sy == W);
    p.per_view_bytes = (size_t)H * W * C * (dtype == CNRMA_BF16 ? 2 : 4);
    for (int v0 = 0; v0 < views; v0 += kMaxViewsPerLaunch) {
        const int nv = (views - v0 < kMaxViewsPerLaunch) ? (views - v0) : kMaxViewsPerLaunch;
        for (int i = 0; i < nv; ++i) p.views[i] = views_host[v0 + i];
        p.dst = static_cast<unsigned char *>(dst) + (size_t)v0 * p.per_view_bytes;
        const dim3 grid((H * W + 63) / 64, (C + 31) / 32, nv);
        if (dtype == CNRMA_BF16) to_channels_last_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
        else to_channels_last_kernel<float><<<grid, 256, 0, stream>>>(p);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace cnrma
