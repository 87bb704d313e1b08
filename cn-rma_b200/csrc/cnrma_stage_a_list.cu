// Stage A, list kernel: the same fused back-projection as cnrma_stage_a.cu (bit-identical results), organised for
// instruction economy.  It serves feature rows below 512 bytes -- the reference's own 32-channel maps are 128-byte
// rows -- and scenes with many more voxels than pixels, where the TMA kernel's per-voxel projection rounds and
// per-row issue / drain overhead, not DRAM, set the time (ref test config, 256x256x96 x 50 views x 32 ch: 6.5 ms).
//
//   phase 1  lane <-> voxel: a warp takes a batch of up to 32 consecutive voxels of the sweep order (slabs of z-slices, cnrma_common.cuh) and walks the
//            views in order; the camera matrix of the current view is warp-uniform (broadcast shared-memory reads),
//            every lane projects its own voxel and appends the pixel of each view that sees it to its voxel's list in
//            shared memory.  One projection serves 32 voxels: ~1.2 warp instructions per (voxel, view) instead of ~3.
//   phase 2  lane group <-> voxel: groups of G = row_bytes/16/VPL lanes walk the lists (kept in view order, so the
//            fp32 sums are formed exactly like the reference's `self.volume + volume` loop), 32/G voxels at a time,
//            with coalesced 16-byte loads of the channels-last rows, U rows in flight per group before the first add.
#include <cstdlib>

#include "cnrma_internal.cuh"

namespace cnrma {

struct ListParams {
    GridDev g;
    int V, C, H, W, nvox;
    int64_t stride_y_bytes, stride_x_bytes;
    float stride;
    const float *proj;
    int64_t proj_stride;
    float *volume;
    int64_t vsv, vsc;
    int32_t *count;
    uint8_t *valid;
    uint32_t flags;
    int vec_store;
    int chunk_base, write_count;
    int nb;     // voxels per warp batch (<= 32)
    int lcap;   // list capacity per voxel (odd, >= V)
    SweepOrder sweep;   // traversal order of the voxels (cnrma_common.cuh)
    OutputRoute route;  // view-sharded output over peer memory (n_owners == 0: plain output)
    // uniform: the views are equally spaced in one allocation, so a list entry is the row's offset from views[0] in
    // 16-byte units (one multiply-add to decode); otherwise entries pack (view, py, px) and go through the pointer table
    int uniform;
    int reserve_ctas;   // host only: CTA slots left free for a kernel that runs beside this one
    uint32_t view_stride16, stride_y16, stride_x16;
    const void *views[kMaxViewsPerLaunch];
};

constexpr int kListThreads = 256;
constexpr int kListUnroll = 4;
#ifndef CNRMA_LONG_CAP
#define CNRMA_LONG_CAP 161
#endif
#ifndef CNRMA_LONG_MINB
#define CNRMA_LONG_MINB 3   // 3 CTAs of 80 registers per SM (4 x 64 registers: 21.9 ms on cfg 5 against 19.1; 6 or 8 rows in flight at 105-116 registers: 21.3 / 20.8)
#endif
constexpr int kLongListCap = CNRMA_LONG_CAP;   // entries per voxel list of the long-list kernel (odd); longer lists go out in segments

template <int G, int VPL, typename T, bool UNIFORM>
__global__ void __launch_bounds__(kListThreads) aggregate_views_list_kernel(const __grid_constant__ ListParams p) {
    using V16 = Vec16<T>;
    constexpr int E = V16::kElems;
    constexpr int kWarps = kListThreads / kWarp;
    constexpr int kVPW = kWarp / G;   // voxels gathered concurrently by one warp

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sP = reinterpret_cast<float *>(smem_raw);                                                    // [V][12]
    const unsigned char **sView = reinterpret_cast<const unsigned char **>(smem_raw + sizeof(float) * 12 * p.V);   // [V]
    uint32_t *sList = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float) * 12 * p.V + sizeof(void *) * p.V);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *my_lists = sList + (size_t)warp * p.nb * p.lcap;

    for (int i = threadIdx.x; i < 12 * p.V; i += blockDim.x) {
        const int v = i / 12, k = i % 12;
        float val = __ldg(p.proj + (int64_t)v * p.proj_stride + k);
        if (k < 8) val = __fdiv_rn(val, p.stride);   // rows 0-1 / stride (rm.py:238-239)
        sP[i] = val;
    }
    const int chunk = p.chunk_base + blockIdx.y;
    constexpr int kChunkBytes = G * VPL * 16;
    for (int i = threadIdx.x; i < p.V; i += blockDim.x)
        sView[i] = static_cast<const unsigned char *>(p.views[i]) + (size_t)chunk * kChunkBytes;
    __syncthreads();

    const int grp = lane / G, lig = lane % G;
    const unsigned char *view0 = (p.V > 0) ? sView[0] : nullptr;
    const int c0 = chunk * (kChunkBytes / (int)sizeof(T)) + lig * E;   // first channel of this lane (+ q*G*E)
    const int units = (p.nvox + p.nb - 1) / p.nb;
    const int warps_total = gridDim.x * kWarps;
    const float fW = (float)p.W - 0.5f, fH = (float)p.H - 0.5f;

    for (int u = blockIdx.x * kWarps + warp; u < units; u += warps_total) {
        // ---- phase 1: lane <-> voxel -------------------------------------------------------------------------
        const int it = u * p.nb + lane;                       // index in the sweep order
        const bool active = lane < p.nb && it < p.nvox;
        int vx, vy, vz;
        sweep_voxel(p.sweep, active ? it : 0, vx, vy, vz);
        const int vox = (vx * p.g.ny + vy) * p.g.nz + vz;    // voxel order of datasets/tsdf.py:24-29
        const float wx = world_coord(vx + p.g.x0, p.g.vs, p.g.ox);
        const float wy = world_coord(vy + p.g.y0, p.g.vs, p.g.oy);
        const float wz = world_coord(vz + p.g.z0, p.g.vs, p.g.oz);
        uint32_t *lst = my_lists + lane * p.lcap;
        int cnt = 0;
        for (int v = 0; v < p.V; ++v) {
            const float4 a = *reinterpret_cast<const float4 *>(sP + 12 * v);       // warp-uniform: broadcast reads
            const float4 b = *reinterpret_cast<const float4 *>(sP + 12 * v + 4);
            const float4 c = *reinterpret_cast<const float4 *>(sP + 12 * v + 8);
            const float cx = row_dot4(a.x, a.y, a.z, a.w, wx, wy, wz, 1.0f);
            const float cy = row_dot4(b.x, b.y, b.z, b.w, wx, wy, wz, 1.0f);
            const float cz = row_dot4(c.x, c.y, c.z, c.w, wx, wy, wz, 1.0f);
            // Cheap superset of the frustum test on the un-divided coordinates: a visible voxel has cz > 0 and
            // -0.5 <= cx/cz <= W-0.5, -0.5 <= cy/cz <= H-0.5 up to the rounding of the quotient, i.e.
            // cx + 0.5 cz >= 0, (W-0.5) cz - cx >= 0, ... up to ~1e-7 cz; the slack used here is 1e-3 cz.  Neighbouring
            // voxels mostly agree, so when no lane of the warp passes, the division / rounding / exact test is skipped.
            const float slack = 1.0e-3f * cz;
            const bool maybe = active && (cz > 0.0f) && (cx + 0.5f * cz >= -slack) && (fW * cz - cx >= -slack) &&
                               (cy + 0.5f * cz >= -slack) && (fH * cz - cy >= -slack);
            if (!__any_sync(0xffffffffu, maybe)) continue;
            float rx, ry;
            rounded_pixel(cx, cy, cz, rx, ry);
            if (active && in_frustum(rx, ry, cz, p.H, p.W))
                lst[cnt++] = UNIFORM ? ((uint32_t)v * p.view_stride16 + (uint32_t)(int)ry * p.stride_y16 + (uint32_t)(int)rx * p.stride_x16)
                                     : (((uint32_t)v << 20) | ((uint32_t)(int)ry << 10) | (uint32_t)(int)rx);
        }
        __syncwarp();

        // ---- phase 2: lane group <-> voxel ---------------------------------------------------------------------
        for (int b0 = 0; b0 < p.nb; b0 += kVPW) {
            const int j = b0 + grp;                                       // voxel (lane of phase 1) of this group
            const int jcnt = __shfl_sync(0xffffffffu, cnt, j & 31);
            const int jvox = __shfl_sync(0xffffffffu, vox, j & 31);
            const bool jact = __shfl_sync(0xffffffffu, (int)active, j & 31) != 0 && j < p.nb;
            const int n = jact ? jcnt : 0;
            const uint32_t *jl = my_lists + (j < p.nb ? j : 0) * p.lcap;
            float acc[VPL][E];
            int total = n;
            if ((p.flags & CNRMA_AGG_ACCUMULATE) && jact) {
                total += (p.flags & CNRMA_AGG_COUNT_F32) ? (int)reinterpret_cast<const float *>(p.count)[jvox] : p.count[jvox];
#pragma unroll
                for (int q = 0; q < VPL; ++q)
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        acc[q][e] = p.volume[(int64_t)jvox * p.vsv + (int64_t)(c0 + q * G * E + e) * p.vsc];
            } else {
#pragma unroll
                for (int q = 0; q < VPL; ++q)
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[q][e] = 0.0f;
            }
            int nmax = n;   // longest list among the groups of this warp
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
            for (int k = 0; k < nmax; k += kListUnroll) {
                uint4 raw[kListUnroll][VPL];
#pragma unroll
                for (int uu = 0; uu < kListUnroll; ++uu) {
                    if (k + uu < n) {
                        const uint32_t e = jl[k + uu];
                        const unsigned char *src =
                            UNIFORM ? (view0 + (int64_t)e * 16 + lig * 16)
                                    : (sView[e >> 20] + (int64_t)((e >> 10) & 1023u) * p.stride_y_bytes +
                                       (int64_t)(e & 1023u) * p.stride_x_bytes + lig * 16);
#pragma unroll
                        for (int q = 0; q < VPL; ++q) raw[uu][q] = V16::load_raw(reinterpret_cast<const T *>(src + q * G * 16));
                    }
                }
#pragma unroll
                for (int uu = 0; uu < kListUnroll; ++uu) {
                    if (k + uu < n) {
#pragma unroll
                        for (int q = 0; q < VPL; ++q) {
                            const V16 val = V16::widen(raw[uu][q]);
#pragma unroll
                            for (int e = 0; e < E; ++e) acc[q][e] = __fadd_rn(acc[q][e], val.v[e]);
                        }
                    }
                }
            }
            if (!jact) continue;
            if (p.flags & CNRMA_AGG_MEAN) {
                const float fn = (float)total;   // rm.py:251: fp32 sum / int64 count, 0 where count == 0
                const float y = __frcp_rn(fn);
#pragma unroll
                for (int q = 0; q < VPL; ++q)
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[q][e] = (total > 0) ? div_by_count(acc[q][e], fn, y) : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < VPL; ++q) {
                const int c = c0 + q * G * E;
                if (p.vec_store) {
                    float *dst = (p.route.n_owners > 0 ? route_row(p.route, jvox) : p.volume + (int64_t)jvox * p.vsv) + c;
#pragma unroll
                    for (int e = 0; e < E; e += 4)
                        __stcs(reinterpret_cast<float4 *>(dst + e), make_float4(acc[q][e], acc[q][e + 1], acc[q][e + 2], acc[q][e + 3]));
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) p.volume[(int64_t)jvox * p.vsv + (int64_t)(c + e) * p.vsc] = acc[q][e];
                }
            }
            if (lig == 0 && blockIdx.y == 0 && p.write_count) {
                if (p.route.n_owners > 0) route_row(p.route, jvox)[p.C] = (float)total;
                else if (p.flags & CNRMA_AGG_COUNT_F32) reinterpret_cast<float *>(p.count)[jvox] = (float)total;
                else p.count[jvox] = total;
                if (p.valid != nullptr) p.valid[jvox] = (uint8_t)(total > 0);
            }
        }
        __syncwarp();   // the lists are rewritten by the next unit
    }
}

// ---- long view lists -----------------------------------------------------------------------------------------------
// Beyond ~96 views the per-voxel lists above no longer fit 32 voxels per warp; splitting the views into batches that
// accumulate through the volume in HBM (one read-modify-write of every voxel row per batch, and a DRAM-latency load in
// front of every gather round) cost cfg 5 (300 views, 256-byte bf16 rows, 6.3 M voxels) 25.6 ms in five batches.
// This kernel walks ALL views of a unit in one go:
//   * a unit is 8 consecutive voxels of the sweep; phase 1 maps lane <-> (voxel, one of 4 consecutive views), so one
//     warp instruction still projects 32 (voxel, view) pairs; the camera matrices of the four views come from four
//     conflict-free shared-memory broadcasts;
//   * lists of `lcap` entries per voxel (view order kept: the four lanes of a voxel append in lane order); the usual
//     scene fills well under half of that.  A unit whose list fills up is served in SEGMENTS: the sums gathered so far
//     go to the volume un-averaged and are read back when the next segment's gathers are added -- the same fp32 chain,
//     paid only by the voxels that need it;
//   * phase 2 as above (lane group <-> voxel, kListUnroll rows in flight), sums held in registers from the first view to
//     the last, divided once, written once.
// cfg 5: 25.6 -> 19.0 ms (three CTAs of 80 registers per SM; forcing four CTAs with 64 registers and lists of 129 entries
// was measured: 21.9 ms).  The kernel issues 1860 warp instructions per voxel -- 300 projections and ~72 gathered rows
// whose bf16 halves are widened one by one -- at 57 % of the issue slots: instruction-bound with the rest lost to L2
// latency (profiles/r02_kernels.md).
constexpr int kLongVoxels = 8;     // voxels per unit
#ifndef CNRMA_LONG_UNROLL
#define CNRMA_LONG_UNROLL 4
#endif
constexpr int kLongUnroll = CNRMA_LONG_UNROLL;   // rows in flight per lane group in phase 2
constexpr int kLongStep = 4;       // views per phase-1 step (lane = sub * 8 + voxel)

template <int G, int VPL, typename T, bool UNIFORM>
__global__ void __launch_bounds__(kListThreads, (VPL == 1) ? CNRMA_LONG_MINB : 1) aggregate_views_long_kernel(const __grid_constant__ ListParams p) {
    using V16 = Vec16<T>;
    constexpr int E = V16::kElems;
    constexpr int kWarps = kListThreads / kWarp;
    constexpr int kVPW = kWarp / G;   // voxels gathered concurrently by one warp (<= 8)
    static_assert(kVPW <= kLongVoxels && kLongVoxels % kVPW == 0, "lane groups must tile the unit");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sP = reinterpret_cast<float *>(smem_raw);                                                    // [V][12]
    const unsigned char **sView = reinterpret_cast<const unsigned char **>(smem_raw + sizeof(float) * 12 * p.V);   // [V]
    uint32_t *sList = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float) * 12 * p.V + sizeof(void *) * p.V);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *my_lists = sList + (size_t)warp * kLongVoxels * p.lcap;

    for (int i = threadIdx.x; i < 12 * p.V; i += blockDim.x) {
        const int v = i / 12, k = i % 12;
        float val = __ldg(p.proj + (int64_t)v * p.proj_stride + k);
        if (k < 8) val = __fdiv_rn(val, p.stride);   // rows 0-1 / stride (rm.py:238-239)
        sP[i] = val;
    }
    const int chunk = p.chunk_base + blockIdx.y;
    constexpr int kChunkBytes = G * VPL * 16;
    for (int i = threadIdx.x; i < p.V; i += blockDim.x)
        sView[i] = static_cast<const unsigned char *>(p.views[i]) + (size_t)chunk * kChunkBytes;
    __syncthreads();

    const int grp = lane / G, lig = lane % G;
    const int vj = lane & (kLongVoxels - 1), sub = lane >> 3;            // phase 1: voxel of the unit, view slot
    const unsigned same = 0x01010101u << vj;                            // the four lanes of my voxel
    const unsigned below = same & ((1u << lane) - 1u);
    const unsigned char *view0 = (p.V > 0) ? sView[0] : nullptr;
    const int c0 = chunk * (kChunkBytes / (int)sizeof(T)) + lig * E;
    const int units = (p.nvox + kLongVoxels - 1) / kLongVoxels;
    const int warps_total = gridDim.x * kWarps;
    const float fW = (float)p.W - 0.5f, fH = (float)p.H - 0.5f;
    const bool accumulate_in = (p.flags & CNRMA_AGG_ACCUMULATE) != 0;

    for (int u = blockIdx.x * kWarps + warp; u < units; u += warps_total) {
        const int it = u * kLongVoxels + vj;
        const bool active = it < p.nvox;
        int vx, vy, vz;
        sweep_voxel(p.sweep, active ? it : 0, vx, vy, vz);
        const int vox = (vx * p.g.ny + vy) * p.g.nz + vz;    // voxel order of datasets/tsdf.py:24-29
        const float wx = world_coord(vx + p.g.x0, p.g.vs, p.g.ox);
        const float wy = world_coord(vy + p.g.y0, p.g.vs, p.g.oy);
        const float wz = world_coord(vz + p.g.z0, p.g.vs, p.g.oz);
        uint32_t *lst = my_lists + vj * p.lcap;
        int total = 0;          // views that see my voxel so far (all four lanes of a voxel hold the same value)
        if (accumulate_in && active)
            total = (p.flags & CNRMA_AGG_COUNT_F32) ? (int)reinterpret_cast<const float *>(p.count)[vox] : p.count[vox];
        bool first = !accumulate_in;
        int v0 = 0;
        do {
            // ---- phase 1: lane <-> (voxel, view), four views per step, until the views or a list run out -------------
            int cnt = 0;        // entries in my voxel's list (this segment)
            bool full = false;
            for (; v0 < p.V && !full; v0 += kLongStep) {
                const int v = v0 + sub;
                const bool vin = v < p.V;
                const float *P = sP + 12 * (vin ? v : 0);
                const float4 a = *reinterpret_cast<const float4 *>(P);
                const float4 b = *reinterpret_cast<const float4 *>(P + 4);
                const float4 c = *reinterpret_cast<const float4 *>(P + 8);
                const float cx = row_dot4(a.x, a.y, a.z, a.w, wx, wy, wz, 1.0f);
                const float cy = row_dot4(b.x, b.y, b.z, b.w, wx, wy, wz, 1.0f);
                const float cz = row_dot4(c.x, c.y, c.z, c.w, wx, wy, wz, 1.0f);
                const float slack = 1.0e-3f * cz;   // cheap superset of the frustum test (see the kernel above)
                const bool maybe = active && vin && (cz > 0.0f) && (cx + 0.5f * cz >= -slack) && (fW * cz - cx >= -slack) &&
                                   (cy + 0.5f * cz >= -slack) && (fH * cz - cy >= -slack);
                if (!__any_sync(0xffffffffu, maybe)) continue;
                float rx, ry;
                rounded_pixel(cx, cy, cz, rx, ry);
                const bool ok = maybe && in_frustum(rx, ry, cz, p.H, p.W);
                const unsigned hits = __ballot_sync(0xffffffffu, ok);
                if (ok)
                    lst[cnt + __popc(hits & below)] =
                        UNIFORM ? ((uint32_t)v * p.view_stride16 + (uint32_t)(int)ry * p.stride_y16 + (uint32_t)(int)rx * p.stride_x16)
                                : (((uint32_t)v << 20) | ((uint32_t)(int)ry << 10) | (uint32_t)(int)rx);
                cnt += __popc(hits & same);
                full = __any_sync(0xffffffffu, cnt + kLongStep > p.lcap);   // the next step could overflow a list
            }
            total += cnt;
            const bool last = v0 >= p.V;
            __syncwarp();

            // ---- phase 2: lane group <-> voxel ------------------------------------------------------------------------
            for (int b0 = 0; b0 < kLongVoxels; b0 += kVPW) {
                const int j = b0 + grp;                                       // voxel of the unit served by this group
                const int n = __shfl_sync(0xffffffffu, cnt, j);               // lane j (< 8) holds voxel j's values
                const int jvox = __shfl_sync(0xffffffffu, vox, j);
                const int jtotal = __shfl_sync(0xffffffffu, total, j);
                const bool jact = __shfl_sync(0xffffffffu, (int)active, j) != 0;
                const uint32_t *jl = my_lists + j * p.lcap;
                float acc[VPL][E];
                if (!first && jact) {
#pragma unroll
                    for (int q = 0; q < VPL; ++q)
#pragma unroll
                        for (int e = 0; e < E; ++e)
                            acc[q][e] = p.volume[(int64_t)jvox * p.vsv + (int64_t)(c0 + q * G * E + e) * p.vsc];
                } else {
#pragma unroll
                    for (int q = 0; q < VPL; ++q)
#pragma unroll
                        for (int e = 0; e < E; ++e) acc[q][e] = 0.0f;
                }
                const int nn = jact ? n : 0;
                int nmax = nn;   // longest list among the groups of this warp
#pragma unroll
                for (int o = 16; o >= G; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
                for (int k = 0; k < nmax; k += kLongUnroll) {
                    uint4 raw[kLongUnroll][VPL];
#pragma unroll
                    for (int uu = 0; uu < kLongUnroll; ++uu) {
                        if (k + uu < nn) {
                            const uint32_t e = jl[k + uu];
                            const unsigned char *src =
                                UNIFORM ? (view0 + (int64_t)e * 16 + lig * 16)
                                        : (sView[e >> 20] + (int64_t)((e >> 10) & 1023u) * p.stride_y_bytes +
                                           (int64_t)(e & 1023u) * p.stride_x_bytes + lig * 16);
#pragma unroll
                            for (int q = 0; q < VPL; ++q) raw[uu][q] = V16::load_raw(reinterpret_cast<const T *>(src + q * G * 16));
                        }
                    }
#pragma unroll
                    for (int uu = 0; uu < kLongUnroll; ++uu) {
                        if (k + uu < nn) {
#pragma unroll
                            for (int q = 0; q < VPL; ++q) {
                                const V16 val = V16::widen(raw[uu][q]);
#pragma unroll
                                for (int e = 0; e < E; ++e) acc[q][e] = __fadd_rn(acc[q][e], val.v[e]);
                            }
                        }
                    }
                }
                if (!jact) continue;
                if (last && (p.flags & CNRMA_AGG_MEAN)) {
                    const float fn = (float)jtotal;   // rm.py:251: fp32 sum / int64 count, 0 where count == 0
                    const float y = __frcp_rn(fn);
#pragma unroll
                    for (int q = 0; q < VPL; ++q)
#pragma unroll
                        for (int e = 0; e < E; ++e) acc[q][e] = (jtotal > 0) ? div_by_count(acc[q][e], fn, y) : 0.0f;
                }
#pragma unroll
                for (int q = 0; q < VPL; ++q) {
                    const int c = c0 + q * G * E;
                    if (p.vec_store) {
                        float *dst = p.volume + (int64_t)jvox * p.vsv + c;
#pragma unroll
                        for (int e = 0; e < E; e += 4) {
                            const float4 o4 = make_float4(acc[q][e], acc[q][e + 1], acc[q][e + 2], acc[q][e + 3]);
                            if (last) __stcs(reinterpret_cast<float4 *>(dst + e), o4);
                            else *reinterpret_cast<float4 *>(dst + e) = o4;     // read back by the next segment
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < E; ++e) p.volume[(int64_t)jvox * p.vsv + (int64_t)(c + e) * p.vsc] = acc[q][e];
                    }
                }
                if (last && lig == 0 && blockIdx.y == 0 && p.write_count) {
                    if (p.flags & CNRMA_AGG_COUNT_F32) reinterpret_cast<float *>(p.count)[jvox] = (float)jtotal;
                    else p.count[jvox] = jtotal;
                    if (p.valid != nullptr) p.valid[jvox] = (uint8_t)(jtotal > 0);
                }
            }
            first = false;
            __syncwarp();   // the lists are rewritten by the next segment / unit; partial sums written above are re-read
        } while (v0 < p.V);
    }
}

// ---- launch -----------------------------------------------------------------------------------------------------

template <int G, int VPL, typename T, bool UNIFORM>
static cudaError_t launch_list(const ListParams &p, int chunks, cudaStream_t stream) {
    const size_t smem = sizeof(float) * 12 * p.V + sizeof(void *) * p.V +
                        sizeof(uint32_t) * (size_t)(kListThreads / kWarp) * p.nb * p.lcap;
    auto kernel = aggregate_views_list_kernel<G, VPL, T, UNIFORM>;
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
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kListThreads, smem);
        if (err != cudaSuccess) return err;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        cache.dev = dev;
        cache.smem = smem;
        cache.ctas = sms * per_sm;
    }
    const int units = (p.nvox + p.nb - 1) / p.nb;
    const int needed = (units + (kListThreads / kWarp) - 1) / (kListThreads / kWarp);
    int persistent = cache.ctas - p.reserve_ctas;
    if (persistent < 1) persistent = 1;
    const dim3 grid(needed < persistent ? needed : persistent, chunks);
    kernel<<<grid, kListThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_list_gv(const ListParams &p, int g, int vpl, int chunks, cudaStream_t stream) {
#define CNRMA_LIST_CASE(GG, VV) \
    if (g == GG && vpl == VV)  \
        return p.uniform ? launch_list<GG, VV, T, true>(p, chunks, stream) : launch_list<GG, VV, T, false>(p, chunks, stream);
    CNRMA_LIST_CASE(1, 1) CNRMA_LIST_CASE(2, 1) CNRMA_LIST_CASE(4, 1) CNRMA_LIST_CASE(8, 1) CNRMA_LIST_CASE(16, 1)
    CNRMA_LIST_CASE(32, 1) CNRMA_LIST_CASE(32, 2) CNRMA_LIST_CASE(1, 3) CNRMA_LIST_CASE(2, 3) CNRMA_LIST_CASE(4, 3)
    CNRMA_LIST_CASE(8, 3) CNRMA_LIST_CASE(16, 3) CNRMA_LIST_CASE(32, 3) CNRMA_LIST_CASE(32, 4)
#undef CNRMA_LIST_CASE
    return cudaErrorInvalidValue;
}

template <int G, int VPL, typename T, bool UNIFORM>
static cudaError_t launch_long(const ListParams &p, int chunks, cudaStream_t stream) {
    const size_t smem = sizeof(float) * 12 * p.V + sizeof(void *) * p.V +
                        sizeof(uint32_t) * (size_t)(kListThreads / kWarp) * kLongVoxels * p.lcap;
    auto kernel = aggregate_views_long_kernel<G, VPL, T, UNIFORM>;
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
        err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kListThreads, smem);
        if (err != cudaSuccess) return err;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        cache.dev = dev;
        cache.smem = smem;
        cache.ctas = sms * per_sm;
    }
    const int units = (p.nvox + kLongVoxels - 1) / kLongVoxels;
    const int needed = (units + (kListThreads / kWarp) - 1) / (kListThreads / kWarp);
    int persistent = cache.ctas - p.reserve_ctas;
    if (persistent < 1) persistent = 1;
    const dim3 grid(needed < persistent ? needed : persistent, chunks);
    kernel<<<grid, kListThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_long_gv(const ListParams &p, int g, int vpl, int chunks, cudaStream_t stream) {
#define CNRMA_LONG_CASE(GG, VV) \
    if (g == GG && vpl == VV)  \
        return p.uniform ? launch_long<GG, VV, T, true>(p, chunks, stream) : launch_long<GG, VV, T, false>(p, chunks, stream);
    CNRMA_LONG_CASE(4, 1) CNRMA_LONG_CASE(8, 1) CNRMA_LONG_CASE(16, 1) CNRMA_LONG_CASE(32, 1) CNRMA_LONG_CASE(4, 3)
    CNRMA_LONG_CASE(8, 3) CNRMA_LONG_CASE(16, 3)
#undef CNRMA_LONG_CASE
    return cudaErrorInvalidValue;
}

// True when the long-list kernel serves this row shape: rows below 512 bytes made of a multiple of four 16-byte vectors
// (lane groups of 4, 8 or 16 lanes, one or three vectors per lane).
bool long_list_supports(int channels, int dtype) {
    const int row_bytes = channels * ((dtype == CNRMA_BF16) ? 2 : 4);
    return row_bytes < 512 && row_bytes % 64 == 0 && tuning().agg_list_views <= 0;
}

// True when the list kernel can serve this shape (pixel coordinates and view ids are packed into 32 bits).
bool list_kernel_supports(int V, int H, int W) { return V <= 4096 && H <= 1024 && W <= 1024 && V <= kMaxViewsPerLaunch; }

cudaError_t run_aggregate_list(const GridDev &g, const cnrma_features &f, int v0, int nv, const float *proj,
                               int64_t proj_stride, float stride, uint32_t flags, float *volume, int64_t vsv, int64_t vsc,
                               int32_t *count, uint8_t *valid, cudaStream_t stream, const OutputRoute *route,
                               int reserve_ctas) {
    const int esz = (f.dtype == CNRMA_BF16) ? 2 : 4;
    const int nvec = f.channels * esz / 16;
    // lanes per voxel: the largest power of two <= 32 dividing nvec; the rest as vectors per lane (<= 4) and chunks
    int G = 1;
    while (G < 32 && nvec % (G * 2) == 0) G *= 2;
    const int q = nvec / G;
    int vpl = 1;
    for (int d = 4; d >= 1; --d)
        if (q % d == 0 && (d == 1 || d == 3 || G == 32)) { vpl = d; break; }
    const int chunks = q / vpl;
    ListParams p;
    p.g = g;
    p.V = nv; p.C = f.channels; p.H = f.height; p.W = f.width;
    p.nvox = g.nx * g.ny * g.nz;
    p.stride_y_bytes = f.stride_y * esz;
    p.stride_x_bytes = f.stride_x * esz;
    p.stride = stride;
    p.proj = proj;
    p.proj_stride = proj_stride;
    p.volume = volume;
    p.vsv = vsv; p.vsc = vsc;
    p.count = count;
    p.valid = valid;
    p.flags = flags;
    p.vec_store = (vsc == 1) && (vsv % 4 == 0) && (reinterpret_cast<uintptr_t>(volume) % 16 == 0);
    p.route.n_owners = 0;
    if (route != nullptr) {
        p.route = *route;
        p.vec_store = 1;
    }
    p.sweep = make_sweep(g.nx, g.ny, g.nz, sweep_thickness(g.ny, g.nz, nv, f.channels * esz));
    p.reserve_ctas = reserve_ctas > 0 ? reserve_ctas : 0;
    // many views: the long-list kernel (all views of a unit in one go); its lists hold up to kLongListCap entries
    const bool long_lists = nv > kListViewsMax && route == nullptr && long_list_supports(f.channels, f.dtype);
    p.lcap = nv | 1;                                      // odd: the lanes' list writes hit different banks
    if (long_lists && p.lcap > kLongListCap) p.lcap = kLongListCap;
    int nb = 32;                                          // voxels per warp batch: lists must fit ~8 KB per warp
    while (nb > 1 && (size_t)nb * p.lcap * 4 > 8192) nb >>= 1;
    if (nb < 32 / G) nb = 32 / G;                         // at least one full gather round
    p.nb = long_lists ? kLongVoxels : nb;
    for (int i = 0; i < nv; ++i) p.views[i] = f.view_ptrs_host[v0 + i];
    // equally spaced views (one [V,...] tensor): entries become 32-bit offsets in 16-byte units
    p.uniform = 0;
    p.view_stride16 = p.stride_y16 = p.stride_x16 = 0;
    if (nv > 0 && p.stride_y_bytes % 16 == 0 && p.stride_x_bytes % 16 == 0) {
        const intptr_t base = reinterpret_cast<intptr_t>(p.views[0]);
        const intptr_t step = nv > 1 ? reinterpret_cast<intptr_t>(p.views[1]) - base : 0;
        bool ok = step >= 0 && step % 16 == 0;
        for (int i = 2; i < nv && ok; ++i) ok = reinterpret_cast<intptr_t>(p.views[i]) - base == (intptr_t)i * step;
        const int64_t span = (int64_t)(nv - 1) * step + (int64_t)f.height * p.stride_y_bytes + (int64_t)f.width * p.stride_x_bytes;
        if (ok && span / 16 < ((int64_t)1 << 32)) {
            p.uniform = 1;
            p.view_stride16 = (uint32_t)(step / 16);
            p.stride_y16 = (uint32_t)(p.stride_y_bytes / 16);
            p.stride_x16 = (uint32_t)(p.stride_x_bytes / 16);
        }
    }
    auto launch = [&](const ListParams &lp, int nchunks) -> cudaError_t {
        if (long_lists)
            return (f.dtype == CNRMA_BF16) ? launch_long_gv<__nv_bfloat16>(lp, G, vpl, nchunks, stream)
                                           : launch_long_gv<float>(lp, G, vpl, nchunks, stream);
        return (f.dtype == CNRMA_BF16) ? launch_list_gv<__nv_bfloat16>(lp, G, vpl, nchunks, stream)
                                       : launch_list_gv<float>(lp, G, vpl, nchunks, stream);
    };
    p.chunk_base = 0;
    p.write_count = 1;
    if (!(flags & CNRMA_AGG_ACCUMULATE) || chunks == 1) return launch(p, chunks);
    for (int c = 0; c < chunks; ++c) {   // see run_aggregate in cnrma_stage_a.cu: the count is rewritten by the last chunk only
        p.chunk_base = c;
        p.write_count = (c == chunks - 1);
        const cudaError_t err = launch(p, 1);
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

}  // namespace cnrma
