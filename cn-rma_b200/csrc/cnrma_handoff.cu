// Point-cloud hand-off: what RayMarching.switch_pointcloud does between the RMA lift
// and the sparse detector (rm.py:339-407) -- `coord + offset`, then the random sub-sampling mask of sample_points
// (fcaf3d_transforms.py:283-296, max_points = 500000) applied with one torch.masked_select PER COLUMN (C + 3 of
// them, rm.py:380-402).  Here: one ordered compaction of whole rows.
//
//   mask_prefix    exclusive prefix sum of the keep mask over the M rows (block sums -> scan -> per-row prefix)
//   select_rows    kept rows copied in order to out[prefix[row]], the offset added to x, y, z (one rounding, like
//                  the reference's `coord + offsets[b]`)
// The same prefix array lets cnrma_rma_fill write ONLY the kept rows (fill_rows_kernel, `sel_*` parameters), which
// removes ~90 % of the point-row traffic when max_points << M.
#include "cnrma_internal.cuh"

namespace cnrma {

constexpr int kSelThreads = 256;

__global__ void __launch_bounds__(kSelThreads) mask_block_sums_kernel(const uint8_t *__restrict__ mask, int64_t M,
                                                                      int32_t *__restrict__ blk) {
    __shared__ int s[kSelThreads / kWarp];
    const int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x;
    int v = (i < M && mask[i]) ? 1 : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < kSelThreads / kWarp; ++k) t += s[k];
        blk[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan of the block sums (int64 offsets) and the total
__global__ void __launch_bounds__(1024) scan_i32_kernel(const int32_t *__restrict__ blk, int64_t *__restrict__ blk_off,
                                                        int64_t n, int64_t *__restrict__ total) {
    __shared__ int64_t s[1024];
    const int t = threadIdx.x;
    const int64_t per = (n + 1023) / 1024;
    const int64_t lo = (int64_t)t * per, hi = (lo + per < n) ? lo + per : n;
    int64_t sum = 0;
    for (int64_t i = lo; i < hi; ++i) sum += blk[i];
    s[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int64_t add = (t >= o) ? s[t - o] : 0;
        __syncthreads();
        s[t] += add;
        __syncthreads();
    }
    int64_t run = s[t] - sum;
    for (int64_t i = lo; i < hi; ++i) {
        blk_off[i] = run;
        run += blk[i];
    }
    if (t == 1023) *total = s[t];
}

__global__ void __launch_bounds__(kSelThreads) mask_prefix_kernel(const uint8_t *__restrict__ mask, int64_t M,
                                                                  const int64_t *__restrict__ blk_off,
                                                                  int32_t *__restrict__ prefix) {
    __shared__ int s[kSelThreads / kWarp];
    const int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x;
    const int v = (i < M && mask[i]) ? 1 : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nn = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nn;
    }
    if (lane == 31) s[warp] = incl;
    __syncthreads();
    int64_t base = blk_off[blockIdx.x];
    for (int k = 0; k < warp; ++k) base += s[k];
    if (i < M) prefix[i] = (int32_t)(base + incl - v);
}

// one warp per 32 rows; kept rows are copied whole, coalesced
__global__ void __launch_bounds__(kSelThreads) select_rows_kernel(const float *__restrict__ rows, int64_t row_stride,
                                                                  int cols, int64_t M, const uint8_t *__restrict__ mask,
                                                                  const int32_t *__restrict__ prefix, float ox, float oy,
                                                                  float oz, float *__restrict__ out, int64_t out_stride,
                                                                  int64_t capacity) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (kSelThreads / kWarp) + (threadIdx.x >> 5);
    const int64_t r0 = warp * 32;
    if (r0 >= M) return;
    const int64_t mine = r0 + lane;
    const bool keep = mine < M && mask[mine];
    const int dst_mine = keep ? prefix[mine] : -1;
    const unsigned bits = __ballot_sync(0xffffffffu, keep);
    for (unsigned b = bits; b; b &= b - 1) {
        const int k = __ffs(b) - 1;
        const int64_t dst = __shfl_sync(0xffffffffu, dst_mine, k);
        if (dst >= capacity) continue;
        const float *src = rows + (r0 + k) * row_stride;
        float *d = out + dst * out_stride;
        for (int c = lane; c < cols; c += 32) {
            float v = __ldg(src + c);
            if (c < 3) v = __fadd_rn(v, c == 0 ? ox : (c == 1 ? oy : oz));   // coord + offset (rm.py:365)
            __stcs(d + c, v);
        }
    }
}

size_t handoff_workspace_bytes(int64_t M) {
    const int64_t blocks = (M + kSelThreads - 1) / kSelThreads;
    return ((size_t)blocks * 4 + 255) / 256 * 256 + ((size_t)blocks * 8 + 255) / 256 * 256 + 256;
}

// prefix[row] for all rows and *total (device int64) = number of kept rows
cudaError_t run_mask_prefix(const uint8_t *mask, int64_t M, void *workspace, int32_t *prefix, int64_t *total,
                            cudaStream_t stream) {
    const int64_t blocks = (M + kSelThreads - 1) / kSelThreads;
    unsigned char *base = static_cast<unsigned char *>(workspace);
    int32_t *blk = reinterpret_cast<int32_t *>(base);
    int64_t *blk_off = reinterpret_cast<int64_t *>(base + ((size_t)blocks * 4 + 255) / 256 * 256);
    if (M == 0) return cudaMemsetAsync(total, 0, sizeof(int64_t), stream);
    mask_block_sums_kernel<<<(unsigned)blocks, kSelThreads, 0, stream>>>(mask, M, blk);
    scan_i32_kernel<<<1, 1024, 0, stream>>>(blk, blk_off, blocks, total);
    mask_prefix_kernel<<<(unsigned)blocks, kSelThreads, 0, stream>>>(mask, M, blk_off, prefix);
    return cudaGetLastError();
}

cudaError_t run_select_rows(const float *rows, int64_t row_stride, int cols, int64_t M, const uint8_t *mask,
                            const int32_t *prefix, const float *offset3_host, float *out, int64_t out_stride,
                            int64_t capacity, cudaStream_t stream) {
    if (M == 0) return cudaSuccess;
    const int64_t warps = (M + 31) / 32;
    const unsigned blocks = (unsigned)((warps + (kSelThreads / kWarp) - 1) / (kSelThreads / kWarp));
    select_rows_kernel<<<blocks, kSelThreads, 0, stream>>>(rows, row_stride, cols, M, mask, prefix, offset3_host[0],
                                                          offset3_host[1], offset3_host[2], out, out_stride, capacity);
    return cudaGetLastError();
}

}  // namespace cnrma

// ---- device-side sampler ---------------------------------------------------------------------------------------------
// sample_points draws its mask with np.random.choice(N, max_points, replace=False) on the host
// (fcaf3d_transforms.py:290): a full permutation of ~6 M indices, ~0.2 s per scene -- two orders of magnitude more
// than the whole lift.  This is the on-device equivalent in distribution (NOT the same random stream, so it is an
// option, never the parity path): every row gets the key (hash32(seed, i) << 32 | i), the k smallest keys are kept.
// Keys are distinct by construction, so exactly k rows are selected; the subset is uniform up to the quality of
// the hash.  Selection = one 16-bit histogram pass over the keys, a scan of the 65536 bins to find the bin that
// contains the k-th key, a rank computation inside that bin (a few hundred keys), and a final compare pass.
namespace cnrma {

constexpr int kSampleBins = 65536;
constexpr int kSampleCand = 8192;   // capacity for the keys of the boundary bin (n / 65536 on average)

struct SampleState {        // lives in the caller's workspace
    unsigned int hist[kSampleBins];
    unsigned long long threshold;   // keys <= threshold are kept
    unsigned int bin;               // boundary bin
    unsigned int below;             // keys in the bins before it
    unsigned int ncand;
    unsigned int overflow;
    unsigned long long cand[kSampleCand];
};

__device__ __forceinline__ unsigned long long sample_key(unsigned long long seed, unsigned int i) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (z & 0xFFFFFFFF00000000ull) | i;
}

// The row count either comes from the host (`n`) or, for the sync-free hand-off, from the march's result block in
// device memory (`n_dev`, clamped to the `n` rows the buffers hold).
__device__ __forceinline__ unsigned int sample_rows(unsigned int n, const long long *n_dev) {
    if (n_dev == nullptr) return n;
    const long long m = *n_dev;
    return m < 0 ? 0u : (m < (long long)n ? (unsigned int)m : n);
}

__global__ void __launch_bounds__(256) sample_hist_kernel(unsigned long long seed, unsigned int n, const long long *n_dev,
                                                          SampleState *st) {
    n = sample_rows(n, n_dev);
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&st->hist[sample_key(seed, i) >> 48], 1u);
}

__global__ void __launch_bounds__(1024) sample_find_bin_kernel(unsigned int k, unsigned int n, const long long *n_dev,
                                                               SampleState *st) {
    __shared__ unsigned int s[1024];
    const int t = threadIdx.x;
    n = sample_rows(n, n_dev);
    if (k >= n) {                       // nothing to drop (only reachable with a device-side row count): keep every row
        if (t == 0) {
            st->threshold = ~0ull;
            st->bin = kSampleBins;      // no key lives there: collect / threshold become no-ops
            st->below = 0;
            st->ncand = 0;
            st->overflow = 0;
        }
        return;
    }
    unsigned int sum = 0;
    for (int b = t * 64; b < t * 64 + 64; ++b) sum += st->hist[b];
    s[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned int add = (t >= o) ? s[t - o] : 0;
        __syncthreads();
        s[t] += add;
        __syncthreads();
    }
    const unsigned int before = s[t] - sum;     // keys in the strips before this one
    if (k > before && k <= s[t]) {              // the k-th smallest key lives in this strip
        unsigned int run = before;
        for (int b = t * 64; b < t * 64 + 64; ++b) {
            if (k <= run + st->hist[b]) {
                st->bin = b;
                st->below = run;
                break;
            }
            run += st->hist[b];
        }
    }
    if (t == 0) {
        st->ncand = 0;
        st->overflow = 0;
    }
}

__global__ void __launch_bounds__(256) sample_collect_kernel(unsigned long long seed, unsigned int n, const long long *n_dev,
                                                             SampleState *st) {
    n = sample_rows(n, n_dev);
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = sample_key(seed, i);
    if ((unsigned int)(key >> 48) == st->bin) {
        const unsigned int slot = atomicAdd(&st->ncand, 1u);
        if (slot < kSampleCand) st->cand[slot] = key;
        else st->overflow = 1;
    }
}

__global__ void __launch_bounds__(1024) sample_threshold_kernel(unsigned int k, SampleState *st) {
    // the (k - below)-th smallest key of the boundary bin: rank by counting (a few hundred keys)
    if (st->bin >= (unsigned int)kSampleBins) return;   // every row is kept (see sample_find_bin_kernel)
    const unsigned int m = min(st->ncand, (unsigned int)kSampleCand);
    const unsigned int want = k - st->below;    // 1-based rank inside the bin
    for (unsigned int a = threadIdx.x; a < m; a += blockDim.x) {
        const unsigned long long key = st->cand[a];
        unsigned int rank = 1;
        for (unsigned int b = 0; b < m; ++b) rank += (st->cand[b] < key);
        if (rank == want) st->threshold = key;
    }
}

__global__ void __launch_bounds__(256) sample_mask_kernel(unsigned long long seed, unsigned int cap, const long long *n_dev,
                                                          const SampleState *st, uint8_t *__restrict__ mask) {
    const unsigned int n = sample_rows(cap, n_dev);
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) mask[i] = (uint8_t)(i < n && sample_key(seed, i) <= st->threshold);   // rows beyond the count: dropped
}

size_t sample_workspace_bytes() { return (sizeof(SampleState) + 255) / 256 * 256; }

// n: the number of rows (n_dev == nullptr) or the capacity of `mask` when the row count is read from *n_dev.
cudaError_t run_sample_mask(int64_t n, int64_t k, unsigned long long seed, void *workspace, uint8_t *mask,
                            cudaStream_t stream, const long long *n_dev) {
    if (n == 0) return cudaSuccess;
    if (n_dev == nullptr && k >= n) return cudaMemsetAsync(mask, 1, (size_t)n, stream);
    if (k <= 0) return cudaMemsetAsync(mask, 0, (size_t)n, stream);
    SampleState *st = static_cast<SampleState *>(workspace);
    cudaError_t err = cudaMemsetAsync(st->hist, 0, sizeof(st->hist), stream);
    if (err != cudaSuccess) return err;
    const unsigned int blocks = (unsigned int)((n + 255) / 256);
    sample_hist_kernel<<<blocks, 256, 0, stream>>>(seed, (unsigned int)n, n_dev, st);
    sample_find_bin_kernel<<<1, 1024, 0, stream>>>((unsigned int)(k < n ? k : n), (unsigned int)n, n_dev, st);
    sample_collect_kernel<<<blocks, 256, 0, stream>>>(seed, (unsigned int)n, n_dev, st);
    sample_threshold_kernel<<<1, 1024, 0, stream>>>((unsigned int)(k < n ? k : n), st);
    sample_mask_kernel<<<blocks, 256, 0, stream>>>(seed, (unsigned int)n, n_dev, st, mask);
    return cudaGetLastError();
}

}  // namespace cnrma

// ---- quantisation of the hand-off (rm.py:330-332) ---------------------------------------------------------------
// The reference hands `coords / voxel_size_fcaf3d` (0.01 m) to MinkowskiEngine 0.5.4 (`ME.utils.batch_sparse_collate`
// then `ME.SparseTensor`; third-party, not vendored): the collate step stores the float coordinates into an int32 tensor
// (truncation toward zero), and the sparse tensor keeps ONE row per occupied cell (its default quantisation mode picks
// an arbitrary duplicate -- whichever thread wins on the GPU).  Here the deterministic member of that family: the FIRST
// row of every cell in row order survives, survivors stay in row order.
//   insert   open-addressing hash table keyed by the packed cell (3 x 21 bits); every row atomicMin's its index into its
//            cell's slot
//   keep     a row survives iff it is the minimum of its cell
//   compact  survivors (whole rows + their int32 cells) written in order through the mask's prefix sum
namespace cnrma {

constexpr unsigned long long kQuantEmpty = ~0ull;
constexpr int kQuantRange = 1 << 20;   // |cell index| below 2^20 on every axis (10 km at 0.01 m)

__device__ __forceinline__ bool quant_cell(const float *__restrict__ row, float vs, int q[3]) {
    bool ok = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float f = __fdiv_rn(__ldg(row + a), vs);     // coord / voxel_size: IEEE division (torch CPU, rm.py:331)
        ok = ok && (f > -(float)kQuantRange) && (f < (float)kQuantRange);   // also false for NaN
        q[a] = ok ? (int)f : 0;                            // float -> int32: truncation toward zero
    }
    return ok;
}

__device__ __forceinline__ unsigned long long quant_key(const int q[3]) {
    return ((unsigned long long)(unsigned)(q[0] + kQuantRange) << 42) | ((unsigned long long)(unsigned)(q[1] + kQuantRange) << 21) |
           (unsigned long long)(unsigned)(q[2] + kQuantRange);
}

__device__ __forceinline__ unsigned int quant_hash(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned int)k;
}

__global__ void __launch_bounds__(256) quantize_insert_kernel(const float *__restrict__ rows, int64_t stride, int64_t n,
                                                              float vs, unsigned long long *keys, int *vals,
                                                              unsigned int slot_mask, int *bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int q[3];
    if (!quant_cell(rows + i * stride, vs, q)) {
        *bad = 1;
        return;
    }
    const unsigned long long key = quant_key(q);
    unsigned int h = quant_hash(key) & slot_mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(keys + h, kQuantEmpty, key);
        if (prev == kQuantEmpty || prev == key) {
            atomicMin(vals + h, (int)i);
            return;
        }
        h = (h + 1) & slot_mask;
    }
}

__global__ void __launch_bounds__(256) quantize_keep_kernel(const float *__restrict__ rows, int64_t stride, int64_t n, float vs,
                                                            const unsigned long long *__restrict__ keys,
                                                            const int *__restrict__ vals, unsigned int slot_mask,
                                                            uint8_t *__restrict__ keep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int q[3];
    if (!quant_cell(rows + i * stride, vs, q)) {
        keep[i] = 0;
        return;
    }
    const unsigned long long key = quant_key(q);
    unsigned int h = quant_hash(key) & slot_mask;
    while (keys[h] != key) h = (h + 1) & slot_mask;        // the key was inserted by the first pass
    keep[i] = (uint8_t)(vals[h] == (int)i);
}

// one warp per 32 rows; surviving rows are copied whole (coalesced) and their cells written as int32
__global__ void __launch_bounds__(kSelThreads) quantize_compact_kernel(const float *__restrict__ rows, int64_t stride, int cols,
                                                                       int64_t n, float vs, const uint8_t *__restrict__ keep,
                                                                       const int32_t *__restrict__ prefix,
                                                                       float *__restrict__ out, int64_t out_stride,
                                                                       int32_t *__restrict__ cells, int64_t capacity) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (kSelThreads / kWarp) + (threadIdx.x >> 5);
    const int64_t r0 = warp * 32;
    if (r0 >= n) return;
    const int64_t mine = r0 + lane;
    const bool k = mine < n && keep[mine];
    const int dst_mine = k ? prefix[mine] : -1;
    if (k && dst_mine < capacity && cells != nullptr) {
        int q[3];
        quant_cell(rows + mine * stride, vs, q);
        cells[(int64_t)dst_mine * 3 + 0] = q[0];
        cells[(int64_t)dst_mine * 3 + 1] = q[1];
        cells[(int64_t)dst_mine * 3 + 2] = q[2];
    }
    const unsigned bits = __ballot_sync(0xffffffffu, k);
    for (unsigned b = bits; b; b &= b - 1) {
        const int j = __ffs(b) - 1;
        const int64_t dst = __shfl_sync(0xffffffffu, dst_mine, j);
        if (dst >= capacity) continue;
        const float *src = rows + (r0 + j) * stride;
        float *d = out + dst * out_stride;
        for (int c = lane; c < cols; c += 32) __stcs(d + c, __ldg(src + c));
    }
}

static unsigned int quant_slots(int64_t n) {
    unsigned int s = 1024;
    while ((int64_t)s < 2 * n) s <<= 1;
    return s;
}

size_t quantize_workspace_bytes(int64_t n) {
    const size_t slots = quant_slots(n);
    return slots * 8 + slots * 4 + 256;
}

cudaError_t run_quantize_mark(const float *rows, int64_t stride, int64_t n, float vs, void *workspace, uint8_t *keep,
                              cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const unsigned int slots = quant_slots(n);
    unsigned char *base = static_cast<unsigned char *>(workspace);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(base);
    int *vals = reinterpret_cast<int *>(base + (size_t)slots * 8);
    int *bad = reinterpret_cast<int *>(base + (size_t)slots * 12);
    cudaError_t err = cudaMemsetAsync(keys, 0xFF, (size_t)slots * 8, stream);
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(vals, 0x7F, (size_t)slots * 4, stream);   // 0x7F7F7F7F: larger than any row index
    if (err != cudaSuccess) return err;
    err = cudaMemsetAsync(bad, 0, sizeof(int), stream);
    if (err != cudaSuccess) return err;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    quantize_insert_kernel<<<blocks, 256, 0, stream>>>(rows, stride, n, vs, keys, vals, slots - 1, bad);
    quantize_keep_kernel<<<blocks, 256, 0, stream>>>(rows, stride, n, vs, keys, vals, slots - 1, keep);
    return cudaGetLastError();
}

cudaError_t run_quantize_compact(const float *rows, int64_t stride, int cols, int64_t n, float vs, const uint8_t *keep,
                                 const int32_t *prefix, float *out, int64_t out_stride, int32_t *cells, int64_t capacity,
                                 cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int64_t warps = (n + 31) / 32;
    const unsigned blocks = (unsigned)((warps + (kSelThreads / kWarp) - 1) / (kSelThreads / kWarp));
    quantize_compact_kernel<<<blocks, kSelThreads, 0, stream>>>(rows, stride, cols, n, vs, keep, prefix, out, out_stride, cells,
                                                                capacity);
    return cudaGetLastError();
}

}  // namespace cnrma
