// N2 — SparseMatrix::transpose (sparsematrix.rs:174-183) on the device.
//
// The reference builds the transpose entry by entry: rows i ascending, a row's entries in storage order,
// `ret.set(col, i, val)`.  On the assembly format (SparseMatIndexList, whose `set` appends to the row's chain) followed
// by `to_crs()` this yields, for row j of the result, the entries of column j ordered by source row i — equal i (a
// duplicate (i, j), which IndexList-built matrices cannot hold) aside, that is the order of their positions k in the
// source arrays.  So the transpose is a STABLE sort of k = 0..nnz-1 by column[k]:
//
//   keys = columns, payload = k           LSD radix sort, 8-bit digits, ceil(bits(n_cols) / 8) passes, stable and
//                                         free of global atomics, hence deterministic (own kernels; no CUB)
//   offsets_T[j] = lower_bound(keys, j)   binary search in the sorted keys
//   columns_T[t] = row of perm[t]         binary search in offset_rows;  values_T[t] = values[perm[t]]
//
// Result dimensions as the reference produces them: n_rows = largest column + 1 (IndexList::n_rows = pos_start.len()),
// n_cols = last non-empty source row + 1 (`n_cols = max j + 1`), 0 x 0 when there are no entries (to_crs, sparsemat_crs.rs:25).
// Integer / copy work: bit-exact against the oracle's restatement (tests/test_gpu_transpose.py).
#include "common.cuh"

namespace smb {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;                          // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;        // 4096 keys per CTA, warp w owns keys [512 w, 512 (w + 1))
constexpr int kRsBins = 256;

// Pass 1 of a digit: per-tile histogram -> hist[digit * n_tiles + tile].
template <class K>
__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const K* __restrict__ keys, uint64_t n, unsigned shift, uint64_t* __restrict__ hist, uint64_t n_tiles) {
    __shared__ unsigned cnt[kRsBins];
    cnt[threadIdx.x] = 0u;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kRsTile;
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint64_t k = base + (uint64_t)r * kRsThreads + threadIdx.x;
        if (k < n) atomicAdd(&cnt[(unsigned)(keys[k] >> shift) & 255u], 1u);          // counts do not depend on the order
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * n_tiles + blockIdx.x] = cnt[threadIdx.x];
}

// Pass 2 of a digit: stable scatter.  `offs` is the exclusive scan of hist (digit-major), so offs[d * n_tiles + t] is where
// tile t's first key with digit d goes.  Inside a tile, warp w handles its 512 consecutive keys in 16 rounds of 32; a key's
// rank among equal digits = (equal digits in earlier warps) + (in earlier rounds of this warp) + (in lower lanes of this
// round) — the order of the keys in memory.
template <class K, class P, bool IOTA>
__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const K* __restrict__ keys_in, const P* __restrict__ pay_in, K* __restrict__ keys_out, P* __restrict__ pay_out,
                  uint64_t n, unsigned shift, const uint64_t* __restrict__ offs, uint64_t n_tiles) {
    constexpr int W = kRsThreads / 32;
    __shared__ unsigned warp_cnt[W][kRsBins];         // running count per warp and digit, then its exclusive prefix over warps
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < W * kRsBins; i += kRsThreads) (&warp_cnt[0][0])[i] = 0u;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kRsTile + (uint64_t)w * (32 * kRsItems);
    K key[kRsItems];
    unsigned rank[kRsItems];
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint64_t k = base + (uint64_t)r * 32 + lane;
        const bool live = k < n;
        key[r] = live ? keys_in[k] : K(0);
        const unsigned d = live ? ((unsigned)(key[r] >> shift) & 255u) : 256u;         // dead lanes match each other only
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned before = __popc(peers & ((1u << lane) - 1u));
        unsigned prev = 0;
        if (live) prev = warp_cnt[w][d];
        __syncwarp();
        if (live && before == 0) warp_cnt[w][d] = prev + __popc(peers);               // one lane per digit updates the count
        __syncwarp();
        rank[r] = prev + before;
    }
    __syncthreads();
    // exclusive prefix over the warps, per digit: thread d handles digit d
    {
        unsigned run = 0;
        const unsigned d = threadIdx.x;
#pragma unroll
        for (int ww = 0; ww < W; ++ww) { const unsigned c = warp_cnt[ww][d]; warp_cnt[ww][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const uint64_t k = base + (uint64_t)r * 32 + lane;
        if (k < n) {
            const unsigned d = (unsigned)(key[r] >> shift) & 255u;
            const uint64_t at = offs[(uint64_t)d * n_tiles + blockIdx.x] + warp_cnt[w][d] + rank[r];
            keys_out[at] = key[r];
            pay_out[at] = IOTA ? (P)k : pay_in[k];
        }
    }
}

// keys[k] = (K)columns[k]: the sort key, narrowed when the column range needs fewer bits than the index type has
template <class I, class K>
__global__ void transpose_keys_kernel(const I* __restrict__ cols, uint64_t n, K* __restrict__ keys) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < n; k += stride) keys[k] = (K)cols[k];
}

// offsets_T[j] = number of sorted keys < j, j = 0..n_out_rows (index type I).
template <class K, class I>
__global__ void transpose_offsets_kernel(const K* __restrict__ keys, uint64_t n, uint64_t n_out_rows, I* __restrict__ offs_t) {
    const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j > n_out_rows) return;
    uint64_t lo = 0, hi = n;
    while (lo < hi) { const uint64_t mid = lo + ((hi - lo) >> 1); if ((uint64_t)keys[mid] < j) lo = mid + 1; else hi = mid; }
    offs_t[j] = (I)lo;
}

// columns_T[t] = source row of entry perm[t] (the r with offs[r] <= k < offs[r + 1]); values_T[t] = values[perm[t]].
template <class T, class I, class P>
__global__ void transpose_gather_kernel(const T* __restrict__ vals, const I* __restrict__ offs, uint64_t n_rows, const P* __restrict__ perm,
                                        uint64_t n, T* __restrict__ vals_t, I* __restrict__ cols_t) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += stride) {
        const uint64_t k = (uint64_t)perm[t];
        uint64_t lo = 0, hi = n_rows;                 // largest r with offs[r] <= k  (empty rows repeat an offset: take the last)
        while (hi - lo > 1) { const uint64_t mid = lo + ((hi - lo) >> 1); if ((uint64_t)offs[mid] <= k) lo = mid; else hi = mid; }
        cols_t[t] = (I)lo;
        vals_t[t] = vals[k];
    }
}

template <class T, class I, class K, class P>
static smb200_status transpose_typed(const smb200_crs* a, smb200_crs** out) {
    smb200_ctx* ctx = a->ctx;
    const uint64_t n = a->nnz;
    cudaStream_t st = ctx->stream;
    K* keys[2] = {nullptr, nullptr};
    P* pay[2] = {nullptr, nullptr};
    uint64_t* hist = nullptr;
    smb200_crs* t = nullptr;
    auto cleanup = [&] {
        for (int i = 0; i < 2; ++i) { if (keys[i]) cudaFree(keys[i]); if (pay[i]) cudaFree(pay[i]); }
        if (hist) cudaFree(hist);
    };
#define TR_CUDA(expr) do { cudaError_t te__ = (expr); if (te__ != cudaSuccess) { cleanup(); if (t) smb200_crs_free(t); SMB_CUDA(te__); } } while (0)
#define TR_TRY(expr) do { smb200_status s__ = (expr); if (s__ != SMB200_OK) { cleanup(); if (t) smb200_crs_free(t); return s__; } } while (0)
    const uint64_t n_tiles = (n + kRsTile - 1) / kRsTile;
    for (int i = 0; i < 2; ++i) { TR_CUDA(cudaMalloc(&keys[i], n * sizeof(K))); TR_CUDA(cudaMalloc(&pay[i], n * sizeof(P))); }
    TR_CUDA(cudaMalloc(&hist, (n_tiles * kRsBins + 1) * sizeof(uint64_t)));
    // keys: the columns, narrowed to K when the index type is wider than the column range needs
    transpose_keys_kernel<I, K><<<(unsigned)ctx->sm_count * 8, 256, 0, st>>>((const I*)a->columns, n, keys[0]);
    count_launch();
    unsigned bits = 1;
    while (bits < 64 && (a->n_cols - 1) >> bits) ++bits;
    const unsigned passes = (bits + 7) / 8;
    int cur = 0;
    for (unsigned p = 0; p < passes; ++p) {
        const unsigned shift = 8 * p;
        rs_hist_kernel<K><<<(unsigned)n_tiles, kRsThreads, 0, st>>>(keys[cur], n, shift, hist, n_tiles);
        count_launch();
        TR_TRY(exclusive_scan_inplace(ctx, SMB200_U64, hist, n_tiles * kRsBins, nullptr));
        if (p == 0) rs_scatter_kernel<K, P, true><<<(unsigned)n_tiles, kRsThreads, 0, st>>>(keys[cur], nullptr, keys[cur ^ 1], pay[cur ^ 1], n, shift, hist, n_tiles);
        else rs_scatter_kernel<K, P, false><<<(unsigned)n_tiles, kRsThreads, 0, st>>>(keys[cur], pay[cur], keys[cur ^ 1], pay[cur ^ 1], n, shift, hist, n_tiles);
        count_launch();
        TR_CUDA(cudaGetLastError());
        cur ^= 1;
    }
    // dimensions as the reference's loop leaves them
    K last_key = 0;
    TR_CUDA(cudaMemcpyAsync(&last_key, keys[cur] + (n - 1), sizeof(K), cudaMemcpyDeviceToHost, st));
    TR_CUDA(cudaStreamSynchronize(st));
    const uint64_t rows_t = (uint64_t)last_key + 1;
    uint64_t cols_t = a->n_rows;                       // last non-empty source row + 1: the row of entry nnz - 1
    {
        std::vector<unsigned char> tail;               // offsets are monotone: search from the end on the host, a chunk at a time
        const size_t is = sizeof(I);
        uint64_t hi = a->n_rows;
        while (hi > 0) {
            const uint64_t lo = hi > 65536 ? hi - 65536 : 0;
            tail.resize((hi - lo) * is);
            TR_CUDA(cudaMemcpyAsync(tail.data(), (const char*)a->offsets + lo * is, (hi - lo) * is, cudaMemcpyDeviceToHost, st));
            TR_CUDA(cudaStreamSynchronize(st));
            const I* o = (const I*)tail.data();
            bool found = false;
            for (uint64_t r = hi; r-- > lo;) if ((uint64_t)o[r - lo] < n) { cols_t = r + 1; found = true; break; }
            if (found) break;
            hi = lo;
        }
    }
    TR_TRY(crs_alloc(ctx, a->vt, a->it, rows_t, cols_t, n, &t));
    transpose_offsets_kernel<K, I><<<(unsigned)((rows_t + 1 + 255) / 256), 256, 0, st>>>(keys[cur], n, rows_t, (I*)t->offsets);
    count_launch();
    transpose_gather_kernel<T, I, P><<<(unsigned)ctx->sm_count * 8, 256, 0, st>>>((const T*)a->values, (const I*)a->offsets, a->n_rows, pay[cur], n,
                                                                                 (T*)t->values, (I*)t->columns);
    count_launch();
    TR_CUDA(cudaGetLastError());
    TR_CUDA(cudaStreamSynchronize(st));
    cleanup();
    const smb200_status fs = crs_finalize(t, false);
    if (fs != SMB200_OK) { smb200_crs_free(t); return fs; }
#undef TR_CUDA
#undef TR_TRY
    *out = t;
    return SMB200_OK;
}

}  // namespace smb

using namespace smb;

extern "C" smb200_status smb200_crs_transpose(const smb200_crs* a, smb200_crs** out) {
    SMB_REQUIRE(a && out, SMB200_ERR_INVALID, "crs_transpose: NULL argument");
    SMB_REQUIRE(a->x_extra == 0, SMB200_ERR_UNSUPPORTED, "crs_transpose: the local block of a distributed matrix has remapped columns");
    *out = nullptr;
    SMB_CUDA(cudaSetDevice(a->ctx->device));
    if (a->nnz == 0) return crs_alloc(a->ctx, a->vt, a->it, 0, 0, 0, out);       // to_crs of an IndexList without entries: 0 x 0
    const bool k32 = a->n_cols <= 0xFFFFFFFFull, p32 = a->nnz <= 0xFFFFFFFFull;
#define TR_DISPATCH(T, I)                                                                                            \
    (k32 ? (p32 ? transpose_typed<T, I, uint32_t, uint32_t>(a, out) : transpose_typed<T, I, uint32_t, uint64_t>(a, out)) \
         : transpose_typed<T, I, uint64_t, uint64_t>(a, out))
    if (a->vt == SMB200_F64) return a->it == SMB200_U64 ? TR_DISPATCH(double, uint64_t) : transpose_typed<double, uint32_t, uint32_t, uint32_t>(a, out);
    return a->it == SMB200_U64 ? TR_DISPATCH(float, uint64_t) : transpose_typed<float, uint32_t, uint32_t, uint32_t>(a, out);
#undef TR_DISPATCH
}
