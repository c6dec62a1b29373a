// SparseMatCRS<T,I> on the device: storage, upload/validation/download, and K8 — the IndexList -> CRS
// conversion (sparsemat_crs.rs:24-50) done on the GPU with the reference's exact layout.
#include "common.cuh"

namespace smb {

// ---- exclusive scan (in place, index type), used by K8 and the generators ----------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <class I>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const I* __restrict__ data, uint64_t n, uint64_t* __restrict__ sums) {
    __shared__ uint64_t warp_sums[kScanThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += (uint64_t)data[base + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += warp_sums[w];
        sums[blockIdx.x] = t;
    }
}

// Scans each tile locally and adds the tile's global prefix (tile_prefix may be NULL for one tile).
template <class I>
__global__ void __launch_bounds__(kScanThreads) scan_tiles(I* __restrict__ data, uint64_t n, const uint64_t* __restrict__ tile_prefix) {
    __shared__ uint64_t warp_sums[kScanThreads / 32];
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t item[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        item[k] = (base + k < n) ? (uint64_t)data[base + k] : 0;
        s += item[k];
    }
    // inclusive scan of the per-thread sums inside the warp
    uint64_t incl = s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint64_t warp_prefix = 0;
    for (int w = 0; w < warp; ++w) warp_prefix += warp_sums[w];
    uint64_t run = (tile_prefix ? tile_prefix[blockIdx.x] : 0) + warp_prefix + (incl - s);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = (I)run;
        run += item[k];
    }
}

template <class I>
static smb200_status scan_impl(smb200_ctx* ctx, I* data, uint64_t n) {
    if (n == 0) return SMB200_OK;
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1) {
        scan_tiles<I><<<1, kScanThreads, 0, ctx->stream>>>(data, n, nullptr);
        count_launch();
        SMB_CUDA(cudaGetLastError());
        return SMB200_OK;
    }
    uint64_t* sums = nullptr;
    SMB_CUDA(cudaMalloc(&sums, tiles * sizeof(uint64_t)));
    scan_tile_sums<I><<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(data, n, sums);
    count_launch();
    smb200_status s = scan_impl<uint64_t>(ctx, sums, tiles);
    if (s == SMB200_OK) {
        scan_tiles<I><<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(data, n, sums);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) s = SMB200_ERR_CUDA;
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(sums);
    if (s != SMB200_OK) SMB_FAIL(s, "exclusive scan failed");
    return SMB200_OK;
}

// data[0..n) <- exclusive prefix sums; if total != NULL it receives the sum of all inputs
// (callers pass n = n_rows + 1 with a trailing zero so that data[n_rows] becomes the total).
smb200_status exclusive_scan_inplace(smb200_ctx* ctx, int it, void* data, uint64_t n, uint64_t* total) {
    uint64_t last_in = 0;
    if (total && n) {
        if (it == SMB200_U64) { SMB_CUDA(cudaMemcpyAsync(&last_in, (uint64_t*)data + (n - 1), 8, cudaMemcpyDeviceToHost, ctx->stream)); }
        else { uint32_t t = 0; SMB_CUDA(cudaMemcpyAsync(&t, (uint32_t*)data + (n - 1), 4, cudaMemcpyDeviceToHost, ctx->stream)); SMB_CUDA(cudaStreamSynchronize(ctx->stream)); last_in = t; }
        SMB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (it == SMB200_U64) SMB_TRY(scan_impl<uint64_t>(ctx, (uint64_t*)data, n));
    else SMB_TRY(scan_impl<uint32_t>(ctx, (uint32_t*)data, n));
    if (total) {
        uint64_t last_out = 0;
        if (n) {
            if (it == SMB200_U64) { SMB_CUDA(cudaMemcpyAsync(&last_out, (uint64_t*)data + (n - 1), 8, cudaMemcpyDeviceToHost, ctx->stream)); SMB_CUDA(cudaStreamSynchronize(ctx->stream)); }
            else { uint32_t t = 0; SMB_CUDA(cudaMemcpyAsync(&t, (uint32_t*)data + (n - 1), 4, cudaMemcpyDeviceToHost, ctx->stream)); SMB_CUDA(cudaStreamSynchronize(ctx->stream)); last_out = t; }
        }
        *total = last_out + last_in;
    }
    return SMB200_OK;
}

// ---- validation / row statistics ----------------------------------------------------------------------
struct CrsCheck {
    unsigned long long max_row_len;
    unsigned int bad_offsets;
    unsigned int bad_columns;
};

template <class I>
__global__ void row_stats_kernel(const I* __restrict__ off, uint64_t n_rows, uint64_t nnz, CrsCheck* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long local_max = 0;
    unsigned int bad = 0;
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint64_t a = (uint64_t)off[r], b = (uint64_t)off[r + 1];
        if (b < a || b > nnz) bad = 1;
        else if (b - a > local_max) local_max = b - a;
        if (r == 0 && a != 0) bad = 1;
        if (r == n_rows - 1 && b != nnz) bad = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, local_max, o);
        local_max = t > local_max ? t : local_max;
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (local_max) atomicMax(&out->max_row_len, local_max);
        if (bad) atomicOr(&out->bad_offsets, 1u);
    }
}

template <class I>
__global__ void col_check_kernel(const I* __restrict__ cols, uint64_t nnz, uint64_t n_cols, CrsCheck* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned int bad = 0;
    for (uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; k < nnz; k += stride)
        if ((uint64_t)cols[k] >= n_cols) bad = 1;
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&out->bad_columns, 1u);
}

smb200_status crs_alloc(smb200_ctx* ctx, int vt, int it, uint64_t n_rows, uint64_t n_cols, uint64_t nnz, smb200_crs** out) {
    SMB_REQUIRE(ctx && out, SMB200_ERR_INVALID, "crs: NULL argument");
    SMB_REQUIRE(vt == SMB200_F32 || vt == SMB200_F64, SMB200_ERR_INVALID, "crs: bad value type %d", vt);
    SMB_REQUIRE(it == SMB200_U32 || it == SMB200_U64, SMB200_ERR_INVALID, "crs: bad index type %d", it);
    if (it == SMB200_U32) {
        // offset_rows is Vec<I>: nnz itself must be representable and != I::MAX (the UNSET sentinel)
        SMB_REQUIRE(nnz < 0xFFFFFFFFull && n_cols <= 0xFFFFFFFFull && n_rows < 0xFFFFFFFFull, SMB200_ERR_INVALID,
                    "crs: sizes do not fit the u32 index type (n_rows=%llu n_cols=%llu nnz=%llu)",
                    (unsigned long long)n_rows, (unsigned long long)n_cols, (unsigned long long)nnz);
    }
    *out = nullptr;
    SMB_CUDA(cudaSetDevice(ctx->device));
    smb200_crs* m = new smb200_crs();
    m->ctx = ctx; m->vt = vt; m->it = it; m->n_rows = n_rows; m->n_cols = n_cols; m->nnz = nnz;
    ctx_retain(ctx);
    smb200_status s = SMB200_OK;
    if (nnz) {
        s = dev_alloc(&m->values, nnz * vsize(vt));
        if (s == SMB200_OK) s = dev_alloc(&m->columns, nnz * isize(it));
    }
    if (s == SMB200_OK && (n_rows || nnz)) s = dev_alloc(&m->offsets, (n_rows + 1) * isize(it));
    if (s != SMB200_OK) { smb200_crs_free(m); return s; }
    // keep the padding defined: 128-bit tail loads read (and discard) it
    if (m->values) cudaMemsetAsync((char*)m->values + nnz * vsize(vt), 0, kPadBytes, ctx->stream);
    if (m->columns) cudaMemsetAsync((char*)m->columns + nnz * isize(it), 0, kPadBytes, ctx->stream);
    if (m->offsets) cudaMemsetAsync((char*)m->offsets + (n_rows + 1) * isize(it), 0, kPadBytes, ctx->stream);
    *out = m;
    return SMB200_OK;
}

smb200_status crs_finalize(smb200_crs* m, bool validate) {
    smb200_ctx* ctx = m->ctx;
    plan_free(m->plan);
    hostpipe_free(m->hp);
    m->max_row_len = 0;
    if (m->n_rows == 0) return SMB200_OK;
    CrsCheck* dchk = nullptr;
    SMB_CUDA(cudaMalloc(&dchk, sizeof(CrsCheck)));
    cudaMemsetAsync(dchk, 0, sizeof(CrsCheck), ctx->stream);
    const unsigned g = (unsigned)ctx->sm_count * 8;
    if (m->it == SMB200_U64) row_stats_kernel<uint64_t><<<g, 256, 0, ctx->stream>>>((const uint64_t*)m->offsets, m->n_rows, m->nnz, dchk);
    else row_stats_kernel<uint32_t><<<g, 256, 0, ctx->stream>>>((const uint32_t*)m->offsets, m->n_rows, m->nnz, dchk);
    count_launch();
    if (validate && m->nnz) {
        if (m->it == SMB200_U64) col_check_kernel<uint64_t><<<g, 256, 0, ctx->stream>>>((const uint64_t*)m->columns, m->nnz, m->n_cols, dchk);
        else col_check_kernel<uint32_t><<<g, 256, 0, ctx->stream>>>((const uint32_t*)m->columns, m->nnz, m->n_cols, dchk);
        count_launch();
    }
    CrsCheck h;
    cudaError_t e = cudaMemcpyAsync(&h, dchk, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dchk);
    SMB_CUDA(e);
    SMB_REQUIRE(!h.bad_offsets, SMB200_ERR_INVALID,
                "crs: offset_rows must start at 0, be non-decreasing and end at nnz=%llu", (unsigned long long)m->nnz);
    SMB_REQUIRE(!h.bad_columns, SMB200_ERR_INVALID, "crs: a column index is >= n_cols=%llu", (unsigned long long)m->n_cols);
    m->max_row_len = h.max_row_len;
    return SMB200_OK;
}

// ---- K8: IndexList -> CRS ------------------------------------------------------------------------------
// One thread per row walks the row's chain (indexlist.rs:94-112).  Pass 1 counts, the scan turns the
// counts into offset_rows, pass 2 copies (column, value) pairs in chain order (sparsemat_crs.rs:29-36).
// Pure integer / copy work: the result must equal the reference's arrays bit for bit.
template <class I>
__global__ void chain_len_kernel(const I* __restrict__ pos_start, const I* __restrict__ next, uint64_t n_rows,
                                 uint64_t nnz, I* __restrict__ lens, unsigned int* __restrict__ bad) {
    const I unset = (I)~(I)0;
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) { lens[r] = 0; return; }
    uint64_t count = 0;
    for (I p = pos_start[r]; p != unset; p = next[p]) {
        if ((uint64_t)p >= nnz || count >= nnz) { atomicOr(bad, 1u); break; }   // dangling link / cycle
        ++count;
    }
    lens[r] = (I)count;
}

template <class T, class I>
__global__ void chain_fill_kernel(const I* __restrict__ pos_start, const I* __restrict__ next,
                                  const I* __restrict__ columns, const T* __restrict__ values, uint64_t n_rows,
                                  const I* __restrict__ offsets, I* __restrict__ out_cols, T* __restrict__ out_vals) {
    const I unset = (I)~(I)0;
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    uint64_t k = (uint64_t)offsets[r];
    const uint64_t end = (uint64_t)offsets[r + 1];
    for (I p = pos_start[r]; p != unset && k < end; p = next[p], ++k) {
        out_cols[k] = columns[p];
        out_vals[k] = values[p];
    }
}

template <class T>
__global__ void scale_values_kernel(T* __restrict__ v, uint64_t n, T s) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) v[i] *= s;
}

template <class T, class I>
static smb200_status from_indexlist_impl(smb200_ctx* ctx, smb200_crs* m, const void* h_columns, const void* h_values,
                                         const void* h_pos_start, const void* h_next) {
    const uint64_t n_rows = m->n_rows, nnz = m->nnz;
    I *d_cols = nullptr, *d_pos = nullptr, *d_next = nullptr;
    T* d_vals = nullptr;
    unsigned int* d_bad = nullptr;
    smb200_status s = SMB200_OK;
    auto cleanup = [&] {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_cols); cudaFree(d_pos); cudaFree(d_next); cudaFree(d_vals); cudaFree(d_bad);
    };
#define K8_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { set_error("from_indexlist: %s", cudaGetErrorString(e_)); cleanup(); return e_ == cudaErrorMemoryAllocation ? SMB200_ERR_OOM : SMB200_ERR_CUDA; } } while (0)
    K8_CUDA(cudaMalloc(&d_cols, nnz * sizeof(I)));
    K8_CUDA(cudaMalloc(&d_next, nnz * sizeof(I)));
    K8_CUDA(cudaMalloc(&d_vals, nnz * sizeof(T)));
    K8_CUDA(cudaMalloc(&d_pos, n_rows * sizeof(I)));
    K8_CUDA(cudaMalloc(&d_bad, sizeof(unsigned int)));
    K8_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(unsigned int), ctx->stream));
    K8_CUDA(cudaMemcpyAsync(d_cols, h_columns, nnz * sizeof(I), cudaMemcpyHostToDevice, ctx->stream));
    K8_CUDA(cudaMemcpyAsync(d_next, h_next, nnz * sizeof(I), cudaMemcpyHostToDevice, ctx->stream));
    K8_CUDA(cudaMemcpyAsync(d_vals, h_values, nnz * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    K8_CUDA(cudaMemcpyAsync(d_pos, h_pos_start, n_rows * sizeof(I), cudaMemcpyHostToDevice, ctx->stream));
    const unsigned g = (unsigned)((n_rows + 1 + 255) / 256);
    chain_len_kernel<I><<<g, 256, 0, ctx->stream>>>(d_pos, d_next, n_rows, nnz, (I*)m->offsets, d_bad);
    count_launch();
    uint64_t total = 0;
    s = exclusive_scan_inplace(ctx, m->it, m->offsets, n_rows + 1, &total);
    if (s != SMB200_OK) { cleanup(); return s; }
    unsigned int bad = 0;
    K8_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->stream));
    K8_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bad || total != nnz) {
        cleanup();
        SMB_FAIL(SMB200_ERR_INVALID, "from_indexlist: chains do not cover the %llu entries exactly once (walked %llu%s)",
                 (unsigned long long)nnz, (unsigned long long)total, bad ? ", dangling link" : "");
    }
    chain_fill_kernel<T, I><<<g, 256, 0, ctx->stream>>>(d_pos, d_next, d_cols, d_vals, n_rows, (const I*)m->offsets,
                                                         (I*)m->columns, (T*)m->values);
    count_launch();
    K8_CUDA(cudaGetLastError());
#undef K8_CUDA
    cleanup();
    return SMB200_OK;
}

}  // namespace smb

using namespace smb;

extern "C" {

smb200_status smb200_crs_upload(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t n_rows, uint64_t n_cols,
                                uint64_t nnz, const void* values, const void* columns, const void* offset_rows,
                                smb200_crs** out) {
    SMB_REQUIRE(out, SMB200_ERR_INVALID, "crs_upload: out is NULL");
    SMB_REQUIRE((values && columns) || nnz == 0, SMB200_ERR_INVALID, "crs_upload: NULL values/columns");
    SMB_REQUIRE(offset_rows || n_rows == 0, SMB200_ERR_INVALID, "crs_upload: NULL offset_rows");
    smb200_crs* m = nullptr;
    SMB_TRY(crs_alloc(ctx, vt, it, n_rows, n_cols, nnz, &m));
    cudaError_t e = cudaSuccess;
    if (nnz) {
        e = cudaMemcpyAsync(m->values, values, nnz * vsize(vt), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->columns, columns, nnz * isize(it), cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e == cudaSuccess && n_rows) e = cudaMemcpyAsync(m->offsets, offset_rows, (n_rows + 1) * isize(it), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { smb200_crs_free(m); SMB_CUDA(e); }
    smb200_status s = crs_finalize(m, true);
    if (s != SMB200_OK) { smb200_crs_free(m); return s; }
    *out = m;
    return SMB200_OK;
}

smb200_status smb200_crs_from_indexlist(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t n_rows,
                                        uint64_t n_cols, uint64_t nnz, const void* columns, const void* values,
                                        const void* pos_start, const void* index_list, smb200_crs** out) {
    SMB_REQUIRE(out, SMB200_ERR_INVALID, "crs_from_indexlist: out is NULL");
    if (nnz == 0) {
        // sparsemat_crs.rs:25,47-49 — an IndexList without entries converts to SparseMatCRS::new(): 0 x 0.
        return crs_alloc(ctx, vt, it, 0, 0, 0, out);
    }
    SMB_REQUIRE(columns && values && pos_start && index_list, SMB200_ERR_INVALID, "crs_from_indexlist: NULL array");
    SMB_REQUIRE(n_rows > 0, SMB200_ERR_INVALID, "crs_from_indexlist: entries without rows");
    smb200_crs* m = nullptr;
    SMB_TRY(crs_alloc(ctx, vt, it, n_rows, n_cols, nnz, &m));
    smb200_status s;
    if (vt == SMB200_F64) {
        s = it == SMB200_U64 ? from_indexlist_impl<double, uint64_t>(ctx, m, columns, values, pos_start, index_list)
                             : from_indexlist_impl<double, uint32_t>(ctx, m, columns, values, pos_start, index_list);
    } else {
        s = it == SMB200_U64 ? from_indexlist_impl<float, uint64_t>(ctx, m, columns, values, pos_start, index_list)
                             : from_indexlist_impl<float, uint32_t>(ctx, m, columns, values, pos_start, index_list);
    }
    if (s == SMB200_OK) s = crs_finalize(m, true);
    if (s != SMB200_OK) { smb200_crs_free(m); return s; }
    *out = m;
    return SMB200_OK;
}

smb200_status smb200_crs_free(smb200_crs* m) {
    if (!m) return SMB200_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    plan_free(m->plan);
    hostpipe_free(m->hp);
    cg_free(m->cg);
    if (m->values) cudaFree(m->values);
    if (m->columns) cudaFree(m->columns);
    if (m->offsets) cudaFree(m->offsets);
    smb200_ctx* ctx = m->ctx;
    delete m;
    ctx_release(ctx);
    return SMB200_OK;
}

smb200_status smb200_crs_dims(const smb200_crs* m, uint64_t* out3) {
    SMB_REQUIRE(m && out3, SMB200_ERR_INVALID, "crs_dims: NULL argument");
    out3[0] = m->n_rows; out3[1] = m->n_cols - m->x_extra; out3[2] = m->nnz;
    return SMB200_OK;
}

smb200_status smb200_crs_types(const smb200_crs* m, int32_t* vt, int32_t* it) {
    SMB_REQUIRE(m, SMB200_ERR_INVALID, "crs_types: NULL argument");
    if (vt) *vt = m->vt;
    if (it) *it = m->it;
    return SMB200_OK;
}

smb200_status smb200_crs_download(const smb200_crs* m, void* values, void* columns, void* offset_rows) {
    SMB_REQUIRE(m, SMB200_ERR_INVALID, "crs_download: NULL argument");
    cudaStream_t st = m->ctx->stream;
    if (values && m->nnz) SMB_CUDA(cudaMemcpyAsync(values, m->values, m->nnz * vsize(m->vt), cudaMemcpyDeviceToHost, st));
    if (columns && m->nnz) SMB_CUDA(cudaMemcpyAsync(columns, m->columns, m->nnz * isize(m->it), cudaMemcpyDeviceToHost, st));
    if (offset_rows && m->offsets) SMB_CUDA(cudaMemcpyAsync(offset_rows, m->offsets, (m->n_rows + 1) * isize(m->it), cudaMemcpyDeviceToHost, st));
    SMB_CUDA(cudaStreamSynchronize(st));
    return SMB200_OK;
}

smb200_status smb200_crs_scale(smb200_crs* m, double s) {
    SMB_REQUIRE(m, SMB200_ERR_INVALID, "crs_scale: NULL argument");
    if (m->nnz == 0) return SMB200_OK;
    const unsigned g = (unsigned)m->ctx->sm_count * 8;
    if (m->vt == SMB200_F64) scale_values_kernel<double><<<g, 256, 0, m->ctx->stream>>>((double*)m->values, m->nnz, s);
    else scale_values_kernel<float><<<g, 256, 0, m->ctx->stream>>>((float*)m->values, m->nnz, (float)s);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    // a value-indexed ring plan multiplies the blocks' dictionaries, not m->values: every dictionary entry takes the same
    // single rounding as the value it stands for (range plans are never value-indexed, spmv.cu ring_plan)
    if (m->plan.vdict && m->plan.n_blocks) {
        const uint64_t nd = m->plan.n_blocks * 256;
        if (m->vt == SMB200_F64) scale_values_kernel<double><<<g, 256, 0, m->ctx->stream>>>((double*)m->plan.vdict, nd, s);
        else scale_values_kernel<float><<<g, 256, 0, m->ctx->stream>>>((float*)m->plan.vdict, nd, (float)s);
        count_launch();
        SMB_CUDA(cudaGetLastError());
    }
    // a band-split plan multiplies its own copies of the values (bandsplit.cu): they are scaled with the matrix, so that
    // `scale` affects every later product as in the reference (sparsemat_crs.rs:153-157)
    for (smb200_crs* part : m->plan.parts) SMB_TRY(smb200_crs_scale(part, s));
    for (auto& hp : m->hp.plans)
        for (smb200_crs* part : hp.parts) SMB_TRY(smb200_crs_scale(part, s));
    return SMB200_OK;
}

}  // extern "C"
