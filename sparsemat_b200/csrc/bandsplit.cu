// BANDSPLIT — column-band blocked SpMV for matrices whose x does not fit L2 and whose columns have no locality
// (BASELINE.json configs[2]: power-law rows, uniform random columns, N = 50 M, x = 400 MB).
//
// Why: with scattered columns every 8-byte gather of x misses L2 and costs a whole DRAM access (measured on the stream
// kernel, profiles/c3cg_ncu_r01.csv: 4.6x the algorithmic bytes, DRAM saturated at 5.1 TB/s, 13 % of the roofline).
// The only way to go faster is to make the gathers hit L2, i.e. to visit the columns band by band:
//
//     y = sum_b A_b x_b,    A_b = the entries of A whose column lies in band b = [b W, (b+1) W)
//
// Plan time (once per matrix, on the device): A is split into the n_bands CRS matrices A_b — same rows, entries of a row
// in their storage order, columns rebased to the band (so they fit u32 even for a u64 matrix: 12 instead of 16 bytes per
// f64 entry).  Product: one stream-kernel launch per band, the first writes y, the others add to it; x_b (W elements,
// about half of L2) stays cache-resident for the duration of its launch while A_b streams past it.
//
// Traffic per product: sum_b [nnz_b (sizeof T + 4) + (n_rows + 1) 4] + x + (2 n_bands - 1) n_rows sizeof T; for
// configs[2] with 7 bands about 17 GB against 68 GB of DRAM traffic today.  Rows are summed band-major (storage order
// inside a band), so results differ from the reference's storage-order sum by reassociation only — the same class as
// the multi-lane kernels, inside the north-star tolerance (1e-5 f32 / 1e-12 f64 relative to sum |a||x|).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace smb {

constexpr int kMaxBands = 64;

struct BandPtrs {
    void* offs[kMaxBands];     // per band: offset / count array of the part (IP)
    void* cols[kMaxBands];     // per band: columns of the part (IP)
    void* vals[kMaxBands];     // per band: values of the part (T)
};

// One thread per row: how many entries of the row fall into each band.  lens_b[r] for every band; lens_b[n_rows] = 0 so that
// the exclusive scan leaves the band's total there.
template <class I, class IP>
__global__ void band_count_kernel(const I* __restrict__ cols, const I* __restrict__ offs, uint64_t n_rows, uint64_t width,
                                  int n_bands, BandPtrs out) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r > n_rows) return;
    if (r == n_rows) {
        for (int b = 0; b < n_bands; ++b) static_cast<IP*>(out.offs[b])[r] = 0;
        return;
    }
    unsigned long long cnt[kMaxBands];
    for (int b = 0; b < n_bands; ++b) cnt[b] = 0;
    const uint64_t a = (uint64_t)offs[r], e = (uint64_t)offs[r + 1];
    for (uint64_t k = a; k < e; ++k) ++cnt[(uint64_t)cols[k] / width];
    for (int b = 0; b < n_bands; ++b) static_cast<IP*>(out.offs[b])[r] = (IP)cnt[b];
}

// One thread per row: copy the row's entries, in storage order, to the parts; columns rebased to their band.
template <class T, class I, class IP>
__global__ void band_fill_kernel(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                                 uint64_t n_rows, uint64_t width, int n_bands, BandPtrs out) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    unsigned long long pos[kMaxBands];
    for (int b = 0; b < n_bands; ++b) pos[b] = (unsigned long long)static_cast<const IP*>(out.offs[b])[r];
    const uint64_t a = (uint64_t)offs[r], e = (uint64_t)offs[r + 1];
    for (uint64_t k = a; k < e; ++k) {
        const uint64_t c = (uint64_t)cols[k];
        const uint64_t b = c / width;
        const unsigned long long at = pos[b]++;
        static_cast<IP*>(out.cols[b])[at] = (IP)(c - b * width);
        static_cast<T*>(out.vals[b])[at] = vals[k];
    }
}

template <class T, class I, class IP>
static smb200_status bandsplit_typed(smb200_crs* m, SpmvPlan& p, uint64_t width, int n_bands, int part_it) {
    smb200_ctx* ctx = m->ctx;
    const uint64_t n_rows = m->n_rows;
    // 1. counts -> offsets, in temporary arrays (the parts can only be allocated once their sizes are known)
    std::vector<IP*> tmp((size_t)n_bands, nullptr);
    auto free_tmp = [&] { for (IP* t : tmp) if (t) cudaFree(t); };
    BandPtrs ptrs;
    memset(&ptrs, 0, sizeof ptrs);
    for (int b = 0; b < n_bands; ++b) {
        cudaError_t e = cudaMalloc(&tmp[b], (n_rows + 1) * sizeof(IP));
        if (e != cudaSuccess) { free_tmp(); SMB_CUDA(e); }
        ptrs.offs[b] = tmp[b];
    }
    const unsigned g = (unsigned)((n_rows + 1 + 127) / 128);
    band_count_kernel<I, IP><<<g, 128, 0, ctx->stream>>>((const I*)m->columns, (const I*)m->offsets, n_rows, width, n_bands, ptrs);
    count_launch();
    std::vector<uint64_t> nnz_b((size_t)n_bands, 0);
    for (int b = 0; b < n_bands; ++b) {
        smb200_status s = exclusive_scan_inplace(ctx, part_it, tmp[b], n_rows + 1, &nnz_b[b]);
        if (s != SMB200_OK) { free_tmp(); return s; }
    }
    // 2. the parts
    uint64_t total = 0;
    for (int b = 0; b < n_bands; ++b) {
        const uint64_t lo = (uint64_t)b * width;
        const uint64_t w_b = std::min<uint64_t>(width, m->n_cols - lo);
        smb200_crs* part = nullptr;
        smb200_status s = crs_alloc(ctx, m->vt, part_it, n_rows, w_b, nnz_b[b], &part);
        if (s != SMB200_OK) { free_tmp(); return s; }
        p.parts.push_back(part);
        cudaError_t e = cudaMemcpyAsync(part->offsets, tmp[b], (n_rows + 1) * sizeof(IP), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) { free_tmp(); SMB_CUDA(e); }
        ptrs.offs[b] = part->offsets;
        ptrs.cols[b] = part->columns;
        ptrs.vals[b] = part->values;
        total += nnz_b[b];
    }
    if (total != m->nnz) { free_tmp(); SMB_FAIL(SMB200_ERR_INVALID, "bandsplit: the bands hold %llu of %llu entries", (unsigned long long)total, (unsigned long long)m->nnz); }
    band_fill_kernel<T, I, IP><<<(unsigned)((n_rows + 127) / 128), 128, 0, ctx->stream>>>(
        (const T*)m->values, (const I*)m->columns, (const I*)m->offsets, n_rows, width, n_bands, ptrs);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    free_tmp();
    SMB_CUDA(e);
    // 3. every part is an ordinary CRS matrix multiplied by the stream kernel (balanced over rows + non-zeros)
    for (smb200_crs* part : p.parts) {
        SMB_TRY(crs_finalize(part, false));
        part->want_variant = SMB200_SPMV_STREAM;
        part->want_flags = p.flags;
        SMB_TRY(plan_build(part));
    }
    return SMB200_OK;
}

// Band width in elements.  Measured on B200 (scripts/probes/l2_gather_probe.cu, scripts/c3_sweep.py): random 8-byte
// gathers keep hitting L2 up to a window of ~64 MB of the 126 MB while evict-first streams pass through it, and the product
// is fastest with the fewest bands that respect that (C3: 6 bands of 64 MB 6.25 ms, 7 of 55 MB 6.28, 9 of 43 MB 7.7, 12 of
// 32 MB 8.8): every extra band costs a read-modify-write of y and a pass over the row offsets.  So: the smallest number of
// equal bands of at most 0.53 L2.
uint64_t bandsplit_width(const smb200_crs* m) {
    const char* e = getenv("SMB200_BANDSPLIT_WIDTH");
    uint64_t w = (e && *e) ? strtoull(e, nullptr, 10) : 0;
    if (w == 0) {
        const uint64_t max_w = (uint64_t)((double)m->ctx->l2_bytes * 0.53) / vsize(m->vt);
        const uint64_t nb = (m->n_cols + max_w - 1) / (max_w ? max_w : 1);
        w = ((m->n_cols + (nb ? nb : 1) - 1) / (nb ? nb : 1) + 1023) / 1024 * 1024;
    }
    w = w / 1024 * 1024;
    if (w < 1024) w = 1024;
    // no more than kMaxBands bands
    const uint64_t min_w = ((m->n_cols + kMaxBands - 1) / kMaxBands + 1023) / 1024 * 1024;
    return std::max(w, min_w);
}

// Builds p.parts for the whole matrix.  p.variant is set to BANDSPLIT on success; a matrix that fits one band keeps `fallback`.
smb200_status bandsplit_build(smb200_crs* m, SpmvPlan& p, int fallback_variant) {
    const uint64_t width = bandsplit_width(m);
    const uint64_t n_bands = (m->n_cols + width - 1) / width;
    if (n_bands < 2 || m->nnz == 0) { p.variant = fallback_variant; return SMB200_OK; }
    SMB_REQUIRE(n_bands <= (uint64_t)kMaxBands, SMB200_ERR_UNSUPPORTED, "bandsplit: %llu bands", (unsigned long long)n_bands);
    // band-relative columns and per-band offsets fit u32 unless the matrix itself is beyond u32 offsets
    const int part_it = (m->nnz < 0xFFFFFFFEull && width <= 0xFFFFFFFEull && m->n_rows < 0xFFFFFFFEull) ? SMB200_U32 : m->it;
    p.band_width = width;
    smb200_status s;
    if (m->vt == SMB200_F64) {
        if (m->it == SMB200_U64) s = part_it == SMB200_U32 ? bandsplit_typed<double, uint64_t, uint32_t>(m, p, width, (int)n_bands, part_it)
                                                           : bandsplit_typed<double, uint64_t, uint64_t>(m, p, width, (int)n_bands, part_it);
        else s = bandsplit_typed<double, uint32_t, uint32_t>(m, p, width, (int)n_bands, part_it);
    } else {
        if (m->it == SMB200_U64) s = part_it == SMB200_U32 ? bandsplit_typed<float, uint64_t, uint32_t>(m, p, width, (int)n_bands, part_it)
                                                           : bandsplit_typed<float, uint64_t, uint64_t>(m, p, width, (int)n_bands, part_it);
        else s = bandsplit_typed<float, uint32_t, uint32_t>(m, p, width, (int)n_bands, part_it);
    }
    if (s != SMB200_OK) {
        for (smb200_crs* part : p.parts) smb200_crs_free(part);
        p.parts.clear();
        return s;
    }
    p.variant = SMB200_SPMV_BANDSPLIT;
    p.n_blocks = 0;
    for (smb200_crs* part : p.parts) p.n_blocks += part->plan.n_blocks;
    return SMB200_OK;
}

// Bytes one band-split product moves (smb200_plan_info.stream_bytes).
uint64_t bandsplit_stream_bytes(const smb200_crs* m, const SpmvPlan& p) {
    uint64_t bytes = m->n_cols * vsize(m->vt);                                            // every band of x once
    for (const smb200_crs* part : p.parts)
        bytes += part->nnz * (vsize(part->vt) + isize(part->it)) + (part->n_rows + 1) * isize(part->it);
    const uint64_t nb = p.parts.size();
    bytes += (2 * nb - 1) * m->n_rows * vsize(m->vt);                                     // y: written by every band, read by all but the first
    return bytes;
}

}  // namespace smb
