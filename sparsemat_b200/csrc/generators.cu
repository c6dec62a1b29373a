// Device-side generators for the BASELINE.json workloads, so that the 10^8..10^9
// non-zero matrices never cross PCIe.  The host oracle (oracle/generators.hpp) reproduces the same
// arrays bit for bit; tests compare them.
#include "common.cuh"
#include "rng.cuh"

namespace smb {

struct LaplaceGeom {
    uint64_t nx, ny, nz, plane;
    uint64_t row_lo, row_hi;
    // column numbering: global, or local = [owned | pad | ghost plane below | ghost plane above], ghosts from ghost_base on
    int local_cols;
    uint64_t n_lo_ghost;
    uint64_t ghost_base;
};

__device__ __forceinline__ uint64_t laplace_col(const LaplaceGeom& g, uint64_t c) {
    if (!g.local_cols) return c;
    if (c >= g.row_lo && c < g.row_hi) return c - g.row_lo;
    if (c < g.row_lo) return g.ghost_base + (c - (g.row_lo - g.n_lo_ghost));
    return g.ghost_base + g.n_lo_ghost + (c - g.row_hi);
}

template <class I>
__global__ void laplace_len_kernel(LaplaceGeom g, I* __restrict__ lens) {
    const uint64_t n = g.row_hi - g.row_lo;
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t > n) return;
    if (t == n) { lens[t] = 0; return; }
    const uint64_t r = g.row_lo + t;
    const uint64_t ix = r % g.nx, iy = (r / g.nx) % g.ny, iz = r / g.plane;
    uint64_t c = 1 + (ix > 0) + (ix + 1 < g.nx) + (iy > 0) + (iy + 1 < g.ny);
    if (g.nz > 1) c += (iz > 0) + (iz + 1 < g.nz);
    lens[t] = (I)c;
}

template <class T, class I>
__global__ void laplace_fill_kernel(LaplaceGeom g, const I* __restrict__ offsets, T* __restrict__ values, I* __restrict__ columns) {
    const uint64_t n = g.row_hi - g.row_lo;
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint64_t r = g.row_lo + t;
    const uint64_t ix = r % g.nx, iy = (r / g.nx) % g.ny, iz = r / g.plane;
    const T diag = (T)(g.nz > 1 ? 6.0 : 4.0), off = (T)(-1.0);
    uint64_t k = (uint64_t)offsets[t];
    if (g.nz > 1 && iz > 0)        { columns[k] = (I)laplace_col(g, r - g.plane); values[k++] = off; }
    if (iy > 0)                    { columns[k] = (I)laplace_col(g, r - g.nx);    values[k++] = off; }
    if (ix > 0)                    { columns[k] = (I)laplace_col(g, r - 1);       values[k++] = off; }
    columns[k] = (I)laplace_col(g, r); values[k++] = diag;
    if (ix + 1 < g.nx)             { columns[k] = (I)laplace_col(g, r + 1);       values[k++] = off; }
    if (iy + 1 < g.ny)             { columns[k] = (I)laplace_col(g, r + g.nx);    values[k++] = off; }
    if (g.nz > 1 && iz + 1 < g.nz) { columns[k] = (I)laplace_col(g, r + g.plane); values[k++] = off; }
}

static uint64_t laplace_nnz_rows(uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_lo, uint64_t row_hi) {
    // closed form over whole z-planes plus a loop over the ragged ends (host, O(plane) at worst)
    const uint64_t plane = nx * ny;
    auto row_len = [&](uint64_t r) {
        const uint64_t ix = r % nx, iy = (r / nx) % ny, iz = r / plane;
        uint64_t c = 1 + (ix > 0) + (ix + 1 < nx) + (iy > 0) + (iy + 1 < ny);
        if (nz > 1) c += (iz > 0) + (iz + 1 < nz);
        return c;
    };
    const uint64_t per_plane_inplane = plane + 2 * (nx - 1) * ny + 2 * nx * (ny - 1);
    uint64_t total = 0, r = row_lo;
    while (r < row_hi && r % plane != 0) total += row_len(r++);
    while (r + plane <= row_hi) {
        const uint64_t iz = r / plane;
        total += per_plane_inplane + (nz > 1 ? plane * ((iz > 0) + (iz + 1 < nz)) : 0);
        r += plane;
    }
    while (r < row_hi) total += row_len(r++);
    return total;
}

// Shared by smb200_gen_laplace and the distributed z-slab builder (dist.cu).
smb200_status gen_laplace_block(smb200_ctx* ctx, int vt, int it, uint64_t nx, uint64_t ny, uint64_t nz, uint64_t row_lo,
                                uint64_t row_hi, int local_cols, uint64_t n_lo_ghost, uint64_t n_cols_out,
                                smb200_crs** out, uint64_t ghost_base) {
    SMB_REQUIRE(nx && ny && nz, SMB200_ERR_INVALID, "gen_laplace: empty grid");
    const uint64_t N = nx * ny * nz;
    SMB_REQUIRE(row_lo <= row_hi && row_hi <= N, SMB200_ERR_INVALID, "gen_laplace: bad row range");
    const uint64_t n = row_hi - row_lo;
    const uint64_t nnz = laplace_nnz_rows(nx, ny, nz, row_lo, row_hi);
    smb200_crs* m = nullptr;
    SMB_TRY(crs_alloc(ctx, vt, it, n, n_cols_out, nnz, &m));
    if (n == 0) { *out = m; return SMB200_OK; }
    LaplaceGeom g{nx, ny, nz, nx * ny, row_lo, row_hi, local_cols, n_lo_ghost, ghost_base ? ghost_base : row_hi - row_lo};
    const unsigned grid = (unsigned)((n + 1 + 255) / 256);
    if (it == SMB200_U64) laplace_len_kernel<uint64_t><<<grid, 256, 0, ctx->stream>>>(g, (uint64_t*)m->offsets);
    else laplace_len_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>(g, (uint32_t*)m->offsets);
    count_launch();
    uint64_t total = 0;
    smb200_status s = exclusive_scan_inplace(ctx, it, m->offsets, n + 1, &total);
    if (s == SMB200_OK && total != nnz) { set_error("gen_laplace: internal nnz mismatch %llu vs %llu", (unsigned long long)total, (unsigned long long)nnz); s = SMB200_ERR_INVALID; }
    if (s != SMB200_OK) { smb200_crs_free(m); return s; }
    if (vt == SMB200_F64) {
        if (it == SMB200_U64) laplace_fill_kernel<double, uint64_t><<<grid, 256, 0, ctx->stream>>>(g, (const uint64_t*)m->offsets, (double*)m->values, (uint64_t*)m->columns);
        else laplace_fill_kernel<double, uint32_t><<<grid, 256, 0, ctx->stream>>>(g, (const uint32_t*)m->offsets, (double*)m->values, (uint32_t*)m->columns);
    } else {
        if (it == SMB200_U64) laplace_fill_kernel<float, uint64_t><<<grid, 256, 0, ctx->stream>>>(g, (const uint64_t*)m->offsets, (float*)m->values, (uint64_t*)m->columns);
        else laplace_fill_kernel<float, uint32_t><<<grid, 256, 0, ctx->stream>>>(g, (const uint32_t*)m->offsets, (float*)m->values, (uint32_t*)m->columns);
    }
    count_launch();
    if (cudaGetLastError() != cudaSuccess) { smb200_crs_free(m); SMB_FAIL(SMB200_ERR_CUDA, "gen_laplace: launch failed"); }
    s = crs_finalize(m, false);
    if (s != SMB200_OK) { smb200_crs_free(m); return s; }
    *out = m;
    return SMB200_OK;
}

template <class I>
__global__ void powerlaw_len_kernel(uint64_t seed, uint64_t n_rows, uint64_t max_len, I* __restrict__ lens) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i > n_rows) return;
    if (i == n_rows) { lens[i] = 0; return; }
    const double u = rng::u01(rng::rng1(seed, i));
    double l = floor(8.0 / sqrt(u));
    if (!(l >= 1.0)) l = 1.0;
    if (l > (double)max_len) l = (double)max_len;
    lens[i] = (I)(uint64_t)l;
}

// one warp per row
template <class T, class I>
__global__ void powerlaw_fill_kernel(uint64_t seed_col, uint64_t seed_val, uint64_t n_rows, uint64_t n_cols,
                                     const I* __restrict__ offsets, T* __restrict__ values, I* __restrict__ columns) {
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp; i < n_rows; i += n_warps) {
        const uint64_t b = (uint64_t)offsets[i], e = (uint64_t)offsets[i + 1];
        const uint64_t hc = rng::rng1(seed_col, i), hv = rng::rng1(seed_val, i);
        for (uint64_t k = b + lane; k < e; k += 32) {
            columns[k] = (I)(rng::mix(hc + (k - b)) % n_cols);
            values[k] = (T)rng::pm1(rng::mix(hv + (k - b)));
        }
    }
}

}  // namespace smb

using namespace smb;

extern "C" {

smb200_status smb200_gen_laplace(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t nx, uint64_t ny,
                                 uint64_t nz, uint64_t row_lo, uint64_t row_hi, smb200_crs** out) {
    SMB_REQUIRE(ctx && out, SMB200_ERR_INVALID, "gen_laplace: NULL argument");
    return gen_laplace_block(ctx, vt, it, nx, ny, nz, row_lo, row_hi, 0, 0, nx * ny * nz, out);
}

smb200_status smb200_gen_powerlaw(smb200_ctx* ctx, smb200_vtype vt, smb200_itype it, uint64_t n_rows, uint64_t n_cols,
                                  uint64_t seed_len, uint64_t seed_col, uint64_t seed_val, uint64_t max_len,
                                  smb200_crs** out) {
    SMB_REQUIRE(ctx && out, SMB200_ERR_INVALID, "gen_powerlaw: NULL argument");
    SMB_REQUIRE(n_rows && n_cols && max_len, SMB200_ERR_INVALID, "gen_powerlaw: empty shape");
    *out = nullptr;
    SMB_CUDA(cudaSetDevice(ctx->device));
    // pass 1: lengths -> offsets (temporary, u64 so that nnz is known before the typed allocation)
    uint64_t* d_off = nullptr;
    SMB_CUDA(cudaMalloc(&d_off, (n_rows + 1) * sizeof(uint64_t)));
    const unsigned grid = (unsigned)((n_rows + 1 + 255) / 256);
    powerlaw_len_kernel<uint64_t><<<grid, 256, 0, ctx->stream>>>(seed_len, n_rows, max_len, d_off);
    count_launch();
    uint64_t nnz = 0;
    smb200_status s = exclusive_scan_inplace(ctx, SMB200_U64, d_off, n_rows + 1, &nnz);
    if (s != SMB200_OK) { cudaFree(d_off); return s; }
    smb200_crs* m = nullptr;
    s = crs_alloc(ctx, vt, it, n_rows, n_cols, nnz, &m);
    if (s != SMB200_OK) { cudaFree(d_off); return s; }
    if (it == SMB200_U64) {
        cudaMemcpyAsync(m->offsets, d_off, (n_rows + 1) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream);
    } else {
        powerlaw_len_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>(seed_len, n_rows, max_len, (uint32_t*)m->offsets);
        count_launch();
        uint64_t chk = 0;
        s = exclusive_scan_inplace(ctx, SMB200_U32, m->offsets, n_rows + 1, &chk);
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_off);
    if (s != SMB200_OK) { smb200_crs_free(m); return s; }
    const unsigned fgrid = (unsigned)ctx->sm_count * 16;
    if (vt == SMB200_F64) {
        if (it == SMB200_U64) powerlaw_fill_kernel<double, uint64_t><<<fgrid, 256, 0, ctx->stream>>>(seed_col, seed_val, n_rows, n_cols, (const uint64_t*)m->offsets, (double*)m->values, (uint64_t*)m->columns);
        else powerlaw_fill_kernel<double, uint32_t><<<fgrid, 256, 0, ctx->stream>>>(seed_col, seed_val, n_rows, n_cols, (const uint32_t*)m->offsets, (double*)m->values, (uint32_t*)m->columns);
    } else {
        if (it == SMB200_U64) powerlaw_fill_kernel<float, uint64_t><<<fgrid, 256, 0, ctx->stream>>>(seed_col, seed_val, n_rows, n_cols, (const uint64_t*)m->offsets, (float*)m->values, (uint64_t*)m->columns);
        else powerlaw_fill_kernel<float, uint32_t><<<fgrid, 256, 0, ctx->stream>>>(seed_col, seed_val, n_rows, n_cols, (const uint32_t*)m->offsets, (float*)m->values, (uint32_t*)m->columns);
    }
    count_launch();
    if (cudaGetLastError() != cudaSuccess) { smb200_crs_free(m); SMB_FAIL(SMB200_ERR_CUDA, "gen_powerlaw: launch failed"); }
    s = crs_finalize(m, false);
    if (s != SMB200_OK) { smb200_crs_free(m); return s; }
    *out = m;
    return SMB200_OK;
}

}  // extern "C"
