// Jacobi-preconditioned CG — additive, NOT in the reference (the LinearSolver trait of
// linearsolver.rs:6-10 has one implementation, the unpreconditioned ConjugateGradient).  The solver is cg.cu's: the same
// three fused kernels per iteration with the inverse diagonal as one more operand (z = D^-1 r is never stored), alpha, beta,
// r.z, r.r and the stop test on the device, iterations replayed from a CUDA graph — no host synchronisation in the loop.
// This file holds the diagonal extraction and the entry point.  The arithmetic follows the reference's conventions:
// everything in T, products rounded before they are added (never an FMA), the stop test sqrt(r.r) < tol (absolute;
// `relative` scales tol by ||b||) evaluated in f64 after the x / r update.
//
//   r = b - A x;  z = D^-1 r;  p = z;  rz = r.z
//   loop:  Ap = A p;  alpha = rz / p.Ap;  x += p alpha;  r -= Ap alpha;  stop if sqrt(r.r) < tol
//          z = D^-1 r;  beta = r.z / rz;  p = p beta + z
//
// Algorithmic bytes per iteration: SpMV + 11 N sizeof(T).  Checked on hardware against the same recurrence in numpy
// (tests/test_gpu_pcg.py; there is no reference implementation to compare with).
#include "common.cuh"
#include "reduce.cuh"

#include <cmath>

namespace smb {

// d[i] = A[i][i]: the first stored entry of row i whose column is i (what SparseMatCRS::get(i, i) returns,
// sparsemat_crs.rs:54-67,136-143), zero when the row holds none.
template <class T, class I>
__global__ void diagonal_kernel(const T* __restrict__ vals, const I* __restrict__ cols, const I* __restrict__ offs,
                                uint64_t n_rows, T* __restrict__ d) {
    const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    T v = T(0);
    for (uint64_t k = (uint64_t)offs[r], e = (uint64_t)offs[r + 1]; k < e; ++k)
        if ((uint64_t)cols[k] == r) { v = vals[k]; break; }
    d[r] = v;
}

// inv[i] = 1 / d[i]; counts zeros (a zero pivot makes the preconditioner undefined).
template <class T>
__global__ void reciprocal_kernel(const T* d, uint64_t n, T* inv, unsigned long long* zeros) {   // in place: inv may be d
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T v = d[i];
    if (v == T(0)) { atomicAdd(zeros, 1ull); inv[i] = T(0); }
    else inv[i] = T(1) / v;
}

}  // namespace smb

using namespace smb;

extern "C" {

smb200_status smb200_crs_diagonal(const smb200_crs* a, smb200_vec* d) {
    SMB_REQUIRE(a && d, SMB200_ERR_INVALID, "crs_diagonal: NULL argument");
    SMB_REQUIRE(d->vt == a->vt && d->ctx == a->ctx, SMB200_ERR_INVALID, "crs_diagonal: vector of another type or context");
    SMB_REQUIRE(d->n >= a->n_rows, SMB200_ERR_DIM, "Dimension mismatch");
    if (a->n_rows == 0) return SMB200_OK;
    const unsigned g = (unsigned)((a->n_rows + 255) / 256);
    cudaStream_t st = a->ctx->stream;
    if (a->vt == SMB200_F64) {
        if (a->it == SMB200_U64) diagonal_kernel<double, uint64_t><<<g, 256, 0, st>>>((const double*)a->values, (const uint64_t*)a->columns, (const uint64_t*)a->offsets, a->n_rows, (double*)d->d);
        else diagonal_kernel<double, uint32_t><<<g, 256, 0, st>>>((const double*)a->values, (const uint32_t*)a->columns, (const uint32_t*)a->offsets, a->n_rows, (double*)d->d);
    } else {
        if (a->it == SMB200_U64) diagonal_kernel<float, uint64_t><<<g, 256, 0, st>>>((const float*)a->values, (const uint64_t*)a->columns, (const uint64_t*)a->offsets, a->n_rows, (float*)d->d);
        else diagonal_kernel<float, uint32_t><<<g, 256, 0, st>>>((const float*)a->values, (const uint32_t*)a->columns, (const uint32_t*)a->offsets, a->n_rows, (float*)d->d);
    }
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

smb200_status smb200_pcg_jacobi_solve(smb200_crs* a, const smb200_vec* b, smb200_vec* x, double tol, int32_t relative,
                                      uint64_t iter_max, smb200_cg_stats* stats) {
    SMB_REQUIRE(a && b && x, SMB200_ERR_INVALID, "pcg_jacobi_solve: NULL argument");
    SMB_REQUIRE(b->vt == a->vt && x->vt == a->vt, SMB200_ERR_INVALID, "pcg_jacobi_solve: value types differ");
    SMB_REQUIRE(b->ctx == a->ctx && x->ctx == a->ctx, SMB200_ERR_INVALID, "pcg_jacobi_solve: operands belong to different contexts");
    // the reference's checks, in its order (linearsolver.rs:30-36)
    SMB_REQUIRE(a->n_rows == a->n_cols, SMB200_ERR_NOT_SQUARE, "Matrix is not symmetric");
    SMB_REQUIRE(a->n_rows == b->n && a->n_rows == x->n, SMB200_ERR_SIZE_MISMATCH, "Matrix and vector size mismatch");
    smb200_ctx* ctx = a->ctx;
    const uint64_t n = a->n_rows;
    if (n == 0) { if (stats) memset(stats, 0, sizeof *stats); return SMB200_OK; }
    SMB_CUDA(cudaSetDevice(ctx->device));
    // D^-1, kept with the matrix's solver workspace (recomputed per solve: the values may have been scaled since)
    CgWork& w = a->cg;
    SMB_TRY(cg_prepare(ctx, w, a->vt, n, n, iter_max));       // first: (re)allocating the workspace releases an older dinv
    if (!w.dinv || w.dinv_n != n) {
        cudaStreamSynchronize(ctx->stream);
        if (w.dinv) { cudaFree(w.dinv); w.dinv = nullptr; }
        SMB_TRY(dev_alloc(&w.dinv, n * vsize(a->vt)));
        w.dinv_n = n;
    }
    smb200_vec dv;
    dv.ctx = ctx; dv.vt = a->vt; dv.n = n; dv.cap = n; dv.d = w.dinv; dv.owned = false;
    SMB_TRY(smb200_crs_diagonal(a, &dv));
    unsigned long long* d_zeros = nullptr;
    SMB_CUDA(cudaMalloc(&d_zeros, sizeof(unsigned long long)));
    cudaMemsetAsync(d_zeros, 0, sizeof(unsigned long long), ctx->stream);
    {
        const unsigned g = (unsigned)((n + 255) / 256);
        if (a->vt == SMB200_F64) reciprocal_kernel<double><<<g, 256, 0, ctx->stream>>>((const double*)w.dinv, n, (double*)w.dinv, d_zeros);
        else reciprocal_kernel<float><<<g, 256, 0, ctx->stream>>>((const float*)w.dinv, n, (float*)w.dinv, d_zeros);
        count_launch();
    }
    unsigned long long h_zeros = 0;
    cudaError_t e = cudaMemcpyAsync(&h_zeros, d_zeros, sizeof h_zeros, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_zeros);
    SMB_CUDA(e);
    SMB_REQUIRE(h_zeros == 0, SMB200_ERR_INVALID, "pcg_jacobi_solve: %llu rows have no (or a zero) diagonal entry", h_zeros);
    // the CG machinery of cg.cu with one more operand: scalars on the device, iterations replayed from a CUDA graph, no
    // host synchronisation inside the loop
    return cg_solve_impl(a, b, x, tol, relative, iter_max, stats, w.dinv);
}

}  // extern "C"
