// Jacobi-preconditioned CG — additive, NOT in the reference (the LinearSolver trait of
// linearsolver.rs:6-10 has one implementation, the unpreconditioned ConjugateGradient).  First version: composed from
// the validated primitives (smb200_spmv, vec_dot, vec_axpy, vec_scale_add) plus two small kernels (diagonal, elementwise
// product); the scalars alpha / beta live on the host, so every iteration synchronises three times.  The arithmetic
// follows the reference's conventions: everything in T, products rounded before they are added (never an FMA), the stop
// test sqrt(r.r) < tol (absolute; `relative` scales tol by ||b||) evaluated in f64 after the x / r update.
//
//   r = b - A x;  z = D^-1 r;  p = z;  rz = r.z
//   loop:  Ap = A p;  alpha = rz / p.Ap;  x += p alpha;  r -= Ap alpha;  stop if sqrt(r.r) < tol
//          z = D^-1 r;  beta = r.z / rz;  p = p beta + z
//
// Checked on hardware against the same recurrence in numpy (tests/test_gpu_pcg.py); not yet tuned: a device-scalar,
// graph-replayed version like cg.cu's is the obvious next step.
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

// z[i] = w[i] * r[i]
template <class T>
__global__ void multiply_kernel(const T* __restrict__ w, const T* __restrict__ r, uint64_t n, T* __restrict__ z) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) z[i] = mul_rn(w[i], r[i]);
}

template <class T>
static smb200_status multiply_launch(smb200_ctx* ctx, const void* w, const void* r, uint64_t n, void* z) {
    if (n == 0) return SMB200_OK;
    uint64_t g = (n + 255) / 256, cap = (uint64_t)ctx->sm_count * 8;
    multiply_kernel<T><<<(unsigned)(g < cap ? g : cap), 256, 0, ctx->stream>>>((const T*)w, (const T*)r, n, (T*)z);
    count_launch();
    SMB_CUDA(cudaGetLastError());
    return SMB200_OK;
}

struct VecGuard {          // frees the solver's temporaries on every exit path
    smb200_vec* v[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    ~VecGuard() { for (smb200_vec* p : v) if (p) smb200_vec_free(p); }
};

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
    const smb200_vtype vt = (smb200_vtype)a->vt;
    const uint64_t launches0 = g_launches;
    smb200_cg_stats st_out;
    memset(&st_out, 0, sizeof st_out);
    if (n == 0) { if (stats) *stats = st_out; return SMB200_OK; }

    VecGuard tmp;
    for (int k = 0; k < 5; ++k) SMB_TRY(smb200_vec_create(ctx, vt, n, &tmp.v[k]));
    smb200_vec *r = tmp.v[0], *z = tmp.v[1], *p = tmp.v[2], *ap = tmp.v[3], *dinv = tmp.v[4];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    SMB_CUDA(cudaEventCreate(&ev0));
    if (cudaEventCreate(&ev1) != cudaSuccess) { cudaEventDestroy(ev0); SMB_FAIL(SMB200_ERR_CUDA, "pcg_jacobi_solve: cudaEventCreate failed"); }
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
    SMB_CUDA(cudaEventRecord(ev0, ctx->stream));

    // D^-1
    SMB_TRY(smb200_crs_diagonal(a, dinv));
    unsigned long long* d_zeros = nullptr;
    SMB_CUDA(cudaMalloc(&d_zeros, sizeof(unsigned long long)));
    cudaMemsetAsync(d_zeros, 0, sizeof(unsigned long long), ctx->stream);
    {
        const unsigned g = (unsigned)((n + 255) / 256);
        if (vt == SMB200_F64) reciprocal_kernel<double><<<g, 256, 0, ctx->stream>>>((const double*)dinv->d, n, (double*)dinv->d, d_zeros);
        else reciprocal_kernel<float><<<g, 256, 0, ctx->stream>>>((const float*)dinv->d, n, (float*)dinv->d, d_zeros);
        count_launch();
    }
    unsigned long long h_zeros = 0;
    cudaError_t e = cudaMemcpyAsync(&h_zeros, d_zeros, sizeof h_zeros, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_zeros);
    SMB_CUDA(e);
    SMB_REQUIRE(h_zeros == 0, SMB200_ERR_INVALID, "pcg_jacobi_solve: %llu rows have no (or a zero) diagonal entry", h_zeros);

    auto mul = [&](const smb200_vec* w, const smb200_vec* rr, smb200_vec* out) {
        return vt == SMB200_F64 ? multiply_launch<double>(ctx, w->d, rr->d, n, out->d) : multiply_launch<float>(ctx, w->d, rr->d, n, out->d);
    };
    // scalars are rounded to T like the reference's (alpha, beta and the dot products are `T`)
    auto to_t = [&](double v) { return vt == SMB200_F64 ? v : (double)(float)v; };

    double bb = 0.0, threshold = tol;
    if (relative) { SMB_TRY(smb200_vec_dot(b, b, &bb)); threshold = tol * std::sqrt(bb); }
    // r = b - A x   (b.clone() - mat.mvp(x), linearsolver.rs:38)
    SMB_TRY(smb200_spmv(a, x, ap));
    SMB_TRY(smb200_vec_copy(r, b));
    SMB_TRY(smb200_vec_sub(r, ap));
    SMB_TRY(mul(dinv, r, z));
    SMB_TRY(smb200_vec_copy(p, z));
    double rz = 0.0, rr = 0.0;
    SMB_TRY(smb200_vec_dot(r, z, &rz));
    SMB_TRY(smb200_vec_dot(r, r, &rr));
    double res = std::sqrt(rr);
    uint64_t it = 0;
    int converged = 0;
    for (uint64_t k = 0; k < iter_max; ++k) {
        SMB_TRY(smb200_spmv(a, p, ap));
        double pap = 0.0;
        SMB_TRY(smb200_vec_dot(p, ap, &pap));
        const double alpha = to_t(rz / pap);
        SMB_TRY(smb200_vec_axpy(x, alpha, p));                 // *x += p.clone() * alpha
        SMB_TRY(smb200_vec_axpy(r, -alpha, ap));               // r -= Ap * alpha  (the product's sign flip is exact)
        SMB_TRY(smb200_vec_dot(r, r, &rr));
        it = k + 1;
        res = std::sqrt(rr);
        if (res < threshold) { converged = 1; break; }
        SMB_TRY(mul(dinv, r, z));
        double rz_new = 0.0;
        SMB_TRY(smb200_vec_dot(r, z, &rz_new));
        const double beta = to_t(rz_new / rz);
        SMB_TRY(smb200_vec_scale_add(p, beta, z));             // p = p * beta + z
        rz = rz_new;
    }
    SMB_CUDA(cudaEventRecord(ev1, ctx->stream));
    SMB_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.f;
    SMB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    st_out.iterations = it;
    st_out.final_residual = res;
    st_out.converged = converged;
    st_out.device_ms = ms;
    st_out.launches = g_launches - launches0;
    if (stats) *stats = st_out;
    return SMB200_OK;
}

}  // extern "C"
