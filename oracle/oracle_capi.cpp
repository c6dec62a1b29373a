// ORACLE — TEST INFRASTRUCTURE ONLY (see sparsemat_oracle.hpp).
//
// extern "C" surface over the restatement so that tests/ and bench.py (cpu_baseline / --impl
// reference) can drive it through ctypes.  Functions are suffixed with the value/index combo:
// _f32u32, _f64u32, _f32u64, _f64u64.  Return value 0 = ok, 1 = the reference would have panicked
// (message retrievable with orc_last_panic()).
#include "generators.hpp"
#include "sparsemat_oracle.hpp"

#include <chrono>
#include <cstring>
#include <string>

using namespace oracle;

namespace {
thread_local std::string g_panic;

template <class F> int guarded(F&& f) {
    try { f(); return 0; }
    catch (const Panic& p) { g_panic = p.what(); return 1; }
    catch (const std::exception& e) { g_panic = std::string("oracle internal: ") + e.what(); return 2; }
}

template <class T, class I>
int il_apply(SparseMatIndexList<T, I>* m, std::uint64_t n, const std::uint64_t* i, const std::uint64_t* j,
             const T* v, const std::uint8_t* op) {
    return guarded([&] {
        for (std::uint64_t k = 0; k < n; ++k) {
            if (op && op[k]) m->add_to(i[k], j[k], v[k]);
            else m->set(i[k], j[k], v[k]);
        }
    });
}

template <class T, class I>
void il_export(const SparseMatIndexList<T, I>* m, I* columns, T* values, I* pos_start, I* next) {
    if (m->nnz()) {
        std::memcpy(columns, m->columns.data(), m->nnz() * sizeof(I));
        std::memcpy(values, m->values.data(), m->nnz() * sizeof(T));
        std::memcpy(next, m->chains.next.data(), m->nnz() * sizeof(I));
    }
    if (m->n_rows()) std::memcpy(pos_start, m->chains.pos_start.data(), m->n_rows() * sizeof(I));
}

template <class T, class I>
void crs_export(const SparseMatCRS<T, I>* m, T* values, I* columns, I* offsets) {
    if (m->nnz()) {
        std::memcpy(values, m->values.data(), m->nnz() * sizeof(T));
        std::memcpy(columns, m->columns.data(), m->nnz() * sizeof(I));
    }
    if (!m->offset_rows.empty()) std::memcpy(offsets, m->offset_rows.data(), m->offset_rows.size() * sizeof(I));
}

// Raw-array IndexList -> CRS (sparsemat_crs.rs:24-36) for large inputs that never existed as an
// oracle object (e.g. arrays produced by the product's host-side assembler).
template <class T, class I>
int to_crs_raw(std::uint64_t n_rows, std::uint64_t nnz, const I* columns, const T* values, const I* pos_start,
               const I* next, T* out_values, I* out_columns, I* out_offsets) {
    return guarded([&] {
        std::uint64_t k = 0;
        for (std::uint64_t r = 0; r < n_rows; ++r) {
            out_offsets[r] = static_cast<I>(k);
            for (I p = pos_start[r]; p != unset<I>(); p = next[static_cast<std::size_t>(p)]) {
                if (static_cast<std::uint64_t>(p) >= nnz || k >= nnz) throw Panic("index out of bounds: chain");
                out_columns[k] = columns[static_cast<std::size_t>(p)];
                out_values[k] = values[static_cast<std::size_t>(p)];
                ++k;
            }
        }
        out_offsets[n_rows] = static_cast<I>(k);
    });
}

template <class T, class I>
int mvp_checked(std::uint64_t n_rows, std::uint64_t n_cols, std::uint64_t nnz, const T* values, const I* columns,
                const I* offsets, const T* x, std::uint64_t nx, T* y) {
    return guarded([&] {
        CrsView<T, I> a{n_rows, n_cols, nnz, values, columns, offsets};
        DenseVec<T> xv(std::vector<T>(x, x + nx));
        DenseVec<T> yv = mvp(a, xv);
        std::memcpy(y, yv.v.data(), yv.v.size() * sizeof(T));
    });
}

template <class T, class I>
int bilinear_raw(std::uint64_t n_rows, std::uint64_t n_cols, std::uint64_t nnz, const T* values, const I* columns,
                 const I* offsets, const T* lhs, std::uint64_t nl, const T* rhs, std::uint64_t nr, T* out) {
    return guarded([&] {
        CrsView<T, I> a{n_rows, n_cols, nnz, values, columns, offsets};
        DenseVec<T> l(std::vector<T>(lhs, lhs + nl)), r(std::vector<T>(rhs, rhs + nr));
        *out = bilinear(a, l, r);
    });
}

// linearsolver.rs:27-61 on raw CRS arrays.  Same arithmetic as ConjugateGradient::solve in the
// header (sequential folds, scale-then-add), written on flat arrays so that the CPU baseline is
// not slowed by per-element bounds checks the optimiser cannot remove.
template <class T, class I>
int cg_raw(std::uint64_t n_rows, std::uint64_t n_cols, const T* values, const I* columns, const I* offsets,
           const T* b, std::uint64_t nb, T* x, std::uint64_t nx, double tol, int relative, std::uint64_t iter_max,
           unsigned n_threads, std::uint64_t* iters, double* final_res, int* converged, double* history,
           std::uint64_t history_cap) {
    return guarded([&] {
        if (n_rows != n_cols) throw Panic("Matrix is not symmetric");
        if (n_rows != nb || n_rows != nx) throw Panic("Matrix and vector size mismatch");
        const std::size_t n = n_rows;
        auto spmv = [&](const T* in, T* out) {
            if (n_threads > 1) mvp_crs_threads(n, values, columns, offsets, in, out, n_threads);
            else mvp_crs_raw(n, values, columns, offsets, in, out);
        };
        auto dot = [&](const T* u, const T* w) { T s = T(0); for (std::size_t k = 0; k < n; ++k) s += u[k] * w[k]; return s; };
        double threshold = tol;
        if (relative) threshold = tol * std::sqrt(static_cast<double>(dot(b, b)));
        std::vector<T> r(n), p(n), ap(n), tmp(n);
        spmv(x, ap.data());
        for (std::size_t k = 0; k < n; ++k) r[k] = b[k] - ap[k];
        p = r;
        T rr = dot(r.data(), r.data());
        std::uint64_t it = 0;
        double res = std::sqrt(static_cast<double>(rr));
        int conv = 0;
        for (std::uint64_t k = 0; k < iter_max; ++k) {
            spmv(p.data(), ap.data());
            const T alpha = rr / dot(p.data(), ap.data());
            for (std::size_t q = 0; q < n; ++q) { tmp[q] = p[q] * alpha; }     // p.clone() * alpha
            for (std::size_t q = 0; q < n; ++q) { x[q] += tmp[q]; }            // *x += ...
            for (std::size_t q = 0; q < n; ++q) { tmp[q] = ap[q] * alpha; }    // mat_p * alpha
            for (std::size_t q = 0; q < n; ++q) { r[q] -= tmp[q]; }            // r -= ...
            const T rr_prev = rr;
            rr = dot(r.data(), r.data());
            it = k + 1;
            res = std::sqrt(static_cast<double>(rr));
            if (history && k < history_cap) history[k] = res;
            if (res < threshold) { conv = 1; break; }
            const T beta = rr / rr_prev;
            for (std::size_t q = 0; q < n; ++q) { p[q] *= beta; }              // p.scale(beta)
            for (std::size_t q = 0; q < n; ++q) { p[q] += r[q]; }              // p.add(&r)
        }
        if (iters) *iters = it;
        if (final_res) *final_res = res;
        if (converged) *converged = conv;
    });
}
// What `SparseMatPar<SparseMatIndexList>::mvp` executes today (sparsemat_par.rs:71-140 + the default mvp,
// sparsematrix.rs:146-158): the Laplacian is assembled through SparseMatPar::set (block dispatch, IndexList chains,
// ascending columns per row), then the serial default mvp walks every row through locate() and the block's chain.
// Returns the best-of-`reps` seconds per product (and the assembly time): the CPU baseline of the sparsemat_par path.
template <class T, class I>
int par_laplace_mvp(std::uint64_t n_blocks, std::uint64_t nx, std::uint64_t ny, std::uint64_t nz, const T* x, T* y,
                    unsigned reps, double* sec_per_mvp, double* sec_assemble) {
    return guarded([&] {
        using clock = std::chrono::steady_clock;
        const std::uint64_t n = nx * ny * nz;
        std::vector<T> vals(gen::laplace_nnz(nx, ny, nz));
        std::vector<I> cols(vals.size()), offs(n + 1);
        gen::laplace_rows<T, I>(nx, ny, nz, 0, n, vals.data(), cols.data(), offs.data());
        const auto t0 = clock::now();
        SparseMatPar<SparseMatIndexList<T, I>> par(n_blocks, n);
        for (std::uint64_t r = 0; r < n; ++r)
            for (std::uint64_t k = offs[r]; k < offs[r + 1]; ++k) par.set(r, static_cast<std::size_t>(cols[k]), vals[k]);
        if (sec_assemble) *sec_assemble = std::chrono::duration<double>(clock::now() - t0).count();
        DenseVec<T> xv(std::vector<T>(x, x + n));
        double best = 1e300;
        for (unsigned rep = 0; rep < (reps ? reps : 1u); ++rep) {
            const auto t1 = clock::now();
            DenseVec<T> yv = mvp(par, xv);
            const double dt = std::chrono::duration<double>(clock::now() - t1).count();
            if (dt < best) best = dt;
            if (y && rep == 0) std::memcpy(y, yv.v.data(), yv.v.size() * sizeof(T));
        }
        if (sec_per_mvp) *sec_per_mvp = best;
    });
}
}  // namespace

extern "C" {

const char* orc_last_panic() { return g_panic.c_str(); }
std::uint64_t orc_laplace_nnz(std::uint64_t nx, std::uint64_t ny, std::uint64_t nz) { return gen::laplace_nnz(nx, ny, nz); }
std::uint64_t orc_powerlaw_row_len(std::uint64_t seed, std::uint64_t i, std::uint64_t max_len) {
    return gen::powerlaw_row_len(seed, i, max_len);
}
void orc_powerlaw_row_lens(std::uint64_t seed, std::uint64_t n, std::uint64_t max_len, std::uint64_t* out) {
    for (std::uint64_t i = 0; i < n; ++i) out[i] = gen::powerlaw_row_len(seed, i, max_len);
}
// sparsemat_par.rs:31-35
int orc_par_locate(std::uint64_t n_blocks, std::uint64_t max_rows, std::uint64_t row, std::uint64_t* out2) {
    return guarded([&] {
        if (n_blocks == 0) throw Panic("attempt to divide by zero");
        const std::uint64_t r = max_rows / n_blocks;
        if (r == 0) throw Panic("attempt to divide by zero");
        const std::uint64_t b = std::min<std::uint64_t>(row / r, n_blocks);
        out2[0] = b;
        out2[1] = row - b * r;
    });
}

#define ORC_VEC(T, S)                                                                                              \
    void orc_powerlaw_sample_mvp_##S(std::uint64_t seed_len, std::uint64_t seed_col, std::uint64_t seed_val,        \
                                     std::uint64_t n_cols, std::uint64_t max_len, std::uint64_t n_sample,           \
                                     const std::uint64_t* rows, const T* x, T* y, double* abs_sum,                  \
                                     std::uint64_t* lens) {                                                         \
        gen::powerlaw_sample_mvp<T>(seed_len, seed_col, seed_val, n_cols, max_len, n_sample, rows, x, y, abs_sum, lens); } \
    void orc_uniform_##S(std::uint64_t seed, std::uint64_t n, T* out) { gen::uniform_pm1<T>(seed, n, out); }        \
    T orc_dot_##S(std::uint64_t n, const T* x, const T* y) {                                                        \
        T s = T(0); for (std::uint64_t k = 0; k < n; ++k) s += x[k] * y[k]; return s; }                             \
    T orc_norm2sq_##S(std::uint64_t n, const T* x) {                                                                \
        T s = T(0); for (std::uint64_t k = 0; k < n; ++k) s += x[k] * x[k]; return s; }                             \
    double orc_norm_##S(std::uint64_t n, const T* x) { return std::sqrt(static_cast<double>(orc_norm2sq_##S(n, x))); } \
    int orc_add_##S(std::uint64_t n, T* x, std::uint64_t m, const T* y) {                                           \
        return guarded([&] { if (n < m) throw Panic("Dimension mismatch"); for (std::uint64_t k = 0; k < m; ++k) x[k] += y[k]; }); } \
    int orc_sub_##S(std::uint64_t n, T* x, std::uint64_t m, const T* y) {                                           \
        return guarded([&] { if (n < m) throw Panic("Dimension mismatch"); for (std::uint64_t k = 0; k < m; ++k) x[k] -= y[k]; }); } \
    void orc_scale_##S(std::uint64_t n, T* x, T s) { for (std::uint64_t k = 0; k < n; ++k) x[k] *= s; }

ORC_VEC(float, f32)
ORC_VEC(double, f64)

#define ORC_MAT(T, I, S)                                                                                           \
    void* orc_il_new_##S() { return new SparseMatIndexList<T, I>(); }                                               \
    void orc_il_free_##S(void* h) { delete static_cast<SparseMatIndexList<T, I>*>(h); }                             \
    int orc_il_apply_##S(void* h, std::uint64_t n, const std::uint64_t* i, const std::uint64_t* j, const T* v,      \
                         const std::uint8_t* op) {                                                                  \
        return il_apply(static_cast<SparseMatIndexList<T, I>*>(h), n, i, j, v, op); }                               \
    T orc_il_get_##S(void* h, std::uint64_t i, std::uint64_t j) {                                                   \
        return static_cast<SparseMatIndexList<T, I>*>(h)->get(i, j); }                                              \
    void orc_il_dims_##S(void* h, std::uint64_t* out3) {                                                            \
        auto* m = static_cast<SparseMatIndexList<T, I>*>(h);                                                        \
        out3[0] = m->n_rows(); out3[1] = m->n_cols(); out3[2] = m->nnz(); }                                         \
    void orc_il_export_##S(void* h, I* columns, T* values, I* pos_start, I* next) {                                 \
        il_export(static_cast<SparseMatIndexList<T, I>*>(h), columns, values, pos_start, next); }                   \
    void* orc_il_transpose_##S(void* h) {                                                                           \
        return new SparseMatIndexList<T, I>(transpose(*static_cast<SparseMatIndexList<T, I>*>(h))); }               \
    void* orc_il_to_crs_##S(void* h) {                                                                              \
        return new SparseMatCRS<T, I>(SparseMatCRS<T, I>::from_indexlist(*static_cast<SparseMatIndexList<T, I>*>(h))); } \
    void orc_crs_free_##S(void* h) { delete static_cast<SparseMatCRS<T, I>*>(h); }                                  \
    void orc_crs_dims_##S(void* h, std::uint64_t* out4) {                                                           \
        auto* m = static_cast<SparseMatCRS<T, I>*>(h);                                                              \
        out4[0] = m->n_rows(); out4[1] = m->n_cols(); out4[2] = m->nnz(); out4[3] = m->offset_rows.size(); }        \
    void orc_crs_export_##S(void* h, T* values, I* columns, I* offsets) {                                           \
        crs_export(static_cast<SparseMatCRS<T, I>*>(h), values, columns, offsets); }                                \
    int orc_to_crs_raw_##S(std::uint64_t n_rows, std::uint64_t nnz, const I* columns, const T* values,              \
                           const I* pos_start, const I* next, T* ov, I* oc, I* oo) {                                \
        return to_crs_raw<T, I>(n_rows, nnz, columns, values, pos_start, next, ov, oc, oo); }                       \
    void orc_mvp_##S(std::uint64_t n_rows, const T* values, const I* columns, const I* offsets, const T* x, T* y) { \
        mvp_crs_raw<T, I>(n_rows, values, columns, offsets, x, y); }                                                \
    void orc_mvp_threads_##S(std::uint64_t n_rows, const T* values, const I* columns, const I* offsets, const T* x, \
                             T* y, unsigned n_threads) {                                                            \
        mvp_crs_threads<T, I>(n_rows, values, columns, offsets, x, y, n_threads); }                                 \
    int orc_mvp_checked_##S(std::uint64_t n_rows, std::uint64_t n_cols, std::uint64_t nnz, const T* values,         \
                            const I* columns, const I* offsets, const T* x, std::uint64_t nx, T* y) {               \
        return mvp_checked<T, I>(n_rows, n_cols, nnz, values, columns, offsets, x, nx, y); }                        \
    int orc_bilinear_##S(std::uint64_t n_rows, std::uint64_t n_cols, std::uint64_t nnz, const T* values,            \
                         const I* columns, const I* offsets, const T* lhs, std::uint64_t nl, const T* rhs,          \
                         std::uint64_t nr, T* out) {                                                                \
        return bilinear_raw<T, I>(n_rows, n_cols, nnz, values, columns, offsets, lhs, nl, rhs, nr, out); }          \
    int orc_cg_##S(std::uint64_t n_rows, std::uint64_t n_cols, const T* values, const I* columns, const I* offsets, \
                   const T* b, std::uint64_t nb, T* x, std::uint64_t nx, double tol, int relative,                  \
                   std::uint64_t iter_max, unsigned n_threads, std::uint64_t* iters, double* final_res,             \
                   int* converged, double* history, std::uint64_t history_cap) {                                    \
        return cg_raw<T, I>(n_rows, n_cols, values, columns, offsets, b, nb, x, nx, tol, relative, iter_max,        \
                            n_threads, iters, final_res, converged, history, history_cap); }                        \
    int orc_par_laplace_mvp_##S(std::uint64_t n_blocks, std::uint64_t nx, std::uint64_t ny, std::uint64_t nz, const T* x, \
                                T* y, unsigned reps, double* sec_per_mvp, double* sec_assemble) {                    \
        return par_laplace_mvp<T, I>(n_blocks, nx, ny, nz, x, y, reps, sec_per_mvp, sec_assemble); }                \
    void orc_laplace_rows_##S(std::uint64_t nx, std::uint64_t ny, std::uint64_t nz, std::uint64_t row_lo,           \
                              std::uint64_t row_hi, T* values, I* columns, I* offsets) {                            \
        gen::laplace_rows<T, I>(nx, ny, nz, row_lo, row_hi, values, columns, offsets); }                            \
    void orc_powerlaw_fill_##S(std::uint64_t seed_col, std::uint64_t seed_val, std::uint64_t n_rows,                \
                               std::uint64_t n_cols, const I* offsets, T* values, I* columns) {                     \
        gen::powerlaw_fill<T, I>(seed_col, seed_val, n_rows, n_cols, offsets, values, columns); }

ORC_MAT(float, std::uint32_t, f32u32)
ORC_MAT(double, std::uint32_t, f64u32)
ORC_MAT(float, std::uint64_t, f32u64)
ORC_MAT(double, std::uint64_t, f64u64)

}  // extern "C"
