// C++ host-side mirror of the reference's public interface for the hot path, on top of the C ABI
// (include/smb200.h).  The reference is a Rust crate and no Rust toolchain exists in this image, so
// this header plays the role of the patched crate: same type and method names, same argument
// meaning, same failure behaviour (the reference's panics surface as sparsemat::Panic carrying the
// reference's message).  rust/ holds the equivalent Rust overlay as source.
//
//   reference (src/…)                          here
//   SparseMatIndexList<T,I>  sparsemat_indexlist.rs   sparsemat::SparseMatIndexList<T,I>   (host assembly)
//   SparseMatIndexList::to_crs :61-63                 .to_crs(ctx)  -> device conversion (K8)
//   SparseMatCRS<T,I>        sparsemat_crs.rs          sparsemat::SparseMatCRS<T,I>         (device resident)
//   SparseMatrix::mvp        sparsematrix.rs:146-158   .mvp(x) / operator*                  (CUDA SpMV)
//   DenseVec<T>              densevec.rs               sparsemat::DenseVec<T>               (device resident)
//   Vector::{inner_prod,norm_squared,norm} vector.rs:50-63
//   ConjugateGradient        linearsolver.rs:12-61     sparsemat::ConjugateGradient
//   SparseMatPar             sparsemat_par.rs          sparsemat::par_locate (partition contract)
//
// Assembly (set / add_to / get_mut) stays on the host exactly like the reference; everything that the
// solver spends time in runs on the GPU.  Nothing here falls back to the CPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/smb200.h"

namespace sparsemat {

struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };

inline void check(smb200_status s) {
    if (s == SMB200_OK) return;
    const char* msg = smb200_last_error();
    switch (s) {
        case SMB200_ERR_DIM: throw Panic("Dimension mismatch");
        case SMB200_ERR_NOT_SQUARE: throw Panic("Matrix is not symmetric");
        case SMB200_ERR_SIZE_MISMATCH: throw Panic("Matrix and vector size mismatch");
        default: throw std::runtime_error(std::string("smb200: ") + (msg ? msg : "unknown error"));
    }
}

template <class T> constexpr smb200_vtype vtype_of() {
    static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value, "FloatType: f32 or f64");
    return std::is_same<T, double>::value ? SMB200_F64 : SMB200_F32;
}
template <class I> constexpr smb200_itype itype_of() {
    static_assert(std::is_same<I, uint32_t>::value || std::is_same<I, uint64_t>::value, "IndexType on the GPU path: u32 or u64");
    return std::is_same<I, uint64_t>::value ? SMB200_U64 : SMB200_U32;
}

// One CUDA context (device + stream) shared by the objects created from it.
class Context {
public:
    explicit Context(int device = 0, void* stream = nullptr) {
        smb200_ctx* c = nullptr;
        check(smb200_ctx_create(device, stream, &c));
        h_.reset(c, [](smb200_ctx* p) { smb200_ctx_destroy(p); });
    }
    smb200_ctx* get() const { return h_.get(); }
    void sync() const { check(smb200_ctx_sync(h_.get())); }
private:
    std::shared_ptr<smb200_ctx> h_;
};

// ---- DenseVec<T> ------------------------------------------------------------------------------------------
template <class T>
class DenseVec {
public:
    DenseVec(const Context& ctx, uint64_t n) : ctx_(ctx) {                       // with_capacity + zeros
        smb200_vec* v = nullptr;
        check(smb200_vec_create(ctx.get(), vtype_of<T>(), n, &v));
        h_.reset(v, [](smb200_vec* p) { smb200_vec_free(p); });
    }
    static DenseVec from_vec(const Context& ctx, const std::vector<T>& host) {   // densevec.rs:30-34
        DenseVec r(ctx, host.size());
        check(smb200_vec_upload(r.raw(), host.data(), host.size()));
        return r;
    }
    DenseVec clone() const {                                                      // #[derive(Clone)]
        smb200_vec* v = nullptr;
        check(smb200_vec_clone(raw(), &v));
        return DenseVec(ctx_, v);
    }
    uint64_t dim() const { uint64_t n = 0; check(smb200_vec_dim(raw(), &n)); return n; }   // :36-38
    std::vector<T> to_vec() const {
        std::vector<T> out(dim());
        check(smb200_vec_download(raw(), out.data(), out.size()));
        return out;
    }
    T get(uint64_t i) const {                                                     // :40-42 (bounds-checked)
        if (i >= dim()) throw Panic("index out of bounds");
        return to_vec()[i];
    }
    void add(const DenseVec& rhs) { check(smb200_vec_add(raw(), rhs.raw())); }    // :51-58
    void sub(const DenseVec& rhs) { check(smb200_vec_sub(raw(), rhs.raw())); }    // :60-67
    void scale(T s) { check(smb200_vec_scale(raw(), (double)s)); }                // :69-73
    T inner_prod(const DenseVec& rhs) const { double d; check(smb200_vec_dot(raw(), rhs.raw(), &d)); return (T)d; }   // vector.rs:50-53
    T norm_squared() const { double d; check(smb200_vec_norm2sq(raw(), &d)); return (T)d; }                           // vector.rs:56-58
    double norm() const { double d; check(smb200_vec_norm(raw(), &d)); return d; }                                    // vector.rs:61-63
    // operators of densevec.rs:76-140 (the binary forms clone the left operand, like the reference)
    DenseVec& operator+=(const DenseVec& r) { add(r); return *this; }
    DenseVec& operator-=(const DenseVec& r) { sub(r); return *this; }
    DenseVec& operator*=(T s) { scale(s); return *this; }
    DenseVec operator+(const DenseVec& r) const { DenseVec t = clone(); t.add(r); return t; }
    DenseVec operator-(const DenseVec& r) const { DenseVec t = clone(); t.sub(r); return t; }
    DenseVec operator*(T s) const { DenseVec t = clone(); t.scale(s); return t; }
    T operator*(const DenseVec& r) const { return inner_prod(r); }
    smb200_vec* raw() const { return h_.get(); }
    const Context& context() const { return ctx_; }
private:
    DenseVec(const Context& ctx, smb200_vec* v) : ctx_(ctx) { h_.reset(v, [](smb200_vec* p) { smb200_vec_free(p); }); }
    Context ctx_;
    std::shared_ptr<smb200_vec> h_;
};

// ---- SparseMatCRS<T,I>: device resident ---------------------------------------------------------------------
template <class T, class I>
class SparseMatCRS {
public:
    // Equivalent of the overlay's `SparseMatCRS::raw_parts()` hand-over: finished arrays in.
    static SparseMatCRS from_raw_parts(const Context& ctx, uint64_t n_rows, uint64_t n_cols, const std::vector<T>& values,
                                       const std::vector<I>& columns, const std::vector<I>& offset_rows) {
        // the reference derives these lengths from its Vecs and cannot get them wrong; raw arrays can
        if (columns.size() != values.size()) throw Panic("SparseMatCRS::from_raw_parts: columns and values differ in length");
        if (offset_rows.size() != (n_rows || !values.empty() ? n_rows + 1 : offset_rows.size()))
            throw Panic("SparseMatCRS::from_raw_parts: offset_rows must hold n_rows + 1 entries");
        smb200_crs* m = nullptr;
        check(smb200_crs_upload(ctx.get(), vtype_of<T>(), itype_of<I>(), n_rows, n_cols, values.size(), values.data(),
                                columns.data(), offset_rows.data(), &m));
        return SparseMatCRS(ctx, m);
    }
    // sparsematrix.rs:174-183 — row j of the result: column j's entries ordered by source row
    SparseMatCRS transpose() const {
        smb200_crs* m = nullptr;
        check(smb200_crs_transpose(raw(), &m));
        return SparseMatCRS(ctx_, m);
    }
    // Binary CRS container (additive: the reference has text/PBM writers only, sparsematrix.rs:304-338): the three arrays
    // byte for byte behind a checksummed header.  load() refuses files of another value/index type.
    void save(const std::string& path) const { check(smb200_crs_save(raw(), path.c_str())); }
    static SparseMatCRS load(const Context& ctx, const std::string& path) {
        int32_t vt = 0, it = 0;
        check(smb200_crsfile_info(path.c_str(), &vt, &it, nullptr));
        if (vt != (int32_t)vtype_of<T>() || it != (int32_t)itype_of<I>()) throw Panic("SparseMatCRS::load: the file holds another value/index type");
        smb200_crs* m = nullptr;
        check(smb200_crs_load(ctx.get(), path.c_str(), &m));
        return SparseMatCRS(ctx, m);
    }
    uint64_t n_rows() const { return dims()[0]; }
    uint64_t n_cols() const { return dims()[1]; }
    uint64_t n_non_zero_entries() const { return dims()[2]; }
    bool empty() const { return n_rows() == 0; }
    double density() const { return (double)n_non_zero_entries() / (double)(n_rows() * n_cols()); }   // sparsematrix.rs:237-241
    // sparsematrix.rs:146-158 — returns a fresh vector of dim n_rows
    DenseVec<T> mvp(const DenseVec<T>& rhs) const {
        DenseVec<T> y(ctx_, n_rows());
        check(smb200_spmv(raw(), rhs.raw(), y.raw()));
        return y;
    }
    DenseVec<T> operator*(const DenseVec<T>& rhs) const { return mvp(rhs); }        // sparsematrix.rs:435-443
    T inner_prod(const DenseVec<T>& lhs, const DenseVec<T>& rhs) const {            // sparsematrix.rs:161-171
        double d; check(smb200_bilinear(raw(), lhs.raw(), rhs.raw(), &d)); return (T)d;
    }
    void scale(T s) { check(smb200_crs_scale(raw(), (double)s)); }                  // sparsemat_crs.rs:153-157
    struct RawParts { std::vector<T> values; std::vector<I> columns; std::vector<I> offset_rows; };
    RawParts raw_parts() const {
        RawParts p;
        const uint64_t nr = n_rows(), nz = n_non_zero_entries();
        p.values.resize(nz); p.columns.resize(nz); p.offset_rows.resize(nz || nr ? nr + 1 : 0);
        check(smb200_crs_download(raw(), p.values.data(), p.columns.data(), p.offset_rows.empty() ? nullptr : p.offset_rows.data()));
        return p;
    }
    // sparsemat_crs.rs:102-110 — row slice as (column, value) pairs; rows past the end are empty
    std::vector<std::pair<I, T>> iter_row(uint64_t row) const {
        std::vector<std::pair<I, T>> out;
        if (row >= n_rows()) return out;
        RawParts p = raw_parts();
        for (uint64_t k = p.offset_rows[row]; k < p.offset_rows[row + 1]; ++k) out.emplace_back(p.columns[k], p.values[k]);
        return out;
    }
    smb200_crs* raw() const { return h_.get(); }
    const Context& context() const { return ctx_; }
    SparseMatCRS(const Context& ctx, smb200_crs* m) : ctx_(ctx) { h_.reset(m, [](smb200_crs* p) { smb200_crs_free(p); }); }
private:
    std::vector<uint64_t> dims() const { std::vector<uint64_t> d(3); check(smb200_crs_dims(raw(), d.data())); return d; }
    Context ctx_;
    std::shared_ptr<smb200_crs> h_;
};

// ---- IndexList / SparseMatIndexList: host-side assembly -----------------------------------------------------
// Same arrays as indexlist.rs:26-29 (pos_start, index_list) with I::MAX as UNSET.  A per-row tail
// pointer makes push O(1) instead of the reference's chain walk (indexlist.rs:74-80); the arrays it
// produces are identical.
template <class I>
struct IndexList {
    static constexpr I UNSET = std::numeric_limits<I>::max();
    std::vector<I> pos_start, index_list, tail;
    uint64_t n_rows() const { return pos_start.size(); }
    uint64_t n_entries() const { return index_list.size(); }
    uint64_t push(uint64_t row) {
        if (row >= pos_start.size()) { pos_start.resize(row + 1, UNSET); tail.resize(row + 1, UNSET); }
        const I id = (I)index_list.size();
        if (id == UNSET) throw Panic("assertion failed: index != UNSET");
        index_list.push_back(UNSET);
        if (pos_start[row] == UNSET) pos_start[row] = id; else index_list[tail[row]] = id;
        tail[row] = id;
        return (uint64_t)id;
    }
};

template <class T, class I>
class SparseMatIndexList {
public:
    static constexpr I UNSET = std::numeric_limits<I>::max();
    uint64_t n_rows() const { return list_.n_rows(); }
    uint64_t n_cols() const { return n_cols_; }
    uint64_t n_non_zero_entries() const { return columns_.size(); }
    T get(uint64_t i, uint64_t j) const { const uint64_t e = find(i, j); return e == npos ? T(0) : values_[e]; }
    T& get_mut(uint64_t i, uint64_t j) {                                            // sparsemat_indexlist.rs:158-164
        uint64_t e = find(i, j);
        if (e == npos) {
            if (j >= n_cols_) n_cols_ = j + 1;
            e = list_.push(i);
            columns_.push_back((I)j);
            values_.push_back(T(0));
        }
        return values_[e];
    }
    void set(uint64_t i, uint64_t j, T v) { get_mut(i, j) = v; }                    // sparsematrix.rs:226-228
    void add_to(uint64_t i, uint64_t j, T v) { get_mut(i, j) += v; }                // sparsematrix.rs:231-233
    std::vector<std::pair<I, T>> iter_row(uint64_t row) const {                     // chain order
        if (row >= n_rows()) throw Panic("index out of bounds");
        std::vector<std::pair<I, T>> out;
        for (I p = list_.pos_start[row]; p != UNSET; p = list_.index_list[p]) out.emplace_back(columns_[p], values_[p]);
        return out;
    }
    // sparsemat_indexlist.rs:61-63 — conversion runs on the device (smb200_crs_from_indexlist)
    SparseMatCRS<T, I> to_crs(const Context& ctx) const {
        smb200_crs* m = nullptr;
        check(smb200_crs_from_indexlist(ctx.get(), vtype_of<T>(), itype_of<I>(), n_rows(), n_cols_, columns_.size(),
                                        columns_.data(), values_.data(), list_.pos_start.data(), list_.index_list.data(), &m));
        return SparseMatCRS<T, I>(ctx, m);
    }
    const std::vector<I>& columns() const { return columns_; }
    const std::vector<T>& values() const { return values_; }
    const IndexList<I>& chains() const { return list_; }
private:
    static constexpr uint64_t npos = ~0ull;
    uint64_t find(uint64_t i, uint64_t j) const {                                   // sparsemat_indexlist.rs:29-42
        if (i >= n_rows()) return npos;
        const I col = (I)j;
        for (I p = list_.pos_start[i]; p != UNSET; p = list_.index_list[p])
            if (columns_[p] == col) return (uint64_t)p;
        return npos;
    }
    uint64_t n_cols_ = 0;
    std::vector<I> columns_;
    std::vector<T> values_;
    IndexList<I> list_;
};

// ---- ConjugateGradient (linearsolver.rs:12-61) ----------------------------------------------------------------
struct CgStats { uint64_t iterations = 0; double final_residual = 0.0; bool converged = false; float device_ms = 0.f; };

class ConjugateGradient {
public:
    ConjugateGradient() = default;                                                   // Default: 1e-12, 10_000 (:17-24)
    ConjugateGradient(double tol, uint64_t iter_max, bool relative = false)         // additive constructor
        : tol_(tol), iter_max_(iter_max), relative_(relative) {}
    template <class T, class I>
    void solve(const SparseMatCRS<T, I>& mat, const DenseVec<T>& b, DenseVec<T>& x) const { solve_with_stats(mat, b, x); }
    template <class T, class I>
    CgStats solve_with_stats(const SparseMatCRS<T, I>& mat, const DenseVec<T>& b, DenseVec<T>& x) const {
        smb200_cg_stats st;
        check(smb200_cg_solve(mat.raw(), b.raw(), x.raw(), tol_, relative_ ? 1 : 0, iter_max_, &st));
        CgStats out;
        out.iterations = st.iterations; out.final_residual = st.final_residual; out.converged = st.converged != 0;
        out.device_ms = st.device_ms;
        return out;
    }
private:
    double tol_ = 1e-12;
    uint64_t iter_max_ = 10000;
    bool relative_ = false;
};

// Additive (the reference has no preconditioner): CG preconditioned with the inverse diagonal.  Same constructor
// arguments, checks and panics as ConjugateGradient.
class JacobiPCG {
public:
    JacobiPCG() = default;
    JacobiPCG(double tol, uint64_t iter_max, bool relative = false) : tol_(tol), iter_max_(iter_max), relative_(relative) {}
    template <class T, class I>
    CgStats solve_with_stats(const SparseMatCRS<T, I>& mat, const DenseVec<T>& b, DenseVec<T>& x) const {
        smb200_cg_stats st;
        check(smb200_pcg_jacobi_solve(mat.raw(), b.raw(), x.raw(), tol_, relative_ ? 1 : 0, iter_max_, &st));
        CgStats out;
        out.iterations = st.iterations; out.final_residual = st.final_residual; out.converged = st.converged != 0;
        out.device_ms = st.device_ms;
        return out;
    }
    template <class T, class I>
    void solve(const SparseMatCRS<T, I>& mat, const DenseVec<T>& b, DenseVec<T>& x) const { solve_with_stats(mat, b, x); }
private:
    double tol_ = 1e-12;
    uint64_t iter_max_ = 10000;
    bool relative_ = false;
};

// sparsemat_par.rs:31-35 — the row-block contract the multi-GPU partitioner follows.
inline std::pair<uint64_t, uint64_t> par_locate(uint64_t n_blocks, uint64_t max_n_rows, uint64_t row) {
    uint64_t b = 0, r = 0;
    if (smb200_par_locate(n_blocks, max_n_rows, row, &b, &r) != SMB200_OK) throw Panic("attempt to divide by zero");
    return {b, r};
}

}  // namespace sparsemat
