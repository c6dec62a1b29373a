// ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code, never shipped, never on the GPU path.
//
// CPU restatement (C++17) of the part of lostinc0de/sparsemat that lies on the CRS SpMV / CG hot
// path, used as the parity checker for the CUDA library (libsmb200) and as the timed CPU baseline.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Parity status: PINNED.  oracle/kat.cpp replays the reference's own unit tests (src/lib.rs:36-218)
// against this restatement and requires the same exact floating-point answers (34.544f, 20.16f,
// 17.9632f, 0.0909, "0 2.24 4.12 ", iteration orders, densities, chain order).
//
// Arithmetic rules that make the restatement faithful (the reference is safe Rust; rustc neither
// contracts a*b+c into an FMA nor re-associates float sums):
//   * build with  -O2 -ffp-contract=off  and never -ffast-math (see oracle/Makefile);
//   * every sum is a sequential left-to-right fold starting from +0.0 (Rust `Iterator::sum`,
//     `sum += ...` loops);
//   * axpy-like updates are scale-then-add: two roundings, as `p.clone() * alpha` then `+=`.
//
// Citations `file:line` refer to /root/reference/src/.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace oracle {

// A Rust `panic!` on the reference side surfaces here as this exception (message kept verbatim).
struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };

// types.rs:14-51 — index casts are plain `as` casts: silent truncation on narrowing.
template <class I> inline I as_index(std::size_t v) { return static_cast<I>(v); }
template <class I> constexpr I unset() { return std::numeric_limits<I>::max(); }  // sparsematrix.rs:68

// ------------------------------------------------------------------------------------------------
// indexlist.rs:26-112 — per-row singly linked chains threaded through one flat `next` array.
template <class I>
struct IndexList {
    std::vector<I> pos_start;   // first entry of each row, or UNSET
    std::vector<I> next;        // `index_list` in the reference: successor in the same row, or UNSET

    std::size_t n_entries() const { return next.size(); }        // :52-54
    std::size_t n_rows() const { return pos_start.size(); }      // :57-59

    // :62-83 — append an entry for `row`; O(row length) because the chain is walked to its tail.
    std::size_t push(std::size_t row) {
        if (row >= pos_start.size()) pos_start.resize(row + 1, unset<I>());
        const I fresh = as_index<I>(n_entries());
        if (fresh == unset<I>()) throw Panic("assertion failed: index != UNSET");   // :68
        next.push_back(unset<I>());
        if (pos_start[row] == unset<I>()) {
            pos_start[row] = fresh;
        } else {
            std::size_t at = static_cast<std::size_t>(pos_start[row]);
            while (next[at] != unset<I>()) at = static_cast<std::size_t>(next[at]);
            next[at] = fresh;
        }
        return static_cast<std::size_t>(fresh);
    }

    // :85-111 — visit entry ids of `row` in chain order.  Like the reference, row >= n_rows is an
    // out-of-bounds panic (vector index), not an empty iteration.
    template <class F> void for_row(std::size_t row, F&& f) const {
        if (row >= pos_start.size()) throw Panic("index out of bounds: IndexList::iter_row");
        for (I p = pos_start[row]; p != unset<I>(); p = next[static_cast<std::size_t>(p)])
            if (!f(static_cast<std::size_t>(p))) break;
    }
};

// ------------------------------------------------------------------------------------------------
// densevec.rs:5-140 + vector.rs:5-64
template <class T>
struct DenseVec {
    std::vector<T> v;

    DenseVec() = default;
    explicit DenseVec(std::vector<T> init) : v(std::move(init)) {}          // from_vec :30-34
    std::size_t dim() const { return v.size(); }                            // :36-38
    T get(std::size_t i) const {                                            // :40-42 (bounds-checked)
        if (i >= v.size()) throw Panic("index out of bounds: DenseVec::get");
        return v[i];
    }
    T& get_mut(std::size_t i) {                                             // :44-49 (grows with zeros)
        if (i >= v.size()) v.resize(i + 1, T(0));
        return v[i];
    }
    void set(std::size_t i, T val) { get_mut(i) = val; }                    // vector.rs:40-42
    void add_to(std::size_t i, T val) { get_mut(i) += val; }                // vector.rs:45-47

    void add(const DenseVec& rhs) {                                         // :51-58
        if (dim() < rhs.dim()) throw Panic("Dimension mismatch");
        for (std::size_t k = 0; k < rhs.v.size(); ++k) v[k] += rhs.v[k];
    }
    void sub(const DenseVec& rhs) {                                         // :60-67
        if (dim() < rhs.dim()) throw Panic("Dimension mismatch");
        for (std::size_t k = 0; k < rhs.v.size(); ++k) v[k] -= rhs.v[k];
    }
    void scale(T s) { for (T& e : v) e *= s; }                              // :69-73

    // vector.rs:50-53 — zip stops at the shorter operand; sequential fold from 0.
    T inner_prod(const DenseVec& rhs) const {
        const std::size_t n = std::min(v.size(), rhs.v.size());
        T acc = T(0);
        for (std::size_t k = 0; k < n; ++k) acc += v[k] * rhs.v[k];
        return acc;
    }
    T norm_squared() const {                                                // vector.rs:56-58
        T acc = T(0);
        for (const T& e : v) acc += e * e;
        return acc;
    }
    double norm() const { return std::sqrt(static_cast<double>(norm_squared())); }  // vector.rs:61-63

    // operators, densevec.rs:76-140: the binary forms clone the left operand first.
    DenseVec plus(const DenseVec& rhs) const { DenseVec r = *this; r.add(rhs); return r; }   // :98-107
    DenseVec minus(const DenseVec& rhs) const { DenseVec r = *this; r.sub(rhs); return r; }  // :109-118
    DenseVec times(T s) const { DenseVec r = *this; r.scale(s); return r; }                  // :121-130
};

// ------------------------------------------------------------------------------------------------
// Shared default methods of trait SparseMatrix (sparsematrix.rs:62-339).  `M` provides
//   n_rows(), n_cols(), nnz(), for_row(row, f(col, val) -> bool), get(i,j), get_mut(i,j).

// sparsematrix.rs:146-158 — THE hot loop: per row, `sum += x[col] * val` in storage order.
template <class M, class T>
DenseVec<T> mvp(const M& a, const DenseVec<T>& x) {
    DenseVec<T> y;
    y.v.reserve(a.n_rows());
    const std::size_t n = a.n_rows();
    for (std::size_t i = 0; i < n; ++i) {
        T sum = T(0);
        a.for_row(i, [&](auto col, T val) {
            sum += x.get(static_cast<std::size_t>(col)) * val;
            return true;
        });
        y.set(i, sum);
    }
    return y;
}

// sparsematrix.rs:161-171 — bilinear form, one running sum over the whole matrix,
// association (lhs[i] * val) * rhs[j].
template <class M, class T>
T bilinear(const M& a, const DenseVec<T>& lhs, const DenseVec<T>& rhs) {
    T sum = T(0);
    for (std::size_t i = 0; i < a.n_rows(); ++i)
        a.for_row(i, [&](auto col, T val) {
            sum += lhs.get(i) * val * rhs.get(static_cast<std::size_t>(col));
            return true;
        });
    return sum;
}

template <class M> double density(const M& a) {                             // :237-241
    return static_cast<double>(a.nnz()) / static_cast<double>(a.n_rows() * a.n_cols());
}

// sparsematrix.rs:124-143 — add/sub another matrix entry by entry through get_mut.
template <class M, class S> void mat_add(M& self, const S& rhs) {
    for (std::size_t i = 0; i < rhs.n_rows(); ++i)
        rhs.for_row(i, [&](auto col, auto val) { self.get_mut(i, static_cast<std::size_t>(col)) += val; return true; });
}
template <class M, class S> void mat_sub(M& self, const S& rhs) {
    for (std::size_t i = 0; i < rhs.n_rows(); ++i)
        rhs.for_row(i, [&](auto col, auto val) { self.get_mut(i, static_cast<std::size_t>(col)) -= val; return true; });
}

// sparsematrix.rs:28-59 + :75-81 — whole-matrix iteration order (row-major, storage order in a row).
template <class M, class T>
std::vector<std::tuple<std::size_t, std::size_t, T>> iter_all(const M& a) {
    std::vector<std::tuple<std::size_t, std::size_t, T>> out;
    for (std::size_t i = 0; i < a.n_rows(); ++i)
        a.for_row(i, [&](auto col, T val) { out.emplace_back(i, static_cast<std::size_t>(col), val); return true; });
    return out;
}

// sparsematrix.rs:272-301 — row rendered with explicit zeros after sorting by column.
// Rust's `{}` for floats prints the shortest round-trip decimal; callers pass a formatter.
template <class M, class Fmt>
std::string to_string_row(const M& a, std::size_t i, Fmt&& fmt) {
    using Col = typename M::index_type;
    using Val = typename M::value_type;
    std::vector<std::pair<Col, Val>> row;
    a.for_row(i, [&](Col c, Val v) { row.emplace_back(c, v); return true; });
    std::stable_sort(row.begin(), row.end(), [](auto& l, auto& r) { return l.first < r.first; });
    std::string s;
    Col j = 0;
    for (auto& [c, v] : row) {
        while (j < c) { s += "0 "; ++j; }
        s += fmt(v);
        s += " ";
        ++j;
    }
    return s;
}

// ------------------------------------------------------------------------------------------------
// sparsemat_indexlist.rs:14-207 — the assembly format.
template <class T, class I>
struct SparseMatIndexList {
    using value_type = T;
    using index_type = I;
    std::size_t ncols = 0;
    std::vector<I> columns;
    std::vector<T> values;
    IndexList<I> chains;
    std::vector<I> rows;        // column-iterator side info, copied verbatim by to_crs
    IndexList<I> chains_col;

    std::size_t n_rows() const { return chains.n_rows(); }                  // :137-139
    std::size_t n_cols() const { return ncols; }
    std::size_t nnz() const { return columns.size(); }
    bool empty() const { return n_rows() == 0; }                            // sparsematrix.rs:119-121

    template <class F> void for_row(std::size_t row, F&& f) const {         // :119-124,173-188
        chains.for_row(row, [&](std::size_t e) { return f(columns[e], values[e]); });
    }

    static constexpr std::size_t npos() { return static_cast<std::size_t>(unset<I>()); }

    std::size_t find_index(std::size_t i, std::size_t j) const {            // :29-42
        const I col = as_index<I>(j);
        std::size_t hit = npos();
        if (i < n_rows())
            chains.for_row(i, [&](std::size_t e) {
                if (columns[e] == col) { hit = e; return false; }
                return true;
            });
        return hit;
    }
    std::size_t push(std::size_t i, std::size_t j, T val) {                 // :45-53
        if (j >= ncols) ncols = j + 1;
        const std::size_t e = chains.push(i);
        columns.push_back(as_index<I>(j));
        values.push_back(val);
        return e;
    }
    T get(std::size_t i, std::size_t j) const {                             // :149-156
        const std::size_t e = find_index(i, j);
        return e == npos() ? T(0) : values[e];
    }
    T& get_mut(std::size_t i, std::size_t j) {                              // :158-164
        std::size_t e = find_index(i, j);
        if (e == npos()) e = push(i, j, T(0));
        return values[e];
    }
    void set(std::size_t i, std::size_t j, T val) { get_mut(i, j) = val; }  // sparsematrix.rs:226-228
    void add_to(std::size_t i, std::size_t j, T val) { get_mut(i, j) += val; }  // :231-233
    void scale(T s) { for (T& e : values) e *= s; }                         // :166-170

    void assemble_column_info() {                                           // :71-84
        rows.resize(columns.size(), unset<I>());
        for (std::size_t i = 0; i < n_rows(); ++i)
            chains.for_row(i, [&](std::size_t e) { rows[e] = as_index<I>(i); return true; });
        for (const I& c : columns) chains_col.push(static_cast<std::size_t>(c));
    }
    // :86-97 — visit (row, value) of a column in column-chain order.
    template <class F> void for_col(std::size_t col, F&& f) const {
        if (rows.size() != columns.size())
            throw Panic("Column iterator not available - use assemble_column_info()");
        chains_col.for_row(col, [&](std::size_t e) { return f(rows[e], values[e]); });
    }
    void sort_row(std::size_t i) {                                          // :102-109
        std::vector<std::pair<I, T>> cv;
        for_row(i, [&](I c, T v) { cv.emplace_back(c, v); return true; });
        std::stable_sort(cv.begin(), cv.end(), [](auto& l, auto& r) { return l.first < r.first; });
        std::size_t k = 0;
        chains.for_row(i, [&](std::size_t e) { columns[e] = cv[k].first; values[e] = cv[k].second; ++k; return true; });
    }
};

// sparsematrix.rs:174-183 — `transpose`: ret = with_capacity(nnz); rows ascending, a row's entries in storage order,
// ret.set(col, i, val).  Restated on the assembly format (the CRS `push` path of the reference is O(nnz) per insert and
// loses rows through its first-insert quirk, sparsemat_crs.rs:71-92, so nobody transposes through it).
template <class T, class I>
SparseMatIndexList<T, I> transpose(const SparseMatIndexList<T, I>& a) {
    SparseMatIndexList<T, I> ret;
    for (std::size_t i = 0; i < a.n_rows(); ++i)
        a.for_row(i, [&](I c, T v) { ret.set(static_cast<std::size_t>(c), i, v); return true; });
    return ret;
}

// ------------------------------------------------------------------------------------------------
// sparsemat_crs.rs:9-223 — the compute format.
template <class T, class I>
struct SparseMatCRS {
    using value_type = T;
    using index_type = I;
    std::size_t nrows = 0, ncols = 0;
    std::vector<T> values;
    std::vector<I> columns;
    std::vector<I> offset_rows;       // stored in the index type, not usize (:14)
    std::vector<I> rows;
    IndexList<I> chains_col;

    std::size_t n_rows() const { return nrows; }
    std::size_t n_cols() const { return ncols; }
    std::size_t nnz() const { return columns.size(); }
    bool empty() const { return nrows == 0; }

    // :24-50 — IndexList -> CRS.  Rows ascending; inside a row the chain (= insertion) order is
    // kept, nothing is sorted; empty rows repeat the running offset; an empty source gives 0x0.
    static SparseMatCRS from_indexlist(const SparseMatIndexList<T, I>& src) {
        SparseMatCRS out;
        if (src.nnz() == 0) return out;
        out.values.reserve(src.nnz());
        out.columns.reserve(src.nnz());
        out.offset_rows.reserve(src.n_rows() + 1);
        for (std::size_t i = 0; i < src.n_rows(); ++i) {
            out.offset_rows.push_back(as_index<I>(out.columns.size()));
            src.for_row(i, [&](I c, T v) { out.columns.push_back(c); out.values.push_back(v); return true; });
        }
        out.offset_rows.push_back(as_index<I>(out.columns.size()));
        out.rows = src.rows;                 // column_info(): verbatim clones (:37)
        out.chains_col = src.chains_col;
        out.nrows = src.n_rows();
        out.ncols = src.n_cols();
        return out;
    }

    // :102-110 — a row is the zip of two slices; rows past the end are empty, not an error.
    template <class F> void for_row(std::size_t row, F&& f) const {
        if (row >= nrows) return;
        const std::size_t b = static_cast<std::size_t>(offset_rows[row]);
        const std::size_t e = static_cast<std::size_t>(offset_rows[row + 1]);
        for (std::size_t k = b; k < e; ++k)
            if (!f(columns[k], values[k])) break;
    }

    static constexpr std::size_t npos() { return static_cast<std::size_t>(unset<I>()); }

    std::size_t find_index(std::size_t i, std::size_t j) const {            // :54-67
        std::size_t hit = npos();
        if (i < nrows) {
            const std::size_t b = static_cast<std::size_t>(offset_rows[i]);
            const std::size_t e = static_cast<std::size_t>(offset_rows[i + 1]);
            for (std::size_t k = b; k < e; ++k)
                if (static_cast<std::size_t>(columns[k]) == j) { hit = k; break; }
        }
        return hit;
    }
    // :71-92 — insertion at the row's *start* offset (newest first inside a row).  Reproduced
    // literally, including the quirk that the very first push does not set n_rows (:75-76).
    std::size_t push(std::size_t i, std::size_t j, T val) {
        if (j >= ncols) ncols = j + 1;
        if (offset_rows.empty()) {
            offset_rows.resize(i + 2, I(0));
        } else if (i >= nrows) {
            const I last = offset_rows.back();
            offset_rows.resize(i + 2, last);
            nrows = i + 1;
        }
        if (i >= offset_rows.size()) throw Panic("index out of bounds: SparseMatCRS::push");
        if (offset_rows[i] == unset<I>()) throw Panic("Maximum number of entries reached");
        const std::size_t at = static_cast<std::size_t>(offset_rows[i]);
        columns.insert(columns.begin() + static_cast<std::ptrdiff_t>(at), as_index<I>(j));
        values.insert(values.begin() + static_cast<std::ptrdiff_t>(at), val);
        for (std::size_t k = i + 1; k < offset_rows.size(); ++k) offset_rows[k] += I(1);
        return at;
    }
    T get(std::size_t i, std::size_t j) const {                             // :136-143
        const std::size_t e = find_index(i, j);
        return e == npos() ? T(0) : values[e];
    }
    T& get_mut(std::size_t i, std::size_t j) {                              // :145-151
        std::size_t e = find_index(i, j);
        if (e == npos()) e = push(i, j, T(0));
        return values[e];
    }
    void set(std::size_t i, std::size_t j, T val) { get_mut(i, j) = val; }
    void add_to(std::size_t i, std::size_t j, T val) { get_mut(i, j) += val; }
    void scale(T s) { for (T& e : values) e *= s; }                         // :153-157

    void assemble_column_info() {                                           // :180-191
        for (std::size_t i = 0; i < nrows; ++i) {
            const std::size_t b = static_cast<std::size_t>(offset_rows[i]);
            const std::size_t e = static_cast<std::size_t>(offset_rows[i + 1]);
            for (std::size_t k = b; k < e; ++k) {
                chains_col.push(static_cast<std::size_t>(columns[k]));
                rows.push_back(as_index<I>(i));
            }
        }
    }
    template <class F> void for_col(std::size_t col, F&& f) const {         // :193-221
        if (rows.size() != columns.size())
            throw Panic("Column iterator not available - use assemble_column_info()");
        chains_col.for_row(col, [&](std::size_t e) { return f(rows[e], values[e]); });
    }

    // sparsematrix.rs:186-210 — product via the right operand's column iterator; needed only to
    // replay the 17.9632 known answer (src/lib.rs:101-102).  Out of the GPU scope.
    template <class R> SparseMatCRS prod(const R& rhs) const {
        if (n_rows() != rhs.n_cols() || n_cols() != rhs.n_rows()) throw Panic("Dimension mismatch");
        SparseMatCRS out;
        for (std::size_t i = 0; i < n_rows(); ++i) {
            std::vector<std::pair<I, T>> cv;
            for_row(i, [&](I c, T v) { cv.emplace_back(c, v); return true; });
            std::stable_sort(cv.begin(), cv.end(), [](auto& l, auto& r) { return l.first < r.first; });
            for (std::size_t j = 0; j < rhs.n_cols(); ++j) {
                T sum = T(0);
                rhs.for_col(j, [&](I row, T vr) {
                    for (auto& [c, v] : cv) {
                        if (!(c <= row)) break;
                        if (c == row) sum += v * vr;
                    }
                    return true;
                });
                if (sum != T(0)) out.set(i, j, sum);
            }
        }
        return out;
    }
};

// ------------------------------------------------------------------------------------------------
// sparsemat_par.rs:12-140 — 1-D row blocks: block b owns global rows [b*R, (b+1)*R) under local
// row ids and *global* column ids.  What executes today is the serial default mvp through this
// dispatch; the threaded mvp_par is commented out in the reference (:37-68).
template <class M>
struct SparseMatPar {
    using value_type = typename M::value_type;
    using index_type = typename M::index_type;
    std::size_t rows_per_block = 0, n_blocks = 0;
    std::vector<M> blocks;

    SparseMatPar(std::size_t nb, std::size_t max_rows)                      // :20-28
        : rows_per_block(nb ? max_rows / nb : 0), n_blocks(nb), blocks(nb) {
        if (nb == 0) throw Panic("attempt to divide by zero");
    }
    std::pair<std::size_t, std::size_t> locate(std::size_t row) const {     // :31-35
        if (rows_per_block == 0) throw Panic("attempt to divide by zero");
        const std::size_t b = std::min(row / rows_per_block, n_blocks);     // clamps to n_blocks (sic)
        return {b, row - b * rows_per_block};
    }
    M& block_at(std::size_t b) {
        if (b >= blocks.size()) throw Panic("index out of bounds: SparseMatPar block");
        return blocks[b];
    }
    const M& block_at(std::size_t b) const {
        if (b >= blocks.size()) throw Panic("index out of bounds: SparseMatPar block");
        return blocks[b];
    }
    template <class F> void for_row(std::size_t row, F&& f) const {         // :86-89
        auto [b, r] = locate(row);
        block_at(b).for_row(r, f);
    }
    std::size_t n_rows() const {                                            // :95-107
        std::size_t last = 0;
        for (std::size_t b = 0; b < blocks.size(); ++b) {
            if (blocks[b].empty()) break;
            last = b;
        }
        return last * rows_per_block + blocks[last].n_rows();
    }
    std::size_t n_cols() const {                                            // :109-115
        std::size_t c = 0;
        for (auto& m : blocks) c = std::max(c, m.n_cols());
        return c;
    }
    std::size_t nnz() const {                                               // :117-123
        std::size_t n = 0;
        for (auto& m : blocks) n += m.nnz();
        return n;
    }
    auto get(std::size_t i, std::size_t j) const { auto [b, r] = locate(i); return block_at(b).get(r, j); }
    auto& get_mut(std::size_t i, std::size_t j) { auto [b, r] = locate(i); return block_at(b).get_mut(r, j); }
    void set(std::size_t i, std::size_t j, value_type v) { get_mut(i, j) = v; }
    void add_to(std::size_t i, std::size_t j, value_type v) { get_mut(i, j) += v; }
};

// ------------------------------------------------------------------------------------------------
// linearsolver.rs:12-61 — unpreconditioned CG exactly as written, including its clones' arithmetic
// (scale-then-add), the absolute stop test sqrt(rr) < tol evaluated in f64, and no r0 check.
struct CgStats {
    std::size_t iterations = 0;      // number of loop bodies executed
    double final_residual = 0.0;     // sqrt(rr) at exit
    bool converged = false;
};

struct ConjugateGradient {
    double tol = 1e-12;              // Default :17-24
    std::size_t iter_max = 10000;
    // Additive (NOT in the reference): when true the stop test is sqrt(rr) < tol * ||b||.
    bool relative = false;

    template <class M, class T>
    CgStats solve(const M& a, const DenseVec<T>& b, DenseVec<T>& x,
                  std::vector<double>* history = nullptr) const {
        if (a.n_rows() != a.n_cols()) throw Panic("Matrix is not symmetric");                 // :30-32
        if (a.n_rows() != b.dim() || a.n_rows() != x.dim())
            throw Panic("Matrix and vector size mismatch");                                   // :33-36
        const double threshold = relative ? tol * b.norm() : tol;
        CgStats st;
        DenseVec<T> r = b.minus(mvp(a, x));                                                   // :38
        DenseVec<T> p = r;                                                                    // :39
        T rr = r.norm_squared();                                                              // :40
        st.final_residual = std::sqrt(static_cast<double>(rr));
        for (std::size_t k = 0; k < iter_max; ++k) {                                          // :41
            DenseVec<T> ap = mvp(a, p);                                                       // :43
            const T alpha = rr / p.inner_prod(ap);                                            // :45
            x.add(p.times(alpha));                                                            // :47
            r.sub(ap.times(alpha));                                                           // :49
            const T rr_prev = rr;                                                             // :50
            rr = r.norm_squared();                                                            // :51
            st.iterations = k + 1;
            st.final_residual = std::sqrt(static_cast<double>(rr));
            if (history) history->push_back(st.final_residual);
            if (st.final_residual < threshold) { st.converged = true; break; }                // :52
            const T beta = rr / rr_prev;                                                      // :56
            p.scale(beta);                                                                    // :58
            p.add(r);                                                                         // :59
        }
        return st;
    }
};

// ------------------------------------------------------------------------------------------------
// Raw-array forms of the same loops (what the bench times; identical arithmetic to mvp() above
// for a SparseMatCRS, without the per-element bounds checks so the CPU baseline is not handicapped).
template <class T, class I>
void mvp_crs_raw(std::size_t n_rows, const T* values, const I* columns, const I* offset_rows,
                 const T* x, T* y) {
    for (std::size_t i = 0; i < n_rows; ++i) {
        T sum = T(0);
        const std::size_t b = static_cast<std::size_t>(offset_rows[i]);
        const std::size_t e = static_cast<std::size_t>(offset_rows[i + 1]);
        for (std::size_t k = b; k < e; ++k) sum += x[static_cast<std::size_t>(columns[k])] * values[k];
        y[i] = sum;
    }
}

// EXTENSION, not reference behaviour: the thread-per-block completion of the commented-out
// mvp_par (sparsemat_par.rs:39-67) — shared read-only x, each thread writes its own row range.
// Per-row arithmetic is the same sequential fold, so results equal mvp_crs_raw bit for bit.
template <class T, class I>
void mvp_crs_threads(std::size_t n_rows, const T* values, const I* columns, const I* offset_rows,
                     const T* x, T* y, unsigned n_threads) {
    if (n_threads <= 1 || n_rows < 2 * static_cast<std::size_t>(n_threads)) {
        mvp_crs_raw(n_rows, values, columns, offset_rows, x, y);
        return;
    }
    std::vector<std::thread> pool;
    const std::size_t per = n_rows / n_threads;                 // R = max_n_rows / n_blocks (:21)
    for (unsigned t = 0; t < n_threads; ++t) {
        const std::size_t lo = t * per;
        const std::size_t hi = (t + 1 == n_threads) ? n_rows : lo + per;
        pool.emplace_back([=] {
            for (std::size_t i = lo; i < hi; ++i) {
                T sum = T(0);
                const std::size_t b = static_cast<std::size_t>(offset_rows[i]);
                const std::size_t e = static_cast<std::size_t>(offset_rows[i + 1]);
                for (std::size_t k = b; k < e; ++k) sum += x[static_cast<std::size_t>(columns[k])] * values[k];
                y[i] = sum;
            }
        });
    }
    for (auto& th : pool) th.join();
}

// View over raw CRS arrays so ConjugateGradient::solve / mvp() can run on them.
template <class T, class I>
struct CrsView {
    using value_type = T;
    using index_type = I;
    std::size_t nrows, ncols, n_nz;
    const T* values; const I* columns; const I* offset_rows;
    std::size_t n_rows() const { return nrows; }
    std::size_t n_cols() const { return ncols; }
    std::size_t nnz() const { return n_nz; }
    template <class F> void for_row(std::size_t row, F&& f) const {
        if (row >= nrows) return;
        const std::size_t b = static_cast<std::size_t>(offset_rows[row]);
        const std::size_t e = static_cast<std::size_t>(offset_rows[row + 1]);
        for (std::size_t k = b; k < e; ++k)
            if (!f(columns[k], values[k])) break;
    }
};

}  // namespace oracle
