// Replays the hot-path parts of the reference's unit tests (/root/reference/src/lib.rs) through the C++
// host mirror (sparsemat_b200/host/sparsemat.hpp) on a real GPU.  Reads like the reference's tests on
// purpose.  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <string>

#include "../../sparsemat_b200/host/sparsemat.hpp"

using namespace sparsemat;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond) do { ++g_checks; if (!(cond)) { ++g_fail; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } } while (0)

static void check_cg(const Context& ctx) {                                  // lib.rs:36-52
    SparseMatIndexList<double, uint32_t> mat;
    mat.set(0, 0, 4.0); mat.set(0, 1, 1.0); mat.set(1, 0, 1.0); mat.set(1, 1, 3.0);
    auto b = DenseVec<double>::from_vec(ctx, {1.0, 2.0});
    auto x = DenseVec<double>::from_vec(ctx, {2.0, 1.0});
    ConjugateGradient cg;                                                   // default(): 1e-12, 10_000
    auto crs = mat.to_crs(ctx);
    CgStats st = cg.solve_with_stats(crs, b, x);
    CHECK(std::floor(x.get(0) * 10000.0) / 10000.0 == 0.0909);
    CHECK(std::fabs(x.get(1) - 7.0 / 11.0) < 1e-12);
    CHECK(st.iterations >= 2 && st.iterations <= 3);
}

static void check_sparsemat_indexlist(const Context& ctx) {                 // lib.rs:54-112 (hot-path parts)
    SparseMatIndexList<float, uint32_t> sp;
    sp.add_to(0, 1, 4.2f);
    sp.add_to(1, 2, 4.12f);
    sp.add_to(2, 2, 2.12f);
    sp.add_to(1, 1, 1.12f);
    sp.get_mut(1, 1) += 1.12f;
    sp.get_mut(0, 2) += 0.12f;
    sp.get_mut(0, 0) = 8.12f;
    sp.set(0, 0, 7.12f);
    CHECK(sp.get(0, 0) == 7.12f);
    auto row2 = sp.iter_row(2);
    CHECK(row2.size() == 1 && row2[0].first == 2 && row2[0].second == 2.12f);
    auto sp_crs = sp.to_crs(ctx);
    auto v = DenseVec<float>::from_vec(ctx, {2.0f, 4.8f, 1.2f});
    auto mvp = sp_crs * v;
    CHECK(mvp.get(0) == 34.544f);                                           // lib.rs:80-82
    CHECK(mvp.dim() == 3);
    CHECK(sp_crs.density() == 6.0 / 9.0);                                   // lib.rs:83
    auto parts = sp_crs.raw_parts();                                        // to_crs keeps insertion order
    CHECK((parts.columns == std::vector<uint32_t>{1, 2, 0, 2, 1, 2}));
    CHECK((parts.values == std::vector<float>{4.2f, 0.12f, 7.12f, 4.12f, 1.12f + 1.12f, 2.12f}));
    CHECK((parts.offset_rows == std::vector<uint32_t>{0, 3, 5, 6}));
    auto row1 = sp_crs.iter_row(1);                                         // "0 2.24 4.12 " once sorted (lib.rs:97-98)
    CHECK(row1.size() == 2 && row1[0].first == 2 && row1[0].second == 4.12f && row1[1].first == 1 && row1[1].second == 2.24f);
}

static void check_sparsemat_crs(const Context& ctx) {                       // lib.rs:114-154
    // state reached by the reference's insertion script (hand-traced)
    auto m = SparseMatCRS<float, uint32_t>::from_raw_parts(ctx, 4, 4, {4.2f, 4.12f, 2.12f, 5.12f, 1.12f}, {1, 2, 2, 3, 2}, {0, 1, 2, 3, 5});
    auto row0 = m.iter_row(0);
    CHECK(row0.size() == 1 && row0[0].first == 1 && row0[0].second == 4.2f);
    CHECK(m.iter_row(5).empty());                                           // past the end: empty, no panic
    auto v = DenseVec<float>::from_vec(ctx, {2.0f, 4.8f, 1.2f, 3.4f});
    auto mvp = m * v;
    CHECK(mvp.get(0) == 20.16f);                                            // lib.rs:150-152
    CHECK(m.density() == 5.0 / 16.0);
}

static void check_panics(const Context& ctx) {
    auto a = SparseMatCRS<double, uint32_t>::from_raw_parts(ctx, 1, 2, {1.0, 1.0}, {0, 1}, {0, 2});
    auto b = DenseVec<double>::from_vec(ctx, {1.0});
    auto x = DenseVec<double>::from_vec(ctx, {0.0});
    std::string msg;
    try { ConjugateGradient().solve(a, b, x); } catch (const Panic& p) { msg = p.what(); }
    CHECK(msg == "Matrix is not symmetric");                                // linearsolver.rs:30-32
    auto sq = SparseMatCRS<double, uint32_t>::from_raw_parts(ctx, 2, 2, {1.0, 1.0}, {0, 1}, {0, 1, 2});
    msg.clear();
    try { ConjugateGradient().solve(sq, b, x); } catch (const Panic& p) { msg = p.what(); }
    CHECK(msg == "Matrix and vector size mismatch");                        // linearsolver.rs:33-36
    auto s = DenseVec<double>::from_vec(ctx, {1.0});
    auto l = DenseVec<double>::from_vec(ctx, {1.0, 2.0});
    msg.clear();
    try { s.add(l); } catch (const Panic& p) { msg = p.what(); }
    CHECK(msg == "Dimension mismatch");                                     // densevec.rs:52-54
    msg.clear();
    try { (void)sq.mvp(s); } catch (const Panic& p) { msg = p.what(); }     // x shorter than n_cols
    CHECK(msg == "Dimension mismatch");
    CHECK(par_locate(4, 16, 9) == std::make_pair(uint64_t(2), uint64_t(1)));   // sparsemat_par.rs:31-35
}

int main() {
    try {
        Context ctx(0);
        check_cg(ctx);
        check_sparsemat_indexlist(ctx);
        check_sparsemat_crs(ctx);
        check_panics(ctx);
    } catch (const std::exception& e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf("replay_reference_tests: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
