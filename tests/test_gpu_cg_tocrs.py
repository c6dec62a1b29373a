"""GPU parity of (a) the IndexList -> CRS conversion (sparsemat_indexlist.rs:61-63 -> sparsemat_crs.rs:24-50),
bit-exact, and (b) ConjugateGradient::solve (linearsolver.rs:27-61) against the oracle's solver."""
import os
import subprocess

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- (a) to_crs -------------------------------------------------------------------------------------------
def _laplace2d_entries(nx, ny):
    """(i, j, v) triples of the 2-D 5-point Laplacian in ascending-column order per row."""
    r = np.arange(nx * ny, dtype=np.uint64)
    ix, iy = r % nx, r // nx
    parts = [(r[iy > 0], r[iy > 0] - nx, -1.0), (r[ix > 0], r[ix > 0] - 1, -1.0), (r, r, 4.0),
             (r[ix + 1 < nx], r[ix + 1 < nx] + 1, -1.0), (r[iy + 1 < ny], r[iy + 1 < ny] + nx, -1.0)]
    i = np.concatenate([p[0] for p in parts])
    j = np.concatenate([p[1] for p in parts])
    v = np.concatenate([np.full(p[0].size, p[2]) for p in parts])
    order = np.lexsort((j, i))
    return i[order], j[order], v[order]


@pytest.mark.parametrize("vdt,idt", [(np.float64, np.uint32), (np.float32, np.uint64)])
@pytest.mark.parametrize("scramble", [False, True])
def test_to_crs_is_bit_exact_and_keeps_insertion_order(smb, orc, ctx, vdt, idt, scramble):
    i, j, v = _laplace2d_entries(96, 64)
    if scramble:                                        # deterministic shuffle (seed 0xC0FFEE) of the insertion order
        p = np.random.default_rng(0xC0FFEE).permutation(i.size)
        i, j, v = i[p], j[p], v[p]
    sp = smb.SparseMatIndexList(vdt, idt)
    sp.set(i, j, v.astype(vdt))
    sp.add_to(i[:100], j[:100], np.full(100, 0.5, vdt))   # updates of existing entries do not append
    ref = orc.IndexListMat(vdt, idt)
    ref.set(i, j, v.astype(vdt))
    ref.add_to(i[:100], j[:100], np.full(100, 0.5, vdt))
    assert sp._dims() == ref.dims()
    for got, want in zip(sp.raw_arrays(), ref.raw_arrays()):    # host assembler == oracle assembler
        assert np.array_equal(got, want)
    a = sp.to_crs(ctx)
    n_rows, n_cols, wv, wc, wo = ref.to_crs()
    assert (a.n_rows(), a.n_cols(), a.n_non_zero_entries()) == (n_rows, n_cols, wv.size)
    gv, gc, go = a.raw_parts()
    assert gv.tobytes() == wv.tobytes() and np.array_equal(gc, wc) and np.array_equal(go, wo)
    if scramble:                                        # rows are NOT sorted: the shuffle survives the conversion
        rows_sorted = [np.all(np.diff(gc[int(go[r]):int(go[r + 1])].astype(np.int64)) > 0) for r in range(n_rows)]
        assert not all(rows_sorted)
    x = orc.uniform(vdt, 1, n_cols)
    assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), orc.mvp(wv, wc, wo, x))


def test_to_crs_edge_cases(smb, orc, ctx):
    # no entries at all -> SparseMatCRS::new(): 0 x 0 (sparsemat_crs.rs:25,47-49)
    sp = smb.SparseMatIndexList(np.float64, np.uint32)
    a = sp.to_crs(ctx)
    assert (a.n_rows(), a.n_cols(), a.n_non_zero_entries()) == (0, 0, 0)
    # empty interior rows repeat the offset; explicit zeros are stored and counted
    sp.set([0, 5, 5, 2], [3, 0, 9, 2], [1.5, 0.0, -2.0, 4.0])
    ref = orc.IndexListMat(np.float64, np.uint32)
    ref.set([0, 5, 5, 2], [3, 0, 9, 2], [1.5, 0.0, -2.0, 4.0])
    a = sp.to_crs(ctx)
    n_rows, n_cols, wv, wc, wo = ref.to_crs()
    assert (n_rows, n_cols) == (6, 10) and (a.n_rows(), a.n_cols()) == (6, 10)
    gv, gc, go = a.raw_parts()
    assert np.array_equal(gv, wv) and np.array_equal(gc, wc) and np.array_equal(go, wo)
    assert list(go) == [0, 1, 1, 2, 2, 2, 4]
    # raw-array entry point (what a Rust binding passes) with a corrupted chain is rejected, not walked forever
    cols, vals, pos, nxt = sp.raw_arrays()
    bad = nxt.copy()
    bad[1] = 1                                                            # self-loop
    with pytest.raises(smb.SmbError):
        smb.crs_from_indexlist_arrays(ctx, 6, 10, cols, vals, pos, bad)


def test_binary_crs_container_through_the_device(smb, orc, ctx, tmp_path):
    """README workflow with a file in between: assemble in IndexList order -> to_crs on the device -> save; a later run loads
    the file and gets the same three arrays bit for bit (in-row insertion order included) and the same product; a file written
    on the host alone (crsfile_write) loads too; a file whose arrays are not a CRS matrix is refused by the upload validation."""
    rng = np.random.default_rng(8)
    for vdt, idt in [(np.float32, np.uint32), (np.float64, np.uint64)]:
        sp = smb.SparseMatIndexList(vdt, idt)
        i, j = rng.integers(0, 500, 4000), rng.integers(0, 700, 4000)
        sp.set(i, j, rng.uniform(-1, 1, 4000).astype(vdt))
        a = sp.to_crs(ctx)
        path = tmp_path / f"a_{np.dtype(vdt).name}.smbcrs"
        a.save(path)
        b = smb.SparseMatCRS.load(ctx, path)
        assert (b.n_rows(), b.n_cols(), b.n_non_zero_entries()) == (a.n_rows(), a.n_cols(), a.n_non_zero_entries())
        assert b.dtype == a.dtype and b.itype == a.itype
        for got, want in zip(b.raw_parts(), a.raw_parts()):
            assert got.tobytes() == want.tobytes()
        x = rng.uniform(-1, 1, a.n_cols()).astype(vdt)
        xd = smb.DenseVec.from_vec(ctx, x)
        assert np.array_equal(b.mvp(xd).to_numpy(), a.mvp(xd).to_numpy())
        # the host-only reader sees what the device wrote
        n_rows, n_cols, v, c, o = smb.crsfile_read(path)
        assert (n_rows, n_cols) == (a.n_rows(), a.n_cols()) and np.array_equal(orc.mvp(v, c, o, x), a.mvp(xd).to_numpy())
    # the 0 x 0 matrix
    e = smb.SparseMatIndexList(np.float64, np.uint32).to_crs(ctx)
    e.save(tmp_path / "e.smbcrs")
    assert smb.SparseMatCRS.load(ctx, tmp_path / "e.smbcrs").n_rows() == 0
    # host-written file -> device
    vals, cols, offs = orc.laplace(np.float64, np.uint32, 9, 7, 5)
    smb.crsfile_write(tmp_path / "lap.smbcrs", 315, 315, vals, cols, offs)
    lap = smb.SparseMatCRS.load(ctx, tmp_path / "lap.smbcrs")
    x = orc.uniform(np.float64, 3, 315)
    assert np.array_equal(lap.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), orc.mvp(vals, cols, offs, x))
    # a well-formed container around arrays that are not a CRS matrix (a column >= n_cols)
    smb.crsfile_write(tmp_path / "bad.smbcrs", 315, 100, vals, cols, offs)
    with pytest.raises(smb.SmbError):
        smb.SparseMatCRS.load(ctx, tmp_path / "bad.smbcrs")


# ---- (b) CG -----------------------------------------------------------------------------------------------
def test_reference_cg_known_answer(smb, ctx):
    """lib.rs:36-52: [[4,1],[1,3]] x = [1,2], x0 = [2,1], default solver -> x[0] floors to 0.0909."""
    sp = smb.SparseMatIndexList(np.float64, np.uint32)
    sp.set([0, 0, 1, 1], [0, 1, 0, 1], [4.0, 1.0, 1.0, 3.0])
    a = sp.to_crs(ctx)
    b = smb.DenseVec.from_vec(ctx, [1.0, 2.0])
    x = smb.DenseVec.from_vec(ctx, [2.0, 1.0])
    st = smb.ConjugateGradient.default().solve_with_stats(a, b, x)
    assert np.floor(x.get(0) * 10000.0) / 10000.0 == 0.0909
    assert abs(x.get(1) - 7.0 / 11.0) < 1e-12 and st["converged"] and st["iterations"] <= 3


@pytest.mark.parametrize("vdt,tol,n", [(np.float64, 1e-9, 20), (np.float32, 1e-3, 16)])
def test_cg_matches_the_oracle_solver(smb, orc, ctx, vdt, tol, n):
    a = smb.SparseMatCRS.laplace(ctx, vdt, np.uint32, n, n, n)
    vals, cols, offs = orc.laplace(vdt, np.uint32, n, n, n)
    N = n ** 3
    xstar = orc.uniform(vdt, 6, N)
    b = orc.mvp(vals, cols, offs, xstar)
    for variant in (smb.SPMV_AUTO, smb.SPMV_VECTOR, smb.SPMV_STREAM_TMA, smb.SPMV_STREAM_PIPE, smb.SPMV_RING):
        a.configure(variant)
        x = smb.DenseVec(ctx, N, vdt)
        st = smb.ConjugateGradient(tol, 2000, relative=True).solve_with_stats(a, smb.DenseVec.from_vec(ctx, b), x)
        xo = np.zeros(N, vdt)
        so = orc.cg(N, N, vals, cols, offs, b, xo, tol=tol, relative=True, iter_max=2000, history_cap=2000)
        assert st["converged"] and so["converged"]
        assert abs(int(st["iterations"]) - so["iterations"]) <= max(2, so["iterations"] // 50), (st, so["iterations"])
        # residual history tracks the oracle's over the first iterations (reductions are re-ordered)
        h = smb.ConjugateGradient.history(a)
        k = min(20, h.size, so["history"].size)
        assert np.allclose(h[:k], so["history"][:k], rtol=1e-10 if vdt == np.float64 else 1e-3)
        # true residual recomputed by the oracle's mvp
        got = x.to_numpy()
        r = b.astype(np.float64) - orc.mvp(vals, cols, offs, got).astype(np.float64)
        assert np.linalg.norm(r) / np.linalg.norm(b.astype(np.float64)) <= (10 * tol if vdt == np.float64 else 5e-3)
        assert np.allclose(got, xo, rtol=0, atol=(1e-7 if vdt == np.float64 else 2e-2))


@pytest.mark.parametrize("vdt,tol,n", [(np.float64, 1e-9, 40), (np.float32, 1e-4, 24)])
def test_single_reduction_cg_tracks_the_oracle_solver(smb, orc, monkeypatch, vdt, tol, n):
    """smb200_dist_cg_solve_sr on one rank: the loop of linearsolver.rs:41-60 rearranged so that r.r and (A r).r come out of
    one reduction (additive — the reference has no such variant, so the check is against the oracle's plain CG: same
    iteration count within 2 %, same residual history over the first iterations, same solution, true residual recomputed
    by the oracle).  Without a halo the scalar step runs inside the fused dot's finalize kernel; SMB200_DIST_SELF=1 sends the
    one rank through the distributed kernels (ring launch with the halo protocol, then push | rows | wait), where the
    scalar step is the one-warp kernel behind the product."""
    c = smb.Context(0)
    c.comm_init(0, 1, None)
    N = n ** 3
    vals, cols, offs = orc.laplace(vdt, np.uint32, n, n, n)
    b_h = orc.mvp(vals, cols, offs, orc.uniform(vdt, 6, N))
    xo = np.zeros(N, vdt)
    so = orc.cg(N, N, vals, cols, offs, b_h, xo, tol=tol, relative=True, iter_max=2000, history_cap=2000)
    assert so["converged"]
    for variant, self_halo in ((smb.SPMV_AUTO, "0"), (smb.SPMV_AUTO, "1"), (smb.SPMV_VECTOR, "1")):
        monkeypatch.setenv("SMB200_DIST_SELF", self_halo)
        a = smb.DistCRS.laplace(c, vdt, np.uint32, n, n, n)
        assert a.info()["p2p"] == (self_halo == "1")
        a.local.configure(variant)
        b, x = a.new_vec(), a.new_vec()
        b.upload(b_h)
        for rep in range(2):                                     # the second solve replays the captured batch
            x.fill(0.0)
            st = smb.ConjugateGradient(tol, 2000, relative=True, single_reduce=True).solve_with_stats(a, b, x)
            assert st["converged"], st
            assert abs(int(st["iterations"]) - so["iterations"]) <= max(2, so["iterations"] // 50), (st, so["iterations"])
            h = smb.ConjugateGradient.history(a.local)
            assert h.size == st["iterations"]
            k = min(20, h.size, so["history"].size)
            assert np.allclose(h[:k], so["history"][:k], rtol=1e-9 if vdt == np.float64 else 1e-3)
            got = x.to_numpy()
            r = b_h.astype(np.float64) - orc.mvp(vals, cols, offs, got).astype(np.float64)
            assert np.linalg.norm(r) / np.linalg.norm(b_h.astype(np.float64)) <= (10 * tol if vdt == np.float64 else 5e-3)
            assert np.allclose(got, xo, rtol=0, atol=(1e-7 if vdt == np.float64 else 2e-2))
        # the reference's loop on the same object afterwards: the workspace and the captured batch are keyed by the variant
        x.fill(0.0)
        st = smb.ConjugateGradient(tol, 2000, relative=True).solve_with_stats(a, b, x)
        assert st["converged"] and abs(int(st["iterations"]) - so["iterations"]) <= max(2, so["iterations"] // 50)
        assert np.allclose(x.to_numpy(), xo, rtol=0, atol=(1e-7 if vdt == np.float64 else 2e-2))
    with pytest.raises(ValueError):
        smb.ConjugateGradient(tol, 10, single_reduce=True).solve_with_stats(smb.SparseMatCRS.laplace(c, vdt, np.uint32, 4, 4, 4),
                                                                           smb.DenseVec(c, 64, vdt), smb.DenseVec(c, 64, vdt))


@pytest.mark.parametrize("vdt", [np.float64, np.float32])
def test_cg_on_borrowed_misaligned_vectors(smb, orc, ctx, vdt):
    """b and x wrapped around caller memory one element past a 16-byte boundary (smb200_vec_wrap): the solver's fused
    kernels must take their element path instead of faulting on 128-bit accesses, and give the same answer."""
    n = 12
    N = n ** 3
    a = smb.SparseMatCRS.laplace(ctx, vdt, np.uint32, n, n, n)
    vals, cols, offs = orc.laplace(vdt, np.uint32, n, n, n)
    b = orc.mvp(vals, cols, offs, orc.uniform(vdt, 6, N))
    es = np.dtype(vdt).itemsize
    big_b, big_x = smb.DenseVec(ctx, N + 8, vdt), smb.DenseVec(ctx, N + 8, vdt)
    big_b.upload(np.concatenate([np.zeros(1, vdt), b, np.zeros(7, vdt)]))
    bw = smb.DenseVec.wrap(ctx, big_b.device_ptr() + es, N, vdt)
    xw = smb.DenseVec.wrap(ctx, big_x.device_ptr() + es, N, vdt)
    tol = 1e-9 if vdt == np.float64 else 1e-3
    st = smb.ConjugateGradient(tol, 2000, relative=True).solve_with_stats(a, bw, xw)
    x_al = smb.DenseVec(ctx, N, vdt)
    st2 = smb.ConjugateGradient(tol, 2000, relative=True).solve_with_stats(a, smb.DenseVec.from_vec(ctx, b), x_al)
    assert st["converged"] and abs(int(st["iterations"]) - int(st2["iterations"])) <= 2
    # same elementwise arithmetic; only the partial sums of r.r are grouped differently on the element path
    assert np.allclose(xw.to_numpy(), x_al.to_numpy(), rtol=0, atol=1e-7 if vdt == np.float64 else 2e-2)
    assert big_x.to_numpy()[0] == 0 and np.all(big_x.to_numpy()[N + 1:] == 0)   # nothing written outside the borrowed range


def test_cg_iter_max_and_panics(smb, ctx):
    a = smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 12, 12, 12)
    N = 12 ** 3
    b = smb.DenseVec(ctx, N, np.float64)
    b.fill(1.0)
    x = smb.DenseVec(ctx, N, np.float64)
    st = smb.ConjugateGradient(1e-30, 7).solve_with_stats(a, b, x)
    assert st["iterations"] == 7 and not st["converged"]                  # `for _k in 0..iter_max` (linearsolver.rs:41)
    for batch in ("1", "3"):                                               # batching must not change the trajectory
        os.environ["SMB200_CG_BATCH"] = batch
        x2 = smb.DenseVec(ctx, N, np.float64)
        st2 = smb.ConjugateGradient(1e-30, 7).solve_with_stats(a, b, x2)
        assert st2["iterations"] == 7 and np.array_equal(x2.to_numpy(), x.to_numpy())
    del os.environ["SMB200_CG_BATCH"]
    with pytest.raises(smb.Panic, match="Matrix and vector size mismatch"):   # linearsolver.rs:33-36
        smb.ConjugateGradient().solve(a, smb.DenseVec(ctx, N - 1, np.float64), x)
    rect = smb.SparseMatCRS.from_raw_parts(ctx, 2, 3, np.ones(2), np.array([0, 2], np.uint32), np.array([0, 1, 2], np.uint32))
    with pytest.raises(smb.Panic, match="Matrix is not symmetric"):           # linearsolver.rs:30-32
        smb.ConjugateGradient().solve(rect, smb.DenseVec(ctx, 2, np.float64), smb.DenseVec(ctx, 2, np.float64))


def test_plain_c_client_of_the_abi():
    """tests/c/abi_example.c: C99, include/smb200.h only — the reference's known answers through the raw C ABI."""
    exe = os.path.join(ROOT, "build", "abi_example")
    assert os.path.exists(exe), "build() did not produce build/abi_example"
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "abi_example: ok" in res.stdout, res.stdout + res.stderr


def test_cpp_host_mirror_replays_the_reference_tests():
    """tests/cpp/replay_reference_tests.cpp: lib.rs's hot-path tests written against the C++ mirror of the crate."""
    exe = os.path.join(ROOT, "build", "replay_reference_tests")
    assert os.path.exists(exe), "build() did not produce build/replay_reference_tests"
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
