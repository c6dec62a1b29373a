"""GPU: the additive solver layer — smb200_crs_diagonal and Jacobi-preconditioned CG.  The reference has
no preconditioner, so the checker is the same recurrence in f64 numpy/scipy: iteration counts within +-3, the TRUE residual
||b - A x|| recomputed on the host, fewer iterations than the unpreconditioned solver on a badly scaled SPD matrix, and the
reference's panics (linearsolver.rs:30-36)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _numpy_pcg(a, b, x, tol, iter_max):
    """The same recurrence in f64 numpy (scipy CSR product)."""
    dinv = 1.0 / a.diagonal()
    r = b - a @ x
    z = dinv * r
    p = z.copy()
    rz = r @ z
    it = 0
    for k in range(iter_max):
        ap = a @ p
        alpha = rz / (p @ ap)
        x += alpha * p
        r -= alpha * ap
        it = k + 1
        if np.sqrt(r @ r) < tol:
            break
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return it


def _scaled_laplacian(orc, n1, seed):
    """S L S with a strongly varying positive diagonal scaling S: SPD, and Jacobi undoes most of the scaling."""
    import scipy.sparse as sp
    vals, cols, offs = orc.laplace(np.float64, np.uint32, n1, n1, n1)
    n = n1 ** 3
    s = 10.0 ** np.random.default_rng(seed).uniform(-1.5, 1.5, n)
    rows = np.repeat(np.arange(n), np.diff(offs.astype(np.int64)))
    vals = vals * s[rows] * s[cols.astype(np.int64)]
    return n, vals, cols, offs, sp.csr_matrix((vals, cols.astype(np.int64), offs.astype(np.int64)), shape=(n, n))


def test_crs_diagonal(smb, orc, ctx):
    for vdt, idt in [(np.float32, np.uint32), (np.float64, np.uint64)]:
        vals, cols, offs = orc.laplace(vdt, idt, 9, 7, 5)
        a = smb.SparseMatCRS.from_raw_parts(ctx, 315, 315, vals, cols, offs)
        assert np.array_equal(a.diagonal().to_numpy(), np.full(315, 6.0, vdt))
    # rows without a diagonal entry give 0; the FIRST stored (i, i) wins when a raw upload holds two
    vals = np.array([1.0, 2.0, 3.0, 4.0, 5.0])
    cols = np.array([1, 0, 1, 1, 2], np.uint32)
    offs = np.array([0, 1, 4, 5], np.uint32)
    a = smb.SparseMatCRS.from_raw_parts(ctx, 3, 3, vals, cols, offs)
    assert list(a.diagonal().to_numpy()) == [0.0, 3.0, 5.0]


def test_jacobi_pcg_against_numpy_and_plain_cg(smb, orc, ctx):
    n, vals, cols, offs, a_sp = _scaled_laplacian(orc, 14, 3)
    a = smb.SparseMatCRS.from_raw_parts(ctx, n, n, vals, cols, offs)
    xs = np.random.default_rng(4).uniform(-1, 1, n)
    b = a_sp @ xs
    tol = 1e-9 * float(np.linalg.norm(b))
    x = smb.DenseVec(ctx, n, np.float64)
    st = smb.JacobiPCG(tol, 2000).solve_with_stats(a, smb.DenseVec.from_vec(ctx, b), x)
    it_np = _numpy_pcg(a_sp, b, np.zeros(n), tol, 2000)
    assert st["converged"] and abs(int(st["iterations"]) - it_np) <= 3, (st, it_np)
    # fused like cg.cu: four launches per iteration (SpMV + p.Ap, its finalize, r update, x / p update), replayed from a graph
    # in batches of 8 that may overrun the stop by two batches; no per-iteration host round trip
    assert int(st["launches"]) <= 4 * (int(st["iterations"]) + 16) + 24, st
    got = x.to_numpy()
    assert float(np.linalg.norm(b - a_sp @ got)) <= 10 * tol                       # the TRUE residual
    # relative mode gives the same stop; the preconditioner pays off on this matrix
    x2 = smb.DenseVec(ctx, n, np.float64)
    st2 = smb.JacobiPCG(1e-9, 2000, relative=True).solve_with_stats(a, smb.DenseVec.from_vec(ctx, b), x2)
    assert st2["converged"] and abs(int(st2["iterations"]) - int(st["iterations"])) <= 1
    x3 = smb.DenseVec(ctx, n, np.float64)
    st3 = smb.ConjugateGradient(tol, 20000).solve_with_stats(a, smb.DenseVec.from_vec(ctx, b), x3)
    assert st3["converged"] and int(st["iterations"]) < int(st3["iterations"]), (st, st3)
    # f32
    a32 = smb.SparseMatCRS.from_raw_parts(ctx, n, n, vals.astype(np.float32), cols, offs)
    x4 = smb.DenseVec(ctx, n, np.float32)
    st4 = smb.JacobiPCG(1e-4, 2000, relative=True).solve_with_stats(a32, smb.DenseVec.from_vec(ctx, b.astype(np.float32)), x4)
    assert st4["converged"]
    assert float(np.linalg.norm(b - a_sp @ x4.to_numpy().astype(np.float64))) <= 1e-3 * float(np.linalg.norm(b))


def test_jacobi_pcg_panics_and_zero_pivot(smb, orc, ctx):
    vals, cols, offs = orc.laplace(np.float64, np.uint32, 4, 4, 1)
    a = smb.SparseMatCRS.from_raw_parts(ctx, 16, 16, vals, cols, offs)
    with pytest.raises(smb.Panic, match="Matrix and vector size mismatch"):        # linearsolver.rs:33-36
        smb.JacobiPCG().solve(a, smb.DenseVec(ctx, 15, np.float64), smb.DenseVec(ctx, 16, np.float64))
    rect = smb.SparseMatCRS.from_raw_parts(ctx, 2, 3, np.array([1.0, 2.0]), np.array([0, 2], np.uint32), np.array([0, 1, 2], np.uint32))
    with pytest.raises(smb.Panic, match="Matrix is not symmetric"):                # linearsolver.rs:30-32
        smb.JacobiPCG().solve(rect, smb.DenseVec(ctx, 2, np.float64), smb.DenseVec(ctx, 2, np.float64))
    hole = smb.SparseMatCRS.from_raw_parts(ctx, 2, 2, np.array([1.0, 2.0]), np.array([0, 0], np.uint32), np.array([0, 1, 2], np.uint32))
    with pytest.raises(smb.SmbError):                                              # row 1 has no diagonal entry
        smb.JacobiPCG().solve(hole, smb.DenseVec(ctx, 2, np.float64), smb.DenseVec(ctx, 2, np.float64))
