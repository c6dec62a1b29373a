"""CPU: independent cross-checks of the oracle (SURVEY.md §8c "secondary cross-check").

The oracle is pinned by the reference's own known answers (test_oracle_golden.py).  Here its compiled loops are checked
on seeded random inputs against (a) a ten-line pure-Python/numpy-scalar restatement of the same reference lines — same
order, same roundings, so bit-for-bit — and (b) scipy.sparse, whose csr_matvec also sums a row sequentially, within the
north-star tolerance.  Small sizes: the Python loops are slow on purpose."""
import numpy as np
import pytest

import cases

scipy_sparse = pytest.importorskip("scipy.sparse")

COMBOS = [(np.float32, np.uint32), (np.float64, np.uint32), (np.float32, np.uint64), (np.float64, np.uint64)]


def py_mvp(values, columns, offsets, x):
    """sparsematrix.rs:146-158 with numpy scalars of the value type: `sum += rhs.get(j) * val`, two roundings per entry."""
    T = values.dtype.type
    y = np.empty(offsets.size - 1, values.dtype)
    for i in range(y.size):
        s = T(0)
        for k in range(int(offsets[i]), int(offsets[i + 1])):
            s = T(s + T(x[int(columns[k])] * values[k]))
        y[i] = s
    return y


def py_to_crs(n_rows, rows, cols, vals):
    """IndexList semantics (indexlist.rs:62-83, sparsemat_indexlist.rs:29-53,158-164) + to_crs (sparsemat_crs.rs:24-36) on
    plain Python lists: first (i, j) creates the entry at the row's tail, later sets of the same (i, j) overwrite it."""
    per_row = [[] for _ in range(n_rows)]
    for i, j, v in zip(rows, cols, vals):
        for e in per_row[i]:
            if e[0] == j:
                e[1] = v
                break
        else:
            per_row[i].append([j, v])
    offs, c, v = [0], [], []
    for r in per_row:
        c += [e[0] for e in r]
        v += [e[1] for e in r]
        offs.append(len(c))
    return np.array(v), np.array(c), np.array(offs)


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_mvp_against_python_restatement_and_scipy(orc, vdt, idt):
    for seed, (n_rows, n_cols, max_len) in enumerate([(60, 45, 9), (200, 300, 40), (7, 3, 3), (1, 1, 1)]):
        n_rows, n_cols, vals, cols, offs = cases.ragged(100 + seed, n_rows, n_cols, max_len, vdt, idt)
        x = np.random.default_rng(seed).uniform(-1, 1, n_cols).astype(vdt)
        got = orc.mvp(vals, cols, offs, x)
        assert np.array_equal(got, py_mvp(vals, cols, offs, x))                       # same order, same roundings
        assert np.array_equal(got, orc.mvp(vals, cols, offs, x, threads=3))           # the threaded completion too
        a = scipy_sparse.csr_matrix((vals.astype(np.float64), cols.astype(np.int64), offs.astype(np.int64)), shape=(n_rows, n_cols))
        ref = a @ x.astype(np.float64)
        scale = cases.abs_rowsum(vals, cols, offs, x) + np.finfo(np.float64).tiny
        assert float(np.max(np.abs(got.astype(np.float64) - ref) / scale, initial=0.0)) <= cases.TOL[np.dtype(vdt)]


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_indexlist_to_crs_against_python_restatement(orc, vdt, idt):
    rng = np.random.default_rng(77)
    for n_rows, n_cols, n_ops in [(30, 20, 400), (5, 5, 60), (100, 7, 300)]:
        rows = rng.integers(0, n_rows, n_ops)
        cols = rng.integers(0, n_cols, n_ops)
        vals = rng.uniform(-1, 1, n_ops).astype(vdt)
        m = orc.IndexListMat(vdt, idt)
        m.set(rows, cols, vals)
        r, c, gv, gc, go = m.to_crs()
        top = int(rows.max()) + 1                                                     # n_rows = pos_start.len() (indexlist.rs:66-68)
        wv, wc, wo = py_to_crs(top, rows.tolist(), cols.tolist(), vals.tolist())
        assert (r, c) == (top, int(cols.max()) + 1)
        assert np.array_equal(go.astype(np.int64), wo) and np.array_equal(gc.astype(np.int64), wc)
        assert np.array_equal(gv, wv.astype(vdt))


@pytest.mark.parametrize("vdt", [np.float32, np.float64])
def test_vector_kernels_against_python_restatement(orc, vdt):
    """vector.rs:50-58 (sequential fold from zero), densevec.rs:51-73 (elementwise, prefix only)."""
    T = np.dtype(vdt).type
    rng = np.random.default_rng(3)
    x, y = rng.uniform(-1, 1, 257).astype(vdt), rng.uniform(-1, 1, 257).astype(vdt)
    s = T(0)
    for a, b in zip(x, y):
        s = T(s + T(a * b))
    assert orc.dot(x, y) == s
    s = T(0)
    for a in x:
        s = T(s + T(a * a))
    assert orc.norm2sq(x) == s
    z = x.copy()
    orc.vec_add(z, y[:100])
    assert np.array_equal(z[:100], x[:100] + y[:100]) and np.array_equal(z[100:], x[100:])
    with pytest.raises(orc.OraclePanic):
        orc.vec_sub(z[:10].copy(), y)                                                 # "Dimension mismatch" (densevec.rs:61-63)


def test_single_reduction_recurrences_against_the_oracle_solver(orc):
    """The scalar recurrences of the single-reduction loop (sparsemat_b200/csrc/cg_sr.cuh: beta = g'/g, alpha = g'/(d - beta g'/alpha),
    p = r + beta p, s = w + beta s, x += alpha p, r -= alpha s, with g' = r.r and d = (A r).r taken behind ONE product) restated in
    numpy and run beside the oracle's ConjugateGradient::solve (linearsolver.rs:27-61) on a 3-D Laplacian: same residual
    history over the first iterations, same iteration count within 2 %, same solution.  Guards the algebra the CUDA kernels
    implement; their own parity tests are tests/test_gpu_cg_tocrs.py::test_single_reduction_* and tests/test_gpu_dist.py."""
    n = 14
    N = n ** 3
    vals, cols, offs = orc.laplace(np.float64, np.uint32, n, n, n)
    b = orc.mvp(vals, cols, offs, orc.uniform(np.float64, 6, N))
    tol = 1e-10
    xo = np.zeros(N)
    so = orc.cg(N, N, vals, cols, offs, b, xo, tol=tol, relative=True, iter_max=1000, history_cap=1000)
    assert so["converged"]
    thresh = tol * float(np.sqrt(b @ b))
    A = lambda v: orc.mvp(vals, cols, offs, np.ascontiguousarray(v))  # noqa: E731
    x = np.zeros(N)
    r = b - A(x)
    w = A(r)
    g, d = float(r @ r), float(w @ r)
    alpha, beta = g / d, 0.0
    p, s = np.zeros(N), np.zeros(N)
    hist = []
    for _ in range(1000):
        p = r + beta * p
        s = w + beta * s
        x = x + alpha * p
        r = r - alpha * s
        w = A(r)
        gn, d = float(r @ r), float(w @ r)
        hist.append(np.sqrt(gn))
        if hist[-1] < thresh:
            break
        beta = gn / g
        alpha = gn / (d - beta * gn / alpha)
        g = gn
    assert abs(len(hist) - so["iterations"]) <= max(2, so["iterations"] // 50), (len(hist), so["iterations"])
    k = min(20, len(hist), so["history"].size)
    assert np.allclose(hist[:k], so["history"][:k], rtol=1e-9)
    assert np.allclose(x, xo, rtol=0, atol=1e-8)
