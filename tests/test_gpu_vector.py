"""GPU parity of the dense-vector kernels the CG solver lives in: DenseVec::{add,sub,scale}
(densevec.rs:51-73), their compositions at linearsolver.rs:47-59, and Vector::{inner_prod,norm_squared,norm}
(vector.rs:50-63).  Elementwise results are bit-exact (separate multiply and add roundings, never an FMA);
reductions are re-ordered (fixed-order tree, partials combined in f64).  The reference's own left-to-right
fold in T carries a forward error of up to n*eps(T)/2 * sum|x_i y_i|, so parity with it is held to
max(1e-5 | 1e-12, n*eps/2) * sum|x_i y_i|, while the GPU value itself must be within 1e-5 (f32) / 1e-12 (f64)
of the exactly rounded sum."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

SIZES = [0, 1, 3, 4, 5, 255, 1024, 4097, 1_000_003]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_elementwise_ops_are_bit_exact(smb, orc, ctx, dt):
    rng = np.random.default_rng(1)
    for n in SIZES:
        x = rng.uniform(-3, 3, n).astype(dt)
        y = rng.uniform(-3, 3, n).astype(dt)
        s = dt(rng.uniform(-2, 2))
        xd, yd = smb.DenseVec.from_vec(ctx, x), smb.DenseVec.from_vec(ctx, y)
        # add / sub / scale against the oracle's restatement
        w = x.copy(); orc.vec_add(w, y)
        t = xd.clone(); t.add(yd)
        assert np.array_equal(t.to_numpy(), w)
        w = x.copy(); orc.vec_sub(w, y)
        t = xd.clone(); t.sub(yd)
        assert np.array_equal(t.to_numpy(), w)
        w = x.copy(); orc.vec_scale(w, s)
        t = xd.clone(); t.scale(s)
        assert np.array_equal(t.to_numpy(), w)
        # *x += p.clone() * alpha  (linearsolver.rs:47): product rounded first, then the add
        w = y.copy(); orc.vec_scale(w, s); z = x.copy(); orc.vec_add(z, w)
        t = xd.clone(); t.axpy(s, yd)
        assert np.array_equal(t.to_numpy(), z)
        # p.scale(beta); p.add(&r)  (linearsolver.rs:58-59)
        z = x.copy(); orc.vec_scale(z, s); orc.vec_add(z, y)
        t = xd.clone(); t.scale_add(s, yd)
        assert np.array_equal(t.to_numpy(), z)
        # operators (densevec.rs:76-130)
        assert np.array_equal((xd + yd).to_numpy(), x + y)
        assert np.array_equal((xd - yd).to_numpy(), x - y)
        assert np.array_equal((xd * s).to_numpy(), x * s)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_shorter_rhs_touches_only_its_prefix(smb, ctx, dt):
    """densevec.rs:51-58: `for (i, val) in rhs.iter().enumerate()` — only rhs.dim() entries change; a longer
    rhs panics with "Dimension mismatch"."""
    x = np.arange(10, dtype=dt)
    y = np.ones(4, dt)
    xd, yd = smb.DenseVec.from_vec(ctx, x), smb.DenseVec.from_vec(ctx, y)
    xd.add(yd)
    want = x.copy(); want[:4] += 1
    assert np.array_equal(xd.to_numpy(), want)
    with pytest.raises(smb.Panic, match="Dimension mismatch"):
        yd.add(xd)
    with pytest.raises(smb.Panic, match="Dimension mismatch"):
        yd.sub(xd)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_reductions_within_tolerance(smb, orc, ctx, dt):
    rng = np.random.default_rng(2)
    tol = cases.TOL[np.dtype(dt)]
    for n in SIZES + [16_777_216]:
        x = rng.uniform(-1, 1, n).astype(dt)
        y = rng.uniform(-1, 1, n + 3).astype(dt)        # zip stops at the shorter vector (vector.rs:50-53)
        xd, yd = smb.DenseVec.from_vec(ctx, x), smb.DenseVec.from_vec(ctx, y)
        scale = float(np.sum(np.abs(x.astype(np.float64) * y[:n].astype(np.float64))))
        ptol = max(tol, 0.5 * n * float(np.finfo(dt).eps))                    # the reference fold's own error bound
        got, want = xd.inner_prod(yd), orc.dot(x, y)
        exact = float(np.sum(x.astype(np.float64) * y[:n].astype(np.float64)))
        assert abs(float(got) - exact) <= tol * scale, (n, got, exact)
        assert abs(float(got) - float(want)) <= ptol * scale, (n, got, want)
        assert float(xd * yd) == float(got)                                   # Mul<DenseVec> (densevec.rs:133-140)
        got, want = xd.norm_squared(), orc.norm2sq(x)
        n2 = float(np.sum(x.astype(np.float64) ** 2))
        assert abs(float(got) - n2) <= tol * n2, (n, got, n2)
        assert abs(float(got) - float(want)) <= ptol * n2, (n, got, want)
        assert xd.norm() == np.sqrt(np.float64(got))                          # vector.rs:61-63: sqrt in f64
        # deterministic: the same launch shape gives the same bits
        assert xd.inner_prod(yd) == xd.inner_prod(yd)


def test_clone_and_from_vec_round_trip(smb, ctx):
    x = np.random.default_rng(3).uniform(-1, 1, 12345)
    xd = smb.DenseVec.from_vec(ctx, x)
    assert xd.dim() == x.size and np.array_equal(xd.to_numpy(), x)
    c = xd.clone()
    xd.scale(0.0)
    assert np.array_equal(c.to_numpy(), x) and not np.any(xd.to_numpy())
    with pytest.raises(smb.Panic):
        xd.get(12345)
    u = smb.DenseVec(ctx, 1000, np.float32)
    u.fill_uniform(2)
    from oracle import oracle_py as orc
    assert np.array_equal(u.to_numpy(), orc.uniform(np.float32, 2, 1000))
