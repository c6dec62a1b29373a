"""GPU parity of the hot path: SparseMatrix::mvp over SparseMatCRS (sparsematrix.rs:146-158,
sparsemat_crs.rs:102-110) — libsmb200's SpMV kernels against the oracle on identical inputs.

Bar (BASELINE.json north_star): bit-exact wherever a row is summed in storage order (scalar kernel; the
stream kernels for rows of <= 64 entries), otherwise |got - want| <= tol * (|A||x|)_i with tol = 1e-5 (f32) /
1e-12 (f64)."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

F32, F64, U32, U64 = np.float32, np.float64, np.uint32, np.uint64
COMBOS = [(F32, U32), (F64, U32), (F32, U64), (F64, U64)]


def _variants(smb):
    return [(smb.SPMV_SCALAR, 0), (smb.SPMV_VECTOR, 2), (smb.SPMV_VECTOR, 4), (smb.SPMV_VECTOR, 8), (smb.SPMV_VECTOR, 16),
            (smb.SPMV_VECTOR, 32), (smb.SPMV_STREAM, 0), (smb.SPMV_STREAM_TMA, 0), (smb.SPMV_BANDED, 0), (smb.SPMV_STREAM_PIPE, 0), (smb.SPMV_RING, 0), (smb.SPMV_AUTO, 0)]


def _check(smb, orc, ctx, case, seed=11, x_extra=0):
    n_rows, n_cols, vals, cols, offs = case
    x = np.random.default_rng(seed).uniform(-1, 1, n_cols + x_extra).astype(vals.dtype)
    want = orc.mvp(vals, cols, offs, x)
    scale = cases.abs_rowsum(vals, cols, offs, x) + np.finfo(np.float64).tiny
    tol = cases.TOL[vals.dtype]
    max_len = int(np.max(np.diff(offs.astype(np.int64)))) if n_rows else 0
    a = smb.SparseMatCRS.from_raw_parts(ctx, n_rows, n_cols, vals, cols, offs)
    assert (a.n_rows(), a.n_cols(), a.n_non_zero_entries()) == (n_rows, n_cols, vals.size)
    xd = smb.DenseVec.from_vec(ctx, x)
    for variant, lanes in _variants(smb):
        a.configure(variant, lanes)
        info = a.plan_info()
        y = a.mvp(xd)
        assert y.dim() == n_rows                       # mvp returns a vector of dim n_rows (sparsematrix.rs:148-155)
        got = y.to_numpy()
        name = f"{info['variant_name']}/{lanes} {vals.dtype}/{cols.dtype} rows={n_rows} nnz={vals.size}"
        # storage-order row sums: one thread per row (stream kernels up to 64 entries, the ring with one lane per row)
        exact = variant == smb.SPMV_SCALAR or (info["variant"] >= smb.SPMV_STREAM and max_len <= 64 and
                                               not (info["variant"] == smb.SPMV_RING and info["lanes"] > 1))
        if exact:
            assert np.array_equal(got, want), f"{name}: not bit-exact ({np.count_nonzero(got != want)} rows differ)"
        else:
            err = np.abs(got.astype(np.float64) - want.astype(np.float64)) / scale
            assert np.all(np.isfinite(got)) and float(err.max(initial=0.0)) <= tol, f"{name}: rel err {err.max():.3e}"
        # the end-to-end entry point (host buffers) gives the same bits as the device-resident call
        if variant in (smb.SPMV_AUTO, smb.SPMV_STREAM):
            assert np.array_equal(a.mvp_host(x), got), name + " (host path)"


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_laplace_2d_and_3d_generated_on_device(smb, orc, ctx, vdt, idt):
    """The bench workloads at toy size: device generator == oracle generator bit for bit, then SpMV."""
    for nx, ny, nz in [(33, 17, 1), (20, 12, 9), (1, 1, 1), (5, 1, 1), (2, 2, 2)]:
        a = smb.SparseMatCRS.laplace(ctx, vdt, idt, nx, ny, nz)
        vals, cols, offs = orc.laplace(vdt, idt, nx, ny, nz)
        gv, gc, go = a.raw_parts()
        assert np.array_equal(gv, vals) and np.array_equal(gc, cols) and np.array_equal(go, offs)
        _check(smb, orc, ctx, (nx * ny * nz, nx * ny * nz, vals, cols, offs))


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_ragged_rows_with_empty_rows(smb, orc, ctx, vdt, idt):
    _check(smb, orc, ctx, cases.ragged(1, 3000, 2500, 40, vdt, idt))
    _check(smb, orc, ctx, cases.ragged(2, 257, 4096, 130, vdt, idt))      # rows past the 64-entry storage-order limit


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_power_law_rows(smb, orc, ctx, vdt, idt):
    _check(smb, orc, ctx, cases.powerlaw(3, 20000, 20000, 6000, vdt, idt))


@pytest.mark.parametrize("vdt,idt", [(F32, U32), (F64, U64)])
def test_one_giant_row_among_short_ones(smb, orc, ctx, vdt, idt):
    for where in (0, None, 1499):
        _check(smb, orc, ctx, cases.giant_row(4, 1500, 50000, 120_000, vdt, idt, where=where))


@pytest.mark.parametrize("vdt,idt", [(F32, U32), (F64, U32)])
def test_banded_matrix_uses_the_shared_memory_x_window(smb, orc, ctx, vdt, idt):
    case = cases.banded(5, 40000, 600, 9, vdt, idt)
    a = smb.SparseMatCRS.from_raw_parts(ctx, *case)
    a.configure(smb.SPMV_BANDED)
    assert a.plan_info()["variant"] == smb.SPMV_BANDED
    _check(smb, orc, ctx, case)


@pytest.mark.parametrize("vdt,idt", [(F32, U32), (F64, U64)])
def test_edge_shapes(smb, orc, ctx, vdt, idt):
    _check(smb, orc, ctx, cases.all_empty(7, 3, vdt, idt))                 # rows without entries: y = 0
    _check(smb, orc, ctx, cases.ragged(6, 1, 10, 10, vdt, idt, empty_frac=0.0))
    _check(smb, orc, ctx, cases.ragged(7, 50, 1, 5, vdt, idt))             # a single column
    _check(smb, orc, ctx, cases.ragged(8, 300, 200, 12, vdt, idt), x_extra=17)   # x longer than n_cols is fine
    # 0 x 0 matrix (SparseMatCRS::new()): mvp gives the empty vector
    a = smb.SparseMatCRS.from_raw_parts(ctx, 0, 0, np.zeros(0, vdt), np.zeros(0, idt), np.zeros(0, idt))
    assert a.mvp(smb.DenseVec(ctx, 0, vdt)).dim() == 0


@pytest.mark.parametrize("chunks", ["3", "7"])
def test_host_buffer_pipeline_chunks(smb, orc, ctx, chunks, monkeypatch):
    """smb200_spmv_host cuts the rows into chunks (H2D pieces of x | compute | D2H slices of y overlap); forced here on
    small matrices: banded (a chunk waits for a few pieces of x), unstructured (waits for all of x), ragged ends."""
    monkeypatch.setenv("SMB200_HOST_CHUNKS", chunks)
    for case in (cases.banded(12, 9000, 300, 7, F32, U32), cases.ragged(13, 5000, 7000, 25, F64, U64),
                 cases.powerlaw(14, 4000, 4000, 3000, F64, U32), cases.all_empty(2500, 10, F32, U32)):
        n_rows, n_cols, vals, cols, offs = case
        x = np.random.default_rng(3).uniform(-1, 1, n_cols).astype(vals.dtype)
        a = smb.SparseMatCRS.from_raw_parts(ctx, n_rows, n_cols, vals, cols, offs)
        want = a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy()
        hx, hy = smb.pinned_empty(n_cols, vals.dtype), smb.pinned_empty(n_rows, vals.dtype)
        hx[:] = x
        for _ in range(2):
            hy[:] = -7.0
            a.mvp_host(hx, hy)
            assert np.array_equal(hy, want)
        assert np.array_equal(a.mvp_host(x), want)                            # pageable host memory works too
        a.configure(smb.SPMV_SCALAR)                                          # re-planning rebuilds the chunk plans
        assert np.array_equal(a.mvp_host(hx, hy), orc.mvp(vals, cols, offs, x))


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_ring_kernel_windows_and_fallbacks(smb, orc, ctx, vdt, idt):
    """RING stages everything by TMA, including up to four windows of x per block.  Stencils get windows; a short-row matrix
    with scattered columns gets none (global gathers); a borrowed, misaligned x must not be bulk-copied; rows of 33..256 entries take
    several lanes per row (tolerance), longer ones fall back to STREAM.  One lane per row: bit-exact (storage-order sums)."""
    for nx, ny, nz in [(40, 30, 1), (24, 20, 18)]:
        vals, cols, offs = orc.laplace(vdt, idt, nx, ny, nz)
        n = nx * ny * nz
        a = smb.SparseMatCRS.from_raw_parts(ctx, n, n, vals, cols, offs).configure(smb.SPMV_RING)
        info = a.plan_info()
        assert info["variant"] == smb.SPMV_RING and info["n_xwin_blocks"] == info["n_blocks"]   # every block got its windows
        # ... and with them 16-bit window positions instead of its columns (plan-time index compression); the Laplacian has
        # two distinct values, so every block also gets a dictionary and 8-bit value codes (value indexing), and the padding
        # of the sliced-ELLPACK stage order is small
        assert info["nnz_c16"] == vals.size and info["nnz_v8"] == vals.size
        assert vals.size <= info["sell_entries"] <= 1.2 * vals.size + 4096
        assert info["stream_bytes"] < info["algorithmic_bytes"] - vals.size * (np.dtype(idt).itemsize - 2)
        # without value indexing: the packed CRS-order stage with 16-bit block-relative row offsets
        os.environ["SMB200_RING_V8"] = "0"
        try:
            a.configure(smb.SPMV_RING)
            info = a.plan_info()
            assert info["nnz_c16"] == vals.size and info["rows_o16"] == n and info["nnz_v8"] == 0 and info["sell_entries"] == 0
            assert info["stream_bytes"] == info["algorithmic_bytes"] - (vals.size + n) * (np.dtype(idt).itemsize - 2)
            x0 = orc.uniform(vdt, 21, n)
            assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x0)).to_numpy(), orc.mvp(vals, cols, offs, x0))
            # value indexing in CRS stage order (no sliced-ELLPACK)
            os.environ["SMB200_RING_V8"] = "1"
            os.environ["SMB200_RING_SELL"] = "0"
            a.configure(smb.SPMV_RING)
            info = a.plan_info()
            assert info["nnz_v8"] == vals.size and info["sell_entries"] == 0 and info["rows_o16"] == n
            assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x0)).to_numpy(), orc.mvp(vals, cols, offs, x0))
        finally:
            os.environ.pop("SMB200_RING_V8", None)
            os.environ.pop("SMB200_RING_SELL", None)
        a.configure(smb.SPMV_RING)
        x = orc.uniform(vdt, 21, n)
        want = orc.mvp(vals, cols, offs, x)
        xd = smb.DenseVec.from_vec(ctx, x)
        assert np.array_equal(a.mvp(xd).to_numpy(), want)
        # borrowed device memory, 4 or 8 bytes off 16-byte alignment, no padding behind it: same bits
        big = smb.DenseVec.from_vec(ctx, np.concatenate([[0], x]).astype(vdt))
        xw = smb.DenseVec.wrap(ctx, big.device_ptr() + np.dtype(vdt).itemsize, n, vdt)
        assert np.array_equal(a.mvp(xw).to_numpy(), want)
        lhs = orc.uniform(vdt, 22, n)
        got = a.inner_prod(smb.DenseVec.from_vec(ctx, lhs), xd)               # fused dot: weights staged by TMA too
        ref = float(np.sum(lhs.astype(np.float64) * want.astype(np.float64)))
        assert abs(float(got) - ref) <= 1e-5 * float(np.sum(np.abs(lhs.astype(np.float64) * want.astype(np.float64))))
    case = cases.ragged(41, 6000, 900_000, 9, vdt, idt)                        # scattered columns: no windows
    a = smb.SparseMatCRS.from_raw_parts(ctx, *case).configure(smb.SPMV_RING)
    assert a.plan_info()["variant"] == smb.SPMV_RING and a.plan_info()["n_xwin_blocks"] == 0
    assert a.plan_info()["nnz_c16"] == 0 and a.plan_info()["rows_o16"] == 0
    assert a.plan_info()["stream_bytes"] == a.plan_info()["algorithmic_bytes"]
    x = np.random.default_rng(1).uniform(-1, 1, case[1]).astype(vdt)
    assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), orc.mvp(case[2], case[3], case[4], x))
    long_rows = cases.ragged(42, 300, 5000, 90, vdt, idt)                      # rows of 33..256 entries: several lanes per row
    a = smb.SparseMatCRS.from_raw_parts(ctx, *long_rows).configure(smb.SPMV_RING)
    assert a.plan_info()["variant"] == smb.SPMV_RING and a.plan_info()["lanes"] == 4
    x = np.random.default_rng(2).uniform(-1, 1, long_rows[1]).astype(vdt)
    want = orc.mvp(long_rows[2], long_rows[3], long_rows[4], x)
    got = a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy()
    scale = cases.abs_rowsum(long_rows[2], long_rows[3], long_rows[4], x) + np.finfo(np.float64).tiny
    assert float((np.abs(got.astype(np.float64) - want.astype(np.float64)) / scale).max()) <= cases.TOL[np.dtype(vdt)]
    longer = cases.ragged(43, 300, 5000, 400, vdt, idt)                        # rows > 256 entries: not RING's business
    a = smb.SparseMatCRS.from_raw_parts(ctx, *longer).configure(smb.SPMV_RING)
    assert a.plan_info()["variant"] == smb.SPMV_STREAM


def test_reference_known_answers(smb, ctx):
    """lib.rs:80-82 (34.544, storage order [col1, col2, col0]) and lib.rs:150-152 (20.16), f32, assert_eq!."""
    a = smb.SparseMatCRS.from_raw_parts(ctx, 3, 3, np.array([4.2, 0.12, 7.12, 4.12, 2.24, 2.12], F32),
                                        np.array([1, 2, 0, 2, 1, 2], U32), np.array([0, 3, 5, 6], U32))
    v = smb.DenseVec.from_vec(ctx, np.array([2.0, 4.8, 1.2], F32))
    for variant in (smb.SPMV_AUTO, smb.SPMV_SCALAR, smb.SPMV_STREAM, smb.SPMV_STREAM_TMA, smb.SPMV_BANDED, smb.SPMV_STREAM_PIPE, smb.SPMV_RING):
        a.configure(variant)
        assert (a * v).get(0) == F32(34.544)
    assert a.density() == 6.0 / 9.0                                        # lib.rs:83
    b = smb.SparseMatCRS.from_raw_parts(ctx, 4, 4, np.array([4.2, 4.12, 2.12, 5.12, 1.12], F32),
                                        np.array([1, 2, 2, 3, 2], U32), np.array([0, 1, 2, 3, 5], U32))
    w = smb.DenseVec.from_vec(ctx, np.array([2.0, 4.8, 1.2, 3.4], F32))
    assert (b * w).get(0) == F32(20.16)
    assert b.density() == 5.0 / 16.0                                       # lib.rs:153
    assert b.iter_row(5) == []                                             # lib.rs:148-149: past the end -> empty


@pytest.mark.parametrize("seed", range(6))
def test_ring_windows_on_random_multi_diagonal_matrices(smb, orc, ctx, seed):
    """Randomised: 1-9 diagonals at random offsets (clustered or far apart, so blocks need 1..>4 windows; beyond four the block
    gathers from global memory), ragged ends, rectangular shapes, all type combinations, unsorted entries inside a row."""
    rng = np.random.default_rng(1000 + seed)
    vdt, idt = COMBOS[seed % 4]
    n_rows = int(rng.integers(3000, 60000))
    n_cols = int(n_rows * rng.uniform(0.7, 1.6))
    ndiag = int(rng.integers(1, 10))
    offsets = np.unique(np.concatenate([[0], rng.integers(-n_cols // 2, n_cols // 2, ndiag - 1) if ndiag > 1 else []]).astype(np.int64))
    if seed % 2:                                                           # clusters of neighbouring diagonals
        offsets = np.unique(np.concatenate([offsets, offsets + 1, offsets - 1]))
    rows = np.repeat(np.arange(n_rows, dtype=np.int64), offsets.size)
    cols = rows * n_cols // n_rows + np.tile(offsets, n_rows)
    keep = (cols >= 0) & (cols < n_cols) & (rng.random(cols.size) > 0.05)
    rows, cols = rows[keep], cols[keep]
    perm = np.lexsort((rng.random(rows.size), rows))                       # shuffle inside each row
    rows, cols = rows[perm], cols[perm]
    offs = np.zeros(n_rows + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=n_rows), out=offs[1:])
    vals = rng.uniform(-1, 1, cols.size).astype(vdt)
    a = smb.SparseMatCRS.from_raw_parts(ctx, n_rows, n_cols, vals, cols.astype(idt), offs.astype(idt)).configure(smb.SPMV_RING)
    info = a.plan_info()
    assert info["variant"] == smb.SPMV_RING
    x = rng.uniform(-1, 1, n_cols).astype(vdt)
    want = orc.mvp(vals, cols.astype(idt), offs.astype(idt), x)
    assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), want), (seed, info)
    a.configure(smb.SPMV_AUTO)
    assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), want), (seed, a.plan_info())
    # the compressed (16-bit window positions) and the uncompressed ring read different bytes and give the same bits
    assert 0 <= info["nnz_c16"] <= vals.size and (info["nnz_c16"] > 0) == (info["n_xwin_blocks"] > 0)
    import os
    try:
        os.environ["SMB200_RING_O16"] = "0"                                # packed (if every block has windows), full-width offsets
        a.configure(smb.SPMV_RING)
        assert a.plan_info()["rows_o16"] == 0
        assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), want), (seed, "o16 off", a.plan_info())
        os.environ["SMB200_RING_PACK"] = "0"                               # worst-case stage capacities, mixed blocks
        a.configure(smb.SPMV_RING)
        assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), want), (seed, "unpacked", a.plan_info())
        os.environ["SMB200_RING_C16"] = "0"                                # full-width columns everywhere
        a.configure(smb.SPMV_RING)
        assert a.plan_info()["nnz_c16"] == 0
        assert np.array_equal(a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), want), (seed, "uncompressed")
    finally:
        os.environ.pop("SMB200_RING_C16", None)
        os.environ.pop("SMB200_RING_PACK", None)
        os.environ.pop("SMB200_RING_O16", None)


def test_sparsemat_par_known_answer_and_blocked_product(smb, orc, ctx):
    """lib.rs:180-202 (check_sparsemat_par): with_sub_matrices(4, 16), the indexlist script, mvp row 0 == 34.544, density
    6/9 — through the completed mvp_par on the GPU; then a matrix that fills several blocks against the oracle."""
    mp = smb.SparseMatPar.with_sub_matrices(4, 16, np.float32, np.uint32)
    mp.add_to(0, 1, 4.2); mp.add_to(1, 2, 4.12); mp.add_to(2, 2, 2.12); mp.add_to(1, 1, 1.12)
    mp.add_to(1, 1, 1.12); mp.add_to(0, 2, 0.12); mp.set(0, 0, 8.12); mp.set(0, 0, 7.12)
    assert mp.get(0, 0) == F32(7.12) and mp.get(0, 1) == F32(4.2)
    y = mp.mvp(smb.DenseVec.from_vec(ctx, np.array([2.0, 4.8, 1.2], F32)))
    assert y.dim() == 3 and y.get(0) == F32(34.544)
    assert mp.density() == 6.0 / 9.0
    with pytest.raises(smb.Panic):                                         # row 16 -> block 4 of 4 (the clamp quirk)
        mp.set(16, 0, 1.0)
    # 4 blocks of 250 rows, every block full: equals the product of the assembled global matrix, bit for bit
    n_rows, n_cols, vals, cols, offs = cases.ragged(51, 1000, 800, 12, F64, U32, empty_frac=0.0)
    i = np.repeat(np.arange(n_rows), np.diff(offs.astype(np.int64)))
    big = smb.SparseMatPar(4, 1000, F64, U32)
    big.set(np.append(i, np.arange(0, 1000, 250) + 249), np.append(cols, [0, 0, 0, 0]), np.append(vals, [0.5, 0.5, 0.5, 0.5]))
    ref = orc.IndexListMat(F64, U32)
    ref.set(np.append(i, np.arange(0, 1000, 250) + 249), np.append(cols, [0, 0, 0, 0]), np.append(vals, [0.5, 0.5, 0.5, 0.5]))
    _, _, wv, wc, wo = ref.to_crs()
    x = orc.uniform(F64, 3, 800)
    assert big.n_rows() == 1000 and big.n_non_zero_entries() == wv.size
    assert np.array_equal(big.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), orc.mvp(wv, wc, wo, x))


def test_dimension_mismatch_panics_like_the_reference(smb, ctx):
    a = smb.SparseMatCRS.laplace(ctx, F64, U32, 8, 8, 1)
    with pytest.raises(smb.Panic):                                         # rhs.get(col) out of bounds (densevec.rs:40-42)
        a.mvp(smb.DenseVec(ctx, 63, F64))
    with pytest.raises(smb.SmbError):                                      # malformed CRS never reaches a kernel
        smb.SparseMatCRS.from_raw_parts(ctx, 2, 2, np.ones(2), np.array([0, 2], U32), np.array([0, 1, 2], U32))
    with pytest.raises(smb.SmbError):
        smb.SparseMatCRS.from_raw_parts(ctx, 2, 2, np.ones(2), np.array([0, 1], U32), np.array([0, 2, 1], U32))


def test_bilinear_form(smb, orc, ctx):
    """SparseMatrix::inner_prod (sparsematrix.rs:161-171): lhs^T A rhs."""
    n_rows, n_cols, vals, cols, offs = cases.ragged(9, 5000, 4000, 20, F64, U32)
    rng = np.random.default_rng(10)
    lhs, rhs = rng.uniform(-1, 1, n_rows), rng.uniform(-1, 1, n_cols)
    want = orc.bilinear(n_rows, n_cols, vals, cols, offs, lhs, rhs)
    a = smb.SparseMatCRS.from_raw_parts(ctx, n_rows, n_cols, vals, cols, offs)
    got = a.inner_prod(smb.DenseVec.from_vec(ctx, lhs), smb.DenseVec.from_vec(ctx, rhs))
    scale = float(np.sum(np.abs(lhs) * cases.abs_rowsum(vals, cols, offs, rhs)))
    assert abs(float(got) - float(want)) <= 1e-12 * scale


def test_full_size_headline_workload_properties(smb, ctx):
    """C2 at full size (256^3, f32/u32, 117M nnz): size-independent checks.
    A * ones is known in closed form: 6 - (number of in-grid neighbours), exactly representable in f32;
    linearity A(2x) == 2*A(x) bit for bit (scaling by 2 is exact); every kernel family agrees bit for bit
    (all rows have <= 7 entries, summed in storage order)."""
    n = 256
    a = smb.SparseMatCRS.laplace(ctx, F32, U32, n, n, n)
    assert a.n_non_zero_entries() == 117_047_296 and a.n_rows() == n ** 3
    ones = smb.DenseVec(ctx, n ** 3, F32)
    ones.fill(1.0)
    y = a.mvp(ones).to_numpy().reshape(n, n, n)
    idx = np.arange(n)
    edge = ((idx == 0).astype(np.float32) + (idx == n - 1).astype(np.float32))
    want = edge[:, None, None] + edge[None, :, None] + edge[None, None, :]
    assert np.array_equal(y, want)
    x = smb.DenseVec(ctx, n ** 3, F32)
    x.fill_uniform(2)
    ref = None
    for variant in (smb.SPMV_STREAM, smb.SPMV_STREAM_TMA, smb.SPMV_STREAM_PIPE, smb.SPMV_RING, smb.SPMV_SCALAR, smb.SPMV_AUTO):
        a.configure(variant)
        got = a.mvp(x).to_numpy()
        if ref is None:
            ref = got
        else:
            assert np.array_equal(got, ref), smb.VARIANT_NAMES[variant]
    x2 = x.clone()
    x2.scale(2.0)
    assert np.array_equal(a.mvp(x2).to_numpy(), 2.0 * ref)


def test_small_matrix_sweep_all_variants_types_and_entry_points():
    """scripts/sanitize_small.py: every kernel family x {f32,f64} x {u32,u64} on small ragged / power-law / banded / giant-row /
    empty matrices, plus bilinear, CG, vector ops, to_crs and the chunked host path (regression: a banded plan whose dynamic +
    static shared memory crossed 48 KB without the opt-in attribute failed to launch)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "sanitize_small.py")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "sanitize_small: ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_bandsplit_column_bands(smb, orc, ctx, vdt, idt):
    """BANDSPLIT (AUTO picks it when x is far larger than L2 and the columns have no locality; forced here): the matrix is cut into column bands at plan time and multiplied band by band, the
    first launch writing y and the others adding to it.  Tiny bands are forced here so that small matrices split into many.
    Band-major row sums: within the north-star tolerance of the oracle (not bit-exact); a row whose entries all fall into one
    band IS bit-exact; fused dot and CG ride the last band."""
    import os
    tol = cases.TOL[np.dtype(vdt)]
    try:
        for width, case in [(1024, cases.ragged(51, 3000, 5000, 30, vdt, idt)), (2048, cases.powerlaw(52, 4000, 9000, 3000, vdt, idt)),
                            (1024, cases.giant_row(53, 300, 20000, 30000, vdt, idt)), (1024, cases.all_empty(40, 5000, vdt, idt)),
                            (4096, cases.banded(54, 9000, 50, 9, vdt, idt))]:
            os.environ["SMB200_BANDSPLIT_WIDTH"] = str(width)
            n_rows, n_cols, vals, cols, offs = case
            a = smb.SparseMatCRS.from_raw_parts(ctx, *case).configure(smb.SPMV_BANDSPLIT)
            info = a.plan_info()
            if vals.size == 0:
                assert info["variant"] == smb.SPMV_STREAM                                  # nothing to split
            else:
                assert info["variant"] == smb.SPMV_BANDSPLIT and info["launches_per_spmv"] == -(-n_cols // width), info
            x = np.random.default_rng(7).uniform(-1, 1, n_cols).astype(vdt)
            xd = smb.DenseVec.from_vec(ctx, x)
            want = orc.mvp(vals, cols, offs, x)
            y = smb.DenseVec(ctx, n_rows, vdt)
            y.fill(123.0)                                                                  # the first band must overwrite, not add
            got = a.mvp(xd, out=y).to_numpy()
            scale = cases.abs_rowsum(vals, cols, offs, x) + np.finfo(np.float64).tiny
            err = np.abs(got.astype(np.float64) - want.astype(np.float64)) / scale
            assert np.all(np.isfinite(got)) and float(err.max(initial=0.0)) <= tol, (width, float(err.max(initial=0.0)))
            assert np.array_equal(a.mvp(xd).to_numpy(), got)                               # deterministic
            if vals.size:
                # scale() reaches the band parts' copies of the values (sparsemat_crs.rs:153-157): exactly 2x (a power of two)
                a.scale(2.0)
                assert np.array_equal(a.mvp(xd).to_numpy(), got * vdt(2.0)), "scale() did not reach the band-split plan"
                a.scale(0.5)
                assert np.array_equal(a.mvp(xd).to_numpy(), got)
            lhs = np.random.default_rng(8).uniform(-1, 1, n_rows).astype(vdt)
            bil = float(a.inner_prod(smb.DenseVec.from_vec(ctx, lhs), xd))
            ref = float(np.sum(lhs.astype(np.float64) * want.astype(np.float64)))
            assert abs(bil - ref) <= 1e-5 * float(np.sum(np.abs(lhs.astype(np.float64)) * scale)) + 1e-30
        # the banded case: every row lives in one or two bands -> rows inside one band are bit-exact
        one_band = (cols.astype(np.int64) // width)
        o = offs.astype(np.int64)
        single = np.array([o[r] == o[r + 1] or one_band[o[r]:o[r + 1]].min() == one_band[o[r]:o[r + 1]].max() for r in range(n_rows)])
        assert single.sum() > n_rows // 2 and np.array_equal(got[single], want[single])
        # CG through the band-split product (f64 only: the solver's tolerance)
        if vdt == np.float64:
            os.environ["SMB200_BANDSPLIT_WIDTH"] = "1024"
            lap = smb.SparseMatCRS.laplace(ctx, vdt, idt, 16, 16, 16).configure(smb.SPMV_BANDSPLIT)
            assert lap.plan_info()["variant"] == smb.SPMV_BANDSPLIT and lap.plan_info()["launches_per_spmv"] == 4
            v64, c64, o64 = orc.laplace(vdt, idt, 16, 16, 16)
            b_host = orc.uniform(vdt, 6, 4096)
            xs = smb.DenseVec(ctx, 4096, vdt)
            st = smb.ConjugateGradient(1e-10, 500).solve_with_stats(lap, smb.DenseVec.from_vec(ctx, b_host), xs)
            xo = np.zeros(4096)
            so = orc.cg(4096, 4096, v64, c64, o64, b_host, xo, tol=1e-10, iter_max=500)
            assert st["converged"] and abs(int(st["iterations"]) - so["iterations"]) <= 2, (st, so)
            assert np.allclose(xs.to_numpy(), xo, rtol=1e-8, atol=1e-10)
    finally:
        os.environ.pop("SMB200_BANDSPLIT_WIDTH", None)


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_ring_kernel_long_rows_multi_lane(smb, orc, ctx, vdt, idt):
    """FEM-like rows of 33..256 entries (27-point stencil x dof unknowns per node) stay on the TMA ring: 2 / 4 / 8 threads per
    row, strided partial sums + a fixed shuffle tree — deterministic, inside the north-star tolerance of the oracle's
    storage-order sum (not bit-exact); rows of <= 32 entries keep one thread per row and stay bit-exact."""
    tol = cases.TOL[np.dtype(vdt)]
    for dof, lanes in ((1, 1), (2, 2), (3, 4), (6, 8)):
        case = cases.fem_like(14, 11, 9, dof, vdt, idt, seed=dof)
        n_rows, n_cols, vals, cols, offs = case
        a = smb.SparseMatCRS.from_raw_parts(ctx, *case)
        info = a.plan_info()
        assert info["variant"] == smb.SPMV_RING and info["lanes"] == lanes, info            # AUTO keeps the ring
        assert info["nnz_c16"] == vals.size                                                   # every block streams 16-bit columns
        x = orc.uniform(vdt, 21, n_cols)
        xd = smb.DenseVec.from_vec(ctx, x)
        want = orc.mvp(vals, cols, offs, x)
        got = a.mvp(xd).to_numpy()
        if lanes == 1:
            assert np.array_equal(got, want)
        else:
            scale = cases.abs_rowsum(vals, cols, offs, x) + np.finfo(np.float64).tiny
            err = np.abs(got.astype(np.float64) - want.astype(np.float64)) / scale
            assert float(err.max()) <= tol, (dof, float(err.max()))
            assert np.array_equal(a.mvp(xd).to_numpy(), got)                                  # deterministic
            # same numbers as the stream kernel within the tolerance, and the fused dot agrees
            lhs = orc.uniform(vdt, 22, n_rows)
            bil = float(a.inner_prod(smb.DenseVec.from_vec(ctx, lhs), xd))
            ref = float(np.sum(lhs.astype(np.float64) * want.astype(np.float64)))
            assert abs(bil - ref) <= 1e-5 * float(np.sum(np.abs(lhs.astype(np.float64)) * scale)) + 1e-30


def test_value_indexing_is_exact_and_optional(smb, orc, ctx):
    """Plan-time value indexing (8-bit codes into per-block dictionaries of <= 256 distinct values): taken for matrices with few
    distinct values, left alone otherwise; results bit-identical either way; scale() reaches the dictionaries; special values
    (-0.0, inf, nan payloads) survive the dictionary."""
    n = 20
    vals, cols, offs = orc.laplace(F64, U32, n, n, n)
    N = n ** 3
    x = orc.uniform(F64, 5, N)
    a = smb.SparseMatCRS.from_raw_parts(ctx, N, N, vals, cols, offs)
    assert a.plan_info()["nnz_v8"] == vals.size
    xd = smb.DenseVec.from_vec(ctx, x)
    assert np.array_equal(a.mvp(xd).to_numpy(), orc.mvp(vals, cols, offs, x))
    a.scale(0.37)                                                            # sparsemat_crs.rs:153-157: every later product sees it
    assert np.array_equal(a.mvp(xd).to_numpy(), orc.mvp(vals * 0.37, cols, offs, x))
    # a handful of special values in an otherwise two-valued matrix
    v2 = vals.copy()
    v2[5], v2[77], v2[1234], v2[4000] = -0.0, np.inf, np.nan, 5e-324
    b = smb.SparseMatCRS.from_raw_parts(ctx, N, N, v2, cols, offs)
    assert b.plan_info()["nnz_v8"] == vals.size
    got, want = b.mvp(xd).to_numpy(), orc.mvp(v2, cols, offs, x)
    assert got.tobytes() == want.tobytes()                                   # NaN rows included: the same bits
    # random values: more than 256 distinct ones per block -> the plan keeps full-width values
    v3 = np.random.default_rng(3).uniform(-1, 1, vals.size)
    c = smb.SparseMatCRS.from_raw_parts(ctx, N, N, v3, cols, offs)
    info = c.plan_info()
    assert info["variant"] == smb.SPMV_RING and info["nnz_v8"] == 0 and info["sell_entries"] == 0 and info["nnz_c16"] == vals.size
    assert np.array_equal(c.mvp(xd).to_numpy(), orc.mvp(v3, cols, offs, x))
    # 300 distinct values in ONE block only: still no value indexing anywhere (the stage format is per plan)
    v4 = vals.copy()
    v4[:300] = np.arange(300) * 0.5 + 7.0
    d = smb.SparseMatCRS.from_raw_parts(ctx, N, N, v4, cols, offs)
    assert d.plan_info()["nnz_v8"] == 0
    assert np.array_equal(d.mvp(xd).to_numpy(), orc.mvp(v4, cols, offs, x))


def test_l2_persisting_window_over_x(smb, orc, ctx):
    """SMB200_FLAG_L2_PERSIST_X: an L2 access-policy window (persisting) over x for kernels that gather it from global memory.
    A cache policy only: the same bits with the window on, after switching it off again, and on a second matrix whose x
    lives elsewhere (the window follows the operand)."""
    case = cases.ragged(91, 20000, 300000, 24, F64, U32)
    n_rows, n_cols, vals, cols, offs = case
    x = orc.uniform(F64, 4, n_cols)
    want = orc.mvp(vals, cols, offs, x)
    a = smb.SparseMatCRS.from_raw_parts(ctx, *case)
    xd = smb.DenseVec.from_vec(ctx, x)
    for variant in (smb.SPMV_STREAM, smb.SPMV_SCALAR):
        a.configure(variant, 0, smb.FLAG_L2_PERSIST_X)
        assert a.plan_info()["flags"] & smb.FLAG_L2_PERSIST_X
        assert np.array_equal(a.mvp(xd).to_numpy(), want)
        assert np.array_equal(a.mvp(xd).to_numpy(), want)                    # the window is already set: reused
        x2 = smb.DenseVec.from_vec(ctx, x)                                   # another x: the window moves
        assert np.array_equal(a.mvp(x2).to_numpy(), want)
        a.configure(variant, 0, 0)
        assert not (a.plan_info()["flags"] & smb.FLAG_L2_PERSIST_X)
        assert np.array_equal(a.mvp(xd).to_numpy(), want)                    # window cleared
    # a second matrix without the flag right after one with it: no stale window
    a.configure(smb.SPMV_STREAM, 0, smb.FLAG_L2_PERSIST_X)
    a.mvp(xd)
    b = smb.SparseMatCRS.laplace(ctx, F64, U32, 20, 20, 20)
    xb = orc.uniform(F64, 5, 8000)
    vb, cb, ob = orc.laplace(F64, U32, 20, 20, 20)
    assert np.array_equal(b.mvp(smb.DenseVec.from_vec(ctx, xb)).to_numpy(), orc.mvp(vb, cb, ob, xb))
