"""GPU parity of SparseMatrix::transpose (sparsematrix.rs:174-183) on the device (csrc/transpose.cu: a stable radix sort of
the entries by column) against the oracle's restatement of the reference loop on the assembly format followed by to_crs:
bit-exact layout (integer / copy work), including the result's dimensions."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
COMBOS = [(np.float32, np.uint32), (np.float64, np.uint32), (np.float32, np.uint64), (np.float64, np.uint64)]


def _dedup(n_rows, n_cols, vals, cols, offs):
    """Drop repeated (row, col) pairs (keeping the first): the reference's assembly cannot hold them (get_mut finds the
    existing entry), and transpose of such a matrix is what the oracle can restate."""
    o = offs.astype(np.int64)
    keep = np.ones(vals.size, bool)
    for r in range(n_rows):
        seen = set()
        for k in range(o[r], o[r + 1]):
            c = int(cols[k])
            if c in seen:
                keep[k] = False
            seen.add(c)
    lens = np.array([keep[o[r]:o[r + 1]].sum() for r in range(n_rows)], np.int64)
    no = np.zeros(n_rows + 1, np.int64)
    np.cumsum(lens, out=no[1:])
    return n_rows, n_cols, vals[keep], cols[keep], no.astype(offs.dtype)


def _oracle_transpose(orc, vdt, idt, n_rows, vals, cols, offs):
    il = orc.IndexListMat(vdt, idt)
    o = offs.astype(np.int64)
    rows = np.repeat(np.arange(n_rows, dtype=np.uint64), o[1:] - o[:-1])
    if vals.size:
        il.set(rows, cols.astype(np.uint64), vals)             # rows ascending, storage order inside a row
    return il.transpose().to_crs()


@pytest.mark.parametrize("vdt,idt", COMBOS)
def test_transpose_is_bit_exact(smb, orc, ctx, vdt, idt):
    mats = [_dedup(*cases.ragged(61, 700, 900, 12, vdt, idt)),                     # unsorted columns, empty rows, trailing empty row
            _dedup(*cases.powerlaw(62, 500, 300, 200, vdt, idt)),                  # long rows -> long columns
            _dedup(*cases.banded(63, 2500, 40, 7, vdt, idt)),
            _dedup(*cases.giant_row(64, 60, 9000, 5000, vdt, idt)),                # more than one radix tile, > 1 digit
            cases.all_empty(30, 40, vdt, idt)]
    for n_rows, n_cols, vals, cols, offs in mats:
        a = smb.SparseMatCRS.from_raw_parts(ctx, n_rows, n_cols, vals, cols, offs)
        t = a.transpose()
        wr, wc, wv, wcols, woffs = _oracle_transpose(orc, vdt, idt, n_rows, vals, cols, offs)
        assert (t.n_rows(), t.n_cols(), t.n_non_zero_entries()) == (wr, wc, wv.size)
        gv, gc, go = t.raw_parts()
        assert gv.tobytes() == wv.tobytes() and np.array_equal(gc, wcols) and np.array_equal(go, woffs)
        if vals.size:
            # (A^T)^T has A's entries, rows sorted by column (a stable sort of each row); and A^T x == the oracle's product
            tt = t.transpose()
            assert tt.n_non_zero_entries() == vals.size
            x = orc.uniform(vdt, 3, t.n_cols())
            t.configure(smb.SPMV_SCALAR)                                  # one thread per row: the reference's storage-order sums
            assert np.array_equal(t.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy(), orc.mvp(wv, wcols, woffs, x))


def test_transpose_of_the_reference_test_matrix(smb, orc, ctx):
    """The 3 x 3 matrix of src/lib.rs:54-66, assembled in the reference's order, transposed on the device."""
    sp = smb.SparseMatIndexList(np.float32, np.uint32)
    ref = orc.IndexListMat(np.float32, np.uint32)
    for i, j, v in [(0, 1, 4.2), (0, 2, 0.12), (1, 2, 4.12), (1, 1, 2.24), (2, 0, 1.12), (0, 0, 7.12)]:
        sp.set(i, j, np.float32(v))
        ref.set(i, j, np.float32(v))
    t = sp.to_crs(ctx).transpose()
    wr, wc, wv, wcols, woffs = ref.transpose().to_crs()
    gv, gc, go = t.raw_parts()
    assert (t.n_rows(), t.n_cols()) == (wr, wc) == (3, 3)
    assert np.array_equal(gv, wv) and np.array_equal(gc, wcols) and np.array_equal(go, woffs)
    assert [(int(c), float(v)) for c, v in t.iter_row(2)] == [(0, float(np.float32(0.12))), (1, float(np.float32(4.12)))]


def test_transpose_full_size_properties(smb, ctx):
    """128^3 Laplacian (14.6 M entries, symmetric with sorted rows): the transpose equals the matrix itself, array by array."""
    n = 128
    a = smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, n, n, n)
    t = a.transpose()
    for got, want in zip(t.raw_parts(), a.raw_parts()):
        assert np.array_equal(got, want)
