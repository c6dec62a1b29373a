"""CPU: pins the oracle (oracle/sparsemat_oracle.hpp via liboracle.so) to the reference's own known-answer
tests (tests/golden/reference_kats.json, transcribed from /root/reference/src/lib.rs) and to the committed
seeded fixtures (tests/golden/oracle_fixtures.npz, written by tests/golden/make_fixtures.py)."""
import json
import os

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_kats.json")) as f:
    KATS = json.load(f)
F32, U32 = np.float32, np.uint32


def _replay_script(mat, ops):
    for op, i, j, v in ops:
        getattr(mat, op)(i, j, v)


def test_kat_binary_replays_all_seven_reference_tests(orc):
    out = orc.run_kat()                                                  # oracle/kat.cpp: lib.rs:36-231 literally
    assert "0 failed" in out


def test_indexlist_script_and_mvp_known_answer(orc):
    k = KATS["indexlist_script"]
    sp = orc.IndexListMat(F32, U32)
    _replay_script(sp, k["ops"])
    n_rows, n_cols, nnz = sp.dims()
    assert (n_rows, n_cols, nnz) == (3, 3, 6)
    assert float(nnz) / float(n_rows * n_cols) == k["density"]["num"] / k["density"]["den"]
    nr, nc, vals, cols, offs = sp.to_crs()
    lay = k["to_crs_layout"]
    assert list(cols) == lay["columns"] and list(offs) == lay["offset_rows"]
    # 2.24 is 1.12f + 1.12f accumulated in f32 (add_to twice), the literals are f32 literals in the reference
    want_vals = np.array(lay["values_f32"], F32)
    want_vals[4] = F32(1.12) + F32(1.12)
    assert np.array_equal(vals, want_vals)
    it = [(r, int(cols[p]), vals[p]) for r in range(nr) for p in range(int(offs[r]), int(offs[r + 1]))]
    for got, want in zip(it[:4], k["iter_first4"]["value"]):
        assert got[0] == want[0] and got[1] == want[1] and got[2] == F32(want[2])
    y = orc.mvp(vals, cols, offs, np.array(k["mvp"]["x"], F32))
    assert y[k["mvp"]["row"]] == F32(k["mvp"]["value"])                  # assert_eq!(mvp.get(0), 34.544)
    # summation order matters: ascending-column order gives a different f32
    order = np.argsort(cols[:3])
    y_sorted = orc.mvp(vals[:3][order], cols[:3][order], np.array([0, 3], U32), np.array(k["mvp"]["x"], F32))
    assert y_sorted[0] != F32(k["mvp"]["value"])
    # row 1 printed in sorted order is "0 2.24 4.12 " (lib.rs:94-98): stored [(2,4.12),(1,2.24)]
    row1 = sorted((int(cols[p]), vals[p]) for p in range(int(offs[1]), int(offs[2])))
    assert "0 " + "".join(str(v) + " " for _, v in row1) == k["to_crs_row1_sorted_string"]["value"]


def test_crs_direct_known_answer(orc):
    k = KATS["crs_direct"]
    vals, cols, offs = np.array(k["values_f32"], F32), np.array(k["columns"], U32), np.array(k["offset_rows"], U32)
    y = orc.mvp_checked(k["n_rows"], k["n_cols"], vals, cols, offs, np.array(k["mvp"]["x"], F32))
    assert y[0] == F32(k["mvp"]["value"])                                # 20.16
    assert vals.size / (k["n_rows"] * k["n_cols"]) == k["density"]["num"] / k["density"]["den"]
    with pytest.raises(orc.OraclePanic):                                 # x shorter than a referenced column -> panic
        orc.mvp_checked(k["n_rows"], k["n_cols"], vals, cols, offs, np.array([1.0, 2.0], F32))


def test_cg_known_answer(orc):
    k = KATS["cg"]
    sp = orc.IndexListMat(np.float64, U32)
    for i, j, v in k["entries"]:
        sp.set(i, j, v)
    nr, nc, vals, cols, offs = sp.to_crs()
    x = np.array(k["x0"])
    st = orc.cg(nr, nc, vals, cols, offs, np.array(k["b"]), x, tol=k["tol"], iter_max=k["iter_max"])
    assert np.floor(x[0] * 10000.0) / 10000.0 == k["x0_floor_1e4"]
    assert abs(x[1] - 7.0 / 11.0) < 1e-15 and st["iterations"] == 2 and st["converged"]
    with pytest.raises(orc.OraclePanic, match="Matrix and vector size mismatch"):
        orc.cg(nr, nc, vals, cols, offs, np.array([1.0]), x)


def test_par_contract(orc, smb):
    k = KATS["par"]
    for row, blk, loc in k["locate"]:
        assert orc.par_locate(k["n_blocks"], k["max_n_rows"], row) == (blk, loc)
        assert smb.SparseMatPar.with_sub_matrices(k["n_blocks"], k["max_n_rows"]).get_block_and_row_id(row) == (blk, loc)
    # the reference clamps to n_blocks, not n_blocks - 1 (sparsemat_par.rs:32): row 16 maps to block 4
    assert orc.par_locate(4, 16, 16) == (4, 0)
    assert smb.SparseMatPar(4, 16).get_block_and_row_id(16) == (4, 0)
    with pytest.raises(orc.OraclePanic):                                 # R = 3 / 4 = 0 -> division by zero
        orc.par_locate(4, 3, 1)
    with pytest.raises(smb.Panic):
        smb.SparseMatPar(4, 3).get_block_and_row_id(1)


def test_threaded_mvp_equals_serial_bit_for_bit(orc):
    n_rows, n_cols, vals, cols, offs = cases.powerlaw(1, 30000, 30000, 3000, np.float32, np.uint32)
    x = orc.uniform(np.float32, 9, n_cols)
    assert np.array_equal(orc.mvp(vals, cols, offs, x, threads=4), orc.mvp(vals, cols, offs, x))


def test_par_path_as_shipped_equals_crs_mvp_bit_for_bit(orc):
    """SparseMatPar<SparseMatIndexList> assembled through set() + the serial default mvp through the block dispatch
    (sparsemat_par.rs:71-140, sparsematrix.rs:146-158) sums every row in insertion order: the same bits as to_crs + mvp."""
    for vdt, idt in [(np.float32, np.uint32), (np.float64, np.uint64)]:
        for nx, ny, nz, nb in [(20, 12, 9, 4), (33, 17, 1, 3), (5, 1, 1, 5), (16, 16, 16, 16)]:
            n = nx * ny * nz
            x = orc.uniform(vdt, 7, n)
            y, sec, asm = orc.par_laplace_mvp(vdt, idt, nb, nx, ny, nz, x)
            v, c, o = orc.laplace(vdt, idt, nx, ny, nz)
            assert np.array_equal(y, orc.mvp(v, c, o, x)) and sec > 0 and asm > 0
    with pytest.raises(orc.OraclePanic):                                   # R = max_n_rows / n_blocks == 0 (sparsemat_par.rs:21,32)
        orc.par_laplace_mvp(np.float32, np.uint32, 8, 2, 2, 1, orc.uniform(np.float32, 7, 4))


def test_committed_fixtures(orc):
    """Oracle outputs for seeded inputs, committed: guards the oracle itself against drift."""
    fx = np.load(os.path.join(HERE, "golden", "oracle_fixtures.npz"))
    import golden.make_fixtures as mk
    fresh = mk.compute(orc)
    assert set(fresh) == set(fx.files)
    for name in fx.files:
        assert fresh[name].dtype == fx[name].dtype and fresh[name].tobytes() == fx[name].tobytes(), name


def test_generators_closed_forms(orc):
    for nx, ny, nz in [(7, 5, 1), (6, 5, 4), (1, 1, 1)]:
        vals, cols, offs = orc.laplace(np.float64, np.uint32, nx, ny, nz)
        n = nx * ny * nz
        assert offs[-1] == vals.size == orc.lib().orc_laplace_nnz(nx, ny, nz)
        dense = np.zeros((n, n))
        for r in range(n):
            sl = slice(int(offs[r]), int(offs[r + 1]))
            assert np.all(np.diff(cols[sl].astype(np.int64)) > 0)        # ascending columns inside a row
            dense[r, cols[sl]] = vals[sl]
        assert np.array_equal(dense, dense.T) and np.all(np.diag(dense) == (6.0 if nz > 1 else 4.0))
        # partial row ranges are slices of the full operator
        lo, hi = n // 3, n - n // 4
        pv, pc, po = orc.laplace(np.float64, np.uint32, nx, ny, nz, lo, hi)
        assert np.array_equal(pv, vals[int(offs[lo]):int(offs[hi])]) and np.array_equal(pc, cols[int(offs[lo]):int(offs[hi])])
        assert np.array_equal(po.astype(np.int64), offs[lo:hi + 1].astype(np.int64) - int(offs[lo]))
    v, c, o = orc.powerlaw(np.float64, np.uint64, 5000, max_len=400)
    lens = np.diff(o.astype(np.int64))
    assert lens.min() >= 8 and lens.max() <= 400 and 10 < lens.mean() < 20 and c.max() < 5000
