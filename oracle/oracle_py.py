"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes access to oracle/liboracle.so (the C++17 CPU restatement
of the reference, see sparsemat_oracle.hpp).  May be imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs — never by the sparsemat_b200 package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
KAT = os.path.join(HERE, "kat")


def build(force: bool = False):
    if force or not (os.path.exists(LIB) and os.path.exists(KAT)):
        subprocess.run(["make", "-C", HERE, "-j4", "all"], check=True, capture_output=True)


def _load():
    build()
    return C.CDLL(LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
        _lib.orc_last_panic.restype = C.c_char_p
        _lib.orc_laplace_nnz.restype = C.c_uint64
        _lib.orc_laplace_nnz.argtypes = [C.c_uint64] * 3
        _lib.orc_powerlaw_row_len.restype = C.c_uint64
        _lib.orc_powerlaw_row_len.argtypes = [C.c_uint64] * 3
    return _lib


class OraclePanic(RuntimeError):
    pass


def _chk(rc):
    if rc != 0:
        raise OraclePanic(lib().orc_last_panic().decode())


def suffix(vdt, idt=None) -> str:
    v = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(vdt)]
    if idt is None:
        return v
    return v + {np.dtype(np.uint32): "u32", np.dtype(np.uint64): "u64"}[np.dtype(idt)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u64(x):
    return C.c_uint64(int(x))


def fn(name, sfx):
    return getattr(lib(), f"{name}_{sfx}")


# ---- generators -------------------------------------------------------------------------------------------
def laplace(vdt, idt, nx, ny, nz=1, row_lo=0, row_hi=None):
    """(values, columns, offsets) of rows [row_lo,row_hi) of the Dirichlet Laplacian; global columns."""
    n = nx * ny * nz
    if row_hi is None:
        row_hi = n
    if row_lo == 0 and row_hi == n:
        nnz = lib().orc_laplace_nnz(nx, ny, nz)
    else:
        nnz = 7 * (row_hi - row_lo)          # upper bound, trimmed below
    values = np.empty(nnz, vdt)
    columns = np.empty(nnz, idt)
    offsets = np.empty(row_hi - row_lo + 1, idt)
    fn("orc_laplace_rows", suffix(vdt, idt))(_u64(nx), _u64(ny), _u64(nz), _u64(row_lo), _u64(row_hi), _p(values),
                                             _p(columns), _p(offsets))
    real = int(offsets[-1])
    return values[:real], columns[:real], offsets


def uniform(vdt, seed, n):
    out = np.empty(n, vdt)
    fn("orc_uniform", suffix(vdt))(_u64(seed), _u64(n), _p(out))
    return out


def powerlaw(vdt, idt, n_rows, n_cols=None, seed_len=3, seed_col=4, seed_val=5, max_len=1_000_000):
    n_cols = n_cols or n_rows
    lens = np.empty(n_rows, np.uint64)
    lib().orc_powerlaw_row_lens(_u64(seed_len), _u64(n_rows), _u64(max_len), _p(lens))
    offsets = np.zeros(n_rows + 1, np.uint64)
    np.cumsum(lens, out=offsets[1:])
    nnz = int(offsets[-1])
    offsets = offsets.astype(idt)
    values = np.empty(nnz, vdt)
    columns = np.empty(nnz, idt)
    fn("orc_powerlaw_fill", suffix(vdt, idt))(_u64(seed_col), _u64(seed_val), _u64(n_rows), _u64(n_cols), _p(offsets),
                                              _p(values), _p(columns))
    return values, columns, offsets


def powerlaw_sample_mvp(vdt, n_cols, rows, x, seed_len=3, seed_col=4, seed_val=5, max_len=1_000_000):
    """The reference's mvp on a sample of rows of the implicit power-law matrix (regenerated from the seeds):
    returns (y[rows], sum |a||x| per row, row lengths)."""
    rows = np.ascontiguousarray(rows, np.uint64)
    y = np.empty(rows.size, vdt)
    ab = np.empty(rows.size, np.float64)
    lens = np.empty(rows.size, np.uint64)
    fn("orc_powerlaw_sample_mvp", suffix(vdt))(_u64(seed_len), _u64(seed_col), _u64(seed_val), _u64(n_cols), _u64(max_len),
                                               _u64(rows.size), _p(rows), _p(x), _p(y), _p(ab), _p(lens))
    return y, ab, lens


def powerlaw_row_lens(n_rows, seed_len=3, max_len=1_000_000):
    lens = np.empty(n_rows, np.uint64)
    lib().orc_powerlaw_row_lens(_u64(seed_len), _u64(n_rows), _u64(max_len), _p(lens))
    return lens


# ---- the hot path -------------------------------------------------------------------------------------------
def mvp(values, columns, offsets, x, threads: int = 1):
    """SparseMatrix::mvp over SparseMatCRS (sparsematrix.rs:146-158), raw arrays."""
    n_rows = offsets.size - 1
    y = np.empty(n_rows, values.dtype)
    sfx = suffix(values.dtype, columns.dtype)
    if threads > 1:
        fn("orc_mvp_threads", sfx)(_u64(n_rows), _p(values), _p(columns), _p(offsets), _p(x), _p(y), C.c_uint(threads))
    else:
        fn("orc_mvp", sfx)(_u64(n_rows), _p(values), _p(columns), _p(offsets), _p(x), _p(y))
    return y


def mvp_checked(n_rows, n_cols, values, columns, offsets, x):
    """Bounds-checked trait path (panics like the reference when x is too short)."""
    y = np.empty(n_rows, values.dtype)
    _chk(fn("orc_mvp_checked", suffix(values.dtype, columns.dtype))(
        _u64(n_rows), _u64(n_cols), _u64(values.size), _p(values), _p(columns), _p(offsets), _p(x), _u64(x.size), _p(y)))
    return y


def bilinear(n_rows, n_cols, values, columns, offsets, lhs, rhs):
    out = np.zeros(1, values.dtype)
    _chk(fn("orc_bilinear", suffix(values.dtype, columns.dtype))(
        _u64(n_rows), _u64(n_cols), _u64(values.size), _p(values), _p(columns), _p(offsets), _p(lhs), _u64(lhs.size),
        _p(rhs), _u64(rhs.size), _p(out)))
    return out[0]


def dot(x, y):
    f = fn("orc_dot", suffix(x.dtype))
    f.restype = C.c_float if x.dtype == np.float32 else C.c_double
    return x.dtype.type(f(_u64(min(x.size, y.size)), _p(x), _p(y)))


def norm2sq(x):
    f = fn("orc_norm2sq", suffix(x.dtype))
    f.restype = C.c_float if x.dtype == np.float32 else C.c_double
    return x.dtype.type(f(_u64(x.size), _p(x)))


def vec_add(x, y):
    _chk(fn("orc_add", suffix(x.dtype))(_u64(x.size), _p(x), _u64(y.size), _p(y)))


def vec_sub(x, y):
    _chk(fn("orc_sub", suffix(x.dtype))(_u64(x.size), _p(x), _u64(y.size), _p(y)))


def vec_scale(x, s):
    f = fn("orc_scale", suffix(x.dtype))
    f.argtypes = [C.c_uint64, C.c_void_p, C.c_float if x.dtype == np.float32 else C.c_double]
    f(x.size, _p(x), float(s))


def cg(n_rows, n_cols, values, columns, offsets, b, x, tol=1e-12, relative=False, iter_max=10_000, threads=1,
       history_cap=0):
    """ConjugateGradient::solve (linearsolver.rs:27-61) on raw CRS arrays; x is updated in place."""
    iters = C.c_uint64()
    res = C.c_double()
    conv = C.c_int()
    hist = np.zeros(max(1, history_cap), np.float64)
    _chk(fn("orc_cg", suffix(values.dtype, columns.dtype))(
        _u64(n_rows), _u64(n_cols), _p(values), _p(columns), _p(offsets), _p(b), _u64(b.size), _p(x), _u64(x.size),
        C.c_double(tol), C.c_int(int(relative)), _u64(iter_max), C.c_uint(threads), C.byref(iters), C.byref(res),
        C.byref(conv), _p(hist) if history_cap else None, _u64(history_cap)))
    return {"iterations": iters.value, "final_residual": res.value, "converged": bool(conv.value),
            "history": hist[:min(history_cap, iters.value)]}


# ---- assembly format ------------------------------------------------------------------------------------------
class IndexListMat:
    """SparseMatIndexList (sparsemat_indexlist.rs) restated; used to check to_crs bit for bit."""

    def __init__(self, vdt, idt):
        self.vdt, self.idt = np.dtype(vdt), np.dtype(idt)
        self.sfx = suffix(vdt, idt)
        f = fn("orc_il_new", self.sfx)
        f.restype = C.c_void_p
        self.h = C.c_void_p(f())

    def __del__(self):
        try:
            fn("orc_il_free", self.sfx)(self.h)
        except Exception:
            pass

    def apply(self, i, j, v, op):
        i = np.ascontiguousarray(np.atleast_1d(i), np.uint64)
        j = np.ascontiguousarray(np.atleast_1d(j), np.uint64)
        v = np.ascontiguousarray(np.atleast_1d(v), self.vdt)
        ops = np.full(i.size, op, np.uint8)
        _chk(fn("orc_il_apply", self.sfx)(self.h, _u64(i.size), _p(i), _p(j), _p(v), _p(ops)))

    def set(self, i, j, v):
        self.apply(i, j, v, 0)

    def add_to(self, i, j, v):
        self.apply(i, j, v, 1)

    def dims(self):
        d = np.zeros(3, np.uint64)
        fn("orc_il_dims", self.sfx)(self.h, _p(d))
        return int(d[0]), int(d[1]), int(d[2])

    def raw_arrays(self):
        r, _, z = self.dims()
        columns, values = np.empty(z, self.idt), np.empty(z, self.vdt)
        pos_start, nxt = np.empty(r, self.idt), np.empty(z, self.idt)
        fn("orc_il_export", self.sfx)(self.h, _p(columns), _p(values), _p(pos_start), _p(nxt))
        return columns, values, pos_start, nxt

    def transpose(self) -> "IndexListMat":
        """SparseMatrix::transpose (sparsematrix.rs:174-183) on the assembly format."""
        f = fn("orc_il_transpose", self.sfx)
        f.restype = C.c_void_p
        t = IndexListMat.__new__(IndexListMat)
        t.vdt, t.idt, t.sfx = self.vdt, self.idt, self.sfx
        t.h = C.c_void_p(f(self.h))
        return t

    def to_crs(self):
        """(n_rows, n_cols, values, columns, offset_rows) exactly as from_sparsemat_index lays them out."""
        f = fn("orc_il_to_crs", self.sfx)
        f.restype = C.c_void_p
        h = C.c_void_p(f(self.h))
        d = np.zeros(4, np.uint64)
        fn("orc_crs_dims", self.sfx)(h, _p(d))
        values, columns = np.empty(int(d[2]), self.vdt), np.empty(int(d[2]), self.idt)
        offsets = np.empty(int(d[3]), self.idt)
        fn("orc_crs_export", self.sfx)(h, _p(values), _p(columns), _p(offsets))
        fn("orc_crs_free", self.sfx)(h)
        return int(d[0]), int(d[1]), values, columns, offsets


def to_crs_raw(n_rows, columns, values, pos_start, nxt):
    nnz = columns.size
    ov, oc = np.empty(nnz, values.dtype), np.empty(nnz, columns.dtype)
    oo = np.empty(n_rows + 1, columns.dtype)
    _chk(fn("orc_to_crs_raw", suffix(values.dtype, columns.dtype))(
        _u64(n_rows), _u64(nnz), _p(columns), _p(values), _p(pos_start), _p(nxt), _p(ov), _p(oc), _p(oo)))
    return ov, oc, oo


def par_laplace_mvp(vdt, idt, n_blocks, nx, ny, nz, x, reps=1):
    """The sparsemat_par path as shipped (sparsemat_par.rs:71-140): the Laplacian assembled through SparseMatPar::set into
    IndexList blocks, then the SERIAL default mvp through the block dispatch.  Returns (y, seconds per product, seconds of
    assembly)."""
    n = nx * ny * nz
    x = np.ascontiguousarray(x, dtype=vdt)
    assert x.size >= n
    y = np.empty(n, vdt)
    sec, asm = C.c_double(), C.c_double()
    _chk(fn("orc_par_laplace_mvp", suffix(vdt, idt))(_u64(n_blocks), _u64(nx), _u64(ny), _u64(nz), _p(x), _p(y), C.c_uint(reps),
                                                   C.byref(sec), C.byref(asm)))
    return y, sec.value, asm.value


def par_locate(n_blocks, max_rows, row):
    out = np.zeros(2, np.uint64)
    _chk(lib().orc_par_locate(_u64(n_blocks), _u64(max_rows), _u64(row), _p(out)))
    return int(out[0]), int(out[1])


def run_kat() -> str:
    build()
    res = subprocess.run([KAT], capture_output=True, text=True)
    if res.returncode != 0:
        raise AssertionError("oracle KAT replay failed:\n" + res.stdout + res.stderr)
    return res.stdout
