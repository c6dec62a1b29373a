"""Python mirror of the reference's public interface for the hot path, over the C ABI.

Same names and argument meaning as lostinc0de/sparsemat (paths relative to /root/reference/src/):

    SparseMatIndexList   sparsemat_indexlist.rs   host assembly (set / add_to / get), ``to_crs()`` on the GPU
    SparseMatCRS         sparsemat_crs.rs          device-resident CRS; ``mvp`` = the CUDA SpMV
    DenseVec             densevec.rs, vector.rs    device-resident dense vector
    ConjugateGradient    linearsolver.rs           fused CG on the device
    SparseMatPar         sparsemat_par.rs          the row-block contract; ``DistCRS`` is its multi-GPU form

Everything compute runs through libsmb200; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi as F
from ._ffi import Panic, SmbError, check, lib  # noqa: F401


class Context:
    """One CUDA device + stream.  ``stream`` may be an existing ``cudaStream_t`` handle (int)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = C.c_void_p()
        check(lib.smb200_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib.smb200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib.smb200_ctx_sync(self._h))

    def flush_l2(self):
        check(lib.smb200_ctx_flush_l2(self._h))

    def devinfo(self) -> dict:
        d = F.DevInfo()
        check(lib.smb200_ctx_devinfo(self._h, C.byref(d)))
        return {"device": d.device, "sm_count": d.sm_count, "cc": (d.cc_major, d.cc_minor), "l2_bytes": d.l2_bytes,
                "l2_persist_max_bytes": d.l2_persist_max_bytes, "hbm_bytes": d.hbm_bytes, "name": d.name.decode()}

    def event(self) -> "Event":
        return Event(self)

    # multi-GPU (one process per GPU): NCCL communicator bootstrap
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib.smb200_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, uid: bytes | None):
        buf = C.create_string_buffer(uid, 128) if uid is not None else None
        check(lib.smb200_comm_init(self._h, rank, world, buf))
        self.rank, self.world = rank, world


class Event:
    def __init__(self, ctx: Context):
        h = C.c_void_p()
        check(lib.smb200_event_create(ctx._h, C.byref(h)))
        self._h = h

    def record(self):
        check(lib.smb200_event_record(self._h))
        return self

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        check(lib.smb200_event_elapsed_ms(self._h, stop._h, C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self._h:
                lib.smb200_event_destroy(self._h)
        except Exception:
            pass


def pinned_empty(n: int, dtype) -> np.ndarray:
    """Page-locked host array (for the end-to-end path)."""
    dt = np.dtype(dtype)
    p = C.c_void_p()
    check(lib.smb200_host_alloc(max(1, n * dt.itemsize), C.byref(p)))
    buf = (C.c_char * max(1, n * dt.itemsize)).from_address(p.value)
    buf._smb_pinned = _PinnedOwner(p)  # the array keeps `buf` alive, `buf` keeps the allocation alive
    return np.frombuffer(buf, dtype=dt, count=n)


class _PinnedOwner:
    def __init__(self, p):
        self.p = p

    def __del__(self):
        try:
            lib.smb200_host_free(self.p)
        except Exception:
            pass


# --------------------------------------------------------------------------------------------------------------
class DenseVec:
    """densevec.rs:5-140 / vector.rs:5-64 on the device."""

    def __init__(self, ctx: Context, n: int, dtype=np.float64, _handle=None):
        self.ctx = ctx
        self.dtype = np.dtype(dtype)
        if _handle is None:
            h = C.c_void_p()
            check(lib.smb200_vec_create(ctx._h, F.vtype_of(dtype), n, C.byref(h)))
            _handle = h
        self._h = _handle

    @classmethod
    def from_vec(cls, ctx: Context, values) -> "DenseVec":            # densevec.rs:30-34
        a = np.ascontiguousarray(values)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        v = cls(ctx, a.size, a.dtype)
        check(lib.smb200_vec_upload(v._h, F.ptr(a), a.size))
        ctx.sync()
        return v

    @classmethod
    def wrap(cls, ctx: Context, device_ptr: int, n: int, dtype) -> "DenseVec":
        h = C.c_void_p()
        check(lib.smb200_vec_wrap(ctx._h, F.vtype_of(dtype), n, C.c_void_p(device_ptr), C.byref(h)))
        return cls(ctx, n, dtype, _handle=h)

    def __del__(self):
        try:
            if self._h:
                lib.smb200_vec_free(self._h)
                self._h = None
        except Exception:
            pass

    def dim(self) -> int:                                              # densevec.rs:36-38
        n = C.c_uint64()
        check(lib.smb200_vec_dim(self._h, C.byref(n)))
        return n.value

    def device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib.smb200_vec_device_ptr(self._h, C.byref(p)))
        return p.value or 0

    def to_numpy(self) -> np.ndarray:                                  # iter_ref().as_slice()
        out = np.empty(self.dim(), self.dtype)
        check(lib.smb200_vec_download(self._h, F.ptr(out), out.size))
        return out

    def upload(self, values: np.ndarray):
        a = np.ascontiguousarray(values, self.dtype)
        check(lib.smb200_vec_upload(self._h, F.ptr(a), a.size))
        self.ctx.sync()

    def get(self, i: int):                                             # densevec.rs:40-42 (bounds-checked)
        if i < 0 or i >= self.dim():
            raise Panic("index out of bounds")
        return self.to_numpy()[i]

    def clone(self) -> "DenseVec":
        h = C.c_void_p()
        check(lib.smb200_vec_clone(self._h, C.byref(h)))
        return DenseVec(self.ctx, 0, self.dtype, _handle=h)

    def fill(self, value: float):
        check(lib.smb200_vec_fill(self._h, float(value)))

    def fill_uniform(self, seed: int):
        check(lib.smb200_vec_fill_uniform(self._h, seed))

    def add(self, rhs: "DenseVec"):                                    # densevec.rs:51-58
        check(lib.smb200_vec_add(self._h, rhs._h))

    def sub(self, rhs: "DenseVec"):                                    # densevec.rs:60-67
        check(lib.smb200_vec_sub(self._h, rhs._h))

    def scale(self, s: float):                                         # densevec.rs:69-73
        check(lib.smb200_vec_scale(self._h, float(s)))

    def axpy(self, alpha: float, x: "DenseVec"):                       # *self += x.clone() * alpha
        check(lib.smb200_vec_axpy(self._h, float(alpha), x._h))

    def scale_add(self, beta: float, r: "DenseVec"):                   # self.scale(beta); self.add(&r)
        check(lib.smb200_vec_scale_add(self._h, float(beta), r._h))

    def inner_prod(self, rhs: "DenseVec"):                             # vector.rs:50-53
        out = C.c_double()
        check(lib.smb200_vec_dot(self._h, rhs._h, C.byref(out)))
        return self.dtype.type(out.value)

    def norm_squared(self):                                            # vector.rs:56-58
        out = C.c_double()
        check(lib.smb200_vec_norm2sq(self._h, C.byref(out)))
        return self.dtype.type(out.value)

    def norm(self) -> float:                                           # vector.rs:61-63 (always f64)
        out = C.c_double()
        check(lib.smb200_vec_norm(self._h, C.byref(out)))
        return out.value

    # operators, densevec.rs:76-140
    def __iadd__(self, rhs):
        self.add(rhs)
        return self

    def __isub__(self, rhs):
        self.sub(rhs)
        return self

    def __imul__(self, s):
        self.scale(s)
        return self

    def __add__(self, rhs):
        r = self.clone()
        r.add(rhs)
        return r

    def __sub__(self, rhs):
        r = self.clone()
        r.sub(rhs)
        return r

    def __mul__(self, rhs):
        if isinstance(rhs, DenseVec):
            return self.inner_prod(rhs)
        r = self.clone()
        r.scale(rhs)
        return r


# --------------------------------------------------------------------------------------------------------------
def _check_crs_lengths(n_rows, values, columns, offset_rows) -> None:
    """The reference derives these lengths from its Vecs (sparsemat_crs.rs:9-17) and cannot get them wrong; raw arrays can,
    and a short array would be read past its end by the upload."""
    if columns.size != values.size:
        raise ValueError(f"columns has {columns.size} entries, values has {values.size}")
    want = n_rows + 1 if (n_rows or values.size) else offset_rows.size
    if offset_rows.size != want:
        raise ValueError(f"offset_rows has {offset_rows.size} entries, expected n_rows + 1 = {want}")


def crsfile_write(path, n_rows, n_cols, values, columns, offset_rows) -> None:
    """Host only: write CRS arrays (sparsemat_crs.rs:9-17 layout) as a binary container.  No device needed."""
    values = np.ascontiguousarray(values)
    columns = np.ascontiguousarray(columns)
    offset_rows = np.ascontiguousarray(offset_rows, columns.dtype)
    _check_crs_lengths(n_rows, values, columns, offset_rows)
    check(lib.smb200_crsfile_write(os.fsencode(path), F.vtype_of(values.dtype), F.itype_of(columns.dtype), n_rows, n_cols,
                                   values.size, F.ptr(values) if values.size else None, F.ptr(columns) if columns.size else None,
                                   F.ptr(offset_rows) if n_rows else None))


def crsfile_read(path):
    """Host only: (n_rows, n_cols, values, columns, offset_rows) of a binary CRS container; raises SmbError(ERR_IO) for
    foreign, truncated or damaged files."""
    vt, it = C.c_int32(), C.c_int32()
    d = (C.c_uint64 * 3)()
    check(lib.smb200_crsfile_info(os.fsencode(path), C.byref(vt), C.byref(it), d))
    values = np.empty(d[2], F.VDTYPES[vt.value])
    columns = np.empty(d[2], F.IDTYPES[it.value])
    offsets = np.empty(d[0] + 1 if d[0] else 0, F.IDTYPES[it.value])
    check(lib.smb200_crsfile_read(os.fsencode(path), F.ptr(values) if values.size else None, values.nbytes,
                                  F.ptr(columns) if columns.size else None, columns.nbytes,
                                  F.ptr(offsets) if offsets.size else None, offsets.nbytes))
    return int(d[0]), int(d[1]), values, columns, offsets


class SparseMatCRS:
    """sparsemat_crs.rs:9-17 on the device; ``mvp`` is sparsematrix.rs:146-158."""

    def __init__(self, ctx: Context, handle, borrowed: bool = False):
        self.ctx = ctx
        self._h = handle
        self._borrowed = borrowed
        vt, it = C.c_int32(), C.c_int32()
        check(lib.smb200_crs_types(self._h, C.byref(vt), C.byref(it)))
        self.dtype = F.VDTYPES[vt.value]
        self.itype = F.IDTYPES[it.value]

    def __del__(self):
        try:
            if self._h and not self._borrowed:
                lib.smb200_crs_free(self._h)
            self._h = None
        except Exception:
            pass

    @classmethod
    def from_raw_parts(cls, ctx, n_rows, n_cols, values, columns, offset_rows) -> "SparseMatCRS":
        values = np.ascontiguousarray(values)
        columns = np.ascontiguousarray(columns)
        offset_rows = np.ascontiguousarray(offset_rows, columns.dtype)
        _check_crs_lengths(n_rows, values, columns, offset_rows)
        h = C.c_void_p()
        check(lib.smb200_crs_upload(ctx._h, F.vtype_of(values.dtype), F.itype_of(columns.dtype), n_rows, n_cols,
                                    values.size, F.ptr(values), F.ptr(columns), F.ptr(offset_rows), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def laplace(cls, ctx, dtype, itype, nx, ny, nz=1, row_lo=0, row_hi=None) -> "SparseMatCRS":
        """Dirichlet Laplacian generated on the device (BASELINE.json configs C1/C2/C4/C5)."""
        if row_hi is None:
            row_hi = nx * ny * nz
        h = C.c_void_p()
        check(lib.smb200_gen_laplace(ctx._h, F.vtype_of(dtype), F.itype_of(itype), nx, ny, nz, row_lo, row_hi, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def powerlaw(cls, ctx, dtype, itype, n_rows, n_cols=None, seed_len=3, seed_col=4, seed_val=5, max_len=1_000_000):
        """Power-law row lengths (BASELINE.json config C3)."""
        h = C.c_void_p()
        check(lib.smb200_gen_powerlaw(ctx._h, F.vtype_of(dtype), F.itype_of(itype), n_rows, n_cols or n_rows, seed_len,
                                      seed_col, seed_val, max_len, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def load(cls, ctx, path) -> "SparseMatCRS":
        """Read a binary CRS container written by ``save`` / ``crsfile_write`` and upload it (validated like any upload)."""
        h = C.c_void_p()
        check(lib.smb200_crs_load(ctx._h, os.fsencode(path), C.byref(h)))
        return cls(ctx, h)

    def save(self, path) -> None:
        """Download the three CRS arrays and write them, byte for byte, behind a checksummed header (include/smb200.h)."""
        check(lib.smb200_crs_save(self._h, os.fsencode(path)))

    def _dims(self):
        d = (C.c_uint64 * 3)()
        check(lib.smb200_crs_dims(self._h, d))
        return d[0], d[1], d[2]

    def n_rows(self) -> int:
        return self._dims()[0]

    def n_cols(self) -> int:
        return self._dims()[1]

    def n_non_zero_entries(self) -> int:
        return self._dims()[2]

    def empty(self) -> bool:
        return self.n_rows() == 0

    def density(self) -> float:                                        # sparsematrix.rs:237-241
        r, c, z = self._dims()
        return float(z) / float(r * c)

    def raw_parts(self):
        """(values, columns, offset_rows) copied back from the device — bit-exact layout checks."""
        r, _, z = self._dims()
        values = np.empty(z, self.dtype)
        columns = np.empty(z, self.itype)
        offsets = np.empty(r + 1 if (r or z) else 0, self.itype)
        check(lib.smb200_crs_download(self._h, F.ptr(values), F.ptr(columns), F.ptr(offsets) if offsets.size else None))
        return values, columns, offsets

    def iter_row(self, row: int):                                      # sparsemat_crs.rs:102-110
        if row >= self.n_rows():
            return []
        v, c, o = self.raw_parts()
        return [(c[k], v[k]) for k in range(int(o[row]), int(o[row + 1]))]

    def configure(self, variant=F.SPMV_AUTO, lanes=0, flags=0):
        check(lib.smb200_crs_configure(self._h, variant, lanes, flags))
        return self

    def plan_info(self) -> dict:
        p = F.PlanInfo()
        check(lib.smb200_crs_plan_info(self._h, C.byref(p)))
        d = {k: getattr(p, k) for k, _ in F.PlanInfo._fields_}
        d["variant_name"] = F.VARIANT_NAMES.get(p.variant, "?")
        return d

    def diagonal(self) -> DenseVec:
        """d[i] = get(i, i) (first stored entry of row i with column i, else 0) as a device vector."""
        d = DenseVec(self.ctx, self.n_rows(), self.dtype)
        check(lib.smb200_crs_diagonal(self._h, d._h))
        return d

    def scale(self, s: float):                                         # sparsemat_crs.rs:153-157
        check(lib.smb200_crs_scale(self._h, float(s)))

    def mvp(self, rhs: DenseVec, out: DenseVec | None = None) -> DenseVec:
        """y = A x.  Like the reference it returns a fresh vector of dim n_rows unless ``out`` is given."""
        y = out if out is not None else DenseVec(self.ctx, self.n_rows(), self.dtype)
        check(lib.smb200_spmv(self._h, rhs._h, y._h))
        return y

    def mvp_host(self, x: np.ndarray, y: np.ndarray | None = None) -> np.ndarray:
        """Same product through host buffers (H2D x, SpMV, D2H y) — the end-to-end call."""
        x = np.ascontiguousarray(x, self.dtype)
        if y is None:
            y = np.empty(self.n_rows(), self.dtype)
        check(lib.smb200_spmv_host(self._h, F.ptr(x), x.size, F.ptr(y)))
        return y

    def transpose(self) -> "SparseMatCRS":                             # sparsematrix.rs:174-183
        h = C.c_void_p()
        check(lib.smb200_crs_transpose(self._h, C.byref(h)))
        return SparseMatCRS(self.ctx, h)

    def inner_prod(self, lhs: DenseVec, rhs: DenseVec):                # sparsematrix.rs:161-171
        out = C.c_double()
        check(lib.smb200_bilinear(self._h, lhs._h, rhs._h, C.byref(out)))
        return self.dtype.type(out.value)

    def __mul__(self, rhs):                                            # sparsematrix.rs:435-443
        if isinstance(rhs, DenseVec):
            return self.mvp(rhs)
        return NotImplemented

    __matmul__ = __mul__


class SparseMatIndexList:
    """sparsemat_indexlist.rs:14-207: host-side assembly, converted on the device by ``to_crs``."""

    def __init__(self, dtype=np.float64, itype=np.uint32):
        self.dtype = np.dtype(dtype)
        self.itype = np.dtype(itype)
        h = C.c_void_p()
        check(lib.smb200_il_create(F.vtype_of(dtype), F.itype_of(itype), C.byref(h)))
        self._h = h

    @classmethod
    def new(cls, dtype=np.float64, itype=np.uint32):
        return cls(dtype, itype)

    with_capacity = new

    def __del__(self):
        try:
            if self._h:
                lib.smb200_il_free(self._h)
                self._h = None
        except Exception:
            pass

    def _apply(self, i, j, v, op):
        i = np.ascontiguousarray(np.atleast_1d(i), np.uint64)
        j = np.ascontiguousarray(np.atleast_1d(j), np.uint64)
        v = np.ascontiguousarray(np.atleast_1d(v), self.dtype)
        assert i.size == j.size == v.size
        check(lib.smb200_il_apply(self._h, i.size, F.ptr(i), F.ptr(j), F.ptr(v), op))

    def set(self, i, j, val):                                          # sparsematrix.rs:226-228
        self._apply(i, j, val, 0)

    def add_to(self, i, j, val):                                       # sparsematrix.rs:231-233
        self._apply(i, j, val, 1)

    set_many = set
    add_to_many = add_to

    def get(self, i: int, j: int):
        out = C.c_double()
        check(lib.smb200_il_get(self._h, i, j, C.byref(out)))
        return self.dtype.type(out.value)

    def _dims(self):
        d = (C.c_uint64 * 3)()
        check(lib.smb200_il_dims(self._h, d))
        return d[0], d[1], d[2]

    def n_rows(self):
        return self._dims()[0]

    def n_cols(self):
        return self._dims()[1]

    def n_non_zero_entries(self):
        return self._dims()[2]

    def density(self):
        r, c, z = self._dims()
        return float(z) / float(r * c)

    def raw_arrays(self):
        """(columns, values, pos_start, index_list) — indexlist.rs:26-29 layout, I::MAX = UNSET."""
        r, _, z = self._dims()
        columns = np.empty(z, self.itype)
        values = np.empty(z, self.dtype)
        pos_start = np.empty(r, self.itype)
        index_list = np.empty(z, self.itype)
        check(lib.smb200_il_export(self._h, F.ptr(columns), F.ptr(values), F.ptr(pos_start), F.ptr(index_list)))
        return columns, values, pos_start, index_list

    def to_crs(self, ctx: Context) -> SparseMatCRS:                    # sparsemat_indexlist.rs:61-63
        h = C.c_void_p()
        check(lib.smb200_il_to_crs(self._h, ctx._h, C.byref(h)))
        return SparseMatCRS(ctx, h)


def crs_from_indexlist_arrays(ctx, n_rows, n_cols, columns, values, pos_start, index_list) -> SparseMatCRS:
    """smb200_crs_from_indexlist on caller-provided IndexList arrays (what the Rust overlay passes)."""
    columns = np.ascontiguousarray(columns)
    values = np.ascontiguousarray(values)
    pos_start = np.ascontiguousarray(pos_start, columns.dtype)
    index_list = np.ascontiguousarray(index_list, columns.dtype)
    h = C.c_void_p()
    check(lib.smb200_crs_from_indexlist(ctx._h, F.vtype_of(values.dtype), F.itype_of(columns.dtype), n_rows, n_cols,
                                        columns.size, F.ptr(columns), F.ptr(values), F.ptr(pos_start), F.ptr(index_list),
                                        C.byref(h)))
    return SparseMatCRS(ctx, h)


class ConjugateGradient:
    """linearsolver.rs:12-61.  ``ConjugateGradient()`` is ``Default``: tol 1e-12 (absolute), 10 000 iterations.
    ``single_reduce`` (additive, distributed matrices only): the loop rearranged so that an iteration has one all-reduce."""

    def __init__(self, tol: float = 1e-12, iter_max: int = 10_000, relative: bool = False, single_reduce: bool = False):
        self.tol, self.iter_max, self.relative, self.single_reduce = tol, iter_max, relative, single_reduce
        self.last_stats = None

    @classmethod
    def default(cls):
        return cls()

    def solve(self, mat, b: DenseVec, x: DenseVec) -> None:
        self.solve_with_stats(mat, b, x)

    def solve_with_stats(self, mat, b: DenseVec, x: DenseVec) -> dict:
        st = F.CgStats()
        if self.single_reduce and not isinstance(mat, DistCRS):
            raise ValueError("single_reduce is the distributed solver's option (DistCRS)")
        if isinstance(mat, DistCRS):
            fn = lib.smb200_dist_cg_solve_sr if self.single_reduce else lib.smb200_dist_cg_solve
            check(fn(mat._h, b._h, x._h, self.tol, int(self.relative), self.iter_max, C.byref(st)))
        else:
            check(lib.smb200_cg_solve(mat._h, b._h, x._h, self.tol, int(self.relative), self.iter_max, C.byref(st)))
        self.last_stats = {"iterations": st.iterations, "final_residual": st.final_residual,
                           "converged": bool(st.converged), "device_ms": st.device_ms, "launches": st.launches}
        return self.last_stats

    @staticmethod
    def history(mat: SparseMatCRS) -> np.ndarray:
        n = C.c_uint64()
        check(lib.smb200_cg_history(mat._h, None, 0, C.byref(n)))
        out = np.empty(n.value, np.float64)
        if n.value:
            check(lib.smb200_cg_history(mat._h, out.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n)))
        return out


class JacobiPCG:
    """Additive (the reference has no preconditioner): CG preconditioned with the inverse diagonal.  Same constructor,
    checks and panics as ``ConjugateGradient``."""

    def __init__(self, tol: float = 1e-12, iter_max: int = 10_000, relative: bool = False):
        self.tol, self.iter_max, self.relative = tol, iter_max, relative
        self.last_stats = None

    def solve(self, mat: SparseMatCRS, b: DenseVec, x: DenseVec) -> None:
        self.solve_with_stats(mat, b, x)

    def solve_with_stats(self, mat: SparseMatCRS, b: DenseVec, x: DenseVec) -> dict:
        st = F.CgStats()
        check(lib.smb200_pcg_jacobi_solve(mat._h, b._h, x._h, self.tol, int(self.relative), self.iter_max, C.byref(st)))
        self.last_stats = {"iterations": st.iterations, "final_residual": st.final_residual,
                           "converged": bool(st.converged), "device_ms": st.device_ms, "launches": st.launches}
        return self.last_stats


# ---- partition contract + multi-GPU ---------------------------------------------------------------------------
class _ParHandle:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        try:
            if self.h:
                lib.smb200_par_free(self.h)
                self.h = None
        except Exception:
            pass


class SparseMatPar:
    """sparsemat_par.rs:12-140: 1-D row blocks of R = max_n_rows / n_blocks rows, each a SparseMatIndexList with LOCAL
    row ids and GLOBAL column ids, assembled on the host like the reference.  ``mvp`` is the reference's commented-out
    ``mvp_par`` (sparsemat_par.rs:37-68) completed on the GPU: every block is converted by ``to_crs`` on the device and
    multiplied against the shared x; block b writes y[b*R ..).  The reference's quirks are kept: the block id is clamped
    to n_blocks (not n_blocks - 1), so rows >= n_blocks*R panic, and R == 0 divides by zero."""

    def __init__(self, n_blocks: int, max_n_rows: int, dtype=np.float64, itype=np.uint32):
        if n_blocks == 0:
            raise Panic("attempt to divide by zero")                      # sparsemat_par.rs:21
        self.n_blocks, self.max_n_rows = n_blocks, max_n_rows
        self.n_rows_sub_matrix = max_n_rows // n_blocks
        self.dtype, self.itype = np.dtype(dtype), np.dtype(itype)
        self.sub_matrices = [SparseMatIndexList(dtype, itype) for _ in range(n_blocks)]
        self._device = None                                               # (ctx, [SparseMatCRS per block]) after to_device

    @classmethod
    def with_sub_matrices(cls, n_blocks, max_n_rows, dtype=np.float64, itype=np.uint32):
        return cls(n_blocks, max_n_rows, dtype, itype)

    @classmethod
    def with_capacity(cls, cap, dtype=np.float64, itype=np.uint32):      # sparsemat_par.rs:91-93
        return cls(4, cap, dtype, itype)

    def get_block_and_row_id(self, row: int):                             # sparsemat_par.rs:31-35
        b, r = C.c_uint64(), C.c_uint64()
        st = lib.smb200_par_locate(self.n_blocks, self.max_n_rows, row, C.byref(b), C.byref(r))
        if st != F.OK:
            raise Panic("attempt to divide by zero")
        return b.value, r.value

    def _block(self, block_id: int) -> SparseMatIndexList:
        if block_id >= len(self.sub_matrices):
            raise Panic("index out of bounds")                            # Vec indexing with the clamped block id
        return self.sub_matrices[block_id]

    def _apply(self, i, j, v, op):
        self._device = None
        i = np.atleast_1d(np.asarray(i, np.uint64))
        j = np.atleast_1d(np.asarray(j, np.uint64))
        v = np.atleast_1d(np.asarray(v, self.dtype))
        if self.n_rows_sub_matrix == 0:
            raise Panic("attempt to divide by zero")
        blocks = np.minimum(i // np.uint64(self.n_rows_sub_matrix), np.uint64(self.n_blocks))
        if np.any(blocks >= self.n_blocks):
            raise Panic("index out of bounds")
        # keep the caller's order inside every block (insertion order is what to_crs freezes)
        for b in np.unique(blocks):
            m = blocks == b
            self.sub_matrices[int(b)]._apply(i[m] - b * np.uint64(self.n_rows_sub_matrix), j[m], v[m], op)

    def set(self, i, j, val):
        self._apply(i, j, val, 0)

    def add_to(self, i, j, val):
        self._apply(i, j, val, 1)

    def get(self, i: int, j: int):
        b, r = self.get_block_and_row_id(i)
        return self._block(b).get(r, j)

    def n_rows(self) -> int:                                              # sparsemat_par.rs:95-107
        last = 0
        for b, m in enumerate(self.sub_matrices):
            if m.n_rows() == 0:
                break
            last = b
        return last * self.n_rows_sub_matrix + self.sub_matrices[last].n_rows()

    def n_cols(self) -> int:
        return max(m.n_cols() for m in self.sub_matrices)

    def n_non_zero_entries(self) -> int:
        return sum(m.n_non_zero_entries() for m in self.sub_matrices)

    def density(self) -> float:                                           # sparsematrix.rs:237-241
        return float(self.n_non_zero_entries()) / float(self.n_rows() * self.n_cols())

    def to_device(self, ctx: Context):
        """The device container (smb200_par_*): to_crs() of every block on its owner's GPU, bit-exact layout per block.  With a
        communicator on the context, block b lives on rank b * world / n_blocks and every rank makes this same call."""
        h = C.c_void_p()
        check(lib.smb200_par_create(ctx._h, self.n_blocks, self.max_n_rows, F.vtype_of(self.dtype), F.itype_of(self.itype), C.byref(h)))
        dev = _ParHandle(h)
        for b, m in enumerate(self.sub_matrices):
            r, c, z = m._dims()
            cols, vals, pos, nxt = m.raw_arrays()
            check(lib.smb200_par_set_block_indexlist(h, b, r, c, z, F.ptr(cols), F.ptr(vals), F.ptr(pos), F.ptr(nxt)))
        self._device = (ctx, dev)
        return self

    def owner(self, block_id: int) -> int:
        """Rank that holds block `block_id` on the device (after to_device)."""
        r = C.c_int32()
        check(lib.smb200_par_owner(self._device[1].h, block_id, C.byref(r)))
        return r.value

    def mvp(self, rhs: DenseVec) -> DenseVec:
        """y = A x: the completed mvp_par (sparsemat_par.rs:37-68) — every device multiplies its blocks against the shared x
        (the `Arc<rhs>`), the slices of y are gathered on every rank."""
        ctx = rhs.ctx
        if self._device is None or self._device[0] is not ctx:
            self.to_device(ctx)
        y = DenseVec(ctx, self.n_rows(), self.dtype)
        st = lib.smb200_par_mvp(self._device[1].h, rhs._h, y._h)
        if st == F.ERR_INVALID and "index out of bounds" in F.last_error():
            # the default mvp would call IndexList::iter_row past a short block's last row (indexlist.rs:88)
            raise Panic("index out of bounds")
        check(st)
        return y

    def __mul__(self, rhs):
        if isinstance(rhs, DenseVec):
            return self.mvp(rhs)
        return NotImplemented


def partition_rows(n_rows: int, world: int, align: int = 1) -> np.ndarray:
    out = np.empty(world + 1, np.uint64)
    check(lib.smb200_partition_rows(n_rows, world, align, out.ctypes.data_as(C.POINTER(C.c_uint64))))
    return out


def partition_rows_by_nnz(offset_rows: np.ndarray, world: int) -> np.ndarray:
    offset_rows = np.ascontiguousarray(offset_rows)
    out = np.empty(world + 1, np.uint64)
    check(lib.smb200_partition_rows_by_nnz(F.itype_of(offset_rows.dtype), offset_rows.size - 1, F.ptr(offset_rows), world,
                                           out.ctypes.data_as(C.POINTER(C.c_uint64))))
    return out


def ghost_plan(columns_global: np.ndarray, world: int, rank: int, bounds: np.ndarray):
    """Host-side ghost analysis of one rank: (columns_local, ghosts, ghosts_per_owner)."""
    cols = np.ascontiguousarray(columns_global)
    bounds = np.ascontiguousarray(bounds, np.uint64)
    n = C.c_uint64()
    per = np.zeros(world, np.uint64)
    u64p = C.POINTER(C.c_uint64)
    check(lib.smb200_ghost_plan(F.itype_of(cols.dtype), cols.size, F.ptr(cols), world, rank, bounds.ctypes.data_as(u64p),
                                None, None, C.byref(n), per.ctypes.data_as(u64p)))
    ghosts = np.empty(n.value, np.uint64)
    local = np.empty_like(cols)
    check(lib.smb200_ghost_plan(F.itype_of(cols.dtype), cols.size, F.ptr(cols), world, rank, bounds.ctypes.data_as(u64p),
                                F.ptr(local), ghosts.ctypes.data_as(u64p), C.byref(n), per.ctypes.data_as(u64p)))
    return local, ghosts, per


class DistCRS:
    """Row-block partitioned matrix, one rank per GPU."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle
        lh = C.c_void_p()
        check(lib.smb200_dist_local(self._h, C.byref(lh)))
        self.local = SparseMatCRS(ctx, lh, borrowed=True)
        self.dtype = self.local.dtype

    @classmethod
    def laplace(cls, ctx, dtype, itype, nx, ny, nz) -> "DistCRS":
        h = C.c_void_p()
        check(lib.smb200_dist_laplace(ctx._h, F.vtype_of(dtype), F.itype_of(itype), nx, ny, nz, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_local_block(cls, ctx, n_global, bounds, values, columns_global, offset_rows_local) -> "DistCRS":
        bounds = np.ascontiguousarray(bounds, np.uint64)
        values = np.ascontiguousarray(values)
        cols = np.ascontiguousarray(columns_global)
        offs = np.ascontiguousarray(offset_rows_local, cols.dtype)
        h = C.c_void_p()
        check(lib.smb200_dist_create(ctx._h, F.vtype_of(values.dtype), F.itype_of(cols.dtype), n_global,
                                     bounds.ctypes.data_as(C.POINTER(C.c_uint64)), values.size, F.ptr(values), F.ptr(cols),
                                     F.ptr(offs), C.byref(h)))
        return cls(ctx, h)

    def __del__(self):
        try:
            if self._h:
                self.local._h = None
                lib.smb200_dist_free(self._h)
                self._h = None
        except Exception:
            pass

    def dims(self) -> dict:
        d = (C.c_uint64 * 4)()
        check(lib.smb200_dist_dims(self._h, d))
        return {"n_local": d[0], "n_ghost": d[1], "nnz_local": d[2], "row_lo": d[3]}

    def new_vec(self) -> DenseVec:
        h = C.c_void_p()
        check(lib.smb200_dist_vec_create(self._h, C.byref(h)))
        return DenseVec(self.ctx, 0, self.dtype, _handle=h)

    def mvp(self, x: DenseVec, out: DenseVec | None = None) -> DenseVec:
        y = out if out is not None else self.new_vec()
        check(lib.smb200_dist_spmv(self._h, x._h, y._h))
        return y

    def dot(self, x: DenseVec, y: DenseVec) -> float:
        out = C.c_double()
        check(lib.smb200_dist_dot(self._h, x._h, y._h, C.byref(out)))
        return out.value

    def barrier(self) -> None:
        """Stream-ordered barrier over the ranks (the host does not wait)."""
        check(lib.smb200_dist_barrier(self._h))

    def info(self) -> dict:
        d = (C.c_uint64 * 6)()
        check(lib.smb200_dist_info(self._h, d))
        return {"p2p": bool(d[0]), "neighbours": int(d[1]), "products": int(d[2]), "peer_timeout": bool(d[3]),
                "wait_ns_total": int(d[4]), "waits": int(d[5])}
