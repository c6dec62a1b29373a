"""ctypes binding of libsmb200 (include/smb200.h, include/smb200_host.h).

The shared library is built in-tree by ``sparsemat_b200.build`` (nvcc, sm_100a) and lives next to this
file in ``lib/``.  There is no CPU fallback: if the library is missing the import fails loudly, and if
no CUDA device is usable every compute call raises ``SmbError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMB200_LIB") or os.path.join(_HERE, "lib", "libsmb200.so")   # SMB200_LIB: A/B builds of the same ABI

OK, ERR_INVALID, ERR_DIM, ERR_CUDA, ERR_NOT_SQUARE, ERR_SIZE_MISMATCH, ERR_NCCL, ERR_UNSUPPORTED, ERR_OOM, ERR_IO = range(10)
F32, F64 = 0, 1
U32, U64 = 0, 1
SPMV_AUTO, SPMV_SCALAR, SPMV_VECTOR, SPMV_STREAM, SPMV_STREAM_TMA, SPMV_BANDED, SPMV_STREAM_PIPE, SPMV_RING, SPMV_BANDSPLIT = range(9)
FLAG_L2_PERSIST_X = 1
VARIANT_NAMES = {0: "auto", 1: "scalar", 2: "vector", 3: "stream", 4: "stream_tma", 5: "banded", 6: "stream_pipe", 7: "ring", 8: "bandsplit"}


class Panic(RuntimeError):
    """The reference would have panicked here; the message is the reference's own."""


class SmbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"smb200 status {status}: {msg}")
        self.status = status


class DevInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("l2_bytes", C.c_int64), ("l2_persist_max_bytes", C.c_int64), ("hbm_bytes", C.c_int64),
                ("name", C.c_char * 128)]


class PlanInfo(C.Structure):
    _fields_ = [("variant", C.c_int32), ("lanes", C.c_int32), ("flags", C.c_uint32), ("n_blocks", C.c_uint64),
                ("n_rows", C.c_uint64), ("n_cols", C.c_uint64), ("nnz", C.c_uint64), ("max_row_len", C.c_uint64),
                ("mean_row_len", C.c_double), ("algorithmic_bytes", C.c_uint64), ("launches_per_spmv", C.c_uint64),
                ("n_xwin_blocks", C.c_uint64), ("nnz_c16", C.c_uint64), ("stream_bytes", C.c_uint64),
                ("rows_o16", C.c_uint64), ("plan_bytes", C.c_uint64), ("plan_ms", C.c_double), ("nnz_v8", C.c_uint64), ("sell_entries", C.c_uint64)]


class CgStats(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("final_residual", C.c_double), ("converged", C.c_int32),
                ("device_ms", C.c_float), ("launches", C.c_uint64)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C sparsemat_b200/csrc` or `python __graft_entry__.py` (nvcc, sm_100a). "
            "sparsemat_b200 has no CPU fallback.")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

_p = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_u64 = C.c_uint64
_u64p = C.POINTER(C.c_uint64)
_dp = C.POINTER(C.c_double)
_i32 = C.c_int32

# name -> (restype, argtypes); every symbol of the two headers is listed (tests check the export table)
PROTOTYPES = {
    "smb200_version": (_i32, []),
    "smb200_last_error": (C.c_char_p, []),
    "smb200_launch_count": (_u64, []),
    "smb200_ctx_create": (_i32, [_i32, _p, _pp]),
    "smb200_ctx_destroy": (_i32, [_p]),
    "smb200_ctx_sync": (_i32, [_p]),
    "smb200_ctx_devinfo": (_i32, [_p, C.POINTER(DevInfo)]),
    "smb200_ctx_flush_l2": (_i32, [_p]),
    "smb200_event_create": (_i32, [_p, _pp]),
    "smb200_event_record": (_i32, [_p]),
    "smb200_event_elapsed_ms": (_i32, [_p, _p, C.POINTER(C.c_float)]),
    "smb200_event_destroy": (_i32, [_p]),
    "smb200_host_alloc": (_i32, [C.c_size_t, _pp]),
    "smb200_host_free": (_i32, [_p]),
    "smb200_vec_create": (_i32, [_p, _i32, _u64, _pp]),
    "smb200_vec_wrap": (_i32, [_p, _i32, _u64, _p, _pp]),
    "smb200_vec_free": (_i32, [_p]),
    "smb200_vec_dim": (_i32, [_p, _u64p]),
    "smb200_vec_device_ptr": (_i32, [_p, _pp]),
    "smb200_vec_upload": (_i32, [_p, _p, _u64]),
    "smb200_vec_download": (_i32, [_p, _p, _u64]),
    "smb200_vec_clone": (_i32, [_p, _pp]),
    "smb200_vec_copy": (_i32, [_p, _p]),
    "smb200_vec_fill": (_i32, [_p, C.c_double]),
    "smb200_vec_fill_uniform": (_i32, [_p, _u64]),
    "smb200_vec_add": (_i32, [_p, _p]),
    "smb200_vec_sub": (_i32, [_p, _p]),
    "smb200_vec_scale": (_i32, [_p, C.c_double]),
    "smb200_vec_axpy": (_i32, [_p, C.c_double, _p]),
    "smb200_vec_scale_add": (_i32, [_p, C.c_double, _p]),
    "smb200_vec_dot": (_i32, [_p, _p, _dp]),
    "smb200_vec_norm2sq": (_i32, [_p, _dp]),
    "smb200_vec_norm": (_i32, [_p, _dp]),
    "smb200_crs_upload": (_i32, [_p, _i32, _i32, _u64, _u64, _u64, _p, _p, _p, _pp]),
    "smb200_crs_from_indexlist": (_i32, [_p, _i32, _i32, _u64, _u64, _u64, _p, _p, _p, _p, _pp]),
    "smb200_crs_free": (_i32, [_p]),
    "smb200_crsfile_write": (_i32, [C.c_char_p, _i32, _i32, _u64, _u64, _u64, _p, _p, _p]),
    "smb200_crsfile_info": (_i32, [C.c_char_p, C.POINTER(_i32), C.POINTER(_i32), _u64p]),
    "smb200_crsfile_read": (_i32, [C.c_char_p, _p, _u64, _p, _u64, _p, _u64]),
    "smb200_crs_save": (_i32, [_p, C.c_char_p]),
    "smb200_crs_load": (_i32, [_p, C.c_char_p, _pp]),
    "smb200_crs_dims": (_i32, [_p, _u64p]),
    "smb200_crs_types": (_i32, [_p, C.POINTER(_i32), C.POINTER(_i32)]),
    "smb200_crs_download": (_i32, [_p, _p, _p, _p]),
    "smb200_crs_scale": (_i32, [_p, C.c_double]),
    "smb200_crs_configure": (_i32, [_p, _i32, _i32, C.c_uint32]),
    "smb200_crs_plan_info": (_i32, [_p, C.POINTER(PlanInfo)]),
    "smb200_gen_laplace": (_i32, [_p, _i32, _i32, _u64, _u64, _u64, _u64, _u64, _pp]),
    "smb200_gen_powerlaw": (_i32, [_p, _i32, _i32, _u64, _u64, _u64, _u64, _u64, _u64, _pp]),
    "smb200_spmv": (_i32, [_p, _p, _p]),
    "smb200_spmv_host": (_i32, [_p, _p, _u64, _p]),
    "smb200_crs_transpose": (_i32, [_p, _pp]),
    "smb200_bilinear": (_i32, [_p, _p, _p, _dp]),
    "smb200_cg_solve": (_i32, [_p, _p, _p, C.c_double, _i32, _u64, C.POINTER(CgStats)]),
    "smb200_cg_history": (_i32, [_p, _dp, _u64, _u64p]),
    "smb200_crs_diagonal": (_i32, [_p, _p]),
    "smb200_pcg_jacobi_solve": (_i32, [_p, _p, _p, C.c_double, _i32, _u64, C.POINTER(CgStats)]),
    "smb200_par_locate": (_i32, [_u64, _u64, _u64, _u64p, _u64p]),
    "smb200_par_create": (_i32, [_p, _u64, _u64, _i32, _i32, _pp]),
    "smb200_par_free": (_i32, [_p]),
    "smb200_par_owner": (_i32, [_p, _u64, C.POINTER(_i32)]),
    "smb200_par_set_block_indexlist": (_i32, [_p, _u64, _u64, _u64, _u64, _p, _p, _p, _p]),
    "smb200_par_set_block_crs": (_i32, [_p, _u64, _u64, _u64, _u64, _p, _p, _p]),
    "smb200_par_dims": (_i32, [_p, _u64p]),
    "smb200_par_block": (_i32, [_p, _u64, _pp]),
    "smb200_par_mvp": (_i32, [_p, _p, _p]),
    "smb200_partition_rows": (_i32, [_u64, C.c_uint32, _u64, _u64p]),
    "smb200_partition_rows_by_nnz": (_i32, [_i32, _u64, _p, C.c_uint32, _u64p]),
    "smb200_ghost_plan": (_i32, [_i32, _u64, _p, C.c_uint32, C.c_uint32, _u64p, _p, _u64p, _u64p, _u64p]),
    "smb200_comm_unique_id": (_i32, [_p]),
    "smb200_comm_init": (_i32, [_p, _i32, _i32, _p]),
    "smb200_comm_destroy": (_i32, [_p]),
    "smb200_dist_create": (_i32, [_p, _i32, _i32, _u64, _u64p, _u64, _p, _p, _p, _pp]),
    "smb200_dist_laplace": (_i32, [_p, _i32, _i32, _u64, _u64, _u64, _pp]),
    "smb200_dist_free": (_i32, [_p]),
    "smb200_dist_dims": (_i32, [_p, _u64p]),
    "smb200_dist_local": (_i32, [_p, _pp]),
    "smb200_dist_vec_create": (_i32, [_p, _pp]),
    "smb200_dist_spmv": (_i32, [_p, _p, _p]),
    "smb200_dist_dot": (_i32, [_p, _p, _p, _dp]),
    "smb200_dist_barrier": (_i32, [_p]),
    "smb200_dist_info": (_i32, [_p, _u64p]),
    "smb200_dist_cg_solve": (_i32, [_p, _p, _p, C.c_double, _i32, _u64, C.POINTER(CgStats)]),
    "smb200_dist_cg_solve_sr": (_i32, [_p, _p, _p, C.c_double, _i32, _u64, C.POINTER(CgStats)]),
    # smb200_host.h
    "smb200_il_create": (_i32, [_i32, _i32, _pp]),
    "smb200_il_free": (_i32, [_p]),
    "smb200_il_apply": (_i32, [_p, _u64, _p, _p, _p, _i32]),
    "smb200_il_get": (_i32, [_p, _u64, _u64, _dp]),
    "smb200_il_dims": (_i32, [_p, _u64p]),
    "smb200_il_export": (_i32, [_p, _p, _p, _p, _p]),
    "smb200_il_to_crs": (_i32, [_p, _p, _pp]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.smb200_last_error() or b"").decode("utf-8", "replace")


def check(status: int) -> None:
    """Raise for a non-zero status.  The reference's panics keep the reference's messages."""
    if status == OK:
        return
    if status == ERR_DIM:
        raise Panic("Dimension mismatch")               # densevec.rs:52-54
    if status == ERR_NOT_SQUARE:
        raise Panic("Matrix is not symmetric")          # linearsolver.rs:30-32
    if status == ERR_SIZE_MISMATCH:
        raise Panic("Matrix and vector size mismatch")  # linearsolver.rs:33-36
    raise SmbError(status, last_error())


def vtype_of(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise TypeError(f"value type must be float32 or float64 (FloatType), got {dt}")


def itype_of(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.uint32:
        return U32
    if dt == np.uint64:
        return U64
    raise TypeError(f"index type on the GPU path must be uint32 or uint64, got {dt}")


VDTYPES = {F32: np.dtype(np.float32), F64: np.dtype(np.float64)}
IDTYPES = {U32: np.dtype(np.uint32), U64: np.dtype(np.uint64)}


def ptr(a: np.ndarray | None):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
