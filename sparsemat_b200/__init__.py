"""sparsemat_b200 — the B200 (sm_100a) engine behind lostinc0de/sparsemat's CRS SpMV / CG hot path.

The product is ``lib/libsmb200.so`` (hand-written CUDA + the C ABI of ``include/smb200.h``); this
package is the thin Python mirror of the reference's interface used by the tests and the bench.
"""
from .api import (  # noqa: F401
    ConjugateGradient, Context, DenseVec, DistCRS, Event, JacobiPCG, Panic, SmbError, SparseMatCRS, SparseMatIndexList,
    SparseMatPar, crs_from_indexlist_arrays, crsfile_read, crsfile_write, ghost_plan, partition_rows, partition_rows_by_nnz, pinned_empty,
)
from ._ffi import (  # noqa: F401
    FLAG_L2_PERSIST_X, SPMV_AUTO, SPMV_BANDED, SPMV_BANDSPLIT, SPMV_RING, SPMV_SCALAR, SPMV_STREAM, SPMV_STREAM_PIPE, SPMV_STREAM_TMA, SPMV_VECTOR, VARIANT_NAMES,
)

__version__ = "0.1.0"
