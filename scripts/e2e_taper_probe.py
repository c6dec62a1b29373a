#!/usr/bin/env python
"""GPU: smb200_spmv_host on C2 per chunk tapering mode (SMB200_HOST_TAPER = 0 uniform, 1 half-size ends, 2 geometric ends)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

ctx = smb.Context(0)
for taper, chunks in [(1, 0), (2, 0), (0, 0), (2, 6), (2, 12)]:
    os.environ["SMB200_HOST_TAPER"] = str(taper)
    os.environ["SMB200_HOST_CHUNKS"] = str(chunks)
    a = smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, 256, 256, 256)
    n = a.n_rows()
    hx, hy = smb.pinned_empty(n, np.float32), smb.pinned_empty(n, np.float32)
    hx[:] = 1.0
    for _ in range(3):
        a.mvp_host(hx, hy)
    t0 = time.perf_counter()
    for _ in range(20):
        a.mvp_host(hx, hy)
    dt = (time.perf_counter() - t0) / 20
    print(f"smb200_spmv_host C2, taper={taper} chunks={chunks or 'auto'}: {dt * 1e3:.3f} ms/step", flush=True)
    del a
