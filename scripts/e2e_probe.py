#!/usr/bin/env python
"""GPU: end-to-end host-buffer SpMV (smb200_spmv_host) on C2 for the chunk count in SMB200_HOST_CHUNKS."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb
ctx = smb.Context(0)
n = 256 ** 3
a = smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, 256, 256, 256)
hx, hy = smb.pinned_empty(n, np.float32), smb.pinned_empty(n, np.float32)
hx[:] = np.random.default_rng(0).uniform(-1, 1, n).astype(np.float32)
for _ in range(3):
    a.mvp_host(hx, hy)
ctx.sync()
t0 = time.perf_counter()
for _ in range(20):
    a.mvp_host(hx, hy)
dt = (time.perf_counter() - t0) / 20
print(f"chunks={os.environ.get('SMB200_HOST_CHUNKS', 'default')}: {dt * 1e3:.3f} ms/step, {a.plan_info()['algorithmic_bytes'] / dt / 1e9:.0f} GB/s", flush=True)
