#!/usr/bin/env python
"""GPU: config C1 end to end — 2-D 5-point Laplacian 1024^2 f64/u32 assembled through the IndexList API on the host,
converted by to_crs() on the device (K8), checked bit for bit against the oracle's conversion, then one SpMV."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sparsemat_b200 as smb  # noqa: E402
from oracle import oracle_py as orc  # noqa: E402  (checker only)
from test_gpu_cg_tocrs import _laplace2d_entries  # noqa: E402

nx = ny = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = smb.Context(0)
i, j, v = _laplace2d_entries(nx, ny)
for scramble in (False, True):
    if scramble:
        p = np.random.default_rng(0xC0FFEE).permutation(i.size)
        i, j, v = i[p], j[p], v[p]
    sp = smb.SparseMatIndexList(np.float64, np.uint32)
    t0 = time.perf_counter()
    sp.set(i, j, v)
    t_asm = time.perf_counter() - t0
    cols, vals, pos, nxt = sp.raw_arrays()
    ctx.sync()
    t0 = time.perf_counter()
    a = sp.to_crs(ctx)
    ctx.sync()
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    ov, oc, oo = orc.to_crs_raw(sp.n_rows(), cols, vals, pos, nxt)
    t_cpu = time.perf_counter() - t0
    gv, gc, go = a.raw_parts()
    ok = gv.tobytes() == ov.tobytes() and np.array_equal(gc, oc) and np.array_equal(go, oo)
    x = orc.uniform(np.float64, 1, a.n_cols())
    y = a.mvp(smb.DenseVec.from_vec(ctx, x)).to_numpy()
    ok_mvp = np.array_equal(y, orc.mvp(ov, oc, oo, x))
    print(f"C1 {nx}x{ny} scramble={scramble}: nnz={vals.size} host assembly {t_asm*1e3:.0f} ms; to_crs GPU (H2D + 3 kernels) {t_gpu*1e3:.1f} ms "
          f"vs oracle CPU {t_cpu*1e3:.1f} ms; layout bit-exact={ok}; mvp bit-exact={ok_mvp}", flush=True)
