#!/usr/bin/env python
"""GPU: BASELINE.json configs[2] (power-law rows, f64/u64, N = 50 M) through one kernel family, a few products — the
command profiled by ncu for the C3 work.  usage: python scripts/c3_probe.py [variant=bandsplit] [reps=2] [n_rows=50000000]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "bandsplit"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000_000
ctx = smb.Context(0)
a = smb.SparseMatCRS.powerlaw(ctx, np.float64, np.uint64, n)
x = smb.DenseVec(ctx, n, np.float64)
x.fill_uniform(7)
y = smb.DenseVec(ctx, n, np.float64)
t0 = time.perf_counter()
a.configure({"bandsplit": smb.SPMV_BANDSPLIT, "stream": smb.SPMV_STREAM, "auto": smb.SPMV_AUTO}[variant])
ctx.sync()
pi = a.plan_info()
print(f"plan {pi['variant_name']} built in {time.perf_counter() - t0:.2f} s, {pi['launches_per_spmv']} launches per product, "
      f"plan bytes {pi['plan_bytes'] / 1e9:.2f} GB", flush=True)
a.mvp(x, out=y)
ctx.sync()
e0 = ctx.event().record()
for _ in range(reps):
    a.mvp(x, out=y)
e1 = ctx.event().record()
ms = e0.elapsed_ms(e1) / reps
print(f"{variant}: {ms * 1e3:.1f} us per product, effective {pi['algorithmic_bytes'] / ms / 1e6:.1f} GB/s, "
      f"moved {pi['stream_bytes'] / ms / 1e6:.1f} GB/s", flush=True)
