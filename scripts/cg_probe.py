#!/usr/bin/env python
"""GPU: CG on C4 (256^3 f64/u32) for a fixed number of iterations; prints iter/s.  Used under ncu for the launch list."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ctx = smb.Context(0)
a = smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, n, n, n)
xs = smb.DenseVec(ctx, n ** 3, np.float64)
xs.fill_uniform(6)
b = a.mvp(xs)
for rep in range(2):
    x0 = smb.DenseVec(ctx, n ** 3, np.float64)
    st = smb.ConjugateGradient(1e-30, iters).solve_with_stats(a, b, x0)
    B = a.plan_info()["algorithmic_bytes"] + 9 * n ** 3 * 8
    print(f"rep {rep}: {st['iterations']} iterations in {st['device_ms']:.2f} ms -> {st['iterations'] / st['device_ms'] * 1e3:.1f} it/s, "
          f"{B * st['iterations'] / st['device_ms'] / 1e6:.0f} GB/s, launches {st['launches']}", flush=True)
