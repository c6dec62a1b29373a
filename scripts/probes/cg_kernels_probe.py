#!/usr/bin/env python
"""GPU: a short eager (no graph) CG run on C4's matrix for a per-kernel launch list:
    SMB200_CG_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/probes/cg_kernels_probe.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sparsemat_b200 as smb  # noqa: E402

ctx = smb.Context(0)
n = 256
a = smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, n, n, n)
xs = smb.DenseVec(ctx, n ** 3, np.float64)
xs.fill_uniform(6)
b = a.mvp(xs)
x0 = smb.DenseVec(ctx, n ** 3, np.float64)
st = smb.ConjugateGradient(1e-30, int(sys.argv[1]) if len(sys.argv) > 1 else 12).solve_with_stats(a, b, x0)
print(st)
