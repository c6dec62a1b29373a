#!/usr/bin/env python
"""GPU (run plain, then under ncu --metrics dram bytes): one C3-style power-law SpMV (N = 20 M, f64/u64, stream kernel) and
a few CG iterations on C4 (256^3 f64/u32, ring + vector kernels) — DRAM traffic evidence for DESIGN.md §7."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb
ctx = smb.Context(0)
n = 20_000_000
a = smb.SparseMatCRS.powerlaw(ctx, np.float64, np.uint64, n)
x = smb.DenseVec(ctx, n, np.float64); x.fill_uniform(1)
y = smb.DenseVec(ctx, n, np.float64)
info = a.plan_info()
for _ in range(3):
    a.mvp(x, out=y)
ctx.sync()
print("C3/2.5:", info["variant_name"], "nnz", info["nnz"], "algorithmic bytes", info["algorithmic_bytes"], flush=True)
del a, x, y
m = smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 256, 256, 256)
b = smb.DenseVec(ctx, 256 ** 3, np.float64); b.fill_uniform(6)
os.environ["SMB200_CG_GRAPH"] = "0"
st = smb.ConjugateGradient(1e-30, 6).solve_with_stats(m, b, smb.DenseVec(ctx, 256 ** 3, np.float64))
print("CG:", st, "spmv bytes", m.plan_info()["algorithmic_bytes"], flush=True)
