#!/usr/bin/env python
"""GPU, meant to run under ncu: the headline SpMV (C2 = 256^3 f32/u32 Laplacian, AUTO = packed ring) launched a few times,
nothing else — a cheap target for `ncu --set full -k regex:spmv_ring -s 3 -c 1` when bench.py under ncu is too slow.
usage: ncu ... python scripts/ncu_ring_c2.py [f32|f64]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

vdt = np.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else np.float32
ctx = smb.Context(0)
a = smb.SparseMatCRS.laplace(ctx, vdt, np.uint32, 256, 256, 256)
x = smb.DenseVec(ctx, a.n_cols(), vdt)
x.fill_uniform(2)
y = smb.DenseVec(ctx, a.n_rows(), vdt)
for _ in range(6):
    a.mvp(x, out=y)
ctx.sync()
pi = a.plan_info()
print({k: pi[k] for k in ("variant_name", "n_blocks", "algorithmic_bytes", "stream_bytes", "nnz_c16", "rows_o16")})
