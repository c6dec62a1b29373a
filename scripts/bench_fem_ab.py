#!/usr/bin/env python
"""GPU: rows of 27..162 entries (27-point stencil x dof unknowns per node, f32/u32 and f64/u32): the TMA ring with several
threads per row against the stream kernel.  usage: python scripts/bench_fem_ab.py [n=64]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases  # noqa: E402
import sparsemat_b200 as smb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = smb.Context(0)
for vdt in (np.float32, np.float64):
    for dof in (1, 2, 3):
        case = cases.fem_like(n, n, n, dof, vdt, np.uint32, seed=1)
        a = smb.SparseMatCRS.from_raw_parts(ctx, *case)
        x = smb.DenseVec(ctx, case[1], vdt)
        x.fill_uniform(3)
        y = smb.DenseVec(ctx, case[0], vdt)
        res = {}
        for name, var in (("auto", smb.SPMV_AUTO), ("stream", smb.SPMV_STREAM)):
            a.configure(var)
            pi = a.plan_info()
            for _ in range(5):
                a.mvp(x, out=y)
            ctx.sync()
            e0 = ctx.event().record()
            for _ in range(50):
                a.mvp(x, out=y)
            e1 = ctx.event().record()
            ms = e0.elapsed_ms(e1) / 50
            res[name] = (ms, pi, y.to_numpy())
            print(f"{np.dtype(vdt).name} dof={dof} rows={case[0]} nnz={case[2].size} max_row={pi['max_row_len']} {name:6s} -> {pi['variant_name']:6s} "
                  f"lanes={pi['lanes']} {ms * 1e3:8.1f} us  algorithmic {pi['algorithmic_bytes'] / ms / 1e6:7.1f} GB/s  streamed "
                  f"{pi['stream_bytes'] / ms / 1e6:7.1f} GB/s", flush=True)
        d = np.abs(res["auto"][2].astype(np.float64) - res["stream"][2].astype(np.float64)).max()
        print(f"    max |auto - stream| = {d:.3e}", flush=True)
        del a, x, y
