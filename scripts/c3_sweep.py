#!/usr/bin/env python
"""GPU: BASELINE.json configs[2] (power-law rows, f64/u64, N = 50 M): band-split product over band widths, with the tuned
band kernel (L2 eviction hints, prefetched y) and with the plain stream kernel per band.  usage: python scripts/c3_sweep.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
ctx = smb.Context(0)
a = smb.SparseMatCRS.powerlaw(ctx, np.float64, np.uint64, n)
x = smb.DenseVec(ctx, n, np.float64)
x.fill_uniform(7)
y = smb.DenseVec(ctx, n, np.float64)


def run(label, reps=5):
    pi = a.plan_info()
    for _ in range(2):
        a.mvp(x, out=y)
    ctx.sync()
    e0 = ctx.event().record()
    for _ in range(reps):
        a.mvp(x, out=y)
    e1 = ctx.event().record()
    ms = e0.elapsed_ms(e1) / reps
    print(f"{label:44s} launches={pi['launches_per_spmv']:2d} {ms * 1e3:9.1f} us  effective {pi['algorithmic_bytes'] / ms / 1e6:7.1f} GB/s  "
          f"moved {pi['stream_bytes'] / 1e9:6.2f} GB at {pi['stream_bytes'] / ms / 1e6:7.1f} GB/s  plan {pi['plan_ms']:.0f} ms", flush=True)
    return y.to_numpy()


a.configure(smb.SPMV_STREAM)
ref = run("stream")
for w in [int(v) for v in os.environ.get("WIDTHS", "4194304,5600000,7200000,8400000,10000000").split(",")]:
    os.environ["SMB200_BANDSPLIT_WIDTH"] = str(w)
    a.configure(smb.SPMV_BANDSPLIT)
    for bk in (1, 0):
        os.environ["SMB200_BAND_KERNEL"] = str(bk)
        got = run(f"bandsplit W={w} ({w * 8 / 2**20:.0f} MiB) band_kernel={bk}")
    print(f"    max |diff| vs stream {np.abs(got - ref).max():.3e}", flush=True)
