#!/usr/bin/env python
"""GPU: BASELINE.json configs[2] (power-law rows, uniform random columns, f64/u64, N = 50 M) — the stream kernel against the
band-split product (column bands of about half an L2 of x).  Effective GB/s counts the ALGORITHMIC bytes of the CRS format
in both arms.  usage: python scripts/bench_bandsplit_c3.py [n_rows] [width ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
widths = [int(w) for w in sys.argv[2:]] or [0]
ctx = smb.Context(0)
t0 = time.perf_counter()
a = smb.SparseMatCRS.powerlaw(ctx, np.float64, np.uint64, n)
ctx.sync()
print(f"generated N={n} nnz={a.n_non_zero_entries()} in {time.perf_counter() - t0:.1f} s", flush=True)
x = smb.DenseVec(ctx, n, np.float64)
x.fill_uniform(2)
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
    B = pi["algorithmic_bytes"]
    print(f"{label:28s} plan={pi['variant_name']:9s} launches/spmv={pi['launches_per_spmv']:2d} {ms * 1e3:9.1f} us  effective {B / ms / 1e6:7.1f} GB/s  "
          f"moved {pi['stream_bytes'] / 1e9:6.2f} GB -> {pi['stream_bytes'] / ms / 1e6:7.1f} GB/s", flush=True)
    return y.to_numpy()


a.configure(smb.SPMV_STREAM)
ref = run("stream")
scale = None
for w in widths:
    if w:
        os.environ["SMB200_BANDSPLIT_WIDTH"] = str(w)
    else:
        os.environ.pop("SMB200_BANDSPLIT_WIDTH", None)
    t0 = time.perf_counter()
    a.configure(smb.SPMV_BANDSPLIT)
    ctx.sync()
    dt = time.perf_counter() - t0
    got = run(f"bandsplit width={w or 'L2/2'} (plan {dt:.2f} s)")
    d = np.abs(got - ref)
    print(f"    max |diff| vs stream {d.max():.3e}, max |y| {np.abs(ref).max():.3e}", flush=True)
# the same with an L2 persisting window over the band of x that is being gathered
os.environ.pop("SMB200_BANDSPLIT_WIDTH", None)
a.configure(smb.SPMV_BANDSPLIT, 0, smb.FLAG_L2_PERSIST_X)
run("bandsplit L2/2 + persisting x")
