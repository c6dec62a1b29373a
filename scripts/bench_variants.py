#!/usr/bin/env python
"""GPU: times every SpMV kernel family on the BASELINE.json single-GPU configs (CUDA events, L2-cold where the
matrix fits in L2).  usage: python scripts/bench_variants.py [c1] [c2] [c2f64] [c3small] ..."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

PEAK = 6540.2
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def time_spmv(ctx, a, x, y, reps, flush):
    for _ in range(3):
        a.mvp(x, out=y)
    ctx.sync()
    if flush:
        tot = 0.0
        for _ in range(reps):
            ctx.flush_l2()
            e0, e1 = ctx.event().record(), None
            a.mvp(x, out=y)
            e1 = ctx.event().record()
            tot += e0.elapsed_ms(e1)
        return tot / reps
    e0 = ctx.event().record()
    for _ in range(reps):
        a.mvp(x, out=y)
    e1 = ctx.event().record()
    return e0.elapsed_ms(e1) / reps


def run(ctx, name, a, variants, reps=50, flush=False):
    info = a.plan_info()
    B = info["algorithmic_bytes"]
    x = smb.DenseVec(ctx, a.n_cols(), a.dtype)
    x.fill_uniform(2)
    y = smb.DenseVec(ctx, a.n_rows(), a.dtype)
    ref = None
    for v, lanes in variants:
        try:
            a.configure(v, lanes)
        except smb.SmbError as e:
            print(f"{name:10s} {smb.VARIANT_NAMES[v]:>11s}/{lanes:<2d}  unsupported: {e}")
            continue
        pi = a.plan_info()
        ms = time_spmv(ctx, a, x, y, reps, flush)
        got = y.to_numpy()
        if ref is None:
            ref = got
        err = float(np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64)))) if got.size else 0.0
        print(f"{name:10s} {pi['variant_name']:>11s}/{pi['lanes']:<2d} {ms * 1e3:9.1f} us  {B / ms / 1e6:8.1f} GB/s  "
              f"{100 * B / ms / 1e6 / PEAK:5.1f}% of measured  {2 * info['nnz'] / ms / 1e6:8.1f} GFLOP/s  maxdiff {err:.2e}"
              f"{'  (L2 flushed)' if flush else ''}", flush=True)


def main():
    which = sys.argv[1:] or ["c2"]
    ctx = smb.Context(0)
    V = [(smb.SPMV_STREAM, 0), (smb.SPMV_RING, 0), (smb.SPMV_STREAM_TMA, 0), (smb.SPMV_STREAM_PIPE, 0), (smb.SPMV_VECTOR, 8), (smb.SPMV_VECTOR, 4),
         (smb.SPMV_SCALAR, 0), (smb.SPMV_AUTO, 0)]
    for w in which:
        if w == "c1":
            a = smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 1024, 1024, 1)
            run(ctx, "C1 warm", a, V + [(smb.SPMV_BANDED, 0)], reps=200)
            run(ctx, "C1 cold", a, V + [(smb.SPMV_BANDED, 0)], reps=20, flush=True)
        elif w == "c2":
            run(ctx, "C2 f32", smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, 256, 256, 256), V)
        elif w == "c2f64":
            run(ctx, "C4 f64", smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 256, 256, 256), V)
        elif w == "c3":
            a = smb.SparseMatCRS.powerlaw(ctx, np.float64, np.uint64, 50_000_000)
            run(ctx, "C3 f64u64", a, [(smb.SPMV_STREAM, 0), (smb.SPMV_STREAM_PIPE, 0), (smb.SPMV_VECTOR, 16), (smb.SPMV_VECTOR, 8), (smb.SPMV_AUTO, 0)], reps=10)
        elif w == "c3small":
            a = smb.SparseMatCRS.powerlaw(ctx, np.float64, np.uint64, 5_000_000)
            run(ctx, "C3/10", a, [(smb.SPMV_STREAM, 0), (smb.SPMV_STREAM_PIPE, 0), (smb.SPMV_VECTOR, 16), (smb.SPMV_VECTOR, 8), (smb.SPMV_AUTO, 0)], reps=10)
        elif w == "c5":
            run(ctx, "C5 f32", smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, 512, 512, 512), V, reps=20)



if __name__ == "__main__":
    main()
