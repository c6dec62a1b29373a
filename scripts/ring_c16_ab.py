#!/usr/bin/env python
"""GPU: A/B of the ring kernel with and without plan-time index compression (16-bit window positions instead of the
column array), on the C2 / C4 / C1 matrices.  Effective GB/s counts ALGORITHMIC bytes in both arms;
`streamed` is what the kernel really moves.  usage: python scripts/ring_c16_ab.py [c2] [c4] [c1] [cg]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402


def time_spmv(ctx, a, x, y, reps):
    for _ in range(5):
        a.mvp(x, out=y)
    ctx.sync()
    e0 = ctx.event().record()
    for _ in range(reps):
        a.mvp(x, out=y)
    e1 = ctx.event().record()
    return e0.elapsed_ms(e1) / reps


ARMS = [("packed fill=70 (default)", {}), ("packed fill=80", {"SMB200_RING_FILL": "80"}),
        ("packed, full-width offsets", {"SMB200_RING_O16": "0"}),
        ("c16, worst-case stage", {"SMB200_RING_PACK": "0"}), ("full-width columns", {"SMB200_RING_C16": "0"})]


def ab(ctx, name, a, reps=200, caps=None):
    x = smb.DenseVec(ctx, a.n_cols(), a.dtype)
    x.fill_uniform(2)
    y = smb.DenseVec(ctx, a.n_rows(), a.dtype)
    ref = None
    for label, env in ARMS:
        for k in ("SMB200_RING_FILL", "SMB200_RING_PACK", "SMB200_RING_C16", "SMB200_RING_O16"):
            os.environ.pop(k, None)
        os.environ.update(env)
        a.configure(smb.SPMV_RING)
        pi = a.plan_info()
        ms = time_spmv(ctx, a, x, y, reps)
        got = y.to_numpy()
        if ref is None:
            ref = got
        same = bool(np.array_equal(got, ref))
        B, S = pi["algorithmic_bytes"], pi["stream_bytes"]
        print(f"{name:8s} {label:26s} blocks={pi['n_blocks']:6d} nnz_c16={pi['nnz_c16']:>10d} "
              f"{ms * 1e3:8.1f} us  effective {B / ms / 1e6:7.1f} GB/s  streamed {S / ms / 1e6:7.1f} GB/s  identical={same}", flush=True)
    for k in ("SMB200_RING_FILL", "SMB200_RING_PACK", "SMB200_RING_C16", "SMB200_RING_O16"):
        os.environ.pop(k, None)


def main():
    which = sys.argv[1:] or ["c2", "c4"]
    ctx = smb.Context(0)
    for w in which:
        if w == "c2":
            ab(ctx, "C2 f32", smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, 256, 256, 256))
        elif w == "c4":
            ab(ctx, "C4 f64", smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 256, 256, 256))
        elif w == "c1":
            ab(ctx, "C1 f64", smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 1024, 1024, 1), reps=500)
        elif w == "cg":
            for c16 in ("1", "0"):
                os.environ["SMB200_RING_C16"] = c16   # (1 = default plan: packed)
                a = smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 256, 256, 256)
                n = a.n_rows()
                xs = smb.DenseVec(ctx, n, np.float64)
                xs.fill_uniform(6)
                b = a.mvp(xs)
                x = smb.DenseVec(ctx, n, np.float64)
                st = smb.ConjugateGradient(1e-8, 10_000, relative=True).solve_with_stats(a, b, x)
                x.fill(0.0)
                st = smb.ConjugateGradient(1e-8, 10_000, relative=True).solve_with_stats(a, b, x)
                print(f"CG C4 c16={c16}: {st['iterations']} it, {st['device_ms']:.1f} ms, {st['iterations'] / st['device_ms'] * 1e3:.0f} it/s", flush=True)
            os.environ.pop("SMB200_RING_C16", None)


if __name__ == "__main__":
    main()
