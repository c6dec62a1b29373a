#!/usr/bin/env python
"""GPU: every kernel family on small matrices of all four type combinations, the vector kernels, CG, to_crs and the host
pipeline, each checked against the oracle — small on purpose so that the same command can be run under
`compute-sanitizer --tool memcheck` where that tool is available (it is closed on this pool's boxes: gpurun answers that the
tool is unavailable, so here the script only runs plain, from tests/test_gpu_spmv.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("SMB200_HOST_CHUNKS", "3")
import cases  # noqa: E402
import sparsemat_b200 as smb  # noqa: E402

ctx = smb.Context(0)
V = [(smb.SPMV_SCALAR, 0), (smb.SPMV_VECTOR, 4), (smb.SPMV_VECTOR, 32), (smb.SPMV_STREAM, 0), (smb.SPMV_STREAM_TMA, 0),
     (smb.SPMV_BANDED, 0), (smb.SPMV_STREAM_PIPE, 0), (smb.SPMV_RING, 0)]
for vdt, idt in ((np.float32, np.uint32), (np.float64, np.uint32), (np.float32, np.uint64), (np.float64, np.uint64)):
    mats = [cases.ragged(1, 700, 650, 30, vdt, idt), cases.powerlaw(2, 3000, 3000, 5000, vdt, idt), cases.banded(3, 5000, 200, 7, vdt, idt),
            cases.giant_row(4, 300, 20000, 30000, vdt, idt), cases.all_empty(9, 4, vdt, idt)]
    for n_rows, n_cols, vals, cols, offs in mats:
        a = smb.SparseMatCRS.from_raw_parts(ctx, n_rows, n_cols, vals, cols, offs)
        x = smb.DenseVec(ctx, n_cols, vdt)
        x.fill_uniform(1)
        ref = None
        for v, lanes in V:
            a.configure(v, lanes)
            print(f"{np.dtype(vdt)}/{np.dtype(idt)} rows={n_rows} nnz={vals.size} variant={smb.VARIANT_NAMES[v]}/{lanes} plan={a.plan_info()['variant_name']} blocks={a.plan_info()['n_blocks']}", flush=True)
            y = a.mvp(x).to_numpy()
            ref = y if ref is None else ref
            assert np.allclose(y, ref, rtol=1e-4, atol=1e-4)
        a.configure(smb.SPMV_AUTO)
        a.mvp_host(x.to_numpy())
        lhs = smb.DenseVec(ctx, n_rows, vdt)
        lhs.fill_uniform(2)
        a.inner_prod(lhs, x)
    lap = smb.SparseMatCRS.laplace(ctx, vdt, idt, 12, 11, 10)
    for v in (smb.SPMV_AUTO, smb.SPMV_STREAM, smb.SPMV_STREAM_PIPE, smb.SPMV_RING, smb.SPMV_VECTOR):
        lap.configure(v)
        b = smb.DenseVec(ctx, 1320, vdt)
        b.fill(1.0)
        xs = smb.DenseVec(ctx, 1320, vdt)
        st = smb.ConjugateGradient(1e-6 if vdt == np.float32 else 1e-12, 400, relative=True).solve_with_stats(lap, b, xs)
        assert st["converged"], st
    u, w = smb.DenseVec(ctx, 1003, vdt), smb.DenseVec(ctx, 1003, vdt)
    u.fill_uniform(3); w.fill_uniform(4)
    u.add(w); u.sub(w); u.scale(1.5); u.axpy(0.5, w); u.scale_add(0.25, w); u.inner_prod(w); u.norm()
    sp = smb.SparseMatIndexList(vdt, idt)
    rng = np.random.default_rng(5)
    sp.set(rng.integers(0, 400, 3000), rng.integers(0, 300, 3000), rng.uniform(-1, 1, 3000).astype(vdt))
    sp.to_crs(ctx).mvp(smb.DenseVec(ctx, 300, vdt))
g = smb.SparseMatCRS.powerlaw(ctx, np.float64, np.uint64, 20000, max_len=4000)
g.mvp(smb.DenseVec(ctx, 20000, np.float64))
ctx.sync()
print("sanitize_small: ok, launches", smb.api.lib.smb200_launch_count())
