#!/usr/bin/env python
"""GPU, torchrun: BASELINE.json configs[4] — the 3-D 7-point Laplacian 512^3 row-partitioned (z-slabs) over P GPUs, strong
scaling: SpMV f32/u32 (aggregate effective GB/s, max over ranks of CUDA-event time) and CG f64/u32 iter/s."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ctx = smb.Context(local)
box = [smb.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, 0)
ctx.comm_init(rank, world, box[0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = n ** 3
nnz = 7 * N - 6 * n * n


def maxr(v):
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"config": f"3-D 7-point Laplacian {n}^3, z-slab row blocks", "n_gpus": world}
a = smb.DistCRS.laplace(ctx, np.float32, np.uint32, n, n, n)
x, y = a.new_vec(), a.new_vec()
x.fill_uniform(2 + rank)
for _ in range(5):
    a.mvp(x, out=y)
ctx.sync(); dist.barrier()
e0 = ctx.event().record()
steps = 50
for _ in range(steps):
    a.mvp(x, out=y)
e1 = ctx.event().record()
ms = maxr(e0.elapsed_ms(e1)) / steps
B = nnz * 8 + (N + 1) * 4 + 2 * N * 4
out["spmv_f32"] = {"ms_per_spmv": ms, "aggregate_gbs": B / ms / 1e6, "gflops": 2 * nnz / ms / 1e6, "kernel": a.local.plan_info()["variant_name"]}
del a, x, y
a = smb.DistCRS.laplace(ctx, np.float64, np.uint32, n, n, n)
xs = a.new_vec()
xs.fill_uniform(6 + rank)
b = a.mvp(xs)
smb.ConjugateGradient(1e-30, 17).solve_with_stats(a, b, a.new_vec())
ctx.sync(); dist.barrier()
st = smb.ConjugateGradient(1e-30, 100).solve_with_stats(a, b, a.new_vec())
cg_ms = maxr(st["device_ms"])
Bcg = nnz * 12 + (N + 1) * 4 + 2 * N * 8 + 9 * N * 8
out["cg_f64"] = {"iterations": int(st["iterations"]), "ms": cg_ms, "iter_per_s": st["iterations"] / cg_ms * 1e3, "effective_gbs": Bcg * st["iterations"] / cg_ms / 1e6}
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
