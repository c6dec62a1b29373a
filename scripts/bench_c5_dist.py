#!/usr/bin/env python
"""GPU, torchrun: BASELINE.json configs[4] — the 3-D 7-point Laplacian 512^3 row-partitioned (z-slabs) over P GPUs, STRONG
scaling: SpMV f32/u32 (aggregate effective GB/s, max over ranks of CUDA-event time behind a device-side barrier) and CG
f64/u32 iter/s, with every rank's slice of y compared bit for bit with the oracle's rows of the global operator.

    python -m torch.distributed.run --nproc-per-node P scripts/bench_c5_dist.py [n=512] [steps=50]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparsemat_b200 as smb  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
ctx = smb.Context(local)
box = [smb.Context.comm_unique_id() if rank == 0 and world > 1 else None]
dist.broadcast_object_list(box, 0)
ctx.comm_init(rank, world, box[0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
N = n ** 3
nnz = 7 * N - 6 * n * n


def maxr(v):
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = {"config": f"3-D 7-point Laplacian {n}^3, z-slab row blocks, strong scaling", "n_gpus": world}
a = smb.DistCRS.laplace(ctx, np.float32, np.uint32, n, n, n)
d = a.dims()
x, y = a.new_vec(), a.new_vec()
x.fill_uniform(2 + rank)
for _ in range(5):
    a.mvp(x, out=y)
ctx.sync()
dist.barrier()
a.barrier()
e0 = ctx.event().record()
for _ in range(steps):
    a.mvp(x, out=y)
e1 = ctx.event().record()
ms = maxr(e0.elapsed_ms(e1)) / steps
B = nnz * 8 + (N + world) * 4 + 2 * N * 4
pi = a.local.plan_info()
out["spmv_f32"] = {"ms_per_spmv": ms, "aggregate_gbs": B / ms / 1e6, "gflops": 2 * nnz / ms / 1e6, "kernel": pi["variant_name"] + ("_sell" if pi["sell_entries"] else ""),
                   "data_path": "peer memory" if a.info()["p2p"] or world == 1 else "nccl", "steps": steps}
# parity: this rank's rows of the global operator through the oracle (x of rank q = uniform(seed 2 + q, its slab))
from oracle import oracle_py as orc  # noqa: E402
plane = n * n
lo, hi = d["row_lo"], d["row_lo"] + d["n_local"]
vals, cols, offs = orc.laplace(np.float32, np.uint32, n, n, n, lo, hi)
bounds = [n * q // world * plane for q in range(world + 1)]
parts, base = [], lo
if rank > 0:
    parts.append(orc.uniform(np.float32, 2 + rank - 1, bounds[rank] - bounds[rank - 1])[-plane:])
    base = lo - plane
parts.append(orc.uniform(np.float32, 2 + rank, d["n_local"]))
if rank + 1 < world:
    parts.append(orc.uniform(np.float32, 2 + rank + 1, bounds[rank + 2] - bounds[rank + 1])[:plane])
want = orc.mvp(vals, (cols.astype(np.int64) - base).astype(np.uint32), offs, np.ascontiguousarray(np.concatenate(parts)),
               threads=max(1, (os.cpu_count() or 1) // world))
ok = torch.tensor([1.0 if np.array_equal(y.to_numpy(), want) else 0.0], dtype=torch.float64)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
out["spmv_f32"]["y_bit_exact_all_ranks"] = bool(ok.item() == 1.0)
del a, x, y, vals, cols, offs, want
a = smb.DistCRS.laplace(ctx, np.float64, np.uint32, n, n, n)
xs = a.new_vec()
xs.fill_uniform(6 + rank)
b = a.mvp(xs)
x0 = a.new_vec()
iters = 100
smb.ConjugateGradient(1e-1, iters, relative=True).solve_with_stats(a, b, x0)       # untimed: workspace + iteration graph
x0.fill(0.0)
ctx.sync()
dist.barrier()
st = smb.ConjugateGradient(1e-30, iters).solve_with_stats(a, b, x0)
cg_ms = maxr(st["device_ms"])
Bcg = nnz * 12 + (N + world) * 4 + 2 * N * 8 + 9 * N * 8
out["cg_f64"] = {"iterations": int(st["iterations"]), "ms": cg_ms, "iter_per_s": st["iterations"] / cg_ms * 1e3,
                 "effective_gbs": Bcg * st["iterations"] / cg_ms / 1e6, "final_residual": st["final_residual"]}
x0.fill(0.0)
smb.ConjugateGradient(1e-1, iters, relative=True, single_reduce=True).solve_with_stats(a, b, x0)
x0.fill(0.0)
ctx.sync()
dist.barrier()
st = smb.ConjugateGradient(1e-30, iters, single_reduce=True).solve_with_stats(a, b, x0)
sr_ms = maxr(st["device_ms"])
out["cg_f64_single_reduce"] = {"iterations": int(st["iterations"]), "ms": sr_ms, "iter_per_s": st["iterations"] / sr_ms * 1e3,
                               "final_residual": st["final_residual"]}
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
