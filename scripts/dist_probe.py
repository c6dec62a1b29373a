#!/usr/bin/env python
"""GPU, launched with torchrun: times the pieces of the distributed path (SpMV, dot, CG iteration) per rank."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sparsemat_b200 as smb  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(os.environ.get("PROBE_BACKEND", "gloo"), device_id=torch.device("cuda", local) if os.environ.get("PROBE_BACKEND") == "nccl" else None)
if os.environ.get("PROBE_BACKEND") == "nccl":
    dist.barrier()
ctx = smb.Context(local)
box = [smb.Context.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, 0)
ctx.comm_init(rank, world, box[0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
a = smb.DistCRS.laplace(ctx, np.float64, np.uint32, n, n, n * world)
x, y = a.new_vec(), a.new_vec()
x.fill_uniform(3 + rank)


def timed(label, fn, reps):
    for _ in range(3):
        fn()
    ctx.sync(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    dt = (time.perf_counter() - t0) / reps
    if rank == 0:
        print(f"{label}: {dt * 1e6:.1f} us", flush=True)


timed("dist spmv f64", lambda: a.mvp(x, out=y), 100)
timed("dist dot (kernel + all-reduce + D2H sync)", lambda: a.dot(x, y), 100)
b = a.mvp(x)
for iters in (8, 64, 200):
    x0 = a.new_vec()
    ctx.sync(); dist.barrier()
    t0 = time.perf_counter()
    st = smb.ConjugateGradient(1e-30, iters).solve_with_stats(a, b, x0)
    wall = time.perf_counter() - t0
    if rank == 0:
        print(f"dist cg {iters} iterations: device {st['device_ms']:.2f} ms ({st['device_ms'] / iters * 1e3:.0f} us/iter), wall {wall * 1e3:.2f} ms, "
              f"launches {st['launches']}", flush=True)
dist.barrier()
