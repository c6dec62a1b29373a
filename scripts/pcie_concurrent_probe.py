#!/usr/bin/env python
"""GPU box, N ranks (torch.distributed.run): host-link bandwidth with ALL ranks copying at once — what bounds the
end-to-end (host buffers) leg of bench.py at N > 1.  Per rank: 64 MiB pinned H2D alone, D2H alone, both directions at
once on two streams; every phase starts behind a barrier so the ranks' copies overlap."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 64 << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def phase(h2d, d2h, reps=20):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9


res = [phase(True, False), phase(False, True), phase(True, True)]
t = torch.tensor(res, device="cuda", dtype=torch.float64)
if world > 1:
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(t)
    t /= world
else:
    lo = hi = t
if rank == 0:
    print(f"ranks={world}  per-rank GB/s (mean [min..max])  H2D alone {t[0]:.1f} [{lo[0]:.1f}..{hi[0]:.1f}]   D2H alone {t[1]:.1f} "
          f"[{lo[1]:.1f}..{hi[1]:.1f}]   both at once, each direction {t[2]:.1f} [{lo[2]:.1f}..{hi[2]:.1f}]   "
          f"=> a step that moves 64 MiB each way takes >= {64 * 1.048576 / 1e3 / float(t[2]) * 1e3:.2f} ms", flush=True)
if world > 1:
    dist.destroy_process_group()
