#!/usr/bin/env python
"""GPU: what the host link can do — H2D alone, D2H alone, both at once (pinned memory, two streams), 64 MiB each."""
import time
import torch
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(label, do_in, do_out, reps=20, pieces=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step = n // pieces
        for k in range(pieces):
            sl = slice(k * step, (k + 1) * step)
            if do_in:
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    gb = (n * (do_in + do_out)) / dt / 1e9
    print(f"{label}: {dt * 1e3:.3f} ms per round, {gb:.1f} GB/s total", flush=True)


run("H2D 64 MiB", True, False)
run("D2H 64 MiB", False, True)
run("H2D + D2H concurrently", True, True)
run("H2D + D2H concurrently, 32 pieces", True, True, pieces=32)
