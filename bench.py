#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: CRS SpMV effective HBM GB/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (libsmb200, sm_100a kernels)
    python bench.py --impl reference [...]                         the reference's CPU path (oracle port)

A "step" is one pass of the hot path — y = A x, SparseMatrix::mvp (sparsematrix.rs:146-158) — over the
BASELINE.json configs[1] workload: the 3-D 7-point Dirichlet Laplacian 256^3, f32 values / u32 indices
(N = 16,777,216 rows, 117,047,296 non-zeros), synthetic, generated on the device.  With N > 1 GPUs every
rank owns a 256x256x256 z-slab of the 256x256x(256 N) operator (weak scaling, one halo plane per
neighbour exchanged per step over NCCL), launched one process per GPU by torch.distributed.run.

Reported per step: algorithmic bytes
    B = nnz*(sizeof T + sizeof I) + (n_rows+1)*sizeof I + n_cols*sizeof T + n_rows*sizeof T
`value`  = B_total * K / t with every input resident in HBM (CUDA events on the library's stream, max over ranks);
`e2e`    = the same through the host-buffer entry point smb200_spmv_host / vec_upload+dist_spmv+vec_download:
           pinned x H2D, SpMV, y D2H inside the timed region;
`roofline` = the SpMV kernel alone against MEASURED_PEAKS.json's HBM copy bandwidth;
`cpu_baseline` = the oracle's restatement of the reference (1 core = what the reference executes; all cores =
           our completion of its commented-out mvp_par) on the same workload on this box's host cores.
Inputs (1.14 GB per GPU) are ~9x the 126 MB L2, so no L2 flush is needed between timed steps.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NX = NY = 256
NZ_PER_GPU = 256
METRIC = "crs_spmv_effective_hbm_gbs"
UNIT = "GB/s"


def algorithmic_bytes(n_rows, n_cols, nnz, vsize, isize):
    return nnz * (vsize + isize) + (n_rows + 1) * isize + n_cols * vsize + n_rows * vsize


def laplace_nnz(nx, ny, nz):
    n = nx * ny * nz + 2 * (nx - 1) * ny * nz + 2 * nx * (ny - 1) * nz
    if nz > 1:
        n += 2 * nx * ny * (nz - 1)
    return n


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per SpMV launch from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def __enter__(self):
        if self.nv:
            def loop():
                while not self._stop.is_set():
                    self._once()
                    time.sleep(0.002)
            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference(steps: int, warmup: int, want_cg: bool):
    """The reference's CPU path through the oracle port (the Rust crate cannot be built here: no cargo/rustc).
    Each step = one SpMV over the full C2 workload.  Returns the cpu_baseline object and per-step seconds."""
    from oracle import oracle_py as orc
    cores = os.cpu_count() or 1
    vals, cols, offs = orc.laplace(np.float32, np.uint32, NX, NY, NZ_PER_GPU)
    n = NX * NY * NZ_PER_GPU
    x = orc.uniform(np.float32, 2, n)
    B = algorithmic_bytes(n, n, vals.size, 4, 4)

    def timed(threads, reps, warm):
        for _ in range(warm):
            orc.mvp(vals, cols, offs, x, threads=threads)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.mvp(vals, cols, offs, x, threads=threads)
            ts.append(time.perf_counter() - t0)
        return ts

    serial = timed(1, max(3, min(steps, 5)), 1)
    par = timed(cores, max(3, min(steps, 20)), max(1, min(warmup, 3)))
    out = {
        "value": B / np.mean(par) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"full workload (256^3 f32/u32, {vals.size} nnz), {len(par)} SpMVs with the thread-per-row-block completion of "
                  f"mvp_par (sparsemat_par.rs:39-67, an extension: the shipped reference is single-threaded)",
        "serial_value": B / np.mean(serial) / 1e9, "serial_cores": 1,
        "serial_sample": f"{len(serial)} SpMVs of the restated SparseMatCRS mvp on 1 core (what the reference executes)",
        "gflops": 2 * vals.size / np.mean(par) / 1e9, "serial_gflops": 2 * vals.size / np.mean(serial) / 1e9,
    }
    # the sparsemat_par path as shipped: SparseMatPar<SparseMatIndexList>, serial default mvp through the block dispatch
    # (sparsemat_par.rs:71-140) — bounded sample: the 128^3 operator in 16 row blocks, assembled through the restated API
    pn = 128
    _, psec, pasm = orc.par_laplace_mvp(np.float32, np.uint32, 16, pn, pn, pn, orc.uniform(np.float32, 2, pn ** 3), reps=3)
    pB = algorithmic_bytes(pn ** 3, pn ** 3, laplace_nnz(pn, pn, pn), 4, 4)
    out.update({"par_serial_value": pB / psec / 1e9, "par_serial_cores": 1, "par_assemble_s": pasm,
                "par_serial_sample": "3 products of the 128^3 f32/u32 Laplacian held as SparseMatPar<SparseMatIndexList> in 16 row "
                                     "blocks, serial default mvp through the block dispatch (what sparsemat_par.rs executes today); "
                                     "best of 3"})
    if want_cg:
        v64 = vals.astype(np.float64)
        b = orc.mvp(v64, cols, offs, orc.uniform(np.float64, 6, n))
        xs = np.zeros(n)
        cap = 5
        t0 = time.perf_counter()
        orc.cg(n, n, v64, cols, offs, b, xs, tol=1e-8, relative=True, iter_max=cap, threads=1)
        out["cg_iter_per_s_serial"] = cap / (time.perf_counter() - t0)
        out["cg_sample"] = f"{cap} iterations of the restated ConjugateGradient::solve on 256^3 f64/u32, 1 core (cap stated)"
    return out, float(np.mean(par))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    base, sec = cpu_reference(args.steps, args.warmup, want_cg=False)
    n = NX * NY * NZ_PER_GPU
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3-D 7-point Laplacian 256^3 f32/u32 CRS SpMV (BASELINE.json configs[1])", "n_rows": n,
                   "nnz": laplace_nnz(NX, NY, NZ_PER_GPU), "l2": "inputs larger than L2"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops": base["gflops"], "wall_s": time.perf_counter() - t0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    import sparsemat_b200 as smb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libsmb200 has no CPU fallback")
    torch.cuda.set_device(local)
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = smb.Context(local)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()

    vdt, idt = np.float32, np.uint32
    n_local = NX * NY * NZ_PER_GPU
    if multi:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(smb.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
        a = smb.DistCRS.laplace(ctx, vdt, idt, NX, NY, NZ_PER_GPU * world)
        d = a.dims()
        assert d["n_local"] == n_local
        nnz_local, n_ghost = d["nnz_local"], d["n_ghost"]
        x, y = a.new_vec(), a.new_vec()
        plan = a.local.plan_info()
        step = lambda: a.mvp(x, out=y)            # noqa: E731
    else:
        a = smb.SparseMatCRS.laplace(ctx, vdt, idt, NX, NY, NZ_PER_GPU)
        nnz_local, n_ghost = a.n_non_zero_entries(), 0
        x, y = smb.DenseVec(ctx, n_local, vdt), smb.DenseVec(ctx, n_local, vdt)
        plan = a.plan_info()
        step = lambda: a.mvp(x, out=y)            # noqa: E731
    x.fill_uniform(2 + rank)
    # algorithmic bytes of this rank's block: its own rows, its own slice of x (ghost planes are NVLink traffic,
    # reported separately, not counted)
    B_local = algorithmic_bytes(n_local, n_local, nnz_local, 4, 4)
    B_total = torch.tensor([float(B_local)], device="cuda", dtype=torch.float64)
    nnz_total = torch.tensor([float(nnz_local)], device="cuda", dtype=torch.float64)
    if multi:
        dist.all_reduce(B_total)
        dist.all_reduce(nnz_total)
    B_total, nnz_total = float(B_total.item()), float(nnz_total.item())

    # ---- device-resident timing: W warm-up steps, then exactly K timed steps -----------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev0, ev1 = ctx.event(), ctx.event()
    launches0 = smb.api.lib.smb200_launch_count()
    with ClockSampler(local) as clk:
        if multi:
            a.barrier()                # device-side alignment: every rank's stream starts the timed region together
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        clk._once()                    # the K launches are queued and running: one sample from this thread for certain
        ms = ev0.elapsed_ms(ev1)
        barrier()
    launches = smb.api.lib.smb200_launch_count() - launches0
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if multi:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = B_total / (ms_per_step * 1e-3) / 1e9

    # ---- end to end: pinned host x -> device, SpMV, y -> pinned host, every step -------------------------------
    e2e_steps = max(3, min(args.steps, 20))
    hx = smb.pinned_empty(n_local, vdt)
    hy = smb.pinned_empty(n_local, vdt)
    hx[:] = x.to_numpy()

    def e2e_step():
        if multi:
            x.upload(hx)
            a.mvp(x, out=y)
            smb.api.check(smb.api.lib.smb200_vec_download(y._h, smb.api.F.ptr(hy), hy.size))
        else:
            a.mvp_host(hx, hy)
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device="cuda", dtype=torch.float64)
    if multi:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = B_total / float(te.item()) / 1e9
    checksum = float(np.sum(hy.astype(np.float64)))

    # ---- the SpMV kernel alone (roofline): N = 1 -> identical to the timed region above -------------------------
    peak, peak_src = measured_peak()
    kernel_ms = ms_per_step if not multi else None
    if multi:
        # the local product of the interior+boundary launches, no exchange: time the three launches of one step
        loc = a.local
        xl = a.new_vec()
        xl.fill_uniform(99)
        yl = a.new_vec()
        for _ in range(3):
            a.mvp(xl, out=yl)
        barrier()
        k0, k1 = ctx.event(), ctx.event()
        k0.record()
        for _ in range(20):
            a.mvp(xl, out=yl)
        k1.record()
        kernel_ms = k0.elapsed_ms(k1) / 20
        del loc
    roof = {"bound": "hbm", "achieved": B_local / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": B_local / (kernel_ms * 1e-3) / 1e9 / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
            "frac_of_nominal_8tbs": B_local / (kernel_ms * 1e-3) / 1e9 / 8000.0, "kernel": f"spmv_{plan['variant_name']}",
            "algorithmic_bytes_per_launch": B_local, "kernel_ms": kernel_ms,
            # plan-time index compression (16-bit window positions instead of the u32 column array): what the kernel
            # really streams; `achieved` above counts the ALGORITHMIC bytes of the CRS format
            "streamed_bytes_per_launch": int(plan["stream_bytes"]), "nnz_with_16bit_columns": int(plan["nnz_c16"]),
            "streamed_gbs": plan["stream_bytes"] / (kernel_ms * 1e-3) / 1e9,
            "frac_streamed": plan["stream_bytes"] / (kernel_ms * 1e-3) / 1e9 / peak}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "3-D 7-point Laplacian 256^3 f32/u32 CRS SpMV per GPU (BASELINE.json configs[1]); "
                               f"global grid 256x256x{NZ_PER_GPU * world}, 1-D z-slab row blocks",
                   "n_rows": n_local * world, "nnz": int(nnz_total), "bytes_per_step": int(B_total),
                   "parallelism": f"rowblock{world}", "l2": "inputs (1.14 GB/GPU) larger than the 126 MB L2; no flush",
                   "kernel": plan["variant_name"], "halo_elems_per_gpu": int(n_ghost)},
        "gflops": 2.0 * nnz_total / (ms_per_step * 1e-3) / 1e9,
        "frac_of_measured_hbm_peak": value / (peak * world), "frac_of_nominal_8tbs": value / (8000.0 * world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_local * 4 * world),
                "d2h_bytes_per_step": int(n_local * 4 * world), "steps": e2e_steps, "ms_per_step": float(te.item()) * 1e3,
                "checksum_y": checksum},
        "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof,
    }

    # ---- extras at N = 1 on rank 0: CG iter/s (config 4) and the CPU baseline ------------------------------------
    if not multi and not args.no_cg:
        del a, x, y
        a64 = smb.SparseMatCRS.laplace(ctx, np.float64, idt, NX, NY, NZ_PER_GPU)
        xs = smb.DenseVec(ctx, n_local, np.float64)
        xs.fill_uniform(6)
        b = a64.mvp(xs)
        x0 = smb.DenseVec(ctx, n_local, np.float64)
        smb.ConjugateGradient(1e-30, 17).solve_with_stats(a64, b, smb.DenseVec(ctx, n_local, np.float64))   # untimed: graph capture
        st = smb.ConjugateGradient(1e-8, 5000, relative=True).solve_with_stats(a64, b, x0)
        Bcg = algorithmic_bytes(n_local, n_local, nnz_local, 8, 4) + 9 * n_local * 8
        its = max(1, int(st["iterations"]))
        line["cg"] = {"workload": "CG on the 256^3 f64/u32 Laplacian to 1e-8 relative residual (BASELINE.json configs[3])",
                      "iterations": its, "converged": st["converged"], "final_residual": st["final_residual"],
                      "device_ms": st["device_ms"], "iter_per_s": its / (st["device_ms"] * 1e-3),
                      "effective_gbs": Bcg * its / (st["device_ms"] * 1e-3) / 1e9,
                      "frac_of_measured_hbm_peak": Bcg * its / (st["device_ms"] * 1e-3) / 1e9 / peak,
                      "launches": int(st["launches"])}
    if multi and not args.no_cg:
        # CG iter/s at N GPUs (BASELINE.json metric): weak scaling, every rank a 256^3 f64/u32 slab; fixed iteration count
        del a, x, y
        a64 = smb.DistCRS.laplace(ctx, np.float64, idt, NX, NY, NZ_PER_GPU * world)
        xs = a64.new_vec()
        xs.fill_uniform(6 + rank)
        b = a64.mvp(xs)
        x0 = a64.new_vec()
        cg_iters = 200
        # one short untimed solve first: NCCL connects its all-reduce channels lazily and the iteration graph is captured once
        smb.ConjugateGradient(1e-30, 17).solve_with_stats(a64, b, a64.new_vec())
        barrier()
        st = smb.ConjugateGradient(1e-30, cg_iters).solve_with_stats(a64, b, x0)
        tms = torch.tensor([st["device_ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        Bcg = (algorithmic_bytes(n_local, n_local, a64.dims()["nnz_local"], 8, 4) + 9 * n_local * 8) * world
        line["cg"] = {"workload": f"{cg_iters} CG iterations on the 256x256x{NZ_PER_GPU * world} f64/u32 Laplacian, z-slab row blocks "
                                  "(BASELINE.json configs[3] per GPU, weak scaling)",
                      "iterations": int(st["iterations"]), "device_ms": float(tms.item()),
                      "iter_per_s": int(st["iterations"]) / (float(tms.item()) * 1e-3),
                      "effective_gbs": Bcg * int(st["iterations"]) / (float(tms.item()) * 1e-3) / 1e9,
                      "frac_of_measured_hbm_peak": Bcg * int(st["iterations"]) / (float(tms.item()) * 1e-3) / 1e9 / (peak * world)}
    if not multi and not args.no_cpu:
        base, _ = cpu_reference(5, 1, want_cg=not args.no_cg)
        line["cpu_baseline"] = base
    elif rank == 0:
        line["cpu_baseline"] = None if multi else {"skipped": "--no-cpu"}
    if rank == 0:
        emit(line)
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL banners, warnings) was
    diverted to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                       # libraries that print to fd 1 (e.g. "NCCL version ...") land on stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cg", action="store_true", help="skip the CG (config 4) extra")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
