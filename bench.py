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
`value`  = B_total * K / t with every input resident in HBM (CUDA events on the library's stream, max over ranks; at
           N > 1 the timed region starts behind a device-side barrier so that it does not depend on the step count);
`e2e`    = the same through the host-buffer entry point smb200_spmv_host / vec_upload+dist_spmv+vec_download:
           pinned x H2D, SpMV, y D2H inside the timed region;
`roofline` = the SpMV kernel alone against MEASURED_PEAKS.json's HBM copy bandwidth: `frac` counts the bytes the kernel
           physically streams (16-bit plan-time columns/offsets), `frac_effective` the algorithmic CRS bytes;
`parity_checked` = the y that came back in the e2e leg equals the oracle's mvp on the same x bit for bit (every rank
           checks its own row block against the oracle's rows of the global operator);
`cpu_baseline` = the oracle's restatement of the reference (1 core = what the reference executes; all cores =
           our completion of its commented-out mvp_par) on the same workload on this box's host cores;
`extras` (N = 1) = the other BASELINE.json configs: C1 (1024^2 f64/u32 via IndexList -> to_crs), C3 (power-law 50 M
           rows u64), C5's matrix on one GPU (512^3), each checked against the oracle.
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
def workload_config(world: int) -> dict:
    """The `config` object: a pure function of the GPU count, identical in both arms."""
    nz = NZ_PER_GPU * world
    n_local = NX * NY * NZ_PER_GPU
    nnz = laplace_nnz(NX, NY, nz)
    b_total = nnz * 8 + world * (n_local + 1) * 4 + 2 * world * n_local * 4
    return {"workload": "3-D 7-point Laplacian 256^3 f32/u32 CRS SpMV per GPU (BASELINE.json configs[1]); "
                        f"global grid 256x256x{nz}, 1-D z-slab row blocks",
            "n_rows": n_local * world, "nnz": nnz, "bytes_per_step": b_total, "parallelism": f"rowblock{world}",
            "l2": "inputs (1.14 GB/GPU) larger than the 126 MB L2; no flush",
            "halo_elems_per_interior_gpu": 0 if world == 1 else NX * NY * (2 if world > 2 else 1)}


def oracle_slab(rank: int, world: int, threads: int):
    """Oracle side of the parity check: rows [lo, hi) of the global 256x256x(256 world) operator applied to the global x
    (rank q's slice = uniform(seed 2 + q)).  Returns (x_local, y_local)."""
    from oracle import oracle_py as orc
    plane, n_local = NX * NY, NX * NY * NZ_PER_GPU
    lo, hi = rank * n_local, (rank + 1) * n_local
    vals, cols, offs = orc.laplace(np.float32, np.uint32, NX, NY, NZ_PER_GPU * world, lo, hi)
    x_own = orc.uniform(np.float32, 2 + rank, n_local)
    parts, base = [], lo
    if rank > 0:
        parts.append(orc.uniform(np.float32, 2 + rank - 1, n_local)[-plane:])
        base = lo - plane
    parts.append(x_own)
    if rank + 1 < world:
        parts.append(orc.uniform(np.float32, 2 + rank + 1, n_local)[:plane])
    x_ext = np.ascontiguousarray(np.concatenate(parts))
    cols_local = (cols.astype(np.int64) - base).astype(np.uint32)
    return x_own, orc.mvp(vals, cols_local, offs, x_ext, threads=threads)


def cpu_reference(steps: int, warmup: int, want_cg: bool):
    """The reference's CPU path through the oracle port (the Rust crate cannot be built here: no cargo/rustc).
    Each step = one SpMV over one GPU's share of the workload (the 256^3 operator).  Returns the cpu_baseline object,
    per-step seconds, and (x, y) of the oracle for the parity check."""
    from oracle import oracle_py as orc
    cores = os.cpu_count() or 1
    vals, cols, offs = orc.laplace(np.float32, np.uint32, NX, NY, NZ_PER_GPU)
    n = NX * NY * NZ_PER_GPU
    x = orc.uniform(np.float32, 2, n)
    B = algorithmic_bytes(n, n, vals.size, 4, 4)
    y_keep = [None]

    def timed(threads, reps, warm):
        for _ in range(warm):
            orc.mvp(vals, cols, offs, x, threads=threads)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            y_keep[0] = orc.mvp(vals, cols, offs, x, threads=threads)
            ts.append(time.perf_counter() - t0)
        return ts

    serial = timed(1, max(3, min(steps, 5)), 1)
    par = timed(cores, max(3, min(steps, 20)), max(1, min(warmup, 3)))
    out = {
        "value": B / np.mean(par) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"one GPU's share of the workload (256^3 f32/u32, {vals.size} nnz), {len(par)} SpMVs with the thread-per-row-block "
                  f"completion of mvp_par (sparsemat_par.rs:39-67, an extension: the shipped reference is single-threaded, "
                  f"see serial_value)",
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
    return out, float(np.mean(par)), (x, y_keep[0])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    t0 = time.perf_counter()
    base, sec, _ = cpu_reference(args.steps, args.warmup, want_cg=False)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gflops": base["gflops"], "wall_s": time.perf_counter() - t0,
        "note": "value = all host cores on the builder's threaded completion of the reference's commented-out mvp_par; the "
                "reference as shipped is single-threaded: cpu_baseline.serial_value",
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------
def time_spmv(smb, ctx, a, x, y, steps, warm=3, flush=False):
    """Mean ms of one SpMV (CUDA events on the library stream); flush=True evicts L2 before every product and
    times each product on its own."""
    for _ in range(warm):
        a.mvp(x, out=y)
    ctx.sync()
    if not flush:
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        for _ in range(steps):
            a.mvp(x, out=y)
        e1.record()
        return e0.elapsed_ms(e1) / steps
    ts = []
    for _ in range(steps):
        ctx.flush_l2()
        e0, e1 = ctx.event(), ctx.event()
        e0.record()
        a.mvp(x, out=y)
        e1.record()
        ts.append(e0.elapsed_ms(e1))
    return float(np.mean(ts))


def extras_single_gpu(smb, ctx, peak):
    """The other BASELINE.json configs on one GPU, each checked against the oracle (bounded: a few seconds each)."""
    from oracle import oracle_py as orc
    out = {}
    # ---- C1: 2-D 5-point Poisson 1024^2 f64/u32, assembled through SparseMatIndexList -> to_crs (configs[0]) -----------
    try:
        nx = 1024
        t0 = time.perf_counter()
        vals, cols, offs = orc.laplace(np.float64, np.uint32, nx, nx, 1)
        il = smb.SparseMatIndexList(np.float64, np.uint32)
        o64 = offs.astype(np.int64)
        rows = np.repeat(np.arange(nx * nx, dtype=np.uint64), (o64[1:] - o64[:-1]))
        il.set_many(rows, cols.astype(np.uint64), vals)             # `set` per entry, row-major, ascending columns
        t_asm = time.perf_counter() - t0
        t0 = time.perf_counter()
        a = il.to_crs(ctx)
        ctx.sync()
        t_conv = time.perf_counter() - t0
        gv, gc, go = a.raw_parts()
        layout_ok = bool(np.array_equal(gv, vals) and np.array_equal(gc, cols) and np.array_equal(go, offs))
        xh = orc.uniform(np.float64, 1, nx * nx)
        x, y = smb.DenseVec.from_vec(ctx, xh), smb.DenseVec(ctx, nx * nx, np.float64)
        warm = time_spmv(smb, ctx, a, x, y, 50)
        cold = time_spmv(smb, ctx, a, x, y, 10, flush=True)
        want = orc.mvp(vals, cols, offs, xh)
        B = algorithmic_bytes(nx * nx, nx * nx, vals.size, 8, 4)
        out["c1"] = {"workload": "2-D 5-point Laplacian 1024^2 f64/u32 via SparseMatIndexList -> to_crs (configs[0])",
                     "to_crs_layout_bit_exact": layout_ok, "y_bit_exact": bool(np.array_equal(y.to_numpy(), want)),
                     "kernel": a.plan_info()["variant_name"], "ms_warm_l2": warm, "ms_cold_l2_flushed": cold,
                     "gbs_warm": B / warm / 1e6, "gbs_cold": B / cold / 1e6, "frac_cold_of_measured_peak": B / cold / 1e6 / peak,
                     "assemble_host_s": t_asm, "to_crs_device_s": t_conv, "algorithmic_bytes": B,
                     "note": "84 MB fits the 126 MB L2: warm reads L2, cold is after an L2 flush"}
        del a, x, y, il
    except Exception as e:  # noqa: BLE001
        out["c1"] = {"error": repr(e)}
    # ---- C5's matrix on ONE GPU: 512^3 f32/u32 (the 8-GPU strong-scaling base) --------------------------------------------
    try:
        n5 = 512
        a = smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, n5, n5, n5)
        x, y = smb.DenseVec(ctx, n5 ** 3, np.float32), smb.DenseVec(ctx, n5 ** 3, np.float32)
        x.fill_uniform(2)
        ms = time_spmv(smb, ctx, a, x, y, 20)
        # oracle on a slab in the middle: rows of z-planes [255, 258), needs x planes [254, 259)
        plane = n5 * n5
        lo, hi = 255 * plane, 258 * plane
        vals, cols, offs = orc.laplace(np.float32, np.uint32, n5, n5, n5, lo, hi)
        xh = orc.uniform(np.float32, 2, n5 ** 3)
        want = orc.mvp(vals, (cols.astype(np.int64) - (lo - plane)).astype(np.uint32), offs,
                       np.ascontiguousarray(xh[lo - plane:hi + plane]), threads=os.cpu_count() or 1)
        got = y.to_numpy()[lo:hi]
        B = algorithmic_bytes(n5 ** 3, n5 ** 3, a.n_non_zero_entries(), 4, 4)
        out["c5_single_gpu"] = {"workload": "3-D 7-point Laplacian 512^3 f32/u32 on one GPU (configs[4] at P = 1)", "ms": ms,
                                "gbs": B / ms / 1e6, "frac_effective": B / ms / 1e6 / peak, "gflops": 2 * a.n_non_zero_entries() / ms / 1e6,
                                "kernel": a.plan_info()["variant_name"], "algorithmic_bytes": B,
                                "y_bit_exact_on_sample": bool(np.array_equal(got, want)),
                                "sample": "3 z-planes (786,432 rows) in the middle of the grid against the oracle's rows"}
        del a, x, y, xh, got, want
    except Exception as e:  # noqa: BLE001
        out["c5_single_gpu"] = {"error": repr(e)}
    # ---- C3: power-law rows, N = 50 M, u64 indices (configs[2]) --------------------------------------------------------------
    try:
        n3 = 50_000_000
        a = smb.SparseMatCRS.powerlaw(ctx, np.float64, np.uint64, n3)
        x, y = smb.DenseVec(ctx, n3, np.float64), smb.DenseVec(ctx, n3, np.float64)
        x.fill_uniform(7)
        t0 = time.perf_counter()
        pi = a.plan_info()
        ms = time_spmv(smb, ctx, a, x, y, 5, warm=2)
        nnz = a.n_non_zero_entries()
        B = algorithmic_bytes(n3, n3, nnz, 8, 8)
        # parity: 1 M sampled rows incl. the 1000 longest, regenerated from the seeds by the oracle
        lens = orc.powerlaw_row_lens(n3)
        rng = np.random.default_rng(12345)
        longest = np.argpartition(lens, n3 - 1000)[n3 - 1000:]
        rows = np.unique(np.concatenate([longest, rng.integers(0, n3, 1_000_000)])).astype(np.uint64)
        xh = x.to_numpy()
        yo, ab, _ = orc.powerlaw_sample_mvp(np.float64, n3, rows, xh)
        yg = y.to_numpy()[rows.astype(np.int64)]
        err = np.abs(yg.astype(np.float64) - yo.astype(np.float64))
        rel = float(np.max(err / np.maximum(ab, 1e-300)))
        out["c3"] = {"workload": "power-law rows N = 50 M, mean 16 nnz/row, f64/u64 (configs[2])", "nnz": nnz, "ms": ms,
                     "gbs": B / ms / 1e6, "frac_effective": B / ms / 1e6 / peak, "gflops": 2 * nnz / ms / 1e6,
                     "kernel": pi["variant_name"], "launches_per_spmv": int(pi["launches_per_spmv"]),
                     "streamed_bytes": int(pi["stream_bytes"]), "frac_streamed": pi["stream_bytes"] / ms / 1e6 / peak,
                     "plan_ms": pi["plan_ms"], "plan_bytes": int(pi["plan_bytes"]), "algorithmic_bytes": B,
                     "parity_rows_checked": int(rows.size), "parity_max_err_over_abs_rowsum": rel,
                     "parity_ok": bool(rel <= 1e-12), "parity_bit_exact_rows": int(np.sum(yg == yo)),
                     "x_seed_matches_oracle": bool(np.array_equal(xh, orc.uniform(np.float64, 7, n3)))}
        del a, x, y, xh, lens
    except Exception as e:  # noqa: BLE001
        out["c3"] = {"error": repr(e)}
    return out


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
    dist_info = None
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
        dist_info = a.info()
    else:
        a = smb.SparseMatCRS.laplace(ctx, vdt, idt, NX, NY, NZ_PER_GPU)
        nnz_local, n_ghost = a.n_non_zero_entries(), 0
        x, y = smb.DenseVec(ctx, n_local, vdt), smb.DenseVec(ctx, n_local, vdt)
        plan = a.plan_info()
    step = lambda: a.mvp(x, out=y)            # noqa: E731
    x.fill_uniform(2 + rank)
    cfg = workload_config(world)
    # algorithmic bytes of this rank's block: its own rows, its own slice of x (ghost planes are NVLink traffic,
    # reported separately, not counted)
    B_local = algorithmic_bytes(n_local, n_local, nnz_local, 4, 4)
    B_total = torch.tensor([float(B_local)], device="cuda", dtype=torch.float64)
    nnz_total = torch.tensor([float(nnz_local)], device="cuda", dtype=torch.float64)
    if multi:
        dist.all_reduce(B_total)
        dist.all_reduce(nnz_total)
    B_total, nnz_total = float(B_total.item()), float(nnz_total.item())
    assert int(B_total) == cfg["bytes_per_step"] and int(nnz_total) == cfg["nnz"], (B_total, nnz_total, cfg)

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
    launches = smb.api.lib.smb200_launch_count() - launches0 - (1 if multi else 0)     # minus the barrier kernel
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

    # ---- parity at full size: the y of the e2e leg against the oracle's mvp on the same x, bit for bit ------------------
    parity = None
    base = None
    if not args.no_parity:
        cores = os.cpu_count() or 1
        if multi:
            xo, yo = oracle_slab(rank, world, max(1, cores // world))
        else:
            # N = 1: the cpu_baseline leg computes the oracle's y on this very x; reuse it instead of throwing it away
            base, _, (xo, yo) = cpu_reference(5, 1, want_cg=not args.no_cg) if not args.no_cpu else (None, None, oracle_slab(0, 1, cores))
        ok = torch.tensor([1.0 if (np.array_equal(hx, xo) and np.array_equal(hy, yo)) else 0.0], device="cuda", dtype=torch.float64)
        if multi:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        parity = {"checked": True, "y_bit_exact_all_ranks": bool(ok.item() == 1.0), "rows_per_rank": n_local,
                  "oracle": "oracle mvp (storage-order row sums) on the rank's rows of the global operator, x = uniform(seed 2 + rank)",
                  "checksum_y_rank0": checksum, "checksum_y_oracle_rank0": float(np.sum(yo.astype(np.float64)))}

    # ---- the SpMV kernel alone (roofline) ---------------------------------------------------------------------------
    # One launch per product at any N (at N > 1 the same launch also pushes / awaits the halo planes), so the kernel's mean
    # duration is the step time of the timed region above, measured with CUDA events on the launching stream.
    peak, peak_src = measured_peak()
    kernel_ms = ms_per_step
    phys = int(plan["stream_bytes"])
    roof = {"bound": "hbm", "achieved": B_local / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            # frac: bytes the kernel physically streams (plan-time compression: 16-bit window positions instead of columns,
            # 8-bit value codes + per-block dictionaries instead of values, a length byte per row instead of an offset; ncu DRAM
            # traffic: `traffic`) / time / peak.  frac_effective: the ALGORITHMIC bytes of the CRS format / time /
            # peak — above 1 because the kernel moves fewer bytes than the format holds.
            "frac": phys / (kernel_ms * 1e-3) / 1e9 / peak, "frac_effective": B_local / (kernel_ms * 1e-3) / 1e9 / peak,
            "achieved_physical": phys / (kernel_ms * 1e-3) / 1e9,
            "traffic": ncu_traffic(), "peak_source": peak_src,
            "frac_of_nominal_8tbs": phys / (kernel_ms * 1e-3) / 1e9 / 8000.0,
            "kernel": f"spmv_{plan['variant_name']}" + ("_sell" if plan["sell_entries"] else "") +
                      (" (one launch per product incl. the in-kernel halo push / wait)" if multi else ""),
            "algorithmic_bytes_per_launch": B_local, "physical_bytes_per_launch": phys, "kernel_ms": kernel_ms,
            "nnz_with_16bit_columns": int(plan["nnz_c16"]), "rows_with_16bit_offsets": int(plan["rows_o16"]),
            # plan-time value indexing (8-bit codes into per-block dictionaries of <= 256 distinct values; the 7-point Laplacian
            # has two) and the sliced-ELLPACK stage order of the compressed entries (padded entries counted in physical bytes)
            "nnz_with_8bit_value_codes": int(plan["nnz_v8"]), "sell_padded_entries": int(plan["sell_entries"]),
            # the plan is built once per matrix, outside the timed region: its cost, stated
            "plan_ms": plan["plan_ms"], "plan_extra_bytes": int(plan["plan_bytes"]),
            "plan_extra_frac_of_crs": plan["plan_bytes"] / float(B_local)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "gflops": 2.0 * nnz_total / (ms_per_step * 1e-3) / 1e9,
        "frac_of_measured_hbm_peak": value / (peak * world), "frac_of_nominal_8tbs": value / (8000.0 * world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n_local * 4 * world),
                "d2h_bytes_per_step": int(n_local * 4 * world), "steps": e2e_steps, "ms_per_step": float(te.item()) * 1e3,
                "checksum_y": checksum},
        "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof,
        "parity_checked": bool(parity and parity["y_bit_exact_all_ranks"]), "parity": parity,
        "timed_region": "device-side barrier, K products, CUDA events on the library stream, max over ranks" if multi
                        else "K products, CUDA events on the library stream",
    }
    if multi:
        di = a.info() if a is not None else dist_info
        line["dist"] = {"halo_wait_avg_us_rank0": (di["wait_ns_total"] / di["waits"] / 1e3) if di["waits"] else 0.0,
                        "halo_waits_per_product_rank0": di["waits"] / max(1, di["products"]),
                        "data_path": "peer memory over NVLink (CUDA IPC): halo planes stored by the product kernel, no NCCL call per step"
                                     if dist_info["p2p"] else "NCCL send/recv fallback",
                        "neighbours_rank0": dist_info["neighbours"], "halo_elems_rank0": int(n_ghost)}

    # ---- extras at N = 1 on rank 0: CG iter/s (config 4), the other configs, the CPU baseline ------------------------
    if not multi and not args.no_cg:
        del a, x, y
        a64 = smb.SparseMatCRS.laplace(ctx, np.float64, idt, NX, NY, NZ_PER_GPU)
        xs = smb.DenseVec(ctx, n_local, np.float64)
        xs.fill_uniform(6)
        b = a64.mvp(xs)
        x0 = smb.DenseVec(ctx, n_local, np.float64)
        # untimed first solve with the same operands and limits: allocates the workspace and captures the iteration graph
        # (keyed by x, the plan and the batch), so the timed solve replays it
        smb.ConjugateGradient(1e-1, 5000, relative=True).solve_with_stats(a64, b, x0)
        x0.fill(0.0)
        st = smb.ConjugateGradient(1e-8, 5000, relative=True).solve_with_stats(a64, b, x0)
        Bcg = algorithmic_bytes(n_local, n_local, nnz_local, 8, 4) + 9 * n_local * 8
        its = max(1, int(st["iterations"]))
        line["cg"] = {"workload": "CG on the 256^3 f64/u32 Laplacian to 1e-8 relative residual (BASELINE.json configs[3])",
                      "iterations": its, "converged": st["converged"], "final_residual": st["final_residual"],
                      "device_ms": st["device_ms"], "iter_per_s": its / (st["device_ms"] * 1e-3),
                      "effective_gbs": Bcg * its / (st["device_ms"] * 1e-3) / 1e9,
                      "frac_of_measured_hbm_peak": Bcg * its / (st["device_ms"] * 1e-3) / 1e9 / peak,
                      "launches": int(st["launches"])}
        if not args.no_parity:
            # true residual of the GPU's solution, recomputed by the oracle's CPU mvp (CG parity (c))
            from oracle import oracle_py as orc
            v64, c64, o64 = orc.laplace(np.float64, np.uint32, NX, NY, NZ_PER_GPU)
            bh, xh = b.to_numpy(), x0.to_numpy()
            r = bh - orc.mvp(v64, c64, o64, xh, threads=os.cpu_count() or 1)
            true_rel = float(np.linalg.norm(r) / np.linalg.norm(bh))
            line["cg"].update({"true_relative_residual_oracle": true_rel, "true_residual_checked": bool(true_rel <= 1.05e-8)})
            del v64, c64, o64
        del a64, xs, b, x0
    if multi and not args.no_cg:
        # CG iter/s at N GPUs (BASELINE.json metric): weak scaling, every rank a 256^3 f64/u32 slab; fixed iteration count
        del a, x, y
        a64 = smb.DistCRS.laplace(ctx, np.float64, idt, NX, NY, NZ_PER_GPU * world)
        xs = a64.new_vec()
        xs.fill_uniform(6 + rank)
        b = a64.mvp(xs)
        x0 = a64.new_vec()
        cg_iters = 200
        # one short untimed solve first with the same operands and limits: workspace allocated, iteration graph captured
        smb.ConjugateGradient(1e-1, cg_iters, relative=True).solve_with_stats(a64, b, x0)
        x0.fill(0.0)
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
                      "frac_of_measured_hbm_peak": Bcg * int(st["iterations"]) / (float(tms.item()) * 1e-3) / 1e9 / (peak * world),
                      "final_residual": st["final_residual"]}
        # the same iterations through the single-reduction loop (one all-reduce per iteration, smb200_dist_cg_solve_sr)
        x0.fill(0.0)
        smb.ConjugateGradient(1e-1, cg_iters, relative=True, single_reduce=True).solve_with_stats(a64, b, x0)
        x0.fill(0.0)
        barrier()
        st = smb.ConjugateGradient(1e-30, cg_iters, single_reduce=True).solve_with_stats(a64, b, x0)
        tms = torch.tensor([st["device_ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        line["cg"]["single_reduce"] = {"iterations": int(st["iterations"]), "device_ms": float(tms.item()),
                                       "iter_per_s": int(st["iterations"]) / (float(tms.item()) * 1e-3),
                                       "final_residual": st["final_residual"]}
    if not multi and not args.no_extras:
        a = x = y = None               # (already gone if the CG leg ran) free the C2 matrix before the larger extras
        line["extras"] = extras_single_gpu(smb, ctx, peak)
    if not multi:
        if args.no_cpu:
            line["cpu_baseline"] = {"skipped": "--no-cpu"}
        else:
            if base is None:
                base, _, _ = cpu_reference(5, 1, want_cg=not args.no_cg)
            line["cpu_baseline"] = base
    elif rank == 0:
        line["cpu_baseline"] = None
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
    ap.add_argument("--no-parity", action="store_true", help="skip the full-size oracle comparison")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1 / C3 / C5 extras (N = 1)")
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
