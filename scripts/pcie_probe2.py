#!/usr/bin/env python
"""GPU: host link with warm-up — H2D from default-pinned vs write-combined pinned memory, piece sizes, two H2D streams,
and the e2e entry point (smb200_spmv_host) per chunk count.  CUDA events, 64 MiB."""
import ctypes as C
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch  # noqa: F401  (loads libcudart)

rt = None
for pat in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = C.CDLL(pat)
        break
    except OSError:
        pass
if rt is None:
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    rt = C.CDLL(cands[0])
vp = C.c_void_p
rt.cudaHostAlloc.argtypes = [C.POINTER(vp), C.c_size_t, C.c_uint]
rt.cudaMalloc.argtypes = [C.POINTER(vp), C.c_size_t]
rt.cudaMemcpyAsync.argtypes = [vp, vp, C.c_size_t, C.c_int, vp]
rt.cudaStreamCreate.argtypes = [C.POINTER(vp)]
rt.cudaEventCreate.argtypes = [C.POINTER(vp)]
rt.cudaEventRecord.argtypes = [vp, vp]
rt.cudaEventSynchronize.argtypes = [vp]
rt.cudaStreamWaitEvent.argtypes = [vp, vp, C.c_uint]
rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), vp, vp]
rt.cudaStreamSynchronize.argtypes = [vp]
N = 64 << 20


def chk(e):
    assert e == 0, e


def halloc(flags):
    p = vp()
    chk(rt.cudaHostAlloc(C.byref(p), N, flags))
    C.memset(p, 1, N)
    return p


def mk(fn):
    p = vp()
    chk(fn(C.byref(p)))
    return p


h_def, h_wc, h_out = halloc(0), halloc(4), halloc(0)
d_a, d_b = vp(), vp()
chk(rt.cudaMalloc(C.byref(d_a), N))
chk(rt.cudaMalloc(C.byref(d_b), N))
s1, s2, s3 = mk(rt.cudaStreamCreate), mk(rt.cudaStreamCreate), mk(rt.cudaStreamCreate)
e0, e1, e2 = mk(rt.cudaEventCreate), mk(rt.cudaEventCreate), mk(rt.cudaEventCreate)


def timed(label, body, reps=20):
    for _ in range(3):
        body()
    for s in (s1, s2, s3):
        chk(rt.cudaStreamSynchronize(s))
    best, tot = 1e9, 0.0
    for _ in range(reps):
        chk(rt.cudaEventRecord(e0, s1))
        chk(rt.cudaStreamWaitEvent(s2, e0, 0))
        chk(rt.cudaStreamWaitEvent(s3, e0, 0))
        body()
        chk(rt.cudaEventRecord(e1, s2))
        chk(rt.cudaEventRecord(e2, s3))
        chk(rt.cudaStreamWaitEvent(s1, e1, 0))
        chk(rt.cudaStreamWaitEvent(s1, e2, 0))
        chk(rt.cudaEventRecord(e1, s1))
        chk(rt.cudaEventSynchronize(e1))
        ms = C.c_float()
        chk(rt.cudaEventElapsedTime(C.byref(ms), e0, e1))
        best = min(best, ms.value)
        tot += ms.value
    print(f"{label:58s} mean {tot / reps:7.3f} ms  best {best:7.3f} ms  ({N / best / 1e6:6.1f} GB/s one way)", flush=True)


def h2d(src, pieces=1, stream=None, two=False):
    def body():
        step = N // pieces
        for k in range(pieces):
            st = stream or s1
            if two and (k & 1):
                st = s2
            chk(rt.cudaMemcpyAsync(d_a.value + k * step, src.value + k * step, step, 1, st))
    return body


def d2h(pieces=1):
    def body():
        step = N // pieces
        for k in range(pieces):
            chk(rt.cudaMemcpyAsync(h_out.value + k * step, d_b.value + k * step, step, 2, s3))
    return body


timed("H2D default pinned", h2d(h_def))
timed("H2D write-combined pinned", h2d(h_wc))
timed("H2D default pinned, 8 pieces", h2d(h_def, 8))
timed("H2D default pinned, 8 pieces on two streams", h2d(h_def, 8, two=True))
timed("H2D default pinned, 2 halves on two streams", h2d(h_def, 2, two=True))
timed("D2H", d2h())
timed("D2H 8 pieces", d2h(8))
both = lambda: (h2d(h_def)(), d2h()())
timed("H2D + D2H concurrently", both)
both8 = lambda: (h2d(h_def, 8)(), d2h(8)())
timed("H2D + D2H concurrently, 8 pieces each", both8)
bothwc = lambda: (h2d(h_wc)(), d2h()())
timed("H2D (write-combined) + D2H concurrently", bothwc)

# ---- the library's end-to-end entry point on C2, per chunk count --------------------------------------------------------
import time
import sparsemat_b200 as smb  # noqa: E402
ctx = smb.Context(0)
for chunks in (0, 2, 4, 8, 16):
    os.environ["SMB200_HOST_CHUNKS"] = str(chunks)
    a = smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, 256, 256, 256)
    n = a.n_rows()
    hx, hy = smb.pinned_empty(n, np.float32), smb.pinned_empty(n, np.float32)
    hx[:] = 1.0
    for _ in range(3):
        a.mvp_host(hx, hy)
    t0 = time.perf_counter()
    for _ in range(20):
        a.mvp_host(hx, hy)
    dt = (time.perf_counter() - t0) / 20
    print(f"smb200_spmv_host C2, chunks={chunks or 'auto'}: {dt * 1e3:.3f} ms/step", flush=True)
    del a
