import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import sparsemat_b200 as smb
from test_gpu_cg_tocrs import _laplace2d_entries
ctx = smb.Context(0)
i, j, v = _laplace2d_entries(1024, 1024)
sp = smb.SparseMatIndexList(np.float64, np.uint32)
sp.set(i, j, v)
cols, vals, pos, nxt = sp.raw_arrays()
for rep in range(4):
    ctx.sync(); t0 = time.perf_counter()
    a = smb.crs_from_indexlist_arrays(ctx, sp.n_rows(), sp.n_cols(), cols, vals, pos, nxt)
    ctx.sync(); t1 = time.perf_counter()
    print(f"rep {rep}: crs_from_indexlist {1e3*(t1-t0):.1f} ms", flush=True)
    t0 = time.perf_counter(); a2 = sp.to_crs(ctx); ctx.sync(); t1 = time.perf_counter()
    print(f"rep {rep}: il_to_crs {1e3*(t1-t0):.1f} ms", flush=True)
