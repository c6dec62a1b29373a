#!/usr/bin/env python
"""GPU (one device): what does the halo code inside the ring kernel cost by itself?  The 256^3 f32/u32 slab through
(a) the plain product, (b) the distributed kernel path with zero neighbours (SMB200_DIST_SELF=1)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SMB200_DIST_SELF"] = "1"
import sparsemat_b200 as smb  # noqa: E402

ctx = smb.Context(0)
ctx.comm_init(0, 1, None)
n = 256
for rep in range(1):
    a = smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, n, n, n)
    d = smb.DistCRS.laplace(ctx, np.float32, np.uint32, n, n, n)
    x, y = smb.DenseVec(ctx, n ** 3, np.float32), smb.DenseVec(ctx, n ** 3, np.float32)
    xd, yd = d.new_vec(), d.new_vec()
    x.fill_uniform(2)
    xd.fill_uniform(2)
    for name, f in (("plain", lambda: a.mvp(x, out=y)), ("dist-self", lambda: d.mvp(xd, out=yd)),
                    ("local-plain", lambda: d.local.mvp(xd, out=yd)), ("plain-xd", lambda: a.mvp(xd, out=yd)),
                    ("dist-self-x", lambda: d.mvp(x, out=y))):
        for _ in range(10):
            f()
        ctx.sync()
        e0 = ctx.event().record()
        for _ in range(300):
            f()
        e1 = ctx.event().record()
        print(f"{name:10s} {e0.elapsed_ms(e1) / 300 * 1e3:8.2f} us", flush=True)
    assert np.array_equal(y.to_numpy(), yd.to_numpy())
    print("info", d.info(), flush=True)
