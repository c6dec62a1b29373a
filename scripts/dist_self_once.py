#!/usr/bin/env python
"""GPU (one device): a few products through the DISTRIBUTED instantiation of the ring kernel with zero neighbours
(SMB200_DIST_SELF=1) — the command ncu profiles to show what the halo code costs inside the kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SMB200_DIST_SELF"] = "1"
import sparsemat_b200 as smb  # noqa: E402

ctx = smb.Context(0)
ctx.comm_init(0, 1, None)
d = smb.DistCRS.laplace(ctx, np.float32, np.uint32, 256, 256, 256)
x, y = d.new_vec(), d.new_vec()
x.fill_uniform(2)
for _ in range(6):
    d.mvp(x, out=y)
ctx.sync()
print("ok", d.info())
