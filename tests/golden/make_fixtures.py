"""Writes tests/golden/oracle_fixtures.npz: outputs of the oracle (oracle/liboracle.so) on small seeded inputs.

    python tests/golden/make_fixtures.py

The reference is a Rust crate and cannot be run in this image, so these are NOT reference outputs; the
reference's own known answers live in reference_kats.json.  This file freezes the oracle's behaviour
(summation order, two-rounding axpy, CG trajectory) so that a later edit of the restatement cannot drift
unnoticed; tests/test_oracle_golden.py recomputes and compares byte for byte."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def compute(orc):
    import cases
    out = {}
    for tag, vdt, idt in (("f32u32", np.float32, np.uint32), ("f64u64", np.float64, np.uint64)):
        n_rows, n_cols, vals, cols, offs = cases.ragged(21, 200, 150, 70, vdt, idt)
        x = orc.uniform(vdt, 7, n_cols)
        out[f"mvp_ragged_{tag}"] = orc.mvp(vals, cols, offs, x)
        out[f"dot_{tag}"] = np.array([orc.dot(x, orc.uniform(vdt, 8, n_cols)), orc.norm2sq(x)])
        lv, lc, lo = orc.laplace(vdt, idt, 6, 5, 4)
        xs = orc.uniform(vdt, 6, 120)
        b = orc.mvp(lv, lc, lo, xs)
        sol = np.zeros(120, vdt)
        st = orc.cg(120, 120, lv, lc, lo, b, sol, tol=1e-6 if vdt == np.float32 else 1e-12, iter_max=300, history_cap=300)
        out[f"cg_x_{tag}"] = sol
        out[f"cg_history_{tag}"] = st["history"]
        out[f"cg_iters_{tag}"] = np.array([st["iterations"]], np.int64)
    v, c, o = orc.powerlaw(np.float64, np.uint64, 300, max_len=50)
    out["powerlaw_offsets"] = o
    out["powerlaw_columns"] = c
    out["powerlaw_values"] = v
    out["uniform_f32_seed2"] = orc.uniform(np.float32, 2, 64)
    return out


if __name__ == "__main__":
    from oracle import oracle_py
    np.savez_compressed(os.path.join(HERE, "oracle_fixtures.npz"), **compute(oracle_py))
    print("wrote oracle_fixtures.npz")
