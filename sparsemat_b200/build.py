"""Build libsmb200.so in-tree with nvcc for sm_100a (`make -C sparsemat_b200/csrc`).  Rebuild helper for an existing tree:
importing this module imports the package, which needs the library — a fresh checkout builds through
`python __graft_entry__.py` (or make) instead."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose: bool = False, jobs: int = 8) -> str:
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), f"-j{jobs}"]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        sys.stderr.write((res.stdout or "") + (res.stderr or ""))
        raise RuntimeError("building libsmb200.so failed")
    out = os.path.join(HERE, "lib", "libsmb200.so")
    if not os.path.exists(out):
        raise RuntimeError(f"{out} missing after build")
    return out


if __name__ == "__main__":
    print(build(verbose=True))
