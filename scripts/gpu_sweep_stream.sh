#!/bin/bash
set -u
mkdir -p gpurun_out
if [ -z "${NOTEST:-}" ]; then timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log; fi
{
for cap in ${CAPS:-4608 3072 2048 6144}; do
  echo "== STREAM cap=$cap"
  SMB200_STREAM_CAP=$cap timeout 120 python - <<'PY'
import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
import sparsemat_b200 as smb
from bench_variants import run
ctx = smb.Context(0)
V = [(smb.SPMV_STREAM, 0), (smb.SPMV_STREAM_TMA, 0)]
run(ctx, "C2 f32", smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, 256, 256, 256), V)
run(ctx, "C4 f64", smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 256, 256, 256), V)
PY
done
} > gpurun_out/sweep_stream.log 2>&1
cat gpurun_out/sweep_stream.log
