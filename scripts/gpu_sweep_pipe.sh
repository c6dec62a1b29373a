#!/bin/bash
# GPU: parity first, then a parameter sweep of the persistent TMA-ring SpMV on C2 / C4.
set -u
mkdir -p gpurun_out
if [ -z "${NOTEST:-}" ]; then timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log; fi
{
SWEEP=${SWEEP:-"4096,3,2,256 2048,3,4,256 2048,4,3,256 4096,4,1,512 8192,3,1,512 3072,4,2,256"}
for cfg in $SWEEP; do
  IFS=, read cap st ct th <<< "$cfg"
  echo "== cap=$cap stages=$st ctas=$ct threads=$th"
  SMB200_PIPE_CAP=$cap SMB200_PIPE_STAGES=$st SMB200_PIPE_CTAS=$ct SMB200_PIPE_THREADS=$th SMB200_SPMV_VARIANT=6 timeout 120 python - <<'PY'
import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
import sparsemat_b200 as smb
from bench_variants import run
ctx = smb.Context(0)
run(ctx, "C2 f32", smb.SparseMatCRS.laplace(ctx, np.float32, np.uint32, 256, 256, 256), [(smb.SPMV_STREAM_PIPE, 0)])
run(ctx, "C4 f64", smb.SparseMatCRS.laplace(ctx, np.float64, np.uint32, 256, 256, 256), [(smb.SPMV_STREAM_PIPE, 0)])
PY
done
} > gpurun_out/sweep_pipe.log 2>&1
cat gpurun_out/sweep_pipe.log
