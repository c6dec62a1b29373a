#!/bin/bash
# GPU (N ranks, default 2): A/B of the distributed ring schedules.  An arm is OVERLAP:RESERVE —
#   0:0  halo awaited first, then one launch over all local rows on every SM
#   0:R  interior (leaving R SMs to NCCL) | exchange, then the boundary rows on the main stream
#   1:R  interior (leaving R SMs) | exchange -> boundary rows on the side stream, in the slots the interior left
# usage: scripts/dist_reserve_ab.sh [N] [arms...]
N=${1:-2}; shift
ARMS=${@:-0:0 1:4 1:8}
for A in $ARMS; do
  O=${A%%:*}; R=${A##*:}
  CH=$R; [ "$R" = "0" ] && CH=32
  SMB200_DIST_OVERLAP=$O SMB200_DIST_RESERVE_SMS=$R NCCL_MAX_P2P_NCHANNELS=$CH python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port $((29600 + R + 20 * O)) bench.py --gpus $N --steps 200 --warmup 5 --no-cpu 2>gpurun_out/dist_ab_${O}_$R.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('overlap=$O reserve=$R N=$N: %.1f GB/s  %.4f ms/step  CG %.0f it/s  local kernel %.4f ms' % (d['value'], d['ms_per_step'], d['cg']['iter_per_s'], d['roofline']['kernel_ms']))"
done
