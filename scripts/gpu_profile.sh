#!/bin/bash
# Runs on the B200 box under gpurun: ncu launch list + one full capture of the SpMV kernel for the bench command.
# usage: scripts/gpu_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r01}
KRE=${2:-spmv_}
CMD="python bench.py --steps 20 --warmup 3 --no-cpu --no-cg"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 6 -c 2 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/plain_${TAG}.log
