#!/bin/bash
# ncu --set full of one kernel inside a bench command.
# usage: NAME=msm20 KERNEL=msm_serial_reduce ARGS="--workload msm --logn 20" SKIP=3 COUNT=2 bash scripts/gpu_ncu_kernel.sh
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
NAME=${NAME:-k}; KERNEL=${KERNEL:-msm_serial_reduce}
CMD="python bench.py ${ARGS:-} --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$NAME.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s ${SKIP:-3} -c ${COUNT:-2} -f -o gpurun_out/prof_$NAME $CMD > gpurun_out/ncu_$NAME.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_$NAME.log | cut -c1-300
