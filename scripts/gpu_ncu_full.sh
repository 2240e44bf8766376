#!/bin/bash
# full ncu capture of one kernel inside the proof bench; usage: KERNEL=msm_serial_reduce SKIP=12 bash scripts/gpu_ncu_full.sh
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
MODEL=${MODEL:-small}; KERNEL=${KERNEL:-msm_serial_reduce}
CMD="python bench.py --model $MODEL --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_full_$KERNEL.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s ${SKIP:-12} -c ${COUNT:-3} -f -o gpurun_out/prof_$KERNEL $CMD > gpurun_out/ncu_full_$KERNEL.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_full_$KERNEL.log
