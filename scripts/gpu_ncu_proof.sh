#!/bin/bash
# launch list of one whole proof (per-kernel durations); usage: MODEL=small bash scripts/gpu_ncu_proof.sh
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
MODEL=${MODEL:-small}
CMD="python bench.py --model $MODEL --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_proof_$MODEL.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-150} -c ${COUNT:-1200} --csv --log-file gpurun_out/launches_proof_$MODEL.csv $CMD > gpurun_out/ncu_proof_$MODEL.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_proof_$MODEL.log
