#!/bin/bash
# launch list (per-kernel durations) of one bench command; usage: WL=msm LOGN=17 bash scripts/gpu_ncu_launches.sh
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
WL=${WL:-msm}; LOGN=${LOGN:-17}
CMD="python bench.py --workload $WL --logn $LOGN --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_${WL}_$LOGN.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c ${COUNT:-400} --csv --log-file gpurun_out/launches_${WL}_$LOGN.csv $CMD > gpurun_out/ncu_${WL}_$LOGN.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_${WL}_$LOGN.log
