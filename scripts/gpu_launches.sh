#!/bin/bash
# per-kernel launch lists of ONE proof stream (ncu --metrics gpu__time_duration.sum; cold-cache, serialised times)
# usage: bash scripts/launches.sh tag [model ...]   (env is passed through, e.g. ZG_NTT_STAGEWISE=1)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-launches}; shift || true
MODELS=${*:-large}
O=gpurun_out/$TAG; mkdir -p $O
for m in $MODELS; do
  CMD="python bench.py --model $m --inflight 1 --proofs-per-lane 1 --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > $O/plain_$m.json 2> $O/plain_$m.err || { echo "plain run failed"; tail -5 $O/plain_$m.err; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -s ${LSKIP:-700} -c ${LCOUNT:-1500} --csv --log-file $O/launches_$m.csv $CMD > $O/launches_$m.log 2>&1; echo "launch list $m exit $?"
  python scripts/launch_summary.py $O/launches_$m.csv > $O/launches_${m}_summary.txt 2>&1; head -40 $O/launches_${m}_summary.txt
done
