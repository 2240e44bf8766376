#!/bin/bash
# bench-only gpurun call
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for wl in ${WORKLOADS:-msm ntt}; do
  for logn in ${LOGNS:-20}; do
    timeout 900 python bench.py --workload $wl --logn $logn --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${wl}_$logn.json 2> gpurun_out/bench_${wl}_$logn.err
    echo "bench $wl $logn exit $?"; cat gpurun_out/bench_${wl}_$logn.json; tail -3 gpurun_out/bench_${wl}_$logn.err
  done
done
