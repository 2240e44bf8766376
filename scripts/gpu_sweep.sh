#!/bin/bash
# MSM / NTT sweeps (BASELINE configs[4]: k = 16..24); one JSON line per point -> gpurun_out/sweep.jsonl
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
: > gpurun_out/sweep.jsonl
for wl in ${WORKLOADS:-msm ntt}; do
  for logn in ${LOGNS:-16 18 20 22 24}; do
    extra="--no-cpu-baseline"
    if [ "$logn" = "${CPU_AT:-20}" ]; then extra=""; fi
    timeout 900 python bench.py --workload $wl --logn $logn --steps ${STEPS:-5} --warmup 3 $extra >> gpurun_out/sweep.jsonl 2> gpurun_out/sweep_${wl}_$logn.err || { echo "FAIL $wl $logn"; tail -3 gpurun_out/sweep_${wl}_$logn.err; }
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d = json.loads(l)
    cb = d.get('cpu_baseline')
    print('%-28s %9.3f ms  %12.4g %s  frac %.3f  e2e %.4g %s' % (d['config']['workload'][:28], d['ms_per_step'], d['value'], d['unit'], d['roofline']['frac'], d['e2e']['value'], ('cpu %.4g x%d' % (cb['value'], cb['cores'])) if cb else ''))
PY
