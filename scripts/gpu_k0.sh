#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for m in ${MODELS:-small large}; do
  for k0 in ${K0S:-8 16 32 64}; do
    ZG_MSM_K0=$k0 timeout 300 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k0.json 2> gpurun_out/k0.err
    python -c "
import json; d=json.loads(open('gpurun_out/k0.json').read().strip().splitlines()[-1]); print('proof $m K0=$k0: %.1f proofs/s  latency %.2f ms  acc frac %.3f avg %.3f ms' % (d['value'], d['latency_ms_single_proof'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))" || tail -3 gpurun_out/k0.err
  done
done
