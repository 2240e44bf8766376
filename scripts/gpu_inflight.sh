#!/bin/bash
# proof throughput vs proofs in flight per GPU
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for m in ${MODELS:-small large}; do
  for k in ${INFLIGHT:-1 2 3 4 6}; do
    timeout 900 python bench.py --model $m --inflight $k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/inflight_${m}_$k.json 2> gpurun_out/inflight_${m}_$k.err
    python -c "
import json; d=json.load(open('gpurun_out/inflight_${m}_$k.json')); print('proof $m inflight $k: %.2f ms/step  %.1f proofs/s  e2e %.1f proofs/s  latency %.2f ms  launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['latency_ms_single_proof'], d['gpu_launches']))" || tail -5 gpurun_out/inflight_${m}_$k.err
  done
done
