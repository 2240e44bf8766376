#!/bin/bash
# window-width sweep for the proof workload: ZG_MSM_C overrides the default c = k - 2
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for m in ${MODELS:-small large}; do
  for c in ${CS:-0 14 15 16}; do
    ZG_MSM_C=$c timeout 600 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ctune_${m}_$c.json 2> gpurun_out/ctune_${m}_$c.err
    python -c "
import json; d=json.loads(open('gpurun_out/ctune_${m}_$c.json').read().strip().splitlines()[-1]); print('proof $m c=$c: %.2f ms/step  %.1f proofs/s  latency %.2f ms' % (d['ms_per_step'], d['value'], d['latency_ms_single_proof']), {k: round(v,2) for k,v in d['stage_ms_last_proof'].items()})" || tail -3 gpurun_out/ctune_${m}_$c.err
  done
done
