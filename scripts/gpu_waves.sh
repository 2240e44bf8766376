#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 400 python -m pytest tests/test_gpu_msm.py tests/test_golden_proofs.py -m gpu -x -q 2>&1 | tail -2
for w in ${WAVES:-1 2 4}; do
  for m in small large; do
    ZG_MSM_WAVES=$w timeout 300 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k0.json 2> gpurun_out/k0.err
    python -c "
import json; d=json.loads(open('gpurun_out/k0.json').read().strip().splitlines()[-1]); print('proof $m waves=$w: %.1f proofs/s  latency %.2f ms  acc frac %.3f avg %.3f ms' % (d['value'], d['latency_ms_single_proof'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))" || tail -3 gpurun_out/k0.err
  done
  ZG_MSM_WAVES=$w timeout 300 python bench.py --workload msm --logn 20 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k0.json 2> gpurun_out/k0.err
  python -c "
import json; d=json.loads(open('gpurun_out/k0.json').read().strip().splitlines()[-1]); print('msm 2^20 waves=$w: %.3f ms  acc frac %.3f avg %.3f ms' % (d['ms_per_step'], d['roofline']['frac'], d['roofline']['avg_launch_ms']))" || tail -3 gpurun_out/k0.err
done
