#!/bin/bash
# NTT A/B: stage-wise vs register-radix kernel, stand-alone and inside the k = 17 proof (usage: bash scripts/gpu_r02_check2.sh [tag])
set -u
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-check2}; O=gpurun_out/$TAG; mkdir -p $O
ZG_NTT_FAST=1 timeout 600 python -m pytest tests/test_gpu_ntt.py tests/test_gpu_stages.py -m gpu -x -q > $O/pytest_gpu_fast.log 2>&1; echo "pytest (forced fast NTT) exit $?"; tail -4 $O/pytest_gpu_fast.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest_gpu.log
for l in 18 20 22; do
  ZG_NTT_STAGEWISE=1 timeout 300 python bench.py --workload ntt --logn $l --steps 10 --warmup 3 --no-cpu-baseline > $O/ntt_${l}_stagewise.json 2> $O/ntt_${l}_stagewise.err
  ZG_NTT_FAST=1 timeout 300 python bench.py --workload ntt --logn $l --steps 10 --warmup 3 --no-cpu-baseline > $O/ntt_${l}_fast.json 2> $O/ntt_${l}_fast.err
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_large.json 2> $O/proof_large.err; echo "large exit $?"
ZG_NTT_STAGEWISE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_large_stagewise.json 2> $O/proof_large_stagewise.err; echo "large stagewise exit $?"
timeout 600 python bench.py --model small --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_small.json 2> $O/proof_small.err; echo "small exit $?"
ZG_NTT_STAGEWISE=1 timeout 600 python bench.py --model small --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_small_stagewise.json 2> $O/proof_small_stagewise.err; echo "small stagewise exit $?"
TAG=$TAG O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ['O'] + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print('%-30s %9.3f ms/step %10.4g %s e2e %.4g lat %s frac %s int %s stages %s' % (
            os.path.basename(f), d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'),
            d['roofline'] and round(d['roofline']['frac'], 3), d.get('int_pipe', {}).get('kernel_mulmod_gops'), d.get('stage_ms_last_proof')))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-600:])
PY
