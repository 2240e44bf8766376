#!/bin/bash
# parity tests, then A/B bench lines for the round-2 kernel changes (usage: bash scripts/gpu_r02_check.sh [tag])
set -u
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-check}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $O/pytest_gpu.log
for l in 16 18 20; do
  ZG_NTT_STAGEWISE=1 timeout 300 python bench.py --workload ntt --logn $l --steps 10 --warmup 3 --no-cpu-baseline > $O/ntt_${l}_stagewise.json 2> $O/ntt_${l}_stagewise.err
  timeout 300 python bench.py --workload ntt --logn $l --steps 10 --warmup 3 --no-cpu-baseline > $O/ntt_${l}_fast.json 2> $O/ntt_${l}_fast.err
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_large.json 2> $O/proof_large.err; echo "large exit $?"
ZG_EXT_POW2=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_large_pow2.json 2> $O/proof_large_pow2.err; echo "large pow2 exit $?"
timeout 600 python bench.py --model small --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_small.json 2> $O/proof_small.err; echo "small exit $?"
TAG=$TAG O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ.get('O', 'gpurun_out/%s' % os.environ.get('TAG', 'check')) + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print('%-28s %9.3f ms/step %10.4g %s e2e %.4g lat %s frac %s int %s stages %s' % (
            os.path.basename(f), d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'),
            d['roofline'] and round(d['roofline']['frac'], 3), d.get('int_pipe', {}).get('kernel_mulmod_gops'), d.get('stage_ms_last_proof')))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-600:])
PY
