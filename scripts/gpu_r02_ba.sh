#!/bin/bash
# batched-affine pre-reduction: parity with the rounds forced on, then A/B lines (MSM 2^20 and the k = 17 proof)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-ba}; O=gpurun_out/$TAG; mkdir -p $O
for r in 1 2; do
  ZG_MSM_BA=$r timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_golden_kernels.py tests/test_gpu_prover.py -m gpu -x -q > $O/pytest_ba$r.log 2>&1; echo "pytest BA=$r exit $?"; tail -3 $O/pytest_ba$r.log
done
for r in 0 1 2 3; do
  ZG_MSM_BA=$r timeout 300 python bench.py --workload msm --logn 20 --steps 10 --warmup 3 --no-cpu-baseline > $O/msm20_ba$r.json 2> $O/msm20_ba$r.err; echo "msm20 BA=$r exit $?"
done
for r in 0 1 2; do
  ZG_MSM_BA=$r timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_large_ba$r.json 2> $O/proof_large_ba$r.err; echo "proof BA=$r exit $?"
done
O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ['O'] + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print('%-28s %9.3f ms/step %10.4g %s e2e %.4g lat %s frac %s stages %s' % (os.path.basename(f), d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'), d['roofline'] and round(d['roofline']['frac'], 3), {k: round(v, 2) for k, v in (d.get('stage_ms_last_proof') or {}).items()}))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
