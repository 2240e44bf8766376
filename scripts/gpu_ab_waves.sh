#!/bin/bash
# A/B of the whole-wave grid rule of msm_accumulate_kernel (ZG_MSM_AUTOWAVES=0|1, msm.cu::msm_workspace_layout) on ONE GPU
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/ab_waves; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prover.py tests/test_golden_proofs.py tests/test_golden_kernels.py tests/test_gpu_field.py tests/test_verifier.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -2 $O/pytest.log
run() {  # name, env, args...
  local name=$1 envs=$2; shift 2
  env $envs timeout 600 python bench.py "$@" --no-cpu-baseline > $O/$name.json 2> $O/$name.err; echo "$name exit $?"
}
for a in 0 1; do
  run large_w$a "ZG_MSM_AUTOWAVES=$a" --steps 10 --warmup 3
  run small_w$a "ZG_MSM_AUTOWAVES=$a" --model small --steps 20 --warmup 5
  run tiny_w$a "ZG_MSM_AUTOWAVES=$a" --model tiny --steps 20 --warmup 5
  run msm16_w$a "ZG_MSM_AUTOWAVES=$a" --workload msm --logn 16 --steps 10 --warmup 3
  run msm18_w$a "ZG_MSM_AUTOWAVES=$a" --workload msm --logn 18 --steps 10 --warmup 3
done
run large_w0_b "ZG_MSM_AUTOWAVES=0" --steps 10 --warmup 3
run large_w1_b "ZG_MSM_AUTOWAVES=1" --steps 10 --warmup 3
run verify_small "" --workload verify --model small --steps 20 --warmup 3
O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ['O'] + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print('%-18s %9.3f ms/step %9.4g %s e2e %.4g lat %s frac %s stages %s' % (
            os.path.basename(f), d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'),
            d.get('roofline') and round(d['roofline']['frac'], 3), {k: round(v, 2) for k, v in d.get('stage_ms_last_proof', {}).items()}))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
