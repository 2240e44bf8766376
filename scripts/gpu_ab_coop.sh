#!/bin/bash
# A/B of the cooperative MSM tail (ZG_MSM_TAIL_COOP=0|1, msm_tail_coop.cu) on ONE GPU: all GPU parity tests with it on,
# then bench lines per setting (latency is the figure of interest).
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/ab_coop; mkdir -p $O
ZG_MSM_TAIL_COOP=1 timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_coop.log 2>&1; echo "pytest coop exit $?"; tail -2 $O/pytest_coop.log
run() {  # name, env, args...
  local name=$1 envs=$2; shift 2
  env $envs timeout 300 python bench.py "$@" --no-cpu-baseline > $O/$name.json 2> $O/$name.err; echo "$name exit $?"
}
for a in 1 0; do
  run large_c$a "ZG_MSM_TAIL_COOP=$a" --steps 10 --warmup 3
  run small_c$a "ZG_MSM_TAIL_COOP=$a" --model small --steps 20 --warmup 5
  run tiny_c$a "ZG_MSM_TAIL_COOP=$a" --model tiny --steps 20 --warmup 5
  run msm20_c$a "ZG_MSM_TAIL_COOP=$a" --workload msm --logn 20 --steps 10 --warmup 3
done
O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ['O'] + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print('%-16s %9.3f ms/step %9.4g %s e2e %.4g lat %s frac %s stages %s' % (
            os.path.basename(f), d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'),
            d.get('roofline') and round(d['roofline']['frac'], 3), {k: round(v, 2) for k, v in d.get('stage_ms_last_proof', {}).items()}))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-600:])
PY
