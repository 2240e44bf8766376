#!/bin/bash
# A/B of the lazy-reduction mixed addition of msm_accumulate_kernel on ONE GPU (ZG_MSM_MADD=0|1): parity tests under both
# settings, bench lines per setting.  (ZG_H_LAZY selected two-product Horner steps in k_h_lookup / k_h_gates when this
# script was first run; they measured no gain -- profiles/r02_notes.md -- and were removed; the variable is now ignored.)
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/ab_lazy; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1; nproc > $O/nproc.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_default.log 2>&1; echo "pytest default (lazy) exit $?"; tail -2 $O/pytest_default.log
ZG_MSM_MADD=0 ZG_H_LAZY=0 timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prover.py tests/test_golden_proofs.py tests/test_gpu_stages.py -m gpu -q -x > $O/pytest_plain.log 2>&1; echo "pytest plain exit $?"; tail -2 $O/pytest_plain.log
run() {  # name, env, args...
  local name=$1 envs=$2; shift 2
  env $envs timeout 600 python bench.py "$@" --no-cpu-baseline > $O/$name.json 2> $O/$name.err; echo "$name exit $?"
}
run large_m0h0 "ZG_MSM_MADD=0 ZG_H_LAZY=0" --steps 10 --warmup 3
run large_m1h0 "ZG_MSM_MADD=1 ZG_H_LAZY=0" --steps 10 --warmup 3
run large_m1h1 "ZG_MSM_MADD=1 ZG_H_LAZY=1" --steps 10 --warmup 3
run large_m0h0_b "ZG_MSM_MADD=0 ZG_H_LAZY=0" --steps 10 --warmup 3
run large_m1h1_b "ZG_MSM_MADD=1 ZG_H_LAZY=1" --steps 10 --warmup 3
run small_m0h0 "ZG_MSM_MADD=0 ZG_H_LAZY=0" --model small --steps 20 --warmup 5
run small_m1h1 "ZG_MSM_MADD=1 ZG_H_LAZY=1" --model small --steps 20 --warmup 5
run msm20_m0 "ZG_MSM_MADD=0" --workload msm --logn 20 --steps 10 --warmup 3
run msm20_m1 "ZG_MSM_MADD=1" --workload msm --logn 20 --steps 10 --warmup 3
O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ['O'] + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ip = d.get('int_pipe', {})
        print('%-22s %9.3f ms/step %9.4g %s e2e %.4g lat %s frac %s sqr %s two %s stages %s' % (
            os.path.basename(f), d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'),
            d['roofline'] and round(d['roofline']['frac'], 3), round(ip.get('fr_sqr_dedicated_gops', 0), 1), round(ip.get('fr_two_product_gops', 0), 1),
            {k: round(v, 2) for k, v in d.get('stage_ms_last_proof', {}).items()}))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
