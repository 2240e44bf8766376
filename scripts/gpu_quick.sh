#!/bin/bash
# quick perf check: msm 2^17 / 2^20, proof small / large (no cpu baseline), prints compact lines
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
if [ "${TESTS:-1}" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_SEL:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
fi
for logn in ${MSM_LOGNS:-17 20}; do
  timeout 600 python bench.py --workload msm --logn $logn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_msm_$logn.json 2> gpurun_out/q_msm_$logn.err
  python -c "
import json; d=json.load(open('gpurun_out/q_msm_$logn.json')); print('msm 2^$logn: %.3f ms  %.1f Mpts/s  frac %.3f  mulmod_peak %.1f' % (d['ms_per_step'], d['value']/1e6, d['roofline']['frac'], d['int_pipe']['fr_mulmod_evenodd_gops']))" || tail -3 gpurun_out/q_msm_$logn.err
done
for logn in ${NTT_LOGNS:-20}; do
  timeout 600 python bench.py --workload ntt --logn $logn --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_ntt_$logn.json 2> gpurun_out/q_ntt_$logn.err
  python -c "
import json; d=json.load(open('gpurun_out/q_ntt_$logn.json')); print('ntt 2^$logn: %.3f ms  %.1f GB/s  kernel mulmod %.1f G/s' % (d['ms_per_step'], d['value'], d['int_pipe']['kernel_mulmod_gops']))" || tail -3 gpurun_out/q_ntt_$logn.err
done
for m in ${MODELS:-small large}; do
  timeout 900 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/q_proof_$m.json 2> gpurun_out/q_proof_$m.err
  python -c "
import json; d=json.load(open('gpurun_out/q_proof_$m.json')); print('proof $m: %.2f ms  e2e %.2f proofs/s  launches %d' % (d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])); print('   ', {k: round(v,2) for k,v in d['stage_ms_last_proof'].items()})" || tail -3 gpurun_out/q_proof_$m.err
done
