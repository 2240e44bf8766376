#!/bin/bash
# multi-GPU bench lines: N=<gpus> bash scripts/gpu_multi.sh   (run under gpurun --gpus N)
set -u
mkdir -p gpurun_out/final
cd "${GRAFT_REPO_ROOT:-.}"
N=${N:-2}
run() {  # name, bench args...
  local name=$1; shift
  timeout ${RUN_TIMEOUT:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/final/multi_${name}_n$N.json 2> gpurun_out/final/multi_${name}_n$N.err
  python -c "
import json; d=json.loads(open('gpurun_out/final/multi_${name}_n$N.json').read().strip().splitlines()[-1]); print('$name x$N: %.3f ms/step  %.4g %s  e2e %.4g  scaling %s' % (d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d['scaling']))" || tail -5 gpurun_out/final/multi_${name}_n$N.err
}
for m in ${MODELS:-small large}; do run proof_$m --model $m --steps 5 --warmup 3 --no-cpu-baseline; done
for l in ${SHARD_LOGNS:-17 20 24}; do run msmshard_$l --workload msm_sharded --logn $l --steps 5 --warmup 3 --no-cpu-baseline; done
for l in ${MSM_LOGNS:-20}; do run msm_$l --workload msm --logn $l --steps 5 --warmup 3 --no-cpu-baseline; done
for l in ${NTT_LOGNS:-20}; do run ntt_$l --workload ntt --logn $l --steps 5 --warmup 3 --no-cpu-baseline; done
