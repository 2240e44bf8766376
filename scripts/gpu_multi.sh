#!/bin/bash
# multi-GPU checks and strong-scaling lines; run with gpurun --gpus N -- bash scripts/multi.sh N [tag]
set -u
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}; TAG=${2:-multi_n$N}; O=gpurun_out/$TAG; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > $O/pytest_dist.log 2>&1; echo "pytest dist exit $?"; tail -5 $O/pytest_dist.log
for l in 20 24; do
  timeout 600 $TR bench.py --gpus $N --workload msm_sharded --logn $l --steps 10 --warmup 3 --no-cpu-baseline > $O/msmshard_${l}.json 2> $O/msmshard_${l}.err; echo "msm_sharded $l exit $?"
done
timeout 900 $TR bench.py --gpus $N --model large --shard columns --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_large_columns.json 2> $O/proof_large_columns.err; echo "proof columns exit $?"
timeout 900 $TR bench.py --gpus $N --model large --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_large_weak.json 2> $O/proof_large_weak.err; echo "proof weak exit $?"
O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ['O'] + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print('%-30s n=%d %9.3f ms/step %10.4g %s e2e %.4g lat %s scaling %s' % (os.path.basename(f), d['n_gpus'], d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'), d['scaling']))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-1500:])
PY
