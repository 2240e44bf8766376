#!/bin/bash
# strong scaling of ONE k = 17 proof stream over N GPUs (column-distributed MSM rounds + row-block-distributed extended
# forms / quotient numerator): gpurun --gpus N -- bash scripts/gpu_columns.sh N
set -u
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}; O=gpurun_out/columns_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
ZG_RNG_THREADS=4 timeout 600 $TR bench.py --gpus $N --model large --shard columns --steps 10 --warmup 3 --no-cpu-baseline > $O/proof_large_columns.json 2> $O/proof_large_columns.err; echo "proof columns exit $?"
O=$O python - <<'PY'
import json, os
O = os.environ['O']
try:
    d = json.loads(open(O + '/proof_large_columns.json').read().strip().splitlines()[-1])
    print('columns n=%d' % d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['latency_ms_single_proof'], d['stage_ms_last_proof'])
except Exception as e:
    print('ERR', e, open(O + '/proof_large_columns.err').read()[-1500:])
PY
