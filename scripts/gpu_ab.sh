#!/bin/bash
# A/B lines for tuning knobs on the k = 17 proof (4 lanes): usage bash scripts/ab.sh tag "ENV1=.. ENV2=.." ...
set -u
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-ab}; shift
O=gpurun_out/$TAG; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_prover.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest.log
i=0
for envs in "" "$@"; do
  name=$(echo "base $envs" | tr ' =' '__' | tr -cd 'A-Za-z0-9_')
  env $envs timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/$name.json 2> $O/$name.err; echo "$name exit $?"
  i=$((i+1))
done
O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ['O'] + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print('%-40s %9.3f ms/step %8.4g %s e2e %.4g lat %.3f frac %s stages %s' % (os.path.basename(f), d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'), d['roofline'] and round(d['roofline']['frac'], 3), {k: round(v, 2) for k, v in d.get('stage_ms_last_proof', {}).items()}))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
PY
