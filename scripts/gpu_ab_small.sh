set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/ab_small; mkdir -p $O
for envs in "" "ZG_MSM_WAVES=1" "ZG_MSM_WAVES=2" "ZG_MSM_K0=16" "ZG_MSM_C=14" "ZG_MSM_C=15"; do
  name=$(echo "small $envs" | tr ' =' '__' | tr -cd 'A-Za-z0-9_')
  env $envs timeout 600 python bench.py --model small --steps 10 --warmup 3 --no-cpu-baseline > $O/$name.json 2> $O/$name.err; echo "$name exit $?"
done
O=$O python - <<'PY'
import json, glob, os
for f in sorted(glob.glob(os.environ['O'] + '/*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print('%-32s %9.3f ms/step %8.4g %s e2e %.4g lat %.3f frac %s' % (os.path.basename(f), d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'), d['roofline'] and round(d['roofline']['frac'], 3)))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-500:])
PY
