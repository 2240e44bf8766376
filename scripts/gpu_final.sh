#!/bin/bash
# Round-2 evidence on ONE GPU: parity tests, smoke, the default bench line with the CPU baseline, the other model shapes and sweeps
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/final; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1; nproc > $O/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench default exit $?"
for m in small tiny medium; do
  timeout 600 python bench.py --model $m --steps 20 --warmup 5 ${CPUB:---no-cpu-baseline} > $O/bench_proof_$m.json 2> $O/bench_proof_$m.err; echo "bench $m exit $?"
done
timeout 600 python bench.py --model large --inflight 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_proof_large_1lane.json 2> $O/bench_proof_large_1lane.err
timeout 600 python bench.py --workload verify --model small --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_verify_small.json 2> $O/bench_verify_small.err; echo "verify exit $?"
timeout 600 python bench.py --workload keygen --model large --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_keygen_large.json 2> $O/bench_keygen_large.err; echo "keygen exit $?"
: > $O/sweep.jsonl
for wl in msm ntt; do for l in 16 18 20 22 24; do
  extra="--no-cpu-baseline"; [ "$l" = "20" ] && extra=""
  timeout 900 python bench.py --workload $wl --logn $l --steps 5 --warmup 3 $extra >> $O/sweep.jsonl 2> $O/sweep_${wl}_$l.err || echo "FAIL $wl $l"
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final/bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        cb = d.get('cpu_baseline')
        print('%-36s %9.3f ms/step %10.4g %s  e2e %.4g  lat %s  frac %s  cpu %s' % (f.split('/')[-1], d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'), d['roofline']['frac'] if d.get('roofline') else None, cb and cb['value']))
    except Exception as e:
        print(f, 'ERR', e, open(f.replace('.json', '.err')).read()[-800:])
for l in open('gpurun_out/final/sweep.jsonl'):
    d = json.loads(l); print('%-28s %9.3f ms %12.4g %s frac %.3f int %s' % (d['config']['workload'][:28], d['ms_per_step'], d['value'], d['unit'], d['roofline']['frac'], d.get('int_pipe', {}).get('kernel_mulmod_gops')))
PY
