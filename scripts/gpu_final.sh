#!/bin/bash
# Round-end evidence on ONE GPU: parity tests, smoke, bench lines (ours + reference arm), sweeps, launch lists, ncu captures.
set -u
mkdir -p gpurun_out/final
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/final
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1; nproc > $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $O/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.log
# bench lines with the CPU baseline beside them
for m in small large tiny medium; do
  timeout 900 python bench.py --model $m --steps 10 --warmup 3 > $O/bench_proof_$m.json 2> $O/bench_proof_$m.err; echo "bench $m exit $?"
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_small.json 2> $O/bench_reference_small.err; echo "reference exit $?"
for k in 1 2 6 8; do
  timeout 600 python bench.py --model small --inflight $k --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_proof_small_inflight$k.json 2> /dev/null
done
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
        print('%-40s %9.3f ms/step %10.4g %s  e2e %.4g  lat %s  frac %s  cpu %s' % (f.split('/')[-1], d['ms_per_step'], d['value'], d['unit'], d['e2e']['value'], d.get('latency_ms_single_proof'), d['roofline']['frac'] if d.get('roofline') else None, cb and cb['value']))
    except Exception as e:
        print(f, 'ERR', e)
for l in open('gpurun_out/final/sweep.jsonl'):
    d = json.loads(l); print('%-28s %9.3f ms %12.4g %s frac %.3f' % (d['config']['workload'][:28], d['ms_per_step'], d['value'], d['unit'], d['roofline']['frac']))
PY
