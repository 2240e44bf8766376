#!/bin/bash
# One gpurun call: GPU parity tests, smoke, first bench lines.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
for wl in msm ntt; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "bench $wl exit $?"; cat gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
