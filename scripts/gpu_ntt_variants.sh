#!/bin/bash
# rebuilds ntt.cu with different tile shapes on the GPU box and times the NTT workload
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for v in ${VARIANTS:-"256 2" "128 1" "128 2" "256 1"}; do
  set -- $v
  make -C 0g-halo2_b200/csrc -s EXTRA="-DZG_NTT_THREADS=$1 -DZG_NTT_LOGC=$2 -DZG_NTT_MAX_S=${3:-8}" ntt.o capi.o prover.o > /dev/null 2>&1; touch 0g-halo2_b200/csrc/ntt.cuh
  make -C 0g-halo2_b200/csrc -s EXTRA="-DZG_NTT_THREADS=$1 -DZG_NTT_LOGC=$2 -DZG_NTT_MAX_S=${3:-8}" > gpurun_out/ntt_variant_build.log 2>&1 || { echo "build failed $v"; tail -3 gpurun_out/ntt_variant_build.log; continue; }
  for l in ${LOGNS:-18 20 24}; do
    timeout 300 python bench.py --workload ntt --logn $l --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/nttv.json 2> gpurun_out/nttv.err
    python -c "
import json; d=json.load(open('gpurun_out/nttv.json')); print('threads $1 logc $2 maxS ${3:-8}  ntt 2^$l: %.4f ms %.1f GB/s mulmod %.1f G/s' % (d['ms_per_step'], d['value'], d['int_pipe']['kernel_mulmod_gops']))" || tail -3 gpurun_out/nttv.err
  done
done
