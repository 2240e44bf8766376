#!/bin/bash
# ncu evidence (run AFTER gpu_final.sh has exited 0 for the same commands): launch lists + full captures of the top kernels
set -u
mkdir -p gpurun_out/final
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/final
for m in small large; do
  CMD="python bench.py --model $m --inflight 1 --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > $O/plain_proof_$m.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 900 --csv --log-file $O/launches_proof_$m.csv $CMD > $O/ncu_launches_$m.log 2>&1
  echo "launch list $m exit $?"
done
CMD="python bench.py --model small --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 10 -c 5 -f -o $O/prof_accumulate_proof_small $CMD > $O/ncu_full_acc_small.log 2>&1; echo "full acc small exit $?"
CMD="python bench.py --workload msm --logn 20 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 3 -c 1 -f -o $O/prof_accumulate_msm20 $CMD > $O/ncu_full_acc_msm20.log 2>&1; echo "full acc msm20 exit $?"
CMD="python bench.py --workload ntt --logn 20 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 9 -c 3 -f -o $O/prof_ntt20 $CMD > $O/ncu_full_ntt20.log 2>&1; echo "full ntt20 exit $?"
ls -la $O/*.ncu-rep
