#!/bin/bash
# Round-2 starting point on ONE GPU: default bench line (k = 17 target), small model, launch list and full ncu captures at k = 17.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r02_base; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1; nproc > $O/nproc.txt
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/bench_default.json 2> $O/bench_default.err; echo "bench default exit $?"
timeout 600 python bench.py --model small --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_small.json 2> $O/bench_small.err; echo "bench small exit $?"
CMD="python bench.py --model large --inflight 1 --proofs-per-lane 1 --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 1200 --csv --log-file $O/launches_proof_large.csv $CMD > $O/ncu_launches_large.log 2>&1; echo "launch list exit $?"
CMD="python bench.py --model large --inflight 1 --proofs-per-lane 1 --steps 1 --warmup 3 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 12 -c 5 -f -o $O/prof_accumulate_proof_large $CMD > $O/ncu_full_acc_large.log 2>&1; echo "full acc large exit $?"
ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 60 -c 8 -f -o $O/prof_ntt_proof_large $CMD > $O/ncu_full_ntt_large.log 2>&1; echo "full ntt large exit $?"
ncu --set full --clock-control none --import-source on -k "regex:k_h_|msm_finish|msm_bucket|msm_warp_reduce|msm_serial" -s 40 -c 24 -f -o $O/prof_misc_proof_large $CMD > $O/ncu_full_misc_large.log 2>&1; echo "full misc large exit $?"
ls -la $O
tail -c 1500 $O/bench_default.json
