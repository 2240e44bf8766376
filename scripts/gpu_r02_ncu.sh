#!/bin/bash
# ncu evidence for one k = 17 proof: every major kernel with --set full, exported to CSV ON THE BOX (reports are too large to
# carry back); one single-launch report of the top kernel is kept for the source page.  usage: TAG=before bash scripts/gpu_r02_ncu.sh
set -u
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${TAG:-r02}; O=gpurun_out/ncu_$TAG; mkdir -p $O
CMD="python bench.py --model ${MODEL:-large} --inflight 1 --proofs-per-lane 1 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $O/plain.json 2> $O/plain.err || { echo "plain run failed"; tail -5 $O/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s ${LSKIP:-600} -c 1400 --csv --log-file $O/launches.csv $CMD > $O/launches.log 2>&1; echo "launch list exit $?"
K='regex:msm_accumulate|ntt_|k_h_|msm_finish|msm_bucket|msm_warp_reduce|msm_serial|msm_digits|msm_tail|k_batch_invert|k_eval_partial|k_rank_inputs|k_place'
ncu --set full --clock-control none -k "$K" -s ${FSKIP:-330} -c ${FCOUNT:-170} -f -o $O/full $CMD > $O/full.log 2>&1; echo "full exit $?"
ncu -i $O/full.ncu-rep --page raw --csv > $O/full_raw.csv 2>/dev/null; rm -f $O/full.ncu-rep
python scripts/ncu_summary.py $O/full_raw.csv > $O/full_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 18 -c 1 -f -o $O/acc_one $CMD > $O/acc_one.log 2>&1; echo "acc one exit $?"
ls -la $O; du -sh gpurun_out
