#!/bin/bash
# ncu evidence for one k = 17 proof stream (one lane).  `--set full` replays every profiled launch ~40 times and saves /
# restores the 2-GB key arena around each pass (about 8 s per launch), so the captures are SMALL: a handful of launches per
# kernel family, exported to CSV on the box (reports are too large to carry back; one single-launch report of the top
# kernel is kept for the source page).  usage: bash scripts/ncu.sh [tag]
set -u
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-ncu}; O=gpurun_out/$TAG; mkdir -p $O
CMD="python bench.py --model ${MODEL:-large} --inflight 1 --proofs-per-lane 1 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $O/plain.json 2> $O/plain.err || { echo "plain run failed"; tail -5 $O/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s ${LSKIP:-700} -c 1500 --csv --log-file $O/launches_proof_large.csv $CMD > $O/launches.log 2>&1; echo "launch list exit $?"
python scripts/launch_summary.py $O/launches_proof_large.csv 7 > $O/launches_proof_large_summary.txt 2>&1
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $O/$1 $CMD > $O/$1.log 2>&1; echo "full $1 exit $?"
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $O/$1_raw.csv > $O/ncu_full_$1.txt 2>&1
}
cap accumulate 'msm_accumulate' 17 3           # proof 4: advice / lookups / products rounds
ncu -i $O/accumulate.ncu-rep --page source --csv --kernel-name regex:msm_accumulate > $O/accumulate_source.csv 2>/dev/null
head -c 3000000 $O/accumulate_source.csv > $O/accumulate_source_head.csv; rm -f $O/accumulate_source.csv
rm -f $O/accumulate.ncu-rep
[ "${NCU_ONLY:-}" = "accumulate" ] && { ls -la $O; exit 0; }   # the other kernel families did not change: keep their digests
cap ntt 'ntt_pass' 100 4
rm -f $O/ntt.ncu-rep
cap hkern 'k_h_lookup|k_h_gates|k_h_permutation' 12 3
rm -f $O/hkern.ncu-rep
cap tail 'msm_finish|msm_bucket_l1|msm_bucket_l2|msm_warp_reduce|msm_serial_reduce|msm_digits' 60 8
rm -f $O/tail.ncu-rep
ls -la $O; du -sh gpurun_out
