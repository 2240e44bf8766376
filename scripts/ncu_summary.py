#!/usr/bin/env python
"""Compact summary of an .ncu-rep (ncu -i ... --page raw --csv) for profiles/: one block per captured launch."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fmaheavy.sum",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def main(rep):
    if rep.endswith(".csv"):                 # already exported on the GPU box (ncu -i x.ncu-rep --page raw --csv > x.csv)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print("kernel: %s" % d[idx["Kernel Name"]][:140])
        for w in WANT:
            if w in idx:
                print("  %-82s %s %s" % (w, d[idx[w]], units[idx[w]]))
        # every per-pipe counter the capture holds (names differ between ncu versions: fmaheavy / fma_heavy / imad ...)
        for h in hdr:
            if h in WANT:
                continue
            hl = h.lower()
            if ("pipe_" in hl and ("fma" in hl or "alu" in hl or "imad" in hl or "xu" in hl or "lsu" in hl or "uniform" in hl)) or \
               hl.startswith("launch__occupancy") or hl.startswith("launch__waves") or "registers" in hl or "shared_mem" in hl or \
               hl.startswith("smsp__average_warps_issue_stalled") or hl.startswith("smsp__warps_eligible") or \
               hl.startswith("sm__warps_active") or hl.startswith("lts__t_bytes") or hl.startswith("l1tex__data_bank_conflicts"):
                print("  %-82s %s %s" % (h, d[idx[h]], units[idx[h]]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
