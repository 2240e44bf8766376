#!/usr/bin/env python
"""Static evidence for profiles/: per-kernel register / spill table from the ptxas logs and a SASS instruction histogram
(cuobjdump -sass) of the named kernels -- what each kernel actually issues (IMAD.WIDE vs IMAD vs IADD3, loads, shuffles,
barriers, calls), plus the head of its listing.  usage: python scripts/sass_digest.py OUTDIR"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "0g-halo2_b200", "csrc")
KERNELS = ["msm_accumulate_kernel", "ntt_pass_kernel", "k_h_lookup", "k_h_gates", "k_h_permutation", "msm_finish_kernel",
           "msm_bucket_l1_kernel", "msm_digits_kernel", "k_rs_scatter", "k_bi2_apply", "fp_mul_outlined", "fp_sqr_gen_outlined",
           "fp_mul2_gen_outlined"]


def registers_table(out):
    rows = []
    for log in sorted(glob.glob(os.path.join(CSRC, "*.ptxas.log"))):
        name = None
        for line in open(log):
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                name = m.group(1)
            m2 = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m2:
                stack = m2.groups()
            m3 = re.search(r"Used (\d+) registers", line)
            if m3 and name:
                dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
                dem = re.sub(r"\(.*", "", dem).replace("zg::", "").replace("(anonymous namespace)::", "")
                rows.append((os.path.basename(log).replace(".o.ptxas.log", ".cu"), dem[:70], int(m3.group(1)), stack))
                name = None
    with open(out, "w") as f:
        f.write("ptxas -v, sm_100a (nvcc 12.9): registers per thread, stack frame / spill stores / spill loads in bytes\n")
        f.write("%-16s %-72s %5s  %s\n" % ("file", "kernel", "regs", "stack/spill-st/spill-ld"))
        for r in sorted(rows, key=lambda r: (r[0], -r[2])):
            f.write("%-16s %-72s %5d  %s/%s/%s\n" % (r[0], r[1], r[2], *r[3]))


def sass_digests(outdir):
    for obj in sorted(glob.glob(os.path.join(CSRC, "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        for blk in txt.split("\t\tFunction : ")[1:]:
            fn = blk.split("\n", 1)[0].strip()
            hit = [k for k in KERNELS if k in fn]
            if not hit:
                continue
            ops = collections.Counter()
            lines = []
            for line in blk.splitlines():
                m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
                if m:
                    ops[m.group(1)] += 1
                    lines.append(line.rstrip())
            total = sum(ops.values())
            short = re.sub(r"[^A-Za-z0-9_]", "_", hit[0])
            tag = "_" + re.sub(r"\.o$", "", os.path.basename(obj)) if "outlined" in hit[0] else ""
            with open(os.path.join(outdir, "r02_sass_%s%s.txt" % (short, tag)), "w") as f:
                f.write("%s\n%s: %d SASS instructions (%d bytes)\n\n" % (fn, os.path.basename(obj), total, total * 16))
                grp = collections.Counter()
                for op, c in ops.items():
                    grp[".".join(op.split(".")[:2]) if op.startswith("IMAD") else op.split(".")[0]] += c
                for op, c in grp.most_common(24):
                    f.write("  %-14s %6d  %5.1f%%\n" % (op, c, 100.0 * c / total))
                f.write("\nfirst 80 instructions:\n" + "\n".join(lines[:80]) + "\n")


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")
    registers_table(os.path.join(out, "r02_ptxas_registers.txt"))
    sass_digests(out)
