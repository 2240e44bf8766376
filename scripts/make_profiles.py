#!/usr/bin/env python
"""Copies the round's evidence from gpurun_out/final/ into profiles/ (tracked) and writes profiles/README.md and
profiles/ncu_traffic.json.  Usage: python scripts/make_profiles.py r01"""
import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import launch_summary  # noqa: E402
import ncu_summary  # noqa: E402

SRC = os.path.join(ROOT, "gpurun_out", "final")
DST = os.path.join(ROOT, "profiles")


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def capture(fn, *a):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def ncu_traffic(rep):
    """average dram read+write bytes per launch over the captured launches of a report"""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = [float(d[ir]) * scale[units[ir]] + float(d[iw]) * scale[units[iw]] for d in data]
    return sum(tot) / len(tot), len(tot)


def main(tag):
    os.makedirs(DST, exist_ok=True)
    lines = ["# profiles/ — measured evidence, round %s" % tag.lstrip("r"), "",
             "All numbers from one B200 (`gpurun`), SM clock 1965 MHz unless the JSON's `clocks` says otherwise; bench JSON "
             "lines are copied verbatim (`*_bench_*.json`), ncu launch lists are `--metrics gpu__time_duration.sum "
             "--clock-control none` passes of the SAME bench command (per-launch times are cold-cache and serialised: "
             "use the kernel's SHARE), `*_ncu_full_*.txt` are `scripts/ncu_summary.py` digests of `ncu --set full` captures.", ""]
    # ---- bench lines
    lines += ["## Proof workload (BASELINE configs[0..3]): proofs/s, latency, CPU baseline", "",
              "| model | k | in flight | proofs/s (device) | proofs/s (e2e, host buffers) | single-proof latency ms | CPU oracle proofs/s (cores) | e2e / CPU |",
              "|---|---|---|---|---|---|---|---|"]
    for m in ["tiny", "small", "medium", "large"]:
        f = os.path.join(SRC, "bench_proof_%s.json" % m)
        if not os.path.exists(f):
            continue
        d = last_json(f)
        shutil.copy(f, os.path.join(DST, "%s_bench_proof_%s.json" % (tag, m)))
        cb = d.get("cpu_baseline") or {}
        k = d["config"]["workload"].split("k=")[1].split(")")[0]
        lines.append("| %s | %s | %d | %.1f | %.1f | %.2f | %s | %s |" % (
            m, k, d.get("inflight", 1), d["value"], d["e2e"]["value"], d["latency_ms_single_proof"],
            "%.3f (%d)" % (cb["value"], cb["cores"]) if cb else "-", "%.0fx" % (d["e2e"]["value"] / cb["value"]) if cb else "-"))
    lines += ["", "Proofs in flight per GPU (small model):", "", "| in flight | proofs/s | ms per batch |", "|---|---|---|"]
    for k in [1, 2, 4, 6, 8]:
        f = os.path.join(SRC, "bench_proof_small_inflight%d.json" % k) if k != 4 else os.path.join(SRC, "bench_proof_small.json")
        if os.path.exists(f):
            d = last_json(f)
            lines.append("| %d | %.1f | %.2f |" % (d.get("inflight", k), d["value"], d["ms_per_step"]))
    f = os.path.join(SRC, "bench_reference_small.json")
    if os.path.exists(f):
        d = last_json(f)
        shutil.copy(f, os.path.join(DST, "%s_bench_reference_small.json" % tag))
        lines += ["", "`bench.py --impl reference` (CPU oracle, %d cores): %.3f proofs/s (%.0f ms per proof)." % (
            d["cpu_baseline"]["cores"], d["value"], d["ms_per_step"])]
    # ---- roofline of the dominant kernel
    lines += ["", "## Dominant kernel (`msm_accumulate_kernel`): live CUDA-event timing against the measured IMAD peak", "",
              "| workload | launches | avg launch ms | point additions / launch | achieved GIMAD/s | peak GIMAD/s | frac | share of timed region |",
              "|---|---|---|---|---|---|---|---|"]
    for m in ["small", "large"]:
        f = os.path.join(SRC, "bench_proof_%s.json" % m)
        if os.path.exists(f):
            r = last_json(f)["roofline"]
            lines.append("| proof %s (single lane) | %d | %.3f | %.3g | %.0f | %.0f | %.3f | %.2f |" % (
                m, r["launches"], r["avg_launch_ms"], r["point_additions_per_launch"], r["achieved"], r["peak"], r["frac"],
                r["share_of_timed_region"]))
    sweep = os.path.join(SRC, "sweep.jsonl")
    if os.path.exists(sweep):
        shutil.copy(sweep, os.path.join(DST, "%s_sweep_msm_ntt.jsonl" % tag))
        rows = [json.loads(l) for l in open(sweep)]
        for d in rows:
            if d["metric"].startswith("msm"):
                r = d["roofline"]
                lines.append("| %s | %d | %.3f | %.3g | %.0f | %.0f | %.3f | %.2f |" % (
                    d["config"]["workload"].split(",")[0], r["launches"], r["avg_launch_ms"], r["point_additions_per_launch"],
                    r["achieved"], r["peak"], r["frac"], r["share_of_timed_region"]))
        lines += ["", "## MSM / NTT sweeps (BASELINE configs[4])", "",
                  "| workload | ms | value | unit | whole-step frac of its SURVEY 8(d) roofline | e2e value | CPU oracle (16 cores) |", "|---|---|---|---|---|---|---|"]
        for d in rows:
            cb = d.get("cpu_baseline")
            fr = d.get("step_roofline", d["roofline"])["frac"]
            lines.append("| %s | %.3f | %.4g | %s | %.3f | %.4g | %s |" % (
                d["config"]["workload"].split(",")[0], d["ms_per_step"], d["value"], d["unit"], fr, d["e2e"]["value"],
                "%.4g" % cb["value"] if cb else "-"))
        ntt = [d for d in rows if d["metric"].startswith("ntt")]
        if ntt:
            lines += ["", "NTT against the INTEGER roofline (the binding one, DESIGN.md section 3): kernel mulmod rate / measured "
                      "`fr_mulmod_evenodd_gops`:", ""]
            lines.append(", ".join("2^%s: %.1f / %.1f G mulmod/s = %.2f" % (
                d["config"]["workload"].split("^")[1].split(",")[0], d["int_pipe"]["kernel_mulmod_gops"],
                d["int_pipe"]["fr_mulmod_evenodd_gops"], d["int_pipe"]["kernel_mulmod_gops"] / d["int_pipe"]["fr_mulmod_evenodd_gops"])
                for d in ntt))
    # ---- launch lists
    for m in ["small", "large"]:
        f = os.path.join(SRC, "launches_proof_%s.csv" % m)
        if os.path.exists(f):
            shutil.copy(f, os.path.join(DST, "%s_launches_proof_%s.csv" % (tag, m)))
            txt = capture(launch_summary.main, f, 1.0)
            open(os.path.join(DST, "%s_launches_proof_%s_summary.txt" % (tag, m)), "w").write(txt)
            lines += ["", "## Launch list, proof %s (one lane, %s)" % (m, "%s_launches_proof_%s.csv" % (tag, m)), "", "```"]
            lines += txt.splitlines()[:24] + ["```"]
    # ---- ncu full captures
    traffic = {}
    for name, key in [("prof_accumulate_proof_small", "proof_small"), ("prof_accumulate_msm20", "msm_20"), ("prof_ntt20", None)]:
        rep = os.path.join(SRC, name + ".ncu-rep")
        if not os.path.exists(rep):
            continue
        txt = capture(ncu_summary.main, rep)
        open(os.path.join(DST, "%s_ncu_full_%s.txt" % (tag, name.replace("prof_", ""))), "w").write(txt)
        if key:
            t, cnt = ncu_traffic(rep)
            traffic[key] = t
            lines += ["", "ncu `--set full`, %s: average dram read+write per launch over %d captured launches = %.1f MB "
                      "(`profiles/ncu_traffic.json`, reported as `roofline.traffic`)." % (name, cnt, t / 1e6)]
    if traffic:
        json.dump(traffic, open(os.path.join(DST, "ncu_traffic.json"), "w"), indent=1)
    for f in glob.glob(os.path.join(SRC, "multi_*.json")):
        shutil.copy(f, os.path.join(DST, "%s_%s" % (tag, os.path.basename(f))))
    extra = os.path.join(DST, "%s_notes.md" % tag)
    if os.path.exists(extra):
        lines += ["", open(extra).read()]
    open(os.path.join(DST, "README.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
