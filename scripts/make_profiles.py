#!/usr/bin/env python
"""Copies the round's evidence from gpurun_out/{final,ncu,multi_n*}/ into profiles/ (tracked) and writes profiles/README.md
and profiles/ncu_traffic.json.  Usage: python scripts/make_profiles.py r02"""
import csv
import glob
import io
import json
import os
import shutil
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import launch_summary  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
SRC = os.path.join(OUT, "final")
NCU = os.path.join(OUT, "ncu")
DST = os.path.join(ROOT, "profiles")


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def capture(fn, *a):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn(*a)
    return buf.getvalue()


def ncu_traffic(raw_csv, kernel):
    """average dram read+write bytes per launch over the captured launches of `kernel` in an exported raw page"""
    rows = list(csv.reader(l for l in open(raw_csv) if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = [float(d[ir]) * scale[units[ir]] + float(d[iw]) * scale[units[iw]] for d in data if kernel in d[ik]]
    return (sum(tot) / len(tot), len(tot)) if tot else (None, 0)


def main(tag):
    os.makedirs(DST, exist_ok=True)
    L = ["# profiles/ — measured evidence, round %s" % tag.lstrip("r"), "",
         "All numbers from one B200 (`gpurun`), SM clock 1965 MHz unless the JSON's `clocks` says otherwise; bench JSON lines are "
         "copied verbatim (`%s_bench_*.json`), launch lists are `ncu --metrics gpu__time_duration.sum --clock-control none` passes "
         "of the SAME bench command (per-launch times are cold-cache and serialised: use the kernel's SHARE), `%s_ncu_full_*.txt` "
         "are `scripts/ncu_summary.py` digests of `ncu --set full` captures, `%s_sass_*.txt` / `%s_ptxas_registers.txt` come from "
         "`scripts/sass_digest.py`.  Round-1 files (`r01_*`) are kept for comparison." % (tag, tag, tag, tag), ""]
    # ---- proof lines
    L += ["## Proof workload: proofs/s, image -> proof e2e, latency, CPU baseline", "",
          "`bench.py` default = the north-star target (49input_8192entry_4hash_6bpi shape, k = 17, synthetic stand-in), 4 proofs in "
          "flight, 16 proofs per step; e2e = image -> `zg_wnn_synthesize` -> pinned buffers -> `zg_create_proof` -> proof bytes "
          "through `ProofService.prove_many`, production RNG.", "",
          "| model | k | in flight | proofs/s (advice in HBM) | proofs/s (e2e from the image) | single-proof latency ms | kernel frac | CPU oracle proofs/s (cores) | e2e / CPU |",
          "|---|---|---|---|---|---|---|---|---|"]
    files = [("large (default)", "bench_default.json"), ("large, one lane", "bench_proof_large_1lane.json"), ("medium", "bench_proof_medium.json"),
             ("small", "bench_proof_small.json"), ("tiny", "bench_proof_tiny.json")]
    for name, fn in files:
        f = os.path.join(SRC, fn)
        if not os.path.exists(f):
            continue
        d = last_json(f)
        shutil.copy(f, os.path.join(DST, "%s_%s" % (tag, fn)))
        cb = d.get("cpu_baseline") or {}
        k = d["config"]["workload"].split("k=")[1].split(")")[0]
        L.append("| %s | %s | %d | %.1f | %.1f | %.2f | %.3f | %s | %s |" % (
            name, k, d.get("inflight", 1), d["value"], d["e2e"]["value"], d["latency_ms_single_proof"], d["roofline"]["frac"],
            "%.3f (%d)" % (cb["value"], cb["cores"]) if cb else "-", "%.0fx" % (d["e2e"]["value"] / cb["value"]) if cb else "-"))
    f = os.path.join(SRC, "bench_default.json")
    if os.path.exists(f):
        d = last_json(f)
        L += ["", "Stages of the last proof of the default run (ms): " + ", ".join("%s %.2f" % kv for kv in d["stage_ms_last_proof"].items()) + "."]
    for name, fn, what in [("verify", "bench_verify_small.json", "`Wnn::verify_proof` of one k = 15 proof (host code, one core)"),
                           ("keygen", "bench_keygen_large.json", "`keygen_vk` + `keygen_pk` at k = 17 on the device (host columns in)")]:
        f = os.path.join(SRC, fn)
        if os.path.exists(f):
            d = last_json(f)
            shutil.copy(f, os.path.join(DST, "%s_%s" % (tag, fn)))
            L += ["", "`bench.py --workload %s`: %s: %.2f ms (%.1f %s)." % (name, what, d["ms_per_step"], d["value"], d["unit"])]
    # ---- roofline of the dominant kernel
    L += ["", "## Dominant kernel (`msm_accumulate_kernel`): live CUDA-event timing against the measured IMAD peak", "",
          "| workload | launches | avg launch ms | point additions / launch | achieved GIMAD/s | peak GIMAD/s | frac | share of timed region |",
          "|---|---|---|---|---|---|---|---|"]
    for name, fn in files:
        f = os.path.join(SRC, fn)
        if os.path.exists(f):
            r = last_json(f)["roofline"]
            L.append("| proof %s (single-lane leg) | %d | %.3f | %.3g | %.0f | %.0f | %.3f | %.2f |" % (
                name, r["launches"], r["avg_launch_ms"], r["point_additions_per_launch"], r["achieved"], r["peak"], r["frac"],
                r["share_of_timed_region"]))
    sweep = os.path.join(SRC, "sweep.jsonl")
    if os.path.exists(sweep):
        shutil.copy(sweep, os.path.join(DST, "%s_sweep_msm_ntt.jsonl" % tag))
        rows = [json.loads(l) for l in open(sweep)]
        for d in rows:
            if d["metric"].startswith("msm"):
                r = d["roofline"]
                L.append("| %s | %d | %.3f | %.3g | %.0f | %.0f | %.3f | %.2f |" % (
                    d["config"]["workload"].split(",")[0], r["launches"], r["avg_launch_ms"], r["point_additions_per_launch"],
                    r["achieved"], r["peak"], r["frac"], r["share_of_timed_region"]))
        L += ["", "## MSM / NTT sweeps (BASELINE configs[4])", "",
              "| workload | ms | value | unit | e2e value (host buffers) | CPU oracle (16 cores) |", "|---|---|---|---|---|---|"]
        for d in rows:
            cb = d.get("cpu_baseline")
            L.append("| %s | %.3f | %.4g | %s | %.4g | %s |" % (
                d["config"]["workload"].split(",")[0], d["ms_per_step"], d["value"], d["unit"], d["e2e"]["value"],
                "%.4g" % cb["value"] if cb else "-"))
        ntt = [d for d in rows if d["metric"].startswith("ntt")]
        if ntt:
            L += ["", "NTT against the HBM roofline (the contract's bound, %s GB/s) and against the INTEGER roofline (the binding one, "
                  "DESIGN.md section 3: kernel mulmod rate / measured `fr_mulmod_evenodd_gops`):" % ntt[0]["roofline"]["peak"], ""]
            L.append(", ".join("2^%s: HBM %.3f, integer %.1f / %.1f G mulmod/s = %.2f" % (
                d["config"]["workload"].split("^")[1].split(",")[0], d["roofline"]["frac"], d["int_pipe"]["kernel_mulmod_gops"],
                d["int_pipe"]["fr_mulmod_evenodd_gops"], d["int_pipe"]["kernel_mulmod_gops"] / d["int_pipe"]["fr_mulmod_evenodd_gops"])
                for d in ntt))
    # ---- launch list
    f = os.path.join(NCU, "launches_proof_large.csv")
    if os.path.exists(f):
        shutil.copy(f, os.path.join(DST, "%s_launches_proof_large.csv" % tag))
        nproofs = 7.0
        txt = capture(launch_summary.main, f, nproofs)
        open(os.path.join(DST, "%s_launches_proof_large_summary.txt" % tag), "w").write(txt)
        L += ["", "## Launch list, k = 17 proof (one lane; 1500 launches = 7 proofs, times per proof; %s_launches_proof_large.csv)" % tag, "", "```"]
        L += txt.splitlines()[:26] + ["```"]
    # ---- ncu full digests
    traffic = {}
    if os.path.exists(os.path.join(DST, "ncu_traffic.json")):
        traffic = json.load(open(os.path.join(DST, "ncu_traffic.json")))
    for name in ["accumulate", "ntt", "hkern", "tail"]:
        txt = os.path.join(NCU, "ncu_full_%s.txt" % name)
        if os.path.exists(txt) and os.path.getsize(txt) > 100:
            shutil.copy(txt, os.path.join(DST, "%s_ncu_full_%s_proof_large.txt" % (tag, name)))
    raw = os.path.join(NCU, "accumulate_raw.csv")
    if os.path.exists(raw):
        t, cnt = ncu_traffic(raw, "msm_accumulate")
        if t:
            traffic["proof_large"] = t
            L += ["", "ncu `--set full`, `msm_accumulate_kernel` inside the k = 17 proof: average dram read+write per launch over %d "
                  "captured launches = %.1f MB (`profiles/ncu_traffic.json`, reported as `roofline.traffic`); digest "
                  "`%s_ncu_full_accumulate_proof_large.txt`." % (cnt, t / 1e6, tag)]
    src_csv = os.path.join(NCU, "accumulate_source_head.csv")
    if os.path.exists(src_csv):
        shutil.copy(src_csv, os.path.join(DST, "%s_ncu_source_accumulate_head.csv" % tag))
    if traffic:
        json.dump(traffic, open(os.path.join(DST, "ncu_traffic.json"), "w"), indent=1)
    # ---- multi-GPU
    multi = sorted(glob.glob(os.path.join(OUT, "multi_n*", "*.json")))
    if multi:
        L += ["", "## Multi-GPU (one process per GPU, `torchrun`; NCCL inside the library)", "",
              "| line | GPUs | ms/step | value | e2e | single-proof latency ms | scaling |", "|---|---|---|---|---|---|---|"]
        for f in multi:
            d = last_json(f)
            n = os.path.basename(os.path.dirname(f))
            shutil.copy(f, os.path.join(DST, "%s_%s_%s" % (tag, n, os.path.basename(f))))
            lat = d.get("latency_ms_single_proof")
            L.append("| %s | %d | %.3f | %.4g %s | %.4g | %s | %s |" % (
                os.path.basename(f).replace(".json", ""), d["n_gpus"], d["ms_per_step"], d["value"], d["unit"], d["e2e"]["value"],
                "%.2f" % lat if lat else "-", d["scaling"]))
    extra = os.path.join(DST, "%s_notes.md" % tag)
    if os.path.exists(extra):
        L += ["", open(extra).read()]
    open(os.path.join(DST, "README.md"), "w").write("\n".join(L) + "\n")
    print("\n".join(L))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
