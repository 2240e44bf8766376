#!/usr/bin/env python
"""bench.py -- headline benchmark of the zg_b200 proving backend (contract: see repo prompt).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workloads (BASELINE.json metric: "WNN MNIST proofs/sec & proof latency; MSM pts/s, NTT GB/s"):
  msm   one 2^LOGN-point BN254 G1 MSM per step (uniform scalars)       -> points/s
  ntt   one 2^LOGN-point BN254 Fr NTT per step                         -> GB/s (algorithmic)
`value` is timed with inputs resident in HBM (CUDA events on the launching stream, L2 flushed
between steps); `e2e` goes through the host-pointer C-ABI call (pinned host buffers, H2D and D2H
inside the timed region).  One process per GPU; N > 1 shards independent work units across ranks
(weak scaling, no data-path collective), time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "0g-halo2_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ZG_BENCH_WORKLOAD", "msm"))
    ap.add_argument("--logn", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling during the timed region (recipe in B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def rand_fr(n, seed):
    import cpu_ref  # generating synthetic inputs (host side, outside every timed region)
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 60) - 1)
    return cpu_ref.fr_to_mont(raw)


def synth_bases(n):
    import bn254
    import cpu_ref
    gen = bn254.g1_affine_to_limbs([bn254.G1_GEN])[0]
    return cpu_ref.g1_sequence(gen, n)


def ntt_bytes(logn):
    # 64 B per element per pass, ceil(log2 N / 12) passes
    return 64 * (1 << logn) * ((logn + 11) // 12)


def msm_imad(n, c):
    # ceil(254/c) * (10 N + 2 * 2^(c-1) * 14) * 264 IMAD-class instructions
    return ((254 + c - 1) // c) * (10 * n + 2 * (1 << (c - 1)) * 14) * 264


def run_reference(args, rank, world):
    """CPU arm: the oracle's restatement of upstream best_multiexp / best_fft on all host cores
    (the Rust reference cannot be built in this image: kind = "port")."""
    if rank != 0:
        return
    import bn254
    import cpu_ref
    n = 1 << args.logn
    cores = cpu_ref.num_threads()
    if args.workload == "msm":
        bases, sc = synth_bases(n), rand_fr(n, 7)
        fn = lambda: cpu_ref.best_multiexp(sc, bases)
        metric, unit, units = "msm_points_per_s", "points/s", n
    else:
        a, w = rand_fr(n, 11), bn254.fr_to_limbs([bn254.omega(args.logn)])
        fn = lambda: cpu_ref.best_fft(a, w, args.logn)
        metric, unit, units = "ntt_algorithmic_gbs", "GB/s", ntt_bytes(args.logn) / 1e9
    for _ in range(min(args.warmup, 1)):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    v = units / dt
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 (4x64 Montgomery)", "data": "synthetic",
        "config": {"workload": "%s 2^%d" % (args.workload, args.logn)},
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port",
                         "sample": "full workload, %d steps" % args.steps},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import zg_b200
    import bn254

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a non-default torch stream: the context enqueues on it, so torch CUDA events time our kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = zg_b200.Context(local, stream.cuda_stream)
    hbm_peak, peak_src = peaks()
    logn, n = args.logn, 1 << args.logn
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---- set-up (untimed) ----
    if args.workload == "msm":
        bases = synth_bases(n)
        ctx.srs_load(logn, bases, None)
        sc_host_t = torch.from_numpy(rand_fr(n, 7 + rank).view(np.int64)).pin_memory()
        sc_host = sc_host_t.numpy().view(np.uint64)
        sc_dev = sc_host_t.cuda()
        out_dev = torch.zeros(12, dtype=torch.int64, device="cuda")
        c = int(os.environ.get("ZG_MSM_C", "0")) or max(8, min(16, logn - 2))
        step_dev = lambda: ctx.msm_dev(0, sc_dev.data_ptr(), n, n, 1, out_dev.data_ptr())
        step_e2e = lambda: ctx.msm(0, sc_host)
        metric, unit, units = "msm_points_per_s", "points/s", n
        h2d, d2h = n * 32, 96
        dom_kernel = "msm_serial_reduce_kernel<true>"
    elif args.workload == "ntt":
        a_host_t = torch.from_numpy(rand_fr(n, 11 + rank).view(np.int64)).pin_memory()
        a_host = a_host_t.numpy().view(np.uint64)
        a_dev = a_host_t.cuda()
        b_dev = torch.empty_like(a_dev)
        w = bn254.fr_to_limbs([bn254.omega(logn)])
        step_dev = lambda: ctx.ntt_dev(a_dev.data_ptr(), b_dev.data_ptr(), logn, w)
        step_e2e = lambda: ctx.ntt(a_host, logn, w)
        metric, unit, units = "ntt_algorithmic_gbs", "GB/s", ntt_bytes(logn) / 1e9
        h2d = d2h = n * 32
        dom_kernel = "ntt_pass_kernel"
    else:
        raise SystemExit("unknown workload %s" % args.workload)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps):
        """per-step CUDA events on the launching stream, L2 flushed between steps"""
        tot = 0.0
        for _ in range(steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        return tot  # ms

    for _ in range(max(args.warmup, 3)):
        step_dev()
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ms = timed(step_dev, args.steps)
    barrier()
    launches = ctx.launch_count - l0
    # e2e: host buffers through the plain C-ABI call (H2D + compute + D2H), wall-clocked around a
    # synchronising call
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms / args.steps
    value = units * world / (ms_per_step * 1e-3)
    e2e_value = units * world / (e2e_ms / args.steps * 1e-3)
    # roofline of the dominant kernel
    imad_peak = ctx.bench_int_pipe(0, 4096)
    imad_wide = ctx.bench_int_pipe(1, 4096)
    mulmod_rate = ctx.bench_int_pipe(2, 256)
    mulmod_ptx_rate = ctx.bench_int_pipe(3, 256)
    if args.workload == "ntt":
        ach = ntt_bytes(logn) / 1e9 / (ms_per_step * 1e-3)
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": None, "peak_source": peak_src, "kernel": dom_kernel,
                "note": "254-bit Montgomery butterflies make this kernel integer-pipe bound; see int_pipe"}
        mulmods = (n // 2) * logn
    else:
        ach = msm_imad(n, c) / 1e9 / (ms_per_step * 1e-3)
        roof = {"bound": "int", "achieved": ach, "peak": imad_peak, "unit": "GIMAD/s", "frac": ach / imad_peak,
                "traffic": None, "peak_source": "measured in this run (zg_bench_int_pipe kind 0)",
                "kernel": dom_kernel, "window_c": c}
        mulmods = None
    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 (8x32-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": "%s 2^%d, BN254, uniform scalars" % (args.workload, logn), "l2": "flushed between steps"},
        "roofline": roof,
        "int_pipe": {"imad_gops": imad_peak, "imad_wide_gops": imad_wide, "fr_mulmod_portable_gops": mulmod_rate,
                     "fr_mulmod_ptx_gops": mulmod_ptx_rate},
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if mulmods:
        out["int_pipe"]["kernel_mulmod_gops"] = mulmods / 1e9 / (ms_per_step * 1e-3)
    if not args.no_cpu_baseline:
        import cpu_ref
        cores = cpu_ref.num_threads()
        if args.workload == "msm":
            fn = lambda: cpu_ref.best_multiexp(sc_host, bases)
        else:
            fn = lambda: cpu_ref.best_fft(a_host, w, logn)
        fn()
        reps, t0 = 0, time.perf_counter()
        while reps < 3 or (time.perf_counter() - t0 < 5 and reps < 20):
            fn()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        out["cpu_baseline"] = {"value": units / dt, "unit": unit, "cores": cores, "kind": "port",
                               "sample": "same workload, %d repetitions on host cores (oracle restatement of upstream)" % reps}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
