#!/usr/bin/env python
"""bench.py -- headline benchmark of the zg_b200 proving backend (contract: see repo prompt).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

BASELINE.json metric: "WNN MNIST proofs/sec & proof latency; MSM pts/s, NTT GB/s vs roofline".
Workloads:
  proof (default)  zero_g WNN proofs, create_proof (KZG/BN254, GWC, EvmTranscript) for `--model` on
                   benches/example_image_7.png.  Default model = large: the north-star target, the
                   49input_8192entry_4hash_6bpi shape at k = 17 (synthetic stand-in, the real file is absent upstream;
                   BASELINE configs[3] on one GPU); small / tiny / medium = configs[1] / [0] / [2].  A step =
                   `--inflight` lanes x `--proofs-per-lane` proofs.  `value`: advice columns resident in HBM;
                   `e2e`: image -> native witness synthesis -> proof bytes (what benches/bench.rs:30-36 times).
  msm              one 2^LOGN-point BN254 G1 MSM per step (uniform scalars)      -> points/s
  keygen           keygen_vk + keygen_pk of `--model` on the device (benches/bench.rs bench_key_generation)  -> keygens/s
  verify           Wnn::verify_proof of one proof of `--model` (benches/bench.rs bench_verification; host code) -> verifications/s
  msm_sharded      ONE 2^LOGN-point MSM split by point range over the ranks, partial sums all-gathered (NCCL)
                   and added (strong scaling, BASELINE configs[3])                -> points/s
  ntt              one 2^LOGN-point BN254 Fr NTT per step                        -> GB/s (algorithmic)
`value` is timed with inputs resident in HBM (CUDA events on the launching stream, L2 flushed between
steps); `e2e` goes through the host-pointer C-ABI call (pinned host buffers in, result bytes out).
One process per GPU; N > 1 proves / transforms independent units per rank (weak scaling, no data-path
collective), time = max over ranks.
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

GOLD = os.path.join(ROOT, "tests", "golden")
MODELS = {
    "tiny": ("model_28input_256entry_1hash_1bpi.hdf5", 14),
    "small": ("model_28input_1024entry_2hash_2bpi.hdf5", 15),
    "medium": ("model_28input_2048entry_2hash_3bpi.hdf5", 15),
    "large": (None, 17),          # synthetic same-shape stand-in (the real file is absent upstream)
}
SRS_SECRET = 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ZG_BENCH_WORKLOAD", "proof"))
    ap.add_argument("--model", default=os.environ.get("ZG_BENCH_MODEL", "large"), choices=list(MODELS),
                    help="proof / keygen workloads: large (default) = the north-star target, the 49input_8192entry_4hash_6bpi "
                         "shape at k = 17 (BASELINE configs[3] on one GPU); small = configs[1]; tiny = configs[0]; medium = configs[2]")
    ap.add_argument("--logn", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("ZG_BENCH_INFLIGHT", "4")),
                    help="proof workload: independent proofs in flight per GPU (one context + stream + host thread each); "
                         "a step is one batch of that many proofs")
    ap.add_argument("--synth", default=os.environ.get("ZG_BENCH_SYNTH", "native"), choices=["cached", "native"],
                    help="proof workload, e2e leg: `native` (default) = every e2e proof starts from the IMAGE, as "
                         "benches/bench.rs:30-36 `wnn.proof(&pk, &params, &img)` does: zg_wnn_synthesize (C++ host code) fills the "
                         "lane's pinned advice buffers inside the timed region; `cached` = witnesses synthesized once, untimed")
    ap.add_argument("--proofs-per-lane", type=int, default=int(os.environ.get("ZG_BENCH_PPL", "4")),
                    help="proof workload: proofs every lane proves back to back in one step (a step = inflight x this many "
                         "proofs); 4 = 16 proofs per step with 4 lanes: the lanes stagger inside a step as they do in a long-running service")
    ap.add_argument("--shard", default=os.environ.get("ZG_BENCH_SHARD", "proofs"), choices=["proofs", "columns"],
                    help="proof workload with N > 1 GPUs: `proofs` = independent proofs per GPU (weak scaling, no data-path "
                         "collective); `columns` = ONE stream of proofs, every round's commitment MSMs spread over the ranks by "
                         "column and all-gathered over NCCL (strong scaling of one proof, BASELINE configs[3]); one lane")
    ap.add_argument("--images", type=int, default=int(os.environ.get("ZG_BENCH_IMAGES", "1")),
                    help="proof workload: 1 = benches/example_image_7.png for every proof (the reference's bench input); "
                         "K > 1 = K distinct synthetic MNIST-shaped images per rank, proofs cycle through them "
                         "(BASELINE configs[4]); witnesses are synthesized once, untimed")
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
        self.index, self.rows, self.proc = index, [], None

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
    import cpu_ref  # synthetic inputs (host side, outside every timed region)
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


# dram__bytes_read.sum + dram__bytes_write.sum per launch of msm_accumulate_kernel, from the `ncu --set full` captures
# summarised in profiles/ (see profiles/README.md); keyed by workload
NCU_TRAFFIC = {}
_traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
if os.path.exists(_traffic_path):
    NCU_TRAFFIC = json.load(open(_traffic_path))


def kernel_roofline(probe, imad_peak, region_ms, peak_src, traffic=None):
    """msm_accumulate_kernel alone, timed live with CUDA events around every launch (zg_probe_*): algorithmic work =
    one mixed XYZZ addition (10 Montgomery products = 2640 IMAD) per sorted (bucket, point) entry."""
    if not probe or not probe[1]:
        return None
    kms, launches, adds = probe
    ach = adds * 10 * 264 / 1e9 / (kms * 1e-3)
    return {"bound": "int", "achieved": ach, "peak": imad_peak, "unit": "GIMAD/s", "frac": ach / imad_peak,
            "traffic": traffic, "peak_source": peak_src, "kernel": "msm_accumulate_kernel", "launches": launches,
            "avg_launch_ms": kms / launches, "point_additions_per_launch": adds / launches,
            "algorithmic_gimad_per_launch": adds / launches * 2640 / 1e9, "share_of_timed_region": kms / region_ms,
            "note": "bound is the integer (IMAD) pipe, not HBM or tensor: 254-bit Montgomery arithmetic; traffic = ncu dram bytes "
                    "per launch (profiles/ncu_traffic.json) when captured for this workload"}


# ---- the proof workload: model, witness, SRS -----------------------------------------------------------
def load_model(name):
    from zg_b200.io import load_wnn, load_grayscale_image, synthetic_wnn
    fname, k = MODELS[name]
    wnn = synthetic_wnn() if fname is None else load_wnn(os.path.join(GOLD, fname))
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    return wnn, img, k


def workload_name(args):
    if args.workload == "proof":
        fname, k = MODELS[args.model]
        return "zero_g proof, %s (k=%d), example_image_7.png" % (fname or "synthetic 49input_8192entry_4hash_6bpi stand-in", k)
    return "%s 2^%d, BN254, uniform scalars" % (args.workload, args.logn)


def run_reference(args, rank, world):
    """CPU arm: the oracle's restatement of the upstream CPU prover (oracle/halo2_ref.py over
    oracle/zg_oracle.c, OpenMP on all host cores; real best_multiexp commitments).  The Rust reference
    cannot be built in this image, so kind = "port"."""
    if rank != 0:
        return
    import bn254
    import cpu_ref
    cores = cpu_ref.use_all_cores()
    if args.workload == "proof":
        import halo2_ref as H
        wnn, img, k = load_model(args.model)
        srs = H.Srs(k, SRS_SECRET)
        circ0, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
        pk = H.keygen(srs, circ0.cs, asm0)
        _, asm = wnn.synthesize(img, k)
        out = wnn.predict(img)
        seed = [0]

        def fn():
            seed[0] += 1
            return H.create_proof(srs, pk, asm.advice, [out], H.XorShiftRng(bytes([seed[0] % 256] * 16)), real_msm=True)
        metric, unit, units = "proofs_per_s", "proofs/s", 1
    elif args.workload in ("msm", "msm_sharded"):
        n = 1 << args.logn
        bases, sc = synth_bases(n), rand_fr(n, 7)
        fn = lambda: cpu_ref.best_multiexp(sc, bases)
        metric, unit, units = "msm_points_per_s", "points/s", n
    else:
        n = 1 << args.logn
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
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 (4x64-bit Montgomery)", "data": "synthetic",
        "config": {"workload": workload_name(args)},
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port",
                         "sample": "full workload per step, %d steps (witness synthesis excluded)" % args.steps},
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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    # a non-default torch stream: the context enqueues on it, so torch CUDA events time our kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = zg_b200.Context(local, stream.cuda_stream)
    hbm_peak, peak_src = peaks()
    logn, n = args.logn, 1 << args.logn
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    cpu_fn = None
    extra = {}

    # ---- set-up (untimed) ----
    if args.workload == "proof":
        import halo2_ref as H                      # test SRS with a known trapdoor + the checker
        from concurrent.futures import ThreadPoolExecutor
        from zg_b200.prover import ParamsKZG, keygen, advice_to_mont, create_proof_limbs
        from zg_b200.bn254_host import to_limbs
        wnn, img, k = load_model(args.model)
        n = 1 << k
        srs = H.Srs(k, SRS_SECRET)
        shard_cols = args.shard == "columns" and world > 1
        # the inputs every proof cycles through: (advice columns, public outputs) per image
        from zg_b200.io import synthetic_image
        nimg = max(1, args.images)
        imgs = [img] if nimg == 1 else [synthetic_image((0 if shard_cols else rank * nimg) + i, wnn.img_shape()) for i in range(nimg)]
        witnesses = []
        usable_rows = None
        for im in imgs:
            _, a = wnn.synthesize(im, k)
            usable_rows = a.usable_rows
            o = wnn.predict(im)
            witnesses.append((a.advice, o, [to_limbs(list(o))]))
        native = zg_b200.lib.NativeSynthesizer(wnn) if args.synth == "native" else None
        asm_advice, outputs, inst = witnesses[0]

        class DevCol:                              # quacks like a numpy column for create_proof_limbs
            def __init__(self, t):
                self.t = t
                self.ctypes = type("c", (), {"data": t.data_ptr()})
                self.shape = (t.shape[0],)

        # The lanes are the product's own throughput front-end (zg_b200/service.py::ProofService): one context + proving key
        # + host thread per lane, witness synthesis pipelined one image ahead of the proof on a helper thread per lane.
        # The e2e leg calls service.prove_many (image in, proof bytes out); the device leg proves from advice columns that
        # already sit in HBM on the same lanes.
        import itertools
        from zg_b200.service import ProofService
        K = 1 if shard_cols else max(1, args.inflight)
        PPL = max(1, args.proofs_per_lane)
        lane_streams = [stream] + [torch.cuda.Stream() for _ in range(K - 1)]
        contexts = [ctx] + [zg_b200.Context(local, lane_streams[i].cuda_stream) for i in range(1, K)]
        job_counter = itertools.count(1)
        # SPMD ranks must draw the same stream: a ChaCha20 stream keyed by the job number (same generator as production)
        seeded = lambda _job: zg_b200.lib.ChaCha20Rng.from_key(int(next(job_counter)).to_bytes(32, "little"))
        # e2e draws from the production RNG (OS-seeded ChaCha20, the reference's OsRng) unless the ranks must agree (SPMD)
        service = ProofService(wnn, ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2), device=local, lanes=K,
                               rng_factory=seeded if shard_cols else None, contexts=contexts)

        class Lane:
            """the device-resident leg of one service lane"""
            def __init__(self, idx, sl):
                self.ctx, self.pk = sl.ctx, sl.pk
                self.adv_host, self.adv_dev_t, self.adv_dev = [], [], []
                for adv, _, _ in witnesses:
                    ht = [torch.from_numpy(a.view(np.int64)).pin_memory() for a in advice_to_mont(sl.ctx, adv)]
                    dt = [t.cuda() for t in ht]
                    self.adv_host.append([t.numpy().view(np.uint64) for t in ht])
                    self.adv_dev_t.append((ht, dt))
                    self.adv_dev.append([DevCol(t) for t in dt])
                self.seed = (0 if shard_cols else rank * 100000) + idx * 1000     # SPMD: every rank draws the same stream
                self.turn = idx                      # which image this lane proves next

            def rng(self):
                self.seed += 1
                if shard_cols:
                    return zg_b200.lib.ChaCha20Rng.from_key(int(self.seed).to_bytes(32, "little"))
                return zg_b200.lib.XorShift.from_seed(int(self.seed).to_bytes(16, "little"))

            def _next(self):
                w = self.turn % len(witnesses)
                self.turn += K
                return w

            def prove_dev(self):
                w = self._next()
                return w, create_proof_limbs(self.pk, self.adv_dev[w], witnesses[w][2], self.rng())

            def prove_host_cols(self):               # --synth cached: witnesses synthesized ahead of time, host buffers
                w = self._next()
                return w, create_proof_limbs(self.pk, self.adv_host[w], witnesses[w][2], self.rng())

        lanes = [Lane(i, service.lanes[i]) for i in range(K)]
        pk = lanes[0].pk
        if shard_cols:
            from zg_b200 import farm
            farm.join_communicator(ctx, dist)
            ctx.set_distribution(1)
            extra["scaling_override"] = "strong"
            extra["shard"] = "one proof stream; every round's commitments spread over %d GPUs by column, all-gathered (NCCL)" % world
        pool = ThreadPoolExecutor(K) if K > 1 else None

        def lane_run(lane, method):
            return [getattr(lane, method)() for _ in range(PPL)]

        def run_all(method):
            if pool is None:
                return lane_run(lanes[0], method)
            return [r for f in [pool.submit(lane_run, l, method) for l in lanes] for r in f.result()]
        step_dev = lambda: run_all("prove_dev")
        e2e_turn = [0]

        def step_e2e():
            if native is None:
                return run_all("prove_host_cols")
            ws = [(e2e_turn[0] + i) % len(imgs) for i in range(K * PPL)]
            e2e_turn[0] += K * PPL
            return [(w, pr) for w, (pr, _) in zip(ws, service.prove_many([imgs[w] for w in ws]))]

        def run_e2e_stream(steps):
            ws = [(e2e_turn[0] + i) % len(imgs) for i in range(steps * K * PPL)]
            e2e_turn[0] += len(ws)
            service.prove_many([imgs[w] for w in ws])
        if native is not None:
            extra["run_e2e"] = run_e2e_stream
            extra["run_e2e_stream"] = True
        # every measured proof is a real proof: check one per lane against the restated verifier (untimed)
        opk = None
        if rank == 0 or shard_cols:                    # (column sharding: the proofs are collectives, every rank takes part)
            if rank == 0:
                circ_o, asm_o = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
                opk = H.keygen(srs, circ_o.cs, asm_o)
                assert pk.fixed_commitments == opk.fixed_commitments and pk.perm_commitments == opk.perm_commitments
                vparams = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
            seen = set()
            for _ in range((len(witnesses) + K * PPL - 1) // (K * PPL) + 1):
                for w, pr in step_e2e():
                    if rank == 0 and w not in seen:
                        assert H.verify_proof(srs, opk, [witnesses[w][1]], pr), "GPU proof rejected by the restated verifier"
                        assert pk.get_vk().verify(vparams, [witnesses[w][1]], pr), "GPU proof rejected by the product verifier"
                        seen.add(w)
            assert rank != 0 or len(seen) == len(witnesses), "not every image was proven during the check"
        metric, unit, units = "proofs_per_s", "proofs/s", K * PPL
        h2d, d2h = units * (len(lanes[0].adv_host[0]) * n * 32 + len(outputs) * 32), units * 3840
        if shard_cols:
            units = units / world                      # ONE proof stream for the whole job: value = proofs / time
        dom_kernel = "msm_accumulate_kernel"
        extra["inflight"] = K
        extra["proofs_per_step"] = K * PPL
        extra["images"] = "example_image_7.png" if nimg == 1 else "%d synthetic MNIST-shaped images per rank" % nimg
        extra["e2e_starts_from"] = ("image (native witness synthesis timed, pipelined one image ahead per lane by ProofService; the timed steps are one prove_many call)"
                                    if native is not None else "synthesized advice columns")
        extra["e2e_rng"] = ("ChaCha20 keyed by the job number (SPMD ranks must agree)" if shard_cols else
                            "seeded XorShift" if native is None else "ChaCha20 keyed from the OS per proof")
        if not args.no_cpu_baseline and rank == 0:
            cpu_fn = lambda: H.create_proof(srs, opk, asm_advice, [outputs], H.XorShiftRng(bytes(range(16))), real_msm=True)
            cpu_units = 1
    elif args.workload == "msm":
        bases = synth_bases(n)
        ctx.srs_load(logn, bases, None)
        sc_host_t = torch.from_numpy(rand_fr(n, 7 + rank).view(np.int64)).pin_memory()
        sc_host = sc_host_t.numpy().view(np.uint64)
        sc_dev = sc_host_t.cuda()
        out_dev = torch.zeros(12, dtype=torch.int64, device="cuda")
        step_dev = lambda: ctx.msm_dev(0, sc_dev.data_ptr(), n, n, 1, out_dev.data_ptr())
        step_e2e = lambda: ctx.msm(0, sc_host)
        metric, unit, units = "msm_points_per_s", "points/s", n
        h2d, d2h = n * 32, 96
        dom_kernel = "msm_accumulate_kernel"
        if not args.no_cpu_baseline and rank == 0:
            import cpu_ref
            cpu_fn = lambda: cpu_ref.best_multiexp(sc_host, bases)
    elif args.workload == "msm_sharded":
        # BASELINE configs[3]: ONE 2^logn MSM split by point range; rank g keeps bases [g n/G, (g+1) n/G) resident, gets
        # the matching scalar slice, and the G partial sums (96 B each) are all-gathered over NCCL and added (strong scaling)
        from zg_b200 import farm
        assert n % world == 0 and (n // world) & (n // world - 1) == 0, "world size must divide 2^logn into powers of two"
        lo, hi = farm.point_range(n, rank, world)
        bases = synth_bases(n)
        logn_loc = (hi - lo).bit_length() - 1
        ctx.srs_load(logn_loc, np.ascontiguousarray(bases[lo:hi]), None)
        sc_full = rand_fr(n, 7)
        sc_host_t = torch.from_numpy(np.ascontiguousarray(sc_full[lo:hi]).view(np.int64)).pin_memory()
        sc_host = sc_host_t.numpy().view(np.uint64)
        sc_dev = sc_host_t.cuda()
        out_dev = torch.zeros(12, dtype=torch.int64, device="cuda")
        result = [None]
        farm.join_communicator(ctx, dist if world > 1 else None)     # NCCL inside the library (csrc/dist.cu)

        def step_dev():      # local MSM over this rank's point range, ncclAllGather of 96 B, G - 1 adds on the device
            ctx.msm_sharded_dev(0, sc_dev.data_ptr(), hi - lo, hi - lo, 1, out_dev.data_ptr())

        def step_e2e():      # host scalars in, the total (96 B) back on the host of every rank
            result[0] = ctx.msm_sharded(0, sc_host)
        metric, unit, units = "msm_points_per_s", "points/s", n / world      # value = n / time (units * world below)
        h2d, d2h = (hi - lo) * 32, 96 * world
        dom_kernel = "msm_accumulate_kernel"
        extra["scaling_override"] = "strong"
        if logn <= 20:
            step_e2e()                                 # a collective: every rank takes part, rank 0 checks
            step_dev()
        if rank == 0 and logn <= 20:
            import cpu_ref
            exp = cpu_ref.g1_to_affine(cpu_ref.best_multiexp(sc_full, bases))[0]
            for got_jac in (result[0], out_dev.cpu().numpy().view(np.uint64)):
                got = cpu_ref.g1_to_affine(np.ascontiguousarray(got_jac, dtype=np.uint64).reshape(1, 12))[0]
                assert (np.asarray(got).reshape(-1) == np.asarray(exp).reshape(-1)).all(), "sharded MSM differs from the oracle"
        if not args.no_cpu_baseline and rank == 0:
            import cpu_ref
            cpu_fn = lambda: cpu_ref.best_multiexp(sc_full, bases)
            cpu_units = n
    elif args.workload == "ntt":
        a_host_t = torch.from_numpy(rand_fr(n, 11 + rank).view(np.int64)).pin_memory()
        a_host = a_host_t.numpy().view(np.uint64)
        a_dev = a_host_t.cuda()
        b_dev = torch.empty_like(a_dev)
        w = bn254.fr_to_limbs([bn254.omega(logn)])
        step_dev = lambda: ctx.ntt_dev(a_dev.data_ptr(), b_dev.data_ptr(), logn, w)
        step_e2e = lambda: ctx.ntt_inplace(a_host, logn, w)     # pinned buffer, transformed in place
        metric, unit, units = "ntt_algorithmic_gbs", "GB/s", ntt_bytes(logn) / 1e9
        h2d = d2h = n * 32
        dom_kernel = "ntt_pass_kernel"
        if not args.no_cpu_baseline and rank == 0:
            import cpu_ref
            cpu_fn = lambda: cpu_ref.best_fft(a_host, w, logn)
    elif args.workload == "verify":
        # benches/bench.rs:38-45 `bench_verification`: Wnn::verify_proof (src/wnn.rs:265-280) of one proof.  Host work in the
        # reference and here (csrc/verifier.cu: transcript, expression evaluation, one ~70-term MSM, two pairings); the GPU
        # only produces the proof that is verified.
        import halo2_ref as H
        from zg_b200.prover import ParamsKZG, keygen
        wnn, img, k = load_model(args.model)
        n = 1 << k
        srs = H.Srs(k, SRS_SECRET)
        params = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
        circ0, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
        vpk = keygen(ctx, params, circ0.cs, asm0)
        proof, outs = wnn.proof(vpk, params, img)
        vk = vpk.get_vk()
        assert vk.verify(params, [outs], proof)

        def step_dev():
            assert vk.verify(params, [outs], proof)
        step_e2e = step_dev
        metric, unit, units = "verifications_per_s", "verifications/s", 1
        h2d, d2h = 0, 0
        dom_kernel = None
        extra["note"] = "verification is host code by design (as in the reference); no kernel launches in the timed region"
    elif args.workload == "keygen":
        # keygen_vk + keygen_pk on the device (benches/bench.rs:24-28 `bench_key_generation`): commitments of the 16 fixed and
        # 8 sigma columns, their coefficient and extended-coset forms, l_0 / l_last / l_active.  Host circuit synthesis and
        # selector compression are done once, untimed (they stay on the host in the reference too).
        import halo2_ref as H
        from zg_b200.plonk.mock import finalize_fixed
        from zg_b200.plonk.serialize import serialize_cs
        from zg_b200.prover import ParamsKZG, _load_pk, ints_to_canonical
        wnn, img, k = load_model(args.model)
        n = 1 << k
        srs = H.Srs(k, SRS_SECRET)
        params = ParamsKZG(k, srs.g, srs.g_lagrange)
        params.load(ctx)
        circ0, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
        fixed_int = finalize_fixed(circ0.cs, asm0)
        words, constants = serialize_cs(circ0.cs)
        fixed_mont = [ctx.debug_field_op(0, 7, ints_to_canonical(c)) for c in fixed_int]
        mapping = np.ascontiguousarray(np.array(asm0.mapping, dtype=np.uint32).reshape(len(asm0.perm_cols), asm0.n, 2))

        def step_dev():
            h = _load_pk(ctx, params, circ0.cs, words, constants, fixed_mont, mapping, 1)
            ctx._L.zg_pk_free(ctx._h, h)
        step_e2e = step_dev                        # zg_pk_load always takes host columns
        metric, unit, units = "keygens_per_s", "keygens/s", 1
        h2d, d2h = (len(fixed_mont) + 2 * len(asm0.perm_cols) // 8) * n * 32, (len(fixed_mont) + len(asm0.perm_cols)) * 64
        dom_kernel = "msm_accumulate_kernel"
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
            torch.cuda.synchronize()               # every lane's stream is idle when the start event is recorded
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()                                  # (multi-lane steps end with every lane's stream drained)
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        return tot  # ms

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_dev()
    torch.cuda.synchronize()
    all_ctx = [l.ctx for l in lanes] if args.workload == "proof" else [ctx]
    l0 = sum(c.launch_count for c in all_ctx)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    probe = None
    if args.workload in ("msm", "msm_sharded"):
        ctx.probe_enable(True)
    barrier()
    ms = timed(step_dev, args.steps)
    barrier()
    if args.workload in ("msm", "msm_sharded"):
        probe = ctx.probe_read()
        ctx.probe_enable(False)
    launches = sum(c.launch_count for c in all_ctx) - l0
    stage = None
    if args.workload == "proof":
        # single-proof latency (one lane alone), the other half of BASELINE's "proofs/s & proof latency"
        ctx.probe_enable(True)
        lat_ms = timed(lambda: lanes[0].prove_dev(), args.steps) / args.steps
        probe = ctx.probe_read()
        ctx.probe_enable(False)
        stage = pk.stage_ms()
        extra["latency_ms_single_proof"] = lat_ms
    # e2e: host buffers through the plain C-ABI call (H2D + compute + D2H of the result)
    # (proof workload from the image: the timed steps are ONE ProofService.prove_many call over all their images -- the
    #  call a user with a queue of images makes -- so the witness pipeline fills once, not once per step)
    e2e_streams = extra.pop("run_e2e_stream", False)
    run_e2e = extra.pop("run_e2e", None) or (lambda steps: [step_e2e() for _ in range(steps)])
    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
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
    # integer-pipe denominators, measured in this run
    imad_peak = ctx.bench_int_pipe(0, 4096)
    imad_wide = ctx.bench_int_pipe(1, 4096)
    mulmod_rate = ctx.bench_int_pipe(2, 256)
    mulmod_ptx_rate = ctx.bench_int_pipe(3, 256)
    mulmod_eo_rate = ctx.bench_int_pipe(4, 256)
    int_pipe = {"imad_gops": imad_peak, "imad_wide_gops": imad_wide, "fr_mulmod_portable_gops": mulmod_rate,
                "fr_mulmod_ptx_gops": mulmod_ptx_rate, "fr_mulmod_evenodd_gops": mulmod_eo_rate,
                "fr_sqr_dedicated_gops": ctx.bench_int_pipe(8, 256), "fr_two_product_gops": ctx.bench_int_pipe(9, 256)}
    if args.workload == "ntt":
        ach = ntt_bytes(logn) / 1e9 / (ms_per_step * 1e-3)
        roof = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": None, "peak_source": peak_src, "kernel": dom_kernel,
                "note": "254-bit Montgomery butterflies make this kernel integer-pipe bound; see int_pipe"}
        int_pipe["kernel_mulmod_gops"] = (n // 2) * logn / 1e9 / (ms_per_step * 1e-3)
    elif args.workload in ("msm", "msm_sharded"):
        nl = n // world if args.workload == "msm_sharded" else n           # points per GPU
        ll = nl.bit_length() - 1
        c = int(os.environ.get("ZG_MSM_C", "0")) or (16 if ll >= 17 else max(8, ll - 2))
        ach = msm_imad(nl, c) / 1e9 / (ms_per_step * 1e-3)
        step_roof = {"bound": "int", "achieved": ach, "peak": imad_peak, "unit": "GIMAD/s", "frac": ach / imad_peak,
                     "window_c": c, "note": "SURVEY 8(d) per-window formula IMAD(N, c) / step time; it OVER-COUNTS for this fixed-base "
                                            "design (one bucket set instead of one per window) -- not an achieved rate, see `roofline`"}
        roof = kernel_roofline(probe, imad_peak, ms, peak_src="measured in this run (zg_bench_int_pipe kind 0)",
                               traffic=NCU_TRAFFIC.get("%s_%d" % (args.workload, logn)))
        extra["survey_formula_roofline"] = step_roof
    elif args.workload in ("keygen", "verify"):
        roof = None                                # a secondary number: no roofline claim (24 MSMs + 27 extended NTTs + uploads)
    else:
        roof = kernel_roofline(probe, imad_peak, lat_ms * args.steps, peak_src="measured in this run (zg_bench_int_pipe kind 0)",
                               traffic=NCU_TRAFFIC.get("proof_" + args.model))
        extra["stage_ms_last_proof"] = stage
    out = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u256 (8x32-bit Montgomery limbs)", "data": "synthetic",
        "config": dict({"workload": workload_name(args),
                        "l2": "flushed between steps" + ("; the e2e leg streams %.0f MB of freshly synthesized witness columns per step through "
                                                         "the GPU (larger than L2)" % (h2d / 1e6) if e2e_streams else "")},
                       **({"batch": "%d independent proofs in flight per GPU, %d proofs per step" % (extra["inflight"], extra["proofs_per_step"]),
                           "synthesis": "e2e: " + extra["e2e_starts_from"] + "; value: advice columns resident in HBM"}
                          if "inflight" in extra else {})),
        "roofline": roof, "int_pipe": int_pipe,
        "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    # the same dominant kernel against the HBM roofline (the base contract's vocabulary): algorithmic bytes = 72 B per point
    # addition (8-B sorted entry + 64-B table point); far from binding, which is the point
    if roof and roof.get("kernel") == "msm_accumulate_kernel":
        gbs = roof["point_additions_per_launch"] * 72 / 1e9 / (roof["avg_launch_ms"] * 1e-3)
        extra["roofline_hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                 "traffic": roof.get("traffic"), "peak_source": peak_src, "kernel": roof["kernel"],
                                 "note": "same launches as `roofline`; this kernel is bound by the integer pipe, not by HBM"}
    if "scaling_override" in extra:
        out["scaling"] = extra.pop("scaling_override")
    out.update(extra)
    # library tuning switches in force (unset = the built-in defaults: lazy two-product mixed addition and Horner steps)
    out["knobs"] = {k: v for k, v in sorted(os.environ.items()) if k.startswith("ZG_") and not k.startswith("ZG_BENCH_")}
    if cpu_fn is not None:
        import cpu_ref
        cores = cpu_ref.use_all_cores()
        cpu_fn()
        reps, t0 = 0, time.perf_counter()
        while reps < 2 or (time.perf_counter() - t0 < 10 and reps < 20):
            cpu_fn()
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        out["cpu_baseline"] = {"value": (cpu_units if args.workload in ("proof", "msm_sharded") else units) / dt, "unit": unit, "cores": cores, "kind": "port",
                               "sample": "same workload, %d repetitions on host cores (oracle restatement of the upstream "
                                         "CPU prover, OpenMP; witness synthesis excluded)" % reps}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
