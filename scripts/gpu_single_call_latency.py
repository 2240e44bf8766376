#!/usr/bin/env python
"""Latency of ONE `Wnn::proof`-shaped call (image -> witness synthesis -> proof bytes, nothing pipelined): what one iteration of
the reference's bench_proof_generation (benches/bench.rs:30-36) times.  usage: python scripts/gpu_single_call_latency.py [model]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "0g-halo2_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import bench  # noqa: E402
import halo2_ref as H  # noqa: E402
import torch  # noqa: E402
from zg_b200.prover import ParamsKZG  # noqa: E402
from zg_b200.service import ProofService  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "large"
wnn, img, k = bench.load_model(model)
srs = H.Srs(k, bench.SRS_SECRET)
with ProofService(wnn, ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2), device=0, lanes=1) as service:
    for _ in range(3):
        service.prove(img)
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        t0 = time.perf_counter()
        proof, scores = service.prove(img)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    print(json.dumps({"model": model, "k": k, "single_call_ms_median": ts[len(ts) // 2], "min": ts[0], "max": ts[-1],
                      "what": "ProofService.prove(image): native witness synthesis + H2D + zg_create_proof + proof bytes, sequential"}))
