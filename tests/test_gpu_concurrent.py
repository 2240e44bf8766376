"""Several contexts on one GPU driven from host threads (what bench.py --inflight and the proof farm do): every lane
must produce the oracle's proof bytes while the others run, and the live kernel probe must account for every MSM."""
import os
import threading

import numpy as np
import pytest

import halo2_ref as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_three_lanes_prove_concurrently(ctx):
    import torch
    import zg_b200
    from zg_b200.io import load_wnn, load_grayscale_image
    from zg_b200.prover import ParamsKZG, keygen, create_proof
    wnn = load_wnn(os.path.join(GOLD, "model_28input_256entry_1hash_1bpi.hdf5"))
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    k = 14
    srs = H.Srs(k, 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8)
    outputs = wnn.predict(img)
    zero = np.zeros((28, 28), dtype=np.uint8)
    circ0, asm0 = wnn.synthesize(zero, k)
    opk = H.keygen(srs, circ0.cs, asm0)
    _, asm = wnn.synthesize(img, k)
    seeds = [bytes([s] * 16) for s in (1, 2, 3)]
    expected = [H.create_proof(srs, opk, asm.advice, [outputs], H.XorShiftRng(s)) for s in seeds]
    lanes = []
    for i in range(3):
        c = ctx if i == 0 else zg_b200.Context(0, torch.cuda.Stream().cuda_stream)
        params = ParamsKZG(k, srs.g, srs.g_lagrange)
        ci, ai = wnn.synthesize(zero, k)
        lanes.append((c, params, keygen(c, params, ci.cs, ai)))
    got = [[None, None] for _ in lanes]

    def work(i):
        c, params, pk = lanes[i]
        for rep in range(2):                       # two rounds: workspaces are reused while the others are mid-proof
            got[i][rep] = create_proof(params, pk, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(seeds[i]))
    ctx.probe_enable(True)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    kms, launches, adds = ctx.probe_read()
    ctx.probe_enable(False)
    for i in range(3):
        assert got[i][0] == expected[i] and got[i][1] == expected[i], "lane %d differs from the oracle" % i
    # lane 0 ran two proofs = 2 x 5 MSM batches (advice, permuted, products + random, h pieces, W) through the probe
    assert launches == 2 * 5 and adds > 0 and kms > 0
    for c, _, pk in lanes:
        pk.close()
    for c, _, _ in lanes[1:]:
        c.close()


def test_proof_service_image_to_proof():
    """ProofService (zg_b200/service.py): lanes of independent provers, image in -> proof out with the native witness
    synthesis; job 0 must reproduce the committed digest, every other proof must verify."""
    import hashlib
    import json
    import zg_b200
    from zg_b200.io import load_wnn, load_grayscale_image, synthetic_image
    from zg_b200.prover import ParamsKZG
    from zg_b200.service import ProofService
    fix = json.load(open(os.path.join(GOLD, "proofs.json")))
    fname = "model_28input_256entry_1hash_1bpi.hdf5"
    e = fix["models"][fname]
    wnn = load_wnn(os.path.join(GOLD, fname))
    srs = H.Srs(e["k"], int(fix["srs_secret"], 16))
    images = [load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))] + [synthetic_image(i) for i in range(5)]
    seed = bytes(range(16))
    with ProofService(wnn, ParamsKZG(e["k"], srs.g, srs.g_lagrange), device=0, lanes=3,
                      rng_factory=lambda job: zg_b200.lib.XorShift.from_seed(seed)) as svc:
        results = svc.prove_many(images)
        assert hashlib.sha256(results[0][0]).hexdigest() == e["proof_sha256"] and results[0][1] == e["outputs"]
        zero = np.zeros((28, 28), dtype=np.uint8)
        circ0, asm0 = wnn.synthesize(zero, e["k"])
        opk = H.keygen(srs, circ0.cs, asm0)
        assert svc.vk.fixed_commitments == opk.fixed_commitments
        for img, (proof, outputs) in zip(images, results):
            assert outputs == wnn.predict(img)
            assert H.verify_proof(srs, opk, [outputs], proof)
        single = svc.prove(images[0])
        assert single[0] == results[0][0]
