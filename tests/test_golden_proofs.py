"""Committed fixtures tests/golden/proofs.json (made by scripts/make_golden.py from the CPU oracle with a fixed SRS
secret and RNG seed).  CPU leg: the oracle still reproduces its own tiny-model entry (guards the checker against
silent drift).  GPU leg: zg_create_proof reproduces every entry's digest and zg_pk_load the vk commitments."""
import hashlib
import json
import os

import numpy as np
import pytest

import halo2_ref as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
FIX = json.load(open(os.path.join(GOLD, "proofs.json")))
SECRET = int(FIX["srs_secret"], 16)
SEED = bytes(range(16))


def _inputs(fname, k):
    from zg_b200.io import load_wnn, load_grayscale_image
    wnn = load_wnn(os.path.join(GOLD, fname))
    img = load_grayscale_image(os.path.join(GOLD, FIX["image"]))
    return wnn, img, H.Srs(k, SECRET)


def test_oracle_reproduces_tiny_fixture():
    fname = "model_28input_256entry_1hash_1bpi.hdf5"
    e = FIX["models"][fname]
    wnn, img, srs = _inputs(fname, e["k"])
    assert wnn.predict(img) == e["outputs"] == [9, 6, 13, 10, 17, 10, 9, 26, 11, 16]   # tests/integration_test.rs:13-20
    circ0, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), e["k"])
    pk = H.keygen(srs, circ0.cs, asm0)
    assert hashlib.sha256(repr((pk.fixed_commitments, pk.perm_commitments)).encode()).hexdigest() == e["vk_commitments_sha256"]
    _, asm = wnn.synthesize(img, e["k"])
    proof = H.create_proof(srs, pk, asm.advice, [e["outputs"]], H.XorShiftRng(SEED))
    assert len(proof) == e["proof_len"] and hashlib.sha256(proof).hexdigest() == e["proof_sha256"]
    assert H.verify_proof(srs, pk, [e["outputs"]], proof)                               # pairing check, no secret used
    assert H.verify_proof(srs, pk, [e["outputs"]], proof, use_trapdoor=True)           # the G1 shortcut agrees
    assert not H.verify_proof(srs, pk, [[o + 1 for o in e["outputs"]]], proof)          # wrong public outputs
    bad = bytearray(proof)
    bad[-1] ^= 1                                                                         # a bit of the last opening witness
    assert not H.verify_proof(srs, pk, [e["outputs"]], bytes(bad))


@pytest.mark.gpu
@pytest.mark.parametrize("fname", sorted(FIX["models"]))
def test_gpu_reproduces_fixture(ctx, fname):
    import zg_b200
    from zg_b200.prover import ParamsKZG, keygen, create_proof
    e = FIX["models"][fname]
    wnn, img, srs = _inputs(fname, e["k"])
    params = ParamsKZG(e["k"], srs.g, srs.g_lagrange)
    circ, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), e["k"])
    pk = keygen(ctx, params, circ.cs, asm0)
    assert hashlib.sha256(repr((pk.fixed_commitments, pk.perm_commitments)).encode()).hexdigest() == e["vk_commitments_sha256"]
    assert hex(pk.transcript_repr) == e["transcript_repr"]
    _, asm = wnn.synthesize(img, e["k"])
    proof = create_proof(params, pk, asm.advice, [wnn.predict(img)], zg_b200.lib.XorShift.from_seed(SEED))
    assert len(proof) == e["proof_len"] and hashlib.sha256(proof).hexdigest() == e["proof_sha256"]
    pk.close()


@pytest.mark.gpu
def test_wnn_proof_native_and_python_witness_agree(ctx):
    """Wnn::proof mirror (src/wnn.rs:232-262): image in, proof out, with the C++ witness synthesis and with the Python
    front-end -- both must hit the committed digest."""
    import zg_b200
    from zg_b200.prover import ParamsKZG
    fname = "model_28input_256entry_1hash_1bpi.hdf5"
    e = FIX["models"][fname]
    wnn, img, srs = _inputs(fname, e["k"])
    params = ParamsKZG(e["k"], srs.g, srs.g_lagrange)
    pk = wnn.generate_proving_key(ctx, params)
    for native in (True, False):
        proof, outputs = wnn.proof(pk, params, img, zg_b200.lib.XorShift.from_seed(SEED), native=native)
        assert outputs == e["outputs"]
        assert hashlib.sha256(proof).hexdigest() == e["proof_sha256"], "native=%s" % native
    pk.close()
