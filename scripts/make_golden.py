#!/usr/bin/env python
"""Regenerates tests/golden/proofs.json: SHA-256 of the oracle's proof bytes (and the proof length, public outputs and
vk commitments digest) for the checked-in models with the fixed test SRS secret and XorShift seed.  The oracle is a CPU
restatement (parity unpinned against the Rust crates); this fixture pins the ORACLE ITSELF so that
neither it nor the GPU prover can drift silently: `-m "not gpu"` re-derives the tiny entry, `-m gpu` checks the GPU
proof of every entry against it."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "0g-halo2_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import halo2_ref as H  # noqa: E402
from zg_b200.io import load_grayscale_image, load_wnn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SECRET = 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8
SEED = bytes(range(16))
MODELS = [("model_28input_256entry_1hash_1bpi.hdf5", 14), ("model_28input_1024entry_2hash_2bpi.hdf5", 15),
          ("model_28input_2048entry_2hash_3bpi.hdf5", 15)]


def entry(fname, k):
    wnn = load_wnn(os.path.join(GOLD, fname))
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    srs = H.Srs(k, SECRET)
    circ0, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    pk = H.keygen(srs, circ0.cs, asm0)
    _, asm = wnn.synthesize(img, k)
    outputs = wnn.predict(img)
    proof = H.create_proof(srs, pk, asm.advice, [outputs], H.XorShiftRng(SEED))
    assert H.verify_proof(srs, pk, [outputs], proof)
    vk = hashlib.sha256(repr((pk.fixed_commitments, pk.perm_commitments)).encode()).hexdigest()
    return {"k": k, "outputs": [int(o) for o in outputs], "proof_len": len(proof),
            "proof_sha256": hashlib.sha256(proof).hexdigest(), "vk_commitments_sha256": vk,
            "transcript_repr": hex(pk.transcript_repr)}


if __name__ == "__main__":
    out = {"srs_secret": hex(SECRET), "rng": "XorShiftRng seed bytes 0..15", "image": "example_image_7.png",
           "models": {f: entry(f, k) for f, k in MODELS}}
    json.dump(out, open(os.path.join(GOLD, "proofs.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
