"""scripts/compare_ref_dump.py is the tool that pins the oracle against a run of the real Rust crates
(rust/tools/ref_dump.rs).  No Rust toolchain exists here, so this test feeds it a dump in the same JSON format produced by
the oracle itself: the comparison must come out identical, and a corrupted dump must be reported with the stage of the
first differing byte.  (It proves the tool, not parity with Rust -- DESIGN.md section 7 still says "parity unpinned".)"""
import json
import os
import subprocess
import sys

import numpy as np

import halo2_ref as H
from zg_b200.io import load_grayscale_image, load_wnn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _le_hex(v):
    return int(v).to_bytes(32, "little").hex()


def test_compare_ref_dump_on_an_oracle_made_dump(tmp_path):
    wnn = load_wnn(os.path.join(GOLD, "model_28input_256entry_1hash_1bpi.hdf5"))
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    k = 14
    srs = H.Srs(k, 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8)
    circ0, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    pk = H.keygen(srs, circ0.cs, asm0)
    pk.transcript_repr = 0x1234567890ABCDEF                      # stands for the Rust-side value: an input on both sides
    _, asm = wnn.synthesize(img, k)
    outs = wnn.predict(img)
    proof = H.create_proof(srs, pk, asm.advice, [outs], H.XorShiftRng(bytes(range(16))))
    dump = {"k": k, "transcript_repr": _le_hex(pk.transcript_repr),
            "fixed_commitments": [[_le_hex(x), _le_hex(y)] for x, y in pk.fixed_commitments],
            "permutation_commitments": [[_le_hex(x), _le_hex(y)] for x, y in pk.perm_commitments],
            "outputs": [_le_hex(o) for o in outs], "proof": proof.hex()}
    good = tmp_path / "dump.json"
    good.write_text(json.dumps(dump))
    tool = [sys.executable, os.path.join(ROOT, "scripts", "compare_ref_dump.py")]
    r = subprocess.run(tool + [str(good)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DIFFERENT" not in r.stdout and "proof bytes (oracle)" in r.stdout
    # corrupt one byte inside the lookup product commitments: the tool names the stage and fails
    nadv, nlk, nsets = 6, 4, 2
    off = 64 * nadv + 128 * nlk + 64 * nsets + 7
    bad = bytearray(proof)
    bad[off] ^= 1
    dump["proof"] = bytes(bad).hex()
    badf = tmp_path / "bad.json"
    badf.write_text(json.dumps(dump))
    r = subprocess.run(tool + [str(badf)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1
    assert "lookup product commitments (+7)" in r.stdout, r.stdout
