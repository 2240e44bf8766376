#!/usr/bin/env python
"""Compare a dump of the UNMODIFIED Rust reference (rust/tools/ref_dump.rs) with the oracle -- and, with --gpu, with the
B200 prover -- on the same model, image, SRS secret and RNG seed.  Closes "parity unpinned" (DESIGN.md section 7):

    python scripts/compare_ref_dump.py ref_dump_tiny.json [--model tiny] [--gpu]

Reports, in order: public outputs, fixed / permutation commitments of the verifying key, transcript_repr (the dump's
value is then fed to the oracle and the GPU prover as their input), proof length, and the first differing proof byte with
the stage it belongs to (advice commitments, lookup commitments, ..., opening witnesses).  Exit status 0 = identical."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "0g-halo2_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

MODELS = {"tiny": ("model_28input_256entry_1hash_1bpi.hdf5", 14), "small": ("model_28input_1024entry_2hash_2bpi.hdf5", 15),
          "medium": ("model_28input_2048entry_2hash_3bpi.hdf5", 15)}
SECRET = 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8
SEED = bytes(range(16))


def le(hexstr):
    return int.from_bytes(bytes.fromhex(hexstr), "little")


def stage_of(offset, cs, nsets, qdeg):
    """which part of the proof a byte offset falls into (write order of create_proof)"""
    sizes = [("advice commitments", 64 * cs.num_advice), ("lookup permuted commitments", 128 * len(cs.lookups)),
             ("permutation product commitments", 64 * nsets), ("lookup product commitments", 64 * len(cs.lookups)),
             ("random polynomial commitment", 64), ("h piece commitments", 64 * qdeg),
             ("advice evaluations", 32 * len(cs.queries["advice"])), ("fixed evaluations", 32 * len(cs.queries["fixed"])),
             ("random polynomial evaluation", 32), ("sigma evaluations", 32 * len(cs.permutation)),
             ("permutation product evaluations", 32 * (3 * nsets - 1)), ("lookup evaluations", 160 * len(cs.lookups)),
             ("GWC opening witnesses", 1 << 30)]
    pos = 0
    for name, sz in sizes:
        if offset < pos + sz:
            return "%s (+%d)" % (name, offset - pos)
        pos += sz
    return "?"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("--model", default="tiny", choices=list(MODELS))
    ap.add_argument("--image", default=os.path.join(ROOT, "tests", "golden", "example_image_7.png"))
    ap.add_argument("--gpu", action="store_true", help="also run the B200 prover (needs a CUDA device)")
    args = ap.parse_args()
    import bn254
    import halo2_ref as H
    from zg_b200.io import load_grayscale_image, load_wnn
    d = json.load(open(args.dump))
    fname, k = MODELS[args.model]
    assert d["k"] == k, "dump is for k = %d, model %s uses k = %d" % (d["k"], args.model, k)
    wnn = load_wnn(os.path.join(ROOT, "tests", "golden", fname))
    img = load_grayscale_image(args.image)
    ok = True

    def check(name, cond, detail=""):
        nonlocal ok
        print("%-44s %s %s" % (name, "identical" if cond else "DIFFERENT", detail))
        ok = ok and cond
    outs = wnn.predict(img)
    check("public outputs", [le(x) for x in d["outputs"]] == outs)
    srs = H.Srs(k, SECRET)
    circ0, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    opk = H.keygen(srs, circ0.cs, asm0)
    ref_fixed = [(le(x), le(y)) for x, y in d["fixed_commitments"]]
    ref_perm = [(le(x), le(y)) for x, y in d["permutation_commitments"]]
    check("fixed commitments (%d)" % len(ref_fixed), ref_fixed == list(opk.fixed_commitments),
          "" if len(ref_fixed) == len(opk.fixed_commitments) else "(count %d vs %d)" % (len(ref_fixed), len(opk.fixed_commitments)))
    check("permutation commitments (%d)" % len(ref_perm), ref_perm == list(opk.perm_commitments))
    repr_ref = le(d["transcript_repr"])
    print("%-44s %s" % ("transcript_repr", "taken from the dump (Rust Debug-format hash; an input of the prover here)"))
    opk.transcript_repr = repr_ref
    _, asm = wnn.synthesize(img, k)
    oproof = H.create_proof(srs, opk, asm.advice, [outs], H.XorShiftRng(SEED))
    ref_proof = bytes.fromhex(d["proof"])
    cs = opk.cs
    chunk = cs.degree() - 2
    nsets = (len(cs.permutation) + chunk - 1) // chunk
    check("proof length", len(ref_proof) == len(oproof), "(%d vs %d)" % (len(ref_proof), len(oproof)))
    if ref_proof != oproof:
        first = next((i for i in range(min(len(ref_proof), len(oproof))) if ref_proof[i] != oproof[i]), None)
        check("proof bytes (oracle)", False, "first difference at offset %s: %s" % (first, first is not None and stage_of(first, cs, nsets, cs.degree() - 1)))
    else:
        check("proof bytes (oracle)", True)
    check("reference proof accepted by the oracle verifier", H.verify_proof(srs, opk, [outs], ref_proof))
    if args.gpu:
        import zg_b200
        from zg_b200.prover import ParamsKZG, create_proof, keygen
        ctx = zg_b200.Context(0)
        params = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
        circ1, asm1 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
        pk = keygen(ctx, params, circ1.cs, asm1, transcript_repr=repr_ref)
        gproof = create_proof(params, pk, asm.advice, [outs], zg_b200.lib.XorShift.from_seed(SEED))
        check("proof bytes (B200 prover)", gproof == ref_proof)
        check("reference proof accepted by the product verifier", pk.get_vk().verify(params, [outs], ref_proof))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
