"""End-to-end parity of the B200 prover (zg_pk_load + zg_create_proof through the C ABI) against the CPU
oracle's restatement of halo2's keygen / create_proof (oracle/halo2_ref.py) for the WNN circuit:
same SRS, same seeded RNG => identical vk commitments and identical proof BYTES; the restated
verifier accepts the GPU proof and rejects tampering.  Models/images are the reference's own files."""
import os

import numpy as np
import pytest

import bn254
import halo2_ref as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SEED = bytes(range(16))


def _stage_of(offset, cs):
    """which proof section a byte offset falls in (for mismatch reports)."""
    sections = [("advice commitments", 64 * cs.num_advice), ("lookup permuted commitments", 128 * len(cs.lookups)),
                ("permutation z commitments", 64 * 2), ("lookup z commitments", 64 * len(cs.lookups)),
                ("random poly commitment", 64), ("h piece commitments", 64 * (cs.degree() - 1)),
                ("advice evals", 32 * len(cs.queries["advice"])), ("fixed evals", 32 * len(cs.queries["fixed"])),
                ("random eval", 32), ("sigma evals", 32 * len(cs.permutation)), ("perm evals", 32 * 5),
                ("lookup evals", 32 * 5 * len(cs.lookups)), ("gwc witnesses", 64 * 4)]
    pos = 0
    for name, ln in sections:
        if offset < pos + ln:
            return "%s (+%d)" % (name, offset - pos)
        pos += ln
    return "past end"


@pytest.fixture(scope="module")
def tiny():
    from zg_b200.io import load_wnn, load_grayscale_image
    wnn = load_wnn(os.path.join(GOLD, "model_28input_256entry_1hash_1bpi.hdf5"))
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    k = 14
    srs = H.Srs(k, 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8)
    return wnn, img, k, srs


def test_tiny_model_proof_bytes_match_oracle(ctx, tiny):
    import zg_b200
    from zg_b200.prover import ParamsKZG, keygen, create_proof
    wnn, img, k, srs = tiny
    outputs = wnn.predict(img)
    assert outputs == [9, 6, 13, 10, 17, 10, 9, 26, 11, 16]     # tests/integration_test.rs:13-20
    # oracle side
    circ0, asm0 = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    opk = H.keygen(srs, circ0.cs, asm0)
    _, asm = wnn.synthesize(img, k)
    oproof = H.create_proof(srs, opk, asm.advice, [outputs], H.XorShiftRng(SEED))
    assert H.verify_proof(srs, opk, [outputs], oproof)
    # B200 side (fresh constraint system: keygen compresses selectors in place)
    params = ParamsKZG(k, srs.g, srs.g_lagrange)
    circ1, asm1 = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    pk = keygen(ctx, params, circ1.cs, asm1)
    assert pk.fixed_commitments == opk.fixed_commitments
    assert pk.perm_commitments == opk.perm_commitments
    assert pk.transcript_repr == opk.transcript_repr
    proof = create_proof(params, pk, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(SEED))
    assert len(proof) == len(oproof) == 3840
    if proof != oproof:
        first = next(i for i in range(len(proof)) if proof[i] != oproof[i])
        pytest.fail("proof bytes differ from the oracle first at offset %d: %s" % (first, _stage_of(first, circ0.cs)))
    assert H.verify_proof(srs, opk, [outputs], proof)
    # the product's own verifier (Wnn::verify_proof, src/wnn.rs:265-280) on the product's own key: accept, reject a wrong
    # output, reject a flipped byte -- and a proof drawn from the production RNG (OS-seeded ChaCha20) verifies as well
    vparams = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
    assert wnn.verify_proof(proof, pk, vparams, outputs)
    assert not wnn.verify_proof(proof, pk, vparams, [outputs[0] + 1] + outputs[1:])
    bad = bytearray(proof)
    bad[1000] ^= 4
    assert not wnn.verify_proof(bytes(bad), pk, vparams, outputs)
    proof_os, out_os = wnn.proof(pk, params, img)
    assert out_os == outputs and proof_os != proof and wnn.verify_proof(proof_os, pk, vparams, outputs)
    assert H.verify_proof(srs, opk, [outputs], proof_os)
    bad = bytearray(proof)
    bad[1000] ^= 1
    assert not H.verify_proof(srs, opk, [outputs], bytes(bad))
    # a second proof with another seed differs but verifies; same seed reproduces the bytes
    p2 = create_proof(params, pk, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(bytes(range(1, 17))))
    assert p2 != proof and H.verify_proof(srs, opk, [outputs], p2)
    p3 = create_proof(params, pk, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(SEED))
    assert p3 == proof
    print("stage ms", pk.stage_ms())
    pk.close()


def test_lookup_failure_is_reported(ctx, tiny):
    """a witness whose lookup input is not in the table must fail like plonk::Error::ConstraintSystemFailure"""
    import zg_b200
    from zg_b200.prover import ParamsKZG, keygen, create_proof
    wnn, img, k, srs = tiny
    params = ParamsKZG(k, srs.g, srs.g_lagrange)
    circ, asm0 = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    pk = keygen(ctx, params, circ.cs, asm0)
    _, asm = wnn.synthesize(img, k)
    adv = [list(c) for c in asm.advice]
    # break a byte decomposition: a running-sum cell under the range-check lookup becomes huge
    sel_rows = [r for r in range(asm.usable_rows) if asm.selectors[8][r]]
    adv[5][sel_rows[0]] = (adv[5][sel_rows[0]] + (1 << 40)) % bn254.R_MOD
    with pytest.raises(zg_b200.ZgError) as ei:
        create_proof(params, pk, adv, [wnn.predict(img)], zg_b200.lib.XorShift.from_seed(SEED))
    assert ei.value.code == -5
    pk.close()


def _parity(ctx, wnn, img, k, srs):
    import zg_b200
    from zg_b200.prover import ParamsKZG, keygen, create_proof
    outputs = wnn.predict(img)
    zero = np.zeros(wnn.img_shape(), dtype=np.uint8)
    circ0, asm0 = wnn.synthesize(zero, k)
    opk = H.keygen(srs, circ0.cs, asm0)
    _, asm = wnn.synthesize(img, k)
    oproof = H.create_proof(srs, opk, asm.advice, [outputs], H.XorShiftRng(SEED))
    assert H.verify_proof(srs, opk, [outputs], oproof)
    params = ParamsKZG(k, srs.g, srs.g_lagrange)
    circ1, asm1 = wnn.synthesize(zero, k)
    pk = keygen(ctx, params, circ1.cs, asm1)
    assert pk.fixed_commitments == opk.fixed_commitments and pk.perm_commitments == opk.perm_commitments
    proof = create_proof(params, pk, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(SEED))
    if proof != oproof:
        first = next(i for i in range(min(len(proof), len(oproof))) if proof[i] != oproof[i])
        pytest.fail("proof bytes differ from the oracle first at offset %d: %s" % (first, _stage_of(first, circ0.cs)))
    assert H.verify_proof(srs, opk, [outputs], proof)
    ms = pk.stage_ms()
    pk.close()
    return outputs, ms


@pytest.mark.parametrize("name,expected", [
    ("model_28input_1024entry_2hash_2bpi.hdf5", [17, 13, 25, 27, 29, 21, 15, 55, 27, 32]),     # integration_test.rs:30-37
    ("model_28input_2048entry_2hash_3bpi.hdf5", [29, 21, 40, 47, 45, 41, 28, 82, 35, 66]),     # integration_test.rs:47-54
])
def test_small_and_medium_models_match_oracle(ctx, name, expected):
    """BASELINE configs[1] and [2]: k = 15, seeded RNG, proof bytes identical to the CPU restatement."""
    from zg_b200.io import load_wnn, load_grayscale_image
    wnn = load_wnn(os.path.join(GOLD, name))
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    srs = H.Srs(15, 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8)
    outputs, ms = _parity(ctx, wnn, img, 15, srs)
    assert outputs == expected
    print(name, "stage ms", ms)


def test_large_shape_model_matches_oracle(ctx):
    """BASELINE configs[3] shape (49 inputs/filter, 8192 entries, 4 hashes, 6 bits/input, k = 17) on the
    synthetic stand-in for the model file that is absent from the reference checkout, with a synthetic image."""
    from zg_b200.io import synthetic_wnn, synthetic_image
    wnn = synthetic_wnn()
    img = synthetic_image(0)
    srs = H.Srs(17, 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8)
    outputs, ms = _parity(ctx, wnn, img, 17, srs)
    print("large stage ms", ms)


def test_key_files_round_trip_through_the_gpu(ctx, tiny, tmp_path):
    """write_keys / read_pk / read_vk (src/io.rs:159-176): a key exported in halo2's RawBytes layout and loaded again (fixed
    and sigma columns from the file, everything else recomputed on the device) proves the same bytes; its extended
    sections are halo2's zeta-coset forms and equal the oracle's; the vk file drives the host verifier."""
    import zg_b200
    from zg_b200 import io as zio
    from zg_b200.prover import (ParamsKZG, create_proof, export_proving_key, export_verifying_key, keygen, load_proving_key,
                                load_verifying_key)
    wnn, img, k, srs = tiny
    outputs = wnn.predict(img)
    params = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
    circ1, asm1 = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    selectors = [list(s) for s in asm1.selectors]
    pk = keygen(ctx, params, circ1.cs, asm1)
    _, asm = wnn.synthesize(img, k)
    proof = create_proof(params, pk, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(SEED))
    pk_path, vk_path = tmp_path / "pk.bin", tmp_path / "vk.bin"
    with open(pk_path, "wb") as f:
        export_proving_key(pk, selectors, f)
    with open(vk_path, "wb") as f:
        export_verifying_key(pk, selectors, f)
    # the file's sections against the oracle's key (same SRS): halo2's extended forms, not the backend's internal domain
    circ0, asm0 = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    opk = H.keygen(srs, circ0.cs, asm0)
    d = zio.read_pk(open(pk_path, "rb"), len(asm0.perm_cols), len(selectors))
    assert d["l0"].shape[0] == opk.domain.ext_n
    for name, ref in (("fixed_values", opk.fixed_values), ("fixed_polys", opk.fixed_polys), ("fixed_cosets", opk.fixed_cosets),
                      ("perm_values", opk.perm_values), ("perm_polys", opk.perm_polys), ("perm_cosets", opk.perm_cosets)):
        assert all((a == b).all() for a, b in zip(d[name], ref)), name
    assert (d["l0"] == opk.l0).all() and (d["l_last"] == opk.l_last).all() and (d["l_active_row"] == opk.l_active).all()
    # load it back
    circ2, _ = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    pk2 = load_proving_key(ctx, params, circ2.cs, open(pk_path, "rb"))
    assert pk2.fixed_commitments == pk.fixed_commitments and pk2.perm_commitments == pk.perm_commitments
    assert pk2.transcript_repr == pk.transcript_repr
    proof2 = create_proof(params, pk2, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(SEED))
    assert proof2 == proof
    circ3, _ = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    vk = load_verifying_key(circ3.cs, open(vk_path, "rb"))
    assert vk.verify(params, [outputs], proof2)
    pk2.close()
    pk.close()


def test_cloned_key_and_shared_srs(ctx, tiny):
    """zg_srs_share / zg_pk_clone: a second context of the same GPU proves with the first one's window tables and the first
    key's resident columns -- the same bytes as the original -- and keeps working after the original key and context are gone
    (the shared halves are reference-counted)."""
    import zg_b200
    from zg_b200.prover import ParamsKZG, create_proof, keygen
    wnn, img, k, srs = tiny
    outputs = wnn.predict(img)
    _, asm = wnn.synthesize(img, k)
    first = zg_b200.Context(0)
    params = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
    circ, asm0 = wnn.synthesize(np.zeros((28, 28), dtype=np.uint8), k)
    pk = keygen(first, params, circ.cs, asm0)
    proof = create_proof(params, pk, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(SEED))
    second = zg_b200.Context(0)
    with pytest.raises(zg_b200.ZgError):
        pk.clone(second)                                   # no SRS on that context yet
    params2 = ParamsKZG(k, srs.g, srs.g_lagrange)
    params2.share(second, first)
    pk2 = pk.clone(second)
    assert create_proof(params2, pk2, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(SEED)) == proof
    pk.close()
    first.close()
    again = create_proof(params2, pk2, asm.advice, [outputs], zg_b200.lib.XorShift.from_seed(SEED))
    assert again == proof and pk2.get_vk().verify(params, [outputs], again)
    pk2.close()
    second.close()


def test_small_gadget_circuit_proof_matches_oracle(ctx):
    """The prover is not specific to the WNN circuit: the hash gadget's test circuit (src/gadgets/hash.rs:222-372) at k = 9 --
    one lookup, degree-5 gates, 2^10-point extended blocks, the smallest MSM windows -- gives the oracle's bytes and verifies."""
    import zg_b200
    from test_frontend_pinned import hash_circuit
    from zg_b200.plonk.circuit import Assembly, SimpleFloorPlanner
    from zg_b200.prover import ParamsKZG, create_proof, keygen
    k = 9
    srs = H.Srs(k, 0x1234567)

    def synth(x):
        cs, fn = hash_circuit(x)
        asm = Assembly(cs, k)
        fn(SimpleFloorPlanner(asm))
        return cs, asm
    cs0, asm0 = synth(42)
    opk = H.keygen(srs, cs0, asm0)
    oproof = H.create_proof(srs, opk, asm0.advice, [[3]], H.XorShiftRng(SEED))
    assert H.verify_proof(srs, opk, [[3]], oproof)
    cs1, asm1 = synth(42)
    params = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
    pk = keygen(ctx, params, cs1, asm1)
    assert pk.fixed_commitments == opk.fixed_commitments and pk.perm_commitments == opk.perm_commitments
    proof = create_proof(params, pk, asm1.advice, [[3]], zg_b200.lib.XorShift.from_seed(SEED))
    assert proof == oproof
    assert pk.get_vk().verify(params, [[3]], proof) and not pk.get_vk().verify(params, [[2]], proof)
    pk.close()
