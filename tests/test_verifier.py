"""The product's host verifier (csrc/verifier.cu: zg_vk_create / zg_verify_proof / zg_pairing_check, the counterpart of
`Wnn::verify_proof`, /root/reference/src/wnn.rs:265-280) against the oracle: its own optimal-ate pairing vs the oracle's
(bilinearity, non-degeneracy, the KZG identity e([s]P, G2) = e(P, [s]G2)), and accept / reject of whole proofs produced
by the oracle prover, including every tampering the oracle verifier rejects.  CPU only: the verifier needs no GPU."""
import os

import numpy as np
import pytest

import bn254
import cpu_ref
import halo2_ref as H
from zg_b200.bn254_host import to_limbs
from zg_b200.io import load_grayscale_image, load_wnn
from zg_b200.plonk.mock import finalize_fixed
from zg_b200.plonk.serialize import serialize_cs
from zg_b200.prover import ParamsKZG, VerifyingKey, pairing_check

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SECRET = 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8


def g1(k):
    return bn254.g1_affine_to_limbs([bn254.g1_mul(bn254.G1_GEN, k % bn254.R_MOD)])[0]


def g2(k):
    return cpu_ref.g2_mul(cpu_ref.g2_generator(), bn254.fr_to_limbs([k % bn254.R_MOD])[0])


def test_pairing_bilinear_and_nondegenerate():
    a, b = 0x1234567890ABCDEF1234567, 0xFEDCBA9876543210FEDCBA98765
    # e(aP, bQ) * e(-abP, Q) == 1
    assert pairing_check([g1(a), g1(-a * b)], [g2(b), g2(1)])
    # e(P, Q) != 1 and e(aP, bQ) * e(-(ab+1)P, Q) != 1
    assert not pairing_check([g1(1)], [g2(1)])
    assert not pairing_check([g1(a), g1(-(a * b + 1))], [g2(b), g2(1)])
    # the identity on either side contributes 1
    zero1, zero2 = np.zeros(8, dtype=np.uint64), np.zeros(16, dtype=np.uint64)
    assert pairing_check([zero1, g1(5)], [g2(3), zero2])
    # agreement with the oracle's pairing on the same inputs
    for pts1, pts2 in [([g1(a), g1(-a * b)], [g2(b), g2(1)]), ([g1(7), g1(9)], [g2(11), g2(13)])]:
        assert pairing_check(pts1, pts2) == cpu_ref.pairing_check(np.stack(pts1), np.stack(pts2))


def test_pairing_rejects_points_off_the_curve():
    import zg_b200.lib as zl
    bad = g1(3).copy()
    bad[0] ^= np.uint64(1)
    with pytest.raises(zl.ZgError):
        pairing_check([bad], [g2(1)])
    bad2 = g2(3).copy()
    bad2[5] ^= np.uint64(2)
    with pytest.raises(zl.ZgError):
        pairing_check([g1(1)], [bad2])


@pytest.fixture(scope="module")
def tiny_case():
    wnn = load_wnn(os.path.join(GOLD, "model_28input_256entry_1hash_1bpi.hdf5"))
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    k = 14
    srs = H.Srs(k, SECRET)
    circ0, asm0 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    opk = H.keygen(srs, circ0.cs, asm0)
    _, asm = wnn.synthesize(img, k)
    out = wnn.predict(img)
    proof = H.create_proof(srs, opk, asm.advice, [out], H.XorShiftRng(bytes(range(16))))
    # the product's vk from the product's own front-end + the oracle's commitments (bit-equal to the GPU keygen's)
    circ1, asm1 = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    finalize_fixed(circ1.cs, asm1)
    words, constants = serialize_cs(circ1.cs)
    fc = bn254.g1_affine_to_limbs(opk.fixed_commitments)
    pc = bn254.g1_affine_to_limbs(opk.perm_commitments)
    vk = VerifyingKey(k, words, constants, fc, pc, opk.transcript_repr)
    params = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
    return wnn, srs, opk, vk, params, out, proof


def test_verifier_accepts_oracle_proof(tiny_case):
    wnn, srs, opk, vk, params, out, proof = tiny_case
    assert H.verify_proof(srs, opk, [out], proof)
    assert vk.verify(params, [out], proof)
    assert wnn.verify_proof(proof, vk, params, out)


def test_verifier_rejects_wrong_output_and_tampering(tiny_case):
    wnn, srs, opk, vk, params, out, proof = tiny_case
    wrong = list(out)
    wrong[0] += 1
    assert not vk.verify(params, [wrong], proof) and not H.verify_proof(srs, opk, [wrong], proof)
    rng = np.random.default_rng(5)
    # a flipped bit anywhere: commitments (off-curve or wrong point), evaluations, opening witnesses
    positions = [3, 40, 64 * 6 + 10, len(proof) // 2, len(proof) // 2 + 33, len(proof) - 70, len(proof) - 1] + \
        [int(x) for x in rng.integers(0, len(proof), 6)]
    for pos in positions:
        t = bytearray(proof)
        t[pos] ^= 1 << int(rng.integers(0, 8))
        assert not vk.verify(params, [out], bytes(t)), pos
    assert not vk.verify(params, [out], proof[:-1])
    assert not vk.verify(params, [out], proof + b"\0")
    assert not vk.verify(params, [out], b"")
    # a different verifying key (transcript_repr off by one) rejects the same proof
    vk2 = VerifyingKey(vk.k, vk.cs_words, vk.constants, vk.fixed_limbs, vk.perm_limbs, (vk.transcript_repr + 1) % bn254.R_MOD)
    assert not vk2.verify(params, [out], proof)
    # the wrong SRS ([s']G2) fails the pairing check only
    other = H.Srs(4, SECRET + 1)
    params2 = ParamsKZG(params.k, params.g, params.g_lagrange, params.g2, other.s_g2)
    assert not vk.verify(params2, [out], proof)


def test_vk_create_rejects_malformed_blob(tiny_case):
    import zg_b200.lib as zl
    _, _, _, vk, _, _, _ = tiny_case
    with pytest.raises(zl.ZgError):
        VerifyingKey(vk.k, vk.cs_words[:40], vk.constants, vk.fixed_limbs, vk.perm_limbs, 1)
    w = vk.cs_words.copy()
    w[0] ^= 1
    with pytest.raises(zl.ZgError):
        VerifyingKey(vk.k, w, vk.constants, vk.fixed_limbs, vk.perm_limbs, 1)


def test_vk_file_round_trip_verifies(tiny_case, tmp_path):
    """write_keys / read_vk (src/io.rs:159-176): the verifying key written in halo2's RawBytes layout, read back against a
    freshly configured circuit (selectors are re-compressed from the activations in the file) still verifies the proof."""
    import io as pyio
    from zg_b200 import io as zio
    from zg_b200.prover import load_verifying_key
    wnn, srs, opk, vk, params, out, proof = tiny_case
    k = vk.k
    circ, asm = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)       # uncompressed constraint system
    path = tmp_path / "vk.bin"
    with open(path, "wb") as f:
        zio.write_vk(f, k, vk.fixed_limbs, vk.perm_limbs, asm.selectors)
    raw = open(path, "rb").read()
    nsel, m, nf = len(circ.cs.selectors), len(circ.cs.permutation), vk.fixed_limbs.shape[0]
    assert raw[:4] == k.to_bytes(4, "big") and raw[4:8] == nf.to_bytes(4, "big")
    assert len(raw) == 8 + 64 * (nf + m) + nsel * ((1 << k) // 8)
    circ2, _ = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    vk2 = load_verifying_key(circ2.cs, open(path, "rb"), transcript_repr=vk.transcript_repr)
    assert (vk2.cs_words == vk.cs_words).all() and (vk2.constants == vk.constants).all()
    assert vk2.verify(params, [out], proof)
    # truncated / trailing bytes are rejected
    circ3, _ = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    with pytest.raises(ValueError):
        load_verifying_key(circ3.cs, pyio.BytesIO(raw[:-5]))
    circ4, _ = wnn.synthesize(np.zeros(wnn.img_shape(), dtype=np.uint8), k)
    with pytest.raises(ValueError):
        load_verifying_key(circ4.cs, pyio.BytesIO(raw + b"\0"))


def test_verifier_on_a_gadget_circuit():
    """not specific to the WNN circuit: the hash gadget's test circuit (src/gadgets/hash.rs:222-372) at k = 9, one instance
    value (42^3 mod 11 mod 8 = 3)"""
    from test_frontend_pinned import hash_circuit
    from zg_b200.plonk.circuit import Assembly, SimpleFloorPlanner
    k = 9
    srs = H.Srs(k, 0x1234567)

    def synth():
        cs, fn = hash_circuit(42)
        asm = Assembly(cs, k)
        fn(SimpleFloorPlanner(asm))
        return cs, asm
    cs, asm = synth()
    opk = H.keygen(srs, cs, asm)
    proof = H.create_proof(srs, opk, asm.advice, [[3]], H.XorShiftRng(bytes(range(16))))
    cs1, asm1 = synth()
    finalize_fixed(cs1, asm1)
    words, constants = serialize_cs(cs1)
    vk = VerifyingKey(k, words, constants, bn254.g1_affine_to_limbs(opk.fixed_commitments),
                      bn254.g1_affine_to_limbs(opk.perm_commitments), opk.transcript_repr)
    params = ParamsKZG(k, srs.g, srs.g_lagrange, srs.g2, srs.s_g2)
    assert vk.verify(params, [[3]], proof) and H.verify_proof(srs, opk, [[3]], proof)
    assert not vk.verify(params, [[2]], proof) and not H.verify_proof(srs, opk, [[2]], proof)
