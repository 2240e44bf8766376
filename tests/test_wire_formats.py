"""Round trips of the files exchanged with the Rust CLI (zg_b200/io.py mirrors /root/reference/src/io.rs:137-207)."""
import json

import numpy as np
import pytest

import bn254

from zg_b200 import io as zio
from zg_b200.bn254_host import R_MOD, to_limbs


def test_circuit_params_json(tmp_path):
    p = {"p": (1 << 21) - 9, "l": 20, "n_hashes": 2, "bits_per_hash": 10, "bits_per_filter": 28, "n_classes": 10}
    path = str(tmp_path / "circuit_params.json")
    zio.write_circuit_params(p, path)
    assert zio.read_circuit_params(path) == p
    # serde_json keeps the struct's field order (src/gadgets/wnn.rs:245-253)
    assert list(json.load(open(path)).keys()) == ["p", "l", "n_hashes", "bits_per_hash", "bits_per_filter", "n_classes"]


def test_proof_with_output_json(tmp_path):
    proof = bytes(range(256)) * 15
    outputs = [17, 13, 25, 27, 29, 21, 15, 55, 27, 32]            # tests/integration_test.rs:30-37
    path = str(tmp_path / "proof.json")
    zio.write_proof_with_output(proof, outputs, path)
    d = json.load(open(path))
    assert list(d.keys()) == ["proof", "output"] and d["proof"][:4] == [0, 1, 2, 3]
    # an Fr is the array of its 4 Montgomery limbs: 17 * R mod r
    assert d["output"][0] == [int(x) for x in to_limbs([17])[0]]
    p2, o2 = zio.read_proof_with_output(path)
    assert p2 == proof and o2 == outputs


def test_srs_raw_bytes_round_trip(tmp_path):
    k = 4
    rng = np.random.default_rng(1)
    g = rng.integers(0, 1 << 62, size=(16, 8), dtype=np.uint64)
    gl = rng.integers(0, 1 << 62, size=(16, 8), dtype=np.uint64)
    g2, sg2 = bytes(range(128)), bytes(reversed(range(128)))
    path = str(tmp_path / "kzg.srs")
    zio.write_srs(k, g, gl, g2, sg2, path)
    assert (tmp_path / "kzg.srs").stat().st_size == 4 + 2 * 16 * 64 + 256
    k2, g_, gl_, g2_, sg2_ = zio.read_srs(path)
    assert k2 == k and (g_ == g).all() and (gl_ == gl).all() and g2_ == g2 and sg2_ == sg2


def test_encode_calldata_layout():
    outputs = [9, 6, 13, 10, 17, 10, 9, 26, 11, 16]               # tests/integration_test.rs:13-20
    proof = bytes(range(200))
    cd = zio.encode_calldata([outputs], proof)
    assert len(cd) == 32 * len(outputs) + len(proof)
    assert cd[:32] == (9).to_bytes(32, "big") and cd[32 * 9:32 * 10] == (16).to_bytes(32, "big")
    assert cd[32 * len(outputs):] == proof
    assert zio.encode_calldata([[R_MOD + 5]], b"")[-1] == 5        # reduced like Fr


def test_pk_file_layout_round_trip(tmp_path):
    """ProvingKey::write / read in halo2's RawBytes layout (src/io.rs:159-170) on a small gadget circuit: every section comes
    back bit for bit, the big-endian length prefixes sit where the format puts them, inconsistent files are refused."""
    import io as pyio
    import halo2_ref as H
    from test_frontend_pinned import hash_circuit
    from zg_b200 import io as zio
    from zg_b200.plonk.circuit import Assembly, SimpleFloorPlanner
    k = 9
    cs, synth = hash_circuit(42)
    asm = Assembly(cs, k)
    synth(SimpleFloorPlanner(asm))
    srs = H.Srs(k, 0x1234567)
    opk = H.keygen(srs, cs, asm)
    n, ext_n = 1 << k, opk.domain.ext_n
    vk = {"k": k, "fixed_commitments": bn254.g1_affine_to_limbs(opk.fixed_commitments),
          "perm_commitments": bn254.g1_affine_to_limbs(opk.perm_commitments), "selectors": asm.selectors}
    buf = pyio.BytesIO()
    zio.write_pk(buf, vk, opk.l0, opk.l_last, opk.l_active, opk.fixed_values, opk.fixed_polys, opk.fixed_cosets,
                 opk.perm_values, opk.perm_polys, opk.perm_cosets)
    raw = buf.getvalue()
    nf, m, nsel = len(opk.fixed_values), len(opk.perm_values), len(asm.selectors)
    vk_len = 8 + 64 * (nf + m) + nsel * (n // 8)
    assert raw[vk_len:vk_len + 4] == ext_n.to_bytes(4, "big")                       # l0: Polynomial = len (u32 BE) | values
    expect = vk_len + 3 * (4 + 32 * ext_n) + 2 * (4 + nf * (4 + 32 * n)) + (4 + nf * (4 + 32 * ext_n)) + \
        2 * (4 + m * (4 + 32 * n)) + (4 + m * (4 + 32 * ext_n))
    assert len(raw) == expect
    d = zio.read_pk(pyio.BytesIO(raw), m, nsel)
    assert d["k"] == k and d["selectors"] == [list(map(bool, s)) for s in asm.selectors]
    assert (d["fixed_commitments"] == vk["fixed_commitments"]).all() and (d["perm_commitments"] == vk["perm_commitments"]).all()
    for name, ref in (("l0", [opk.l0]), ("l_last", [opk.l_last]), ("l_active_row", [opk.l_active])):
        assert (d[name] == ref[0]).all()
    for name, ref in (("fixed_values", opk.fixed_values), ("fixed_polys", opk.fixed_polys), ("fixed_cosets", opk.fixed_cosets),
                      ("perm_values", opk.perm_values), ("perm_polys", opk.perm_polys), ("perm_cosets", opk.perm_cosets)):
        assert len(d[name]) == len(ref) and all((a == b).all() for a, b in zip(d[name], ref)), name
    with pytest.raises(ValueError):
        zio.read_pk(pyio.BytesIO(raw[:-1]), m, nsel)
    with pytest.raises(ValueError):
        zio.read_pk(pyio.BytesIO(raw + b"\0"), m, nsel)
    with pytest.raises(ValueError):
        zio.read_pk(pyio.BytesIO(raw), m + 1, nsel)
