"""The C-ABI library loads and exports every symbol include/zg_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "zg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(zg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "zg_msm" in syms and "zg_ntt" in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol():
    import zg_b200
    assert os.path.exists(zg_b200.LIB_PATH), "libzg_b200.so not built (run __graft_entry__.build())"
    L = ctypes.CDLL(zg_b200.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, "missing exports: %s" % missing


def test_binding_covers_header():
    import zg_b200.lib as zl
    assert sorted(zl.EXPORTS) == declared_symbols()


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    import torch
    import zg_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zg_b200.ZgError):
        zg_b200.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "0g-halo2_b200")
    bad = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+(oracle|cpu_ref|bn254\b)", src, flags=re.M) or "zg_oracle" in src:
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_oracle_never_imports_product():
    """the checker stands on its own: nothing under oracle/ imports the package (its constraint-system bookkeeping and
    selector compression are oracle/circuit_ref.py, an independent restatement)"""
    bad = []
    for f in os.listdir(os.path.join(ROOT, "oracle")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "oracle", f)).read()
            if re.search(r"^\s*(from|import)\s+zg_b200", src, flags=re.M):
                bad.append(f)
    assert not bad, bad


def test_transcript_keccak256_known_answers():
    """the product's own keccak256 (host code in csrc/prover.cu, the hash of EvmTranscript): published vectors plus
    agreement with the oracle's implementation across the 136-byte rate boundary"""
    import ctypes
    import zg_b200
    import halo2_ref as H
    L = zg_b200.load_library()

    def k(data: bytes) -> bytes:
        out = (ctypes.c_uint8 * 32)()
        L.zg_debug_keccak256(data, len(data), out)
        return bytes(out)
    assert k(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert k(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    for n in (1, 31, 32, 33, 64, 96, 135, 136, 137, 271, 272, 273, 1000):
        data = bytes((7 * i + n) & 0xFF for i in range(n))
        assert k(data) == H.keccak256(data), n


def test_chacha20_rng_known_answer_and_stream():
    """zg_chacha20_* (the production RNG, OsRng's role at src/wnn.rs:256): the block function against RFC 8439 section 2.3.2
    (key 00..1f, block counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00 mapped onto the 64-bit counter / 64-bit nonce
    layout), stream continuity across calls of any size, and two OS-seeded generators never repeating each other."""
    import numpy as np
    import zg_b200.lib as zl
    r = zl.ChaCha20Rng.from_key(bytes(range(32)))
    r.counter = 1 | (0x09000000 << 32)
    r.nonce[0], r.nonce[1] = 0x4A000000, 0
    assert r.draw(8).tobytes().hex() == ("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
                                         "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e")
    a = zl.ChaCha20Rng.from_key(b"\x07" * 32)
    b = zl.ChaCha20Rng.from_key(b"\x07" * 32)
    whole = a.draw(1000)
    parts = np.concatenate([b.draw(k) for k in (1, 7, 8, 9, 120, 855)])
    assert (whole == parts).all() and len(set(whole.tolist())) == 1000
    x, y = zl.ChaCha20Rng.from_os().draw(16), zl.ChaCha20Rng.from_os().draw(16)
    assert (x != y).any()
