"""CPU-side checks of the device code's logic:
  * the host paths of csrc/field.cuh + curve.cuh (4 x 64-bit-limb multiplier / adder / subtractor, almost-inverse, and the
    8 x 32-bit-limb body that the device's variant 0 uses) against Python big integers;
  * the Python models of the NTT pass planner and the MSM segmented-reduction pipeline
    (tests/models/*) against the naive DFT / scalar MSM.
No GPU needed; the kernels themselves are exercised by the -m gpu tests."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import bn254
from bn254 import R_MOD, Q_MOD

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(HERE, "host", "host_harness.cpp")
    so = os.path.join(HERE, "host", "libhost_harness.so")
    hdr = os.path.join(HERE, "..", "0g-halo2_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(hdr, f)) for f in ("field.cuh", "curve.cuh"))
    if not os.path.exists(so) or os.path.getmtime(so) < max(newest, os.path.getmtime(src)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    return ctypes.CDLL(so)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("name,mod", [("fr", R_MOD), ("fq", Q_MOD)])
def test_portable_field_ops(harness, name, mod):
    rnd = random.Random(1)
    n = 1000
    a = [rnd.randrange(mod) for _ in range(n)]
    b = [rnd.randrange(mod) for _ in range(n)]
    a[:6] = [0, 1, mod - 1, mod - 1, 0, 1]
    b[:6] = [0, mod - 1, mod - 1, 1, 5, 1]
    A, B = bn254.ints_to_limbs(a, mod), bn254.ints_to_limbs(b, mod)
    O = np.zeros_like(A)
    for op, f in (("mul", lambda x, y: x * y % mod), ("add", lambda x, y: (x + y) % mod), ("sub", lambda x, y: (x - y) % mod)):
        getattr(harness, "h_%s_%s" % (name, op))(P(A), P(B), P(O), n)
        assert bn254.limbs_to_ints(O, mod) == [f(x, y) for x, y in zip(a, b)], (name, op)


@pytest.mark.parametrize("name,mod", [("fr", R_MOD), ("fq", Q_MOD)])
def test_host_64bit_bodies_edge_operands(harness, name, mod):
    """the host's 4 x 64-bit-limb multiplier / adder / subtractor and the 8 x 32-bit-limb body on operands that stress the
    carries and the final subtraction; Kaliski's almost-inverse on both fields (0 -> 0)"""
    rnd = random.Random(9)
    edge = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, (1 << 64) - 1, 1 << 64, (1 << 128) - 1, 1 << 128,
            (1 << 192) - 1, 1 << 192, (1 << 253) - 1, 1 << 253, mod - (1 << 64), mod - (1 << 128), int("f" * 63, 16) % mod]
    a = [x for x in edge for _ in edge] + [rnd.randrange(mod) for _ in range(500)]
    b = [y for _ in edge for y in edge] + [rnd.randrange(mod) for _ in range(500)]
    A, B = bn254.ints_to_limbs(a, mod), bn254.ints_to_limbs(b, mod)
    O = np.zeros_like(A)
    for op, f in (("mul", lambda x, y: x * y % mod), ("mul_portable", lambda x, y: x * y % mod),
                  ("add", lambda x, y: (x + y) % mod), ("sub", lambda x, y: (x - y) % mod)):
        getattr(harness, "h_%s_%s" % (name, op))(P(A), P(B), P(O), len(a))
        assert bn254.limbs_to_ints(O, mod) == [f(x, y) for x, y in zip(a, b)], (name, op)
    inv_in = edge + [rnd.randrange(mod) for _ in range(300)]
    I = bn254.ints_to_limbs(inv_in, mod)
    OI = np.zeros_like(I)
    getattr(harness, "h_%s_inv" % name)(P(I), P(OI), len(inv_in))
    assert bn254.limbs_to_ints(OI, mod) == [pow(x, -1, mod) if x % mod else 0 for x in inv_in]


def test_portable_inverse_and_mont(harness):
    rnd = random.Random(2)
    a = [0] + [rnd.randrange(R_MOD) for _ in range(30)]
    A = bn254.fr_to_limbs(a)
    O = np.zeros_like(A)
    harness.h_fr_inv(P(A), P(O), len(a))
    assert bn254.fr_from_limbs(O) == [pow(x, -1, R_MOD) if x else 0 for x in a]
    raw = bn254.ints_to_limbs(a, R_MOD, mont=False)
    harness.h_fr_to_mont(P(raw), P(O), len(a))
    assert (O == A).all()
    harness.h_fr_from_mont(P(A), P(O), len(a))
    assert (O == raw).all()


def test_xyzz_group_law(harness):
    rnd = random.Random(3)
    pts = [bn254.g1_mul(bn254.G1_GEN, rnd.randrange(R_MOD)) for _ in range(20)]
    pts[3] = None
    pts[5] = pts[4]                  # forces the doubling branch of the mixed add
    pts[8] = bn254.g1_neg(pts[7])    # forces the cancellation branch
    arr = bn254.g1_affine_to_limbs(pts)
    out = np.zeros((1, 12), dtype=np.uint64)
    harness.h_madd_chain(P(arr), 20, P(out))
    exp = None
    for p in pts:
        exp = bn254.g1_add(exp, p)
    assert bn254.g1_proj_from_limbs(out)[0] == exp
    harness.h_add_two_chains(P(arr), 10, P(arr[10:].copy()), 10, P(out))
    assert bn254.g1_proj_from_limbs(out)[0] == exp
    harness.h_add_two_chains(P(arr), 10, P(arr), 10, P(out))   # full add of equal points
    e2 = None
    for p in pts[:10]:
        e2 = bn254.g1_add(e2, p)
    assert bn254.g1_proj_from_limbs(out)[0] == bn254.g1_add(e2, e2)


def test_ntt_pass_plan_model():
    from ntt_model import ntt_run
    rnd = random.Random(4)
    for log_n in range(1, 9):
        for max_s, logc in ((3, 1), (2, 1), (8, 2), (3, 2), (4, 2)):
            n = 1 << log_n
            a = [rnd.randrange(R_MOD) for _ in range(n)]
            w = bn254.omega(log_n)
            assert ntt_run(a, w, log_n, max_s, logc) == bn254.dft_naive(a, w), (log_n, max_s, logc)


def test_ntt_coset_padding_truncation_model():
    from ntt_model import ntt_run
    rnd = random.Random(5)
    k, ek = 4, 6
    n, N = 1 << k, 1 << ek
    a = [rnd.randrange(R_MOD) for _ in range(n)]
    z = bn254.FR_ZETA
    sc = [1, z, z * z % R_MOD]
    ext = ntt_run(a, bn254.omega(ek), ek, 3, 1, n_in=n, in_scale=sc)
    assert ext == bn254.dft_naive([a[i] * sc[i % 3] % R_MOD for i in range(n)] + [0] * (N - n), bn254.omega(ek))
    winv, ninv, zi = pow(bn254.omega(ek), -1, R_MOD), pow(N, -1, R_MOD), pow(z, -1, R_MOD)
    osc = [ninv, ninv * zi % R_MOD, ninv * zi * zi % R_MOD]
    assert ntt_run(ext, winv, ek, 3, 1, n_out=3 * n, out_scale=osc, mod3=True) == a + [0] * (2 * n)


def test_msm_pipeline_model():
    from msm_model import msm_model
    rnd = random.Random(6)

    def uni():
        return rnd.randrange(R_MOD)

    def adv():  # advice-like: 63 % zero, 30 % < 2^8, 7 % dense
        u = rnd.random()
        return 0 if u < 0.63 else (rnd.randrange(256) if u < 0.93 else rnd.randrange(R_MOD))

    def check(n, c, M, gen, **kw):
        g = [rnd.randrange(1, R_MOD) for _ in range(n)]
        sl = [[gen() for _ in range(n)] for _ in range(M)]
        assert msm_model(sl, g, c, **kw) == [sum(s * x for s, x in zip(sc, g)) % R_MOD for sc in sl], (n, c, M)

    for n, c, M in ((64, 6, 1), (64, 8, 2), (300, 7, 3), (1024, 9, 1)):
        check(n, c, M, uni)
        check(n, c, M, adv)
        check(n, c, M, uni, K0=4, serial_l1_threshold=64)
        check(n, c, M, adv, K0=4, serial_l1_threshold=64)
    # chunked digit sort with several chunks, both bucket level-1 forms, level-1 serial fold
    check(300, 7, 3, adv, J=5, l1_serial=True)
    check(300, 7, 3, uni, J=7, l1_serial=False, K0=4, serial_l1_threshold=64)
    check(1024, 9, 2, uni, J=3, l1_serial=True)
    # mixed-basis batch (the product round: MSM 1 reads the other basis)
    n, c = 200, 7
    g = [rnd.randrange(1, R_MOD) for _ in range(n)]
    g2 = [rnd.randrange(1, R_MOD) for _ in range(n)]
    sl = [[uni() for _ in range(n)] for _ in range(3)]
    exp = [sum(s * x for s, x in zip(sc, (g2 if m == 1 else g))) % R_MOD for m, sc in enumerate(sl)]
    assert msm_model(sl, g, c, alt=g2, alt_mask=0b010, J=4) == exp
    check(128, 6, 1, lambda: R_MOD - 1)
    check(128, 6, 2, lambda: 1)
    check(128, 8, 1, lambda: 0)
