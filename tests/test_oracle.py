"""Pins the CPU oracle (oracle/zg_oracle.c via cpu_ref) against self-evident identities:
Python big-int field arithmetic, the naive DFT and double-and-add MSM (SURVEY.md section 8c:
the reference holds no golden vectors for prover arithmetic -- parity unpinned vs the real crate)."""
import random

import numpy as np
import pytest

import bn254
import cpu_ref
from bn254 import R_MOD, Q_MOD


def test_constants_match_survey_appendix_d():
    bn254._self_check()
    assert bn254.omega(28) == bn254.FR_ROOT_OF_UNITY
    assert pow(bn254.omega(14), 1 << 14, R_MOD) == 1 and pow(bn254.omega(14), 1 << 13, R_MOD) == R_MOD - 1


def test_field_vectors():
    rnd = random.Random(11)
    a = [0, 1, R_MOD - 1] + [rnd.randrange(R_MOD) for _ in range(200)]
    b = [R_MOD - 1, R_MOD - 1, R_MOD - 1] + [rnd.randrange(R_MOD) for _ in range(200)]
    A, B = bn254.fr_to_limbs(a), bn254.fr_to_limbs(b)
    assert bn254.fr_from_limbs(cpu_ref.fr_mul_vec(A, B)) == [x * y % R_MOD for x, y in zip(a, b)]
    assert bn254.fr_from_limbs(cpu_ref.fr_add_vec(A, B)) == [(x + y) % R_MOD for x, y in zip(a, b)]
    assert bn254.fr_from_limbs(cpu_ref.fr_sub_vec(A, B)) == [(x - y) % R_MOD for x, y in zip(a, b)]
    inv = bn254.fr_from_limbs(cpu_ref.fr_batch_invert(A))
    assert inv == [pow(x, -1, R_MOD) if x else 0 for x in a]
    raw = cpu_ref.fr_from_mont(A)
    assert [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in raw] == a
    assert (cpu_ref.fr_to_mont(raw) == A).all()


@pytest.mark.parametrize("log_n", [1, 2, 5, 7])
def test_best_fft_matches_naive_dft(log_n):
    rnd = random.Random(log_n)
    n = 1 << log_n
    a = [rnd.randrange(R_MOD) for _ in range(n)]
    w = bn254.omega(log_n)
    got = bn254.fr_from_limbs(cpu_ref.best_fft(bn254.fr_to_limbs(a), bn254.fr_to_limbs([w]), log_n))
    assert got == bn254.dft_naive(a, w)


def test_best_fft_large_matches_bigint_ntt_and_threads_agree():
    rnd = random.Random(99)
    log_n = 13
    a = [rnd.randrange(R_MOD) for _ in range(1 << log_n)]
    w = bn254.omega(log_n)
    A, W = bn254.fr_to_limbs(a), bn254.fr_to_limbs([w])
    r1 = cpu_ref.best_fft(A, W, log_n, threads=1)
    r8 = cpu_ref.best_fft(A, W, log_n, threads=8)
    assert (r1 == r8).all()
    assert bn254.fr_from_limbs(r1) == bn254.ntt(a, w)


def test_fft_output_is_the_polynomial_at_powers_of_omega():
    """output i of best_fft == Horner evaluation at omega^i (big-int check of the evaluation, then the identity that
    tests/test_gpu_ntt.py uses to check sampled outputs of transforms too large to restate in full)"""
    rnd = random.Random(21)
    log_n = 11
    n = 1 << log_n
    a = [rnd.randrange(R_MOD) for _ in range(n)]
    A = bn254.fr_to_limbs(a)
    w = bn254.omega(log_n)
    f = cpu_ref.best_fft(A, bn254.fr_to_limbs([w]), log_n)
    for i in (0, 1, 2, 777, n // 3, n // 2, n - 2, n - 1):
        x = pow(w, i, R_MOD)
        got = cpu_ref.fr_eval_poly(A, bn254.fr_to_limbs([x])[0])
        assert bn254.fr_from_limbs(got.reshape(1, 4)) == [sum(c * pow(x, j, R_MOD) for j, c in enumerate(a)) % R_MOD]
        assert (got == f[i]).all(), i


def test_best_multiexp_matches_double_and_add():
    rnd = random.Random(5)
    n = 48
    pts = [bn254.g1_mul(bn254.G1_GEN, rnd.randrange(R_MOD)) for _ in range(n)]
    sc = [rnd.randrange(R_MOD) for _ in range(n)]
    sc[:4] = [0, 1, R_MOD - 1, 2]
    pts[5] = None  # identity base
    exp = bn254.g1_msm_naive(sc, pts)
    P, S = bn254.g1_affine_to_limbs(pts), bn254.fr_to_limbs(sc)
    for threads in (1, 3, 8):
        got = bn254.g1_proj_from_limbs(cpu_ref.best_multiexp(S, P, threads))[0]
        assert got == exp
    assert bn254.g1_affine_from_limbs(cpu_ref.g1_to_affine(cpu_ref.msm_naive(S, P)))[0] == exp


def test_edge_cases_empty_and_cancelling():
    g = bn254.G1_GEN
    P = bn254.g1_affine_to_limbs([g, bn254.g1_neg(g), g, g])
    S = bn254.fr_to_limbs([5, 5, 0, 0])
    assert bn254.g1_proj_from_limbs(cpu_ref.best_multiexp(S, P))[0] is None
    S = bn254.fr_to_limbs([1, 0, 1, 0])  # doubling inside a bucket
    assert bn254.g1_proj_from_limbs(cpu_ref.best_multiexp(S, P))[0] == bn254.g1_mul(g, 2)


def test_srs_monomial_is_powers_of_s():
    s = 0x1234567890ABCDEF1234567890ABCDEF
    srs = cpu_ref.srs_monomial(bn254.fr_to_limbs([s])[0], bn254.g1_affine_to_limbs([bn254.G1_GEN])[0], 8)
    assert bn254.g1_affine_from_limbs(srs) == bn254.g1_powers(8, s)
    seq = cpu_ref.g1_sequence(bn254.g1_affine_to_limbs([bn254.G1_GEN])[0], 5000)
    got = bn254.g1_affine_from_limbs(seq[[0, 1, 4095, 4096, 4999]].copy())
    assert got == [bn254.g1_mul(bn254.G1_GEN, k) for k in (1, 2, 4096, 4097, 5000)]


def test_pairing_bilinear_nondegenerate_order_r():
    """pins the verifier's pairing (oracle/zg_oracle.c): generator on the twist and of order r, e(aP, bQ) = e(P, Q)^(ab),
    e(P, Q) != 1, e(P, Q)^r = 1, and the product check used for e(W', [s]G2) = e(R, G2)."""
    L = lambda x: bn254.fr_to_limbs([x])[0]
    g2 = cpu_ref.g2_generator()
    assert cpu_ref.g2_on_curve(g2)
    assert (cpu_ref.g2_mul(g2, L(0)) == 0).all()
    minus = cpu_ref.g2_mul(g2, L(R_MOD - 1))
    assert (minus[:8] == g2[:8]).all() and (minus[8:] != g2[8:]).any()       # [r-1]G2 = -G2
    G1 = bn254.g1_affine_to_limbs([bn254.G1_GEN])[0]
    e = cpu_ref.pairing(G1, g2)
    one = cpu_ref.f12_pow(e, 0)
    assert not (e == one).all()
    assert (cpu_ref.f12_pow(e, R_MOD) == one).all()
    a, b = 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8, 987654321987654321
    aP = bn254.g1_affine_to_limbs([bn254.g1_mul(bn254.G1_GEN, a)])[0]
    bQ = cpu_ref.g2_mul(g2, L(b))
    assert (cpu_ref.pairing(aP, bQ) == cpu_ref.f12_pow(e, a * b % R_MOD)).all()
    negP = bn254.g1_affine_to_limbs([bn254.g1_neg(bn254.G1_GEN)])[0]
    assert cpu_ref.pairing_check(np.stack([aP, negP]), np.stack([g2, cpu_ref.g2_mul(g2, L(a))]))
    assert not cpu_ref.pairing_check(np.stack([aP, negP]), np.stack([g2, cpu_ref.g2_mul(g2, L(a + 1))]))
    zero = np.zeros(8, dtype=np.uint64)
    assert (cpu_ref.pairing(zero, g2) == one).all()                          # identity maps to 1
