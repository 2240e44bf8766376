"""CUDA fixed-base Pippenger MSM (csrc/msm.cu through the C ABI) against the CPU oracle's
best_multiexp restatement: the affine normalisation of every result must be bit-identical."""
import numpy as np
import pytest

import bn254
import cpu_ref
from bn254 import R_MOD

pytestmark = pytest.mark.gpu

GEN = bn254.g1_affine_to_limbs([bn254.G1_GEN])[0]


def rand_fr(n, seed):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 60) - 1)
    return cpu_ref.fr_to_mont(raw)


def advice_like(n, seed):
    """63 % zero, 30 % below 2^8, 7 % dense."""
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    dense = rand_fr(n, seed + 1)
    small = np.zeros((n, 4), dtype=np.uint64)
    small[:, 0] = rng.integers(0, 256, size=n, dtype=np.uint64)
    small = cpu_ref.fr_to_mont(small)
    out = np.where((u < 0.93)[:, None], small, dense)
    out[u < 0.63] = 0
    return np.ascontiguousarray(out)


def affine(jac):
    return cpu_ref.g1_to_affine(np.asarray(jac).reshape(-1, 12))


@pytest.fixture(scope="module")
def srs12(ctx):
    k = 12
    s = bn254.fr_to_limbs([0x1F3C5A7B9D2E4F60718293A4B5C6D7E8])[0]
    g = cpu_ref.srs_monomial(s, GEN, 1 << k)
    gl = cpu_ref.g1_sequence(GEN, 1 << k)   # any second basis: the MSM does not care
    ctx.srs_load(k, g, gl)
    return k, g, gl


def test_msm_small_vs_python_double_and_add(ctx, srs12):
    k, g, gl = srs12
    n = 64
    import random
    rnd = random.Random(9)
    sc = [rnd.randrange(R_MOD) for _ in range(n)]
    sc[:5] = [0, 1, R_MOD - 1, 2, (1 << 253)]
    pts = bn254.g1_affine_from_limbs(g[:n].copy())
    exp = bn254.g1_msm_naive(sc, pts)
    got = bn254.g1_proj_from_limbs(ctx.msm(0, bn254.fr_to_limbs(sc)))[0]
    assert got == exp


@pytest.mark.parametrize("basis", [0, 1])
@pytest.mark.parametrize("kind", ["uniform", "advice", "ones", "minus_one", "zeros", "single"])
def test_msm_matches_oracle(ctx, srs12, basis, kind):
    k, g, gl = srs12
    n = 1 << k
    bases = g if basis == 0 else gl
    if kind == "uniform":
        s = rand_fr(n, 21)
    elif kind == "advice":
        s = advice_like(n, 22)
    elif kind == "ones":
        s = bn254.fr_to_limbs([1] * n)
    elif kind == "minus_one":
        s = bn254.fr_to_limbs([R_MOD - 1] * n)
    elif kind == "zeros":
        s = bn254.fr_to_limbs([0] * n)
    else:
        s = bn254.fr_to_limbs([0] * n)
        s[n - 1] = bn254.fr_to_limbs([R_MOD - 2])[0]
    got = affine(ctx.msm(basis, s))
    exp = affine(cpu_ref.best_multiexp(s, bases))
    assert (got == exp).all()


def test_msm_short_and_ragged_lengths(ctx, srs12):
    k, g, gl = srs12
    for n in (1, 2, 31, 33, 1000, (1 << k) - 1):
        s = rand_fr(n, 30 + n)
        assert (affine(ctx.msm(0, s)) == affine(cpu_ref.best_multiexp(s, g[:n]))).all(), n


def test_msm_batch_matches_oracle(ctx, srs12):
    k, g, gl = srs12
    n = 1 << k
    cols = [rand_fr(n, 50), advice_like(n, 51), bn254.fr_to_limbs([0] * n), advice_like(n, 52), rand_fr(n, 53), rand_fr(n, 54)]
    got = affine(ctx.msm_batch(1, cols))
    for j, c in enumerate(cols):
        assert (got[j] == affine(cpu_ref.best_multiexp(c, gl))[0]).all(), j


@pytest.mark.parametrize("k", [14, 16])
def test_msm_proof_sizes_match_oracle(ctx, k):
    n = 1 << k
    bases = cpu_ref.g1_sequence(GEN, n)
    ctx.srs_load(k, bases, None)
    for seed, s in ((1, rand_fr(n, 60 + k)), (2, advice_like(n, 61 + k))):
        assert (affine(ctx.msm(0, s)) == affine(cpu_ref.best_multiexp(s, bases))).all(), (k, seed)


def test_msm_large_linearity(ctx):
    """2^20 points (BASELINE sweep size): MSM(a) + MSM(b) == MSM(a + b), MSM(2a) == 2 MSM(a),
    and the dense result equals the oracle's."""
    k = 20
    n = 1 << k
    bases = cpu_ref.g1_sequence(GEN, n)
    ctx.srs_load(k, bases, None)
    a, b = rand_fr(n, 71), rand_fr(n, 72)
    ra, rb = ctx.msm(0, a), ctx.msm(0, b)
    rs = ctx.msm(0, cpu_ref.fr_add_vec(a, b))
    pa, pb, ps = (bn254.g1_proj_from_limbs(x)[0] for x in (ra, rb, rs))
    assert bn254.g1_add(pa, pb) == ps
    assert (affine(ra) == affine(cpu_ref.best_multiexp(a, bases))).all()


@pytest.mark.parametrize("logn", [22])
def test_msm_full_size_exact_closed_form(logn):
    """A sweep-sized MSM (BASELINE configs[4], 2^22 points -- beyond what the oracle's best_multiexp finishes in seconds) checked
    EXACTLY: with bases [1]G, [2]G, ..., [n]G the result must be [sum_i s_i (i + 1) mod r] G, one scalar multiplication on the
    host.  Uniform 254-bit scalars and the sparse advice-like mix; every window and bucket of the 2^22 pipeline is exercised.
    Own context: the module's shared one keeps the 2^12 SRS of the other tests."""
    import zg_b200
    n = 1 << logn
    bases = cpu_ref.g1_sequence(GEN, n)
    c = zg_b200.Context(0)
    try:
        c.srs_load(logn, bases, None)
        weights = np.arange(1, n + 1, dtype=object)
        for name, sc in (("uniform", rand_fr(n, 5)), ("advice-like", advice_like(n, 6))):
            canon = cpu_ref.fr_from_mont(sc)
            total = 0
            for limb in range(4):                    # sum_i s_i * (i + 1), limb by limb in Python integers
                total += int((canon[:, limb].astype(object) * weights).sum()) << (64 * limb)
            expect = bn254.g1_mul(bn254.G1_GEN, total % R_MOD)
            got = bn254.g1_proj_from_limbs(c.msm(0, sc))[0]
            assert got == expect, name
    finally:
        c.close()
