"""CUDA NTT family (csrc/ntt.cu through the C ABI) against the CPU oracle's best_fft restatement,
bit-exact, on the same seeded inputs; property checks at sizes the oracle would take too long for."""
import numpy as np
import pytest

import bn254
import cpu_ref
from bn254 import R_MOD

pytestmark = pytest.mark.gpu


def rand_fr(n, seed):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 60) - 1)       # < 2^252 < r: already canonical
    return cpu_ref.fr_to_mont(raw)               # uniform-ish Montgomery elements


def W(x):
    return bn254.fr_to_limbs([x])


@pytest.mark.parametrize("log_n", list(range(1, 13)) + [14, 16, 17])
def test_ntt_matches_oracle(ctx, log_n):
    a = rand_fr(1 << log_n, 11 + log_n)
    w = W(bn254.omega(log_n))
    assert (ctx.ntt(a, log_n, w) == cpu_ref.best_fft(a, w, log_n)).all()


def test_ntt_small_matches_naive_dft(ctx):
    import random
    rnd = random.Random(1)
    for log_n in (1, 2, 3, 5):
        a = [rnd.randrange(R_MOD) for _ in range(1 << log_n)]
        w = bn254.omega(log_n)
        got = bn254.fr_from_limbs(ctx.ntt(bn254.fr_to_limbs(a), log_n, W(w)))
        assert got == bn254.dft_naive(a, w)


def test_ntt_edge_inputs(ctx):
    log_n = 10
    n = 1 << log_n
    w = W(bn254.omega(log_n))
    zeros = bn254.fr_to_limbs([0] * n)
    assert (ctx.ntt(zeros, log_n, w) == zeros).all()
    delta = bn254.fr_to_limbs([1] + [0] * (n - 1))
    assert (ctx.ntt(delta, log_n, w) == bn254.fr_to_limbs([1] * n)).all()
    big = bn254.fr_to_limbs([R_MOD - 1] * n)
    assert (ctx.ntt(big, log_n, w) == cpu_ref.best_fft(big, w, log_n)).all()


def test_ntt_arbitrary_omega(ctx):
    """best_fft takes any omega; use the inverse root and check forward*inverse = n * identity."""
    log_n = 12
    a = rand_fr(1 << log_n, 5)
    w = bn254.omega(log_n)
    f = ctx.ntt(a, log_n, W(w))
    b = ctx.ntt(f, log_n, W(pow(w, -1, R_MOD)))
    assert (b == cpu_ref.fr_scale_vec(a, W(1 << log_n)[0])).all()


@pytest.mark.parametrize("k", [4, 9, 14, 15])
def test_lagrange_to_coeff_matches_oracle(ctx, k):
    n = 1 << k
    a = rand_fr(n, 40 + k)
    winv = W(pow(bn254.omega(k), -1, R_MOD))
    exp = cpu_ref.fr_scale_vec(cpu_ref.best_fft(a, winv, k), W(pow(n, -1, R_MOD))[0])
    assert (ctx.lagrange_to_coeff(a, k) == exp).all()


@pytest.mark.parametrize("k,ext_k", [(4, 6), (9, 12), (14, 17), (15, 18)])
def test_extended_domain_round_trip_matches_oracle(ctx, k, ext_k):
    """coeff_to_extended / extended_to_coeff as halo2 EvaluationDomain restated with oracle pieces:
    distribute_powers_zeta, zero-pad, best_fft(extended_omega); inverse + 1/n + un-zeta + truncate."""
    n, ne = 1 << k, 1 << ext_k
    c = rand_fr(n, 70 + k)
    z = bn254.FR_ZETA
    pw = bn254.fr_to_limbs([1, z, z * z % R_MOD])
    padded = np.zeros((ne, 4), dtype=np.uint64)
    padded[:n] = cpu_ref.fr_scale_mod3(c, pw)
    exp_ext = cpu_ref.best_fft(padded, W(bn254.omega(ext_k)), ext_k)
    got_ext = ctx.coeff_to_extended(c, k, ext_k)
    assert (got_ext == exp_ext).all()
    # back: keep quotient-degree * n coefficients like extended_to_coeff's truncate
    keep = min(ne, 5 * n)
    back = ctx.extended_to_coeff(got_ext, k, ext_k, keep)
    inv = cpu_ref.best_fft(exp_ext, W(pow(bn254.omega(ext_k), -1, R_MOD)), ext_k)
    inv = cpu_ref.fr_scale_vec(inv, W(pow(ne, -1, R_MOD))[0])
    zi = pow(z, -1, R_MOD)
    inv = cpu_ref.fr_scale_mod3(inv, bn254.fr_to_limbs([1, zi, zi * zi % R_MOD]))
    assert (back == inv[:keep]).all()
    assert (back[:n] == c).all() and not back[n:].any()


def test_ntt_batch_dev(ctx):
    import torch
    log_n, batch = 13, 5
    n = 1 << log_n
    a = np.stack([rand_fr(n, 200 + j) for j in range(batch)])
    t = torch.from_numpy(a.view(np.int64)).cuda()
    w = W(bn254.omega(log_n))
    ctx.ntt_dev(t.data_ptr(), t.data_ptr(), log_n, w, batch, n)
    ctx.sync()
    got = t.cpu().numpy().view(np.uint64)
    for j in range(batch):
        assert (got[j] == cpu_ref.best_fft(a[j], w, log_n)).all(), j


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_large_properties(ctx, log_n):
    """sizes where the oracle is slow: inverse(forward(a)) == a, and linearity."""
    import torch
    n = 1 << log_n
    a = rand_fr(n, 300 + log_n)
    b = rand_fr(n, 400 + log_n)
    ta = torch.from_numpy(a.view(np.int64)).cuda()
    tb = torch.from_numpy(b.view(np.int64)).cuda()
    ts = torch.from_numpy(cpu_ref.fr_add_vec(a, b).view(np.int64)).cuda()
    w = W(bn254.omega(log_n))
    for t in (ta, tb, ts):
        ctx.ntt_dev(t.data_ptr(), t.data_ptr(), log_n, w)
    ctx.sync()
    fa, fb, fs = (t.cpu().numpy().view(np.uint64) for t in (ta, tb, ts))
    assert (cpu_ref.fr_add_vec(fa, fb) == fs).all()
    ctx.lagrange_to_coeff_dev(ta.data_ptr(), ta.data_ptr(), log_n)
    ctx.sync()
    assert (ta.cpu().numpy().view(np.uint64) == a).all()
    if log_n == 20:
        assert (fa == cpu_ref.best_fft(a, w, log_n)).all()
    else:
        # exact on sampled outputs: output i of the transform is the polynomial evaluated at omega^i (the oracle's Horner
        # evaluation is O(n) per point; tests/test_oracle.py checks this identity against best_fft on the CPU)
        om = bn254.omega(log_n)
        for i in (0, 1, 2, 12345, n // 3, n // 2, n - 2, n - 1):
            assert (fa[i] == cpu_ref.fr_eval_poly(a, W(pow(om, i, R_MOD))[0])).all(), i
