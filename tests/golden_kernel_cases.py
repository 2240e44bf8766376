"""Seeded cases behind tests/golden/kernels.json: for every per-function entry point of the C ABI, the oracle
computation and the call through libzg_b200.so that must produce the same bytes."""
import hashlib

import numpy as np

import bn254
import cpu_ref
from bn254 import R_MOD

GEN = bn254.g1_affine_to_limbs([bn254.G1_GEN])[0]
SECRET = 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8


def rand_fr(n, seed):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 60) - 1)
    return cpu_ref.fr_to_mont(raw)


def digest(arr) -> str:
    return hashlib.sha256(np.ascontiguousarray(arr, dtype=np.uint64).tobytes()).hexdigest()


def L(x):
    return bn254.fr_to_limbs([x])[0]


K = 10
N = 1 << K


def _srs():
    return cpu_ref.srs_monomial(L(SECRET), GEN, N)


def _sparse(seed):
    s = rand_fr(N, seed)
    rng = np.random.default_rng(seed + 1)
    s[rng.random(N) < 0.6] = 0
    small = cpu_ref.fr_to_mont(np.stack([rng.integers(0, 256, size=N, dtype=np.uint64)] + [np.zeros(N, np.uint64)] * 3, axis=1))
    pick = rng.random(N) < 0.5
    s[pick] = small[pick]
    return np.ascontiguousarray(s)


def _lookup_inputs():
    rng = np.random.default_rng(77)
    vals = rand_fr(200, 78)
    table = vals[rng.integers(0, 200, size=N - 6)]
    table[:200] = vals
    inputs = table[rng.integers(0, N - 6, size=N - 6)]
    return np.ascontiguousarray(inputs), np.ascontiguousarray(table)


def _omega(k):
    return L(bn254.omega(k))


CASES = {
    "msm_dense_2^10": {
        "oracle": lambda: cpu_ref.g1_to_affine(cpu_ref.best_multiexp(rand_fr(N, 1), _srs()).reshape(1, 12)),
        "gpu": lambda ctx: (ctx.srs_load(K, _srs(), None), cpu_ref.g1_to_affine(ctx.msm(0, rand_fr(N, 1)).reshape(1, 12)))[1],
    },
    "msm_sparse_2^10": {
        "oracle": lambda: cpu_ref.g1_to_affine(cpu_ref.best_multiexp(_sparse(2), _srs()).reshape(1, 12)),
        "gpu": lambda ctx: (ctx.srs_load(K, _srs(), None), cpu_ref.g1_to_affine(ctx.msm(0, _sparse(2)).reshape(1, 12)))[1],
    },
    "ntt_2^10": {
        "oracle": lambda: cpu_ref.best_fft(rand_fr(N, 3), _omega(K), K),
        "gpu": lambda ctx: ctx.ntt(rand_fr(N, 3), K, _omega(K)),
    },
    "lagrange_to_coeff_2^10": {
        "oracle": lambda: cpu_ref.fr_scale_vec(cpu_ref.best_fft(rand_fr(N, 4), L(pow(bn254.omega(K), -1, R_MOD)), K),
                                               L(pow(N, -1, R_MOD))),
        "gpu": lambda ctx: ctx.lagrange_to_coeff(rand_fr(N, 4), K),
    },
    "coeff_to_extended_2^10_to_2^13": {
        "oracle": lambda: _coeff_to_extended_oracle(rand_fr(N, 5)),
        "gpu": lambda ctx: ctx.coeff_to_extended(rand_fr(N, 5), K, K + 3),
    },
    "lookup_permute": {
        "oracle": lambda: np.concatenate(cpu_ref.permute_expression_pair(*_lookup_inputs(), N - 6)),
        "gpu": lambda ctx: np.concatenate(ctx.lookup_permute(*_lookup_inputs())),
    },
    "grand_product": {
        "oracle": lambda: cpu_ref.fr_running_product(
            cpu_ref.fr_mul_vec(rand_fr(N, 6), cpu_ref.fr_batch_invert(rand_fr(N, 7))), L(1), N),
        "gpu": lambda ctx: ctx.grand_product(rand_fr(N, 6), rand_fr(N, 7)),
    },
    "eval_poly": {
        "oracle": lambda: np.stack([cpu_ref.fr_eval_poly(rand_fr(N, 8 + j), rand_fr(1, 20)[0]) for j in range(3)]),
        "gpu": lambda ctx: ctx.eval_poly_batch([rand_fr(N, 8 + j) for j in range(3)], rand_fr(1, 20)[0]),
    },
    "kate_division": {
        "oracle": lambda: cpu_ref.fr_kate_division(rand_fr(N, 11), rand_fr(1, 21)[0]),
        "gpu": lambda ctx: ctx.kate_division(rand_fr(N, 11), rand_fr(1, 21)[0]),
    },
}


def _coeff_to_extended_oracle(coeff):
    import halo2_ref as H
    dom = H.Domain(K, 9)          # degree 9 -> extended_k = K + 3
    assert dom.ext_k == K + 3
    return dom.coeff_to_extended(coeff)
