"""Per-stage parity of the single-function C-ABI entry points (csrc/stages.cu, zg_evaluate_h in csrc/prover.cu)
against the CPU oracle: SURVEY.md section 8 rows a7 (evaluate_h), a8 (divide_by_vanishing_poly), a9
(permute_expression_pair), a10 (grand products), a12 (eval_polynomial), a13 (kate_division).  All bit-exact."""
import os

import numpy as np
import pytest

import bn254
import cpu_ref
import halo2_ref as H
from bn254 import R_MOD

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SEED = bytes(range(16))


def rand_fr(n, seed):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 60) - 1)
    return cpu_ref.fr_to_mont(raw)


def L(x):
    return bn254.fr_to_limbs([x])[0]


# ---- a9: permute_expression_pair ----------------------------------------------------------------------------
def _lookup_case(n, n_distinct, seed, wide):
    """table with duplicates (a padded table column) and inputs drawn from it, heavily repeated"""
    rng = np.random.default_rng(seed)
    if wide:
        vals = rand_fr(n_distinct, seed + 1)                      # full 254-bit keys: exercises all sort passes
    else:
        vals = bn254.fr_to_limbs([int(v) for v in rng.integers(0, 1 << 20, size=n_distinct)])
    table = vals[rng.integers(0, n_distinct, size=n)]
    table[:n_distinct] = vals                                       # every value present at least once
    inputs = table[rng.integers(0, n, size=n)]
    inputs[: n // 3] = table[0]                                     # one value dominates (the default lookup row)
    return np.ascontiguousarray(inputs), np.ascontiguousarray(table)


@pytest.mark.parametrize("n,n_distinct,wide", [(1, 1, False), (2, 1, False), (33, 5, False), (1000, 1000, True),
                                               (4097, 300, True), ((1 << 14) - 6, 257, False), ((1 << 15) - 6, 5000, True)])
def test_lookup_permute_matches_oracle(ctx, n, n_distinct, wide):
    a, s = _lookup_case(n, n_distinct, 100 + n, wide)
    pa, ps = ctx.lookup_permute(a, s)
    ea, es = cpu_ref.permute_expression_pair(a, s, n)
    assert (pa == ea).all()
    assert (ps == es).all()


def test_lookup_permute_reports_missing_input(ctx):
    import zg_b200
    a, s = _lookup_case(1000, 50, 7, False)
    a[123] = L(R_MOD - 5)                                           # not in the table
    with pytest.raises(zg_b200.ZgError) as ei:
        ctx.lookup_permute(a, s)
    assert ei.value.code == -5
    with pytest.raises(ValueError):
        cpu_ref.permute_expression_pair(a, s, 1000)


# ---- a10: grand product ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 2049, 1 << 14, (1 << 17) - 5])
def test_grand_product_matches_oracle(ctx, n):
    num, den = rand_fr(n, 200 + n), rand_fr(n, 300 + n)
    z = ctx.grand_product(num, den)
    f = cpu_ref.fr_mul_vec(num, cpu_ref.fr_batch_invert(den))
    assert (z == cpu_ref.fr_running_product(f, L(1), n)).all()


def test_grand_product_of_a_permutation_closes(ctx):
    """domain property: when den is a permutation of num the running product returns to 1."""
    n = 1 << 12
    num = rand_fr(n, 5)
    den = num[np.random.default_rng(6).permutation(n)]
    z = ctx.grand_product(np.concatenate([num, num[:1]]), np.concatenate([den, den[:1]]))
    assert (z[0] == L(1)).all() and (z[n] == L(1)).all()


def test_batch_invert_with_zeros(ctx):
    a = rand_fr(5000, 9)
    a[0] = 0
    a[77] = 0
    a[4999] = 0
    inv = ctx.batch_invert(a)
    assert (inv == cpu_ref.fr_batch_invert(a)).all()
    prod = cpu_ref.fr_mul_vec(a, inv)
    one = L(1)
    nz = np.ones(5000, dtype=bool)
    nz[[0, 77, 4999]] = False
    assert (prod[nz] == one).all() and (inv[~nz] == 0).all()


# ---- a12 / a13: eval_polynomial, kate_division ------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 17, 4096, 4097, 1 << 15, 1 << 19])
def test_eval_poly_batch_matches_oracle(ctx, n):
    polys = [rand_fr(n, 400 + j + n) for j in range(3)]
    polys.append(np.zeros((n, 4), dtype=np.uint64))
    x = rand_fr(1, 11)[0]
    got = ctx.eval_poly_batch(polys, x)
    for j, p in enumerate(polys):
        assert (got[j] == cpu_ref.fr_eval_poly(p, x)).all(), j


@pytest.mark.parametrize("n", [2, 3, 33, 2049, 1 << 15, (1 << 17) + 1])
def test_kate_division_matches_oracle(ctx, n):
    a, z = rand_fr(n, 500 + n), rand_fr(1, 13)[0]
    q = ctx.kate_division(a, z)
    assert (q == cpu_ref.fr_kate_division(a, z)).all()
    # q(X) (X - z) + a(z) == a(X) at a random point
    r = rand_fr(1, 14)[0]
    I = lambda v: bn254.fr_from_limbs(np.asarray(v).reshape(1, 4))[0]
    qa, az, ar = I(cpu_ref.fr_eval_poly(q, r)), I(cpu_ref.fr_eval_poly(a, z)), I(cpu_ref.fr_eval_poly(a, r))
    assert (qa * (I(r) - I(z)) + az) % R_MOD == ar


# ---- a7 / a8: evaluate_h on the WNN circuit -------------------------------------------------------------------------
def test_evaluate_h_matches_oracle(ctx):
    """tiny model (k = 14, extended domain 2^17): the oracle's create_proof is traced for the inputs and output of
    Evaluator::evaluate_h; zg_evaluate_h must return the same 2^17 values, before and after the division by X^n - 1."""
    from zg_b200.io import load_wnn, load_grayscale_image
    from zg_b200.prover import ParamsKZG, keygen
    wnn = load_wnn(os.path.join(GOLD, "model_28input_256entry_1hash_1bpi.hdf5"))
    img = load_grayscale_image(os.path.join(GOLD, "example_image_7.png"))
    k = 14
    srs = H.Srs(k, 0x1F3C5A7B9D2E4F60718293A4B5C6D7E8)
    outputs = wnn.predict(img)
    zero = np.zeros((28, 28), dtype=np.uint8)
    circ0, asm0 = wnn.synthesize(zero, k)
    opk = H.keygen(srs, circ0.cs, asm0)
    _, asm = wnn.synthesize(img, k)
    T = {}
    H.create_proof(srs, opk, asm.advice, [outputs], H.XorShiftRng(SEED), trace=T)
    circ1, asm1 = wnn.synthesize(zero, k)
    pk = keygen(ctx, ParamsKZG(k, srs.g, srs.g_lagrange), circ1.cs, asm1)
    E, ch = T["evaluate_h"], T["challenges"]
    chal = bn254.fr_to_limbs([ch["theta"], ch["beta"], ch["gamma"], ch["y"]])
    args = (E["advice_polys"], E["instance_polys"], E["lookup_input_polys"], E["lookup_table_polys"],
            E["lookup_product_polys"], E["perm_product_polys"], chal)
    ext_n = E["h_numerator"].shape[0]
    assert ext_n == 1 << 17
    h = ctx.evaluate_h(pk._h, ext_n, *args)
    assert (h == E["h_numerator"]).all()
    hd = ctx.evaluate_h(pk._h, ext_n, *args, divide=True)
    assert (hd == E["h_divided"]).all()
    # the lookup columns the oracle permuted inside that proof go through zg_lookup_permute unchanged
    usable = (1 << k) - (circ0.cs.blinding_factors() + 1)
    for lc in T["lookup_columns"]:
        pa, ps = ctx.lookup_permute(lc["ci"][:usable], lc["ct"][:usable])
        assert (pa == lc["pa"][:usable]).all() and (ps == lc["ps"][:usable]).all()
    pk.close()
