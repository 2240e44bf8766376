"""BN254 big-integer oracle (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Obviously-correct Python-int arithmetic for the two BN254 fields, the G1 group, naive
MSM / DFT, and the byte layouts used at the C-ABI boundary.  Everything here is the
published algorithm of `halo2curves` tag 0.3.3 (`bn256::{Fr,Fq,G1}`), which is pinned by
/root/reference/Cargo.toml:14-18,26-28 but is NOT vendored under /root/reference
(un-vendored git dependency).  Constants follow SURVEY.md Appendix D and are re-derived
and self-checked at import time (see `_self_check`).

Parity status: "parity unpinned" for prover arithmetic -- the reference holds no golden
vectors for field/curve/MSM/NTT results.  What pins this file are
self-evident identities (Fermat, curve equation, root-of-unity orders, group laws).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------
# Field constants (halo2curves::bn256::{Fr, Fq})
# ---------------------------------------------------------------------------------------
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # Fr modulus r
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # Fq modulus q
FR_S = 28                      # 2-adicity of r-1
FR_GENERATOR = 7               # multiplicative generator used by halo2curves
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R_MOD - 1) >> FR_S, R_MOD)   # order 2^28
FR_DELTA = pow(FR_GENERATOR, 1 << FR_S, R_MOD)                      # generator of the odd part
FR_ZETA = pow(FR_GENERATOR, 2 * (R_MOD - 1) // 3, R_MOD)             # halo2curves' Fr::ZETA
MONT_R = 1 << 256
FR_R = MONT_R % R_MOD
FR_R2 = (MONT_R * MONT_R) % R_MOD
FR_R3 = (MONT_R * MONT_R * MONT_R) % R_MOD
FQ_R = MONT_R % Q_MOD
FQ_R2 = (MONT_R * MONT_R) % Q_MOD
FR_INV64 = (-pow(R_MOD, -1, 1 << 64)) % (1 << 64)
FQ_INV64 = (-pow(Q_MOD, -1, 1 << 64)) % (1 << 64)
CURVE_B = 3
G1_GEN = (1, 2)


def _self_check() -> None:
    assert FR_ROOT_OF_UNITY == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert FR_DELTA == 0x09226B6E22C6F0CA64EC26AAD4C86E715B5F898E5E963F25870E56BBE533E9A2
    assert FR_ZETA == 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
    assert FR_R == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert FR_R2 == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7
    assert FR_INV64 == 0xC2E1F593EFFFFFFF and FQ_INV64 == 0x87D20782E4866389
    assert pow(FR_ROOT_OF_UNITY, 1 << FR_S, R_MOD) == 1
    assert pow(FR_ROOT_OF_UNITY, 1 << (FR_S - 1), R_MOD) == R_MOD - 1
    assert pow(FR_ZETA, 3, R_MOD) == 1 and FR_ZETA != 1
    assert (G1_GEN[1] ** 2 - G1_GEN[0] ** 3 - CURVE_B) % Q_MOD == 0


_self_check()


def fr_inv(a: int) -> int:
    return pow(a, -1, R_MOD)


def fq_inv(a: int) -> int:
    return pow(a, -1, Q_MOD)


def omega(k: int) -> int:
    """Primitive 2^k-th root of unity as EvaluationDomain::new derives it
    (halo2_proofs poly/domain.rs, tag v2023_04_20: ROOT_OF_UNITY squared S-k times)."""
    assert 0 <= k <= FR_S
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_S - k), R_MOD)


# ---------------------------------------------------------------------------------------
# Byte layouts at the C-ABI boundary: 4 x u64 little-endian limbs, Montgomery form R=2^256
# (halo2curves in-memory layout of Fr / Fq)
# ---------------------------------------------------------------------------------------
def ints_to_limbs(vals, mod: int, mont: bool = True) -> np.ndarray:
    """list of ints -> (n,4) uint64 array (Montgomery form when mont=True)."""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    mask = (1 << 64) - 1
    for i, v in enumerate(vals):
        v %= mod
        if mont:
            v = (v << 256) % mod
        out[i, 0] = v & mask
        out[i, 1] = (v >> 64) & mask
        out[i, 2] = (v >> 128) & mask
        out[i, 3] = (v >> 192) & mask
    return out


def limbs_to_ints(arr: np.ndarray, mod: int, mont: bool = True) -> list:
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(MONT_R, -1, mod)
    out = []
    for row in arr:
        v = int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)
        if mont:
            v = (v * rinv) % mod
        out.append(v)
    return out


def fr_to_limbs(vals) -> np.ndarray:
    return ints_to_limbs(vals, R_MOD)


def fr_from_limbs(arr) -> list:
    return limbs_to_ints(arr, R_MOD)


def g1_affine_to_limbs(points) -> np.ndarray:
    """[(x,y) | None] -> (n,8) uint64; identity encoded as (0,0) like halo2curves G1Affine."""
    flat = []
    for p in points:
        if p is None:
            flat += [0, 0]
        else:
            flat += [p[0], p[1]]
    return ints_to_limbs(flat, Q_MOD).reshape(-1, 8)


def g1_affine_from_limbs(arr) -> list:
    vals = limbs_to_ints(np.asarray(arr).reshape(-1, 4), Q_MOD)
    pts = []
    for i in range(0, len(vals), 2):
        x, y = vals[i], vals[i + 1]
        pts.append(None if (x == 0 and y == 0) else (x, y))
    return pts


def g1_proj_from_limbs(arr) -> list:
    """(n,12) uint64 {x,y,z} -> affine tuples / None.  Accepts both homogeneous-Jacobian
    conventions?  No: the ABI fixes JACOBIAN coordinates (x/z^2, y/z^3) like halo2curves G1."""
    vals = limbs_to_ints(np.asarray(arr).reshape(-1, 4), Q_MOD)
    pts = []
    for i in range(0, len(vals), 3):
        x, y, z = vals[i], vals[i + 1], vals[i + 2]
        if z == 0:
            pts.append(None)
        else:
            zi = fq_inv(z)
            zi2 = zi * zi % Q_MOD
            pts.append((x * zi2 % Q_MOD, y * zi2 * zi % Q_MOD))
    return pts


# ---------------------------------------------------------------------------------------
# G1: y^2 = x^3 + 3 over Fq.  Affine tuples, None = identity.  Jacobian internally.
# ---------------------------------------------------------------------------------------
def g1_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - CURVE_B) % Q_MOD == 0


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q_MOD)


def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    q = Q_MOD
    A = X * X % q
    B = Y * Y % q
    C = B * B % q
    D = 2 * ((X + B) * (X + B) - A - C) % q
    E = 3 * A % q
    F = E * E % q
    X3 = (F - 2 * D) % q
    Y3 = (E * (D - X3) - 8 * C) % q
    Z3 = 2 * Y * Z % q
    return (X3, Y3, Z3)


def _jac_add(P, Q):
    q = Q_MOD
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % q
    Z2Z2 = Z2 * Z2 % q
    U1 = X1 * Z2Z2 % q
    U2 = X2 * Z1Z1 % q
    S1 = Y1 * Z2 * Z2Z2 % q
    S2 = Y2 * Z1 * Z1Z1 % q
    if U1 == U2:
        if S1 == S2:
            return _jac_double(P)
        return (1, 1, 0)
    H = (U2 - U1) % q
    Rr = (S2 - S1) % q
    HH = H * H % q
    HHH = H * HH % q
    V = U1 * HH % q
    X3 = (Rr * Rr - HHH - 2 * V) % q
    Y3 = (Rr * (V - X3) - S1 * HHH) % q
    Z3 = Z1 * Z2 * H % q
    return (X3, Y3, Z3)


def _to_jac(p):
    return (1, 1, 0) if p is None else (p[0], p[1], 1)


def _to_aff(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = fq_inv(Z)
    zi2 = zi * zi % Q_MOD
    return (X * zi2 % Q_MOD, Y * zi2 * zi % Q_MOD)


def g1_add(p, q):
    return _to_aff(_jac_add(_to_jac(p), _to_jac(q)))


def g1_mul(p, k: int):
    k %= R_MOD
    acc = (1, 1, 0)
    base = _to_jac(p)
    while k:
        if k & 1:
            acc = _jac_add(acc, base)
        base = _jac_double(base)
        k >>= 1
    return _to_aff(acc)


def g1_msm_naive(scalars, points):
    """Sum s_i * P_i by double-and-add: the mathematically unique value every MSM must hit."""
    acc = (1, 1, 0)
    for s, p in zip(scalars, points):
        s %= R_MOD
        if s == 0 or p is None:
            continue
        base = _to_jac(p)
        t = (1, 1, 0)
        while s:
            if s & 1:
                t = _jac_add(t, base)
            base = _jac_double(base)
            s >>= 1
        acc = _jac_add(acc, t)
    return _to_aff(acc)


def g1_powers(n: int, s: int, base=G1_GEN):
    """[s^i]G for i<n -- the monomial-basis SRS `ParamsKZG::g` for a KNOWN toy secret s
    (test-only; ParamsKZG::new draws s from OsRng, /root/reference/src/main.rs:232)."""
    out = []
    cur = 1
    for _ in range(n):
        out.append(g1_mul(base, cur))
        cur = cur * s % R_MOD
    return out


# ---------------------------------------------------------------------------------------
# Naive transforms: the unique values best_fft must hit
# ---------------------------------------------------------------------------------------
def dft_naive(a, w):
    """out[i] = sum_j a[j] * w^(i*j)   (best_fft semantics: natural order in and out)."""
    n = len(a)
    out = []
    for i in range(n):
        wi = pow(w, i, R_MOD)
        acc = 0
        cur = 1
        for j in range(n):
            acc = (acc + a[j] * cur) % R_MOD
            cur = cur * wi % R_MOD
        out.append(acc)
    return out


def ntt(a, w):
    """O(n log n) recursive NTT in Python ints, same semantics as dft_naive."""
    n = len(a)
    if n == 1:
        return list(a)
    even = ntt(a[0::2], w * w % R_MOD)
    odd = ntt(a[1::2], w * w % R_MOD)
    out = [0] * n
    cur = 1
    h = n // 2
    for i in range(h):
        t = cur * odd[i] % R_MOD
        out[i] = (even[i] + t) % R_MOD
        out[i + h] = (even[i] - t) % R_MOD
        cur = cur * w % R_MOD
    return out


def eval_poly(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc
