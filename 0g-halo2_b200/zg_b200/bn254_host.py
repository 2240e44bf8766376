"""BN254 constants and limb-layout helpers for the host side of the C ABI.

Values follow halo2curves 0.3.3 `bn256::Fr` / `Fq` (pinned by /root/reference/Cargo.toml:14-18;
SURVEY.md Appendix D).  Host bookkeeping only (domain generators, conversions to the
4 x u64 Montgomery layout): no prover arithmetic happens here."""
from __future__ import annotations

import numpy as np

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
FR_S = 28
FR_ROOT_OF_UNITY = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
FR_DELTA = 0x09226B6E22C6F0CA64EC26AAD4C86E715B5F898E5E963F25870E56BBE533E9A2
FR_ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
_MASK = (1 << 64) - 1


def omega(k: int) -> int:
    """generator of the 2^k-point evaluation domain (EvaluationDomain::new)."""
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_S - k), R_MOD)


def to_limbs(vals, mod: int = R_MOD) -> np.ndarray:
    """ints -> (n,4) uint64 Montgomery limbs."""
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v = ((v % mod) << 256) % mod
        out[i] = (v & _MASK, (v >> 64) & _MASK, (v >> 128) & _MASK, (v >> 192) & _MASK)
    return out


def from_limbs(arr, mod: int = R_MOD) -> list:
    arr = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
    rinv = pow(1 << 256, -1, mod)
    return [((int(r[0]) | (int(r[1]) << 64) | (int(r[2]) << 128) | (int(r[3]) << 192)) * rinv) % mod for r in arr]


def jacobian_to_affine(jac) -> list:
    """(m,12) Jacobian limbs -> [(x,y) | None] canonical ints (C::Curve::batch_normalize)."""
    v = from_limbs(np.asarray(jac).reshape(-1, 4), Q_MOD)
    out = []
    for i in range(0, len(v), 3):
        x, y, z = v[i:i + 3]
        if z == 0:
            out.append(None)
            continue
        zi = pow(z, -1, Q_MOD)
        out.append((x * zi * zi % Q_MOD, y * zi * zi * zi % Q_MOD))
    return out
