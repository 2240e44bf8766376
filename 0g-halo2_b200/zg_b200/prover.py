"""keygen / create_proof on the B200 backend: the host-side mirror of
`halo2_proofs::plonk::{keygen_vk, keygen_pk, create_proof}` and `ParamsKZG<Bn256>` as used at
/root/reference/src/wnn.rs:222-262.  Thin marshalling over the C ABI (zg_pk_load, zg_create_proof);
there is no CPU path."""
from __future__ import annotations

import ctypes
import hashlib

import numpy as np

from . import lib as zl
from .bn254_host import Q_MOD, R_MOD, from_limbs, to_limbs
from .plonk.mock import finalize_fixed
from .plonk.serialize import serialize_cs

_M64 = (1 << 64) - 1


def ints_to_canonical(col) -> np.ndarray:
    """list of canonical ints -> (n,4) uint64 little-endian limbs (NOT Montgomery)."""
    n = len(col)
    out = np.zeros((n, 4), dtype=np.uint64)
    try:
        out[:, 0] = np.array(col, dtype=np.uint64)
        return out
    except OverflowError:
        pass
    o = np.array(col, dtype=object)
    for j in range(4):
        out[:, j] = ((o >> (64 * j)) & _M64).astype(np.uint64)
    return out


class ParamsKZG:
    """ParamsKZG<Bn256>: g = [s^i]G, g_lagrange = [L_i(s)]G as (n,8) uint64 affine limbs.  The SRS is
    an input (ParamsKZG::new / read in the reference, src/main.rs:232, src/io.rs:139-146)."""

    def __init__(self, k: int, g: np.ndarray, g_lagrange: np.ndarray):
        self.k, self.n = k, 1 << k
        self.g = np.ascontiguousarray(g, dtype=np.uint64)
        self.g_lagrange = np.ascontiguousarray(g_lagrange, dtype=np.uint64)
        assert self.g.shape == (self.n, 8) and self.g_lagrange.shape == (self.n, 8)
        self._loaded_on = None

    def load(self, ctx: zl.Context):
        if self._loaded_on is not ctx:
            ctx.srs_load(self.k, self.g, self.g_lagrange)
            self._loaded_on = ctx


def vk_transcript_repr(k, cs, fixed_commitments, perm_commitments) -> int:
    """Stand-in for VerifyingKey::transcript_repr (the real value hashes the Rust Debug rendering of the
    vk and cannot be restated without the crate): Blake2b-512, personal
    "Halo2-Verify-Key", over a canonical serialisation; reduced like from_uniform_bytes."""
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    h.update(("k=%d;adv=%d;fix=%d;inst=%d;deg=%d;perm=%d;lookups=%d;gates=%d" % (
        k, cs.num_advice, cs.num_fixed, cs.num_instance, cs.degree(), len(cs.permutation), len(cs.lookups),
        sum(len(g.polys) for g in cs.gates))).encode())
    for p in list(fixed_commitments) + list(perm_commitments):
        h.update(b"\0" * 64 if p is None else p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little"))
    return int.from_bytes(h.digest(), "little") % R_MOD


class ProvingKey:
    """ProvingKey<G1Affine> resident on one GPU (zg_pk handle) + the VerifyingKey data."""

    def __init__(self, ctx, handle, k, cs, fixed_commitments, perm_commitments, transcript_repr):
        self.ctx, self._h, self.k, self.cs = ctx, handle, k, cs
        self.fixed_commitments, self.perm_commitments = fixed_commitments, perm_commitments
        self.transcript_repr = transcript_repr

    def get_vk(self):
        return self

    def stage_ms(self):
        out = (ctypes.c_float * 8)()
        self.ctx._L.zg_pk_last_stage_ms(self._h, out)
        names = ["advice", "lookups", "products", "quotient", "h_commit", "evals", "gwc", "total"]
        return dict(zip(names, [float(x) for x in out]))

    def close(self):
        if self._h:
            self.ctx._L.zg_pk_free(self.ctx._h, self._h)
            self._h = None


def _affine_points(limbs) -> list:
    v = from_limbs(np.asarray(limbs).reshape(-1, 4), Q_MOD)
    return [None if (v[i] == 0 and v[i + 1] == 0) else (v[i], v[i + 1]) for i in range(0, len(v), 2)]


def _load_pk(ctx, params, cs, words, constants, fixed_mont, mapping, repr_int):
    desc = zl.PkDesc()
    desc.k = params.k
    desc.cs_words, desc.cs_nwords = words.ctypes.data, words.shape[0]
    desc.constants, desc.n_constants = constants.ctypes.data, constants.shape[0]
    ptrs = (ctypes.c_void_p * len(fixed_mont))(*[a.ctypes.data for a in fixed_mont])
    desc.fixed = ctypes.cast(ptrs, ctypes.c_void_p)
    desc.perm_mapping = mapping.ctypes.data
    for j in range(4):
        desc.transcript_repr[j] = int(to_limbs([repr_int])[0][j])
    h = ctypes.c_void_p()
    ctx._ck(ctx._L.zg_pk_load(ctx._h, ctypes.byref(desc), ctypes.byref(h)))
    return h


def keygen(ctx: zl.Context, params: ParamsKZG, cs, asm, transcript_repr: int | None = None) -> ProvingKey:
    """keygen_vk + keygen_pk from a synthesized Assembly (selectors are compressed here)."""
    params.load(ctx)
    fixed_int = finalize_fixed(cs, asm)
    words, constants = serialize_cs(cs)
    fixed_mont = [ctx.debug_field_op(0, 7, ints_to_canonical(c)) for c in fixed_int]
    mapping = np.array(asm.mapping, dtype=np.uint32).reshape(len(asm.perm_cols), asm.n, 2)
    mapping = np.ascontiguousarray(mapping)
    # the vk hash depends on the commitments: load once to obtain them, then set transcript_repr
    h = _load_pk(ctx, params, cs, words, constants, fixed_mont, mapping, 0)
    fc = np.zeros((max(cs.num_fixed, 1), 8), dtype=np.uint64)
    pc = np.zeros((max(len(asm.perm_cols), 1), 8), dtype=np.uint64)
    ctx._ck(ctx._L.zg_pk_commitments(ctx._h, h, fc.ctypes.data, pc.ctypes.data))
    fixed_c = _affine_points(fc[:cs.num_fixed])
    perm_c = _affine_points(pc[:len(asm.perm_cols)])
    if transcript_repr is None:
        transcript_repr = vk_transcript_repr(params.k, cs, fixed_c, perm_c)
    repr_limbs = np.ascontiguousarray(to_limbs([transcript_repr])[0], dtype=np.uint64)
    ctx._ck(ctx._L.zg_pk_set_transcript_repr(ctx._h, h, repr_limbs.ctypes.data))
    return ProvingKey(ctx, h, params.k, cs, fixed_c, perm_c, transcript_repr)


def advice_to_mont(ctx: zl.Context, advice_int) -> list:
    """host advice columns (canonical ints) -> Montgomery limb arrays, the form the C ABI takes."""
    return [ctx.debug_field_op(0, 7, ints_to_canonical(c)) for c in advice_int]


def create_proof_limbs(pk: ProvingKey, advice_mont, instance_mont, rng: zl.XorShift) -> bytes:
    """zg_create_proof on already-marshalled buffers (what bench.py times)."""
    ctx = pk.ctx
    aptr = (ctypes.c_void_p * len(advice_mont))(*[a.ctypes.data for a in advice_mont])
    iptr = (ctypes.c_void_p * max(len(instance_mont), 1))(*[a.ctypes.data for a in instance_mont])
    ilen = (ctypes.c_size_t * max(len(instance_mont), 1))(*[a.shape[0] for a in instance_mont])
    cap = 1 << 16
    buf = (ctypes.c_uint8 * cap)()
    plen = ctypes.c_size_t()
    fill = ctypes.cast(ctx._L.zg_xorshift_fill, ctypes.c_void_p)
    ctx._ck(ctx._L.zg_create_proof(ctx._h, pk._h, aptr, iptr, ilen, fill, ctypes.byref(rng), buf, cap, ctypes.byref(plen)))
    return bytes(buf[:plen.value])


def create_proof(params: ParamsKZG, pk: ProvingKey, advice_int, instances, rng: zl.XorShift) -> bytes:
    """create_proof(params, pk, &[circuit], &[&[instances]], rng, transcript) -> proof bytes."""
    ctx = pk.ctx
    params.load(ctx)
    adv = advice_to_mont(ctx, advice_int)
    inst = [to_limbs(list(v)) for v in instances]
    return create_proof_limbs(pk, adv, inst, rng)
