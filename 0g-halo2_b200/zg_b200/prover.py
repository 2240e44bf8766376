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

    def __init__(self, k: int, g: np.ndarray, g_lagrange: np.ndarray, g2=None, s_g2=None):
        self.k, self.n = k, 1 << k
        self.g = np.ascontiguousarray(g, dtype=np.uint64)
        self.g_lagrange = np.ascontiguousarray(g_lagrange, dtype=np.uint64)
        assert self.g.shape == (self.n, 8) and self.g_lagrange.shape == (self.n, 8)
        # ParamsKZG::{g2, s_g2} (G2Affine, 16 uint64 Montgomery limbs x.c0 | x.c1 | y.c0 | y.c1): only the verifier needs them
        self.g2 = None if g2 is None else np.ascontiguousarray(np.frombuffer(g2, dtype=np.uint64) if isinstance(g2, bytes) else g2,
                                                               dtype=np.uint64).reshape(16)
        self.s_g2 = None if s_g2 is None else np.ascontiguousarray(
            np.frombuffer(s_g2, dtype=np.uint64) if isinstance(s_g2, bytes) else s_g2, dtype=np.uint64).reshape(16)
        self._loaded_on = None

    def load(self, ctx: zl.Context):
        if self._loaded_on is not ctx:
            ctx.srs_load(self.k, self.g, self.g_lagrange)
            self._loaded_on = ctx

    def share(self, ctx: zl.Context, owner: zl.Context):
        """`ctx` uses the device copy (bases + window tables) `owner` already holds: zg_srs_share"""
        ctx.srs_share(owner)
        self._loaded_on = ctx


_warned_stand_in = False


def vk_transcript_repr(k, cs, fixed_commitments, perm_commitments) -> int:
    """Stand-in for VerifyingKey::transcript_repr (the real value hashes the Rust Debug rendering of the
    vk and cannot be restated without the crate): Blake2b-512, personal
    "Halo2-Verify-Key", over a canonical serialisation; reduced like from_uniform_bytes.
    Proofs made under it verify with this repository's verifiers only: pass the real value (`transcript_repr=` of keygen /
    load_proving_key, taken from a Rust-built key) for proofs the Rust or EVM verifier must accept.  Warns once."""
    global _warned_stand_in
    if not _warned_stand_in:
        import warnings
        warnings.warn("zg_b200: using the stand-in vk.transcript_repr; proofs will not verify under the Rust / EVM verifier "
                      "until the real value is supplied (keygen(..., transcript_repr=...))", stacklevel=3)
        _warned_stand_in = True
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Verify-Key")
    h.update(("k=%d;adv=%d;fix=%d;inst=%d;deg=%d;perm=%d;lookups=%d;gates=%d" % (
        k, cs.num_advice, cs.num_fixed, cs.num_instance, cs.degree(), len(cs.permutation), len(cs.lookups),
        sum(len(g.polys) for g in cs.gates))).encode())
    for p in list(fixed_commitments) + list(perm_commitments):
        h.update(b"\0" * 64 if p is None else p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little"))
    return int.from_bytes(h.digest(), "little") % R_MOD


class VerifyingKey:
    """VerifyingKey<G1Affine>: constraint system + fixed / permutation commitments + transcript_repr, and the host-side
    verifier bound to them (zg_vk, csrc/verifier.cu).  Needs no GPU and no context."""

    def __init__(self, k, cs_words, constants, fixed_limbs, perm_limbs, transcript_repr: int):
        self.k, self.transcript_repr = k, transcript_repr
        self.cs_words = np.ascontiguousarray(cs_words, dtype=np.uint32)
        self.constants = np.ascontiguousarray(constants, dtype=np.uint64).reshape(-1, 4)
        self.fixed_limbs = np.ascontiguousarray(fixed_limbs, dtype=np.uint64).reshape(-1, 8)
        self.perm_limbs = np.ascontiguousarray(perm_limbs, dtype=np.uint64).reshape(-1, 8)
        self._L = zl.load_library()
        rl = np.ascontiguousarray(to_limbs([transcript_repr])[0], dtype=np.uint64)
        h = ctypes.c_void_p()
        rc = self._L.zg_vk_create(k, self.cs_words.ctypes.data, self.cs_words.shape[0], self.constants.ctypes.data,
                                  self.constants.shape[0], self.fixed_limbs.ctypes.data, self.perm_limbs.ctypes.data,
                                  rl.ctypes.data, ctypes.byref(h))
        if rc != zl.ZG_OK:
            raise zl.ZgError(rc, "zg_vk_create: malformed verifying key")
        self._h = h

    def verify(self, params: "ParamsKZG", instances, proof: bytes) -> bool:
        """verify_proof(params, vk, SingleStrategy, &[&[instances]], transcript).is_ok()"""
        if params.g2 is None or params.s_g2 is None:
            raise ValueError("ParamsKZG without g2 / s_g2 cannot verify")
        inst = [np.ascontiguousarray(to_limbs([int(x) for x in v]), dtype=np.uint64).reshape(-1, 4) for v in instances]
        iptr = (ctypes.c_void_p * max(len(inst), 1))(*[a.ctypes.data for a in inst])
        ilen = (ctypes.c_size_t * max(len(inst), 1))(*[a.shape[0] for a in inst])
        buf = (ctypes.c_uint8 * max(len(proof), 1)).from_buffer_copy(proof or b"\0")
        g0 = np.ascontiguousarray(params.g[0], dtype=np.uint64)
        rc = self._L.zg_verify_proof(self._h, g0.ctypes.data, params.g2.ctypes.data, params.s_g2.ctypes.data, iptr, ilen,
                                     buf, len(proof))
        if rc == zl.ZG_OK:
            return True
        if rc == zl.ZG_E_VERIFY:
            return False
        raise zl.ZgError(rc, self._L.zg_vk_last_error(self._h).decode())

    def last_error(self) -> str:
        return self._L.zg_vk_last_error(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            self._L.zg_vk_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pairing_check(g1_points, g2_points) -> bool:
    """prod e(P_i, Q_i) == 1 on the host (zg_pairing_check): (m,8) and (m,16) uint64 limb arrays"""
    P = np.ascontiguousarray(g1_points, dtype=np.uint64).reshape(-1, 8)
    Q = np.ascontiguousarray(g2_points, dtype=np.uint64).reshape(-1, 16)
    assert P.shape[0] == Q.shape[0]
    out = ctypes.c_int()
    rc = zl.load_library().zg_pairing_check(P.ctypes.data, Q.ctypes.data, P.shape[0], ctypes.byref(out))
    if rc != zl.ZG_OK:
        raise zl.ZgError(rc, "zg_pairing_check: a point is not on its curve")
    return bool(out.value)


class ProvingKey:
    """ProvingKey<G1Affine> resident on one GPU (zg_pk handle) + the VerifyingKey data."""

    def __init__(self, ctx, handle, k, cs, fixed_commitments, perm_commitments, transcript_repr, vk_parts=None):
        self.ctx, self._h, self.k, self.cs = ctx, handle, k, cs
        self.fixed_commitments, self.perm_commitments = fixed_commitments, perm_commitments
        self.transcript_repr = transcript_repr
        self._vk_parts, self._vk = vk_parts, None

    def get_vk(self) -> VerifyingKey:
        """ProvingKey::get_vk (src/wnn.rs:271)"""
        if self._vk is None:
            words, constants, fc, pc = self._vk_parts
            self._vk = VerifyingKey(self.k, words, constants, fc, pc, self.transcript_repr)
        return self._vk

    def clone(self, ctx: "zl.Context") -> "ProvingKey":
        """zg_pk_clone: a key for another context of the same device that shares this key's resident columns and owns only
        its per-proof workspace (`ctx` must hold an SRS of the same k, e.g. through `ctx.srs_share`)."""
        h = ctypes.c_void_p()
        ctx._ck(ctx._L.zg_pk_clone(ctx._h, self._h, ctypes.byref(h)))
        return ProvingKey(ctx, h, self.k, self.cs, self.fixed_commitments, self.perm_commitments, self.transcript_repr,
                          vk_parts=self._vk_parts)

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
    return ProvingKey(ctx, h, params.k, cs, fixed_c, perm_c, transcript_repr,
                      vk_parts=(words, constants, fc[:cs.num_fixed].copy(), pc[:len(asm.perm_cols)].copy()))


def export_proving_key(pk: ProvingKey, selectors, f) -> None:
    """`write_keys` for the proving key (src/io.rs:159-163): ProvingKey::write(RawBytes) to the binary file object `f`.
    `selectors` are the activations before compression (Assembly.selectors).  The key's columns come back from the
    device (zg_pk_read_column); the extended forms are rebuilt on halo2's zeta coset through the C ABI."""
    from . import io as zio
    ctx, k, cs = pk.ctx, pk.k, pk.cs
    n = 1 << k
    ext_k = k
    while (1 << ext_k) < n * (cs.degree() - 1):
        ext_k += 1

    def col(what, i):
        out = np.zeros((n, 4), dtype=np.uint64)
        ctx._ck(ctx._L.zg_pk_read_column(ctx._h, pk._h, what, i, out.ctypes.data))
        return out
    nf, m = cs.num_fixed, len(cs.permutation)
    fixed_values = [col(0, i) for i in range(nf)]
    fixed_polys = [col(1, i) for i in range(nf)]
    perm_values = [col(2, i) for i in range(m)]
    perm_polys = [col(3, i) for i in range(m)]
    ext = lambda c: ctx.coeff_to_extended(c, k, ext_k)
    bf = cs.blinding_factors()
    one = to_limbs([1])[0]

    def unit_rows(rows):
        v = np.zeros((n, 4), dtype=np.uint64)
        v[list(rows)] = one
        return ext(ctx.lagrange_to_coeff(v, k))
    l0, l_last, l_blind = unit_rows([0]), unit_rows([n - bf - 1]), unit_rows(range(n - bf, n))
    ones = np.tile(one, (1 << ext_k, 1))
    l_active = ctx.debug_field_op(0, 4, ctx.debug_field_op(0, 4, ones, l_last), l_blind)      # 1 - l_last - l_blind (op 4 = sub)
    words, constants, fc, pc = pk._vk_parts
    vk = {"k": k, "fixed_commitments": fc, "perm_commitments": pc, "selectors": selectors}
    zio.write_pk(f, vk, l0, l_last, l_active, fixed_values, fixed_polys, [ext(c) for c in fixed_polys], perm_values,
                 perm_polys, [ext(c) for c in perm_polys])


def export_verifying_key(pk: ProvingKey, selectors, f) -> None:
    """`pk.get_vk().write(writer, RawBytes)` (src/io.rs:162)."""
    from . import io as zio
    _, _, fc, pc = pk._vk_parts
    zio.write_vk(f, pk.k, fc, pc, selectors)


def load_proving_key(ctx: zl.Context, params: ParamsKZG, cs, f, transcript_repr: int | None = None) -> ProvingKey:
    """`read_pk` (src/io.rs:166-170): ProvingKey::read(RawBytes, circuit params) -> a key resident on the GPU.
    `cs` is the circuit's constraint system as `configure` leaves it (BEFORE selector compression, exactly what the
    Rust reader rebuilds from the circuit params); the file supplies the selector activations, the fixed columns and
    the permutation columns.  Coefficient and extended forms are recomputed on the device (its extended domain is
    internal), and the commitments it derives are checked against the file's."""
    from . import io as zio
    d = zio.read_pk(f, len(cs.permutation), len(cs.selectors))
    if d["k"] != params.k:
        raise ValueError("key file is for k = %d, parameters for k = %d" % (d["k"], params.k))
    nf0 = cs.num_fixed
    cs.compress_selectors(d["selectors"])                 # substitutes the selectors; the columns come from the file
    if cs.num_fixed != d["fixed_commitments"].shape[0]:
        raise ValueError("key file does not belong to this circuit (fixed column count)")
    params.load(ctx)
    words, constants = serialize_cs(cs)
    fixed = [np.ascontiguousarray(c) for c in d["fixed_values"]]
    sig = [np.ascontiguousarray(c) for c in d["perm_values"]]
    desc = zl.PkDesc()
    desc.k = params.k
    desc.cs_words, desc.cs_nwords = words.ctypes.data, words.shape[0]
    desc.constants, desc.n_constants = constants.ctypes.data, constants.shape[0]
    fptr = (ctypes.c_void_p * max(len(fixed), 1))(*[a.ctypes.data for a in fixed])
    sptr = (ctypes.c_void_p * max(len(sig), 1))(*[a.ctypes.data for a in sig])
    desc.fixed = ctypes.cast(fptr, ctypes.c_void_p)
    desc.perm_mapping = None
    desc.sigma_values = ctypes.cast(sptr, ctypes.c_void_p)
    h = ctypes.c_void_p()
    ctx._ck(ctx._L.zg_pk_load(ctx._h, ctypes.byref(desc), ctypes.byref(h)))
    fc = np.zeros((max(cs.num_fixed, 1), 8), dtype=np.uint64)
    pc = np.zeros((max(len(sig), 1), 8), dtype=np.uint64)
    ctx._ck(ctx._L.zg_pk_commitments(ctx._h, h, fc.ctypes.data, pc.ctypes.data))
    if not ((fc[:cs.num_fixed] == d["fixed_commitments"]).all() and (pc[:len(sig)] == d["perm_commitments"]).all()):
        ctx._L.zg_pk_free(ctx._h, h)
        raise ValueError("key file: commitments do not match its columns under these parameters")
    fixed_c, perm_c = _affine_points(fc[:cs.num_fixed]), _affine_points(pc[:len(sig)])
    if transcript_repr is None:
        transcript_repr = vk_transcript_repr(params.k, cs, fixed_c, perm_c)
    repr_limbs = np.ascontiguousarray(to_limbs([transcript_repr])[0], dtype=np.uint64)
    ctx._ck(ctx._L.zg_pk_set_transcript_repr(ctx._h, h, repr_limbs.ctypes.data))
    return ProvingKey(ctx, h, params.k, cs, fixed_c, perm_c, transcript_repr,
                      vk_parts=(words, constants, fc[:cs.num_fixed].copy(), pc[:len(sig)].copy()))


def load_verifying_key(cs, f, transcript_repr: int | None = None) -> VerifyingKey:
    """`read_vk` (src/io.rs:173-176): VerifyingKey::read(RawBytes, circuit params).  Host only."""
    from . import io as zio
    d = zio.read_vk(f, len(cs.permutation), len(cs.selectors))
    if f.read(1):
        raise ValueError("key file: trailing bytes")
    cs.compress_selectors(d["selectors"])
    words, constants = serialize_cs(cs)
    fixed_c, perm_c = _affine_points(d["fixed_commitments"]), _affine_points(d["perm_commitments"])
    if transcript_repr is None:
        transcript_repr = vk_transcript_repr(d["k"], cs, fixed_c, perm_c)
    return VerifyingKey(d["k"], words, constants, d["fixed_commitments"], d["perm_commitments"], transcript_repr)


def advice_to_mont(ctx: zl.Context, advice_int) -> list:
    """host advice columns (canonical ints) -> Montgomery limb arrays, the form the C ABI takes."""
    return [ctx.debug_field_op(0, 7, ints_to_canonical(c)) for c in advice_int]


def create_proof_limbs(pk: ProvingKey, advice_mont, instance_mont, rng=None) -> bytes:
    """zg_create_proof on already-marshalled buffers (what bench.py times).  `rng` is any ctypes object whose class names
    its bulk-draw callback in `fill_name` (lib.ChaCha20Rng, lib.XorShift) or a (zg_rng_fill_fn pointer, state) pair;
    None = a fresh OS-seeded ChaCha20 stream per proof, the counterpart of the reference's OsRng (src/wnn.rs:256).
    XorShift is for seeded, reproducible tests and benches only."""
    if rng is None:
        rng = zl.ChaCha20Rng.from_os()
    ctx = pk.ctx
    aptr = (ctypes.c_void_p * len(advice_mont))(*[a.ctypes.data for a in advice_mont])
    iptr = (ctypes.c_void_p * max(len(instance_mont), 1))(*[a.ctypes.data for a in instance_mont])
    ilen = (ctypes.c_size_t * max(len(instance_mont), 1))(*[a.shape[0] for a in instance_mont])
    cap = 1 << 16
    buf = (ctypes.c_uint8 * cap)()
    plen = ctypes.c_size_t()
    if isinstance(rng, tuple):
        fill, state = ctypes.cast(rng[0], ctypes.c_void_p), rng[1]
    else:
        fill, state = ctypes.cast(getattr(ctx._L, rng.fill_name), ctypes.c_void_p), ctypes.byref(rng)
    ctx._ck(ctx._L.zg_create_proof(ctx._h, pk._h, aptr, iptr, ilen, fill, state, buf, cap, ctypes.byref(plen)))
    return bytes(buf[:plen.value])


def create_proof(params: ParamsKZG, pk: ProvingKey, advice_int, instances, rng=None) -> bytes:
    """create_proof(params, pk, &[circuit], &[&[instances]], rng, transcript) -> proof bytes."""
    ctx = pk.ctx
    params.load(ctx)
    adv = advice_to_mont(ctx, advice_int)
    inst = [to_limbs(list(v)) for v in instances]
    return create_proof_limbs(pk, adv, inst, rng)
