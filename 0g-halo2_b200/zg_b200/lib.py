"""ctypes binding of libzg_b200.so (the C ABI declared in include/zg_b200.h).

The same entry points a patched halo2_proofs would bind from Rust.  There is
no CPU fallback: if the shared library or a CUDA device is missing, construction raises.

Array conventions (numpy, dtype uint64):
  Fr / Fq      -> (..., 4)   little-endian limbs, Montgomery form (halo2curves in-memory layout)
  G1Affine     -> (..., 8)   x | y ; identity = all zero
  G1 (Jacobian)-> (..., 12)  x | y | z ; identity z = 0
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzg_b200.so")

ZG_OK = 0
ZG_E_VERIFY = -6
BASIS_MONOMIAL = 0
BASIS_LAGRANGE = 1

EXPORTS = [
    "zg_version", "zg_ctx_create", "zg_ctx_destroy", "zg_last_error", "zg_sync", "zg_launch_count",
    "zg_dev_alloc", "zg_dev_free", "zg_h2d", "zg_d2h",
    "zg_srs_load", "zg_srs_share", "zg_msm", "zg_msm_batch", "zg_msm_dev",
    "zg_ntt", "zg_ntt_dev", "zg_lagrange_to_coeff", "zg_lagrange_to_coeff_dev",
    "zg_coeff_to_extended", "zg_coeff_to_extended_dev", "zg_extended_to_coeff", "zg_extended_to_coeff_dev",
    "zg_bench_int_pipe", "zg_debug_field_op", "zg_debug_keccak256", "zg_probe_enable", "zg_probe_read",
    "zg_xorshift_seed", "zg_xorshift_fill", "zg_chacha20_seed_os", "zg_chacha20_seed", "zg_chacha20_fill", "zg_pk_load", "zg_pk_clone", "zg_pk_read_column", "zg_pk_free", "zg_pk_commitments", "zg_create_proof",
    "zg_pk_last_stage_ms", "zg_pk_set_transcript_repr",
    "zg_wnn_create", "zg_wnn_free", "zg_wnn_last_error", "zg_wnn_synthesize",
    "zg_comm_unique_id", "zg_comm_init", "zg_comm_destroy", "zg_ctx_set_distribution", "zg_msm_sharded_dev", "zg_msm_sharded",
    "zg_vk_create", "zg_vk_free", "zg_vk_last_error", "zg_verify_proof", "zg_pairing_check",
    "zg_lookup_permute", "zg_grand_product", "zg_batch_invert", "zg_eval_poly_batch", "zg_kate_division", "zg_evaluate_h",
]


class ZgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("zg_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library() -> ctypes.CDLL:
    """Loads libzg_b200.so; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libzg_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C 0g-halo2_b200/csrc`")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64, sz, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int
    L.zg_version.restype = ctypes.c_char_p
    L.zg_last_error.restype = ctypes.c_char_p
    L.zg_last_error.argtypes = [vp]
    L.zg_launch_count.restype = u64
    L.zg_launch_count.argtypes = [vp]
    L.zg_ctx_create.argtypes = [ci, vp, ctypes.POINTER(vp)]
    L.zg_ctx_destroy.argtypes = [vp]
    L.zg_ctx_destroy.restype = None
    L.zg_sync.argtypes = [vp]
    L.zg_dev_alloc.argtypes = [vp, sz, ctypes.POINTER(vp)]
    L.zg_dev_free.argtypes = [vp, vp]
    L.zg_h2d.argtypes = [vp, vp, vp, sz]
    L.zg_d2h.argtypes = [vp, vp, vp, sz]
    L.zg_srs_load.argtypes = [vp, u32, vp, vp]
    L.zg_srs_share.argtypes = [vp, vp]
    L.zg_pk_clone.argtypes = [vp, vp, ctypes.POINTER(vp)]
    L.zg_msm.argtypes = [vp, ci, vp, sz, vp]
    L.zg_msm_batch.argtypes = [vp, ci, vp, sz, sz, vp]
    L.zg_msm_dev.argtypes = [vp, ci, vp, sz, sz, sz, vp]
    L.zg_ntt.argtypes = [vp, vp, u32, vp]
    L.zg_ntt_dev.argtypes = [vp, vp, vp, u32, vp, sz, sz]
    L.zg_lagrange_to_coeff.argtypes = [vp, vp, u32]
    L.zg_lagrange_to_coeff_dev.argtypes = [vp, vp, vp, u32, sz, sz]
    L.zg_coeff_to_extended.argtypes = [vp, vp, u32, u32, vp]
    L.zg_coeff_to_extended_dev.argtypes = [vp, vp, sz, u32, u32, vp, sz, sz]
    L.zg_extended_to_coeff.argtypes = [vp, vp, u32, u32, sz, vp]
    L.zg_extended_to_coeff_dev.argtypes = [vp, vp, u32, u32, sz, vp]
    L.zg_bench_int_pipe.argtypes = [vp, ci, u32, ctypes.POINTER(ctypes.c_double)]
    L.zg_debug_field_op.argtypes = [vp, ci, ci, vp, vp, vp, sz]
    L.zg_debug_keccak256.argtypes = [vp, sz, vp]
    L.zg_debug_keccak256.restype = None
    L.zg_probe_enable.argtypes = [vp, ci]
    L.zg_probe_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.zg_xorshift_seed.argtypes = [vp, vp]
    L.zg_xorshift_seed.restype = None
    L.zg_xorshift_fill.argtypes = [vp, vp, sz]
    L.zg_xorshift_fill.restype = None
    L.zg_chacha20_seed_os.argtypes = [vp]
    L.zg_chacha20_seed.argtypes = [vp, vp]
    L.zg_chacha20_seed.restype = None
    L.zg_chacha20_fill.argtypes = [vp, vp, sz]
    L.zg_chacha20_fill.restype = None
    L.zg_pk_load.argtypes = [vp, vp, ctypes.POINTER(vp)]
    L.zg_pk_read_column.argtypes = [vp, vp, ci, u32, vp]
    L.zg_pk_free.argtypes = [vp, vp]
    L.zg_pk_free.restype = None
    L.zg_pk_commitments.argtypes = [vp, vp, vp, vp]
    L.zg_pk_set_transcript_repr.argtypes = [vp, vp, vp]
    L.zg_create_proof.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, sz, ctypes.POINTER(sz)]
    L.zg_pk_last_stage_ms.argtypes = [vp, vp]
    L.zg_wnn_create.argtypes = [vp, ctypes.POINTER(vp)]
    L.zg_wnn_free.argtypes = [vp]
    L.zg_wnn_free.restype = None
    L.zg_wnn_last_error.argtypes = [vp]
    L.zg_wnn_last_error.restype = ctypes.c_char_p
    L.zg_wnn_synthesize.argtypes = [vp, vp, u32, u32, vp, vp]
    L.zg_comm_unique_id.argtypes = [vp]
    L.zg_comm_init.argtypes = [vp, ci, ci, vp]
    L.zg_comm_destroy.argtypes = [vp]
    L.zg_ctx_set_distribution.argtypes = [vp, ci]
    L.zg_msm_sharded_dev.argtypes = [vp, ci, vp, sz, sz, sz, vp]
    L.zg_msm_sharded.argtypes = [vp, ci, vp, sz, vp]
    L.zg_vk_create.argtypes = [u32, vp, sz, vp, sz, vp, vp, vp, ctypes.POINTER(vp)]
    L.zg_vk_free.argtypes = [vp]
    L.zg_vk_free.restype = None
    L.zg_vk_last_error.argtypes = [vp]
    L.zg_vk_last_error.restype = ctypes.c_char_p
    L.zg_verify_proof.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz]
    L.zg_pairing_check.argtypes = [vp, vp, sz, ctypes.POINTER(ci)]
    L.zg_lookup_permute.argtypes = [vp, vp, vp, sz, vp, vp]
    L.zg_grand_product.argtypes = [vp, vp, vp, sz, vp]
    L.zg_batch_invert.argtypes = [vp, vp, sz]
    L.zg_eval_poly_batch.argtypes = [vp, vp, sz, sz, vp, vp]
    L.zg_kate_division.argtypes = [vp, vp, sz, vp, vp]
    L.zg_evaluate_h.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, ci, vp]
    _lib = L
    return L


def _np(a, last):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] != last:
        raise ValueError("expected trailing dimension %d, got shape %s" % (last, a.shape))
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class Context:
    """One GPU, one stream.  `stream` may be a raw cudaStream_t handle (int), e.g.
    torch.cuda.current_stream().cuda_stream, so that torch CUDA events time this context's work."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._L = load_library()
        h = ctypes.c_void_p()
        rc = self._L.zg_ctx_create(device, ctypes.c_void_p(stream) if stream else None, ctypes.byref(h))
        if rc != ZG_OK:
            raise ZgError(rc, "zg_ctx_create failed (no CUDA device %d? there is no CPU fallback)" % device)
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._L.zg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != ZG_OK:
            raise ZgError(rc, self._L.zg_last_error(self._h).decode())

    def sync(self):
        self._ck(self._L.zg_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.zg_launch_count(self._h))

    # ---- SRS / MSM ----
    def srs_load(self, k: int, g=None, g_lagrange=None):
        n = 1 << k
        g = None if g is None else _np(g, 8)
        gl = None if g_lagrange is None else _np(g_lagrange, 8)
        for a in (g, gl):
            if a is not None and a.shape[0] != n:
                raise ValueError("SRS basis must have 2^k points")
        self._ck(self._L.zg_srs_load(self._h, k, None if g is None else _ptr(g), None if gl is None else _ptr(gl)))

    def srs_share(self, other: "Context"):
        """use the SRS (bases + window tables) `other` has loaded on the same device -- no second copy"""
        self._ck(self._L.zg_srs_share(self._h, other._h))

    def msm(self, basis: int, scalars) -> np.ndarray:
        s = _np(scalars, 4)
        out = np.zeros(12, dtype=np.uint64)
        self._ck(self._L.zg_msm(self._h, basis, _ptr(s), s.shape[0], _ptr(out)))
        return out

    def msm_batch(self, basis: int, scalar_list) -> np.ndarray:
        arrs = [_np(s, 4) for s in scalar_list]
        n = arrs[0].shape[0]
        assert all(a.shape[0] == n for a in arrs)
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        out = np.zeros((len(arrs), 12), dtype=np.uint64)
        self._ck(self._L.zg_msm_batch(self._h, basis, ptrs, n, len(arrs), _ptr(out)))
        return out

    def msm_dev(self, basis: int, scalars_ptr: int, stride: int, n: int, count: int, out_ptr: int):
        self._ck(self._L.zg_msm_dev(self._h, basis, scalars_ptr, stride, n, count, out_ptr))

    # ---- multi-GPU (NCCL inside the library; dist.cu) ----
    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        """join an NCCL communicator; `unique_id` = the 128 bytes rank 0 got from `comm_unique_id()`"""
        assert len(unique_id) == 128
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._L.zg_comm_init(self._h, nranks, rank, buf))

    def comm_destroy(self):
        self._ck(self._L.zg_comm_destroy(self._h))

    def set_distribution(self, mode: int):
        """0 = none, 1 = spread every round's commitments of zg_create_proof over the ranks (SPMD)"""
        self._ck(self._L.zg_ctx_set_distribution(self._h, mode))

    def msm_sharded(self, basis: int, scalars_slice) -> np.ndarray:
        s = _np(scalars_slice, 4)
        out = np.zeros(12, dtype=np.uint64)
        self._ck(self._L.zg_msm_sharded(self._h, basis, _ptr(s), s.shape[0], _ptr(out)))
        return out

    def msm_sharded_dev(self, basis: int, scalars_ptr: int, stride: int, n_local: int, count: int, out_ptr: int):
        self._ck(self._L.zg_msm_sharded_dev(self._h, basis, scalars_ptr, stride, n_local, count, out_ptr))

    # ---- NTT family ----
    def ntt(self, a, log_n: int, omega) -> np.ndarray:
        a = _np(a, 4).copy()
        w = _np(omega, 4)
        self._ck(self._L.zg_ntt(self._h, _ptr(a), log_n, _ptr(w)))
        return a

    def ntt_inplace(self, a: np.ndarray, log_n: int, omega):
        """zg_ntt on the caller's buffer (no copy; pass pinned memory for full PCIe speed)"""
        assert a.dtype == np.uint64 and a.flags.c_contiguous and a.shape[-1] == 4
        w = _np(omega, 4)
        self._ck(self._L.zg_ntt(self._h, _ptr(a), log_n, _ptr(w)))
        return a

    def ntt_dev(self, in_ptr: int, out_ptr: int, log_n: int, omega, batch: int = 1, stride: int | None = None):
        w = _np(omega, 4)
        self._ck(self._L.zg_ntt_dev(self._h, in_ptr, out_ptr, log_n, _ptr(w), batch, stride or (1 << log_n)))

    def lagrange_to_coeff(self, a, k: int) -> np.ndarray:
        a = _np(a, 4).copy()
        self._ck(self._L.zg_lagrange_to_coeff(self._h, _ptr(a), k))
        return a

    def lagrange_to_coeff_dev(self, in_ptr, out_ptr, k, batch=1, stride=None):
        self._ck(self._L.zg_lagrange_to_coeff_dev(self._h, in_ptr, out_ptr, k, batch, stride or (1 << k)))

    def coeff_to_extended(self, coeff, k: int, ext_k: int) -> np.ndarray:
        c = _np(coeff, 4)
        out = np.zeros((1 << ext_k, 4), dtype=np.uint64)
        self._ck(self._L.zg_coeff_to_extended(self._h, _ptr(c), k, ext_k, _ptr(out)))
        return out

    def coeff_to_extended_dev(self, in_ptr, in_stride, k, ext_k, out_ptr, out_stride, batch=1):
        self._ck(self._L.zg_coeff_to_extended_dev(self._h, in_ptr, in_stride, k, ext_k, out_ptr, out_stride, batch))

    def extended_to_coeff(self, ext, k: int, ext_k: int, keep: int) -> np.ndarray:
        e = _np(ext, 4)
        out = np.zeros((keep, 4), dtype=np.uint64)
        self._ck(self._L.zg_extended_to_coeff(self._h, _ptr(e), k, ext_k, keep, _ptr(out)))
        return out

    def extended_to_coeff_dev(self, ext_ptr, k, ext_k, keep, out_ptr):
        self._ck(self._L.zg_extended_to_coeff_dev(self._h, ext_ptr, k, ext_k, keep, out_ptr))

    # ---- single prover stages (host arrays) ----
    def lookup_permute(self, a, s):
        """permute_expression_pair over the rows given (no blinding rows): returns (a', s')."""
        a, s = _np(a, 4), _np(s, 4)
        assert a.shape == s.shape
        pa, ps = np.zeros_like(a), np.zeros_like(s)
        self._ck(self._L.zg_lookup_permute(self._h, _ptr(a), _ptr(s), a.shape[0], _ptr(pa), _ptr(ps)))
        return pa, ps

    def grand_product(self, num, den) -> np.ndarray:
        num, den = _np(num, 4), _np(den, 4)
        assert num.shape == den.shape
        z = np.zeros_like(num)
        self._ck(self._L.zg_grand_product(self._h, _ptr(num), _ptr(den), num.shape[0], _ptr(z)))
        return z

    def batch_invert(self, a) -> np.ndarray:
        a = _np(a, 4).copy()
        self._ck(self._L.zg_batch_invert(self._h, _ptr(a), a.shape[0]))
        return a

    def eval_poly_batch(self, polys, x) -> np.ndarray:
        arrs = [_np(p, 4) for p in polys]
        n = arrs[0].shape[0]
        assert all(a.shape[0] == n for a in arrs)
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        xx = _np(x, 4).reshape(4)
        out = np.zeros((len(arrs), 4), dtype=np.uint64)
        self._ck(self._L.zg_eval_poly_batch(self._h, ptrs, n, len(arrs), _ptr(xx), _ptr(out)))
        return out

    def kate_division(self, a, z) -> np.ndarray:
        a = _np(a, 4)
        zz = _np(z, 4).reshape(4)
        q = np.zeros((max(a.shape[0] - 1, 0), 4), dtype=np.uint64)
        self._ck(self._L.zg_kate_division(self._h, _ptr(a), a.shape[0], _ptr(zz), _ptr(q)))
        return q

    def evaluate_h(self, pk_handle, ext_n, advice, instance, lk_in, lk_tab, lk_z, perm_z, challenges, divide=False):
        """Evaluator::evaluate_h; polynomial lists are coefficient-form (n,4) arrays, challenges = (theta, beta, gamma, y)
        as a (4,4) limb array."""
        def plist(lst):
            arrs = [_np(p, 4) for p in lst]
            return arrs, (ctypes.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
        keep = [plist(x) for x in (advice, instance, lk_in, lk_tab, lk_z, perm_z)]
        ch = _np(challenges, 4)
        assert ch.shape == (4, 4)
        out = np.zeros((ext_n, 4), dtype=np.uint64)
        self._ck(self._L.zg_evaluate_h(self._h, pk_handle, *[k[1] for k in keep], _ptr(ch), int(divide), _ptr(out)))
        return out

    # ---- micro-benchmarks ----
    def bench_int_pipe(self, kind: int, iters: int = 4096) -> float:
        v = ctypes.c_double()
        self._ck(self._L.zg_bench_int_pipe(self._h, kind, iters, ctypes.byref(v)))
        return v.value

    def probe_enable(self, on: bool = True):
        self._ck(self._L.zg_probe_enable(self._h, int(on)))

    def probe_read(self):
        """(kernel ms, launches, point additions) of msm_accumulate_kernel since probe_enable / the last read"""
        ms, ln, adds = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_uint64()
        self._ck(self._L.zg_probe_read(self._h, ctypes.byref(ms), ctypes.byref(ln), ctypes.byref(adds)))
        return ms.value, int(ln.value), int(adds.value)

    def debug_field_op(self, field: int, op: int, a, b=None) -> np.ndarray:
        a = _np(a, 4)
        b = a if b is None else _np(b, 4)
        out = np.zeros_like(a)
        self._ck(self._L.zg_debug_field_op(self._h, field, op, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it, the host plumbing broadcasts the 128 bytes)"""
    buf = (ctypes.c_uint8 * 128)()
    rc = load_library().zg_comm_unique_id(buf)
    if rc != ZG_OK:
        raise ZgError(rc, "zg_comm_unique_id: NCCL is not available")
    return bytes(buf)


class PkDesc(ctypes.Structure):
    """zg_pk_desc (include/zg_b200.h)."""
    _fields_ = [("k", ctypes.c_uint32), ("cs_words", ctypes.c_void_p), ("cs_nwords", ctypes.c_size_t),
                ("constants", ctypes.c_void_p), ("n_constants", ctypes.c_size_t), ("fixed", ctypes.c_void_p),
                ("perm_mapping", ctypes.c_void_p), ("transcript_repr", ctypes.c_uint64 * 4),
                ("sigma_values", ctypes.c_void_p)]


class XorShift(ctypes.Structure):
    """zg_xorshift: rand_xorshift::XorShiftRng state; `fill_fn` is the zg_rng_fill_fn to pass with it."""
    _fields_ = [("x", ctypes.c_uint32), ("y", ctypes.c_uint32), ("z", ctypes.c_uint32), ("w", ctypes.c_uint32)]

    @classmethod
    def from_seed(cls, seed: bytes):
        assert len(seed) == 16
        r = cls()
        load_library().zg_xorshift_seed(ctypes.byref(r), seed)
        return r


XorShift.fill_name = "zg_xorshift_fill"


class ChaCha20Rng(ctypes.Structure):
    """zg_chacha20: the production RNG (ChaCha20 keystream keyed from the OS entropy source, the role OsRng plays at
    src/wnn.rs:256).  `from_os()` for proofs, `from_key(32 bytes)` for known-answer tests."""
    _fields_ = [("key", ctypes.c_uint32 * 8), ("counter", ctypes.c_uint64), ("nonce", ctypes.c_uint32 * 2),
                ("have", ctypes.c_uint32), ("buf", ctypes.c_uint8 * 64)]
    fill_name = "zg_chacha20_fill"

    @classmethod
    def from_os(cls):
        r = cls()
        if load_library().zg_chacha20_seed_os(ctypes.byref(r)) != ZG_OK:
            raise ZgError(-1, "the operating system returned no entropy (getrandom)")
        return r

    @classmethod
    def from_key(cls, key: bytes):
        assert len(key) == 32
        r = cls()
        load_library().zg_chacha20_seed(ctypes.byref(r), key)
        return r

    def draw(self, n: int) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint64)
        load_library().zg_chacha20_fill(ctypes.byref(self), out.ctypes.data, n)
        return out


class WnnDesc(ctypes.Structure):
    """zg_wnn_desc (include/zg_b200.h)."""
    _fields_ = [("p", ctypes.c_uint64), ("n_hashes", ctypes.c_uint32), ("bits_per_hash", ctypes.c_uint32),
                ("bits_per_filter", ctypes.c_uint32), ("n_classes", ctypes.c_uint32), ("n_filters", ctypes.c_uint32),
                ("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("bits_per_input", ctypes.c_uint32),
                ("thresholds", ctypes.c_void_p), ("input_permutation", ctypes.c_void_p), ("bloom_bits", ctypes.c_void_p)]


class NativeSynthesizer:
    """zg_wnn_*: native witness synthesis of the WNN circuit (host code; usable without a GPU).  Built from a
    zg_b200.wnn.Wnn; `synthesize(image, k, usable_rows)` returns (six (n,4) uint64 Montgomery columns, class scores)."""

    def __init__(self, wnn):
        self._L = load_library()
        cp = wnn.get_circuit_params()
        thr = np.ascontiguousarray(wnn.binarization_thresholds, dtype=np.uint16)
        perm = np.ascontiguousarray(wnn.input_permutation, dtype=np.uint64)
        bloom = np.ascontiguousarray(wnn.bloom_filters, dtype=np.uint8)
        self._keep = (thr, perm, bloom)
        d = WnnDesc()
        d.p, d.n_hashes, d.bits_per_hash, d.bits_per_filter = cp["p"], cp["n_hashes"], cp["bits_per_hash"], cp["bits_per_filter"]
        d.n_classes, d.n_filters = bloom.shape[0], bloom.shape[1]
        d.width, d.height, d.bits_per_input = thr.shape
        d.thresholds, d.input_permutation, d.bloom_bits = thr.ctypes.data, perm.ctypes.data, bloom.ctypes.data
        assert bloom.shape[2] == 1 << cp["bits_per_hash"]
        h = ctypes.c_void_p()
        rc = self._L.zg_wnn_create(ctypes.byref(d), ctypes.byref(h))
        if rc != ZG_OK:
            raise ZgError(rc, "zg_wnn_create: unsupported model parameters")
        self._h, self.n_classes, self.shape = h, int(d.n_classes), (int(d.width), int(d.height))

    def synthesize(self, image, k: int, usable_rows: int, out=None):
        img = np.ascontiguousarray(image, dtype=np.uint8)
        assert img.shape == self.shape
        n = 1 << k
        cols = out if out is not None else [np.empty((n, 4), dtype=np.uint64) for _ in range(6)]
        ptrs = (ctypes.c_void_p * 6)(*[c.ctypes.data for c in cols])
        scores = (ctypes.c_uint64 * self.n_classes)()
        rc = self._L.zg_wnn_synthesize(self._h, img.ctypes.data, k, usable_rows, ptrs, scores)
        if rc != ZG_OK:
            raise ZgError(rc, self._L.zg_wnn_last_error(self._h).decode())
        return cols, [int(x) for x in scores]

    def close(self):
        if getattr(self, "_h", None):
            self._L.zg_wnn_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
