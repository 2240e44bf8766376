"""ctypes front-end of the C oracle (oracle/zg_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu_tag() -> str:
    """-march=native output is only valid on the CPU that built it; the GPU box has another
    host CPU, so the library name carries a hash of this machine's ISA flags."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


_SO = os.path.join(_HERE, "build", "libzg_oracle-%s.so" % _cpu_tag())


def build(force: bool = False) -> str:
    """Compiles oracle/zg_oracle.c for this machine's CPU.  Safe when several processes start at once (one rank per GPU
    under torchrun): an exclusive file lock serialises them and the library appears under its final name atomically."""
    import fcntl
    src = os.path.join(_HERE, "zg_oracle.c")

    def stale():
        return force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if not stale():
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    with open(_SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if stale():                      # another process may have built it while we waited
                tmp = "%s.tmp.%d" % (_SO, os.getpid())
                subprocess.check_call(["make", "-C", _HERE, "-s", "OUT=" + tmp])
                os.replace(tmp, _SO)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.zgo_num_threads.restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _fr(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == 4
    return a


def num_threads() -> int:
    return lib().zgo_num_threads()


def use_all_cores() -> int:
    """OpenMP threads = host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1 in its workers)."""
    lib().zgo_set_num_threads(ctypes.c_int(os.cpu_count() or 1))
    return num_threads()


def best_multiexp(coeffs: np.ndarray, bases: np.ndarray, threads: int = 0) -> np.ndarray:
    """halo2 best_multiexp; coeffs (n,4) Fr Montgomery, bases (n,8) G1Affine -> (12,) Jacobian."""
    coeffs, bases = _fr(coeffs), np.ascontiguousarray(bases, dtype=np.uint64)
    n = coeffs.shape[0]
    assert bases.shape == (n, 8)
    out = np.zeros(12, dtype=np.uint64)
    lib().zgo_best_multiexp(_p(coeffs), _p(bases), ctypes.c_size_t(n), ctypes.c_int(threads), _p(out))
    return out


def msm_naive(coeffs, bases) -> np.ndarray:
    coeffs, bases = _fr(coeffs), np.ascontiguousarray(bases, dtype=np.uint64)
    out = np.zeros(12, dtype=np.uint64)
    lib().zgo_msm_naive(_p(coeffs), _p(bases), ctypes.c_size_t(coeffs.shape[0]), _p(out))
    return out


def g1_to_affine(jac: np.ndarray) -> np.ndarray:
    jac = np.ascontiguousarray(jac, dtype=np.uint64).reshape(-1, 12)
    out = np.zeros((jac.shape[0], 8), dtype=np.uint64)
    lib().zgo_g1_to_affine(_p(jac), _p(out), ctypes.c_size_t(jac.shape[0]))
    return out


def srs_monomial(s_mont: np.ndarray, gen_affine: np.ndarray, n: int) -> np.ndarray:
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().zgo_srs_monomial(_p(_fr(s_mont)), _p(np.ascontiguousarray(gen_affine, dtype=np.uint64)), ctypes.c_size_t(n), _p(out))
    return out


def g1_mul_many(scalars: np.ndarray, gen_affine: np.ndarray) -> np.ndarray:
    scalars = _fr(scalars)
    out = np.zeros((scalars.shape[0], 8), dtype=np.uint64)
    lib().zgo_g1_mul_many(_p(scalars), _p(np.ascontiguousarray(gen_affine, dtype=np.uint64)), ctypes.c_size_t(scalars.shape[0]), _p(out))
    return out


def best_fft(a: np.ndarray, omega: np.ndarray, log_n: int, threads: int = 0) -> np.ndarray:
    """halo2 best_fft on a copy; a (n,4) Fr Montgomery."""
    a = _fr(a).copy()
    assert a.shape[0] == 1 << log_n
    lib().zgo_best_fft(_p(a), _p(_fr(omega)), ctypes.c_uint32(log_n), ctypes.c_int(threads))
    return a


def _vec2(name):
    def f(a, b):
        a, b = _fr(a), _fr(b)
        o = np.empty_like(a)
        getattr(lib(), name)(_p(a), _p(b), _p(o), ctypes.c_size_t(a.shape[0]))
        return o
    return f


fr_mul_vec = _vec2("zgo_fr_mul_vec")
fr_add_vec = _vec2("zgo_fr_add_vec")
fr_sub_vec = _vec2("zgo_fr_sub_vec")


def fr_scale_vec(a, s):
    a = _fr(a)
    o = np.empty_like(a)
    lib().zgo_fr_scale_vec(_p(a), _p(_fr(s)), _p(o), ctypes.c_size_t(a.shape[0]))
    return o


def fr_scale_mod3(a, pw3):
    a = _fr(a).copy()
    lib().zgo_fr_scale_mod3(_p(a), _p(_fr(pw3)), ctypes.c_size_t(a.shape[0]))
    return a


def fr_batch_invert(a):
    a = _fr(a).copy()
    lib().zgo_fr_batch_invert(_p(a), ctypes.c_size_t(a.shape[0]))
    return a


def fr_eval_poly(c, x):
    c = _fr(c)
    out = np.zeros(4, dtype=np.uint64)
    lib().zgo_fr_eval_poly(_p(c), ctypes.c_size_t(c.shape[0]), _p(_fr(x)), _p(out))
    return out


def fr_from_mont(a):
    a = _fr(a)
    o = np.empty_like(a)
    lib().zgo_fr_from_mont_vec(_p(a), _p(o), ctypes.c_size_t(a.reshape(-1, 4).shape[0]))
    return o


def fr_to_mont(a):
    a = _fr(a)
    o = np.empty_like(a)
    lib().zgo_fr_to_mont_vec(_p(a), _p(o), ctypes.c_size_t(a.reshape(-1, 4).shape[0]))
    return o


def fq_from_mont(a):
    a = _fr(a)
    o = np.empty_like(a)
    lib().zgo_fq_from_mont_vec(_p(a), _p(o), ctypes.c_size_t(a.reshape(-1, 4).shape[0]))
    return o


def fq_to_mont(a):
    a = _fr(a)
    o = np.empty_like(a)
    lib().zgo_fq_to_mont_vec(_p(a), _p(o), ctypes.c_size_t(a.reshape(-1, 4).shape[0]))
    return o


def g1_sequence(gen_affine: np.ndarray, n: int) -> np.ndarray:
    """[1]G, [2]G, ... [n]G affine: cheap synthetic bases for MSM timing."""
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().zgo_g1_sequence(_p(np.ascontiguousarray(gen_affine, dtype=np.uint64)), ctypes.c_size_t(n), _p(out))
    return out


def fr_running_product(f, start, n_out):
    f = _fr(f)
    z = np.zeros((n_out, 4), dtype=np.uint64)
    lib().zgo_fr_running_product(_p(f), _p(_fr(start)), _p(z), ctypes.c_size_t(n_out))
    return z


def fr_kate_division(a, z):
    a = _fr(a)
    q = np.zeros((a.shape[0] - 1, 4), dtype=np.uint64)
    lib().zgo_fr_kate_division(_p(a), ctypes.c_size_t(a.shape[0]), _p(_fr(z)), _p(q))
    return q


def fr_mul_add_scalar(a, s, b):
    a, b = _fr(a), _fr(b)
    o = np.empty_like(a)
    lib().zgo_fr_mul_add_scalar(_p(a), _p(_fr(s)), _p(b), _p(o), ctypes.c_size_t(a.shape[0]))
    return o


def fr_from_u512(words):
    w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1, 8)
    o = np.zeros((w.shape[0], 4), dtype=np.uint64)
    lib().zgo_fr_from_u512(_p(w), _p(o), ctypes.c_size_t(w.shape[0]))
    return o


def xorshift_fr(state: np.ndarray, count: int) -> np.ndarray:
    """advances the 4 x u32 XorShiftRng state in place; returns `count` Fr::random draws (Montgomery)."""
    assert state.dtype == np.uint32 and state.shape == (4,)
    out = np.zeros((count, 4), dtype=np.uint64)
    lib().zgo_xorshift_fr(_p(state), ctypes.c_size_t(count), _p(out))
    return out


def fr_powers(w, start, n: int) -> np.ndarray:
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().zgo_fr_powers(_p(_fr(w)), _p(_fr(start)), _p(out), ctypes.c_size_t(n))
    return out


def permute_expression_pair(a, s, usable: int):
    a, s = _fr(a), _fr(s)
    pa = np.zeros((usable, 4), dtype=np.uint64)
    ps = np.zeros((usable, 4), dtype=np.uint64)
    lib().zgo_permute_expression_pair.restype = ctypes.c_int
    rc = lib().zgo_permute_expression_pair(_p(a), _p(s), ctypes.c_size_t(usable), _p(pa), _p(ps))
    if rc != 0:
        raise ValueError("ConstraintSystemFailure: lookup input not in table")
    return pa, ps


def g1_fixed_base_mul_many(scalars, gen_affine) -> np.ndarray:
    scalars = _fr(scalars)
    out = np.zeros((scalars.shape[0], 8), dtype=np.uint64)
    lib().zgo_g1_fixed_base_mul_many(_p(scalars), _p(np.ascontiguousarray(gen_affine, dtype=np.uint64)),
                                     ctypes.c_size_t(scalars.shape[0]), _p(out))
    return out


# ---- BN254 pairing (oracle/zg_oracle.c, "BN254 optimal-ate pairing") -----------------------------------
_Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_FINAL_EXP = None


def _final_exp_words() -> np.ndarray:
    global _FINAL_EXP
    if _FINAL_EXP is None:
        e = (_Q_MOD ** 12 - 1) // _R_MOD
        n = (e.bit_length() + 63) // 64
        _FINAL_EXP = np.array([(e >> (64 * i)) & ((1 << 64) - 1) for i in range(n)], dtype=np.uint64)
    return _FINAL_EXP


def g2_generator() -> np.ndarray:
    """(16,) uint64: x.c0 | x.c1 | y.c0 | y.c1 of the alt_bn128 G2 generator, Montgomery limbs (halo2curves G2Affine layout)."""
    out = np.zeros(16, dtype=np.uint64)
    lib().zgo_g2_generator(_p(out))
    return out


def g2_mul(point: np.ndarray, scalar_mont) -> np.ndarray:
    out = np.zeros(16, dtype=np.uint64)
    lib().zgo_g2_mul(_p(np.ascontiguousarray(point, dtype=np.uint64)), _p(_fr(scalar_mont)), _p(out))
    return out


def g2_on_curve(point: np.ndarray) -> bool:
    lib().zgo_g2_on_curve.restype = ctypes.c_int
    return bool(lib().zgo_g2_on_curve(_p(np.ascontiguousarray(point, dtype=np.uint64))))


def pairing(p_affine: np.ndarray, q_g2: np.ndarray) -> np.ndarray:
    """e(P, Q) as 12 x 4 Montgomery limbs of its Fq12 coefficients (basis 1, w, ..., w^11)."""
    e = _final_exp_words()
    out = np.zeros((12, 4), dtype=np.uint64)
    lib().zgo_pairing(_p(np.ascontiguousarray(p_affine, dtype=np.uint64)), _p(np.ascontiguousarray(q_g2, dtype=np.uint64)),
                      _p(e), ctypes.c_size_t(e.shape[0]), _p(out))
    return out


def pairing_check(ps: np.ndarray, qs: np.ndarray) -> bool:
    """prod_i e(P_i, Q_i) == 1; ps (n, 8) G1 affine, qs (n, 16) G2 affine"""
    ps = np.ascontiguousarray(ps, dtype=np.uint64).reshape(-1, 8)
    qs = np.ascontiguousarray(qs, dtype=np.uint64).reshape(-1, 16)
    assert ps.shape[0] == qs.shape[0]
    e = _final_exp_words()
    lib().zgo_pairing_check.restype = ctypes.c_int
    return bool(lib().zgo_pairing_check(_p(ps), _p(qs), ctypes.c_size_t(ps.shape[0]), _p(e), ctypes.c_size_t(e.shape[0])))


def f12_mul(a, b) -> np.ndarray:
    out = np.zeros((12, 4), dtype=np.uint64)
    lib().zgo_f12_mul(_p(np.ascontiguousarray(a, dtype=np.uint64)), _p(np.ascontiguousarray(b, dtype=np.uint64)), _p(out))
    return out


def f12_pow(a, e: int) -> np.ndarray:
    n = max(1, (e.bit_length() + 63) // 64)
    w = np.array([(e >> (64 * i)) & ((1 << 64) - 1) for i in range(n)], dtype=np.uint64)
    out = np.zeros((12, 4), dtype=np.uint64)
    lib().zgo_f12_pow(_p(np.ascontiguousarray(a, dtype=np.uint64)), _p(w), ctypes.c_size_t(n), _p(out))
    return out
