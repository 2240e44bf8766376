"""Error convention of the C ABI (include/zg_b200.h): negative ZG_E_* codes with a message, never an abort; the
reference's callers `.unwrap()` halo2's Result, so a code + message is what the Rust shim turns into the panic text."""
import numpy as np
import pytest

import bn254
import cpu_ref

pytestmark = pytest.mark.gpu

GEN = bn254.g1_affine_to_limbs([bn254.G1_GEN])[0]


def _fresh():
    import torch
    import zg_b200
    return zg_b200.Context(0, torch.cuda.Stream().cuda_stream)


def test_state_and_argument_errors():
    import zg_b200
    c = _fresh()
    s = bn254.fr_to_limbs([1, 2, 3, 4])
    with pytest.raises(zg_b200.ZgError) as e:
        c.msm(0, s)                                   # no SRS loaded
    assert e.value.code == -3 and "SRS" in str(e.value)
    c.srs_load(2, cpu_ref.g1_sequence(GEN, 4), None)
    assert (cpu_ref.g1_to_affine(c.msm(0, s).reshape(1, 12)) == cpu_ref.g1_to_affine(
        cpu_ref.best_multiexp(s, cpu_ref.g1_sequence(GEN, 4)).reshape(1, 12))).all()
    with pytest.raises(zg_b200.ZgError) as e:
        c.msm(1, s)                                   # Lagrange basis was not loaded
    assert e.value.code == -3
    with pytest.raises(zg_b200.ZgError) as e:
        c.msm(0, bn254.fr_to_limbs([1] * 5))          # more scalars than bases
    assert e.value.code == -1
    with pytest.raises(zg_b200.ZgError) as e:
        c.ntt(s, 0, bn254.fr_to_limbs([1]))           # log_n out of range
    assert e.value.code == -1
    with pytest.raises(zg_b200.ZgError) as e:
        c.extended_to_coeff(bn254.fr_to_limbs([0] * 8), 2, 3, 9)   # keep > 2^ext_k
    assert e.value.code == -1
    c.close()


def test_pk_load_needs_matching_srs():
    import ctypes
    import zg_b200
    c = _fresh()
    desc = zg_b200.lib.PkDesc()
    desc.k = 5
    h = ctypes.c_void_p()
    rc = c._L.zg_pk_load(c._h, ctypes.byref(desc), ctypes.byref(h))
    assert rc == -3 and not h.value
    assert b"SRS" in c._L.zg_last_error(c._h)
    c.close()


def test_degenerate_sizes_are_accepted(ctx):
    assert ctx.kate_division(bn254.fr_to_limbs([7]), bn254.fr_to_limbs([3])).shape == (0, 4)
    one = bn254.fr_to_limbs([5])
    assert (ctx.grand_product(one, one) == bn254.fr_to_limbs([1])).all()      # z[0] = 1 whatever the fraction
    assert (ctx.batch_invert(bn254.fr_to_limbs([0])) == 0).all()
