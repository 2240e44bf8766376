"""Device field arithmetic (csrc/field.cuh, all multiplier bodies) against Python big integers."""
import random

import numpy as np
import pytest

import bn254
from bn254 import R_MOD, Q_MOD

pytestmark = pytest.mark.gpu


def _vectors(mod, n, seed):
    rnd = random.Random(seed)
    a = [rnd.randrange(mod) for _ in range(n)]
    b = [rnd.randrange(mod) for _ in range(n)]
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << 253) % mod, (1 << 32) - 1, 1 << 32, (1 << 64) - 1]
    k = 0
    for x in edge:
        for y in edge:
            a[k], b[k] = x, y
            k += 1
    return a, b


@pytest.mark.parametrize("field,mod", [(0, R_MOD), (1, Q_MOD)])
def test_field_ops_bit_exact(ctx, field, mod):
    a, b = _vectors(mod, 4096, 17 + field)
    A, B = bn254.ints_to_limbs(a, mod), bn254.ints_to_limbs(b, mod)
    exp_mul = bn254.ints_to_limbs([x * y for x, y in zip(a, b)], mod)
    for op in (0, 1, 2, 8):  # default, portable, row-wise PTX and even/odd carry-chain bodies agree bit for bit
        assert (ctx.debug_field_op(field, op, A, B) == exp_mul).all(), op
    assert (ctx.debug_field_op(field, 3, A, B) == bn254.ints_to_limbs([x + y for x, y in zip(a, b)], mod)).all()
    assert (ctx.debug_field_op(field, 4, A, B) == bn254.ints_to_limbs([x - y for x, y in zip(a, b)], mod)).all()
    # generated bodies (csrc/field_gen.cuh): dedicated squaring, two products under one Montgomery reduction
    assert (ctx.debug_field_op(field, 9, A, B) == bn254.ints_to_limbs([x * x for x in a], mod)).all()
    assert (ctx.debug_field_op(field, 10, A, B) == bn254.ints_to_limbs([x * y + (x + y) * x for x, y in zip(a, b)], mod)).all()
    assert (ctx.debug_field_op(field, 11, A, B) == bn254.ints_to_limbs([x * x - y * y for x, y in zip(a, b)], mod)).all()
    raw = bn254.ints_to_limbs(a, mod, mont=False)
    assert (ctx.debug_field_op(field, 6, A) == raw).all()
    assert (ctx.debug_field_op(field, 7, raw) == A).all()


def test_field_inverse(ctx):
    a, _ = _vectors(R_MOD, 256, 3)
    A = bn254.fr_to_limbs(a)
    exp = bn254.fr_to_limbs([pow(x, -1, R_MOD) if x % R_MOD else 0 for x in a])
    assert (ctx.debug_field_op(0, 5, A) == exp).all()
