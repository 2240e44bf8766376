"""The generated PTX bodies of csrc/field_gen.cuh (dedicated squaring, two products under one Montgomery reduction),
interpreted instruction by instruction with exact carry-flag semantics (scripts/gen_field_ops.py) and checked against
Python big integers.  The GPU-side check of the compiled text is tests/test_gpu_field.py (ops 9-11)."""
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import gen_field_ops as g  # noqa: E402

R = 1 << 256


def _edge(mod):
    return [0, 1, 2, mod - 1, mod - 2, (1 << 253) % mod, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, mod >> 1,
            (1 << 128) - 1, (1 << 224) + 1, mod - (1 << 32), int("55" * 31, 16) % mod, int("aa" * 31, 16) % mod]


def _inputs(**vals):
    d = {}
    for name, v in vals.items():
        for i, x in enumerate(g.limbs(v)):
            d["%s%d" % (name, i)] = x
    return d


def _out(res):
    return g.unlimbs([res["r%d" % i] for i in range(8)])


@pytest.mark.parametrize("field", sorted(g.FIELDS))
def test_sqr_body_matches_big_ints(field):
    mod = g.FIELDS[field]
    rinv = pow(R, -1, mod)
    prog = g.gen_sqr(mod)
    rnd = random.Random(7)
    for a in _edge(mod) + [rnd.randrange(mod) for _ in range(400)]:
        t = _out(prog.run(_inputs(a=a)))
        assert t < 2 * mod, "one conditional subtraction must suffice"
        assert t % mod == a * a * rinv % mod
    # the square costs 36 wide products, the reduction 64 (+ 8 mul.lo for the quotient digits)
    assert (prog.count("mad") + prog.count("mul") - 8) // 2 == 100


@pytest.mark.parametrize("field", sorted(g.FIELDS))
def test_two_product_body_matches_big_ints(field):
    mod = g.FIELDS[field]
    rinv = pow(R, -1, mod)
    prog = g.gen_mul2(mod)
    rnd = random.Random(11)
    edge = _edge(mod) + [mod]        # p itself is a legal operand (fp_neg_lazy(0))
    cases = [(a, b, c, d) for a in edge[:6] + [mod] for b in edge[:6] + [mod] for c in (0, mod - 1, mod) for d in (1, mod - 1, mod)]
    cases += [tuple(rnd.choice(edge) for _ in range(4)) for _ in range(300)]
    cases += [tuple(rnd.randrange(mod) for _ in range(4)) for _ in range(400)]
    worst = 0
    for a, b, c, d in cases:
        t = _out(prog.run(_inputs(a=a, b=b, c=c, d=d)))
        assert t < 3 * mod, "two conditional subtractions must suffice"
        assert t % mod == (a * b + c * d) * rinv % mod
        worst = max(worst, t // mod)
    assert worst >= 1       # the cases do reach the second subtraction's range
    assert (prog.count("mad") + prog.count("mul") - 8) // 2 == 192


def test_programs_never_read_a_dead_carry_flag_or_register():
    # Prog.run raises on a carry flag or register read before it is written; one run per program suffices because the
    # programs are straight-line
    for mod in g.FIELDS.values():
        g.gen_sqr(mod).run(_inputs(a=5))
        g.gen_mul2(mod).run(_inputs(a=5, b=6, c=7, d=8))


def test_committed_header_is_the_generator_output():
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gen_field_ops.py"), "--check"],
                        capture_output=True, text=True)
    assert rc.returncode == 0, rc.stdout + rc.stderr


@pytest.mark.parametrize("field", sorted(g.FIELDS))
def test_karatsuba_body_matches_big_ints(field):
    # generator-only body (not in the shipped header: 6 % fewer heavy-pipe cycles for 3x the ALU work, see the generator)
    mod = g.FIELDS[field]
    rinv = pow(R, -1, mod)
    prog = g.gen_mulk(mod)
    rnd = random.Random(13)
    edge = _edge(mod) + [1 << 128, (1 << 128) + 1, ((1 << 128) - 1) << 125]
    cases = [(a, b) for a in edge for b in edge] + [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(500)]
    for _ in range(100):        # equal halves, halves ordered both ways: every sign of the middle term
        lo, hi = rnd.randrange(1 << 126), rnd.randrange(1 << 125)
        cases += [((hi << 128 | lo) % mod, (lo << 128 | hi) % mod), ((lo << 128 | lo) % mod, rnd.randrange(mod))]
    for a, b in cases:
        t = _out(prog.run(_inputs(a=a, b=b)))
        assert t < 2 * mod and t % mod == a * b * rinv % mod
    assert (prog.count("mad") + prog.count("mul") - 8) // 2 == 112
