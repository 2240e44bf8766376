"""MockProver-style constraint check (dev::MockProver::assert_satisfied as used by
/root/reference/src/wnn.rs:203-210): gates on every usable row, lookup membership, copy constraints.
Host-side, pure Python; never on the proving path."""
from __future__ import annotations

from ..bn254_host import R_MOD
from .circuit import ADVICE, FIXED, INSTANCE, Assembly, ConstraintSystem


def finalize_fixed(cs: ConstraintSystem, asm: Assembly):
    """compress_selectors once; returns the full list of fixed columns (keygen order)."""
    if not getattr(asm, "_compressed", False):
        asm.fixed = asm.fixed + cs.compress_selectors(asm.selectors)
        asm._compressed = True
    return asm.fixed


def assert_satisfied(cs: ConstraintSystem, asm: Assembly, instances):
    n = asm.n
    fixed = finalize_fixed(cs, asm)
    inst = [list(v) + [0] * (n - len(v)) for v in instances]
    cols = {ADVICE: asm.advice, FIXED: fixed, INSTANCE: inst}
    usable = asm.usable_rows

    def getter(row):
        def get(kind, qi):
            col, rot = cs.queries[kind][qi]
            return cols[kind][col][(row + rot) % n]
        return get

    for g in cs.gates:
        for pi, poly in enumerate(g.polys):
            for row in range(usable):
                if poly.evaluate(getter(row)) != 0:
                    raise AssertionError("gate %s[%d] not satisfied at row %d" % (g.name, pi, row))
    for l in cs.lookups:
        table = set()
        for row in range(usable):
            gt = getter(row)
            table.add(tuple(e.evaluate(gt) for e in l.tables))
        for row in range(usable):
            gt = getter(row)
            v = tuple(e.evaluate(gt) for e in l.inputs)
            if v not in table:
                raise AssertionError("lookup %s input %s at row %d not in table" % (l.name, v, row))
    pcols = asm.perm_cols
    for c, col in enumerate(pcols):
        for row in range(n):
            c2, r2 = asm.mapping[c][row]
            if (c2, r2) != (c, row):
                a = cols[col.kind][col.index][row]
                b = cols[pcols[c2].kind][pcols[c2].index][r2]
                if a != b:
                    raise AssertionError("copy constraint violated: %s[%d] != %s[%d]" % (col, row, pcols[c2], r2))
