"""MockProver-style constraint check (dev::MockProver::{verify, assert_satisfied} as used by
/root/reference/src/wnn.rs:203-210 and by every gadget test under src/gadgets/**): gates on every usable row,
lookup membership, copy constraints, instance cells.  Host-side; never on the proving path.

Columns are evaluated as whole numpy object arrays of Python ints (one vector operation per expression node)."""
from __future__ import annotations

import numpy as np

from ..bn254_host import R_MOD
from .circuit import ADVICE, FIXED, INSTANCE, Assembly, ConstraintSystem


def finalize_fixed(cs: ConstraintSystem, asm: Assembly):
    """compress_selectors once; returns the full list of fixed columns (keygen order)."""
    if not getattr(asm, "_compressed", False):
        asm.fixed = asm.fixed + cs.compress_selectors(asm.selectors)
        asm._compressed = True
    return asm.fixed


def _eval_vec(e, col_at, memo):
    key = id(e)
    if key in memo:
        return memo[key]
    k = e.kind
    if k == "const":
        r = e.v
    elif k in (ADVICE, FIXED, INSTANCE):
        r = col_at(k, e.v[0])
    elif k == "neg":
        r = (-_eval_vec(e.a, col_at, memo)) % R_MOD
    elif k == "scaled":
        r = _eval_vec(e.a, col_at, memo) * e.v % R_MOD
    elif k == "sum":
        r = (_eval_vec(e.a, col_at, memo) + _eval_vec(e.b, col_at, memo)) % R_MOD
    elif k == "prod":
        r = _eval_vec(e.a, col_at, memo) * _eval_vec(e.b, col_at, memo) % R_MOD
    else:
        raise ValueError("selector left in expression")
    memo[key] = r
    return r


def verify(cs: ConstraintSystem, asm: Assembly, instances, max_failures: int = 8):
    """MockProver::verify: returns a list of failure descriptions (empty = satisfied)."""
    n = asm.n
    usable = asm.usable_rows
    fixed = finalize_fixed(cs, asm)
    failures = []
    inst = []
    for v in instances:
        if len(v) > usable:
            return ["instance column longer than the usable rows (InstanceTooLarge)"]
        inst.append([x % R_MOD for x in v] + [0] * (n - len(v)))
    while len(inst) < cs.num_instance:
        inst.append([0] * n)
    cols = {ADVICE: asm.advice, FIXED: fixed, INSTANCE: inst}
    arr = {kind: [np.array(c, dtype=object) for c in cols[kind]] for kind in cols}
    qcache = {}

    def col_at(kind, qi):
        key = (kind, qi)
        if key not in qcache:
            col, rot = cs.queries[kind][qi]
            qcache[key] = np.roll(arr[kind][col], -rot)[:usable]
        return qcache[key]

    memo = {}
    for g in cs.gates:
        for pi, poly in enumerate(g.polys):
            r = _eval_vec(poly, col_at, memo)
            bad = np.nonzero(np.broadcast_to(r, (usable,)) != 0)[0]
            for row in bad[:max_failures]:
                failures.append("gate %s[%d] not satisfied at row %d" % (g.name, pi, int(row)))
    for l in cs.lookups:
        def tuples(exprs):
            vs = [np.broadcast_to(_eval_vec(e, col_at, memo), (usable,)) for e in exprs]
            return list(zip(*[v.tolist() for v in vs]))
        table = set(tuples(l.tables))
        for row, v in enumerate(tuples(l.inputs)):
            if v not in table:
                failures.append("lookup %s input %s at row %d not in table" % (l.name, v, row))
                if len(failures) > 4 * max_failures:
                    break
    pcols = asm.perm_cols
    for c, col in enumerate(pcols):
        mp = asm.mapping[c]
        mine = cols[col.kind][col.index]
        for row in range(n):
            c2, r2 = mp[row]
            if c2 != c or r2 != row:
                b = cols[pcols[c2].kind][pcols[c2].index][r2]
                if mine[row] != b:
                    failures.append("copy constraint violated: %s[%d] != %s[%d]" % (col, row, pcols[c2], r2))
                    if len(failures) > 8 * max_failures:
                        return failures
    return failures


def assert_satisfied(cs: ConstraintSystem, asm: Assembly, instances):
    failures = verify(cs, asm, instances)
    if failures:
        raise AssertionError("; ".join(failures[:6]))
