"""Serialises a (selector-compressed) ConstraintSystem into the u32 word stream + Fr constant pool that
crosses the C ABI in `zg_pk_desc` (include/zg_b200.h).  Expressions become RPN programs."""
from __future__ import annotations

import numpy as np

from ..bn254_host import R_MOD, to_limbs
from .circuit import ADVICE, FIXED, INSTANCE, ConstraintSystem, Expr

MAGIC = 0x5A473031
OP_CONST, OP_ADVICE, OP_FIXED, OP_INSTANCE, OP_NEG, OP_ADD, OP_MUL, OP_SCALE, OP_SUB = range(9)
_KIND_OP = {ADVICE: OP_ADVICE, FIXED: OP_FIXED, INSTANCE: OP_INSTANCE}
_KIND_ID = {ADVICE: 0, FIXED: 1, INSTANCE: 2}
MAX_STACK = 12


class _Pool:
    def __init__(self):
        self.vals, self.index = [], {}

    def get(self, v: int) -> int:
        v %= R_MOD
        if v not in self.index:
            self.index[v] = len(self.vals)
            self.vals.append(v)
        return self.index[v]


def _emit(e: Expr, pool: _Pool, out: list) -> int:
    """appends RPN ops for e; returns the stack depth needed."""
    k = e.kind
    if k == "const":
        out.append(OP_CONST | (pool.get(e.v) << 8))
        return 1
    if k in _KIND_OP:
        out.append(_KIND_OP[k] | (e.v[0] << 8))
        return 1
    if k == "neg":
        d = _emit(e.a, pool, out)
        out.append(OP_NEG)
        return d
    if k == "scaled":
        d = _emit(e.a, pool, out)
        out.append(OP_SCALE | (pool.get(e.v) << 8))
        return d
    if k == "sum" and e.b.kind == "neg":
        da = _emit(e.a, pool, out)
        db = _emit(e.b.a, pool, out)
        out.append(OP_SUB)
        return max(da, db + 1)
    if k in ("sum", "prod"):
        da = _emit(e.a, pool, out)
        db = _emit(e.b, pool, out)
        out.append(OP_ADD if k == "sum" else OP_MUL)
        return max(da, db + 1)
    raise ValueError("cannot serialise expression kind %s (selectors must be compressed first)" % k)


def serialize_cs(cs: ConstraintSystem):
    """-> (words uint32 array, constants (m,4) uint64 Montgomery limbs)."""
    pool = _Pool()
    progs = []
    for g in cs.gates:
        for p in g.polys:
            progs.append(p)
    n_gate = len(progs)
    lk = []
    for l in cs.lookups:
        in_first = len(progs)
        progs += l.inputs
        tab_first = len(progs)
        progs += l.tables
        lk.append((in_first, len(l.inputs), tab_first, len(l.tables)))
    ops, off = [], [0]
    for p in progs:
        depth = _emit(p, pool, ops)
        assert depth <= MAX_STACK, "expression needs a deeper evaluation stack (%d)" % depth
        off.append(len(ops))
    w = [MAGIC, cs.num_advice, cs.num_fixed, cs.num_instance, cs.degree(), cs.blinding_factors()]
    for kind in (ADVICE, FIXED, INSTANCE):
        q = cs.queries[kind]
        w.append(len(q))
        for col, rot in q:
            w += [col, rot & 0xFFFFFFFF]
    w.append(len(cs.permutation))
    for c in cs.permutation:
        w += [_KIND_ID[c.kind], c.index]
    w.append(len(progs))
    w += off
    w.append(n_gate)
    w.append(len(lk))
    for t in lk:
        w += list(t)
    w.append(len(ops))
    w += ops
    return np.array(w, dtype=np.uint32), to_limbs(pool.vals)
