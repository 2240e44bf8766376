"""The oracle's OWN constraint-system bookkeeping.  TEST INFRASTRUCTURE ONLY.

A second, independently written restatement of what halo2_proofs tag v2023_04_20 (un-vendored;
/root/reference/Cargo.toml:21-25) does with a circuit description between `configure` and
`keygen_vk`:
  * `Expression::{degree, evaluate}`                         (plonk/circuit.rs)
  * `ConstraintSystem::{degree, blinding_factors}`           (plonk/circuit.rs)
  * `ConstraintSystem::compress_selectors` + `compress_selectors::process`
                                                             (plonk/circuit/compress_selectors.rs)
so that the oracle's keygen / prover / verifier (halo2_ref.py) take nothing but DATA from the
product's front-end: `RefCS.from_frontend` walks a not-yet-compressed constraint system by attribute
access only (kind / a / b / v of expression nodes, the query tables, the permutation columns) and
copies it into the classes below.  Nothing here imports the package.  tests/test_frontend_pinned.py
asserts that this restatement and the product's `zg_b200.plonk.circuit.ConstraintSystem` agree on
degree, blinding factors, fixed columns after selector compression, query tables and every
substituted gate / lookup expression, for the toy circuit of src/gadgets/wnn.rs:401-494 and the
checked-in models.

Expressions are nested tuples:
  ("const", v) ("selector", index, simple) ("advice"|"fixed"|"instance", query_index, column, rotation)
  ("neg", a) ("sum", a, b) ("prod", a, b) ("scaled", a, v)
"""
from __future__ import annotations

R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617

ADVICE, FIXED, INSTANCE = "advice", "fixed", "instance"
_LEAVES = (ADVICE, FIXED, INSTANCE)


def expr_degree(e) -> int:
    t = e[0]
    if t == "const":
        return 0
    if t == "selector" or t in _LEAVES:
        return 1
    if t in ("neg", "scaled"):
        return expr_degree(e[1])
    if t == "sum":
        return max(expr_degree(e[1]), expr_degree(e[2]))
    if t == "prod":
        return expr_degree(e[1]) + expr_degree(e[2])
    raise ValueError(t)


def simple_selector_of(e):
    """Expression::extract_simple_selector: the one simple selector of a polynomial, None if there is none;
    two different ones are a circuit bug upstream (panics)."""
    found = set()

    def walk(x):
        t = x[0]
        if t == "selector":
            if x[2]:
                found.add(x[1])
        elif t in ("neg", "scaled"):
            walk(x[1])
        elif t in ("sum", "prod"):
            walk(x[1])
            walk(x[2])
    walk(e)
    if len(found) > 1:
        raise AssertionError("two simple selectors in one constraint")
    return next(iter(found)) if found else None


def substitute(e, repl):
    t = e[0]
    if t == "selector":
        return repl[e[1]]
    if t == "const" or t in _LEAVES:
        return e
    if t == "neg":
        return ("neg", substitute(e[1], repl))
    if t == "scaled":
        return ("scaled", substitute(e[1], repl), e[2])
    return (t, substitute(e[1], repl), substitute(e[2], repl))


class Node:
    """What halo2_ref.py walks: .kind / .a / .b / .v, like the front-end's nodes, plus evaluate()."""
    __slots__ = ("kind", "a", "b", "v")

    def __init__(self, t):
        self.kind = t[0]
        self.a = self.b = self.v = None
        if self.kind == "const":
            self.v = t[1] % R_MOD
        elif self.kind in _LEAVES:
            self.v = (t[1], t[2], t[3])
        elif self.kind == "neg":
            self.a = Node(t[1])
        elif self.kind == "scaled":
            self.a, self.v = Node(t[1]), t[2] % R_MOD
        elif self.kind in ("sum", "prod"):
            self.a, self.b = Node(t[1]), Node(t[2])
        else:
            raise ValueError("selector left in a compressed expression" if self.kind == "selector" else self.kind)

    def evaluate(self, get) -> int:
        k = self.kind
        if k == "const":
            return self.v
        if k in _LEAVES:
            return get(k, self.v[0])
        if k == "neg":
            return (R_MOD - self.a.evaluate(get)) % R_MOD
        if k == "scaled":
            return self.a.evaluate(get) * self.v % R_MOD
        x, y = self.a.evaluate(get), self.b.evaluate(get)
        return (x + y) % R_MOD if k == "sum" else x * y % R_MOD


class _Polys:
    def __init__(self, name, polys):
        self.name, self.polys = name, polys


class _Lookup:
    def __init__(self, name, inputs, tables):
        self.name, self.inputs, self.tables = name, inputs, tables


class _Col:
    def __init__(self, kind, index):
        self.kind, self.index = kind, index

    def __eq__(self, o):
        return (self.kind, self.index) == (o.kind, o.index)

    def __hash__(self):
        return hash((self.kind, self.index))


def _to_tuple(e):
    """front-end expression node (attribute access only) -> nested tuple"""
    k = e.kind
    if k == "const":
        return ("const", int(e.v) % R_MOD)
    if k == "selector":
        return ("selector", int(e.v[0]), bool(e.v[1]))
    if k in _LEAVES:
        return (k, int(e.v[0]), int(e.v[1]), int(e.v[2]))
    if k == "neg":
        return ("neg", _to_tuple(e.a))
    if k == "scaled":
        return ("scaled", _to_tuple(e.a), int(e.v) % R_MOD)
    if k in ("sum", "prod"):
        return (k, _to_tuple(e.a), _to_tuple(e.b))
    raise ValueError(k)


class RefCS:
    def __init__(self):
        self.num_advice = self.num_fixed = self.num_instance = 0
        self.selector_simple = []                                  # per selector: simple?
        self.queries = {ADVICE: [], FIXED: [], INSTANCE: []}
        self.num_advice_queries = []
        self.gate_polys = []                                       # [(name, [tuple expr])]
        self.lookup_exprs = []                                     # [(name, [inputs], [tables])]
        self.permutation = []
        self.compressed = False
        self.gates, self.lookups = [], []                          # Node form, filled by compress_selectors

    @classmethod
    def from_frontend(cls, cs) -> "RefCS":
        if any(getattr(e, "kind", None) is None for g in cs.gates for e in g.polys):
            raise TypeError("not a front-end constraint system")
        r = cls()
        r.num_advice, r.num_fixed, r.num_instance = int(cs.num_advice), int(cs.num_fixed), int(cs.num_instance)
        r.selector_simple = [bool(s.simple) for s in cs.selectors]
        r.queries = {k: [(int(c), int(rot)) for c, rot in cs.queries[k]] for k in _LEAVES}
        r.num_advice_queries = [int(x) for x in cs.num_advice_queries]
        r.gate_polys = [(g.name, [_to_tuple(p) for p in g.polys]) for g in cs.gates]
        r.lookup_exprs = [(l.name, [_to_tuple(e) for e in l.inputs], [_to_tuple(e) for e in l.tables]) for l in cs.lookups]
        r.permutation = [_Col(c.kind, int(c.index)) for c in cs.permutation]

        def has_selector(e):
            return e[0] == "selector" or any(has_selector(x) for x in e[1:] if isinstance(x, tuple))
        n_sel_used = any(has_selector(p) for _, ps in r.gate_polys for p in ps)
        if r.selector_simple and not n_sel_used and r.gate_polys:
            raise ValueError("RefCS.from_frontend needs the constraint system BEFORE selector compression")
        return r

    # ---- derived quantities -----------------------------------------------------------------------
    def degree(self) -> int:
        d = 3                                                     # permutation::Argument::required_degree
        for _, ins, tabs in self.lookup_exprs:
            di = max([1] + [expr_degree(e) for e in ins])
            dt = max([1] + [expr_degree(e) for e in tabs])
            d = max(d, max(4, 2 + di + dt))
        for _, ps in self.gate_polys:
            for p in ps:
                d = max(d, expr_degree(p))
        return d

    def blinding_factors(self) -> int:
        return max(3, max(self.num_advice_queries + [1])) + 2

    def _query_fixed(self, col, rot=0) -> int:
        q = self.queries[FIXED]
        if (col, rot) not in q:
            q.append((col, rot))
        return q.index((col, rot))

    # ---- selector compression ---------------------------------------------------------------------
    def compress_selectors(self, activations):
        """Returns the new fixed columns (lists of ints) in allocation order and substitutes every selector."""
        assert not self.compressed and len(activations) == len(self.selector_simple)
        nsel = len(activations)
        deg = [0] * nsel
        for _, ps in self.gate_polys:
            for p in ps:
                s = simple_selector_of(p)
                if s is not None:
                    deg[s] = max(deg[s], expr_degree(p))
        max_degree = self.degree()
        n = len(activations[0]) if nsel else 0
        new_cols, repl = [], [None] * nsel

        def alloc():
            col = self.num_fixed
            self.num_fixed += 1
            return (FIXED, self._query_fixed(col, 0), col, 0)

        # degree-0 selectors (complex, or absent from every gate) keep a column of their own, in selector order
        simple = []
        for s in range(nsel):
            if deg[s] == 0:
                repl[s] = alloc()
                new_cols.append([1 if b else 0 for b in activations[s]])
            else:
                simple.append(s)
        rows = {s: frozenset(i for i, b in enumerate(activations[s]) if b) for s in simple}
        # exclusion[i][j], j < i: selectors simple[i] and simple[j] are enabled on a common row
        excl = [[not rows[simple[i]].isdisjoint(rows[simple[j]]) for j in range(i)] for i in range(len(simple))]
        done = [False] * len(simple)
        for i in range(len(simple)):
            if done[i]:
                continue
            done[i] = True
            assert deg[simple[i]] <= max_degree
            d = deg[simple[i]] - 1
            members = [i]
            for j in range(i + 1, len(simple)):
                if d + len(members) == max_degree:
                    break
                if done[j] or any(excl[j][m] for m in members):
                    continue
                nd = max(d, deg[simple[j]] - 1)
                if nd + len(members) + 1 > max_degree:
                    continue
                d = nd
                members.append(j)
                done[j] = True
            q = alloc()
            col = [0] * n
            for pos, mi in enumerate(members):
                root = pos + 1
                e = q
                for r in range(1, len(members) + 1):
                    if r != root:
                        e = ("prod", e, ("sum", ("const", r), ("neg", q)))
                repl[simple[mi]] = e
                for row in rows[simple[mi]]:
                    col[row] = root
            new_cols.append(col)
        self.gate_polys = [(nm, [substitute(p, repl) for p in ps]) for nm, ps in self.gate_polys]
        self.lookup_exprs = [(nm, [substitute(e, repl) for e in ins], [substitute(e, repl) for e in tabs])
                             for nm, ins, tabs in self.lookup_exprs]
        self.gates = [_Polys(nm, [Node(p) for p in ps]) for nm, ps in self.gate_polys]
        self.lookups = [_Lookup(nm, [Node(e) for e in ins], [Node(e) for e in tabs]) for nm, ins, tabs in self.lookup_exprs]
        self.compressed = True
        return new_cols


def frontend_expr_tuple(e):
    """The same nested-tuple form for a (compressed or not) front-end expression: lets tests compare structures."""
    return _to_tuple(e)
