"""Host-side mirror of the halo2_proofs circuit front-end that `zero_g` programs against.

Restates (from the published behaviour of halo2_proofs tag v2023_04_20, an un-vendored
dependency pinned by /root/reference/Cargo.toml:21-25):
  * `plonk::Expression`, `ConstraintSystem` (queries, gates, lookups, permutation columns,
    `degree`, `blinding_factors`, `compress_selectors`),
  * `circuit::floor_planner::SimpleFloorPlanner` (region placement, constants column,
    table back-fill), `Region::{assign_advice, assign_advice_from_constant, copy_advice}`,
  * `permutation::keygen::Assembly::copy` (cycle merge by size).
The reference uses these through src/gadgets/** (e.g. src/gadgets/wnn.rs:334-393).  Witness
synthesis stays on the host by design (BASELINE.json north_star); this module only builds the
columns and the constraint-system description that cross the C ABI.  Field values are Python
ints mod r (canonical)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

from ..bn254_host import R_MOD

ADVICE, FIXED, INSTANCE = "advice", "fixed", "instance"


# --------------------------------------------------------------------------------------------
# Expressions
# --------------------------------------------------------------------------------------------
class Expr:
    """plonk::Expression.  kinds: const, selector, fixed, advice, instance, neg, sum, prod, scaled"""
    __slots__ = ("kind", "a", "b", "v")

    def __init__(self, kind, a=None, b=None, v=None):
        self.kind, self.a, self.b, self.v = kind, a, b, v

    # operator sugar mirroring the Rust impls (Add/Sub/Mul/Neg; Mul<F> = Scaled)
    def __add__(self, o):
        return Expr("sum", self, _e(o))

    def __sub__(self, o):
        return Expr("sum", self, Expr("neg", _e(o)))

    def __neg__(self):
        return Expr("neg", self)

    def __mul__(self, o):
        if isinstance(o, int):
            return Expr("scaled", self, v=o % R_MOD)
        return Expr("prod", self, o)

    def degree(self) -> int:
        k = self.kind
        if k == "const":
            return 0
        if k in ("selector", "fixed", "advice", "instance"):
            return 1
        if k in ("neg", "scaled"):
            return self.a.degree()
        if k == "sum":
            return max(self.a.degree(), self.b.degree())
        return self.a.degree() + self.b.degree()

    def simple_selectors(self, out=None):
        out = set() if out is None else out
        if self.kind == "selector":
            if self.v[1]:
                out.add(self.v[0])
        else:
            for c in (self.a, self.b):
                if isinstance(c, Expr):
                    c.simple_selectors(out)
        return out

    def map_selectors(self, fn):
        k = self.kind
        if k == "selector":
            return fn(self.v)
        if k in ("const", "fixed", "advice", "instance"):
            return self
        if k in ("neg",):
            return Expr(k, self.a.map_selectors(fn))
        if k == "scaled":
            return Expr(k, self.a.map_selectors(fn), v=self.v)
        return Expr(k, self.a.map_selectors(fn), self.b.map_selectors(fn))

    def evaluate(self, get) -> int:
        """get(kind, query_index) -> int (row context is bound by the caller)."""
        k = self.kind
        if k == "const":
            return self.v
        if k in ("fixed", "advice", "instance"):
            return get(k, self.v[0])
        if k == "neg":
            return (-self.a.evaluate(get)) % R_MOD
        if k == "scaled":
            return self.a.evaluate(get) * self.v % R_MOD
        if k == "sum":
            return (self.a.evaluate(get) + self.b.evaluate(get)) % R_MOD
        if k == "prod":
            return self.a.evaluate(get) * self.b.evaluate(get) % R_MOD
        raise ValueError("selector left in expression")


def _e(x):
    return x if isinstance(x, Expr) else Expr("const", v=x % R_MOD)


def Const(v: int) -> Expr:
    return Expr("const", v=v % R_MOD)


@dataclass(frozen=True)
class Column:
    kind: str
    index: int


@dataclass(frozen=True)
class Selector:
    index: int
    simple: bool


@dataclass
class Gate:
    name: str
    polys: List[Expr]


@dataclass
class Lookup:
    name: str
    inputs: List[Expr]
    tables: List[Expr]

    def required_degree(self) -> int:
        ind = max([1] + [e.degree() for e in self.inputs])
        td = max([1] + [e.degree() for e in self.tables])
        return max(4, 2 + ind + td)


class VirtualCells:
    """plonk::VirtualCells: registers queries on the constraint system as they are made."""

    def __init__(self, cs: "ConstraintSystem"):
        self.cs = cs

    def query_selector(self, s: Selector) -> Expr:
        return Expr("selector", v=(s.index, s.simple))

    def query_advice(self, col: Column, rot: int) -> Expr:
        return Expr("advice", v=(self.cs._query(ADVICE, col.index, rot), col.index, rot))

    def query_fixed(self, col: Column, rot: int = 0) -> Expr:
        return Expr("fixed", v=(self.cs._query(FIXED, col.index, rot), col.index, rot))

    def query_instance(self, col: Column, rot: int) -> Expr:
        return Expr("instance", v=(self.cs._query(INSTANCE, col.index, rot), col.index, rot))


class ConstraintSystem:
    def __init__(self):
        self.num_fixed = self.num_advice = self.num_instance = 0
        self.selectors: List[Selector] = []
        self.queries = {ADVICE: [], FIXED: [], INSTANCE: []}     # [(column, rotation)]
        self.num_advice_queries: List[int] = []
        self.gates: List[Gate] = []
        self.lookups: List[Lookup] = []
        self.permutation: List[Column] = []
        self.constants: List[Column] = []

    # --- column allocation -------------------------------------------------------------
    def advice_column(self) -> Column:
        self.num_advice += 1
        self.num_advice_queries.append(0)
        return Column(ADVICE, self.num_advice - 1)

    def fixed_column(self) -> Column:
        self.num_fixed += 1
        return Column(FIXED, self.num_fixed - 1)

    def instance_column(self) -> Column:
        self.num_instance += 1
        return Column(INSTANCE, self.num_instance - 1)

    def lookup_table_column(self) -> Column:      # TableColumn wraps a fixed column
        return self.fixed_column()

    def selector(self) -> Selector:
        s = Selector(len(self.selectors), True)
        self.selectors.append(s)
        return s

    def complex_selector(self) -> Selector:
        s = Selector(len(self.selectors), False)
        self.selectors.append(s)
        return s

    def _query(self, kind, col, rot) -> int:
        q = self.queries[kind]
        if (col, rot) in q:
            return q.index((col, rot))
        q.append((col, rot))
        if kind == ADVICE:
            self.num_advice_queries[col] += 1
        return len(q) - 1

    def enable_equality(self, col: Column):
        self._query(col.kind, col.index, 0)          # query_any_index(column, Rotation::cur())
        if col not in self.permutation:
            self.permutation.append(col)

    def enable_constant(self, col: Column):
        if col not in self.constants:
            self.constants.append(col)
            self.enable_equality(col)

    def create_gate(self, name: str, fn: Callable[[VirtualCells], Tuple[Optional[Expr], List[Expr]]]):
        """fn returns (selector_expr | None, constraints) like Constraints::with_selector."""
        sel, cons = fn(VirtualCells(self))
        polys = [(sel * c) if sel is not None else c for c in cons]
        assert polys, "gates must contain at least one constraint"
        self.gates.append(Gate(name, polys))

    def lookup(self, name: str, fn: Callable[[VirtualCells], List[Tuple[Expr, Column]]]) -> int:
        cells = VirtualCells(self)
        inputs, tables = [], []
        for inp, tab in fn(cells):
            assert not inp.simple_selectors(), "expression containing simple selector supplied to lookup argument"
            inputs.append(inp)
            tables.append(cells.query_fixed(tab, 0))
        self.lookups.append(Lookup(name, inputs, tables))
        return len(self.lookups) - 1

    # --- derived quantities ------------------------------------------------------------------
    def degree(self) -> int:
        d = 3                                                    # permutation::Argument::required_degree
        d = max([d] + [l.required_degree() for l in self.lookups])
        d = max([d] + [p.degree() for g in self.gates for p in g.polys])
        return d

    def blinding_factors(self) -> int:
        return max([3] + self.num_advice_queries) + 2

    def minimum_rows(self) -> int:
        return self.blinding_factors() + 3

    # --- selector compression (plonk/circuit/compress_selectors.rs) ----------------------------
    def compress_selectors(self, activations: List[List[bool]]) -> List[List[int]]:
        """Replaces every selector by a fixed-column expression in gates and lookups; returns the
        new fixed columns' values (appended after the existing fixed columns, in allocation order)."""
        nsel = len(self.selectors)
        assert len(activations) == nsel
        degrees = [0] * nsel
        for g in self.gates:
            for p in g.polys:
                ss = p.simple_selectors()
                assert len(ss) <= 1, "at most one simple selector per constraint"
                for s in ss:
                    degrees[s] = max(degrees[s], p.degree())
        max_degree = self.degree()
        n = len(activations[0]) if nsel else 0
        polys: List[List[int]] = []
        replacement: List[Optional[Expr]] = [None] * nsel
        cells = VirtualCells(self)

        def alloc() -> Expr:
            return cells.query_fixed(self.fixed_column(), 0)

        rest = []
        for s in range(nsel):
            if degrees[s] == 0:                       # complex, or unused in gates: own column
                replacement[s] = alloc()
                polys.append([1 if b else 0 for b in activations[s]])
            else:
                rest.append(s)
        # exclusion matrix over the remaining (simple) selectors
        act_sets = {s: {i for i, b in enumerate(activations[s]) if b} for s in rest}
        added = {s: False for s in rest}
        for ii, s in enumerate(rest):
            if added[s]:
                continue
            added[s] = True
            assert degrees[s] <= max_degree
            d = degrees[s] - 1
            combo = [s]
            for t in rest[ii + 1:]:
                if d + len(combo) == max_degree:
                    break
                if added[t]:
                    continue
                if any(act_sets[t] & act_sets[u] for u in combo):
                    continue
                new_d = max(d, degrees[t] - 1)
                if new_d + len(combo) + 1 > max_degree:
                    continue
                d = new_d
                combo.append(t)
                added[t] = True
            query = alloc()
            col = [0] * n
            for root0, u in enumerate(combo):
                assigned_root = root0 + 1
                expr = query
                for root in range(1, len(combo) + 1):
                    if root != assigned_root:
                        expr = expr * (Const(root) - query)
                replacement[u] = expr
                for i in act_sets[u]:
                    col[i] = assigned_root
            polys.append(col)

        def repl(v):
            return replacement[v[0]]

        for g in self.gates:
            g.polys = [p.map_selectors(repl) for p in g.polys]
        for l in self.lookups:
            l.inputs = [e.map_selectors(repl) for e in l.inputs]
            l.tables = [e.map_selectors(repl) for e in l.tables]
        return polys


# --------------------------------------------------------------------------------------------
# Assignment: SimpleFloorPlanner over a keygen Assembly / prover WitnessCollection in one object
# --------------------------------------------------------------------------------------------
@dataclass
class Cell:
    column: Column
    row: int            # absolute row once the region is placed; region-relative while open
    value: Optional[int]
    region: int = -1


class Assembly:
    """Union of plonk::keygen::Assembly (fixed, selectors, permutation) and the prover's
    WitnessCollection (advice).  One synthesis pass fills both."""

    def __init__(self, cs: ConstraintSystem, k: int):
        self.cs, self.k, self.n = cs, k, 1 << k
        self.usable_rows = self.n - (cs.blinding_factors() + 1)
        self.fixed = [[0] * self.n for _ in range(cs.num_fixed)]
        self.advice = [[0] * self.n for _ in range(cs.num_advice)]
        self.selectors = [[False] * self.n for _ in cs.selectors]
        # permutation::keygen::Assembly
        m = len(cs.permutation)
        self.perm_cols = list(cs.permutation)
        self.mapping = [[(c, r) for r in range(self.n)] for c in range(m)]
        self.aux = [[(c, r) for r in range(self.n)] for c in range(m)]
        self.sizes = [[1] * self.n for _ in range(m)]
        self.copies = 0

    def _check_row(self, row):
        if row >= self.usable_rows:
            raise RuntimeError("not enough rows available (k = %d)" % self.k)   # Error::NotEnoughRowsAvailable

    def copy(self, lcol: Column, lrow: int, rcol: Column, rrow: int):
        self._check_row(lrow)
        self._check_row(rrow)
        lc, rc = self.perm_cols.index(lcol), self.perm_cols.index(rcol)
        left, right = self.aux[lc][lrow], self.aux[rc][rrow]
        self.copies += 1
        if left == right:
            return
        if self.sizes[left[0]][left[1]] < self.sizes[right[0]][right[1]]:
            left, right = right, left
        self.sizes[left[0]][left[1]] += self.sizes[right[0]][right[1]]
        i = right
        while True:
            self.aux[i[0]][i[1]] = left
            i = self.mapping[i[0]][i[1]]
            if i == right:
                break
        self.mapping[lc][lrow], self.mapping[rc][rrow] = self.mapping[rc][rrow], self.mapping[lc][lrow]


class Region:
    def __init__(self, layouter: "SimpleFloorPlanner", index: int):
        self.l, self.index = layouter, index
        self.columns = set()        # RegionColumn: Column | ("sel", i)
        self.row_count = 0
        self.ops = []               # deferred until the region start is known
        self.constants = []         # (value, cell)

    def _touch(self, key, offset):
        self.columns.add(key)
        self.row_count = max(self.row_count, offset + 1)

    def enable_selector(self, sel: Selector, offset: int):
        self._touch(("sel", sel.index), offset)
        self.ops.append(("sel", sel.index, offset))

    def assign_advice(self, col: Column, offset: int, value: Optional[int]) -> Cell:
        self._touch(col, offset)
        cell = Cell(col, offset, None if value is None else value % R_MOD, self.index)
        self.ops.append(("adv", cell))
        return cell

    def assign_advice_from_constant(self, col: Column, offset: int, constant: int) -> Cell:
        cell = self.assign_advice(col, offset, constant)
        self.constants.append((constant % R_MOD, cell))
        return cell

    def copy_advice(self, src: Cell, col: Column, offset: int) -> Cell:
        """AssignedCell::copy_advice: assign the same value, then constrain_equal(new, src)."""
        cell = self.assign_advice(col, offset, src.value)
        self.ops.append(("eq", cell, src))
        return cell

    def constrain_equal(self, a: Cell, b: Cell):
        self.ops.append(("eq", a, b))


class SimpleFloorPlanner:
    """circuit::floor_planner::single_pass::SingleChipLayouter."""

    def __init__(self, asm: Assembly):
        self.asm = asm
        self.cs = asm.cs
        self.col_height = {}        # RegionColumn -> first free row
        self.region_starts: List[int] = []
        self.table_columns = set()
        self.stats = {"regions": 0}

    def assign_region(self, name: str, fn: Callable[[Region], object]):
        idx = len(self.region_starts)
        region = Region(self, idx)
        result = fn(region)
        start = max([0] + [self.col_height.get(c, 0) for c in region.columns])
        self.region_starts.append(start)
        for c in region.columns:
            self.col_height[c] = start + region.row_count
        asm = self.asm
        for op in region.ops:
            if op[0] == "adv":
                cell = op[1]
                cell.row += start
                asm._check_row(cell.row)
                if cell.value is not None:
                    asm.advice[cell.column.index][cell.row] = cell.value
            elif op[0] == "sel":
                row = start + op[2]
                asm._check_row(row)
                asm.selectors[op[1]][row] = True
            else:
                a, b = op[1], op[2]
                asm.copy(a.column, a.row, b.column, b.row)   # rows are absolute by now (ops are ordered)
        if region.constants:
            assert self.cs.constants, "NotEnoughColumnsForConstants"
            ccol = self.cs.constants[0]
            for value, cell in region.constants:
                row = self.col_height.get(ccol, 0)
                asm._check_row(row)
                asm.fixed[ccol.index][row] = value
                asm.copy(ccol, row, cell.column, cell.row)
                self.col_height[ccol] = row + 1
        self.stats["regions"] += 1
        return result

    def assign_table(self, name: str, columns: List[Column], rows: List[Tuple[int, ...]]):
        """Layouter::assign_table for tables given as row tuples over `columns`; back-fills every
        column with its row-0 value up to the usable rows (SimpleTableLayouter + fill_from_row)."""
        for c in columns:
            assert c not in self.table_columns, "table column used twice"
            self.table_columns.add(c)
        first_unused = len(rows)
        assert first_unused <= self.asm.usable_rows, "not enough rows available"
        for j, c in enumerate(columns):
            colv = self.asm.fixed[c.index]
            for i, r in enumerate(rows):
                colv[i] = r[j] % R_MOD
            default = rows[0][j] % R_MOD
            for i in range(first_unused, self.asm.usable_rows):
                colv[i] = default

    def constrain_instance(self, cell: Cell, col: Column, row: int):
        self.asm.copy(cell.column, cell.row, col, row)
