"""Mirror of `halo2_proofs::plonk::{ConstraintSystem, Expression}` (U: halo2_proofs 0.2.0 src/plonk/circuit.rs),
restricted to what the prover hot path consumes: query lists in registration order, gate polynomials as
expression trees, lookup arguments, the permutation column list, `degree()` and `blinding_factors()`.

Selectors are represented as fixed columns (after `compress_selectors` the verifying key only contains fixed
queries).  The output of `to_ir()` is the flat PLONKish IR that both libbzhalo2's prover and the
oracle consume; in a Rust deployment the shim fills the same IR from `pk.vk.cs`."""
from dataclasses import dataclass


class Rotation:
    @staticmethod
    def cur(): return 0
    @staticmethod
    def next(): return 1
    @staticmethod
    def prev(): return -1


class Expression:
    """Expression tree over {Constant, Fixed, Advice, Instance, Negated, Sum, Product, Scaled}."""
    __slots__ = ("kind", "a", "b")

    def __init__(self, kind, a=None, b=None):
        self.kind, self.a, self.b = kind, a, b

    # -- constructors --
    def __neg__(self): return Expression("neg", self)
    def __add__(self, o): return Expression("sum", self, _lift(o))
    def __radd__(self, o): return Expression("sum", _lift(o), self)
    def __sub__(self, o): return Expression("sum", self, Expression("neg", _lift(o)))
    def __rsub__(self, o): return Expression("sum", _lift(o), Expression("neg", self))

    def __mul__(self, o):
        if isinstance(o, int):
            return Expression("scaled", self, o)
        return Expression("product", self, o)

    def __rmul__(self, o):
        if isinstance(o, int):
            return Expression("scaled", self, o)
        return Expression("product", o, self)

    def square(self): return Expression("product", self, self)

    def degree(self):
        k = self.kind
        if k == "const": return 0
        if k in ("fixed", "advice", "instance"): return 1
        if k == "neg": return self.a.degree()
        if k == "sum": return max(self.a.degree(), self.b.degree())
        if k == "product": return self.a.degree() + self.b.degree()
        if k == "scaled": return self.a.degree()
        raise ValueError(k)

    def to_ir(self, modulus):
        k = self.kind
        if k == "const": return ["const", self.a % modulus]
        if k in ("fixed", "advice", "instance"): return [k, self.a, self.b]
        if k == "neg": return ["neg", self.a.to_ir(modulus)]
        if k == "sum": return ["sum", self.a.to_ir(modulus), self.b.to_ir(modulus)]
        if k == "product": return ["product", self.a.to_ir(modulus), self.b.to_ir(modulus)]
        if k == "scaled": return ["scaled", self.a.to_ir(modulus), self.b % modulus]
        raise ValueError(k)


def _lift(o):
    return o if isinstance(o, Expression) else Expression("const", int(o))


def Constant(v): return Expression("const", int(v))
def Advice(col, rot=0): return Expression("advice", col, rot)
def Fixed(col, rot=0): return Expression("fixed", col, rot)
def Instance(col, rot=0): return Expression("instance", col, rot)


class ConstraintSystem:
    """Builds the constraint-system description; mirrors the halo2 `configure` phase."""

    def __init__(self, modulus):
        self.modulus = modulus
        self.num_advice = self.num_fixed = self.num_instance = 0
        self.advice_queries, self.fixed_queries, self.instance_queries = [], [], []
        self.gates = []            # (name, [Expression])
        self.gate_constraint_names = []   # per gate: [constraint name] (Constraints::with_selector's labels; "" when unnamed)
        self.lookups = []          # (name, [input Expression], [table Expression])
        self.permutation = []      # [("advice"|"fixed"|"instance", index)] in enable_equality order
        self.selectors = {}        # fixed column index -> "simple" | "complex"  (selectors are kept as fixed columns, see module doc)
        self.constants_column = None      # enable_constant(column): where assign_advice_from_constant / constrain_constant put their values
        self.table_columns = []    # lookup_table_column(): fixed columns only assign_table may write

    def advice_column(self):
        self.num_advice += 1
        return self.num_advice - 1

    def fixed_column(self):
        self.num_fixed += 1
        return self.num_fixed - 1

    def instance_column(self):
        self.num_instance += 1
        return self.num_instance - 1

    def selector(self):
        """`meta.selector()`: a fixed column holding 0 / 1, switched on per row by `region.enable_selector`."""
        c = self.fixed_column()
        self.selectors[c] = "simple"
        return c

    def complex_selector(self):
        """`meta.complex_selector()`: a selector that may appear in lookup arguments (never combined with others)."""
        c = self.fixed_column()
        self.selectors[c] = "complex"
        return c

    def lookup_table_column(self):
        c = self.fixed_column()
        self.table_columns.append(c)
        return c

    def enable_constant(self, fixed_col):
        self.constants_column = fixed_col
        self.enable_equality("fixed", fixed_col)

    def _query(self, lst, col, rot):
        if (col, rot) not in lst:
            lst.append((col, rot))

    def query_advice(self, col, rot=0):
        self._query(self.advice_queries, col, rot)
        return Advice(col, rot)

    def query_fixed(self, col, rot=0):
        self._query(self.fixed_queries, col, rot)
        return Fixed(col, rot)

    def query_instance(self, col, rot=0):
        self._query(self.instance_queries, col, rot)
        return Instance(col, rot)

    def enable_equality(self, kind, col):
        """`enable_equality` registers a Rotation::cur() query and adds the column to the permutation."""
        {"advice": self.query_advice, "fixed": self.query_fixed, "instance": self.query_instance}[kind](col, 0)
        if (kind, col) not in self.permutation:
            self.permutation.append((kind, col))

    def create_gate(self, name, polys):
        """polys: expressions, or (constraint name, expression) pairs as `Constraints::with_selector` labels them."""
        assert polys, "gates must contain at least one constraint"
        names, exprs = [], []
        for p in polys:
            if isinstance(p, tuple):
                names.append(p[0]); exprs.append(p[1])
            else:
                names.append(""); exprs.append(p)
        self.gates.append((name, exprs))
        self.gate_constraint_names.append(names)

    def gate_queried_cells(self, gate_index):
        """`gate.queried_cells()`: the (kind, column, rotation) cells the gate's polynomials query, selectors excluded."""
        out = []

        def walk(e):
            if e.kind in ("advice", "fixed", "instance"):
                if not (e.kind == "fixed" and e.a in self.selectors) and (e.kind, e.a, e.b) not in out:
                    out.append((e.kind, e.a, e.b))
            elif e.kind in ("neg", "scaled"):
                walk(e.a)
            elif e.kind in ("sum", "product"):
                walk(e.a); walk(e.b)
        for poly in self.gates[gate_index][1]:
            walk(poly)
        return out

    def gate_queried_selectors(self, gate_index):
        out = []

        def walk(e):
            if e.kind == "fixed" and e.a in self.selectors and e.a not in out:
                out.append(e.a)
            elif e.kind in ("neg", "scaled"):
                walk(e.a)
            elif e.kind in ("sum", "product"):
                walk(e.a); walk(e.b)
        for poly in self.gates[gate_index][1]:
            walk(poly)
        return out

    def lookup(self, name, pairs):
        self.lookups.append((name, [p[0] for p in pairs], [p[1] for p in pairs]))

    # -- U: circuit.rs `degree()` / `blinding_factors()` --
    def degree(self):
        degree = 3 if self.permutation else 1            # permutation::Argument::required_degree
        for _, inp, tab in self.lookups:
            di = max([1] + [e.degree() for e in inp])
            dt = max([1] + [e.degree() for e in tab])
            degree = max(degree, 4, 2 + di + dt)
        for _, polys in self.gates:
            for p in polys:
                degree = max(degree, p.degree())
        return degree

    def blinding_factors(self):
        per_col = [0] * max(1, self.num_advice)
        for col, _ in self.advice_queries:
            per_col[col] += 1
        factors = max(3, max(per_col) if self.advice_queries else 1)
        return factors + 2

    def minimum_rows(self):
        return self.blinding_factors() + 3

    def to_ir(self):
        m = self.modulus
        return {
            "modulus": m,
            "num_advice": self.num_advice, "num_fixed": self.num_fixed, "num_instance": self.num_instance,
            "advice_queries": [list(q) for q in self.advice_queries],
            "fixed_queries": [list(q) for q in self.fixed_queries],
            "instance_queries": [list(q) for q in self.instance_queries],
            "gates": [{"name": n, "polys": [p.to_ir(m) for p in polys]} for n, polys in self.gates],
            "lookups": [{"name": n, "input": [e.to_ir(m) for e in i], "table": [e.to_ir(m) for e in t]}
                        for n, i, t in self.lookups],
            "permutation": [list(c) for c in self.permutation],
            "degree": self.degree(),
            "blinding_factors": self.blinding_factors(),
        }


class Assignment:
    """Mirror of the keygen `Assembly` + prover `WitnessCollection`: dense fixed / advice cell values (canonical
    ints), copy constraints and instance values for n = 2^k rows."""

    def __init__(self, cs: ConstraintSystem, k: int):
        self.cs, self.k, self.n = cs, k, 1 << k
        self.usable_rows = self.n - (cs.blinding_factors() + 1)
        self.fixed = [[0] * self.n for _ in range(cs.num_fixed)]
        self.advice = [[0] * self.n for _ in range(cs.num_advice)]
        self.instance = [[] for _ in range(cs.num_instance)]
        self.copies = []

    def _row(self, row):
        if not 0 <= row < self.usable_rows:
            raise IndexError("not enough rows available")     # Error::NotEnoughRowsAvailable
        return row

    def assign_fixed(self, col, row, v):
        self.fixed[col][self._row(row)] = v % self.cs.modulus

    def assign_advice(self, col, row, v):
        self.advice[col][self._row(row)] = v % self.cs.modulus

    def copy(self, a, b):
        """a, b = (kind, col, row)"""
        for kind, col, row in (a, b):
            assert (kind, col) in self.cs.permutation, "Error::ColumnNotInPermutation"
            self._row(row)
        self.copies.append((a, b))

    def set_instance(self, col, values):
        assert len(values) <= self.usable_rows, "Error::InstanceTooLarge"
        self.instance[col] = [v % self.cs.modulus for v in values]

    def cell(self, kind, col, row):
        if kind == "advice": return self.advice[col][row]
        if kind == "fixed": return self.fixed[col][row]
        vals = self.instance[col]
        return vals[row] if row < len(vals) else 0

    def permutation_mapping(self):
        """U: plonk/permutation/keygen.rs `Assembly::copy`: cycle merge, smaller into larger."""
        cols = self.cs.permutation
        m, n = len(cols), self.n
        mapping = [[(i, j) for j in range(n)] for i in range(m)]
        aux = [[(i, j) for j in range(n)] for i in range(m)]
        sizes = [[1] * n for _ in range(m)]
        for (ka, ca, ra), (kb, cb, rb) in self.copies:
            lc, rc = cols.index((ka, ca)), cols.index((kb, cb))
            left, right = aux[lc][ra], aux[rc][rb]
            if left == right:
                continue
            if sizes[left[0]][left[1]] < sizes[right[0]][right[1]]:
                left, right = right, left
            sizes[left[0]][left[1]] += sizes[right[0]][right[1]]
            i = right
            while True:
                aux[i[0]][i[1]] = left
                i = mapping[i[0]][i[1]]
                if i == right:
                    break
            mapping[lc][ra], mapping[rc][rb] = mapping[rc][rb], mapping[lc][ra]
        return mapping

    def check_satisfied(self):
        """MockProver-style row check of every gate / lookup / copy constraint on usable rows (host-side sanity
        for synthetic witnesses; mirrors what `MockProver::verify` does in the reference tests)."""
        p, n = self.cs.modulus, self.n

        def ev(e, row):
            k = e.kind
            if k == "const": return e.a % p
            if k in ("fixed", "advice", "instance"): return self.cell(k, e.a, (row + e.b) % n)
            if k == "neg": return (-ev(e.a, row)) % p
            if k == "sum": return (ev(e.a, row) + ev(e.b, row)) % p
            if k == "product": return ev(e.a, row) * ev(e.b, row) % p
            if k == "scaled": return ev(e.a, row) * e.b % p
        for gi, (name, polys) in enumerate(self.cs.gates):
            for pi, poly in enumerate(polys):
                for row in range(self.usable_rows):
                    if ev(poly, row) != 0:
                        return f"gate {gi} '{name}' poly {pi} row {row}"
        for name, inp, tab in self.cs.lookups:
            table = {tuple(ev(t, r) for t in tab) for r in range(self.usable_rows)}
            for row in range(self.usable_rows):
                if tuple(ev(i, row) for i in inp) not in table:
                    return f"lookup '{name}' row {row}"
        for (ka, ca, ra), (kb, cb, rb) in self.copies:
            if self.cell(ka, ca, ra) != self.cell(kb, cb, rb):
                return f"copy ({ka},{ca},{ra}) != ({kb},{cb},{rb})"
        return None


# ------------------------------------------------------------------------------------------------------
# Layouter: mirror of `SimpleFloorPlanner` / `SingleChipLayouter` (U: halo2_proofs 0.2.0 src/circuit/floor_planner/
# single_pass.rs).  A region is laid out at the earliest row for which none of the columns it touches is in use; the
# region's constants go to the first free rows of the constants column right after the region; tables are assigned from
# row 0 and padded with their first value.  Region indices, names, row extents and cell sets are recorded the way
# `MockProver` records them, because the reference's negative tests pin failures by (region index, name, offset).
# ------------------------------------------------------------------------------------------------------
class Cell:
    """`AssignedCell`: a cell of a region plus its value."""
    __slots__ = ("region", "kind", "col", "offset", "value")

    def __init__(self, region, kind, col, offset, value):
        self.region, self.kind, self.col, self.offset, self.value = region, kind, col, offset, value


class RegionRecord:
    def __init__(self, index, name):
        self.index, self.name = index, name
        self.start = 0
        self.columns = set()             # (kind, col) of the advice / fixed cells assigned in the region
        self.rows = None                 # (min, max) absolute rows of those cells
        self.cells = set()               # (kind, col, absolute row)
        self.enabled_selectors = {}      # selector column -> [absolute rows]


class Region:
    """Records one region's assignments (offsets relative to the region start); `Layouter.assign_region` places and applies them."""

    def __init__(self, layouter, index):
        self.lay, self.index = layouter, index
        self.ops = []                    # ("advice" | "fixed" | "selector", col, offset, value)
        self.equal = []                  # (Cell, Cell)
        self.constants = []              # (value, Cell)
        self.shape_columns, self.row_count = [], 0

    def _touch(self, key, offset):
        if key not in self.shape_columns:
            self.shape_columns.append(key)
        self.row_count = max(self.row_count, offset + 1)

    def enable_selector(self, sel, offset):
        assert sel in self.lay.cs.selectors, "not a selector"
        self._touch(("selector", sel), offset)
        self.ops.append(("selector", sel, offset, 1))

    def assign_advice(self, col, offset, value):
        value %= self.lay.cs.modulus
        self._touch(("advice", col), offset)
        self.ops.append(("advice", col, offset, value))
        return Cell(self.index, "advice", col, offset, value)

    def assign_advice_from_constant(self, col, offset, constant):
        cell = self.assign_advice(col, offset, constant)
        self.constants.append((constant % self.lay.cs.modulus, cell))
        return cell

    def assign_fixed(self, col, offset, value):
        value %= self.lay.cs.modulus
        self._touch(("fixed", col), offset)
        self.ops.append(("fixed", col, offset, value))
        return Cell(self.index, "fixed", col, offset, value)

    def copy_advice(self, cell, col, offset):
        """`AssignedCell::copy_advice`: assign the same value and constrain the two cells equal."""
        new = self.assign_advice(col, offset, cell.value)
        self.equal.append((new, cell))
        return new

    def constrain_equal(self, a, b):
        self.equal.append((a, b))

    def constrain_constant(self, cell, constant):
        self.constants.append((constant % self.lay.cs.modulus, cell))


class Layouter:
    def __init__(self, cs, k):
        self.cs = cs
        self.asg = Assignment(cs, k)
        self.columns = {}                # shape key -> first free row
        self.regions = []                # RegionRecord, in creation order (tables included, as MockProver counts them)
        self.used_tables = set()
        self.asg.regions = self.regions

    def _abs(self, cell):
        return (cell.kind, cell.col, self.regions[cell.region].start + cell.offset)

    def assign_region(self, name, assignment):
        index = len(self.regions)
        reg = Region(self, index)
        result = assignment(reg)
        start = 0
        for key in reg.shape_columns:
            start = max(start, self.columns.get(key, 0))
        for key in reg.shape_columns:
            self.columns[key] = start + reg.row_count
        rec = RegionRecord(index, name)
        rec.start = start
        self.regions.append(rec)
        a = self.asg
        for kind, col, offset, value in reg.ops:
            row = start + offset
            if kind == "selector":
                a.assign_fixed(col, row, 1)
                rec.enabled_selectors.setdefault(col, []).append(row)
                continue
            (a.assign_advice if kind == "advice" else a.assign_fixed)(col, row, value)
            rec.columns.add((kind, col))
            rec.rows = (row, row) if rec.rows is None else (min(rec.rows[0], row), max(rec.rows[1], row))
            rec.cells.add((kind, col, row))
        for x, y in reg.equal:
            a.copy(self._abs(x), self._abs(y))
        if reg.constants:
            assert self.cs.constants_column is not None, "Error::NotEnoughColumnsForConstants"
            cc = self.cs.constants_column
            nxt = self.columns.get(("fixed", cc), 0)
            for value, cell in reg.constants:
                a.assign_fixed(cc, nxt, value)
                a.copy(("fixed", cc, nxt), self._abs(cell))
                nxt += 1
            self.columns[("fixed", cc)] = nxt
        return result

    def assign_table(self, name, column, values):
        """`layouter.assign_table`: rows 0.. of a lookup table column, the rest of the column filled with the first value."""
        assert column in self.cs.table_columns and column not in self.used_tables, "Error::Synthesis (table column reused)"
        self.used_tables.add(column)
        rec = RegionRecord(len(self.regions), name)
        self.regions.append(rec)
        a = self.asg
        for row, v in enumerate(values):
            a.assign_fixed(column, row, v)
            rec.cells.add(("fixed", column, row))
        rec.columns.add(("fixed", column))
        rec.rows = (0, len(values) - 1)
        for row in range(len(values), a.usable_rows):
            a.assign_fixed(column, row, values[0])

    def constrain_instance(self, cell, instance_col, row):
        self.asg.copy(self._abs(cell), ("instance", instance_col, row))


# ------------------------------------------------------------------------------------------------------
# Selector compression: `ConstraintSystem::compress_selectors` + `compress_selectors::process` (U: halo2_proofs 0.2.0
# src/plonk/circuit.rs, src/plonk/circuit/compress_selectors.rs), what `keygen_vk` / `keygen_pk` / `MockProver::run` apply
# after synthesis.  Selectors leave the constraint system: complex selectors (and selectors no gate uses) get a 0 / 1 fixed
# column each; simple selectors that are never on in the same row are packed greedily into shared fixed columns holding
# 1, 2, 3, ... and every use of selector s becomes  q * prod_{i != root(s)} (i - q)  as long as the gate's degree stays within
# the constraint system's.  The verifying key -- fixed column count, fixed query list, gate polynomials -- is the result.
# ------------------------------------------------------------------------------------------------------
def compress_selectors(cs, asg):
    """-> (ConstraintSystem, Assignment) with the selector columns of `cs` replaced as halo2's keygen replaces them.  Real
    fixed columns keep their order; the columns compression allocates follow them, their queries after all other fixed queries."""
    p = cs.modulus
    sel_ids = sorted(cs.selectors)                      # allocation order
    real = [c for c in range(cs.num_fixed) if c not in cs.selectors]
    remap = {c: i for i, c in enumerate(real)}
    out = ConstraintSystem(p)
    out.num_advice, out.num_instance, out.num_fixed = cs.num_advice, cs.num_instance, len(real)
    out.advice_queries, out.instance_queries = list(cs.advice_queries), list(cs.instance_queries)
    out.fixed_queries = [(remap[c], r) for c, r in cs.fixed_queries if c in remap]
    out.permutation = [(k, remap[c] if k == "fixed" else c) for k, c in cs.permutation]
    out.constants_column = remap.get(cs.constants_column)
    out.table_columns = [remap[c] for c in cs.table_columns]
    max_degree = cs.degree()

    def simple_selector_of(e):
        """extract_simple_selector: the one simple selector of a polynomial (None if it has none)."""
        found = set()

        def walk(x):
            if x.kind == "fixed" and cs.selectors.get(x.a) == "simple":
                found.add(x.a)
            elif x.kind in ("neg", "scaled"):
                walk(x.a)
            elif x.kind in ("sum", "product"):
                walk(x.a); walk(x.b)
        walk(e)
        assert len(found) <= 1, "two simple selectors cannot be in the same expression"
        return next(iter(found)) if found else None
    degrees = {s: 0 for s in sel_ids}
    for _, polys in cs.gates:
        for poly in polys:
            s = simple_selector_of(poly)
            if s is not None:
                degrees[s] = max(degrees[s], poly.degree())
    n = asg.n
    activations = {s: [1 if v else 0 for v in asg.fixed[s]] for s in sel_ids}
    new_columns, substitution = [], {}

    def allocate():
        col = out.fixed_column()
        out._query(out.fixed_queries, col, 0)
        return col
    # selectors of degree zero: complex, or in no gate
    simple = []
    for s in sel_ids:
        if degrees[s] == 0:
            col = allocate()
            new_columns.append(list(activations[s]))
            substitution[s] = Fixed(col, 0)
        else:
            simple.append(s)
    rows_of = {s: {i for i, v in enumerate(activations[s]) if v} for s in simple}
    added = set()
    for i, s in enumerate(simple):
        if s in added:
            continue
        added.add(s)
        assert degrees[s] <= max_degree
        d = degrees[s] - 1
        combination = [s]
        for t in simple[i + 1:]:
            if d + len(combination) == max_degree:
                break
            if t in added or any(rows_of[t] & rows_of[u] for u in combination):
                continue
            new_d = max(d, degrees[t] - 1)
            if new_d + len(combination) + 1 > max_degree:
                continue
            d = new_d
            combination.append(t)
            added.add(t)
        col = allocate()
        values = [0] * n
        for root0, t in enumerate(combination):
            root = root0 + 1
            e = Fixed(col, 0)
            for r in range(1, len(combination) + 1):
                if r != root:
                    e = e * (Constant(r) - Fixed(col, 0))
            substitution[t] = e
            for row in rows_of[t]:
                values[row] = root
        new_columns.append(values)

    def rewrite(e):
        k = e.kind
        if k == "fixed":
            return substitution[e.a] if e.a in substitution else Fixed(remap[e.a], e.b)
        if k in ("const", "advice", "instance"):
            return e
        if k == "neg":
            return Expression("neg", rewrite(e.a))
        if k == "scaled":
            return Expression("scaled", rewrite(e.a), e.b)
        return Expression(k, rewrite(e.a), rewrite(e.b))
    out.gates = [(name, [rewrite(q) for q in polys]) for name, polys in cs.gates]
    out.gate_constraint_names = [list(x) for x in cs.gate_constraint_names]
    out.lookups = [(name, [rewrite(q) for q in inp], [rewrite(q) for q in tab]) for name, inp, tab in cs.lookups]
    assert out.degree() <= max_degree, "selector compression raised the degree"
    a2 = Assignment(out, asg.k)
    a2.fixed = [list(asg.fixed[c]) for c in real] + new_columns
    a2.advice, a2.instance = asg.advice, asg.instance
    a2.copies = [tuple((k, remap[c] if k == "fixed" else c, r) for k, c, r in pair) for pair in asg.copies]
    return out, a2
