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
        self.lookups = []          # (name, [input Expression], [table Expression])
        self.permutation = []      # [("advice"|"fixed"|"instance", index)] in enable_equality order

    def advice_column(self):
        self.num_advice += 1
        return self.num_advice - 1

    def fixed_column(self):
        self.num_fixed += 1
        return self.num_fixed - 1

    def instance_column(self):
        self.num_instance += 1
        return self.num_instance - 1

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
        assert polys, "gates must contain at least one constraint"
        self.gates.append((name, list(polys)))

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
