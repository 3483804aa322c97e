"""Mirror of `halo2_proofs::dev::MockProver::verify` (U: halo2_proofs 0.2.0 src/dev.rs, src/dev/failure.rs, src/dev/util.rs)
over the circuit mirrors of this package: every gate, lookup and copy constraint is checked row by row and failures are
reported in the reference's own vocabulary -- `VerifyFailure::{CellNotAssigned, ConstraintNotSatisfied, Lookup, Permutation}`
with gate / constraint index and name, `FailureLocation::{InRegion, OutsideRegion}` and the formatted cell values -- so the
reference's negative tests (R:src/circuits/shot.rs:260-878, R:src/circuits/board.rs:164-877) can be replayed literally.
No cryptography; used by tests and by the synthetic-witness generators as a self-check."""

_ANY_ORDER = {"advice": 0, "fixed": 1, "instance": 2}
_ANY_NAME = {"advice": "Advice", "fixed": "Fixed", "instance": "Instance"}


def format_value(v, p):
    """dev/util.rs `format_value`: 0, 1, -1 by name, otherwise hex without leading zeros."""
    if v == 0:
        return "0"
    if v == 1:
        return "1"
    if v == p - 1:
        return "-1"
    return "0x%x" % v


class MockProver:
    def __init__(self, cs, asg):
        self.cs, self.asg = cs, asg
        self.regions = getattr(asg, "regions", [])
        self.n, self.p = asg.n, cs.modulus
        self.usable = asg.usable_rows

    # ---- evaluation ----
    def _eval(self, e, row):
        k, p = e.kind, self.p
        if k == "const":
            return e.a % p
        if k in ("fixed", "advice", "instance"):
            return self.asg.cell(k, e.a, (row + e.b) % self.n)
        if k == "neg":
            return (-self._eval(e.a, row)) % p
        if k == "sum":
            return (self._eval(e.a, row) + self._eval(e.b, row)) % p
        if k == "product":
            a = self._eval(e.a, row)
            return a * self._eval(e.b, row) % p if a else 0
        if k == "scaled":
            return self._eval(e.a, row) * e.b % p
        raise ValueError(k)

    def _columns_of(self, exprs):
        out = set()

        def walk(e):
            if e.kind in ("advice", "fixed", "instance"):
                out.add((e.kind, e.a))
            elif e.kind in ("neg", "scaled"):
                walk(e.a)
            elif e.kind in ("sum", "product"):
                walk(e.a); walk(e.b)
        for e in exprs:
            walk(e)
        return out

    def _find(self, row, columns):
        """FailureLocation::find: the first region whose rows contain `row` and whose columns meet `columns`."""
        for r in self.regions:
            if r.rows is not None and r.rows[0] <= row <= r.rows[1] and (columns & r.columns):
                return ("InRegion", (r.index, r.name), row - r.rows[0])
        return ("OutsideRegion", row)

    def _cell_values(self, gate_index, poly, row):
        cells = set(self.cs.gate_queried_cells(gate_index))
        found = {}

        def walk(e):
            if e.kind in ("advice", "fixed", "instance"):
                if (e.kind, e.a, e.b) in cells:
                    found[(_ANY_ORDER[e.kind], e.a, e.b)] = ((_ANY_NAME[e.kind], e.a), e.b, format_value(self.asg.cell(e.kind, e.a, (row + e.b) % self.n), self.p))
            elif e.kind in ("neg", "scaled"):
                walk(e.a)
            elif e.kind in ("sum", "product"):
                walk(e.a); walk(e.b)
        walk(poly)
        return [found[k] for k in sorted(found)]

    # ---- verify ----
    def verify(self):
        """-> [] (Ok(())) or the list of failures, ordered like the reference: unassigned cells, gates (gate, row, constraint),
        lookups, permutation (column, row)."""
        cs, errors = self.cs, []
        gate_sel = [cs.gate_queried_selectors(g) for g in range(len(cs.gates))]
        gate_cells = [cs.gate_queried_cells(g) for g in range(len(cs.gates))]
        # 1. every cell a switched-on gate queries must have been assigned inside the region that switched it on
        for r in self.regions:
            for sel, rows in r.enabled_selectors.items():
                for g, (gname, _) in enumerate(cs.gates):
                    if sel not in gate_sel[g]:
                        continue
                    for srow in rows:
                        for kind, col, rot in gate_cells[g]:
                            if kind == "instance":
                                continue
                            crow = (srow + rot) % self.n
                            if (kind, col, crow) not in r.cells:
                                errors.append(("CellNotAssigned", (g, gname), (r.index, r.name), ((_ANY_NAME[kind], col), rot), crow - r.start))
        # 2. gates
        for g, (gname, polys) in enumerate(cs.gates):
            names = cs.gate_constraint_names[g]
            cols = self._columns_of(polys)
            # rows where any selector of the gate is on (a gate whose selectors are all zero on a row evaluates to zero)
            sels = gate_sel[g]
            if sels:
                rows = sorted({row for row in range(self.usable) if any(self.asg.fixed[s][row] for s in sels)})
            else:
                rows = range(self.usable)
            for row in rows:
                for pi, poly in enumerate(polys):
                    if self._eval(poly, row) != 0:
                        errors.append(("ConstraintNotSatisfied", ((g, gname), pi, names[pi]), self._find(row, cols), self._cell_values(g, poly, row)))
        # 3. lookups (rows equal to the table's fill row are skipped, as upstream does)
        for li, (lname, inputs, tables) in enumerate(cs.lookups):
            fill = tuple(self._eval(t, self.usable - 1) for t in tables)
            table = {tuple(self._eval(t, row) for t in tables) for row in range(self.usable)}
            cols = self._columns_of(inputs)
            bad = []
            for row in range(self.usable):
                v = tuple(self._eval(i, row) for i in inputs)
                if v != fill and v not in table:
                    bad.append((v, row))
            for v, row in sorted(bad):
                errors.append(("Lookup", li, self._find(row, cols)))
        # 4. permutation: a cell fails when its value differs from the cell it is mapped to
        mapping = self.asg.permutation_mapping()
        for ci, (kind, col) in enumerate(cs.permutation):
            for row in range(self.n):
                c2, r2 = mapping[ci][row]
                if (c2, r2) == (ci, row):
                    continue
                k2, col2 = cs.permutation[c2]
                if self.asg.cell(kind, col, row) != self.asg.cell(k2, col2, r2):
                    errors.append(("Permutation", (_ANY_NAME[kind], col), self._find(row, {(kind, col)})))
        return errors

    def assert_satisfied(self):
        errs = self.verify()
        assert not errs, errs[:4]
