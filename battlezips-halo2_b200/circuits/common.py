"""Shared pieces of the Shot / Board circuit mirrors."""
import random
from ..plonk.circuit import ConstraintSystem, Assignment, Constant

P = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001   # pallas::Base
LOOKUP_BITS = 10            # R:src/utils/constants.rs:10  LOOKUP_SIZE
BOARD_SIZE = 100            # R:src/utils/board.rs:12


class Layout:
    """Sequential region allocator over an Assignment (stands in for SimpleFloorPlanner; placement does not
    change prover cost)."""

    def __init__(self, cs, k):
        self.cs = cs
        self.asg = Assignment(cs, k)
        self.row = 0
        self.const_row = 0

    def region(self, height):
        start = self.row
        self.row += height
        assert self.row <= self.asg.usable_rows - self.const_row, "Error::NotEnoughRowsAvailable"
        return start

    def constant(self, const_col, value):
        """A cell of the constants column holding `value` (enable_constant); returns ("fixed", col, row).
        Constants grow downward from the last usable row so they never collide with regions."""
        self.const_row += 1
        r = self.asg.usable_rows - self.const_row
        assert r >= self.row, "Error::NotEnoughRowsAvailable"
        self.asg.assign_fixed(const_col, r, value)
        return ("fixed", const_col, r)


def num2bits_configure(cs, bits, lc1, e2):
    """R:src/chips/bitify.rs:55-102 -- selector + 3-constraint running-sum gate."""
    sel = cs.fixed_column()
    s = cs.query_fixed(sel)
    bit = cs.query_advice(bits, 0)
    e2c, e2n = cs.query_advice(e2, 0), cs.query_advice(e2, 1)
    l1c, l1n = cs.query_advice(lc1, 0), cs.query_advice(lc1, 1)
    cs.create_gate("num2bits", [s * (bit * (Constant(1) - bit)),
                                s * (e2c + e2c - e2n),
                                s * (bit * e2c + l1c - l1n)])
    return {"bits": bits, "lc1": lc1, "e2": e2, "sel": sel}


def num2bits_synthesize(lay, cfg, const_col, value_cell, bits):
    """R:src/chips/bitify.rs:104-150: rows 0..B hold bit_i, lc1_i, e2_i = 2^i; lc1_0 = 0 and e2_0 = 1 come from
    the constants column; lc1_B is copy-constrained to the decomposed value.  Returns the bit cells."""
    a = lay.asg
    B = len(bits)
    r0 = lay.region(B + 1)
    lc1, e2 = 0, 1
    cells = []
    for i in range(B + 1):
        if i < B:
            a.assign_fixed(cfg["sel"], r0 + i, 1)
            a.assign_advice(cfg["bits"], r0 + i, bits[i])
            cells.append(("advice", cfg["bits"], r0 + i))
        a.assign_advice(cfg["lc1"], r0 + i, lc1)
        a.assign_advice(cfg["e2"], r0 + i, e2)
        if i < B:
            lc1 = (lc1 + bits[i] * e2) % P
            e2 = e2 * 2 % P
    a.copy(("advice", cfg["lc1"], r0), lay.constant(const_col, 0))
    a.copy(("advice", cfg["e2"], r0), lay.constant(const_col, 1))
    a.copy(("advice", cfg["lc1"], r0 + B), value_cell)
    return cells


# ---- shape-equivalent stand-in for halo2_gadgets' EccChip + LookupRangeCheckConfig ------------
def pallas_add(p1, p2):
    (x1, y1), (x2, y2) = p1, p2
    lam = (y2 - y1) * pow((x2 - x1) % P, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return x3, (lam * (x1 - x3) - y1) % P


def pallas_double(p1):
    x1, y1 = p1
    lam = 3 * x1 * x1 * pow(2 * y1 % P, -1, P) % P
    x3 = (lam * lam - 2 * x1) % P
    return x3, (lam * (x1 - x3) - y1) % P


def ecc_shape_configure(cs, advice, fixed, table):
    """19 gates + 1 lookup with the degree / rotation / constraint-count mix of the halo2_gadgets 0.2.0 gates
    listed in SURVEY App. C.  Gates whose selectors these circuits never enable (variable-base and short
    signed multiplication) are still part of h(X) -- exactly as in the reference -- with all-zero selectors."""
    A = lambda i, r=0: cs.query_advice(advice[i], r)
    F = lambda i, r=0: cs.query_fixed(fixed[i], r)
    h = {"advice": advice, "fixed": fixed, "table": table}
    q = {}

    def sel(name):
        q[name] = cs.fixed_column()
        return cs.query_fixed(q[name])

    # lookup range check: running sum of 10-bit words (LookupRangeCheckConfig::configure), degree-3 input
    q_lookup, q_running, q_bitshift = sel("q_lookup"), sel("q_running"), sel("q_bitshift")
    z_cur, z_next = A(9), A(9, 1)
    running = z_cur - z_next * (1 << LOOKUP_BITS)
    cs.lookup("lookup range check", [(q_lookup * (q_running * running + (Constant(1) - q_running) * z_cur),
                                      cs.query_fixed(table))])
    # (1) short lookup bitshift, deg 3
    cs.create_gate("Short lookup bitshift", [q_bitshift * (A(9, -1) * A(9, 1) - A(9))])
    # (2) witness point: 2 constraints deg 5; (3) non-identity point deg 4
    curve = lambda x, y: y * y - x * x * x - Constant(5)
    s = sel("q_point")
    cs.create_gate("witness point", [s * A(0) * curve(A(0), A(1)), s * A(1) * curve(A(0), A(1))])
    s = sel("q_point_non_id")
    cs.create_gate("witness non-identity point", [s * curve(A(0), A(1))])
    # (4) incomplete addition: real constraints, deg 4
    s = sel("q_add_incomplete")
    xp, yp, xq, yq, xr, yr = A(0), A(1), A(2), A(3), A(0, 1), A(1, 1)
    cs.create_gate("incomplete addition", [
        s * ((xr + xq + xp) * (xp - xq) * (xp - xq) - (yp - yq) * (yp - yq)),
        s * ((yr + yq) * (xp - xq) - (yp - yq) * (xq - xr))])
    # (5) complete addition: 12 constraints, deg <= 6 (stand-in products over the same 9 columns)
    s = sel("q_add")
    cs.create_gate("complete addition", [s * A(j % 9) * A((j + 1) % 9) * A((j + 2) % 9) * A((j + 3) % 9) * A((j + 4) % 9)
                                         for j in range(12)])
    # (6-14) variable-base mul family + overflow / LSB checks: configured, never enabled by these circuits
    for gi, (name, npoly, deg) in enumerate([("q_mul_1 == 1 checks (hi)", 3, 4), ("q_mul_2 == 1 checks (hi)", 5, 5),
                                             ("q_mul_3 == 1 checks (hi)", 4, 5), ("q_mul_1 == 1 checks (lo)", 3, 4),
                                             ("q_mul_2 == 1 checks (lo)", 5, 5), ("q_mul_3 == 1 checks (lo)", 4, 5),
                                             ("Decompose scalar for complete bits of variable-base mul", 2, 3),
                                             ("overflow checks", 4, 4), ("LSB check", 3, 4)]):
        s = sel(f"q_mul_{gi}")
        polys = []
        for j in range(npoly):
            e = s
            for t in range(deg - 1):
                e = e * A((j + t) % 9, (t % 3) - 1 if t < 3 else 0)
            polys.append(e)
        cs.create_gate(name, polys)
    # (15) running-sum range check: q * prod_{i<8}(word - i), word = z_cur - 8 z_next  -- degree 9
    s = sel("q_range_check")
    word = A(4) - A(4, 1) * 8
    rc = s
    for i in range(8):
        rc = rc * (word - Constant(i))
    cs.create_gate("range check", [rc])
    # (16) fixed-base coordinates check: q * (sum_j coeff_j(fixed) window^j - x) (deg 9) and q * (u^2 - y - z)
    s = sel("q_mul_fixed_running_sum")
    window = A(4) - A(4, 1) * 8
    interp, wpow = F(0), window
    for j in range(1, 8):
        interp = interp + F(j) * wpow
        if j < 7:
            wpow = wpow * window
    h["fixed_z"] = cs.fixed_column()
    cs.create_gate("Running sum coordinates check", [s * (interp - A(0)),
                                                     s * (A(5) * A(5) - A(1) - cs.query_fixed(h["fixed_z"]))])
    # (17) full-width fixed-base scalar mul: last-window range check, degree 9
    s = sel("q_mul_fixed_full")
    rc = s
    for i in range(8):
        rc = rc * (A(4) - Constant(i))
    cs.create_gate("Full-width fixed-base scalar mul", [rc, s * (A(5) * A(5) - A(1) - cs.query_fixed(h["fixed_z"]))])
    # (18) short fixed-base mul (never enabled), (19) canonicity checks
    s = sel("q_mul_fixed_short")
    cs.create_gate("Short fixed-base mul gate", [s * A(0) * A(1) * A(2), s * A(3) * (A(3) - Constant(1)) * A(4), s * A(5) * A(6)])
    s = sel("q_mul_fixed_base_field")
    cs.create_gate("Canonicity checks", [s * A(6) * (A(6) - Constant(1)), s * A(7) * (A(7) - Constant(1)) * A(8),
                                         s * (A(6) * A(7) - A(8, 1)), s * A(2) * A(3) * A(6) * A(7),
                                         s * (A(2) + A(3) * (1 << 10) - A(2, 1))])
    h["q"] = q
    return h


def ecc_shape_load_table(lay, h):
    """R:src/chips/pedersen.rs:71-85: table_idx = 0..1023."""
    for i in range(1 << LOOKUP_BITS):
        lay.asg.assign_fixed(h["table"], i, i)


def ecc_shape_synthesize(lay, h, scalar_a, scalar_b, rng):
    """Witness for the enabled stand-in gates: two 85-window fixed-base multiplications (R:src/chips/pedersen.rs
    :104-134 `[v]V + [r]R`), the 10-bit running-sum range check of one scalar, one incomplete and one complete
    addition.  Returns the (x, y) cells of the "commitment" (advice[0], advice[1] of the last row used)."""
    a, adv, fx, q = lay.asg, h["advice"], h["fixed"], h["q"]
    frng = random.Random(0x5EED)    # fixed columns (window tables, z) are circuit constants: witness independent
    G = (P - 1, 2)              # a Pallas point: (-1)^3 + 5 = 2^2
    out_cell = None
    acc = G
    for which, scalar in enumerate((scalar_a, scalar_b)):
        r0 = lay.region(86)
        windows = [(scalar >> (3 * i)) & 7 for i in range(85)]
        z = scalar % (1 << 255)
        for i in range(86):
            row = r0 + i
            a.assign_advice(adv[4], row, z)                 # running sum z_i
            if i < 85:
                w = windows[i]
                coeffs = [frng.randrange(P) for _ in range(8)]
                for j in range(8):
                    a.assign_fixed(fx[j], row, coeffs[j])
                x = sum(c * pow(w, j, P) for j, c in enumerate(coeffs)) % P
                zf = frng.randrange(1 << 20)
                u = rng.randrange(P)
                y = (u * u - zf) % P
                a.assign_fixed(h["fixed_z"], row, zf)
                a.assign_advice(adv[0], row, x)
                a.assign_advice(adv[1], row, y)
                a.assign_advice(adv[5], row, u)
                if i < 84:
                    a.assign_fixed(q["q_range_check"], row, 1)
                    a.assign_fixed(q["q_mul_fixed_running_sum"], row, 1)
                else:
                    a.assign_fixed(q["q_mul_fixed_full"], row, 1)     # last window: z_84 in 0..7 itself
                z >>= 3
        # 10-bit lookup range check of the low 250 bits of the scalar (26 rows)
    r0 = lay.region(27)
    z = scalar_a % (1 << 250)
    for i in range(26):
        a.assign_advice(adv[9], r0 + i, z)
        if i < 25:
            a.assign_fixed(q["q_lookup"], r0 + i, 1)
            a.assign_fixed(q["q_running"], r0 + i, 1)
        z >>= LOOKUP_BITS
    # incomplete addition P + Q = R (real curve arithmetic), then "complete addition" stand-in (zero row)
    Pt, Qt = pallas_double(G), pallas_add(pallas_double(G), G)
    Rt = pallas_add(Pt, Qt)
    r0 = lay.region(3)
    for col, v in zip((0, 1, 2, 3), (Pt[0], Pt[1], Qt[0], Qt[1])):
        a.assign_advice(adv[col], r0, v)
    a.assign_advice(adv[0], r0 + 1, Rt[0])
    a.assign_advice(adv[1], r0 + 1, Rt[1])
    a.assign_fixed(q["q_add_incomplete"], r0, 1)
    a.assign_fixed(q["q_point_non_id"], r0 + 1, 1)
    a.assign_fixed(q["q_point"], r0 + 1, 1)
    r1 = lay.region(2)
    a.assign_fixed(q["q_add"], r1, 1)                        # all-zero row satisfies the product stand-ins
    return ("advice", adv[0], r0 + 1), ("advice", adv[1], r0 + 1), Rt
