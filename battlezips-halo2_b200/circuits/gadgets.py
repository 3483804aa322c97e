"""Mirror of the halo2_gadgets 0.2.0 pieces the reference's Pedersen chip instantiates (R:src/chips/pedersen.rs:49-62,104-134;
pinned at R:Cargo.lock:363-379, not vendored): `LookupRangeCheckConfig<_, 10>` and `EccChip` -- all 19 gates in creation order
(U: halo2_gadgets src/utilities/lookup_range_check.rs, src/ecc/chip.rs, src/ecc/chip/{witness_point, add_incomplete, add, mul,
mul/{incomplete, complete, overflow}, mul_fixed, mul_fixed/{full_width, short, base_field_elem}}.rs,
src/utilities/decompose_running_sum.rs) -- and the witness generation of the three operations the commitment uses:
fixed-base multiplication by a base-field element, full-width fixed-base multiplication, complete addition.

Restated from the published algorithm; constraint names and the region structure follow upstream because the reference's
MockProver tests pin gate indices 21 / 23 (Shot) and 56 (Board) and region index 12 / 35 "complete point addition" behind
these 19 gates and 7 regions (R:src/circuits/shot.rs:303-324,684; R:src/circuits/board.rs:271,867).  The variable-base and
short-signed multiplication gates are configured exactly like upstream configures them and never switched on -- their
polynomials are still part of h(X), as in the reference."""
from ..plonk.circuit import Constant
from . import pallas as E
from . import fixed_bases as FB

P, Q = E.P, E.Q
K = 10                                    # R:src/utils/constants.rs:10  LOOKUP_SIZE
T_P = P - (1 << 254)
T_Q = Q - (1 << 254)


def range_check(word, rng):
    """word * (1 - word) * ... * (rng - 1 - word)"""
    acc = word
    for i in range(1, rng):
        acc = acc * (Constant(i) - word)
    return acc


def bool_check(v):
    return range_check(v, 2)


def ternary(a, b, c):
    return a * b + (Constant(1) - a) * c


# ---- utilities::lookup_range_check ------------------------------------------------------------------------------
class LookupRangeCheck:
    def __init__(self, cs, running_sum, table_idx):
        self.cs, self.running_sum, self.table_idx = cs, running_sum, table_idx
        cs.enable_equality("advice", running_sum)
        self.q_lookup = cs.complex_selector()
        self.q_running = cs.complex_selector()
        self.q_bitshift = cs.selector()
        q_lookup, q_running = cs.query_fixed(self.q_lookup), cs.query_fixed(self.q_running)
        z_cur, z_next = cs.query_advice(running_sum, 0), cs.query_advice(running_sum, 1)
        running_sum_lookup = q_running * (z_cur - z_next * (1 << K))
        short_lookup = (Constant(1) - q_running) * z_cur
        cs.lookup("lookup", [(q_lookup * (running_sum_lookup + short_lookup), cs.query_fixed(table_idx))])
        word, shifted, inv_two_pow_s = cs.query_advice(running_sum, -1), cs.query_advice(running_sum, 0), cs.query_advice(running_sum, 1)
        cs.create_gate("Short lookup bitshift", [("", cs.query_fixed(self.q_bitshift) * (word * (1 << K) * inv_two_pow_s - shifted))])

    def witness_check(self, lay, value, num_words, strict):
        """Range-constrain `value` to num_words * K bits: running sum z_0 = value ... z_num_words, every word looked up."""
        def body(region):
            z0 = region.assign_advice(self.running_sum, 0, value)
            return self.range_check(region, z0, num_words, strict)
        return lay.assign_region(f"Range check {num_words * K} bits", body)

    def range_check(self, region, element, num_words, strict):
        inv = E.inv(1 << K)
        zs, z = [element], element
        for idx in range(num_words):
            word = (element.value >> (K * idx)) & ((1 << K) - 1)
            region.enable_selector(self.q_lookup, idx)
            region.enable_selector(self.q_running, idx)
            z = region.assign_advice(self.running_sum, idx + 1, (z.value - word) * inv)
            zs.append(z)
        if strict:
            region.constrain_constant(zs[-1], 0)
        return zs


# ---- utilities::decompose_running_sum (3-bit windows) -----------------------------------------------------------
class RunningSum:
    def __init__(self, cs, q_range_check, z):
        self.q_range_check, self.z = q_range_check, z
        cs.enable_equality("advice", z)
        word = cs.query_advice(z, 0) - cs.query_advice(z, 1) * FB.H
        cs.create_gate("range check", [("", cs.query_fixed(q_range_check) * range_check(word, FB.H))])

    def copy_decompose(self, region, offset, alpha, strict, num_windows):
        z0 = region.copy_advice(alpha, self.z, offset)
        for idx in range(num_windows):
            region.enable_selector(self.q_range_check, offset + idx)
        inv = E.inv(FB.H)
        zs, z = [z0], z0
        for i in range(num_windows):
            word = (alpha.value >> (FB.WINDOW_BITS * i)) & (FB.H - 1)
            z = region.assign_advice(self.z, offset + i + 1, (z.value - word) * inv)
            zs.append(z)
        if strict:
            region.constrain_constant(zs[-1], 0)
        return zs


# ---- ecc::chip ----------------------------------------------------------------------------------------------------
class EccChip:
    def __init__(self, cs, advices, lagrange_coeffs, range_check_cfg):
        assert len(advices) == 10 and len(lagrange_coeffs) == 8
        self.cs, self.adv, self.lookup = cs, advices, range_check_cfg
        A = lambda i, r=0: cs.query_advice(advices[i], r)
        F = cs.query_fixed
        one = Constant(1)
        b = Constant(E.B)
        # -- witness_point
        self.q_point, self.q_point_non_id = cs.selector(), cs.selector()
        curve_eqn = lambda: A(1).square() - A(0).square() * A(0) - b
        q = F(self.q_point)
        cs.create_gate("witness point", [("x == 0 v on_curve", q * (A(0) * curve_eqn())), ("y == 0 v on_curve", q * (A(1) * curve_eqn()))])
        cs.create_gate("witness non-identity point", [("on_curve", F(self.q_point_non_id) * curve_eqn())])
        # -- add_incomplete: x_p, y_p, x_qr, y_qr = advices[0..4]
        self.q_add_incomplete = cs.selector()
        q = F(self.q_add_incomplete)
        x_p, y_p, x_q, y_q, x_r, y_r = A(0), A(1), A(2), A(3), A(2, 1), A(3, 1)
        cs.create_gate("incomplete addition", [
            ("x_r", q * ((x_r + x_q + x_p) * (x_p - x_q) * (x_p - x_q) - (y_p - y_q).square())),
            ("y_r", q * ((y_r + y_q) * (x_p - x_q) - (y_p - y_q) * (x_q - x_r)))])
        # -- add (complete): lambda, alpha, beta, gamma, delta = advices[4..9]
        self.q_add = cs.selector()
        q = F(self.q_add)
        lam, alpha, beta, gamma, delta = A(4), A(5), A(6), A(7), A(8)
        x_q_minus_x_p, y_q_plus_y_p = x_q - x_p, y_q + y_p
        if_alpha, if_beta, if_gamma, if_delta = x_q_minus_x_p * alpha, x_p * beta, x_q * gamma, y_q_plus_y_p * delta
        nonexc_x = lam.square() - x_p - x_q - x_r
        nonexc_y = lam * (x_p - x_r) - y_p - y_r
        cs.create_gate("complete addition", [
            ("1", q * (x_q_minus_x_p * (x_q_minus_x_p * lam - (y_q - y_p)))),
            ("2", q * ((one - if_alpha) * (y_p * 2 * lam - x_p.square() * 3))),
            ("3a", q * (x_p * x_q * x_q_minus_x_p * nonexc_x)),
            ("3b", q * (x_p * x_q * x_q_minus_x_p * nonexc_y)),
            ("3c", q * (x_p * x_q * y_q_plus_y_p * nonexc_x)),
            ("3d", q * (x_p * x_q * y_q_plus_y_p * nonexc_y)),
            ("4a", q * ((one - if_beta) * (x_r - x_q))),
            ("4b", q * ((one - if_beta) * (y_r - y_q))),
            ("5a", q * ((one - if_gamma) * (x_r - x_p))),
            ("5b", q * ((one - if_gamma) * (y_r - y_p))),
            ("6a", q * ((one - if_alpha - if_delta) * x_r)),
            ("6b", q * ((one - if_alpha - if_delta) * y_r))])
        # -- mul (variable base): configured, never enabled by the board commitment
        self._configure_mul(A, F, one)
        # -- mul_fixed shared config: window = advices[4], u = advices[5]
        self.lagrange_coeffs = lagrange_coeffs
        self.window, self.u = advices[4], advices[5]
        self.q_running_sum = cs.selector()
        self.running_sum = RunningSum(cs, self.q_running_sum, self.window)
        self.fixed_z = cs.fixed_column()
        word = A(4) - A(4, 1) * FB.H
        cs.create_gate("Running sum coordinates check", [(n, F(self.q_running_sum) * e) for n, e in self._coords_check(word)])
        # -- full-width
        self.q_mul_fixed_full = cs.selector()
        q = F(self.q_mul_fixed_full)
        window = A(4)
        cs.create_gate("Full-width fixed-base scalar mul", [(n, q * e) for n, e in self._coords_check(window)] + [("window range check", q * range_check(window, FB.H))])
        # -- short signed (never enabled)
        self.q_mul_fixed_short = cs.selector()
        q = F(self.q_mul_fixed_short)
        y_a, last_window, sign = A(3), A(5), A(4)
        cs.create_gate("Short fixed-base mul gate", [
            ("last_window_check", q * bool_check(last_window)),
            ("sign_check", q * (sign.square() - one)),
            ("y_check", q * ((A(1) - y_a) * (A(1) + y_a))),
            ("negation_check", q * (sign * A(1) - y_a))])
        # -- base-field element: canon_advices = advices[6..9]
        self.canon = advices[6:9]
        self.q_mul_fixed_base_field = cs.selector()
        q = F(self.q_mul_fixed_base_field)
        C = lambda i, r=0: cs.query_advice(self.canon[i], r)
        a_alpha, z_84_alpha, alpha_0 = C(0, -1), C(2, -1), C(1, -1)
        alpha_1, alpha_2 = C(1, 0), C(1, 1)
        alpha_0_prime, z_13_alpha_0_prime = C(0, 0), C(0, 1)
        z_44_alpha, z_43_alpha = C(2, 0), C(2, 1)
        alpha_0_hi_120 = z_44_alpha - z_84_alpha * (1 << 120)
        a_43 = z_43_alpha - z_44_alpha * FB.H
        cs.create_gate("Canonicity checks", [
            ("MSB = 1 => alpha_1 = 0", q * (alpha_2 * alpha_1)),
            ("MSB = 1 => alpha_0_hi_120 = 0", q * (alpha_2 * alpha_0_hi_120)),
            ("MSB = 1 => a_43 = 0 or 1", q * (alpha_2 * bool_check(a_43))),
            ("MSB = 1 => z_13_alpha_0_prime = 0", q * (alpha_2 * z_13_alpha_0_prime)),
            ("alpha_1_range_check", q * range_check(alpha_1, 1 << 2)),
            ("alpha_2_range_check", q * bool_check(alpha_2)),
            ("z_84_alpha_check", q * (z_84_alpha - (alpha_1 + alpha_2 * (1 << 2)))),
            ("alpha_0_check", q * (alpha_0 - (a_alpha - z_84_alpha * (1 << 252)))),
            ("alpha_0_prime check", q * (alpha_0_prime - (alpha_0 + Constant(1 << 130) - Constant(T_P))))])

    def _coords_check(self, window):
        cs = self.cs
        y_p, x_p = cs.query_advice(self.adv[1], 0), cs.query_advice(self.adv[0], 0)
        z, u = cs.query_fixed(self.fixed_z, 0), cs.query_advice(self.u, 0)
        interpolated_x = Constant(0)
        for power in range(FB.H):
            wp = Constant(1)
            for _ in range(power):
                wp = wp * window
            interpolated_x = interpolated_x + wp * cs.query_fixed(self.lagrange_coeffs[power], 0)
        return [("check x", interpolated_x - x_p), ("check y", u.square() - y_p - z),
                ("on-curve", y_p.square() - x_p.square() * x_p - Constant(E.B))]

    def _configure_mul(self, A, F, one):
        cs = self.cs
        # incomplete addition halves: (z, x_a, x_p, y_p, lambda1, lambda2)
        hi = (9, 3, 0, 1, 4, 5)
        lo = (6, 7, 0, 1, 8, 2)
        self.q_mul = []
        for z, x_a, x_p, y_p, l1, l2 in (hi, lo):
            q1, q2, q3 = cs.selector(), cs.selector(), cs.selector()
            self.q_mul.append((q1, q2, q3))
            x_r = lambda rot: A(l1, rot).square() - A(x_a, rot) - A(x_p, rot)
            Y_A = lambda rot: (A(l1, rot) + A(l2, rot)) * (A(x_a, rot) - x_r(rot))

            def for_loop(y_a_next):
                k = A(z, 0) - A(z, -1) * 2
                Y_A_cur = Y_A(0)
                gradient_1 = A(l1, 0) * 2 * (A(x_a, 0) - A(x_p, 0)) - Y_A_cur + (k * 2 - one) * A(y_p, 0) * 2
                secant_line = A(l2, 0).square() - A(x_a, 1) - x_r(0) - A(x_a, 0)
                gradient_2 = A(l2, 0) * 2 * (A(x_a, 0) - A(x_a, 1)) - Y_A_cur - y_a_next
                return [("bool_check", bool_check(k)), ("gradient_1", gradient_1), ("secant_line", secant_line), ("gradient_2", gradient_2)]
            cs.create_gate("q_mul_1 == 1 checks", [("init Y_A", F(q1) * (A(l1, 0) * 2 - Y_A(1)))])
            cs.create_gate("q_mul_2 == 1 checks", [(n, F(q2) * e) for n, e in
                           [("x_p_check", A(x_p, 0) - A(x_p, 1)), ("y_p_check", A(y_p, 0) - A(y_p, 1))] + for_loop(Y_A(1))])
            cs.create_gate("q_mul_3 == 1 checks", [(n, F(q3) * e) for n, e in for_loop(A(l1, 1) * 2)])
        # complete bits: z_complete = advices[9]
        self.q_mul_decompose_var = cs.selector()
        q = F(self.q_mul_decompose_var)
        k = A(9, 1) - A(9, -1) * 2
        base_y, y_p = A(9, 0), A(1, -1)
        cs.create_gate("Decompose scalar for complete bits of variable-base mul",
                       [("bool_check", q * bool_check(k)), ("y_switch", q * ternary(k, base_y - y_p, base_y + y_p))])
        # overflow: advices[6..9]
        self.q_mul_overflow = cs.selector()
        q = F(self.q_mul_overflow)
        z_0, z_130, eta = A(6, -1), A(6, 0), A(6, 1)
        k_254, alpha, s_minus_lo_130 = A(7, -1), A(7, 0), A(7, 1)
        s = A(8, 0)
        cs.create_gate("overflow checks", [
            ("s_check", q * (s - (alpha + k_254 * (1 << 130)))),
            ("recovery", q * (z_0 - alpha - Constant(T_Q))),
            ("lo_zero", q * (k_254 * (z_130 - Constant(1 << 124)))),
            ("s_minus_lo_130_check", q * (k_254 * s_minus_lo_130)),
            ("canonicity", q * ((one - k_254) * (one - z_130 * eta) * s_minus_lo_130))])
        # LSB
        self.q_mul_lsb = cs.selector()
        q = F(self.q_mul_lsb)
        lsb = A(9, 1) - A(9, 0) * 2
        cs.create_gate("LSB check", [
            ("bool_check", q * bool_check(lsb)),
            ("lsb_x", q * (lsb * A(0, 0) + (one - lsb) * (A(0, 0) - A(0, 1)))),
            ("lsb_y", q * (lsb * A(1, 0) + (one - lsb) * (A(1, 0) + A(1, 1))))])

    # ---- synthesis ----
    def _assign_fixed_constants(self, region, offset, base, toggle):
        for w in range(FB.NUM_WINDOWS):
            region.enable_selector(toggle, w + offset)
            for k in range(FB.H):
                region.assign_fixed(self.lagrange_coeffs[k], w + offset, base.lagrange_coeffs[w][k])
            region.assign_fixed(self.fixed_z, w + offset, base.z[w])

    def _process_window(self, region, offset, w, k, base):
        pt = base.table[w][k]
        x = region.assign_advice(self.adv[0], offset + w, pt[0])
        y = region.assign_advice(self.adv[1], offset + w, pt[1])
        region.assign_advice(self.u, offset + w, base.u[w][k])
        return (x, y, pt)

    def _add_incomplete_region(self, region, p, q, offset):
        region.enable_selector(self.q_add_incomplete, offset)
        assert p[2][0] != q[2][0], "Error::Synthesis (exceptional case of incomplete addition)"
        region.copy_advice(p[0], self.adv[0], offset); region.copy_advice(p[1], self.adv[1], offset)
        region.copy_advice(q[0], self.adv[2], offset); region.copy_advice(q[1], self.adv[3], offset)
        r = E.add(p[2], q[2])
        return (region.assign_advice(self.adv[2], offset + 1, r[0]), region.assign_advice(self.adv[3], offset + 1, r[1]), r)

    def _assign_region_inner(self, region, offset, windows, base, toggle):
        self._assign_fixed_constants(region, offset, base, toggle)
        acc = self._process_window(region, offset, 0, windows[0], base)
        for w in range(1, FB.NUM_WINDOWS - 1):
            mul_b = self._process_window(region, offset, w, windows[w], base)
            acc = self._add_incomplete_region(region, mul_b, acc, offset + w)
        mul_b = self._process_window(region, offset, FB.NUM_WINDOWS - 1, windows[-1], base)
        return acc, mul_b

    def add_region(self, region, p, q, offset):
        """add::Config::assign_region: complete addition of two (possibly identity) points; p, q = (x cell, y cell, point)."""
        region.enable_selector(self.q_add, offset)
        region.copy_advice(p[0], self.adv[0], offset); region.copy_advice(p[1], self.adv[1], offset)
        region.copy_advice(q[0], self.adv[2], offset); region.copy_advice(q[1], self.adv[3], offset)
        (x_p, y_p), (x_q, y_q) = (p[0].value, p[1].value), (q[0].value, q[1].value)
        region.assign_advice(self.adv[5], offset, E.inv(x_q - x_p))
        region.assign_advice(self.adv[6], offset, E.inv(x_p))
        region.assign_advice(self.adv[7], offset, E.inv(x_q))
        region.assign_advice(self.adv[8], offset, E.inv(y_q + y_p) if x_q == x_p else 0)
        if x_q != x_p:
            lam = (y_q - y_p) * E.inv(x_q - x_p) % P
        elif y_p != 0:
            lam = 3 * x_p * x_p * E.inv(2 * y_p) % P
        else:
            lam = 0
        region.assign_advice(self.adv[4], offset, lam)
        if x_p == 0:
            r = (x_q, y_q)
        elif x_q == 0:
            r = (x_p, y_p)
        elif x_q == x_p and (y_q + y_p) % P == 0:
            r = (0, 0)
        else:
            x_r = (lam * lam - x_p - x_q) % P
            r = (x_r, (lam * (x_p - x_r) - y_p) % P)
        pt = None if r == (0, 0) else r
        return (region.assign_advice(self.adv[2], offset + 1, r[0]), region.assign_advice(self.adv[3], offset + 1, r[1]), pt)

    def mul_fixed_base_field_elem(self, lay, scalar_cell, base):
        """FixedPointBaseField::mul: [alpha] B for a base-field element alpha (4 regions: running-sum windows + incomplete
        additions, complete addition of the last window, 130-bit range check of alpha_0', canonicity rows)."""
        alpha = scalar_cell.value

        def incomplete(region):
            zs = self.running_sum.copy_decompose(region, 0, scalar_cell, True, FB.NUM_WINDOWS)
            windows = [(alpha >> (FB.WINDOW_BITS * i)) & (FB.H - 1) for i in range(FB.NUM_WINDOWS)]
            acc, mul_b = self._assign_region_inner(region, 0, windows, base, self.q_running_sum)
            return zs, acc, mul_b
        zs, acc, mul_b = lay.assign_region("Base-field elem fixed-base mul (incomplete addition)", incomplete)
        result = lay.assign_region("Base-field elem fixed-base mul (complete addition)", lambda region: self.add_region(region, mul_b, acc, 0))
        assert result[2] == base.mul(alpha % Q)
        z_43, z_44, z_84 = zs[43], zs[44], zs[84]
        alpha_0 = (alpha - z_84.value * (1 << 252)) % P
        alpha_0_prime = (alpha_0 + (1 << 130) - T_P) % P
        rc = self.lookup.witness_check(lay, alpha_0_prime, 13, False)

        def canonicity(region):
            region.enable_selector(self.q_mul_fixed_base_field, 1)
            region.copy_advice(scalar_cell, self.canon[0], 0)
            region.assign_advice(self.canon[1], 0, alpha_0)
            region.copy_advice(z_84, self.canon[2], 0)
            region.copy_advice(rc[0], self.canon[0], 1)
            region.assign_advice(self.canon[1], 1, (alpha >> 252) & 3)
            region.copy_advice(z_44, self.canon[2], 1)
            region.copy_advice(rc[13], self.canon[0], 2)
            region.assign_advice(self.canon[1], 2, (alpha >> 254) & 1)
            region.copy_advice(z_43, self.canon[2], 2)
        lay.assign_region("Canonicity checks", canonicity)
        return result

    def mul_fixed_full_width(self, lay, scalar, base):
        """FixedPoint::mul with a full-width scalar (an element of pallas::Scalar, witnessed lazily as 85 windows)."""
        scalar %= Q

        def incomplete(region):
            windows = [(scalar >> (FB.WINDOW_BITS * i)) & (FB.H - 1) for i in range(FB.NUM_WINDOWS)]
            for idx in range(FB.NUM_WINDOWS):
                region.enable_selector(self.q_mul_fixed_full, idx)
            for idx, wv in enumerate(windows):
                region.assign_advice(self.window, idx, wv)
            return self._assign_region_inner(region, 0, windows, base, self.q_mul_fixed_full)
        acc, mul_b = lay.assign_region("Full-width fixed-base mul (incomplete addition)", incomplete)
        result = lay.assign_region("Full-width fixed-base mul (last window, complete addition)", lambda region: self.add_region(region, mul_b, acc, 0))
        assert result[2] == base.mul(scalar)
        return result

    def add(self, lay, a, b):
        """Point::add"""
        return lay.assign_region("complete point addition", lambda region: self.add_region(region, a, b, 0))


class PedersenCommitmentChip:
    """R:src/chips/pedersen.rs:44-134."""

    def __init__(self, cs, advice, lagrange, table_idx):
        self.table_idx = table_idx
        self.range_check = LookupRangeCheck(cs, advice[9], table_idx)            # R:src/chips/pedersen.rs:56-57
        self.ecc = EccChip(cs, advice, lagrange, self.range_check)               # R:src/chips/pedersen.rs:59

    def synthesize(self, lay, value_cell, trapdoor, load_table=True):
        if load_table:                                                           # R:src/chips/pedersen.rs:71-85
            lay.assign_table("table_idx", self.table_idx, list(range(1 << K)))
        commitment = self.ecc.mul_fixed_base_field_elem(lay, value_cell, FB.board_commit_v())       # [v] BoardCommitV
        blind = self.ecc.mul_fixed_full_width(lay, trapdoor, FB.board_commit_r())                   # [rcv] BoardCommitR
        return self.ecc.add(lay, commitment, blind)                                                 # "cv"
