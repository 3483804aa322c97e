"""ORACLE (test infrastructure, never shipped, never timed as the product).

Python big-integer restatement of the parts of `pasta_curves 0.4.1`
(/root/reference/Cargo.lock:567-579, crate NOT vendored) that the Halo2 IPA
prover hot path sits on:

  * Fp / Fq prime fields           (U: pasta_curves/src/fields/{fp,fq}.rs)
  * Pallas / Vesta  y^2 = x^3 + 5  (U: pasta_curves/src/curves.rs)
  * hash_to_curve (BLAKE2b-XMD, simplified SWU with Z = -13, 3-isogeny)
                                    (U: pasta_curves/src/hashtocurve.rs)

Pinned against the reference's own golden vectors (tests/test_oracle_kat.py):
  * GENERATOR of /root/reference/src/utils/constants/fixed_bases/board_commit_v.rs:5-14
    and board_commit_r.rs:5-14  (hash_to_curve KATs, reference tests :2940-2948)
  * the 85x8 U / 85 Z window tables at board_commit_{v,r}.rs:17-2919
    (scalar-mul + sqrt KATs, reference tests :2950-2960)
Vesta has no KAT anywhere in the reference: "parity unpinned" for Vesta
hash-to-curve (it only selects the URS points; every kernel output is
mathematically determined once the points are fixed).

Everything here is plain `int` arithmetic -- slow, exact, small cases only.
"""
import hashlib

# --------------------------------------------------------------------------
# Fields (moduli as published for the Pasta cycle)
# --------------------------------------------------------------------------
P = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001  # Pallas base = Vesta scalar
Q = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001  # Pallas scalar = Vesta base
S = 32                      # 2-adicity of both fields
GENERATOR = 5               # multiplicative generator used by pasta for both fields
R_MONT = 1 << 256           # Montgomery radix of pasta's 4x64 representation


class Field:
    """Constants of one of the two Pasta fields (U: pasta fields/fp.rs, fq.rs)."""

    def __init__(self, name, modulus, zeta):
        self.name = name
        self.p = modulus
        self.t = (modulus - 1) >> S
        self.root_of_unity = pow(GENERATOR, self.t, modulus)          # 2^32-th primitive root
        self.root_of_unity_inv = pow(self.root_of_unity, -1, modulus)
        self.delta = pow(GENERATOR, 1 << S, modulus)                  # generator of the t-order subgroup
        self.two_inv = pow(2, -1, modulus)
        self.zeta = zeta                                              # primitive cube root of unity
        assert pow(zeta, 3, modulus) == 1 and zeta != 1
        self.R = R_MONT % modulus
        self.R2 = (R_MONT * R_MONT) % modulus

    # Montgomery <-> canonical, for the C ABI's in-memory layout ([u64;4] LE, R = 2^256)
    def to_mont_bytes(self, x):
        return ((x * self.R) % self.p).to_bytes(32, "little")

    def from_mont_bytes(self, b):
        return (int.from_bytes(b, "little") * pow(self.R, -1, self.p)) % self.p

    def to_repr(self, x):
        return (x % self.p).to_bytes(32, "little")

    def from_repr(self, b):
        v = int.from_bytes(b, "little")
        assert v < self.p, "non-canonical field encoding"
        return v

    def from_bytes_wide(self, b64):
        """U: pasta `from_bytes_wide` / `from_u512`: 512-bit LE integer mod p."""
        assert len(b64) == 64
        return int.from_bytes(b64, "little") % self.p

    def inv(self, x):
        return pow(x, -1, self.p) if x % self.p else 0

    def sqrt(self, x):
        """Any square root of x or None (Tonelli-Shanks; sign is the caller's business)."""
        p = self.p
        x %= p
        if x == 0:
            return 0
        if pow(x, (p - 1) // 2, p) != 1:
            return None
        # p - 1 = t * 2^S
        z = self.root_of_unity           # a non-residue's t-th power: generator of the 2-Sylow
        m = S
        c = z
        tt = pow(x, self.t, p)
        r = pow(x, (self.t + 1) // 2, p)
        while tt != 1:
            i, t2 = 0, tt
            while t2 != 1:
                t2 = t2 * t2 % p
                i += 1
            b = pow(c, 1 << (m - i - 1), p)
            m = i
            c = b * b % p
            tt = tt * c % p
            r = r * b % p
        assert r * r % p == x
        return r


# ZETA as named by pasta_curves (fp.rs / fq.rs `const ZETA`).  Which of the two
# primitive cube roots is chosen only selects the evaluation coset of the
# extended domain; committed h(X) coefficients do not depend on it.
FP = Field("Fp", P, 0x12ccca834acdba712caad5dc57aab1b01d1f8bd237ad31491dad5ebdfdfe4ab9)
FQ = Field("Fq", Q, 0x06819a58283e528e511db4d81cf70f5a0fed467d47c033af2aa9d2e050aa0e4f)


# --------------------------------------------------------------------------
# Curves  y^2 = x^3 + 5, Jacobian (X/Z^2, Y/Z^3); identity = Z == 0
# (U: pasta_curves/src/curves.rs `new_curve_impl!`)
# --------------------------------------------------------------------------
class Curve:
    def __init__(self, name, base: Field, scalar: Field, iso_a, iso_b=1265):
        self.name = name            # "pallas" / "vesta" -- also the hash-to-curve curve_id
        self.base = base
        self.scalar = scalar
        self.b = 5
        self.iso_a = iso_a
        self.iso_b = iso_b
        self._iso = None

    # ---- affine helpers (None == identity) ----
    def is_on_curve(self, pt):
        if pt is None:
            return True
        x, y = pt
        p = self.base.p
        return (y * y - (x * x * x + self.b)) % p == 0

    def neg(self, pt):
        if pt is None:
            return None
        return (pt[0], (-pt[1]) % self.base.p)

    def add(self, a, b):
        """Affine addition (complete: handles identity, doubling, inverse)."""
        p = self.base.p
        if a is None:
            return b
        if b is None:
            return a
        x1, y1 = a
        x2, y2 = b
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        y3 = (lam * (x1 - x3) - y1) % p
        return (x3, y3)

    # ---- Jacobian arithmetic for the bulk work (tuples (X, Y, Z)) ----
    def j_identity(self):
        return (0, 1, 0)

    def to_jac(self, pt):
        return (0, 1, 0) if pt is None else (pt[0], pt[1], 1)

    def j_double(self, a):
        p = self.base.p
        X, Y, Z = a
        if Z == 0:
            return a
        A = X * X % p
        B = Y * Y % p
        C = B * B % p
        D = 2 * ((X + B) * (X + B) - A - C) % p
        E = 3 * A % p
        F = E * E % p
        X3 = (F - 2 * D) % p
        Y3 = (E * (D - X3) - 8 * C) % p
        Z3 = 2 * Y * Z % p
        return (X3, Y3, Z3)

    def j_add(self, a, b):
        p = self.base.p
        X1, Y1, Z1 = a
        X2, Y2, Z2 = b
        if Z1 == 0:
            return b
        if Z2 == 0:
            return a
        Z1Z1 = Z1 * Z1 % p
        Z2Z2 = Z2 * Z2 % p
        U1 = X1 * Z2Z2 % p
        U2 = X2 * Z1Z1 % p
        S1 = Y1 * Z2 * Z2Z2 % p
        S2 = Y2 * Z1 * Z1Z1 % p
        if U1 == U2:
            if S1 == S2:
                return self.j_double(a)
            return (0, 1, 0)
        H = (U2 - U1) % p
        I = 4 * H * H % p
        J = H * I % p
        r = 2 * (S2 - S1) % p
        V = U1 * I % p
        X3 = (r * r - J - 2 * V) % p
        Y3 = (r * (V - X3) - 2 * S1 * J) % p
        Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % p
        return (X3, Y3, Z3)

    def j_neg(self, a):
        return (a[0], (-a[1]) % self.base.p, a[2])

    def to_affine(self, a):
        p = self.base.p
        X, Y, Z = a
        if Z % p == 0:
            return None
        zi = pow(Z, -1, p)
        zi2 = zi * zi % p
        return (X * zi2 % p, Y * zi2 * zi % p)

    def j_eq(self, a, b):
        return self.to_affine(a) == self.to_affine(b)

    def j_mul(self, a, k):
        k %= self.scalar.p
        acc = (0, 1, 0)
        for bit in bin(k)[2:] if k else "":
            acc = self.j_double(acc)
            if bit == "1":
                acc = self.j_add(acc, a)
        return acc

    def mul(self, pt, k):
        return self.to_affine(self.j_mul(self.to_jac(pt), k))

    # ---- encodings ----
    def to_bytes(self, pt):
        """Compressed 32 B: x LE, bit 255 = parity of y; identity = zeros."""
        if pt is None:
            return bytes(32)
        x, y = pt
        b = bytearray(x.to_bytes(32, "little"))
        b[31] |= (y & 1) << 7
        return bytes(b)

    def from_bytes(self, b):
        if b == bytes(32):
            return None
        bb = bytearray(b)
        sign = bb[31] >> 7
        bb[31] &= 0x7F
        x = int.from_bytes(bb, "little")
        assert x < self.base.p
        y = self.base.sqrt((x * x * x + self.b) % self.base.p)
        assert y is not None, "not on curve"
        if (y & 1) != sign:
            y = self.base.p - y
        return (x, y)

    # ---- hash to curve (U: pasta_curves/src/hashtocurve.rs) ----
    def _isogeny(self):
        """Degree-3 isogeny iso-curve -> curve, derived (Velu) rather than recalled:
        kernel = the unique rational 3-torsion x-coordinate x0 of
        y^2 = x^3 + iso_a x + iso_b; codomain y^2 = x^3 + 5*3^6 is rescaled by
        (x/9, y/27) onto y^2 = x^3 + 5.  The normalisation is pinned on Pallas by the
        reference GENERATOR KATs (board_commit_{v,r}.rs:5-14)."""
        if self._iso is not None:
            return self._iso
        p = self.base.p
        a, b = self.iso_a, self.iso_b
        # psi_3(x) = 3x^4 + 6a x^2 + 12 b x - a^2 ; find its root in F_p via gcd(x^p - x, psi_3)
        psi = [(-a * a) % p, 12 * b % p, 6 * a % p, 0, 3]        # low -> high
        roots = _poly_roots_deg4(psi, p)
        good = []
        for x0 in roots:
            y0sq = (x0 * x0 * x0 + a * x0 + b) % p
            t = (6 * x0 * x0 + 2 * a) % p
            u = 4 * y0sq % p
            w = (u + x0 * t) % p
            A2 = (a - 5 * t) % p
            B2 = (b - 7 * w) % p
            if A2 == 0 and B2 == 5 * 729 % p:
                good.append((x0, t, u))
        assert len(good) == 1, "expected exactly one rational 3-isogeny onto y^2=x^3+5*3^6"
        self._iso = good[0]
        return self._iso

    def iso_map(self, pt):
        """Apply the 3-isogeny to an affine point of the iso-curve."""
        if pt is None:
            return None
        p = self.base.p
        x0, t, u = self._isogeny()
        x, y = pt
        d = (x - x0) % p
        if d == 0:
            return None
        di = pow(d, -1, p)
        di2 = di * di % p
        X = (x + t * di + u * di2) % p
        Y = y * (1 - t * di2 - 2 * u * di2 * di) % p
        return (X * pow(9, -1, p) % p, Y * pow(27, -1, p) % p)

    def _iso_add(self, A, B):
        """Affine addition on the iso-curve y^2 = x^3 + iso_a x + iso_b."""
        p = self.base.p
        if A is None:
            return B
        if B is None:
            return A
        x1, y1 = A
        x2, y2 = B
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = (3 * x1 * x1 + self.iso_a) * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        y3 = (lam * (x1 - x3) - y1) % p
        return (x3, y3)

    def _map_to_curve_simple_swu(self, u):
        """U: hashtocurve.rs `map_to_curve_simple_swu`, Z = -13, on the iso-curve; affine out."""
        F = self.base
        p = F.p
        a, b = self.iso_a, self.iso_b
        z = (-13) % p
        z_u2 = z * u * u % p
        ta = (z_u2 * z_u2 + z_u2) % p
        num_x1 = b * (ta + 1) % p
        div = a * (z if ta == 0 else (-ta) % p) % p
        x1 = num_x1 * pow(div, -1, p) % p
        gx1 = (x1 * x1 * x1 + a * x1 + b) % p
        y1 = F.sqrt(gx1)
        if y1 is not None:
            x, y = x1, y1
        else:
            x = z_u2 * x1 % p
            gx2 = (x * x * x + a * x + b) % p
            y = F.sqrt(gx2)
            assert y is not None
        if (u & 1) != (y & 1):       # sgn0(u) != sgn0(y)  -> negate
            y = (-y) % p
        return (x, y)

    def hash_to_field(self, domain_prefix: str, message: bytes):
        """U: hashtocurve.rs `hash_to_field` (BLAKE2b-512, 16 zero bytes personal, XMD-style)."""
        cid = self.name.encode()
        dom = domain_prefix.encode()
        assert len(dom) < 256 and 22 + len(cid) + len(dom) < 256
        tail = dom + b"-" + cid + b"_XMD:BLAKE2b_SSWU_RO_" + bytes([22 + len(cid) + len(dom)])

        def H(data):
            return hashlib.blake2b(data, digest_size=64, person=bytes(16)).digest()

        b0 = H(bytes(128) + message + bytes([0, 128, 0]) + tail)
        b1 = H(b0 + b"\x01" + tail)
        b2 = H(bytes(x ^ y for x, y in zip(b0, b1)) + b"\x02" + tail)
        return [int.from_bytes(bb, "big") % self.base.p for bb in (b1, b2)]

    def hash_to_curve(self, domain_prefix: str):
        """U: curves.rs `hash_to_curve`: returns closure message -> affine point."""
        def hasher(message: bytes):
            u0, u1 = self.hash_to_field(domain_prefix, message)
            q0 = self._map_to_curve_simple_swu(u0)
            q1 = self._map_to_curve_simple_swu(u1)
            r = self._iso_add(q0, q1)
            out = self.iso_map(r)
            assert self.is_on_curve(out)
            return out
        return hasher


def _poly_mulmod(a, b, m, p):
    """(a*b) mod m over F_p; polys low->high; m monic-normalised inside."""
    res = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                res[i + j] = (res[i + j] + x * y) % p
    return _poly_mod(res, m, p)


def _poly_mod(a, m, p):
    a = a[:]
    dm = len(m) - 1
    inv_lead = pow(m[-1], -1, p)
    while len(a) - 1 >= dm:
        c = a[-1] * inv_lead % p
        if c:
            for i in range(dm + 1):
                a[len(a) - 1 - dm + i] = (a[len(a) - 1 - dm + i] - c * m[i]) % p
        a.pop()
    while a and a[-1] == 0:
        a.pop()
    return a or [0]


def _poly_gcd(a, b, p):
    while b != [0]:
        a, b = b, _poly_mod(a, b, p)
    inv = pow(a[-1], -1, p)
    return [c * inv % p for c in a]


def _poly_roots_deg4(f, p):
    """All roots in F_p of a small-degree polynomial (distinct-degree split via x^p - x)."""
    # x^p mod f
    result, base, e = [1], [0, 1], p
    while e:
        if e & 1:
            result = _poly_mulmod(result, base, f, p)
        base = _poly_mulmod(base, base, f, p)
        e >>= 1
    xp_minus_x = result + [0] * max(0, 2 - len(result))
    xp_minus_x[1] = (xp_minus_x[1] - 1) % p
    while len(xp_minus_x) > 1 and xp_minus_x[-1] == 0:
        xp_minus_x.pop()
    g = _poly_gcd(f, xp_minus_x, p) if xp_minus_x != [0] else f
    # g splits completely over F_p; peel roots by equal-degree splitting
    return _split_roots(g, p)


def _split_roots(g, p, seed=1):
    if len(g) == 1:
        return []
    if len(g) == 2:
        return [(-g[0] * pow(g[1], -1, p)) % p]
    while True:
        # gcd(g, (x + seed)^((p-1)/2) - 1)
        result, base, e = [1], [seed % p, 1], (p - 1) // 2
        while e:
            if e & 1:
                result = _poly_mulmod(result, base, g, p)
            base = _poly_mulmod(base, base, g, p)
            e >>= 1
        result[0] = (result[0] - 1) % p
        while len(result) > 1 and result[-1] == 0:
            result.pop()
        h = _poly_gcd(g, result, p) if result != [0] else g
        seed += 1
        if 1 < len(h) < len(g):
            # g / h
            quo = _poly_div(g, h, p)
            return _split_roots(h, p, seed) + _split_roots(quo, p, seed)


def _poly_div(a, m, p):
    a = a[:]
    dm = len(m) - 1
    inv_lead = pow(m[-1], -1, p)
    q = [0] * (len(a) - dm)
    for k in range(len(a) - dm - 1, -1, -1):
        c = a[k + dm] * inv_lead % p
        q[k] = c
        for i in range(dm + 1):
            a[k + i] = (a[k + i] - c * m[i]) % p
    return q


# iso-curve coefficients (pasta curves.rs `IsoEp` / `IsoEq`)
PALLAS = Curve("pallas", FP, FQ, 0x18354a2eb0ea8c9c49be2d7258370742b74134581a27a59f92bb4b0b657a014b)
VESTA = Curve("vesta", FQ, FP, 0x267f9b2ee592271a81639c4d96f787739673928c7d01b212c515ad7242eaa6b1)
