"""Host-side Pallas arithmetic for witness generation (what `pasta_curves` does for the reference's chips): affine points over
Fp, y^2 = x^3 + 5, None = identity.  Plain big-integer Python -- synthesis is host work in the drop-in as well
(R:src/chips/shot.rs:308-354 runs inside create_proof as a single-threaded callback)."""

P = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001      # pallas::Base
Q = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001      # pallas::Scalar
B = 5


def inv(a, m=P):
    return pow(a % m, -1, m) if a % m else 0          # ff's `invert().unwrap_or(0)` / Assigned inversion


def is_on_curve(pt):
    return pt is None or (pt[1] * pt[1] - pt[0] * pt[0] * pt[0] - B) % P == 0


def neg(pt):
    return None if pt is None else (pt[0], (-pt[1]) % P)


def add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * inv(2 * y1) % P
    else:
        lam = (y2 - y1) * inv(x2 - x1) % P
    x3 = (lam * lam - x1 - x2) % P
    return x3, (lam * (x1 - x3) - y1) % P


def mul(pt, k):
    k %= Q
    acc = None
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


_TS_S, _TS_T = 32, (P - 1) >> 32
_TS_Z = pow(5, _TS_T, P)               # 5 generates Fp^* (pasta's MULTIPLICATIVE_GENERATOR): a 2^32-th primitive root of unity


def sqrt(a):
    """One square root of a mod P, or None (Tonelli-Shanks; S = 32).  Callers pick the representative they need."""
    a %= P
    if a == 0:
        return 0
    if pow(a, (P - 1) // 2, P) != 1:
        return None
    m, c, t, r = _TS_S, _TS_Z, pow(a, _TS_T, P), pow(a, (_TS_T + 1) // 2, P)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % P
            i += 1
        b = pow(c, 1 << (m - i - 1), P)
        m, c = i, b * b % P
        t, r = t * c % P, r * b % P
    return r
