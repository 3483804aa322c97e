"""GPU parity (through the C ABI): field / curve primitives, best_multiexp, best_fft and the EvaluationDomain
transforms against the oracle -- bit-exact (integer work)."""
import random
import numpy as np
import pytest
import battlezips_halo2_b200 as bz
from battlezips_halo2_b200 import arithmetic as ar
from oracle import pasta

pytestmark = pytest.mark.gpu


def _edge_and_random(F, rnd, n):
    edge = [0, 1, 2, F.p - 1, F.p - 2, (1 << 254), (1 << 254) - 1, F.R % F.p, (F.p - 1) // 2, 0xFFFFFFFF, 1 << 32, (1 << 128) - 1]
    return edge + [rnd.randrange(F.p) for _ in range(n - len(edge))]


@pytest.mark.parametrize("field", [0, 1])
def test_field_ops(field, ctx, oracle_c):
    co = oracle_c
    F = co.FIELDS[field]
    rnd = random.Random(100 + field)
    n = 600
    a = _edge_and_random(F, rnd, n)
    b = list(reversed(_edge_and_random(F, rnd, n)))
    am, bm = co.to_mont(field, a), co.to_mont(field, b)
    p = F.p
    assert co.from_mont(field, ar.field_op(ctx, field, "mul", am, bm)) == [x * y % p for x, y in zip(a, b)]
    assert co.from_mont(field, ar.field_op(ctx, field, "add", am, bm)) == [(x + y) % p for x, y in zip(a, b)]
    assert co.from_mont(field, ar.field_op(ctx, field, "sub", am, bm)) == [(x - y) % p for x, y in zip(a, b)]
    assert co.from_mont(field, ar.field_op(ctx, field, "neg", am)) == [(-x) % p for x in a]
    assert co.from_mont(field, ar.field_op(ctx, field, "sqr", am)) == [x * x % p for x in a]
    # ff::BatchInvert semantics (zeros stay zero); 600 elements = 37 full runs of the Montgomery trick + a ragged tail
    assert co.from_mont(field, ar.field_op(ctx, field, "inv", am)) == [F.inv(x) for x in a]
    assert co.from_mont(field, ar.field_op(ctx, field, "inv", am[:1])) == [F.inv(a[0])]
    # Montgomery conversions are bit-exact with pasta's in-memory form
    raw = co.ints_to_raw(a)
    assert np.array_equal(ar.field_op(ctx, field, "to_mont", raw), am)
    assert np.array_equal(ar.field_op(ctx, field, "from_mont", am), raw)
    wide = np.frombuffer(rnd.randbytes(64 * 200), dtype=np.uint64).reshape(-1, 8).copy()
    wide[0] = 0xFFFFFFFFFFFFFFFF
    wide[1] = 0
    assert np.array_equal(ar.field_op(ctx, field, "from_u512", wide), co.from_u512(field, wide))


@pytest.mark.parametrize("field", [0, 1])
def test_binary_gcd_inverse_on_device(field, ctx, oracle_c):
    co = oracle_c
    F = co.FIELDS[field]
    rnd = random.Random(55 + field)
    a = _edge_and_random(F, rnd, 700) + [1 << 32, 1 << 64, (1 << 254) % F.p]
    am = co.to_mont(field, a)
    assert co.from_mont(field, ar.field_op(ctx, field, "inv_gcd", am)) == [F.inv(x) for x in a]
    assert np.array_equal(ar.field_op(ctx, field, "inv_gcd", am), ar.field_op(ctx, field, "inv", am))


@pytest.mark.parametrize("field", [0, 1])
def test_batch_invert_assigned(field, ctx, oracle_c):
    """Assigned<F> columns (numerator / denominator) -> F, against the restated poly::batch_invert_assigned
    and against plain big-integer division; Zero, Trivial, zero-denominator and ragged-tail cells included."""
    from oracle import halo2 as H
    from oracle.vec import Vec
    co = oracle_c
    F = co.FIELDS[field]
    p = F.p
    rnd = random.Random(7 + field)
    for n in (1, 7, 8, 9, 2048 + 3):
        num = [rnd.randrange(p) for _ in range(n)]
        den = [rnd.choice([1, 1, 1, 0, 2, p - 1, rnd.randrange(p)]) for _ in range(n)]
        num[0], den[0] = 0, 1                                   # Assigned::Zero
        nm, dm = co.to_mont(field, num), co.to_mont(field, den)
        got = ar.batch_invert_assigned(ctx, field, nm, dm)
        assert np.array_equal(got, H.batch_invert_assigned(Vec(field), nm, dm))
        assert co.from_mont(field, got) == [x * F.inv(d) % p for x, d in zip(num, den)]
    trivial = co.to_mont(field, [rnd.randrange(p) for _ in range(64)])          # a column of Trivial cells comes back unchanged
    assert np.array_equal(ar.batch_invert_assigned(ctx, field, trivial, co.to_mont(field, [1] * 64)), trivial)


@pytest.mark.parametrize("curve", [0, 1])
def test_curve_ops(curve, ctx, oracle_c):
    co = oracle_c
    C, sf, bf = co.CURVES[curve]
    h = C.hash_to_curve("bz-test")
    pts = [h(bytes([i])) for i in range(24)]
    A = pts[:12] + [None, pts[0], pts[1], None, pts[2]]
    B = pts[12:] + [pts[3], None, pts[1], None, C.neg(pts[2])]
    am, bm = co.points_to_mont(curve, A), co.points_to_mont(curve, B)
    assert co.points_from_mont(curve, ar.curve_op(ctx, curve, "add", am, bm)) == [C.add(x, y) for x, y in zip(A, B)]
    assert co.points_from_mont(curve, ar.curve_op(ctx, curve, "sub", am, bm)) == [C.add(x, C.neg(y)) for x, y in zip(A, B)]
    assert co.points_from_mont(curve, ar.curve_op(ctx, curve, "double", am, bm)) == [C.add(x, x) for x in A]
    assert co.points_from_mont(curve, ar.curve_op(ctx, curve, "double_add", am, bm)) == [C.add(C.add(x, x), y) for x, y in zip(A, B)]
    ks = [0, 1, 2, 3, 0xFFFFFFFF, 12345, 1 << 31] + list(range(7, 17))
    km = np.zeros((len(A), 8), dtype=np.uint64)
    for i, k in enumerate(ks):
        km[i, 0] = k
    assert co.points_from_mont(curve, ar.curve_op(ctx, curve, "mul_u32", am, km)) == [C.mul(x, k) for x, k in zip(A, ks)]


def _bases(co, curve, n):
    """n distinct points: hash-to-curve for a few, then a doubling/adding chain through the C oracle."""
    C, sf, bf = co.CURVES[curve]
    if n == 0:
        return []
    h = C.hash_to_curve("Halo2-Parameters")
    seed = [h(b"\x00" + i.to_bytes(4, "little")) for i in range(min(n, 8))]
    pts = list(seed)
    acc = seed[0]
    while len(pts) < n:
        acc = C.add(C.add(acc, acc), seed[len(pts) % len(seed)])
        pts.append(acc)
    return pts[:n]


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("n", [0, 1, 2, 33, 257, 2049])
def test_best_multiexp_vs_oracle(curve, n, ctx, oracle_c):
    co = oracle_c
    C, sf, bf = co.CURVES[curve]
    rnd = random.Random(n * 2 + curve)
    pts = _bases(co, curve, n)
    sc = [rnd.randrange(C.scalar.p) for _ in range(n)]
    if n >= 33:
        sc[0], sc[1], sc[2], sc[3] = 0, 1, C.scalar.p - 1, 1 << 254
        pts[5] = None                       # identity base
        pts[7] = pts[6]                     # duplicated base
        pts[9] = C.neg(pts[8]); sc[9] = sc[8]   # base equal to -another, same scalar
    sm, bm = co.to_mont(sf, sc) if n else np.zeros((0, 4), np.uint64), co.points_to_mont(curve, pts) if n else np.zeros((0, 8), np.uint64)
    got = ar.best_multiexp(ctx, curve, sm, bm)
    exp = co.best_multiexp(curve, sm, bm)
    assert co.points_from_mont(curve, co.to_affine(curve, got)) == co.points_from_mont(curve, co.to_affine(curve, exp))


@pytest.mark.parametrize("kind", ["zeros", "ones", "minus_one", "single", "witness_like"])
def test_best_multiexp_edge_scalars(kind, ctx, oracle_c):
    co = oracle_c
    curve, n = 0, 700
    C, sf, bf = co.CURVES[curve]
    rnd = random.Random(42)
    pts = _bases(co, curve, n)
    r = C.scalar.p
    if kind == "zeros":
        sc = [0] * n
    elif kind == "ones":
        sc = [1] * n
    elif kind == "minus_one":
        sc = [r - 1] * n
    elif kind == "single":
        sc = [0] * n; sc[123] = rnd.randrange(r)
    else:  # SURVEY §8d "W": 70% zero, 20% one, 5% small, 5% uniform
        sc = []
        for _ in range(n):
            u = rnd.random()
            sc.append(0 if u < 0.7 else 1 if u < 0.9 else rnd.randrange(2, 1 << 10) if u < 0.95 else rnd.randrange(r))
    sm, bm = co.to_mont(sf, sc), co.points_to_mont(curve, pts)
    got = ar.best_multiexp(ctx, curve, sm, bm)
    exp = co.best_multiexp(curve, sm, bm)
    assert co.points_from_mont(curve, co.to_affine(curve, got)) == co.points_from_mont(curve, co.to_affine(curve, exp))


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9, 11, 12, 13, 14, 15, 17, 21])
def test_best_fft_vs_oracle(field, log_n, ctx, oracle_c):
    co = oracle_c
    F = co.FIELDS[field]
    n = 1 << log_n
    rng = np.random.default_rng(log_n * 2 + field)
    a = co.from_u512(field, rng.integers(0, 2**63, size=(n, 8), dtype=np.uint64))
    om = pow(F.root_of_unity, 1 << (32 - log_n), F.p)
    for w in (om, pow(om, -1, F.p)):
        wm = co.to_mont(field, [w])
        got = ar.best_fft(ctx, field, a, wm, log_n)
        exp = co.best_fft(field, a, wm, log_n)
        assert np.array_equal(got, exp)
        if log_n > 13:
            break
    if log_n >= 2:  # an omega that is not the domain generator is refused, not silently mis-transformed
        with pytest.raises(bz.BzError):
            ar.best_fft(ctx, field, a, co.to_mont(field, [om * om % F.p]), log_n)


@pytest.mark.parametrize("k,degree", [(4, 9), (6, 9), (9, 9), (11, 9), (12, 9), (11, 5), (13, 3)])
def test_evaluation_domain_vs_oracle(k, degree, ctx, oracle_c):
    from oracle.domain import EvaluationDomain
    co = oracle_c
    dom = EvaluationDomain(0, degree, k)
    rng = np.random.default_rng(k)
    a = co.from_u512(0, rng.integers(0, 2**63, size=(dom.n, 8), dtype=np.uint64))
    assert np.array_equal(ar.lagrange_to_coeff(ctx, 0, a, k), dom.lagrange_to_coeff(a))
    ext = ar.coeff_to_extended(ctx, 0, a, k, dom.extended_k)
    assert np.array_equal(ext, dom.coeff_to_extended(a))
    big = co.from_u512(0, rng.integers(0, 2**63, size=(1 << dom.extended_k, 8), dtype=np.uint64))
    got = ar.extended_to_coeff(ctx, 0, big, dom.extended_k)[: dom.n * dom.quotient_poly_degree]
    assert np.array_equal(got, dom.extended_to_coeff(big))
    # round trip property (size independent)
    back = ar.extended_to_coeff(ctx, 0, ext, dom.extended_k)
    assert np.array_equal(back[: dom.n], a) and not back[dom.n:].any()


def test_large_ntt_roundtrip_property(ctx, oracle_c):
    """2^22 (three-pass path): inverse(forward(a)) * 1 == a, and linearity on a sample -- size-independent checks."""
    co = oracle_c
    log_n = 22
    n = 1 << log_n
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2**62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= (1 << 61) - 1          # < p: valid Montgomery residues
    d_a = ctx.to_device(a)
    d_b = ctx.alloc(n * 32)
    ctx._check(ctx.lib.bz_ntt_dev(ctx.h, 0, d_a.ptr, d_b.ptr, log_n, 0, 1))
    ctx._check(ctx.lib.bz_lagrange_to_coeff_dev(ctx.h, 0, d_b.ptr, d_b.ptr, log_n, 1))
    back = d_b.download((n, 4))
    assert np.array_equal(back, a)
    d_a.free(); d_b.free()


def test_best_multiexp_large_skewed(ctx, oracle_c):
    """2^15 points with witness-like scalars (70% zero, 20% one, ...): one bucket holds ~20% of all points, which the
    bucket kernel must spread over many threads (segments + CTA-wide fold), and the result must still be bit-exact."""
    co = oracle_c
    curve, n = 1, 1 << 15
    C, sf, bf = co.CURVES[curve]
    rnd = random.Random(77)
    base = co.points_to_mont(curve, _bases(co, curve, 512))
    bm = np.tile(base, (n // 512, 1))
    r = C.scalar.p
    sc = []
    for _ in range(n):
        u = rnd.random()
        sc.append(0 if u < 0.7 else 1 if u < 0.9 else rnd.randrange(2, 1 << 10) if u < 0.95 else rnd.randrange(r))
    sm = co.to_mont(sf, sc)
    got = ar.best_multiexp(ctx, curve, sm, bm)
    exp = co.best_multiexp(curve, sm, bm)
    assert np.array_equal(co.to_affine(curve, got), co.to_affine(curve, exp))


@pytest.mark.parametrize("curve", [0, 1])
def test_point_sum_dev(curve, ctx, oracle_c):
    """bz_point_sum_dev: the local half of the multi-GPU MSM exchange (sum of the all-gathered Jacobian partials)."""
    import ctypes
    co = oracle_c
    C, sf, bf = co.CURVES[curve]
    pts = _bases(co, curve, 6)
    pts[2] = None                                  # an identity partial (a rank whose range summed to zero)
    pts[4] = C.neg(pts[3])                         # two partials that cancel
    aff = co.points_to_mont(curve, pts)
    one = np.frombuffer(co.FIELDS[bf].to_mont_bytes(1), dtype=np.uint64)
    jac = np.concatenate([aff, np.repeat(one[None, :], len(pts), axis=0)], axis=1).copy()
    jac[2, 8:] = 0                                 # z = 0
    d_j, d_o = ctx.to_device(jac), ctx.alloc(64)
    ctx._check(ctx.lib.bz_point_sum_dev(ctx.h, curve, d_j.ptr, len(pts), d_o.ptr))
    got = d_o.download((8,))
    exp = None
    for p in pts:
        exp = C.add(exp, p)
    assert co.points_from_mont(curve, got[None, :])[0] == exp
    d_j.free(); d_o.free()
