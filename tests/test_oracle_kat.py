"""CPU: pin the oracle (Python big-int AND the C restatement) on the reference's own golden vectors.

Golden data: tests/golden/pallas_fixed_base_kats.json, extracted by tests/golden/make_pallas_kats.py from
/root/reference/src/utils/constants/fixed_bases/board_commit_{v,r}.rs (GENERATOR :5-14, Z :17-26, U :28-2919);
the reference checks them at :2940-2960 with halo2_gadgets' test_zs_and_us."""
import json, os, random
import numpy as np
import pytest
from oracle import pasta

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pallas_fixed_base_kats.json")))


def _le(hexstr):
    return int.from_bytes(bytes.fromhex(hexstr), "little")


@pytest.mark.parametrize("name", ["v", "r"])
def test_hash_to_curve_generator_kat(name):
    h = pasta.PALLAS.hash_to_curve(GOLD["personalization"])
    pt = h(GOLD[name]["message"].encode())
    assert pt == (_le(GOLD[name]["generator_x"]), _le(GOLD[name]["generator_y"]))


def _window_scalar(w, k):
    if w < 84:
        return (k + 2) * 8 ** w
    return (k * 8 ** 84 - sum(2 ** (3 * j + 1) for j in range(84))) % pasta.Q


@pytest.mark.parametrize("name", ["v", "r"])
def test_window_table_kats_python(name):
    """z + y([s]B) == u^2 for a sample of windows (full table is covered by the C oracle below)."""
    g = GOLD[name]
    B = (_le(g["generator_x"]), _le(g["generator_y"]))
    for w in (0, 1, 41, 83, 84):
        for k in range(8):
            y = pasta.PALLAS.mul(B, _window_scalar(w, k))[1]
            u = _le(g["u"][w][k])
            assert (g["z"][w] + y) % pasta.P == u * u % pasta.P


@pytest.mark.parametrize("name", ["v", "r"])
def test_window_table_kats_c_oracle_full(name, oracle_c):
    """All 85 x 8 entries through the C restatement's scalar multiplication (1 360 equalities over both bases)."""
    import ctypes
    co = oracle_c
    g = GOLD[name]
    B = (_le(g["generator_x"]), _le(g["generator_y"]))
    base = co.points_to_mont(1, [B])
    out = np.zeros(8, dtype=np.uint64)
    for w in range(85):
        for k in range(8):
            s = co.to_mont(1, [_window_scalar(w, k)])
            co.lib().orc_point_mul(1, co._p(base), co._p(s), co._p(out))
            y = co.points_from_mont(1, out)[0][1]
            u = _le(g["u"][w][k])
            assert (g["z"][w] + y) % pasta.P == u * u % pasta.P, (w, k)


def test_c_oracle_matches_python_fields(oracle_c):
    co = oracle_c
    rnd = random.Random(7)
    for f, F in co.FIELDS.items():
        a = [rnd.randrange(F.p) for _ in range(64)] + [0, 1, F.p - 1, F.p - 1]
        b = [rnd.randrange(F.p) for _ in range(64)] + [F.p - 1, F.p - 1, F.p - 1, 2]
        am, bm = co.to_mont(f, a), co.to_mont(f, b)
        assert co.from_mont(f, am) == a
        assert am[3].tobytes() == F.to_mont_bytes(a[3])
        assert co.from_mont(f, co.field_mul(f, am, bm)) == [x * y % F.p for x, y in zip(a, b)]
        wide = np.frombuffer(rnd.randbytes(64 * 32), dtype=np.uint64).reshape(-1, 8)
        assert co.from_mont(f, co.from_u512(f, wide)) == [
            int.from_bytes(wide[i].tobytes(), "little") % F.p for i in range(32)]


@pytest.mark.parametrize("curve", [0, 1])
def test_c_best_multiexp_matches_python(curve, oracle_c):
    co = oracle_c
    C, sf, bf = co.CURVES[curve]
    rnd = random.Random(11 + curve)
    h = C.hash_to_curve("Halo2-Parameters")
    n = 45
    pts = [h(b"\x00" + i.to_bytes(4, "little")) for i in range(n)]
    sc = [rnd.randrange(C.scalar.p) for _ in range(n)]
    sc[3], sc[4], sc[5] = 0, 1, C.scalar.p - 1
    pts[7] = None
    pts[9] = pts[8]
    pts[11] = C.neg(pts[10]); sc[11] = sc[10]
    acc = C.j_identity()
    for s, p in zip(sc, pts):
        acc = C.j_add(acc, C.j_mul(C.to_jac(p), s))
    expect = C.to_affine(acc)
    for threads in (1, 3, 8):
        co.set_threads(threads)
        got = co.points_from_mont(curve, co.to_affine(curve, co.best_multiexp(curve, co.to_mont(sf, sc), co.points_to_mont(curve, pts))))[0]
        assert got == expect
    co.set_threads(8)


@pytest.mark.parametrize("field", [0, 1])
def test_c_best_fft_matches_naive_dft(field, oracle_c):
    co = oracle_c
    F = co.FIELDS[field]
    rnd = random.Random(5)
    for logn in (0, 1, 2, 5, 7):
        n = 1 << logn
        om = pow(F.root_of_unity, 1 << (32 - logn), F.p)
        a = [rnd.randrange(F.p) for _ in range(n)]
        expect = [sum(a[j] * pow(om, i * j, F.p) for j in range(n)) % F.p for i in range(n)]
        for threads in (1, 2, 8):
            co.set_threads(threads)
            got = co.from_mont(field, co.best_fft(field, co.to_mont(field, a), co.to_mont(field, [om]), logn))
            assert got == expect
    co.set_threads(8)


def test_domain_roundtrip_and_coset(oracle_c):
    """EvaluationDomain restatement: coeff -> extended -> coeff is the identity; extended values are the
    polynomial evaluated on zeta * w_ext^i."""
    from oracle.domain import EvaluationDomain
    co = oracle_c
    rnd = random.Random(3)
    dom = EvaluationDomain(0, 9, 4)
    assert dom.extended_k == 7
    F = dom.F
    coeffs = [rnd.randrange(F.p) for _ in range(dom.n)]
    cm = co.to_mont(0, coeffs)
    ext = dom.coeff_to_extended(cm)
    vals = co.from_mont(0, ext)
    for i in (0, 1, 5, 127):
        x = F.zeta * pow(dom.extended_omega, i, F.p) % F.p
        assert vals[i] == sum(c * pow(x, j, F.p) for j, c in enumerate(coeffs)) % F.p
    back = co.from_mont(0, dom.extended_to_coeff(ext))
    assert back[: dom.n] == coeffs and all(v == 0 for v in back[dom.n:])
    lag = [rnd.randrange(F.p) for _ in range(dom.n)]
    co_ = co.from_mont(0, dom.lagrange_to_coeff(co.to_mont(0, lag)))
    for i in (0, 3, 15):
        x = pow(dom.omega, i, F.p)
        assert lag[i] == sum(c * pow(x, j, F.p) for j, c in enumerate(co_)) % F.p
