"""Function-level GPU parity for SURVEY 8 rows a6-a11 through the fine-grained C ABI (bz_perm_product, bz_lookup_permute,
bz_lookup_product, bz_pk_quotient, bz_divide_by_vanishing, bz_eval_many, bz_kate_div, bz_axpy, bz_ipa_*): every call is
compared with the intermediate the oracle prover records in its `trace` for the same proof, so a
mismatch names the kernel instead of "first differing 32-byte item #k" of the proof."""
import numpy as np
import pytest
from battlezips_halo2_b200 import arithmetic as ar
from tests.util_prover import Job, tiny_circuit

pytestmark = pytest.mark.gpu


def _job(which):
    if which == "tiny":
        return Job(*tiny_circuit(5))
    from battlezips_halo2_b200.circuits import shot_circuit, board_circuit, board_circuit_scaled
    if which == "scaled13":
        cs, cfg, asg = board_circuit_scaled(13)
    else:
        cs, cfg, asg = (shot_circuit if which == "shot" else board_circuit)(1)
    return Job(cs, asg)


@pytest.fixture(scope="module", params=["tiny", "shot", "board", "scaled13"])
def traced(request, ctx):
    job = _job(request.param)
    tr = {}
    proof = job.oracle_proof(index=5, trace=tr)
    params, pk = job.device_keys(ctx, window_bits=8 if request.param != "scaled13" else 0)
    yield job, tr, params, pk, proof
    pk.close(); params.close()


def test_a6_permutation_product(traced, ctx):
    job, tr, params, pk, _ = traced
    V, ch = job.V, tr["challenges"]
    for s in tr["perm_sets"]:
        got = ar.permutation_product(ctx, 0, job.k, s["values"], s["sigmas"], V.m(ch["beta"]), V.m(ch["gamma"]), V.m(s["delta_omega0"]), V.m(s["z0"]))
        assert np.array_equal(got, s["z_unblinded"])


def test_a7_lookup_permute_and_product(traced, ctx):
    job, tr, params, pk, _ = traced
    V, ch = job.V, tr["challenges"]
    n, usable = 1 << job.k, (1 << job.k) - (job.ir["blinding_factors"] + 1)
    for L in tr["lookups"]:
        a, s = ar.lookup_permute(ctx, 0, job.k, usable, L["compressed_input"], L["compressed_table"])
        assert np.array_equal(a[:usable], L["permuted_input"][:usable]) and not a[usable:].any()
        assert np.array_equal(s[:usable], L["permuted_table"][:usable]) and not s[usable:].any()
        z = ar.lookup_product(ctx, 0, job.k, L["compressed_input"], L["compressed_table"], L["permuted_input"], L["permuted_table"],
                              V.m(ch["beta"]), V.m(ch["gamma"]))
        assert np.array_equal(z, L["z_unblinded"])


def test_a7_lookup_permute_rejects_missing_value(traced, ctx):
    import battlezips_halo2_b200 as bz
    job, tr, params, pk, _ = traced
    usable = (1 << job.k) - (job.ir["blinding_factors"] + 1)
    for L in tr["lookups"][:1]:
        bad = L["compressed_input"].copy()
        bad[0] = job.V.m(0x123456789ABCDEF0123)
        with pytest.raises(bz.BzError) as e:
            ar.lookup_permute(ctx, 0, job.k, usable, bad, L["compressed_table"])
        assert e.value.code == -4


def test_a8_quotient_and_divide_by_vanishing(traced, ctx):
    job, tr, params, pk, _ = traced
    V, ch = job.V, tr["challenges"]
    n = 1 << job.k
    h = pk.quotient(np.stack(tr["polys"]), V.m(ch["theta"]), V.m(ch["beta"]), V.m(ch["gamma"]), V.m(ch["y"]))
    exp = np.concatenate(tr["h_pieces"])
    assert h.shape == exp.shape and np.array_equal(h, exp)
    ext_k = (len(tr["h_extended"]) - 1).bit_length()
    got = ar.divide_by_vanishing_poly(ctx, 0, tr["h_extended_before_division"], job.k, ext_k)
    assert np.array_equal(got, tr["h_extended"])
    assert np.array_equal(ar.extended_to_coeff(ctx, 0, got, ext_k)[: len(exp)], exp)


def test_a9_eval_polynomial_many(traced, ctx):
    job, tr, params, pk, _ = traced
    V = job.V
    polys = [q[2][0] for q in tr["queries"]]
    points = np.stack([V.m(q[1]) for q in tr["queries"]])
    got = ar.eval_polynomials(ctx, 0, polys, points)
    exp = np.stack([V.m(V.eval_polynomial(q[2][0], q[1])) for q in tr["queries"]])
    assert np.array_equal(got, exp)


def test_a10_multiopen_axpy_and_kate_division(traced, ctx):
    job, tr, params, pk, _ = traced
    V, ch = job.V, tr["challenges"]
    n = 1 << job.k
    # q_set accumulation: replay the oracle's order with bz_axpy
    from oracle.halo2 import construct_intermediate_sets
    cmap, point_sets = construct_intermediate_sets(tr["queries"])
    acc = [None] * len(point_sets)
    for d in cmap:
        poly, si = d["first_payload"][0], d["set_index"]
        acc[si] = poly if acc[si] is None else ar.axpy(ctx, 0, acc[si], V.m(ch["x1"]), poly)
    for a, e in zip(acc, tr["q_polys"]):
        assert np.array_equal(a, e)
    # q' = fold over the sets of kate_division by every point of the set
    q_prime = None
    for pts, poly in zip(point_sets, acc):
        quo = poly
        for pt in pts:
            got = ar.kate_division(ctx, 0, quo, V.m(pt))
            assert np.array_equal(got, V.kate_division(quo, pt))
            quo = got
        quo = np.concatenate([quo, V.zeros(n - len(quo))])
        q_prime = quo if q_prime is None else ar.axpy(ctx, 0, q_prime, V.m(ch["x2"]), quo)
    assert np.array_equal(q_prime, tr["q_prime"])


def test_a11_ipa_rounds(traced, ctx, oracle_c):
    job, tr, params, pk, _ = traced
    V, ipa = job.V, tr["ipa"]
    p = V.p
    rounds = [(V.m(r["l_rand"]), V.m(r["r_rand"]), (lambda L, R, u=r["u"]: (V.m(u), V.m(pow(u, -1, p))))) for r in ipa["rounds"]]
    got, c = params.ipa_open(ipa["p_prime"], V.m(ipa["x3"]), V.m(ipa["z"]), rounds)
    for (L, R), r in zip(got, ipa["rounds"]):
        assert oracle_c.points_from_mont(0, np.stack([L, R])) == [r["L"], r["R"]]
    assert V.int1(c) == ipa["c"]
