"""GPU parity for `Params::new` on the device through the C ABI:
hash-to-curve against the reference's own Pallas golden vectors and against the oracle on both curves; the whole URS
(g, g_lagrange, w, u) bit-exact against the committed fixtures (tests/golden/params_vesta_k*.npz, made by the oracle's
Params::new restatement with tests/golden/make_params.py)."""
import json, os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "pallas_fixed_base_kats.json")))


def _le(hexstr):
    return int.from_bytes(bytes.fromhex(hexstr), "little")


@pytest.mark.parametrize("name", ["v", "r"])
def test_hash_to_curve_reference_generator_kat(ctx, oracle_c, name):
    """/root/reference/src/utils/constants/fixed_bases/board_commit_{v,r}.rs:5-14 (reference tests :2940-2948)."""
    from battlezips_halo2_b200 import arithmetic as ar
    got = ar.hash_to_curve(ctx, 1, GOLD["personalization"], [GOLD[name]["message"].encode()])
    exp = oracle_c.points_to_mont(1, [(_le(GOLD[name]["generator_x"]), _le(GOLD[name]["generator_y"]))])
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("msg_len", [0, 1, 5, 33, 119, 200])
def test_hash_to_curve_matches_oracle(ctx, oracle_c, curve, msg_len):
    from battlezips_halo2_b200 import arithmetic as ar
    rng = np.random.default_rng(100 * curve + msg_len)
    msgs = [bytes(rng.integers(0, 256, size=msg_len, dtype=np.uint8)) for _ in range(6)]
    for dom in ("Halo2-Parameters", "z"):
        got = ar.hash_to_curve(ctx, curve, dom, msgs)
        h = oracle_c.CURVES[curve][0].hash_to_curve(dom)
        exp = oracle_c.points_to_mont(curve, [h(m) for m in msgs])
        assert np.array_equal(got, exp), (dom, msg_len)


@pytest.mark.parametrize("k", [5, 11])
def test_params_new_matches_fixture(ctx, k):
    """g, g_lagrange (group inverse FFT), w, u of Params::<vesta::Affine>::new(k) -- /root/reference/benches/shot.rs:58."""
    from battlezips_halo2_b200 import arithmetic as ar
    fx = np.load(os.path.join(HERE, "golden", f"params_vesta_k{k}.npz"))
    got = ar.params_new(ctx, k, curve=0)
    for name in ("g", "w", "u", "g_lagrange"):
        assert np.array_equal(got[name], fx[name]), name


def test_params_new_pallas_small_matches_oracle(ctx, oracle_c):
    """The other curve of the cycle (Params<pallas::Affine>), k = 3, against the oracle's Params::new restatement."""
    from battlezips_halo2_b200 import arithmetic as ar
    from oracle import halo2 as H
    exp = H.Params.new(3, 1, cache=False)
    got = ar.params_new(ctx, 3, curve=1)
    assert np.array_equal(got["g"], exp.g) and np.array_equal(got["g_lagrange"], exp.g_lagrange)
    assert np.array_equal(got["w"], exp.w) and np.array_equal(got["u"], exp.u)


def test_device_urs_proves_and_verifies(ctx):
    """A URS generated on the device feeds Params -> ProvingKey -> create_proof; bytes match the oracle prover (which
    uses the fixture URS) and the restated verifier accepts."""
    from tests.util_prover import Job, tiny_circuit, VK_REPR
    from battlezips_halo2_b200 import arithmetic as ar
    from battlezips_halo2_b200.plonk import prover as PR
    job = Job(*tiny_circuit(5))
    urs = ar.params_new(ctx, 5, curve=0)
    params = PR.Params(ctx, 5, urs["g"], urs["g_lagrange"], urs["w"], urs["u"], window_bits=6)
    pk = PR.ProvingKey(ctx, params, job.ir, job.asg.fixed, job.mapping, VK_REPR)
    proof = PR.create_proofs(pk, [job.instances], job.advice[None], job.wide(0)[None])[0]
    assert proof == job.oracle_proof(index=0)
    assert job.verify(proof)
    pk.close(); params.close()
