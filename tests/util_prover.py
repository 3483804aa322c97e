"""Shared helpers for the prover parity tests: build a job (circuit IR, keys, witness, RNG stream) for both the
oracle (CPU restatement) and the product (libbzhalo2 via ctypes)."""
import os
import numpy as np
from oracle import halo2 as H, c_oracle as co, pasta
import battlezips_halo2_b200  # noqa: F401  (registers the package)
from battlezips_halo2_b200.plonk.circuit import ConstraintSystem, Assignment, Constant

P = pasta.P
VK_REPR = 0x1234567890ABCDEF1234567890ABCDEF          # opaque vk.hash_into scalar


def tiny_circuit(k=5):
    """3 advice, 5 fixed, 1 instance; mul/add gates with a rotation, one lookup, copies across all column kinds."""
    cs = ConstraintSystem(P)
    a, b, c = cs.advice_column(), cs.advice_column(), cs.advice_column()
    q_mul, q_add, tbl, q_lk, const = [cs.fixed_column() for _ in range(5)]
    inst = cs.instance_column()
    for col in (a, b, c):
        cs.enable_equality("advice", col)
    cs.enable_equality("fixed", const)
    cs.enable_equality("instance", inst)
    A, B, Cc = cs.query_advice(a), cs.query_advice(b), cs.query_advice(c)
    An = cs.query_advice(a, 1)
    cs.create_gate("mul", [cs.query_fixed(q_mul) * (A * B - Cc)])
    cs.create_gate("add", [cs.query_fixed(q_add) * (A + B - An), cs.query_fixed(q_add) * (Cc * (Cc - Constant(1))) * 3])
    cs.lookup("range", [(cs.query_fixed(q_lk) * A, cs.query_fixed(tbl))])
    asg = Assignment(cs, k)
    for r in range(min(16, asg.usable_rows)):
        asg.assign_fixed(tbl, r, r)
    asg.assign_fixed(q_mul, 0, 1); asg.assign_advice(a, 0, 3); asg.assign_advice(b, 0, 4); asg.assign_advice(c, 0, 12)
    asg.assign_fixed(q_add, 1, 1); asg.assign_advice(a, 1, 5); asg.assign_advice(b, 1, 6); asg.assign_advice(c, 1, 1)
    asg.assign_advice(a, 2, 11)
    for r in range(3):
        asg.assign_fixed(q_lk, r, 1)
    asg.assign_fixed(const, 0, 12)
    asg.copy(("advice", c, 0), ("fixed", const, 0))
    asg.copy(("advice", a, 2), ("instance", inst, 0))
    asg.assign_advice(b, 3, 4)
    asg.copy(("advice", b, 0), ("advice", b, 3))
    asg.set_instance(inst, [11])
    return cs, asg


def device_urs_checked(ctx, k):
    """Large k: the oracle's Python Params::new would take minutes, so both sides use the URS that Params::new on the
    device returns (bit-exact against the oracle fixtures for k <= 13, tests/test_gpu_params.py) after re-checking it
    here with oracle arithmetic: sampled g[i] against the oracle hash-to-curve, w and u, and the two Lagrange-basis
    identities  sum_i g_lagrange[i] = g[0]  and  sum_i omega^i g_lagrange[i] = g[1]  (1 and X in the Lagrange basis)."""
    from battlezips_halo2_b200 import arithmetic as ar
    urs = ar.params_new(ctx, k, curve=0)
    n = 1 << k
    C = co.CURVES[0][0]
    h = C.hash_to_curve("Halo2-Parameters")
    for i in (0, 1, n // 3, n - 1):
        assert co.points_from_mont(0, urs["g"][i][None, :])[0] == h(b"\x00" + i.to_bytes(4, "little"))
    assert co.points_from_mont(0, urs["w"][None, :])[0] == h(b"\x01") and co.points_from_mont(0, urs["u"][None, :])[0] == h(b"\x02")
    F = co.FIELDS[0]
    om = pow(F.root_of_unity, 1 << (32 - k), F.p)
    ones = np.repeat(co.to_mont(0, [1]), n, axis=0)
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * om % F.p
    assert np.array_equal(co.to_affine(0, co.best_multiexp(0, ones, urs["g_lagrange"]))[0], urs["g"][0])
    assert np.array_equal(co.to_affine(0, co.best_multiexp(0, co.to_mont(0, pw), urs["g_lagrange"]))[0], urs["g"][1])
    return H.Params(k, 0, urs["g"], urs["g_lagrange"], urs["w"], urs["u"])


class Job:
    """Everything both sides need for one circuit: IR, oracle pk, witness arrays, RNG words."""

    def __init__(self, cs, asg, seed=0xB200B200B200B200, params_from_device=None):
        self.cs, self.asg, self.k = cs, asg, asg.k
        self.ir = cs.to_ir()
        fixture = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"params_vesta_k{self.k}.npz")
        if os.path.exists(fixture):          # URS fixtures made by tests/golden/make_params.py (oracle Params::new)
            d = np.load(fixture)
            self.oparams = H.Params(self.k, 0, d["g"], d["g_lagrange"], d["w"], d["u"])
        elif params_from_device is not None:
            self.oparams = device_urs_checked(params_from_device, self.k)
        else:
            self.oparams = H.Params.new(self.k, 0)
        self.mapping = asg.permutation_mapping()
        self.opk = H.keygen(self.oparams, self.ir, asg.fixed, self.mapping, vk_repr=VK_REPR)
        self.V = H.Vec(0)
        self.advice = np.stack([self.V.arr(col) for col in asg.advice])          # (G, n, 4)
        self.instances = asg.instance
        self.seed = seed

    def num_random(self):
        ir, n = self.ir, 1 << self.k
        bf, G, L = ir["blinding_factors"], ir["num_advice"], len(ir["lookups"])
        chunk = ir["degree"] - 2
        nsets = (len(ir["permutation"]) + chunk - 1) // chunk if ir["permutation"] else 0
        return (G * (bf + 1) + G + L * (2 * (bf + 1) + 2) + nsets * (bf + 1) + L * (bf + 1) + n + 1 + (ir["degree"] - 1) + 1 + n + 1 + 2 * self.k)

    def wide(self, index=0):
        return H.splitmix64_wide(self.seed + index, self.num_random())

    def oracle_proof(self, index=0, advice=None, instances=None, trace=None):
        T = H.Blake2bTranscript(0)
        draws = H.Draws(self.wide(index))
        adv = self.advice if advice is None else advice
        H.create_proof(self.oparams, self.opk, self.instances if instances is None else instances, [adv[i] for i in range(len(adv))], draws, T, trace)
        assert draws.pos == self.num_random()
        return T.finalize()

    def verify(self, proof, instances=None):
        return H.verify_proof(self.oparams, self.opk, self.instances if instances is None else instances, proof)

    # ---- product side ----
    def device_keys(self, ctx, window_bits=0):
        from battlezips_halo2_b200.plonk import prover as PR
        op = self.oparams
        params = PR.Params(ctx, self.k, op.g, op.g_lagrange, op.w, op.u, window_bits=window_bits)
        pk = PR.ProvingKey(ctx, params, self.ir, self.asg.fixed, self.mapping, VK_REPR)
        assert pk.num_random == self.num_random()
        return params, pk


def first_diff(a, b):
    if len(a) != len(b):
        return f"length {len(a)} != {len(b)}"
    for i in range(0, len(a), 32):
        if a[i:i + 32] != b[i:i + 32]:
            return f"first differing 32-byte item #{i // 32} of {len(a) // 32}"
    return None
