"""GPU parity for verify_proof on the device through the C ABI: accepts what the restated reference
verifier accepts (oracle and GPU proofs), rejects what it rejects (bit flips in every region of the proof, wrong
instances, malformed encodings, wrong lengths), one verdict per proof of a mixed batch."""
import numpy as np
import pytest
from tests.util_prover import Job, tiny_circuit

pytestmark = pytest.mark.gpu


def _prove(job, pk, indices):
    from battlezips_halo2_b200.plonk import prover as PR
    B = len(indices)
    return PR.create_proofs(pk, [job.instances] * B, np.stack([job.advice] * B), np.stack([job.wide(i) for i in indices]))


def _flip(proof, byte, bit=0):
    b = bytearray(proof)
    b[byte] ^= 1 << bit
    return bytes(b)


def test_tiny_accepts_and_rejects_like_the_oracle(ctx):
    from battlezips_halo2_b200.plonk import prover as PR
    job = Job(*tiny_circuit(5))
    params, pk = job.device_keys(ctx, window_bits=6)
    good = _prove(job, pk, [0, 1, 2])
    assert PR.verify_proofs(pk, [job.instances] * 3, good) == [True, True, True]
    assert PR.verify_proofs(pk, [job.instances], [job.oracle_proof(index=5)]) == [True]
    # one flipped bit in every 32-byte item of the proof (points, evaluations, u-evaluations, L/R, c, f)
    proof = good[0]
    items = len(proof) // 32
    bad = [_flip(proof, 32 * i + (7 * i) % 31, i % 8) for i in range(items)]
    got = PR.verify_proofs(pk, [job.instances] * items, bad)
    exp = [job.verify(p) for p in bad]
    assert got == exp
    assert not any(got)
    # a mixed batch keeps per-proof verdicts
    assert PR.verify_proofs(pk, [job.instances] * 4, [good[1], bad[3], good[2], bad[-1]]) == [True, False, True, False]
    # wrong public input
    assert PR.verify_proofs(pk, [[[12]]], [proof]) == [False] and not job.verify(proof, instances=[[12]])
    # wrong lengths
    assert PR.verify_proofs(pk, [job.instances], [proof[:-32]]) == [False]
    assert PR.verify_proofs(pk, [job.instances], [proof + bytes(32)]) == [False]
    pk.close(); params.close()


def test_malformed_encodings_are_rejected(ctx):
    from battlezips_halo2_b200.plonk import prover as PR
    job = Job(*tiny_circuit(5))
    params, pk = job.device_keys(ctx, window_bits=6)
    proof = _prove(job, pk, [0])[0]
    p = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001
    G = job.ir["num_advice"]
    bad = []
    bad.append(bytes(32) + proof[32:])                                   # identity where a commitment is expected
    bad.append(b"\xff" * 31 + b"\x7f" + proof[32:])                      # x >= q
    x_not_on_curve = None
    for x in range(1, 50):                                              # smallest x with x^3 + 5 a non-residue mod q
        q = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001
        if pow((x ** 3 + 5) % q, (q - 1) // 2, q) != 1:
            x_not_on_curve = x
            break
    bad.append(x_not_on_curve.to_bytes(32, "little") + proof[32:])
    ev_off = len(proof) - 64                                            # the scalar c: p itself is not canonical
    bad.append(proof[:ev_off] + p.to_bytes(32, "little") + proof[ev_off + 32:])
    got = PR.verify_proofs(pk, [job.instances] * len(bad), bad)
    assert got == [False] * len(bad)
    assert [job.verify(b) for b in bad] == got
    pk.close(); params.close()


def test_shot_batch_verifies(ctx):
    """BASELINE config 3 in small: a batch of Shot proofs, every one verified on the device; the oracle agrees on the
    first, and on a tampered copy."""
    from battlezips_halo2_b200.circuits import shot_circuit
    from battlezips_halo2_b200.plonk import prover as PR
    cs, cfg, asg = shot_circuit(0)
    job = Job(cs, asg)
    params, pk = job.device_keys(ctx)
    proofs = _prove(job, pk, list(range(8)))
    assert PR.verify_proofs(pk, [job.instances] * 8, proofs) == [True] * 8
    assert job.verify(proofs[0])
    t = _flip(proofs[3], 1000, 2)
    assert PR.verify_proofs(pk, [job.instances] * 2, [t, proofs[4]]) == [False, True]
    assert not job.verify(t)
    pk.close(); params.close()


def test_board_verifies(ctx):
    from battlezips_halo2_b200.circuits import board_circuit
    from battlezips_halo2_b200.plonk import prover as PR
    cs, cfg, asg = board_circuit(0)
    job = Job(cs, asg)
    params, pk = job.device_keys(ctx)
    proof = _prove(job, pk, [0])[0]
    assert PR.verify_proofs(pk, [job.instances], [proof]) == [True]
    assert PR.verify_proofs(pk, [job.instances], [_flip(proof, 2048)]) == [False]
    pk.close(); params.close()


def test_shot_batch_config3_shape(ctx):
    """BASELINE config 3 in small (the bench runs the full width): 128 independent Shot proofs in one call -- 8 distinct
    witnesses (board pattern, shot cell, hit bit), every proof its own RNG stream, one proving key.  Size-independent
    properties: every proof is accepted by verify_proof on the device, all proofs are distinct, a proof is rejected under
    another job's public inputs; one member is byte-compared with the oracle prover."""
    from battlezips_halo2_b200.circuits import shot_circuit
    from battlezips_halo2_b200.plonk import prover as PR
    jobs = [shot_circuit(i) for i in range(8)]
    job = Job(jobs[0][0], jobs[0][2])
    params, pk = job.device_keys(ctx)
    B = 128
    per_job = [np.stack([job.V.arr(col) for col in jobs[j][2].advice]) for j in range(8)]
    advice = np.stack([per_job[b % 8] for b in range(B)])
    instances = [jobs[b % 8][2].instance for b in range(B)]
    wide = np.stack([job.wide(1000 + b) for b in range(B)])
    proofs = PR.create_proofs(pk, instances, advice, wide)
    assert PR.verify_proofs(pk, instances, proofs) == [True] * B
    assert len(set(proofs)) == B
    b = 77
    assert proofs[b] == job.oracle_proof(index=1000 + b, advice=advice[b], instances=instances[b])
    other = next(j for j in range(8) if jobs[j][2].instance != instances[b])
    assert PR.verify_proofs(pk, [jobs[other][2].instance], [proofs[b]]) == [False]
    pk.close(); params.close()


def test_scaled_board_k15_roundtrip_on_device_only(ctx):
    """BASELINE config 5 in small, without the (slow) oracle: the Board circuit tiled down 2^15 rows -- URS from Params::new on
    the device, keygen on the device, create_proof (two-level grand-product scans, radix-sort lookup permutation), and
    verify_proof on the device accepts; a flipped bit is rejected.  Size-independent property: prover and verifier meet."""
    from battlezips_halo2_b200 import arithmetic as ar
    from battlezips_halo2_b200.circuits import board_circuit_scaled
    from battlezips_halo2_b200.plonk import prover as PR
    from oracle import halo2 as H
    from tests.util_prover import VK_REPR
    k = 15
    cs, cfg, asg = board_circuit_scaled(k)
    ir = cs.to_ir()
    urs = ar.params_new(ctx, k, curve=0)
    params = PR.Params(ctx, k, urs["g"], urs["g_lagrange"], urs["w"], urs["u"])
    pk = PR.ProvingKey(ctx, params, ir, asg.fixed, asg.permutation_mapping(), VK_REPR)
    advice = np.stack([PR.mont(col) for col in asg.advice])
    wide = H.splitmix64_wide(0xB200 + k, pk.num_random)
    proof = PR.create_proofs(pk, [asg.instance], advice[None], wide[None])[0]
    assert PR.verify_proofs(pk, [asg.instance], [proof]) == [True]
    assert PR.verify_proofs(pk, [asg.instance], [_flip(proof, len(proof) // 2)]) == [False]
    pk.close(); params.close()
