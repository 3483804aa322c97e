"""CPU: the restated halo2 prover/verifier (oracle/halo2.py) round-trips and rejects tampering -- the shape of the
reference's own `production` tests (/root/reference/src/circuits/shot.rs:880-941, board.rs:879-933)."""
import pytest
from tests.util_prover import Job, tiny_circuit


@pytest.fixture(scope="module")
def tiny():
    return Job(*tiny_circuit(5))


def test_roundtrip_tiny(tiny):
    proof = tiny.oracle_proof()
    assert tiny.verify(proof)
    assert tiny.oracle_proof() == proof                       # deterministic under the seeded RNG
    assert tiny.oracle_proof(index=1) != proof


def test_tamper_rejected(tiny):
    proof = bytearray(tiny.oracle_proof())
    for pos in (5, 40, len(proof) // 2, len(proof) - 1):
        bad = bytearray(proof); bad[pos] ^= 1
        assert not tiny.verify(bytes(bad))
    assert not tiny.verify(bytes(proof), instances=[[12]])
    assert not tiny.verify(bytes(proof[:-32]))


def test_unsatisfied_witness_does_not_verify(tiny):
    adv = tiny.advice.copy()
    adv[2, 0] = tiny.V.m(13)                                   # 3 * 4 != 13 and breaks the copy to the constant 12
    assert not tiny.verify(tiny.oracle_proof(advice=adv))


def test_lookup_failure_is_an_error(tiny):
    adv = tiny.advice.copy()
    adv[0, 0] = tiny.V.m(1000)                                 # not in the 16-row table
    with pytest.raises(ValueError):
        tiny.oracle_proof(advice=adv)


def test_shot_circuit_shape_and_roundtrip():
    """Shot mirror: 24 gates, 13 permutation columns, degree 9, bf 5, 4238 RNG draws."""
    from battlezips_halo2_b200.circuits import shot_circuit
    cs, cfg, asg = shot_circuit(0)
    assert len(cs.gates) == 24 and cs.degree() == 9 and cs.blinding_factors() == 5
    assert len(cs.permutation) == 13 and cs.num_advice == 11 and len(cs.lookups) == 1
    assert asg.check_satisfied() is None
    job = Job(cs, asg)
    assert job.num_random() == 4238
    proof = job.oracle_proof()
    assert len(proof) == 4000
    assert job.verify(proof)
    bad = bytearray(proof); bad[100] ^= 0x10
    assert not job.verify(bytes(bad))


def test_board_circuit_shape():
    from battlezips_halo2_b200.circuits import board_circuit
    cs, cfg, asg = board_circuit(0)
    assert len(cs.gates) == 57 and cs.degree() == 9 and len(cs.permutation) == 13 and cs.num_advice == 11
    assert cs.blinding_factors() == 7
    assert asg.check_satisfied() is None


def test_oracle_reproduces_committed_golden_proofs(tiny):
    """tests/golden/proofs.npz (made by tests/golden/make_proofs.py) freezes the oracle's proof bytes: any drift in the
    restated protocol order, RNG draw order or transcript shows up here, on the CPU, before a GPU run is spent."""
    import os
    import numpy as np
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proofs.npz"))
    assert tiny.oracle_proof(index=0) == bytes(gold["tiny_k5_idx0"])
    assert tiny.oracle_proof(index=3) == bytes(gold["tiny_k5_idx3"])
    from battlezips_halo2_b200.circuits import shot_circuit
    cs, _, asg = shot_circuit(0)
    job = Job(cs, asg)
    proof = job.oracle_proof(index=0)
    assert proof == bytes(gold["shot_w0_idx0"]) and job.verify(proof)


def test_oracle_batch_invert_assigned_is_division():
    """The restated poly::batch_invert_assigned equals numerator / denominator (0 for a zero denominator) in big integers."""
    import random
    from oracle import c_oracle as co, halo2 as H
    from oracle.vec import Vec
    for f in (0, 1):
        F = co.FIELDS[f]
        p, rnd = F.p, random.Random(f)
        num = [rnd.randrange(p) for _ in range(200)]
        den = [rnd.choice([0, 1, 2, p - 1, rnd.randrange(p)]) for _ in range(200)]
        got = H.batch_invert_assigned(Vec(f), co.to_mont(f, num), co.to_mont(f, den))
        assert co.from_mont(f, got) == [x * F.inv(d) % p for x, d in zip(num, den)]
