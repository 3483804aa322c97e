"""Generate tests/golden/proofs.npz: proof bytes of the ORACLE prover (oracle/halo2.py, the CPU restatement of
halo2_proofs 0.2.0 `create_proof`) for the tiny k = 5 circuit, Shot (k = 11) and Board (k = 12), each under the seeded
SplitMix64 RNG stream `Job.wide(index)`.  They freeze the oracle: tests/test_oracle_prover.py recomputes them on the
CPU, tests/test_gpu_prover.py compares the device proofs with them.  (Upstream pins no proof bytes at all -- the
reference draws from OsRng -- so these are oracle goldens, not reference goldens: "parity unpinned", DESIGN.md.)
Run:  python tests/golden/make_proofs.py"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np


def jobs():
    from tests.util_prover import Job, tiny_circuit
    from battlezips_halo2_b200.circuits import shot_circuit, board_circuit
    yield "tiny_k5_idx0", Job(*tiny_circuit(5)), 0
    yield "tiny_k5_idx3", Job(*tiny_circuit(5)), 3
    cs, _, asg = shot_circuit(0)
    yield "shot_w0_idx0", Job(cs, asg), 0
    cs, _, asg = board_circuit(0)
    yield "board_w0_idx0", Job(cs, asg), 0


if __name__ == "__main__":
    out = {}
    for name, job, idx in jobs():
        proof = job.oracle_proof(index=idx)
        assert job.verify(proof)
        out[name] = np.frombuffer(proof, dtype=np.uint8)
        print(name, len(proof), hashlib.sha256(proof).hexdigest()[:16])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "proofs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))
