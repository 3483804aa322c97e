"""CPU, world_size 2 over gloo: the proof-sharding plumbing used for N > 1 GPUs (no data-path collective; ranks prove
disjoint contiguous job ranges and rank 0 collects the bytes).  The per-rank prover here is the oracle on the tiny
circuit -- the GPU path is identical above `create_proofs`."""
import os, socket
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import battlezips_halo2_b200 as bz  # noqa: F401
from battlezips_halo2_b200.sharding import shard_range, gather_proofs


def test_shard_range_partitions():
    for total in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.util_prover import Job, tiny_circuit
    job = Job(*tiny_circuit(5))
    lo, hi = shard_range(total, rank, world)
    local = [job.oracle_proof(index=i) for i in range(lo, hi)]
    proofs = gather_proofs(local, lo, total)
    if rank == 0:
        ok = all(job.verify(p) for p in proofs) and all(p == job.oracle_proof(index=i) for i, p in enumerate(proofs))
        open(out_path, "w").write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_proving(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), 5, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def _msm_worker(rank, world, port, out_path):
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle as co
    from battlezips_halo2_b200.sharding import allgather_point_sum
    import ctypes
    rng = np.random.default_rng(5)
    n = 300
    C = co.CURVES[0][0]
    h = C.hash_to_curve("Halo2-Parameters")
    bases = co.points_to_mont(0, [h(b"\x00" + i.to_bytes(4, "little")) for i in range(n)])
    scalars = co.from_u512(0, rng.integers(0, 2**63, size=(n, 8), dtype=np.uint64))
    lo, hi = shard_range(n, rank, world)
    partial = co.best_multiexp(0, scalars[lo:hi], bases[lo:hi])

    def add(a, b):
        out = np.zeros(8, dtype=np.uint64)
        co.lib().orc_point_add(0, co._p(np.ascontiguousarray(a)), co._p(np.ascontiguousarray(b)), co._p(out))
        return out
    total = allgather_point_sum(partial, lambda j: co.to_affine(0, j), add, device="cpu")
    full = co.to_affine(0, co.best_multiexp(0, scalars, bases))[0]
    if rank == 0:
        open(out_path, "w").write("ok" if np.array_equal(total, full) else "bad")
    dist.barrier()
    dist.destroy_process_group()


def test_point_range_split_msm_allgather(tmp_path):
    """Large-MSM split by point range + all-gather of 96 B partials + local sum == unsplit MSM."""
    out = str(tmp_path / "msm.txt")
    mp.spawn(_msm_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def _commit_worker(rank, world, port, count, out_path):
    """column-sharded commitments: every rank commits its column range with the oracle, the 64 B results are all-gathered"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from oracle import c_oracle as co, halo2 as H
    from battlezips_halo2_b200.sharding import allgather_commitments
    params = H.Params.new(5, 0)
    rng = np.random.default_rng(5)
    cols = co.from_u512(0, rng.integers(0, 2**63, size=(count * 32, 8), dtype=np.uint64)).reshape(count, 32, 4)
    lo, hi = shard_range(count, rank, world)
    per = (count + world - 1) // world
    local = np.zeros((per, 8), dtype=np.uint64)
    for j in range(lo, hi):
        local[j - lo] = co.to_affine(0, params.commit(cols[j], 1 + j))[0]
    got = allgather_commitments(torch.from_numpy(local.view(np.int64)), count, rank, world).numpy().view(np.uint64)
    exp = np.stack([co.to_affine(0, params.commit(cols[j], 1 + j))[0] for j in range(count)])
    ok = got.shape == (count, 8) and np.array_equal(got, exp)
    if rank == 0:
        open(out_path, "w").write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("count", [5, 8])
def test_column_sharded_commitments_allgather(tmp_path, count):
    out = str(tmp_path / "commit.txt")
    mp.spawn(_commit_worker, args=(2, _free_port(), count, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
