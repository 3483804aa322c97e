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
