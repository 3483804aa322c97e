"""Multi-GPU partitioning of the prover workload: independent proofs are sharded as contiguous index ranges,
one process per GPU, with NO data-path collective -- ranks only exchange finished proof bytes (and timing maxima).
NTT / grand product / IPA of a single proof do not shard: replicas only."""


def shard_range(total, rank, world):
    """Contiguous, balanced [lo, hi) slice of `total` jobs for `rank` of `world` (first `total % world` ranks get +1)."""
    assert 0 <= rank < world
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_proofs(local_proofs, lo, total, group=None):
    """All ranks contribute {job index: proof bytes}; rank 0 returns the ordered list of `total` proofs, others None.
    Uses torch.distributed object gather (control plane only: a Shot proof is 4.6 KB)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    payload = {lo + i: p for i, p in enumerate(local_proofs)}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    assert sorted(merged) == list(range(total)), "shards do not cover the job range exactly once"
    return [merged[i] for i in range(total)]
