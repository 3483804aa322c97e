"""Multi-GPU partitioning of the prover workload: independent proofs are sharded as contiguous index ranges,
one process per GPU, with NO data-path collective -- ranks only exchange finished proof bytes (and timing maxima).
Independent column commitments of ONE large proof shard the same way (column ranges per GPU, the 64 B results all-gathered);
one large MSM is split by point range with an all-gather of the 96 B partials.  NTT / grand product / IPA of a single proof
do not shard: replicas only."""


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


def allgather_point_sum(local_jac, normalize, add, device=None, group=None):
    """The one real exchange on this path: a large MSM is split by POINT RANGE, every rank reduces its
    range to one Jacobian partial (96 B) and the partials are all-gathered (NCCL has no elliptic-curve reduction op)
    and summed locally: G - 1 point additions.

    local_jac : numpy (12,) uint64 Jacobian partial of this rank
    normalize : callable (G,12) uint64 Jacobian -> (G,8) uint64 affine       (GPU: bz_batch_normalize_dev)
    add       : callable ((8,), (8,)) affine -> (8,) affine                   (GPU: bz_curve_op add)
    device    : torch device of the exchange tensors ("cuda" for NCCL, "cpu" for gloo)
    Returns the affine sum (8,) uint64 on every rank."""
    import numpy as np
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(local_jac, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    jac = np.stack([p.cpu().numpy().view(np.uint64) for p in parts])
    aff = normalize(jac)
    acc = aff[0]
    for i in range(1, world):
        acc = add(acc, aff[i])
    return acc


def allgather_point_sum_dev(ctx, curve, jac_tensor, gathered, out_affine, group=None):
    """Device-resident form of `allgather_point_sum` for the NCCL path: `jac_tensor` (12 int64 on the GPU, written by
    bz_msm_dev), `gathered` (world x 12) and `out_affine` (8) are torch CUDA tensors allocated once by the caller.  One
    NCCL all-gather of 96 B per rank, then one kernel (bz_point_sum_dev) adds the partials -- no host round trip."""
    import ctypes
    import torch.distributed as dist
    dist.all_gather_into_tensor(gathered, jac_tensor, group=group)
    ctx._check(ctx.lib.bz_point_sum_dev(ctx.h, curve, ctypes.c_void_p(gathered.data_ptr()), gathered.shape[0], ctypes.c_void_p(out_affine.data_ptr())))
    return out_affine


def allgather_commitments(local, count, rank, world, group=None):
    """Column-sharded commitments: rank r computed the affine commitments of columns shard_range(count, r, world) into
    `local` (a torch int64 tensor of shape (ceil(count / world), 8), rows beyond its share are padding).  One all-gather of
    64 B per column; returns the (count, 8) tensor in column order on every rank."""
    import torch
    import torch.distributed as dist
    per = (count + world - 1) // world
    assert local.shape == (per, 8)
    if world == 1:
        return local[:count]
    flat = torch.empty((world * per, 8), dtype=local.dtype, device=local.device)      # concatenation layout (gloo and NCCL)
    dist.all_gather_into_tensor(flat, local.contiguous(), group=group)
    gathered = flat.view(world, per, 8)
    parts = []
    for r in range(world):
        lo, hi = shard_range(count, r, world)
        parts.append(gathered[r, :hi - lo])
    return torch.cat(parts)
