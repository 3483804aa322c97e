"""Worker of tests/test_gpu_multi.py, launched by `python -m torch.distributed.run --nproc-per-node N` (one rank per GPU, NCCL):
  * one large proof across the GPUs (bz_ctx_set_sharding): the MSMs of every commitment batch dealt out by column / split by
    point range, h(X) evaluated by point range (second pass), exchanged by NCCL all-gather -- the proof bytes must equal the single-GPU proof and the oracle-checked verifier
    must accept;
  * two contexts on two devices inside ONE process (per-device kernel attributes): a 2^13 NTT (64 KB of dynamic
    shared memory) and a table MSM on the second device give the first device's results."""
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import battlezips_halo2_b200 as bz
    from battlezips_halo2_b200 import arithmetic as ar
    from battlezips_halo2_b200.plonk import prover as PR
    from battlezips_halo2_b200.circuits import board_circuit_scaled
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = bz.Context(local, stream=stream.cuda_stream)
    k = 13
    cs, cfg, asg = board_circuit_scaled(k)
    fx = np.load(os.path.join(ROOT, "tests", "golden", f"params_vesta_k{k}.npz"))
    advice = np.stack([PR.mont(col) for col in asg.advice])
    rng = np.random.default_rng(5)
    for general in ("1", "0"):             # the k >= 20 commitment path (bucket MSM over the raw bases), then the table path
        os.environ["BZ_FORCE_GENERAL_MSM"] = general
        params = PR.Params(ctx, k, fx["g"], fx["g_lagrange"], fx["w"], fx["u"], window_bits=8)
        pk = PR.ProvingKey(ctx, params, cs.to_ir(), asg.fixed, asg.permutation_mapping(), 0x1234)
        wide = rng.integers(0, 2**63, size=(pk.num_random, 8), dtype=np.uint64)
        dist.broadcast_object_list(bl := [wide if rank == 0 else None], src=0)
        wide = bl[0]
        single = PR.create_proofs(pk, [asg.instance], advice[None], wide[None])[0]
        # second pass: exchange buffers large enough for the point-range split of h(X) (8n + 4n + 2n points of 32 B, per rank)
        ctx.set_sharding(rank, world, capacity=(1 << 16) if general == "1" else 14 * (1 << k) * 32 // world + (1 << 16))
        sharded = PR.create_proofs(pk, [asg.instance], advice[None], wide[None])[0]
        ctx.set_sharding(0, 1)
        again = PR.create_proofs(pk, [asg.instance], advice[None], wide[None])[0]
        ok = sharded == single == again and PR.verify_proofs(pk, [asg.instance], [sharded]) == [True]
        flags = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        assert int(flags.item()) == 1, f"rank {rank}: sharded proof differs from the single-GPU proof (general={general})"
        pk.close(); params.close()
    if rank == 0 and torch.cuda.device_count() >= 2:
        other = bz.Context(1 - local)
        a = rng.integers(0, 2**62, size=(1 << 13, 4), dtype=np.uint64)
        a[:, 3] &= (1 << 61) - 1
        F_ROOT = 0x2bce74deac30ebda362120830561f81aea322bf2b7bb7584bdad6fabd87ea32f
        P = PR.FP
        om = PR.mont([pow(F_ROOT, 1 << (32 - 13), P)])[0]
        assert np.array_equal(ar.best_fft(ctx, 0, a, om, 13), ar.best_fft(other, 0, a, om, 13)), "NTT differs on the second device"
        os.environ.pop("BZ_FORCE_GENERAL_MSM")
        p0 = PR.Params(ctx, 11, *[np.load(os.path.join(ROOT, "tests", "golden", "params_vesta_k11.npz"))[n] for n in ("g", "g_lagrange", "w", "u")], window_bits=8)
        p1 = PR.Params(other, 11, *[np.load(os.path.join(ROOT, "tests", "golden", "params_vesta_k11.npz"))[n] for n in ("g", "g_lagrange", "w", "u")], window_bits=8)
        poly = a[: 1 << 11]
        assert np.array_equal(p0.commit(poly, a[0], lagrange=True), p1.commit(poly, a[0], lagrange=True)), "table MSM differs on the second device"
        p0.close(); p1.close(); other.close()
    dist.barrier()
    if rank == 0:
        print("multi-gpu ok: sharded proof == single-GPU proof on", world, "GPUs")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
