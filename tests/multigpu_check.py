"""Multi-GPU parity check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/multigpu_check.py
Every rank drives the same crossinterpolate2 with Pi evaluation and global search sharded over the
ranks (NCCL); the result must be identical to the unsharded single-GPU run of the same rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import tci_b200 as T
    from tci_b200.parallel import ShardedEvaluator
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for name, kind, params, ld, kw in [
        ("lorentz", T.LORENTZ, [1.0], [10] * 6, dict(tolerance=1e-8)),
        ("quantics2d", T.QUANTICS2D, [0, 8], [4] * 8, dict(tolerance=1e-9, maxbonddim=40, maxiter=6)),
    ]:
        f = T.BuiltinTarget(kind, params, ld)
        ref, rr, re = T.crossinterpolate2(f, ld, rng=T.CounterRNG(3), **kw)
        for mode, shard in (("peer", "cols"), ("allgather", "cols"), ("peer", "rows")):
            sf = ShardedEvaluator(f, dist, torch, mode=mode, shard=shard)
            tci, ranks, errors = T.crossinterpolate2(sf, ld, rng=T.CounterRNG(3), **kw)
            sf.release()
            same = ranks == rr and errors == re and all(
                np.array_equal(a, b) for a, b in zip(tci.Iset + tci.Jset, ref.Iset + ref.Jset)) and all(
                np.array_equal(a, b) for a, b in zip(tci.sitetensors, ref.sitetensors))
            ok = ok and same
            if rank == 0:
                print(f"{name}/{mode}/{shard}: world={world} rank={ranks[-1]} identical_to_single_gpu={same}")
    # MPO x MPO contraction target (config-5 family, reduced): Pi sharded by row blocks of the left index set, every
    # rank's block stored into rank 0's HBM; the assembled Pi must equal the unsharded evaluation to 1e-12 (the GEMM tiling
    # may depend on the block height), and crossinterpolate2 driven through the sharded evaluator must pick the same pivots.
    g = np.random.default_rng(11)
    ns, D = 8, 12
    bonds = [1] + [D] * (ns - 1) + [1]
    A = [np.asfortranarray(g.random((bonds[i], 2, 2, bonds[i + 1])) * 2 - 1) for i in range(ns)]
    B = [np.asfortranarray(g.random((bonds[i], 2, 2, bonds[i + 1])) * 2 - 1) for i in range(ns)]
    fm = T.Contraction(T.TensorTrain(A), T.TensorTrain(B))
    I = np.stack([g.integers(1, 5, 150) for _ in range(4)], axis=1).astype(np.int64)
    J = np.stack([g.integers(1, 5, 130) for _ in range(4)], axis=1).astype(np.int64)
    full, mx0 = fm.batchevaluate_device(I, J, 0)
    for shard in ("rows", "cols"):
        sm = ShardedEvaluator(fm, dist, torch, mode="peer", shard=shard)
        view, mx = sm.batchevaluate_device(I, J, 0)
        same = abs(mx - mx0) <= 1e-12 * mx0 and (
            rank != 0 or np.max(np.abs(view.to_host() - full.to_host())) <= 1e-12 * mx0)
        ok = ok and same
        if rank == 0:
            print(f"mpo Pi {I.shape[0]}x{J.shape[0]}/peer/{shard}: world={world} identical_to_single_gpu={same}")
        del view
        ld = fm.localdims
        ref, rr, re = T.crossinterpolate2(fm, ld, rng=T.CounterRNG(5), tolerance=1e-8, maxbonddim=30, maxiter=4)
        tci, ranks, errors = T.crossinterpolate2(sm, ld, rng=T.CounterRNG(5), tolerance=1e-8, maxbonddim=30, maxiter=4)
        same = ranks == rr and np.allclose(errors, re, rtol=1e-6, atol=0) and all(
            np.array_equal(a, b) for a, b in zip(tci.Iset + tci.Jset, ref.Iset + ref.Jset))
        ok = ok and same
        if rank == 0:
            print(f"mpo crossinterpolate2/peer/{shard}: world={world} rank={ranks[-1]} identical_to_single_gpu={same}")
        sm.release()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_OK" if t.item() == 1 else "MULTIGPU_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
