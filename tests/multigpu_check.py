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
        for mode in ("peer", "allgather"):
            sf = ShardedEvaluator(f, dist, torch, mode=mode)
            tci, ranks, errors = T.crossinterpolate2(sf, ld, rng=T.CounterRNG(3), **kw)
            sf.release()
            same = ranks == rr and errors == re and all(
                np.array_equal(a, b) for a, b in zip(tci.Iset + tci.Jset, ref.Iset + ref.Jset)) and all(
                np.array_equal(a, b) for a, b in zip(tci.sitetensors, ref.sitetensors))
            ok = ok and same
            if rank == 0:
                print(f"{name}/{mode}: world={world} rank={ranks[-1]} identical_to_single_gpu={same}")
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_OK" if t.item() == 1 else "MULTIGPU_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
