"""Config-5-shape MPO x MPO Pi (nL = nR = 1024) alone: single GPU, or sharded by row / column blocks under torchrun.
usage: python tools/mpo_shard_bench.py   |   python -m torch.distributed.run --nproc-per-node N ... tools/mpo_shard_bench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import tci_b200 as T  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if dist else 0
ctx = T.default_context()
out = bench.extra_mpo_1024(T, ctx, torch, dist, rank, world)
out.update(bench.extra_globalsearch(T, ctx, torch, dist, rank, world))
if rank == 0:
    print(json.dumps(out))
if dist:
    dist.destroy_process_group()
