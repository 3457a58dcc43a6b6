"""Per-GPU timeline of the sharded contraction Pi (TCI_SHARD_DEBUG=1): python tools/shard_timeline.py NGPU"""
import os, sys, time
os.environ["TCI_SHARD_DEBUG"] = "1"
sys.path.insert(0, ".")
import numpy as np
import tci_b200 as T
import bench as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = T.Context(devices=list(range(n)))
fm = T.Contraction(T.TensorTrain(B.mpo_cores(5)), T.TensorTrain(B.mpo_cores(6)), ctx=ctx)
I, J = B.index_sets(B.NL)
for it in range(3):
    t0 = time.perf_counter()
    dev, mx = fm.batchevaluate_device(I, J, 0)
    print(f"call {it}: {(time.perf_counter() - t0) * 1e3:.2f} ms wall", file=sys.stderr)
    del dev
