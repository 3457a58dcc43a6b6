"""Debug helper: the bench's N-GPU headline (MPO Pi through one multi-GPU context) in a plain process."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import tci_b200 as T
from tci_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = T.Context(devices=list(range(n)))
_lib.set_default_context(ctx)
fm = T.Contraction(T.TensorTrain(bench.mpo_cores(5)), T.TensorTrain(bench.mpo_cores(6)), ctx=ctx)
I, J = bench.index_sets(bench.NL)
for k in range(4):
    t0 = time.perf_counter()
    dev, mx = fm.batchevaluate_device(I, J, 0)
    del dev
    print("step", k, (time.perf_counter() - t0) * 1e3, "ms", mx, flush=True)
