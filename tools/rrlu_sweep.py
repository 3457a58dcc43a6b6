"""Per-pivot latency / bandwidth sweep of the rrLU kernel over matrix sizes (GPU box only).
usage: python tools/rrlu_sweep.py [size:rank ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402
from bench import lowrank_host, rrlu_bytes, rrlu_flops  # noqa: E402

cases = [tuple(map(int, a.split(":"))) for a in sys.argv[1:]] or [
    (64, 32), (128, 64), (256, 128), (512, 128), (1024, 256), (1536, 256), (2048, 256), (4096, 256)]
ctx = T.default_context()
print(f"{'size':>6} {'rank':>5} {'kernel ms':>10} {'us/pivot':>9} {'GFLOP/s':>9} {'GB/s(alg)':>10}")
for n, r in cases:
    A = lowrank_host(n, n, r, 3)
    mats = [T.DeviceMatrix.from_host(ctx, A) for _ in range(6)]
    for i in range(3):
        T.rrlu(mats[i], maxrank=r, reltol=1e-12)
    ctx.timers(reset=True)
    for i in range(3, 6):
        lu = T.rrlu(mats[i], maxrank=r, reltol=1e-12)
    ms = ctx.timers(reset=True)["rrlu_kernel"] / 3
    print(f"{n:>6} {r:>5} {ms:>10.3f} {1e3 * ms / lu.npivot:>9.2f} {rrlu_flops(n, n, r) / ms / 1e6:>9.1f} "
          f"{rrlu_bytes(n, n, r) / ms / 1e6:>10.1f}")
