"""The headline step alone (bench.py's MPO Pi at config-5 shape), for quick A/B runs of kernel variants."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tci_b200 as T  # noqa: E402
ctx = T.default_context()
fm = T.Contraction(T.TensorTrain(bench.mpo_cores(5)), T.TensorTrain(bench.mpo_cores(6)), ctx=ctx)
I, J = bench.index_sets(bench.NL)
fl = bench.pi_flops(I, J)
for _ in range(2):
    d, mx = fm.batchevaluate_device(I, J, 0); del d
ctx.timers(reset=True)
for _ in range(3):
    d, mx = fm.batchevaluate_device(I, J, 0); del d
ms = ctx.timers(reset=True)["pi_eval"] / 3
print(f"VARIANT={os.environ.get('TCI_DGEMM_VARIANT','0')}: {ms:.2f} ms  {fl/ms/1e9:.2f} TFLOP/s  max|Pi|={mx:.6e}")
