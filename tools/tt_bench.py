"""Pi evaluation throughput for a tensor-train target (TTCache path, K4)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402

rng = np.random.default_rng(0)
for (n, d, D, nI) in [(20, 4, 256, 1024), (12, 64, 128, 2048), (40, 2, 256, 512)]:
    bonds = [1] + [D] * (n - 1) + [1]
    cores = [np.asfortranarray(rng.standard_normal((bonds[i], d, bonds[i + 1])) / np.sqrt(D)) for i in range(n)]
    f = T.TTCache(T.TensorTrain(cores))
    ctx = f.ctx
    nl = n // 2
    I = np.stack([rng.integers(1, d + 1, nI) for _ in range(nl)], axis=1).astype(np.int64)
    J = np.stack([rng.integers(1, d + 1, nI) for _ in range(n - nl)], axis=1).astype(np.int64)
    for M, (Iu, Ju) in ((0, (I, J)), (2, (I[:, :-1], J[:, 1:]))):
        dev, mx = f.batchevaluate_device(Iu, Ju, M)
        del dev
        ctx.timers(reset=True)
        l0 = ctx.launches
        t0 = time.perf_counter()
        dev, mx = f.batchevaluate_device(Iu, Ju, M)
        dt = time.perf_counter() - t0
        ms = ctx.timers(reset=True)["pi_eval"]
        rows = dev.shape[0]
        env = 2.0 * nI * D * D * (Iu.shape[1] - 1 + Ju.shape[1] - 1)
        fin = 2.0 * rows * D * nI + (2.0 * nI * D * d * D + 2.0 * nI * d * D * d * D if M == 2 else 0)
        print(f"n={n} d={d} D={D} nI=nJ={nI} M={M}: Pi {rows}x{nI}  {ms:.3f} ms (wall {dt * 1e3:.2f}), "
              f"{(env + fin) / ms / 1e9:.2f} TFLOP/s (env {env / 1e9:.2f} GF, product {fin / 1e9:.2f} GF), launches {ctx.launches - l0}")
        del dev
