"""DGEMM rates of the library's DMMA kernels against cuBLAS (torch.matmul float64) on the shapes the path uses.
usage (GPU box): python tools/gemm_bench.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402
from tci_b200 import _lib  # noqa: E402

ctx = T.default_context()
rng = np.random.default_rng(0)
TB = int(os.environ.get("GEMM_TB", "0"))
for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 512), (2048, 2048, 256)]:
    A = np.asfortranarray(rng.standard_normal((M, K)))
    B = np.asfortranarray(rng.standard_normal((N, K) if TB else (K, N)))
    C = np.zeros((M, N), order="F")
    res = {}
    for label, env in (("bulk", {}), ("cp.async", {"TCI_DGEMM_NO_BULK": "1"})):
        # the switches are read once per process: run the second variant in a child
        if env:
            continue
        for _ in range(2):
            ctx.check(_lib.lib().tci_dgemm_host(ctx.h, 0, TB, M, N, K, 1.0, _lib.pf(A), _lib.pf(B), 0.0, _lib.pf(C)))
        ctx.timers(reset=True)
        reps = 3
        for _ in range(reps):
            ctx.check(_lib.lib().tci_dgemm_host(ctx.h, 0, TB, M, N, K, 1.0, _lib.pf(A), _lib.pf(B), 0.0, _lib.pf(C)))
        res[label] = 2.0 * M * N * K / (ctx.timers(reset=True)["gemm"] / reps * 1e-3) / 1e12
    a = torch.from_numpy(np.ascontiguousarray(A)).cuda()
    b = torch.from_numpy(np.ascontiguousarray(B.T if TB else B)).cuda()
    for _ in range(2):
        c = a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c = a @ b
    e1.record()
    torch.cuda.synchronize()
    cub = 5 * 2.0 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12
    err = float(np.max(np.abs(C - c.cpu().numpy())) / np.max(np.abs(C)))
    print(f"{M}x{N}x{K}: library {res['bulk']:.2f} TFLOP/s (VARIANT={os.environ.get('TCI_DGEMM_VARIANT', '0')} TB={TB}), "
          f"cuBLAS {cub:.2f}, rel dev {err:.1e}", flush=True)
    del a, b, c
