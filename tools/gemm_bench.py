"""FP64 GEMM kernel throughput (kernel time from the library's stage timer) next to cuBLAS (torch)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402
from tci_b200 import _lib  # noqa: E402

ctx = T.default_context()
rng = np.random.default_rng(0)
for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 512), (2048, 2048, 256), (1024, 1024, 1024), (512, 512, 512)]:
    A = np.asfortranarray(rng.standard_normal((M, K)))
    B = np.asfortranarray(rng.standard_normal((K, N)))
    C = np.zeros((M, N), order="F")
    for _ in range(2):
        ctx.check(_lib.lib().tci_dgemm_host(ctx.h, 0, 0, M, N, K, 1.0, _lib.pf(A), _lib.pf(B), 0.0, _lib.pf(C)))
    ctx.timers(reset=True)
    for _ in range(3):
        ctx.check(_lib.lib().tci_dgemm_host(ctx.h, 0, 0, M, N, K, 1.0, _lib.pf(A), _lib.pf(B), 0.0, _lib.pf(C)))
    ms = ctx.timers(reset=True)["gemm"] / 3
    err = np.max(np.abs(C[:64, :64] - A[:64] @ B[:, :64])) / np.max(np.abs(C[:64, :64]))
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    for _ in range(2):
        c = a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c = a @ b
    e1.record()
    torch.cuda.synchronize()
    cms = e0.elapsed_time(e1) / 5
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: ours {fl / ms / 1e9:8.1f} GFLOP/s ({ms:.3f} ms, relerr {err:.1e})   cuBLAS {fl / cms / 1e9:8.1f} GFLOP/s")
