"""ComplexF64 building blocks on the device: complex DMMA GEMM rate (stage timer of tci_zgemm_host, copies excluded)
and the cooperative complex rrLU (kernel time) -- run under gpurun."""
import sys
import numpy as np
sys.path.insert(0, ".")
import tci_b200 as T

ctx = T.default_context()
rng = np.random.default_rng(0)
for (M, N, K) in ((2048, 2048, 2048), (4096, 4096, 1024), (1024, 65536, 256)):
    A = np.asfortranarray(rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K)))
    B = np.asfortranarray(rng.standard_normal((K, N)) + 1j * rng.standard_normal((K, N)))
    T.zgemm(A, B)
    ctx.timers(reset=True)
    for _ in range(3):
        T.zgemm(A, B)
    ms = ctx.timers()["gemm"] / 3
    print(f"zgemm {M}x{N}x{K}: {ms:.3f} ms  {8.0 * M * N * K / ms / 1e9:.2f} TFLOP/s (real flops, 8 per complex fma)")
for (m, n, r) in ((512, 512, 128), (2048, 2048, 256), (4096, 4096, 256)):
    s = 2.0 ** (-30.0 * np.arange(r) / r)
    A = ((rng.standard_normal((m, r)) + 1j * rng.standard_normal((m, r))) * s) @ (rng.standard_normal((r, n)) + 1j * rng.standard_normal((r, n)))
    T.rrlu(A, maxrank=r, reltol=1e-12)
    ctx.timers(reset=True)
    lu = T.rrlu(A, maxrank=r, reltol=1e-12)
    ms = ctx.timers()["rrlu_kernel"]
    fl = sum(8.0 * (m - k) * (n - k) for k in range(1, lu.npivot + 1))
    by = sum(32.0 * (m - k) * (n - k) for k in range(1, lu.npivot + 1))
    print(f"zrrlu {m}x{n} r={lu.npivot}: {ms:.3f} ms  {ms * 1e3 / max(lu.npivot, 1):.2f} us/pivot  {fl / ms / 1e6:.1f} GFLOP/s  {by / ms / 1e6:.0f} GB/s (read+write model)")
