"""Round-2 kernels under compute-sanitizer (GPU box only): the cooperative complex rrLU, complex LUCI / solve, the complex
contraction chains + GEMM, the CachedFunction memo (concurrent inserts of one key), split-K DGEMM.
   compute-sanitizer --tool memcheck  python tools/sanitize_round2.py
   compute-sanitizer --tool racecheck python tools/sanitize_round2.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402

rng = np.random.default_rng(0)


def crand(*s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


for m, n, r in ((130, 257, 17), (700, 90, 13), (33, 600, 9)):
    A = (crand(m, r) * 2.0 ** (-2.0 * np.arange(r))) @ crand(r, n)
    for lo in (True, False):
        luci = T.MatrixLUCI(A, maxrank=r, reltol=1e-12, leftorthogonal=lo)
        L, R = luci.left(), luci.right()
        print("zrrlu", m, n, r, lo, luci.npivot, float(np.max(np.abs(L @ R - A))))
P, B = crand(21, 21), crand(50, 21)
print("zrdiv", float(np.max(np.abs(T.rrlu(P, reltol=0.0, abstol=0.0).rdiv(B) - B @ np.linalg.inv(P)))))
bonds = [1, 2, 3, 2, 1]
a = [np.asfortranarray(crand(bonds[i], 2, 3, bonds[i + 1])) for i in range(4)]
b = [np.asfortranarray(crand(bonds[i], 3, 2, bonds[i + 1])) for i in range(4)]
f = T.ZContraction(a, b)
print("zpi", f([[1], [2], [3]], [[1, 2], [4, 4]], 1).shape, abs(f([1, 2, 3, 4])))
g = T.BuiltinTarget(T.SUM, [], [2] * 5)
cf = T.CachedFunction(g, capacity_log2=6)
left, right = [[1, 1]] * 40, [[1, 1]] * 40
print("cache", float(cf(left, right, 1).sum()), cf.stats(), float(cf(left, right, 1).sum()), cf.stats())
A, Bm = rng.standard_normal((70, 5000)), rng.standard_normal((5000, 90))
print("splitk", float(np.max(np.abs(T._lib.gemm(A, Bm) - A @ Bm))))
