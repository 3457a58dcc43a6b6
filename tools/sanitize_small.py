"""Small forced-path factorisations for compute-sanitizer (GPU box only):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

os.environ["TCI_RRLU_NO_RES"] = "1"
os.environ["TCI_RRLU_LAZY_MIN"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402

rng = np.random.default_rng(0)
for m, n, r in ((257, 300, 21), (64, 1200, 9), (900, 130, 13)):
    A = (rng.random((m, r)) * 2.0 ** (-np.arange(r))) @ rng.random((r, n))
    for lo in (True, False):
        lu = T.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo)
        L, U = lu.L, lu.U
        Ap = A[lu.rowpermutation - 1][:, lu.colpermutation - 1]
        print(m, n, r, lo, lu.npivot, float(np.max(np.abs(L @ U - Ap))))
del os.environ["TCI_RRLU_LAZY_MIN"]
A = rng.random((300, 280))
lu = T.rrlu(A, maxrank=40)
print("in place", lu.npivot)
