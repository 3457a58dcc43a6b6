"""Small runs of the tiled / bucket-ordered TT chain, the environment ABI and a global search for compute-sanitizer
(GPU box only):   compute-sanitizer --tool memcheck python tools/sanitize_tt.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402

rng = np.random.default_rng(0)
dims = [3, 5, 1, 4, 6, 2]
bonds = [1, 9, 70, 131, 130, 17, 1]
cores = [np.asfortranarray(rng.random((bonds[i], dims[i], bonds[i + 1])) - 0.3) for i in range(len(dims))]
tt = T.TensorTrain(cores)
pts = np.stack([rng.integers(1, d + 1, 1003) for d in dims], axis=1).astype(np.int64)
a = T.evaluate_points(tt, pts)
os.environ["TCI_TT_NO_BUCKETS"] = "1"
b = T.evaluate_points(tt, pts)
del os.environ["TCI_TT_NO_BUCKETS"]
print("tiled chain", np.array_equal(a, b))
f = T.TTCache(tt)
nl = 3
I, J = pts[:37, :nl], pts[:21, nl:]
lenv = T.DeviceMatrix.empty(f.ctx, f.env_dim(0, nl), len(I))
renv = T.DeviceMatrix.empty(f.ctx, f.env_dim(1, len(dims) - nl), len(J))
f.env_eval_into(lenv, 0, 0, I)
f.env_eval_into(renv, 0, 1, J)
out = T.DeviceMatrix.empty(f.ctx, len(I), len(J))
mx = f.pi_from_envs(lenv, 0, len(I), renv, 0, len(J), out, 0)
print("envs", float(np.max(np.abs(out.to_host() - f(I, J, 0)))), mx)
g = T.BuiltinTarget(T.LORENTZ, [1.0], dims)
finder = T.DefaultGlobalPivotFinder(nsearch=40, maxnglobalpivot=5)
found = finder(T.GlobalPivotSearchInput(dims, tt, 1.0, None, None), g, 1e-3, rng=T.CounterRNG(1))
print("global search", len(found))
h = T.BuiltinTarget(T.QUANTICS2D, [0, 8], [4] * 8)
for nI, nJ in ((5, 3), (70, 33), (300, 129)):
    Ih = np.stack([rng.integers(1, 5, nI) for _ in range(4)], axis=1).astype(np.int64)
    Jh = np.stack([rng.integers(1, 5, nJ) for _ in range(4)], axis=1).astype(np.int64)
    print("pi", h(Ih, Jh, 0).shape)
