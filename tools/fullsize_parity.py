"""Full-size parity of crossinterpolate2 against the CPU oracle for BASELINE configs 1, 3 (fused) and 4 (GPU box).
Prints the deviations the full-size tests in tests/test_gpu_parity.py assert on."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (checker)

g = np.random.default_rng(4)
p4 = np.concatenate([[4], g.integers(1, 1025, 12) / 256.0, g.integers(-512, 513, 4) / 1024.0,
                     (g.integers(-1024, 1025, (4, 12)) / 32.0).flatten()])
CASES = [("c1", T.LORENTZ, [1.0], [10] * 8, dict(tolerance=1e-8)),
         ("c3", T.QUANTICS2D, [0, 20], [4] * 20, dict(tolerance=1e-10, maxbonddim=256)),
         ("c4", T.SEPCOS, p4, [64] * 12, dict(tolerance=1e-12, maxbonddim=512))]
for name, kind, params, ld, kw in CASES:
    f = T.BuiltinTarget(kind, params, ld)
    o = orc.Target.builtin(kind, params, ld)
    t0 = time.perf_counter()
    tci, ranks, errors = T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
    t1 = time.perf_counter()
    res = orc.crossinterpolate2(o, ld, seed=1, **kw)
    t2 = time.perf_counter()
    same_sets = all([tuple(x) for x in tci.Iset[b].tolist()] == res.Iset[b] and
                    [tuple(x) for x in tci.Jset[b].tolist()] == res.Jset[b] for b in range(len(ld)))
    dev_T = max(np.max(np.abs(tci.sitetensors[b] - res.sitetensors[b])) / max(1.0, np.max(np.abs(res.sitetensors[b])))
                for b in range(len(ld)))
    rng = np.random.default_rng(0)
    pts = np.stack([rng.integers(1, d + 1, 2000) for d in ld], axis=1).astype(np.int64)
    fv = f.evaluate_points(pts)
    tg = T.evaluate_points(T.TensorTrain(tci.sitetensors), pts)
    to = np.array([orc.tt_evaluate(res.sitetensors, q) for q in pts])
    scale = tci.maxsamplevalue
    print(f"{name}: gpu {t1 - t0:.3f} s oracle {t2 - t1:.2f} s ranks {ranks} == {res.ranks.tolist()}: "
          f"{[int(r) for r in ranks] == res.ranks.tolist()} sets identical {same_sets} "
          f"errors rel dev {np.max(np.abs(np.array(errors) - res.errors) / np.abs(res.errors)):.2e} "
          f"max site-tensor dev {dev_T:.2e} sum rel dev {abs(T.tci_sum(tci) - res.sum()) / abs(res.sum()):.2e} "
          f"sampled |tt_gpu - tt_oracle|/maxsample {np.max(np.abs(tg - to)) / scale:.2e} "
          f"sampled error gpu {np.max(np.abs(fv - tg)) / scale:.2e} oracle {np.max(np.abs(fv - to)) / scale:.2e}")
