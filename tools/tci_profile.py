"""End-to-end crossinterpolate2 timing with the per-stage split (GPU box only).
usage: python tools/tci_profile.py [c1|c3|c3i|c4s]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c1":
    ld, f, kw = [10] * 8, T.BuiltinTarget(T.LORENTZ, [1.0], [10] * 8), dict(tolerance=1e-8)
elif which == "c3":
    ld = [4] * 20
    f, kw = T.BuiltinTarget(T.QUANTICS2D, [0, 20], ld), dict(tolerance=1e-10, maxbonddim=256)
elif which == "c3i":
    ld = [2] * 40
    f, kw = T.BuiltinTarget(T.QUANTICS2D, [1, 20], ld), dict(tolerance=1e-10, maxbonddim=256)
else:
    ld = [64] * 12
    g = np.random.default_rng(4)
    p = np.concatenate([[4], g.integers(1, 1025, 12) / 256.0, g.integers(-512, 513, 4) / 1024.0,
                        (g.integers(-1024, 1025, (4, 12)) / 32.0).flatten()])
    f, kw = T.BuiltinTarget(T.SEPCOS, p, ld), dict(tolerance=1e-12, maxbonddim=int(sys.argv[2]) if len(sys.argv) > 2 else 64, maxiter=3)
ctx = f.ctx
T.crossinterpolate2(T.BuiltinTarget(T.LORENTZ, [1.0], [10] * 4), [10] * 4, tolerance=1e-6)  # warm up
ctx.timers(reset=True)
l0 = ctx.launches
t0 = time.perf_counter()
tci, ranks, errors = T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
dt = time.perf_counter() - t0
tm = ctx.timers(reset=True)
print(f"{which}: time {dt:.3f} s, iterations {len(ranks)}, ranks {ranks}, errors {['%.2e' % e for e in errors]}")
print("stage ms:", {k: round(v, 1) for k, v in tm.items()}, "sum", round(sum(v for k, v in tm.items() if k != 'rrlu_kernel'), 1),
      "launches", ctx.launches - l0, "bond updates", len(tci.trace), "evals", f.nevals)
print("largest Pi:", max((t[1], t[2]) for t in tci.trace), "linkdims", T.linkdims(tci))
