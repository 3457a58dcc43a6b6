"""cProfile of crossinterpolate2 on the GPU box: where the host time goes.  usage: python tools/tci_cprofile.py [c1|c3]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tci_b200 as T  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c1":
    ld, f, kw = [10] * 8, T.BuiltinTarget(T.LORENTZ, [1.0], [10] * 8), dict(tolerance=1e-8)
else:
    ld = [4] * 20
    f, kw = T.BuiltinTarget(T.QUANTICS2D, [0, 20], ld), dict(tolerance=1e-10, maxbonddim=256)
T.crossinterpolate2(T.BuiltinTarget(T.LORENTZ, [1.0], [10] * 4), [10] * 4, tolerance=1e-6)  # warm up
pr = cProfile.Profile()
pr.enable()
T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
