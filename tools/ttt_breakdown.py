"""Wall-clock breakdown of crossinterpolate2 by driver stage (monkey-patched timers around the host mirror's functions)."""
import sys, time
sys.path.insert(0, ".")
import tci_b200 as T
from tci_b200 import tensorci2 as M
acc = {}
def wrap(name):
    fn = getattr(M, name)
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    setattr(M, name, w)
for nm in ("sweep2site", "fillsitetensors", "sweep1site", "addglobalpivots"):
    wrap(nm)
orig_finder = M.DefaultGlobalPivotFinder.__call__
def finder(self, *a, **k):
    t0 = time.perf_counter()
    try:
        return orig_finder(self, *a, **k)
    finally:
        acc["finder"] = acc.get("finder", 0.0) + time.perf_counter() - t0
M.DefaultGlobalPivotFinder.__call__ = finder
for name, kind, params, ld, kw in (("config1", T.LORENTZ, [1.0], [10] * 8, dict(tolerance=1e-8)),
                                   ("config3", T.QUANTICS2D, [0, 20], [4] * 20, dict(tolerance=1e-10, maxbonddim=256))):
    f = T.BuiltinTarget(kind, params, ld)
    M.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
    acc.clear()
    t0 = time.perf_counter()
    M.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
    tot = time.perf_counter() - t0
    inner = acc.get("sweep2site", 0) - acc.get("fillsitetensors", 0)
    print(name, f"total {tot*1e3:.1f} ms | sweep2site bonds {inner*1e3:.1f} | fillsitetensors {acc.get('fillsitetensors',0)*1e3:.1f} | "
          f"finder {acc.get('finder',0)*1e3:.1f} | addglobalpivots {acc.get('addglobalpivots',0)*1e3:.1f} | sweep1site {acc.get('sweep1site',0)*1e3:.1f}",
          file=sys.stderr)
