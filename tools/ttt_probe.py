"""time-to-tolerance probe for configs 1 and 3 with the stage timers and (TCI_FILL_DEBUG=1) the fillsitetensors phases."""
import os, sys, time
sys.path.insert(0, ".")
import tci_b200 as T
ctx = T.default_context()
for name, kind, params, ld, kw in (("config1", T.LORENTZ, [1.0], [10] * 8, dict(tolerance=1e-8)),
                                   ("config3", T.QUANTICS2D, [0, 20], [4] * 20, dict(tolerance=1e-10, maxbonddim=256))):
    f = T.BuiltinTarget(kind, params, ld)
    T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
    best = 1e9
    for _ in range(3):
        ctx.timers(reset=True)
        l0 = ctx.launches
        t0 = time.perf_counter()
        tci, ranks, errors = T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
        best = min(best, time.perf_counter() - t0)
    print(name, f"{best * 1e3:.2f} ms", "rank", ranks[-1], "launches", ctx.launches - l0,
          {k: round(v, 2) for k, v in ctx.timers().items()}, file=sys.stderr)
