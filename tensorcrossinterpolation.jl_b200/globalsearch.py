"""Host mirror of src/globalsearch.jl:18-94: estimatetrueerror and the greedy coordinate ascent
_floatingzone.  The M = 1 fills of the target and of the tensor train are evaluated on the GPU
(tci_pi_eval on the registered target and on a TTCache of the train)."""
import numpy as np

from .cachedtensortrain import TTCache
from .tensortrain import TensorTrain
from .util import CounterRNG, jl_max


def _floatingzone(ttcache, f, earlystoptol=float("inf"), nsweeps=2**62, initp=None, rng=None):
    """_floatingzone (globalsearch.jl:43-94): for every site in turn move to the local index with the largest
    |f - tt| (first maximum), until the maximal error stops changing or exceeds earlystoptol."""
    if nsweeps <= 0:
        raise RuntimeError("nsweeps should be positive!")
    localdims = ttcache.localdims
    n = len(localdims)
    if initp is None:
        if isinstance(rng, CounterRNG):  # the package-wide injected generator (optimize passes it to the finders)
            pivot = [int(v) for v in rng.start_points(1, localdims)[0]]
        else:
            rng = rng or np.random.default_rng()
            pivot = [int(rng.integers(1, d + 1)) for d in localdims]
    else:
        pivot = [int(x) for x in initp]
    maxerror = abs(f(pivot) - ttcache(pivot))
    sweeps = 0
    while sweeps < nsweeps:
        sweeps += 1
        prev = maxerror
        for ipos in range(n):
            left = np.asarray([pivot[:ipos]], dtype=np.int64).reshape(1, ipos)
            right = np.asarray([pivot[ipos + 1:]], dtype=np.int64).reshape(1, n - ipos - 1)
            exact = f(left, right, 1).reshape(-1)
            pred = ttcache(left, right, 1).reshape(-1)
            err = np.abs(exact - pred)
            pivot[ipos] = int(np.argmax(err)) + 1  # argmax: first maximum
            maxerror = jl_max(float(np.max(err)), maxerror)  # Julia's max propagates NaN
        if maxerror == prev or maxerror > earlystoptol:
            break
    return pivot, maxerror


def estimatetrueerror(tt, f, nsearch=100, initialpoints=None, rng=None):
    """estimatetrueerror(tt, f; nsearch, initialpoints) (globalsearch.jl:18-40): list of (pivot, error)
    sorted by descending error, duplicates removed."""
    if nsearch <= 0 and initialpoints is None:
        raise RuntimeError("No search is performed")
    if initialpoints is None:
        if isinstance(rng, CounterRNG):
            cores = tt.sitetensors
            initialpoints = rng.start_points(nsearch, [int(np.prod(c.shape[1:-1])) for c in cores]).tolist()
        else:
            rng = rng or np.random.default_rng()
            initialpoints = [[int(rng.integers(1, int(np.prod(c.shape[1:-1])) + 1)) for c in tt.sitetensors]
                             for _ in range(nsearch)]
    if not isinstance(tt, TensorTrain):
        tt = TensorTrain(tt.sitetensors)
    ttcache = TTCache(tt, ctx=f.ctx)
    found = [_floatingzone(ttcache, f, initp=p) for p in initialpoints]
    order = sorted(range(len(found)), key=lambda i: -found[i][1])  # sortperm(..., rev=true) is stable
    out, seen = [], set()
    for i in order:
        key = (tuple(found[i][0]), found[i][1])
        if key not in seen:
            seen.add(key)
            out.append((list(found[i][0]), found[i][1]))
    return out
