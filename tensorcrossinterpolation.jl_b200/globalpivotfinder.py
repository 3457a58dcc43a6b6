"""Host mirror of src/globalpivotfinder.jl: the finder interface and the default finder.
The star probes |f(x) - tt(x)| are evaluated on the GPU (K7, tci_globalsearch)."""
import ctypes as C

import numpy as np

from ._lib import core_ptrs, lib, pf, pi
from .tensortrain import _dims3
from .util import CounterRNG


class GlobalPivotSearchInput:  # globalpivotfinder.jl:15-48
    def __init__(self, localdims, current_tt, maxsamplevalue, Iset, Jset):
        self.localdims = list(localdims)
        self.current_tt = current_tt
        self.maxsamplevalue = maxsamplevalue
        self.Iset = Iset
        self.Jset = Jset


class AbstractGlobalPivotFinder:  # :56-87
    def __call__(self, input, f, abstol, verbosity=0, rng=None):
        raise NotImplementedError(f"find_global_pivots not implemented for {type(self)}")


class DefaultGlobalPivotFinder(AbstractGlobalPivotFinder):  # :100-195
    def __init__(self, nsearch=5, maxnglobalpivot=5, tolmarginglobalsearch=10.0):
        self.nsearch = nsearch
        self.maxnglobalpivot = maxnglobalpivot
        self.tolmarginglobalsearch = tolmarginglobalsearch

    def draw(self, input, rng):
        """initial_points of :156.  rng: CounterRNG (shared with the oracle), a numpy Generator, or an
        explicit (nsearch x n) array of start points."""
        if isinstance(rng, CounterRNG):
            return rng.start_points(self.nsearch, input.localdims)
        if isinstance(rng, np.ndarray):
            return np.ascontiguousarray(rng, dtype=np.int64)
        rng = rng or np.random.default_rng()
        return np.stack([rng.integers(1, d + 1, self.nsearch) for d in input.localdims], axis=1).astype(np.int64)

    def __call__(self, input, f, abstol, verbosity=0, rng=None):
        n = len(input.localdims)
        if self.nsearch <= 0 or self.maxnglobalpivot <= 0:
            return np.zeros((0, n), dtype=np.int64)
        starts = np.ascontiguousarray(self.draw(input, rng))
        keep, arr = core_ptrs([c.reshape((c.shape[0], -1, c.shape[-1]), order="F")
                               for c in input.current_tt.sitetensors])
        d3 = _dims3(keep)
        world, rank = getattr(f, "world", 1), getattr(f, "rank", 0)
        mine = np.arange(rank, starts.shape[0], world)  # independent starts, dealt round-robin
        local = np.ascontiguousarray(starts[mine])
        # the reference collects every accepted point and then truncates (:186-188)
        cap = max(len(mine), 1) if world > 1 else self.maxnglobalpivot
        piv = np.zeros((cap, n), dtype=np.int64)
        errs = np.zeros(cap, dtype=np.float64)
        acc = np.zeros(cap, dtype=np.int64)
        nf = C.c_int64(0)
        ctx = f.ctx
        if len(mine):
            ctx.check(lib().tci_globalsearch(ctx.h, f.id, n, pi(d3), arr, pi(local), local.shape[0],
                                             float(abstol) * self.tolmarginglobalsearch, cap, pi(piv), pf(errs),
                                             pi(acc), C.byref(nf)))
        if world > 1:
            from .parallel import allgather_candidates, select_global_pivots
            cands = [(int(mine[acc[q]]), piv[q].tolist(), float(errs[q])) for q in range(nf.value)]
            pts, es = select_global_pivots(allgather_candidates(f.dist, cands, f.group), self.maxnglobalpivot)
            self.last_errors = np.asarray(es, dtype=np.float64)
            if verbosity > 0:
                print(f"Found {len(pts)} global pivots")
            return np.asarray(pts, dtype=np.int64).reshape(len(pts), n)
        if verbosity > 0:
            print(f"Found {nf.value} global pivots")
        self.last_errors = errs[: nf.value].copy()
        return piv[: nf.value].copy()
