"""Host mirror of src/globalpivotfinder.jl: the finder interface and the default finder.
The star probes |f(x) - tt(x)| are evaluated on the GPU (K7, tci_globalsearch)."""
import ctypes as C

import numpy as np

from ._lib import lib, pf, pi
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

    def __call__(self, input, f, abstol, verbosity=0, rng=None, mode=0):
        """mode: tci_globalsearch's evaluation mode (0 auto, 1 ordered chain, 2 prefix / suffix environments)."""
        n = len(input.localdims)
        if self.nsearch <= 0 or self.maxnglobalpivot <= 0:
            return np.zeros((0, n), dtype=np.int64)
        if getattr(f, "is_complex", False):  # ComplexF64 target: two device batches + the library's selection
            from .complexf64 import zfind_global_pivots
            return zfind_global_pivots(self, input, f, abstol, rng=rng, verbosity=verbosity)
        ctx = f.ctx
        tt = input.current_tt
        handle = getattr(tt, "device_handle", None)  # the device-resident cores tci_fill_sitetensors left behind
        if handle is None or handle.ctx is not ctx:
            from .cachedtensortrain import TTCache
            handle = TTCache(tt, ctx=ctx)  # uploaded once for this call
        cap = self.maxnglobalpivot
        piv = np.zeros((cap, n), dtype=np.int64)
        errs = np.zeros(cap, dtype=np.float64)
        acc = np.zeros(cap, dtype=np.int64)
        nf = C.c_int64(0)
        if isinstance(rng, CounterRNG):  # the injected generator: the library draws the starts itself, on the device
            rng.calls += 1
            ctx.check(lib().tci_globalsearch_counter(ctx.h, f.id, handle.id, rng.seed & (2**64 - 1), rng.calls, self.nsearch,
                                                     float(abstol) * self.tolmarginglobalsearch, cap, int(mode), pi(piv),
                                                     pf(errs), pi(acc), C.byref(nf)))
        else:
            starts = np.ascontiguousarray(self.draw(input, rng))
            ctx.check(lib().tci_globalsearch(ctx.h, f.id, handle.id, pi(starts), starts.shape[0],
                                             float(abstol) * self.tolmarginglobalsearch, cap, int(mode), pi(piv),
                                             pf(errs), pi(acc), C.byref(nf)))
        if verbosity > 0:
            print(f"Found {nf.value} global pivots")
        self.last_errors = errs[: nf.value].copy()
        self.last_starts = acc[: nf.value].copy()
        return piv[: nf.value].copy()
