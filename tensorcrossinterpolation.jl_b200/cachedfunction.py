"""Host mirror of src/cachedfunction.jl: CachedFunction as a device-resident memo (tci_target_cached, csrc/cache.cu).

The reference keeps a Dict{UInt128, V} on the host; here the table lives in HBM next to the kernels that fill Pi, so a
cached target is evaluated, looked up and stored without leaving the device.  Keys are the reference's:
key(x) = sum_n coeffs[n] * (x_n - 1), coeffs[n] = prod_{m<n} localdims[m] (cachedfunction.jl:14-17, 177-184)."""
import ctypes as C

import numpy as np

from ._lib import lib, pi
from .batcheval import BatchEvaluator


class CachedFunction(BatchEvaluator):
    """CachedFunction{Float64,UInt128}(f, localdims) (cachedfunction.jl:8-63); f is a device target (BatchEvaluator)."""

    def __init__(self, f, localdims=None, capacity_log2=22):
        if not isinstance(f, BatchEvaluator):
            raise TypeError("CachedFunction: f must be a device target (a host closure cannot run on the GPU)")
        localdims = list(f.localdims) if localdims is None else [int(d) for d in localdims]
        if localdims != list(f.localdims):
            raise ValueError("CachedFunction: localdims must be those of the wrapped target")
        coeffs, c = [], 1
        for d in localdims:  # :14-17
            coeffs.append(c)
            c *= d
        if not sum(cf * (d - 1) for cf, d in zip(coeffs, localdims)) < 2 ** 128 - 1:  # :22-24
            raise RuntimeError("Overflow in CachedFunction. Use ValueType = a bigger type with fixed size, e.g., "
                               "BitIntegers.UInt256")
        tid = C.c_int64(0)
        f.ctx.check(lib().tci_target_cached(f.ctx.h, f.id, int(capacity_log2), C.byref(tid)))
        super().__init__(f.ctx, tid.value, localdims)
        self.f = f  # kept alive: the device wrapper refers to it
        self.coeffs = coeffs

    def _key(self, indexset):  # :177-184
        if len(indexset) != len(self.coeffs):
            raise RuntimeError("Invalid length of indexset")
        return sum(c * (int(x) - 1) for c, x in zip(self.coeffs, indexset))

    encodecachekey = _key

    def decodecachekey(self, key):  # :188-196
        index = []
        for d in self.localdims:
            key, r = divmod(key, d)
            index.append(r + 1)
        return index

    def stats(self):
        """{entries: length(cf.cache), hits, misses, unstored} (tci_target_cache_stats)."""
        out = np.zeros(4, dtype=np.int64)
        self.ctx.check(lib().tci_target_cache_stats(self.ctx.h, self.id, pi(out)))
        return dict(zip(("entries", "hits", "misses", "unstored"), (int(v) for v in out)))

    def __len__(self):
        return len(self.localdims)

    @property
    def ncached(self):
        return self.stats()["entries"]
