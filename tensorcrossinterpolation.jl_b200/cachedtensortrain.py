"""Host mirror of TTCache (src/cachedtensortrain.jl:9-225): a tensor train as the target
function.  The cores are uploaded once (tci_tt_create); environments and the batched Pi are
computed by K4 (csrc/tt.cu), so there is no host-side Dict memo."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import core_ptrs, lib, pi
from .batcheval import BatchEvaluator


class TTCache(BatchEvaluator):
    has_environments = True

    def __init__(self, tt, sitedims=None, ctx=None):
        ctx = ctx or _lib.default_context()
        cores = tt.sitetensors if hasattr(tt, "sitetensors") else list(tt)
        if sitedims is None:
            sitedims = [list(c.shape[1:-1]) for c in cores]
        if len(cores) != len(sitedims):  # cachedtensortrain.jl:16
            raise ValueError("The number of site tensors and site dimensions must be the same.")
        cores3 = []
        for n, (c, sd) in enumerate(zip(cores, sitedims)):
            if int(np.prod(sd)) != int(np.prod(c.shape[1:-1])):  # :18
                raise RuntimeError(f"Site dimensions do not match the site tensor dimensions at {n + 1}.")
            cores3.append(np.asfortranarray(c, dtype=np.float64).reshape((c.shape[0], -1, c.shape[-1]), order="F"))
        keep, arr = core_ptrs(cores3)
        d3 = np.ascontiguousarray(np.array([c.shape for c in keep], dtype=np.int64))
        tid = C.c_int64(0)
        ctx.check(lib().tci_tt_create(ctx.h, len(keep), pi(d3), arr, C.byref(tid)))
        super().__init__(ctx, tid.value, d3[:, 1].tolist())
        self.sitetensors = keep
        self.sitedims = [list(s) for s in sitedims]


def isbatchevaluable(f):  # cachedtensortrain.jl:228-229
    return isinstance(f, BatchEvaluator)
