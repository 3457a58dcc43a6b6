"""Host mirror of TTCache (src/cachedtensortrain.jl:9-225): a tensor train as the target
function.  The cores are uploaded once (tci_tt_create); environments and the batched Pi are
computed by K4 (csrc/tt.cu), so there is no host-side Dict memo."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import core_ptrs, lib, pi
from .batcheval import BatchEvaluator, apply_projector


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

    def batchevaluate(self, leftindexset, rightindexset, M, projector=None):
        """batchevaluate(tt::TTCache, leftindexset, rightindexset, Val(M), projector) (cachedtensortrain.jl:151-215)."""
        if len(leftindexset) * len(rightindexset) == 0:  # :156-158
            return np.zeros((0,) * (M + 2), order="F")
        nl = len(leftindexset[0])
        if len(self.localdims) - nl - len(rightindexset[0]) != M:
            raise RuntimeError(f"Invalid parameter M: {M}")  # :167-169
        if projector is None:
            return super().batchevaluate(leftindexset, rightindexset, M)
        projector = [[int(v) for v in pr] for pr in projector]
        if len(projector) != M:
            raise RuntimeError(f"Invalid length of projector: {projector}, correct length should be M={M}")  # :173-175
        for k, pr in enumerate(projector):
            sd = self.sitedims[nl + k]
            if len(pr) != len(sd):
                raise RuntimeError(f"Invalid projector at {nl + k + 1}: {pr}, the length must be {len(sd)}")  # :177
            if not all(0 <= v <= d for v, d in zip(pr, sd)):
                raise RuntimeError(f"Invalid projector: {pr}")  # :178
        return apply_projector(super().batchevaluate(leftindexset, rightindexset, M), self.sitedims[nl:nl + M], projector)


def isbatchevaluable(f):  # cachedtensortrain.jl:228-229
    return isinstance(f, BatchEvaluator)
