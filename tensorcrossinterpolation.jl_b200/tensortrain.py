"""Host mirror of the TensorTrain container (src/tensortrain.jl:17-93) and of the observables
evaluate / sum (src/abstracttensortrain.jl:124-199).  Cores are Fortran-ordered numpy arrays."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import core_ptrs, lib, pf, pi


class TensorTrain:
    def __init__(self, sitetensors):
        self.sitetensors = [np.asfortranarray(t, dtype=np.float64) for t in sitetensors]
        for i in range(len(self.sitetensors) - 1):  # tensortrain.jl:21-27
            if self.sitetensors[i].shape[-1] != self.sitetensors[i + 1].shape[0]:
                raise ValueError(f"The tensors at {i + 1} and {i + 2} must have consistent dimensions for a "
                                 "tensor train.")

    def __len__(self):
        return len(self.sitetensors)

    def __getitem__(self, i):
        return self.sitetensors[i]

    def __iter__(self):
        return iter(self.sitetensors)


def sitetensors(tt):
    return tt.sitetensors


def linkdims(tt):  # abstracttensortrain.jl:32-34
    if hasattr(tt, "linkdims"):
        return tt.linkdims()
    return [t.shape[0] for t in tt.sitetensors[1:]]


def rank(tt):  # :67-69
    return max(linkdims(tt))


def sitedims(tt):  # :50-52
    return [list(t.shape[1:-1]) for t in tt.sitetensors]


def _dims3(cores):
    return np.ascontiguousarray(np.array([[c.shape[0], int(np.prod(c.shape[1:-1])), c.shape[-1]] for c in cores],
                                         dtype=np.int64))


def evaluate(tt, indexset, ctx=None):
    """evaluate(tt, indexset): ordered left-to-right product (abstracttensortrain.jl:124-132),
    computed on the device in that order (tci_tt_evaluate)."""
    cores = tt.sitetensors
    if len(indexset) != len(cores):
        raise ValueError(f"To evaluate a tt of length {len(cores)}, you have to provide {len(cores)} indices, "
                         f"but there were {len(indexset)}.")
    return float(evaluate_points(tt, [indexset], ctx)[0])


def evaluate_points(tt, points, ctx=None):
    ctx = ctx or getattr(tt, "ctx", None) or _lib.default_context()
    cores = tt.sitetensors
    pts = np.ascontiguousarray(np.asarray(points, dtype=np.int64).reshape(-1, len(cores)))
    keep, arr = core_ptrs(cores)
    d3 = _dims3(keep)
    out = np.zeros(pts.shape[0], dtype=np.float64)
    if pts.shape[0]:
        ctx.check(lib().tci_tt_evaluate(ctx.h, len(keep), pi(d3), arr, pi(pts), pts.shape[0], pf(out)))
    return out


def tt_sum(tt):
    """sum(tt)  abstracttensortrain.jl:164-199 (host: O(n d chi^2), an observable, not on the hot path)."""
    v = np.ones((1,), dtype=np.float64)
    for T in tt.sitetensors:
        T3 = T.reshape((T.shape[0], -1, T.shape[-1]), order="F")
        v = v @ T3.sum(axis=1)
    return float(v[0])


def fulltensor(tt):  # tensortrain.jl:279-292
    out = tt.sitetensors[0]
    for c in tt.sitetensors[1:]:
        out = np.tensordot(out, c, axes=([-1], [0]))
    return out.reshape(out.shape[1:-1])
