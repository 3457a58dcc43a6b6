"""Host mirror of the TensorTrain container (src/tensortrain.jl:17-93) and of the observables
evaluate / sum (src/abstracttensortrain.jl:124-199).  Cores are Fortran-ordered numpy arrays."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import core_ptrs, lib, pf, pi


class TensorTrain:
    def __init__(self, sitetensors):
        dt = np.complex128 if any(np.iscomplexobj(t) for t in sitetensors) else np.float64  # TensorTrain{ValueType,N}
        self.sitetensors = [np.asfortranarray(t, dtype=dt) for t in sitetensors]
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
    if any(np.iscomplexobj(c) for c in cores):  # TensorTrain{ComplexF64}: through the complex chains
        from .complexf64 import ZTTCache
        return ZTTCache(tt, ctx=ctx).evaluate_points(points)
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
    return complex(v[0]) if np.iscomplexobj(v) else float(v[0])


def fulltensor(tt):  # tensortrain.jl:279-292
    out = tt.sitetensors[0]
    for c in tt.sitetensors[1:]:
        out = np.tensordot(out, c, axes=([-1], [0]))
    return out.reshape(out.shape[1:-1])


# ---- arithmetic on tensor trains (tensortrain.jl:91-93, 186-245; abstracttensortrain.jl:164-317) ----
def tensortrain(obj):
    """tensortrain(tci) / tensortrain(tensors) (tensortrain.jl:62-65, 91-93): a TensorTrain holding copies."""
    cores = obj.sitetensors if hasattr(obj, "sitetensors") else list(obj)
    return TensorTrain([np.array(c, dtype=np.float64, order="F") for c in cores])


def sum_dims(tt, dims=None, ctx=None):
    """sum(tt; dims) (abstracttensortrain.jl:164-199).  dims are 1-based site numbers; without dims (or with all
    sites) the result is a scalar, otherwise a TensorTrain over the remaining sites."""
    n = len(tt.sitetensors)
    if dims is None:
        dims = tuple(range(1, n + 1))
    elif isinstance(dims, (int, np.integer)):
        dims = (int(dims),)
    tensors = []
    Tprod = np.eye(1)
    for k, T in enumerate(tt.sitetensors, start=1):
        T3 = T.reshape((T.shape[0], -1, T.shape[-1]), order="F")
        if k in dims:
            Tprod = _lib.gemm(Tprod, T3.sum(axis=1), ctx)
        else:
            P = _lib.gemm(Tprod, T3.reshape((T3.shape[0], -1), order="F"), ctx)
            tensors.append(P.reshape((P.shape[0], T3.shape[1], T3.shape[2]), order="F"))
            Tprod = np.eye(T3.shape[2])
    if not tensors:
        return float(Tprod[0, 0])
    last = tensors[-1]
    P = _lib.gemm(last.reshape((-1, last.shape[2]), order="F"), Tprod, ctx)
    tensors[-1] = P.reshape((last.shape[0], last.shape[1], P.shape[1]), order="F")
    return TensorTrain(tensors)


def _addtttensor(A, B, factorA=1.0, factorB=1.0, lefttensor=False, righttensor=False):
    """_addtttensor (abstracttensortrain.jl:201-218): block-diagonal stacking of two site tensors."""
    if A.ndim != B.ndim:
        raise ValueError("DimensionMismatch: Elementwise addition only works if both tensors have the same indices, "
                         f"but A and B have different numbers ({A.ndim} and {B.ndim}) of indices.")
    o1 = 0 if lefttensor else A.shape[0]
    o3 = 0 if righttensor else A.shape[-1]
    out = np.zeros((o1 + B.shape[0], *A.shape[1:-1], o3 + B.shape[-1]), order="F")
    out[: A.shape[0], ..., : A.shape[-1]] = factorA * A
    out[o1:, ..., o3:] = factorB * B
    return out


def add(lhs, rhs, factorlhs=1.0, factorrhs=1.0, tolerance=0.0, maxbonddim=2**63 - 1, ctx=None):
    """add(lhs, rhs; factorlhs, factorrhs, tolerance, maxbonddim) (abstracttensortrain.jl:238-262)."""
    from .contraction import compress  # (contraction imports this module)
    a, b = lhs.sitetensors, rhs.sitetensors
    if len(a) != len(b):
        raise ValueError(f"DimensionMismatch: Two tensor trains with different length ({len(a)} and {len(b)}) cannot "
                         "be added elementwise.")
    L = len(a)
    tt = TensorTrain([_addtttensor(a[k], b[k], factorlhs if k == L - 1 else 1.0, factorrhs if k == L - 1 else 1.0,
                                   lefttensor=(k == 0), righttensor=(k == L - 1)) for k in range(L)])
    compress(tt, "SVD", tolerance=tolerance, maxbonddim=maxbonddim, ctx=ctx)
    return tt


def subtract(lhs, rhs, tolerance=0.0, maxbonddim=2**63 - 1, ctx=None):  # :271-277
    return add(lhs, rhs, factorrhs=-1.0, tolerance=tolerance, maxbonddim=maxbonddim, ctx=ctx)


def multiply(tt, a):  # tensortrain.jl:186-214: the factor goes into the last site tensor
    out = tensortrain(tt)
    out.sitetensors[-1] = out.sitetensors[-1] * a
    return out


def divide(tt, a):  # :216-229
    out = tensortrain(tt)
    out.sitetensors[-1] = out.sitetensors[-1] / a
    return out


def reverse(tt):  # :231-235
    return TensorTrain([np.asfortranarray(np.moveaxis(T, (0, -1), (-1, 0))) for T in reversed(tt.sitetensors)])


def norm2(tt, ctx=None):
    """LA.norm2 (abstracttensortrain.jl:299-312): squared Frobenius norm.  The reference multiplies chi^2 x chi^2
    transfer matrices; here the chi x chi environment is carried through (two GEMMs per site on the device)."""
    env = np.eye(1)
    for T in tt.sitetensors:
        T3 = np.asfortranarray(T.reshape((T.shape[0], -1, T.shape[-1]), order="F"))
        chil, d, chir = T3.shape
        X = _lib.gemm(env, T3.reshape((chil, d * chir), order="F"), ctx)  # sum_l' env[l, l'] T[l', s, r]
        env = _lib.gemm(np.asfortranarray(T3.reshape((chil * d, chir), order="F").T),
                        X.reshape((chil * d, chir), order="F"), ctx)  # sum_{l, s} T[l, s, r] X[l, s, r']
    return float(env[0, 0])


def norm(tt, ctx=None):  # :314-316
    return float(np.sqrt(norm2(tt, ctx)))


TensorTrain.__call__ = lambda self, indexset: evaluate(self, indexset)
TensorTrain.__add__ = lambda self, other: add(self, other)
TensorTrain.__sub__ = lambda self, other: subtract(self, other)
TensorTrain.__mul__ = lambda self, a: multiply(self, a)
TensorTrain.__rmul__ = lambda self, a: multiply(self, a)
TensorTrain.__truediv__ = lambda self, a: divide(self, a)
