"""Host mirror of src/contraction.jl: the MPO x MPO product as a target (Contraction) and the
contract() dispatcher (:TCI, :naive, :zipup).  Environments, the batched Pi and the site
contractions run on the GPU (K5/K6, csrc/mpo.cu)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DeviceMatrix, core_ptrs, lib, pf, pi
from .batcheval import BatchEvaluator
from .matrixlu import MatrixLUCI, left, npivots, right, rrlu
from .tensortrain import TensorTrain
from .tensorci2 import crossinterpolate2

I64MAX = 2**63 - 1


class Contraction(BatchEvaluator):
    """struct Contraction (contraction.jl:5-62).  The elementwise function `f` (applied to the product at
    contraction.jl:203-205, 330-332) is registered by id, as the targets are -- a device kernel cannot call a host
    closure: ("affine", a, b) for x -> a*x + b (the reference's test function x -> 2x is ("affine", 2.0, 0.0)),
    ("abs",), ("square",)."""
    has_environments = True
    ELEMENTWISE = {"affine": 1, "abs": 2, "square": 3}

    def __init__(self, a, b, f=None, ctx=None):
        ctx = ctx or _lib.default_context()
        if f is not None:
            if callable(f) or not f or f[0] not in self.ELEMENTWISE:
                raise NotImplementedError("Contraction: the elementwise function f must be one of the device functions "
                                          '("affine", a, b), ("abs",), ("square",); host closures cannot run on the GPU')
            f = (f[0], float(f[1]) if len(f) > 1 else 1.0, float(f[2]) if len(f) > 2 else 0.0)
        A = a.sitetensors if hasattr(a, "sitetensors") else list(a)
        B = b.sitetensors if hasattr(b, "sitetensors") else list(b)
        if len(A) != len(B):
            raise ValueError("Tensor trains must have the same length.")  # :41-43
        for n in range(len(A)):
            if A[n].shape[2] != B[n].shape[1]:
                raise RuntimeError(f"Tensor trains must share the identical index at n={n + 1}!")  # :44-48
        ka, pa = core_ptrs(A)
        kb, pb = core_ptrs(B)
        da = np.ascontiguousarray(np.array([c.shape for c in ka], dtype=np.int64))
        db = np.ascontiguousarray(np.array([c.shape for c in kb], dtype=np.int64))
        tid = C.c_int64(0)
        ctx.check(lib().tci_mpo_pair_create(ctx.h, len(ka), pi(da), pa, pi(db), pb, C.byref(tid)))
        self.sitedims = [[int(x.shape[1]), int(y.shape[2])] for x, y in zip(ka, kb)]
        super().__init__(ctx, tid.value, [s[0] * s[1] for s in self.sitedims])
        self.mpo = (ka, kb)
        self.f = f
        if f is not None:
            ctx.check(lib().tci_target_set_elementwise(ctx.h, self.id, self.ELEMENTWISE[f[0]], f[1], f[2]))
            self.has_environments = False  # Pi is f(left^T right): not a product of environments any more

    def batchevaluate(self, leftindexset, rightindexset, M, projector=None):
        """batchevaluate(obj::Contraction, leftindexset, rightindexset, Val(M), projector) (contraction.jl:236-335)."""
        if projector is None:
            return super().batchevaluate(leftindexset, rightindexset, M)
        nl = len(leftindexset[0])
        projector = [[int(v) for v in pr] for pr in projector]
        if len(projector) != M:
            raise RuntimeError(f"Length mismatch: length of projector (={len(projector)}) must be {M}")  # :250
        for k, pr in enumerate(projector):
            if len(pr) != 2:
                raise RuntimeError(f"Invalid projector at {nl + k + 1}: {pr}, the length must be 2")  # :252
            if not all(0 <= v <= d for v, d in zip(pr, self.sitedims[nl + k])):
                raise RuntimeError(f"Invalid projector: {pr}")  # :253
        from .batcheval import apply_projector
        return apply_projector(super().batchevaluate(leftindexset, rightindexset, M), self.sitedims[nl:nl + M], projector)


def _gemm(A, B, ctx):
    if np.iscomplexobj(A) or np.iscomplexobj(B):
        from .complexf64 import zgemm
        return zgemm(A, B, ctx=ctx)
    return _lib.gemm(A, B, ctx)


def _contractsitetensors(a, b, ctx=None):  # contraction.jl:338-349
    ctx = ctx or _lib.default_context()
    if np.iscomplexobj(a) or np.iscomplexobj(b):
        from .complexf64 import zcontractsitetensors
        return zcontractsitetensors(a, b, ctx)
    a = np.asfortranarray(a, dtype=np.float64)
    b = np.asfortranarray(b, dtype=np.float64)
    Da, s1, s2, Dan = a.shape
    Db, s2b, s3, Dbn = b.shape
    if s2 != s2b:
        raise ValueError("shared site dimension mismatch")
    out = np.zeros(Da * Db * s1 * s3 * Dan * Dbn, dtype=np.float64)
    ctx.check(lib().tci_contract_naive_site(ctx.h, pf(a), Da, s1, s2, Dan, pf(b), Db, s3, Dbn, pf(out)))
    return out.reshape((Da * Db, s1, s3, Dan * Dbn), order="F")


def _factorize(A, method, tolerance, maxbonddim, leftorthogonal=False, normalizeerror=True, ctx=None):
    """_factorize (tensortrain.jl:95-137).  A: host matrix or DeviceMatrix.  :LU/:CI on the GPU (K2/K3);
    :SVD on the host (LAPACK), as in the reference -- recompression is outside the hot path."""
    reltol, abstol = 1e-14, 0.0
    if normalizeerror:
        reltol = tolerance
    else:
        abstol = tolerance
    mr = None if maxbonddim >= I64MAX else maxbonddim
    if method in ("LU", "CI"):
        fac = MatrixLUCI(A, abstol=abstol, reltol=reltol, maxrank=mr, leftorthogonal=leftorthogonal, ctx=ctx) \
            if not isinstance(A, DeviceMatrix) else MatrixLUCI(A, abstol=abstol, reltol=reltol, maxrank=mr,
                                                               leftorthogonal=leftorthogonal)
        if method == "CI":
            return fac.left(), fac.right(), fac.npivot
        return left(fac.lu), right(fac.lu), npivots(fac.lu)
    if method == "SVD":
        if isinstance(A, DeviceMatrix):
            A = A.to_host()
        U, S, Vt = np.linalg.svd(A, full_matrices=False)  # (ComplexF64: Vt is V^H, as Julia's svd(A).Vt)
        err = np.array([np.sum(S[n + 1:] ** 2) for n in range(len(S))])
        nerr = err / np.sum(S ** 2)

        def first(cond, default):
            w = np.nonzero(cond)[0]
            return int(w[0]) + 1 if w.size else default

        trunci = min(first(err < abstol ** 2, len(err)), first(nerr < reltol ** 2, len(nerr)), maxbonddim)
        if leftorthogonal:
            return U[:, :trunci], np.diag(S[:trunci]) @ Vt[:trunci, :], trunci
        return U[:, :trunci] * S[:trunci], Vt[:trunci, :], trunci
    raise RuntimeError("Not implemented yet.")


def compress(tt, method="LU", tolerance=1e-12, maxbonddim=I64MAX, normalizeerror=True, ctx=None):
    """compress!(tt, method; tolerance, maxbonddim, normalizeerror) (tensortrain.jl:149-183): a left-to-right
    sweep without truncation followed by a truncating right-to-left sweep.  :LU / :CI factorise on the GPU
    (tci_rrlu + tci_luci_*), :SVD on the host."""
    cores = tt.sitetensors
    n = len(cores)
    for ell in range(n - 1):  # :157-167
        shl = cores[ell].shape
        lft, rgt, newd = _factorize(cores[ell].reshape((-1, shl[-1]), order="F"), method, tolerance=0.0,
                                    maxbonddim=I64MAX, leftorthogonal=True, ctx=ctx)
        cores[ell] = np.asfortranarray(lft).reshape((*shl[:-1], newd), order="F")
        shr = cores[ell + 1].shape
        nxt = _gemm(rgt, cores[ell + 1].reshape((shr[0], -1), order="F"), ctx)
        cores[ell + 1] = np.asfortranarray(nxt).reshape((newd, *shr[1:]), order="F")
    for ell in range(n - 1, 0, -1):  # :170-180
        shr = cores[ell].shape
        lft, rgt, newd = _factorize(cores[ell].reshape((shr[0], -1), order="F"), method, tolerance=tolerance,
                                    maxbonddim=maxbonddim, normalizeerror=normalizeerror, leftorthogonal=False,
                                    ctx=ctx)
        cores[ell] = np.asfortranarray(rgt).reshape((newd, *shr[1:]), order="F")
        shl = cores[ell - 1].shape
        nxt = _gemm(cores[ell - 1].reshape((-1, shl[-1]), order="F"), lft, ctx)
        cores[ell - 1] = np.asfortranarray(nxt).reshape((*shl[:-1], newd), order="F")
    return tt


def contract_naive(a, b, tolerance=0.0, maxbonddim=I64MAX, ctx=None):  # contraction.jl:351-372
    tt = TensorTrain([_contractsitetensors(x, y, ctx) for x, y in zip(a.sitetensors, b.sitetensors)])
    if tolerance > 0 or maxbonddim < I64MAX:
        compress(tt, "SVD", tolerance=tolerance, maxbonddim=maxbonddim, ctx=ctx)
    return tt


def contract_zipup(A, B, tolerance=1e-12, method="SVD", maxbonddim=I64MAX, ctx=None):
    """contract_zipup (contraction.jl:442-486)."""
    ctx = ctx or _lib.default_context()
    if len(A) != len(B):
        raise ValueError("Cannot contract tensor trains with different length.")
    R = np.ones((1, 1, 1), order="F")
    out = []
    N = len(A)
    if any(np.iscomplexobj(c) for c in list(A) + list(B)):  # TensorTrain{ComplexF64,4}
        from .complexf64 import zzipup_site
        for n in range(N):
            chi = R.shape[0]
            _, s1, _, Dan = A[n].shape
            _, _, s3, Dbn = B[n].shape
            if n == N - 1:
                out.append(zzipup_site(ctx, R, A[n], B[n], False).reshape((chi, s1, s3, 1), order="F"))
                break
            lft, rgt, newdim = _factorize(zzipup_site(ctx, R, A[n], B[n], True), method, tolerance=tolerance,
                                          maxbonddim=maxbonddim)
            out.append(np.asfortranarray(lft).reshape((chi, s1, s3, newdim), order="F"))
            R = np.asfortranarray(rgt).reshape((newdim, Dan, Dbn), order="F")
        return TensorTrain(out)
    for n in range(N):
        a = np.asfortranarray(A[n], dtype=np.float64)
        b = np.asfortranarray(B[n], dtype=np.float64)
        chi, Da, Db = R.shape
        _, s1, s2, Dan = a.shape
        _, _, s3, Dbn = b.shape
        Rf = np.asfortranarray(R)
        if n == N - 1:
            Cm = np.zeros(chi * s1 * s3 * Dan * Dbn, dtype=np.float64)
            ctx.check(lib().tci_contract_zipup_site(ctx.h, pf(Rf), chi, Da, Db, pf(a), s1, s2, Dan, pf(b), s3, Dbn,
                                                    pf(Cm), None))
            out.append(Cm.reshape((chi, s1, s3, 1), order="F"))
            break
        h = C.c_void_p()
        ctx.check(lib().tci_contract_zipup_site(ctx.h, pf(Rf), chi, Da, Db, pf(a), s1, s2, Dan, pf(b), s3, Dbn, None,
                                                C.byref(h)))
        Cdev = DeviceMatrix(ctx, h)
        lft, rgt, newdim = _factorize(Cdev, method, tolerance=tolerance, maxbonddim=maxbonddim)
        out.append(np.asfortranarray(lft).reshape((chi, s1, s3, newdim), order="F"))
        R = np.asfortranarray(rgt).reshape((newdim, Dan, Dbn), order="F")
    return TensorTrain(out)


def contract_TCI(A, B, initialpivots=None, f=None, ctx=None, **kwargs):
    """contract_TCI (contraction.jl:399-436); initial pivots must be given explicitly or default to
    the all-ones index (the reference draws them with Julia's rng, :386-397)."""
    if len(A) != len(B):
        raise ValueError("Cannot contract tensor trains with different length.")
    for i in range(len(A)):
        if A[i].shape[2] != B[i].shape[1]:
            raise ValueError("Cannot contract tensor trains with non-matching site dimensions.")
    if any(np.iscomplexobj(c) for c in list(A) + list(B)):  # TensorTrain{ComplexF64,4}: the complex kernels
        from .complexf64 import ZContraction
        mp = ZContraction(A, B, f=f, ctx=ctx)
    else:
        mp = Contraction(A, B, f=f, ctx=ctx)
    localdims = mp.localdims
    if initialpivots is None:
        initialpivots = [[1] * len(localdims)]
    tci, ranks, errors = crossinterpolate2(mp, localdims, initialpivots, **kwargs)
    cores = [t.reshape((t.shape[0], sd[0], sd[1], t.shape[-1]), order="F") for t, sd in
             zip(tci.sitetensors, mp.sitedims)]
    return TensorTrain(cores)


def contract(A, B, algorithm="TCI", tolerance=1e-12, maxbonddim=I64MAX, f=None, **kwargs):  # :515-560
    # MPS x MPO and MPO x MPS (contraction.jl:544-560): a three-leg train gets a site leg of size 1 on the side that is
    # not contracted -- (Dl, 1, s, Dr) on the left, (Dl, s, 1, Dr) on the right -- and the result is folded back
    ca = A.sitetensors if hasattr(A, "sitetensors") else list(A)
    cb = B.sitetensors if hasattr(B, "sitetensors") else list(B)
    a3, b3 = all(c.ndim == 3 for c in ca), all(c.ndim == 3 for c in cb)
    if a3 != b3:
        if a3:
            A = TensorTrain([np.asfortranarray(c).reshape((c.shape[0], 1, c.shape[1], c.shape[2]), order="F") for c in ca])
        else:
            B = TensorTrain([np.asfortranarray(c).reshape((c.shape[0], c.shape[1], 1, c.shape[2]), order="F") for c in cb])
        tt = contract(A, B, algorithm=algorithm, tolerance=tolerance, maxbonddim=maxbonddim, f=f, **kwargs)
        return TensorTrain([np.asfortranarray(c).reshape((c.shape[0], -1, c.shape[-1]), order="F")
                            for c in tt.sitetensors])
    if algorithm == "TCI":
        return contract_TCI(A, B, tolerance=tolerance, maxbonddim=maxbonddim, f=f, **kwargs)
    if algorithm == "naive":
        if f is not None:
            raise RuntimeError("Naive contraction implementation cannot contract matrix product with a function. "
                               "Use algorithm=:TCI instead.")
        return contract_naive(A, B, tolerance=tolerance, maxbonddim=maxbonddim)
    if algorithm == "zipup":
        if f is not None:
            raise RuntimeError("Zipup contraction implementation cannot contract matrix product with a function. "
                               "Use algorithm=:TCI instead.")
        return contract_zipup(A, B, tolerance=tolerance, maxbonddim=maxbonddim, **kwargs)
    raise ValueError(f"Unknown algorithm {algorithm}.")
