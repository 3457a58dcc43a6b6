"""ComplexF64 value type of the TensorCI2 path: ctypes front end of the tci_z* entry points.

The reference is generic in the value type; its contraction and conversion tests run on ComplexF64
(test_contraction.jl:39-46, test_matrixlu.jl:39-52).  A Matrix{ComplexF64} crosses the ABI as interleaved (re, im)
pairs, which is numpy's complex128 memory as well, so arrays are passed without conversion.  Everything computes on
the GPU (csrc/zpath.cu, csrc/zgemm.cu, the chains of csrc/mpo.cu instantiated for complex cores); there is no CPU
fallback.  The classes plug into the unchanged driver of tensorci2.py: a complex evaluator returns ZDeviceMatrix
objects, and rrlu / MatrixLUCI of matrixlu.py dispatch on them."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DeviceMatrix, lib, pi
from .batcheval import BatchEvaluator
from .matrixlu import I64MAX, rrLU
from .util import as_indexset

_SYMS = {
    "tci_zrrlu": (C.c_int, [_lib.VP, _lib.VP, _lib.VP, _lib.i64, _lib.i64, _lib.i64, _lib.f64, _lib.f64, C.c_int,
                            _lib.P_i64, _lib.P_i64, _lib.P_i64, _lib.P_f64, _lib.P_f64, C.POINTER(_lib.VP)]),
    "tci_zlu_fetch": (C.c_int, [_lib.VP, _lib.VP, _lib.VP]),
    "tci_zluci_left": (C.c_int, [_lib.VP, _lib.VP, C.POINTER(_lib.VP)]),
    "tci_zluci_right": (C.c_int, [_lib.VP, _lib.VP, C.POINTER(_lib.VP)]),
    "tci_zlu_rdiv": (C.c_int, [_lib.VP, _lib.VP, _lib.VP, C.POINTER(_lib.VP)]),
    "tci_zmpo_pair_create": (C.c_int, [_lib.VP, _lib.i64, _lib.P_i64, C.POINTER(_lib.VP), _lib.P_i64,
                                       C.POINTER(_lib.VP), _lib.P_i64]),
    "tci_ztt_create": (C.c_int, [_lib.VP, _lib.i64, _lib.P_i64, C.POINTER(_lib.VP), _lib.P_i64]),
    "tci_zpi_eval": (C.c_int, [_lib.VP, _lib.i64, _lib.P_i64, _lib.i64, _lib.i64, _lib.P_i64, _lib.i64, _lib.i64,
                               _lib.i64, _lib.VP, C.POINTER(_lib.VP), _lib.P_f64]),
    "tci_ztarget_eval": (C.c_int, [_lib.VP, _lib.i64, _lib.P_i64, _lib.i64, _lib.VP]),
    "tci_zbond_update": (C.c_int, [_lib.VP, _lib.i64, _lib.P_i64, _lib.i64, _lib.i64, _lib.P_i64, _lib.i64, _lib.i64,
                                   _lib.i64, _lib.f64, _lib.f64, C.c_int, _lib.P_i64, _lib.P_i64, _lib.P_i64,
                                   _lib.P_f64, _lib.P_f64, _lib.P_f64, C.POINTER(_lib.VP)]),
    "tci_zgemm_host": (C.c_int, [_lib.VP, C.c_int, C.c_int, _lib.i64, _lib.i64, _lib.i64, _lib.VP, _lib.VP, _lib.VP]),
    "tci_zcontract_zipup_site": (C.c_int, [_lib.VP, _lib.VP, _lib.i64, _lib.i64, _lib.i64, _lib.VP, _lib.i64, _lib.i64,
                                           _lib.i64, _lib.VP, _lib.i64, _lib.i64, _lib.VP, C.POINTER(_lib.VP)]),
    "tci_zcontract_naive_site": (C.c_int, [_lib.VP, _lib.VP, _lib.i64, _lib.i64, _lib.i64, _lib.i64, _lib.VP, _lib.i64,
                                           _lib.i64, _lib.i64, _lib.VP]),
}
_lib.SYMBOLS.update(_SYMS)
if _lib._lib is not None:  # the library was bound before this module was imported
    for _n, (_r, _a) in _SYMS.items():
        getattr(_lib._lib, _n).restype = _r
        getattr(_lib._lib, _n).argtypes = _a


def _pz(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _zarr(a):
    return np.asfortranarray(a, dtype=np.complex128)


def _zcore_ptrs(cores):
    keep = [_zarr(c) for c in cores]
    return keep, (_lib.VP * len(keep))(*[c.ctypes.data for c in keep])


class ZDeviceMatrix(DeviceMatrix):
    """A complex m x n matrix in HBM: the tci_dmat holds 2m x n doubles (interleaved pairs)."""
    is_complex = True

    @classmethod
    def from_host(cls, ctx, a):
        a = _zarr(a)
        h = _lib.VP()
        ctx.check(lib().tci_dmat_create(ctx.h, 2 * a.shape[0], a.shape[1], C.cast(_pz(a), _lib.P_f64), C.byref(h)))
        return cls(ctx, h)

    @property
    def shape(self):
        m2, n = DeviceMatrix.shape.fget(self)
        return m2 // 2, n

    def to_host(self):
        m, n = self.shape
        out = np.zeros((m, n), dtype=np.complex128, order="F")
        if m * n:
            self.ctx.check(lib().tci_dmat_fetch(self.h, C.cast(_pz(out), _lib.P_f64)))
        return out

    def refold(self, m2, n2):
        h = _lib.VP()
        self.ctx.check(lib().tci_dmat_refold(self.h, 2 * int(m2), int(n2), C.byref(h)))
        return ZDeviceMatrix(self.ctx, h)


class ZrrLU(rrLU):
    """rrLU{ComplexF64} (matrixlu.jl:71-96): complex L / U, real pivot errors."""
    is_complex = True

    def _fetch(self):
        m, n = self._shape
        r = self.npivot
        if self._h is None:
            raise RuntimeError("this rrLU was computed without factors (bond_update(want_factors=False))")
        L = np.zeros((m, r), dtype=np.complex128, order="F")
        U = np.zeros((r, n), dtype=np.complex128, order="F")
        if r:
            self.ctx.check(lib().tci_zlu_fetch(self._h, _pz(L), _pz(U)))
        self._L, self._U = L, U

    def rdiv(self, B, device=False):
        if not isinstance(B, ZDeviceMatrix):
            B = ZDeviceMatrix.from_host(self.ctx, B)
        rows, k = B.shape
        if device:
            h = _lib.VP()
            self.ctx.check(lib().tci_zlu_rdiv(self._h, B.h, None, C.byref(h)))
            return ZDeviceMatrix(self.ctx, h)
        out = np.zeros((rows, k), dtype=np.complex128, order="F")
        self.ctx.check(lib().tci_zlu_rdiv(self._h, B.h, _pz(out), None))
        return out

    def luci_left(self, device=False):
        m, n = self._shape
        if device:
            h = _lib.VP()
            self.ctx.check(lib().tci_zluci_left(self._h, None, C.byref(h)))
            return ZDeviceMatrix(self.ctx, h)
        out = np.zeros((m, self.npivot), dtype=np.complex128, order="F")
        self.ctx.check(lib().tci_zluci_left(self._h, _pz(out), None))
        return out

    def luci_right(self, device=False):
        m, n = self._shape
        if device:
            h = _lib.VP()
            self.ctx.check(lib().tci_zluci_right(self._h, None, C.byref(h)))
            return ZDeviceMatrix(self.ctx, h)
        out = np.zeros((self.npivot, n), dtype=np.complex128, order="F")
        self.ctx.check(lib().tci_zluci_right(self._h, _pz(out), None))
        return out


def zrrlu(A, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True, ctx=None, **_):
    """rrlu(A::Matrix{ComplexF64}; ...) (matrixlu.jl:217-225) through tci_zrrlu."""
    if isinstance(A, ZDeviceMatrix):
        ctx = A.ctx
        m, n = A.shape
        host, dev = None, A.h
    else:
        ctx = ctx or _lib.default_context()
        A = _zarr(A)
        if A.ndim != 2:
            raise ValueError("rrlu expects a matrix")
        m, n = A.shape
        host, dev = A, None
    if maxrank is not None and maxrank <= 0:
        raise ValueError("maxrank must be positive")
    mr = 0 if maxrank is None else int(min(maxrank, I64MAX))
    rowperm = np.zeros(m, dtype=np.int64)
    colperm = np.zeros(n, dtype=np.int64)
    npiv, err = C.c_int64(0), C.c_double(0.0)
    pe = np.zeros(min(m, n) + 1, dtype=np.float64)
    h = _lib.VP()
    rc = lib().tci_zrrlu(ctx.h, _pz(host), dev, m, n, mr, float(reltol), float(abstol), int(bool(leftorthogonal)),
                         pi(rowperm), pi(colperm), C.byref(npiv), C.byref(err), _lib.pf(pe), C.byref(h))
    if dev is not None and rc == 0 and h:
        A.release()
    if rc in (_lib.TCI_ERR_NAN_L, _lib.TCI_ERR_NAN_U):
        raise RuntimeError(lib().tci_last_error(ctx.h).decode())
    ctx.check(rc)
    r = npiv.value
    return ZrrLU(ctx, h, rowperm, colperm, r, err.value, pe[: r + 1].copy(), bool(leftorthogonal), (m, n))


def zgemm(A, B, transA=False, transB=False, ctx=None):
    """op(A) * op(B) for host complex matrices through the library's complex DMMA GEMM (tci_zgemm_host)."""
    ctx = ctx or _lib.default_context()
    A, B = _zarr(A), _zarr(B)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    K2, N = (B.shape[1], B.shape[0]) if transB else B.shape
    if K != K2:
        raise ValueError(f"DimensionMismatch: A has dimensions {A.shape}, B has dimensions {B.shape}")
    out = np.zeros((M, N), dtype=np.complex128, order="F")
    if M and N:
        ctx.check(lib().tci_zgemm_host(ctx.h, int(transA), int(transB), M, N, K, _pz(A), _pz(B), _pz(out)))
    return out


class ZBatchEvaluator(BatchEvaluator):
    """BatchEvaluator{ComplexF64} bound to a complex device target."""
    is_complex = True

    def __call__(self, *args):
        if len(args) == 3:
            return self.batchevaluate(*args)
        (indexset,) = args
        return complex(self.evaluate_points([indexset])[0])

    def evaluate_points(self, points):
        pts = as_indexset(points, len(self.localdims))
        out = np.zeros(pts.shape[0], dtype=np.complex128)
        if pts.shape[0]:
            self.ctx.check(lib().tci_ztarget_eval(self.ctx.h, self.id, pi(pts), pts.shape[0], _pz(out)))
        self.nevals += pts.shape[0]
        return out

    def _pi(self, Iset, Jset, M, want_host, want_dev):
        n = len(self.localdims)
        nI, nJ = len(Iset), len(Jset)
        if nI * nJ == 0:
            return np.zeros((0,) * (M + 2), dtype=np.complex128, order="F"), None, 0.0
        I, J = as_indexset(Iset), as_indexset(Jset)
        nl, nr = I.shape[1], J.shape[1]
        if nl + M + nr != n:
            raise RuntimeError("Invalid number of central indices")
        cd = self.localdims[nl:nl + M]
        Csz = int(np.prod(cd, dtype=np.int64)) if M else 1
        host = np.zeros(nI * Csz * nJ, dtype=np.complex128) if want_host else None
        dev = _lib.VP()
        mx = C.c_double(0.0)
        self.ctx.check(lib().tci_zpi_eval(self.ctx.h, self.id, pi(I), nl, nI, pi(J), nr, nJ, M, _pz(host),
                                          C.byref(dev) if want_dev else None, C.byref(mx)))
        self.nevals += nI * Csz * nJ
        if want_host:
            host = host.reshape((nI, *cd, nJ), order="F")
        return host, (ZDeviceMatrix(self.ctx, dev) if want_dev else None), mx.value

    def bond_update(self, Icombined, Jcombined, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True,
                    want_factors=False, exact=True):
        I, J = as_indexset(Icombined), as_indexset(Jcombined)
        m, n = len(I), len(J)
        if m == 0 or n == 0:
            raise ValueError("rows must not be empty")
        if I.shape[1] + J.shape[1] != len(self.localdims):
            raise RuntimeError("Invalid number of central indices")
        rowperm = np.zeros(m, dtype=np.int64)
        colperm = np.zeros(n, dtype=np.int64)
        npiv, err, mx = C.c_int64(0), C.c_double(0.0), C.c_double(0.0)
        pe = np.zeros(min(m, n) + 1, dtype=np.float64)
        h = _lib.VP()
        mr = 0 if maxrank is None else int(min(maxrank, 2**62))
        self.ctx.check(lib().tci_zbond_update(self.ctx.h, self.id, pi(I), I.shape[1], m, pi(J), J.shape[1], n, mr,
                                              float(reltol), float(abstol), int(bool(leftorthogonal)), pi(rowperm),
                                              pi(colperm), C.byref(npiv), C.byref(err), _lib.pf(pe), C.byref(mx),
                                              C.byref(h) if want_factors else None))
        self.nevals += m * n
        r = npiv.value
        return ZrrLU(self.ctx, h if want_factors else None, rowperm, colperm, r, err.value, pe[: r + 1].copy(),
                     bool(leftorthogonal), (m, n)), mx.value

    def fill_sitetensors(self, Isets, Jsets, want_handle=True):
        """fillsitetensors! (globalsearch.jl:97-103): T_b = Pi1_b P_b^-1 per site (tensorci2.jl:367-394), Pi1 / P evaluated
        into HBM, P factorised to full rank by the complex rrLU, the solve on the device (tci_zlu_rdiv)."""
        n = len(self.localdims)
        Ts, mx = [], 0.0
        for b in range(n):
            Ib, Jb = as_indexset(Isets[b], b), as_indexset(Jsets[b], n - 1 - b)
            nI, d, nJ = len(Ib), self.localdims[b], len(Jb)
            if b == n - 1:
                Pi1, _, m1 = self._pi(Ib, Jb, 1, True, False)
                Ts.append(np.asfortranarray(Pi1.reshape((nI, d, nJ), order="F")))
            else:
                Pi1, m1 = self.batchevaluate_device(Ib, Jb, 1)
                Inext = as_indexset(Isets[b + 1], b + 1)
                if len(Inext) != nJ:
                    raise RuntimeError(f"Pivot matrix at bond {b + 1} is not square!")
                P, _ = self.batchevaluate_device(Inext, Jb, 0)
                lu = zrrlu(P, reltol=0.0, abstol=0.0)
                if lu.npivot != nJ:
                    raise RuntimeError(f"Pivot matrix at bond {b + 1} is singular!")
                Ts.append(np.asfortranarray(lu.rdiv(Pi1).reshape((nI, d, nJ), order="F")))
            mx = max(mx, m1) if not (np.isnan(mx) or np.isnan(m1)) else float("nan")
        handle = ZTTCache(Ts, ctx=self.ctx) if want_handle else None
        return Ts, mx, handle


class ZContraction(ZBatchEvaluator):
    """Contraction(a, b) of two TensorTrain{ComplexF64,4} (contraction.jl:5-62).  f: None or ("affine", a, b) with real
    a, b (the reference's tests use x -> 2x)."""

    def __init__(self, a, b, f=None, ctx=None):
        ctx = ctx or _lib.default_context()
        A = a.sitetensors if hasattr(a, "sitetensors") else list(a)
        B = b.sitetensors if hasattr(b, "sitetensors") else list(b)
        if len(A) != len(B):
            raise ValueError("Tensor trains must have the same length.")
        for n in range(len(A)):
            if A[n].shape[2] != B[n].shape[1]:
                raise RuntimeError(f"Tensor trains must share the identical index at n={n + 1}!")
        if f is not None and (callable(f) or not f or f[0] != "affine"):
            raise NotImplementedError('ComplexF64 Contraction: f must be ("affine", a, b); host closures cannot run on the GPU')
        ka, pa = _zcore_ptrs(A)
        kb, pb = _zcore_ptrs(B)
        da = np.ascontiguousarray(np.array([c.shape for c in ka], dtype=np.int64))
        db = np.ascontiguousarray(np.array([c.shape for c in kb], dtype=np.int64))
        tid = C.c_int64(0)
        ctx.check(lib().tci_zmpo_pair_create(ctx.h, len(ka), pi(da), pa, pi(db), pb, C.byref(tid)))
        self.sitedims = [[int(x.shape[1]), int(y.shape[2])] for x, y in zip(ka, kb)]
        super().__init__(ctx, tid.value, [s[0] * s[1] for s in self.sitedims])
        self.mpo = (ka, kb)
        self.f = f
        if f is not None:
            ctx.check(lib().tci_target_set_elementwise(ctx.h, self.id, 1, float(f[1]) if len(f) > 1 else 1.0,
                                                       float(f[2]) if len(f) > 2 else 0.0))


class ZTTCache(ZBatchEvaluator):
    """TTCache(tt) of a TensorTrain{ComplexF64,3} (cachedtensortrain.jl:9-30); cores (Dl, d, Dr)."""

    def __init__(self, tt, ctx=None):
        ctx = ctx or _lib.default_context()
        cores = tt.sitetensors if hasattr(tt, "sitetensors") else list(tt)
        keep, ptrs = _zcore_ptrs([np.asarray(c).reshape((c.shape[0], -1, c.shape[-1]), order="F") for c in cores])
        d3 = np.ascontiguousarray(np.array([c.shape for c in keep], dtype=np.int64))
        tid = C.c_int64(0)
        ctx.check(lib().tci_ztt_create(ctx.h, len(keep), pi(d3), ptrs, C.byref(tid)))
        super().__init__(ctx, tid.value, [int(c.shape[1]) for c in keep])
        self.cores = keep


def zfind_global_pivots(finder, input, f, abstol, rng=None, verbosity=0):
    """DefaultGlobalPivotFinder call (globalpivotfinder.jl:143-195) for a ComplexF64 target: the star probes of every
    start are evaluated in two device batches (f and the current tensor train), the error is abs(f - tt) = hypot, and
    the first-maximum / threshold / truncation logic is the library's host-side selection (tci_globalsearch_select)."""
    n = len(input.localdims)
    starts = np.ascontiguousarray(finder.draw(input, rng))
    ns = starts.shape[0]
    tt = getattr(input.current_tt, "device_handle", None)
    if tt is None:
        tt = ZTTCache(input.current_tt, ctx=f.ctx)
    ld = np.ascontiguousarray(input.localdims, dtype=np.int64)
    per = int(ld.sum())
    pts = np.repeat(starts, per, axis=0)  # probe q of start s: site p(q), value v(q) replaced
    site = np.concatenate([np.full(int(d), p) for p, d in enumerate(ld)])
    val = np.concatenate([np.arange(1, int(d) + 1) for d in ld])
    pts[np.arange(ns * per), np.tile(site, ns)] = np.tile(val, ns)
    err = np.abs(f.evaluate_points(pts) - tt.evaluate_points(pts)).reshape(ns, per)
    rec_err = np.zeros(ns, dtype=np.float64)
    rec_idx = np.zeros(ns, dtype=np.int64)
    for s in range(ns):  # strict '>' from 0.0 in probe order: the first maximum (:170-178)
        best, at = 0.0, -1
        for q in range(per):
            if err[s, q] > best:
                best, at = err[s, q], q
        rec_err[s], rec_idx[s] = best, at
    cap = finder.maxnglobalpivot
    piv = np.zeros((cap, n), dtype=np.int64)
    errs = np.zeros(cap, dtype=np.float64)
    acc = np.zeros(cap, dtype=np.int64)
    nf = C.c_int64(0)
    rc = lib().tci_globalsearch_select(_lib.pf(rec_err), pi(rec_idx), ns, pi(starts), n, pi(ld),
                                       float(abstol) * finder.tolmarginglobalsearch, cap, pi(piv), _lib.pf(errs), pi(acc),
                                       C.byref(nf))
    if rc != 0:
        raise ValueError("tci_globalsearch_select: bad arguments")
    finder.last_errors = errs[: nf.value].copy()
    finder.last_starts = acc[: nf.value].copy()
    return piv[: nf.value].copy()


def zcontractsitetensors(a, b, ctx=None):
    """_contractsitetensors (contraction.jl:338-349) on ComplexF64 cores (tci_zcontract_naive_site)."""
    ctx = ctx or _lib.default_context()
    a, b = _zarr(a), _zarr(b)
    Da, s1, s2, Dan = a.shape
    Db, s2b, s3, Dbn = b.shape
    if s2 != s2b:
        raise ValueError("shared site dimension mismatch")
    out = np.zeros(Da * Db * s1 * s3 * Dan * Dbn, dtype=np.complex128)
    ctx.check(lib().tci_zcontract_naive_site(ctx.h, _pz(a), Da, s1, s2, Dan, _pz(b), Db, s3, Dbn, _pz(out)))
    return out.reshape((Da * Db, s1, s3, Dan * Dbn), order="F")


def zzipup_site(ctx, R, a, b, want_dev):
    """One zip-up step (contraction.jl:455-464) on ComplexF64 data: the (chi*s1*s3) x (Da'*Db') matrix as a host array
    or as a ZDeviceMatrix for the factorisation that follows."""
    R, a, b = _zarr(R), _zarr(a), _zarr(b)
    chi, Da, Db = R.shape
    _, s1, s2, Dan = a.shape
    _, _, s3, Dbn = b.shape
    if want_dev:
        h = _lib.VP()
        ctx.check(lib().tci_zcontract_zipup_site(ctx.h, _pz(R), chi, Da, Db, _pz(a), s1, s2, Dan, _pz(b), s3, Dbn, None,
                                                 C.byref(h)))
        return ZDeviceMatrix(ctx, h)
    out = np.zeros(chi * s1 * s3 * Dan * Dbn, dtype=np.complex128)
    ctx.check(lib().tci_zcontract_zipup_site(ctx.h, _pz(R), chi, Da, Db, _pz(a), s1, s2, Dan, _pz(b), s3, Dbn, _pz(out),
                                             None))
    return out
