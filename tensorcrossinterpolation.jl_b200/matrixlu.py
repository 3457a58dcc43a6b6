"""Host mirror of src/matrixlu.jl + src/matrixluci.jl: same names and keyword meaning,
the work is done by tci_rrlu / tci_luci_left / tci_luci_right (K2, K3)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DeviceMatrix, lib, pf, pi

I64MAX = 2**63 - 1


class rrLU:
    """matrixlu.jl:71-96.  Fields as in the reference (1-based permutations); L and U are
    fetched lazily from the device-resident factorisation."""

    def __init__(self, ctx, handle, rowperm, colperm, npivot, error, pivoterrors, leftorthogonal, shape):
        self.ctx = ctx
        self._h = handle
        self.rowpermutation = rowperm
        self.colpermutation = colperm
        self.npivot = npivot
        self.error = error
        self._pivoterrors = pivoterrors
        self.leftorthogonal = leftorthogonal
        self._shape = shape
        self._L = None
        self._U = None

    def _fetch(self):
        m, n = self._shape
        r = self.npivot
        L = np.zeros((m, r), dtype=np.float64, order="F")
        U = np.zeros((r, n), dtype=np.float64, order="F")
        if r:
            self.ctx.check(lib().tci_lu_fetch(self._h, pf(L), pf(U)))
        self._L, self._U = L, U

    @property
    def L(self):
        if self._L is None:
            self._fetch()
        return self._L

    @property
    def U(self):
        if self._U is None:
            self._fetch()
        return self._U

    def __del__(self):
        try:
            if self._h:
                lib().tci_lu_destroy(self._h)
        except Exception:
            pass


def rrlu(A, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True, exact=True, ctx=None):
    """rrlu(A; maxrank, reltol, abstol, leftorthogonal)  matrixlu.jl:217-225.
    A: host matrix (copied, like the reference's copy(A)) or a DeviceMatrix (consumed in place)."""
    if isinstance(A, DeviceMatrix):
        ctx = A.ctx
        m, n = A.shape
        host, dev = None, A.h
    else:
        ctx = ctx or _lib.default_context()
        A = np.asfortranarray(A, dtype=np.float64)
        if A.ndim != 2:
            raise ValueError("rrlu expects a matrix")
        m, n = A.shape
        host, dev = A, None
    mr = 0 if maxrank is None else int(min(maxrank, I64MAX))
    if maxrank is not None and maxrank <= 0:
        raise ValueError("maxrank must be positive")
    rowperm = np.zeros(m, dtype=np.int64)
    colperm = np.zeros(n, dtype=np.int64)
    npiv = C.c_int64(0)
    err = C.c_double(0.0)
    pe = np.zeros(min(m, n) + 1, dtype=np.float64)
    h = C.c_void_p()
    rc = lib().tci_rrlu(ctx.h, pf(host), dev, m, n, mr, float(reltol), float(abstol), int(bool(leftorthogonal)),
                        int(bool(exact)), pi(rowperm), pi(colperm), C.byref(npiv), C.byref(err), pf(pe), C.byref(h))
    if dev is not None and (rc == 0 and h):
        A.release()  # now owned by the factorisation handle
    ctx.check(rc)
    r = npiv.value
    return rrLU(ctx, h, rowperm, colperm, r, err.value, pe[: r + 1].copy(), bool(leftorthogonal), (m, n))


def size(lu):
    return lu._shape


def npivots(lu):
    return lu.npivot


def rowindices(lu):  # matrixlu.jl:402-404
    return lu.rowpermutation[: lu.npivot]


def colindices(lu):
    return lu.colpermutation[: lu.npivot]


def pivoterrors(lu):  # matrixlu.jl:414-416
    return lu._pivoterrors


def lastpivoterror(lu):
    return lu.error


def left(lu, permute=True):
    """rrLU: matrixlu.jl:374-382; MatrixLUCI: matrixluci.jl:70-76."""
    if isinstance(lu, MatrixLUCI):
        return lu.left()
    if not permute:
        return lu.L
    out = np.zeros_like(lu.L)
    out[lu.rowpermutation - 1, :] = lu.L
    return out


def right(lu, permute=True):
    if isinstance(lu, MatrixLUCI):
        return lu.right()
    if not permute:
        return lu.U
    out = np.zeros_like(lu.U)
    out[:, lu.colpermutation - 1] = lu.U
    return out


class MatrixLUCI:
    """matrixluci.jl:1-92."""

    def __init__(self, A, **kwargs):
        self.lu = rrlu(A, **kwargs)

    # accessors shared with rrLU
    @property
    def npivot(self):
        return self.lu.npivot

    @property
    def rowpermutation(self):
        return self.lu.rowpermutation

    @property
    def colpermutation(self):
        return self.lu.colpermutation

    @property
    def error(self):
        return self.lu.error

    @property
    def _pivoterrors(self):
        return self.lu._pivoterrors

    @property
    def _shape(self):
        return self.lu._shape

    def left(self, device=False):
        m, n = self.lu._shape
        r = self.lu.npivot
        if device:
            h = C.c_void_p()
            self.lu.ctx.check(lib().tci_luci_left(self.lu._h, None, C.byref(h)))
            return DeviceMatrix(self.lu.ctx, h)
        out = np.zeros((m, r), dtype=np.float64, order="F")
        self.lu.ctx.check(lib().tci_luci_left(self.lu._h, pf(out), None))
        return out

    def right(self, device=False):
        m, n = self.lu._shape
        r = self.lu.npivot
        if device:
            h = C.c_void_p()
            self.lu.ctx.check(lib().tci_luci_right(self.lu._h, None, C.byref(h)))
            return DeviceMatrix(self.lu.ctx, h)
        out = np.zeros((r, n), dtype=np.float64, order="F")
        self.lu.ctx.check(lib().tci_luci_right(self.lu._h, pf(out), None))
        return out
