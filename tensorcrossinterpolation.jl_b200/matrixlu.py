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
        if self._h is None:
            raise RuntimeError("this rrLU was computed without factors (bond_update(want_factors=False))")
        L = np.zeros((m, r), dtype=np.float64, order="F")
        U = np.zeros((r, n), dtype=np.float64, order="F")
        if r:
            self.ctx.check(lib().tci_lu_fetch(self._h, pf(L), pf(U)))
        self._L, self._U = L, U

    def fetch_into(self, L, U):
        """lu.L / lu.U written into caller-provided Fortran-ordered arrays (e.g. views of page-locked memory)."""
        m, n = self._shape
        r = self.npivot
        for a, shp in ((L, (m, r)), (U, (r, n))):
            if not (a.shape == shp and a.dtype == np.float64 and a.flags.f_contiguous):
                raise ValueError(f"fetch_into expects a Fortran-ordered Float64 array of shape {shp}")
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

    def rdiv(self, B, device=False):
        """B / A for the square matrix A this object factorises to full rank (tci_lu_rdiv): the solve of
        setsitetensor! (tensorci2.jl:391).  B: DeviceMatrix or host matrix with size(A, 1) columns."""
        if not isinstance(B, DeviceMatrix):
            B = DeviceMatrix.from_host(self.ctx, np.asfortranarray(B, dtype=np.float64))
        rows, k = B.shape
        if device:
            h = C.c_void_p()
            self.ctx.check(lib().tci_lu_rdiv(self._h, B.h, None, C.byref(h)))
            return DeviceMatrix(self.ctx, h)
        out = np.zeros((rows, k), dtype=np.float64, order="F")
        self.ctx.check(lib().tci_lu_rdiv(self._h, B.h, pf(out), None))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().tci_lu_destroy(self._h)
        except Exception:
            pass


def rrlu(A, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True, exact=True, ctx=None):
    """rrlu(A; maxrank, reltol, abstol, leftorthogonal)  matrixlu.jl:217-225.
    A: host matrix (copied, like the reference's copy(A)) or a DeviceMatrix (consumed in place).  A ComplexF64 matrix
    (complex host array or ZDeviceMatrix) takes the complex kernel (complexf64.zrrlu)."""
    if getattr(A, "is_complex", False) or (not isinstance(A, DeviceMatrix) and np.iscomplexobj(A)):
        from .complexf64 import zrrlu
        return zrrlu(A, maxrank=maxrank, reltol=reltol, abstol=abstol, leftorthogonal=leftorthogonal, ctx=ctx)
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
        self.lu = A if isinstance(A, rrLU) else rrlu(A, **kwargs)  # an rrLU: the factors tci_bond_update left behind

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
        if getattr(self.lu, "is_complex", False):
            return self.lu.luci_left(device)
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
        if getattr(self.lu, "is_complex", False):
            return self.lu.luci_right(device)
        if device:
            h = C.c_void_p()
            self.lu.ctx.check(lib().tci_luci_right(self.lu._h, None, C.byref(h)))
            return DeviceMatrix(self.lu.ctx, h)
        out = np.zeros((r, n), dtype=np.float64, order="F")
        self.lu.ctx.check(lib().tci_luci_right(self.lu._h, pf(out), None))
        return out


class RookLU:
    """rrLU produced by the rook search arrlu (matrixlu.jl:227-293).  Permutations, npivot, error and
    pivoterrors are available immediately; L and U (which need the remaining rows / columns of the
    matrix, matrixlu.jl:274-288) are completed lazily on first access."""

    def __init__(self, last, rowperm, colperm, I0, J0, fsub, shape, leftorthogonal):
        self._last, self._fsub = last, fsub
        self.rowpermutation, self.colpermutation = rowperm, colperm
        self._I0, self._J0 = I0, J0
        self.npivot, self.error = last.npivot, last.error
        self._pivoterrors = last._pivoterrors
        self.leftorthogonal = leftorthogonal
        self._shape = shape
        self._L = self._U = None

    def _complete(self):
        """lu.L / lu.U for the rows / columns the search never visited (matrixlu.jl:274-288): the two triangular
        solves cols2Lmatrix! / rows2Umatrix! run on the device (tci_lu_complete: blocked TRSM + DMMA GEMM)."""
        m, n = self._shape
        r = self.npivot
        L, U = self._last.L, self._last.U
        L11, U11 = L[:r, :r], U[:r, :r]
        need_L, need_U = L.shape[0] < m, U.shape[1] < n
        I2, J2 = self.rowpermutation[r:], self.colpermutation[r:]
        A21 = A12 = None
        L2 = np.zeros((len(I2), r), dtype=np.float64, order="F") if need_L else None
        U2 = np.zeros((r, len(J2)), dtype=np.float64, order="F") if need_U else None
        if r:
            if need_L and len(I2):
                A21 = self._fsub(I2, np.asarray(self._J0, dtype=np.int64), device=True)
            if need_U and len(J2):
                A12 = self._fsub(np.asarray(self._I0, dtype=np.int64), J2, device=True)
            if A21 is not None or A12 is not None:
                ctx = self._last.ctx
                ctx.check(lib().tci_lu_complete(self._last._h, A21.h if A21 is not None else None,
                                                A12.h if A12 is not None else None, pf(L2) if A21 is not None else None,
                                                pf(U2) if A12 is not None else None))
        if need_L:
            L = np.vstack([L11, L2])
        if need_U:
            U = np.hstack([U11, U2])
        self._L, self._U = np.asfortranarray(L), np.asfortranarray(U)

    @property
    def L(self):
        if self._L is None:
            self._complete()
        return self._L

    @property
    def U(self):
        if self._U is None:
            self._complete()
        return self._U


def arrlu(fsub, matrixsize, I0=(), J0=(), maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True,
          numrookiter=5, rng=None):
    """arrlu(ValueType, f, matrixsize, I0, J0; ...) (matrixlu.jl:227-293): rook pivot search that only ever
    evaluates whole rows / columns of the matrix.  `fsub(irows, icols, device=True)` returns the submatrix
    (1-based index arrays) as a DeviceMatrix -- each rook move is one evaluation kernel plus one rrLU kernel.
    The random subsets are drawn from `rng.randindex` (injected, shared with the oracle)."""
    from .util import CounterRNG, pushrandomsubset
    rng = rng or CounterRNG(1)
    m, n = matrixsize
    I0, J0 = [int(x) for x in I0], [int(x) for x in J0]
    rowperm = np.arange(1, m + 1, dtype=np.int64)
    colperm = np.arange(1, n + 1, dtype=np.int64)
    islowrank = False
    maxrank = min(m, n) if maxrank is None else min(int(maxrank), m, n)
    last = None
    while True:
        if leftorthogonal:
            pushrandomsubset(J0, n, max(1, len(J0)), rng)
        else:
            pushrandomsubset(I0, m, max(1, len(I0)), rng)
        for rookiter in range(1, numrookiter + 1):
            colmove = (rookiter % 2 == 0) == leftorthogonal
            if colmove:
                sub = fsub(np.asarray(I0, dtype=np.int64), colperm, device=True)
            else:
                sub = fsub(rowperm, np.asarray(J0, dtype=np.int64), device=True)
            sr, sc = (len(I0), n) if colmove else (m, len(J0))
            last = rrlu(sub, maxrank=maxrank, reltol=reltol, abstol=abstol, leftorthogonal=leftorthogonal)
            pr, pc = last.rowpermutation - 1, last.colpermutation - 1
            # _optimizerrlu! keeps permuting lu.rowpermutation / lu.colpermutation in place (their prefixes)
            rowperm[:sr] = rowperm[:sr][pr]
            colperm[:sc] = colperm[:sc][pc]
            islowrank = islowrank or last.npivot < min(sr, sc)
            ri, ci = rowperm[: last.npivot].tolist(), colperm[: last.npivot].tolist()
            if ri == I0 and ci == J0:
                break
            J0, I0 = ci, ri
        if islowrank or len(I0) >= maxrank:
            break
    r = last.npivot
    if last._shape[0] < m:  # lu.rowpermutation = vcat(I0, setdiff(1:m, I0))
        have = set(I0)
        rowperm = np.array(I0 + [v for v in range(1, m + 1) if v not in have], dtype=np.int64)
    if last._shape[1] < n:
        have = set(J0)
        colperm = np.array(J0 + [v for v in range(1, n + 1) if v not in have], dtype=np.int64)
    return RookLU(last, rowperm, colperm, I0, J0, fsub, (m, n), leftorthogonal)
