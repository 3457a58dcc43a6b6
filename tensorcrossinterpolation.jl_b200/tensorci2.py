"""Host mirror of src/tensorci2.jl: the TCI2 driver.  In the deployed system this code is
the reference's own (unchanged) Julia; here it is restated in Python so that the C-ABI can
be driven and tested end to end without a Julia runtime.  It keeps the reference's names,
keyword arguments, defaults and error texts, and it only ever computes through the GPU
library: Pi / T tensors by tci_pi_eval (kept in HBM and handed straight to tci_rrlu),
factorisations by tci_rrlu / tci_luci_*, global search by tci_globalsearch.

Index sets are int64 arrays (count, len), the array form of Vector{MultiIndex}.
"""
import time

import numpy as np

from .batcheval import BatchEvaluator
from .globalpivotfinder import AbstractGlobalPivotFinder, DefaultGlobalPivotFinder, GlobalPivotSearchInput
from .matrixlu import MatrixLUCI, colindices, pivoterrors, rowindices, rrlu
from .tensortrain import TensorTrain, evaluate_points, tt_sum
from .util import (CounterRNG, as_indexset, forwardsweep, jl_max, kronecker_left, kronecker_right, pushunique, union)

I64MAX = 2**63 - 1


class TensorCI2:
    """mutable struct TensorCI2 (tensorci2.jl:6-40)."""

    def __init__(self, func_or_localdims, localdims=None, initialpivots=None, Iset=None, Jset=None):
        if localdims is None:
            func, localdims = None, func_or_localdims
        else:
            func = func_or_localdims
        if len(localdims) <= 1:
            raise RuntimeError("localdims should have at least 2 elements!")  # :27
        n = len(localdims)
        self.localdims = [int(d) for d in localdims]
        self.Iset = [np.zeros((0, b), dtype=np.int64) for b in range(n)]
        self.Jset = [np.zeros((0, n - 1 - b), dtype=np.int64) for b in range(n)]
        self.sitetensors = [np.zeros((0, d, 0), order="F") for d in self.localdims]
        self.pivoterrors = np.zeros(0)
        self.bonderrors = np.zeros(n - 1)
        self.maxsamplevalue = 0.0
        self.Iset_history = []
        self.Jset_history = []
        self.device_tt = None
        if func is None:
            return
        if Iset is not None:  # tensorci2.jl:58-72
            self.Iset = [as_indexset(s, b) for b, s in enumerate(Iset)]
            self.Jset = [as_indexset(s, n - 1 - b) for b, s in enumerate(Jset)]
            pivots = reconstractglobalpivotsfromijset(self.localdims, self.Iset, self.Jset)
        else:  # :42-53
            if initialpivots is None:
                initialpivots = [[1] * n]
            pivots = as_indexset(initialpivots, n)
            addglobalpivots(self, pivots)
        vals = func.evaluate_points(pivots)
        self.maxsamplevalue = float(np.max(np.abs(vals))) if len(vals) else 0.0
        if not abs(self.maxsamplevalue) > 0.0:
            raise RuntimeError("maxsamplevalue is zero!")
        invalidatesitetensors(self)

    def __len__(self):
        return len(self.localdims)

    def linkdims(self):  # :83-85
        return [self.Iset[b + 1].shape[0] for b in range(len(self) - 1)]

    def __call__(self, indexset):
        return evaluate(self, indexset)


def linkdims(tci):
    return tci.linkdims()


def rank(tci):
    return max(tci.linkdims())


class DeviceSiteTensors:
    """tci.sitetensors while the tensors live in HBM (fillsitetensors! leaves them on the device for the global pivot
    finder, their only consumer inside optimize!): a sequence that fetches a tensor the first time the host reads it
    (tci_tt_fetch_core).  Assigning an element turns the object into an ordinary list of host arrays."""

    def __init__(self, handle, n):
        self.handle, self._cache = handle, [None] * n

    def __len__(self):
        return len(self._cache)

    def __getitem__(self, b):
        if isinstance(b, slice):
            return [self[i] for i in range(*b.indices(len(self)))]
        if self._cache[b] is None:
            self._cache[b] = self.handle.core(b)
        return self._cache[b]

    def __iter__(self):
        return (self[b] for b in range(len(self)))


def invalidatesitetensors(tci):  # :90-95
    tci.sitetensors = [np.zeros((0, 0, 0), order="F") for _ in range(len(tci))]
    tci.device_tt = None  # the device-resident copy of the site tensors (tci_fill_sitetensors) goes with them


def issitetensorsavailable(tci):  # :100-102
    if isinstance(tci.sitetensors, DeviceSiteTensors):
        return True
    return all(t.size != 0 for t in tci.sitetensors)


def updatebonderror(tci, b, error):
    tci.bonderrors[b] = error


def maxbonderror(tci):
    return float(np.max(tci.bonderrors))


def pivoterror(tci):  # :157-159
    return maxbonderror(tci)


def updatepivoterror(tci, errors):  # :143-150 elementwise max with zero padding
    L = max(len(tci.pivoterrors), len(errors))
    a = np.zeros(L)
    a[: len(tci.pivoterrors)] = tci.pivoterrors
    b = np.zeros(L)
    b[: len(errors)] = errors
    out = np.where(np.isnan(a) | np.isnan(b), np.nan, np.maximum(a, b))
    tci.pivoterrors = out


def flushpivoterror(tci):
    tci.pivoterrors = np.zeros(0)


def updateerrors(tci, b, errors):  # :161-169
    updatebonderror(tci, b, errors[-1])
    updatepivoterror(tci, errors)


def reconstractglobalpivotsfromijset(localdims, Isets, Jsets):  # :171-188
    seen, out = set(), []
    for i in range(len(Isets)):
        for I in Isets[i]:
            for J in Jsets[i]:
                for j in range(1, localdims[i] + 1):
                    p = tuple(I.tolist()) + (j,) + tuple(J.tolist())
                    if p not in seen:
                        seen.add(p)
                        out.append(p)
    return as_indexset(out, len(localdims))


def _pushunique_all(arr, items):
    """pushunique!(arr, item) (util.jl:16-20) for every row of `items` in turn."""
    if items.shape[1] == 0:  # the empty multi-index: present as soon as the set has one entry
        return arr if arr.shape[0] else np.zeros((1, 0), dtype=np.int64)
    if arr.shape[0] == 0:
        arr = np.zeros((0, items.shape[1]), dtype=np.int64)
    return union(np.ascontiguousarray(arr), np.ascontiguousarray(items))


def addglobalpivots(tci, pivots):  # :193-213
    pivots = as_indexset(pivots, len(tci))
    if pivots.shape[0] and pivots.shape[1] != len(tci):
        raise ValueError("Please specify a pivot as one index per leg of the MPS.")
    n = len(tci)
    if pivots.shape[0] > 0:  # pushunique! of every pivot's partial indices, pivot by pivot = an order-preserving union
        for b in range(n):
            tci.Iset[b] = _pushunique_all(tci.Iset[b], pivots[:, :b])
            tci.Jset[b] = _pushunique_all(tci.Jset[b], pivots[:, b + 1:])
        invalidatesitetensors(tci)


def filltensor(f, localdims, Iset, Jset, M, device=False):
    """filltensor (tensorci2.jl:290-312).  device=True keeps the result in HBM and returns
    (DeviceMatrix (|I|*prod d) x |J|, max|.|); otherwise the host array (|I|, d..., |J|)."""
    if len(Iset) * len(Jset) == 0:
        return (None, 0.0) if device else np.zeros((0,) * (M + 2), order="F")
    if not isinstance(f, BatchEvaluator):
        raise TypeError("Function `f` is not batch evaluatable")
    N = len(localdims)
    if N - Iset.shape[1] - Jset.shape[1] != M:
        raise RuntimeError("Invalid number of central indices")  # :307
    if device:
        return f.batchevaluate_device(Iset, Jset, M)
    return f(Iset, Jset, M)


def _hostsitetensors(tci):
    if isinstance(tci.sitetensors, DeviceSiteTensors):
        tci.sitetensors = list(tci.sitetensors)


def setsitetensor(tci, b, T):  # :329-338
    _hostsitetensors(tci)
    tci.sitetensors[b] = np.asfortranarray(T).reshape(
        (tci.Iset[b].shape[0], tci.localdims[b], tci.Jset[b].shape[0]), order="F")


def updatemaxsample(tci, mx):  # :397-399 with the max-abs fused into the evaluation kernel
    tci.maxsamplevalue = jl_max(tci.maxsamplevalue, mx)


def setsitetensor_fill(tci, f, b):
    """setsitetensor!(tci, f, b) (tensorci2.jl:367-394): T_b = Pi1 * P^-1.  Pi1 and P are evaluated into HBM,
    P is factorised to full rank by the K2 kernel and the solve runs on the device (tci_lu_rdiv) in place of
    the reference's LAPACK `\\` (:391); only T_b comes back to the host."""
    n = len(tci)
    _hostsitetensors(tci)
    nI, d, nJ = tci.Iset[b].shape[0], tci.localdims[b], tci.Jset[b].shape[0]
    if b == n - 1:
        Pi1, _, mx = f._pi(tci.Iset[b], tci.Jset[b], 1, True, False)
        updatemaxsample(tci, mx)
        tci.sitetensors[b] = Pi1.reshape((nI, d, nJ), order="F")
        return tci.sitetensors[b]
    Pi1, mx = f.batchevaluate_device(tci.Iset[b], tci.Jset[b], 1)
    updatemaxsample(tci, mx)
    k = tci.Iset[b + 1].shape[0]
    if k != nJ:
        raise RuntimeError(f"Pivot matrix at bond {b + 1} is not square!")  # :388
    P, _ = f.batchevaluate_device(tci.Iset[b + 1], tci.Jset[b], 0)
    lu = rrlu(P, reltol=0.0, abstol=0.0)  # never truncates; a singular P surfaces as the NaN error of rrlu
    if lu.npivot != k:
        raise RuntimeError(f"Pivot matrix at bond {b + 1} is singular!")
    tci.sitetensors[b] = lu.rdiv(Pi1).reshape((nI, d, k), order="F")
    return tci.sitetensors[b]


def fillsitetensors(tci, f):
    """fillsitetensors! (globalsearch.jl:97-103): every setsitetensor!(tci, f, b) in ONE library call
    (tci_fill_sitetensors: all Pi1 / P evaluations, factorisations and solves queued back to back, one
    synchronisation); the cores also stay on the device for the global pivot finder."""
    for b in range(len(tci) - 1):
        if tci.Iset[b + 1].shape[0] != tci.Jset[b].shape[0]:
            raise RuntimeError(f"Pivot matrix at bond {b + 1} is not square!")  # :388
    lazy = not getattr(f, "is_complex", False)
    Ts, mx, handle = f.fill_sitetensors(tci.Iset, tci.Jset, want_host=False) if lazy else f.fill_sitetensors(tci.Iset, tci.Jset)
    updatemaxsample(tci, mx)
    tci.sitetensors = DeviceSiteTensors(handle, len(tci)) if lazy else list(Ts)
    tci.device_tt = handle


def _sanitycheck(tci):  # globalsearch.jl:106-112
    for b in range(len(tci) - 1):
        if tci.Iset[b + 1].shape[0] != tci.Jset[b].shape[0]:
            raise RuntimeError(f"Pivot matrix at bond {b + 1} is not square!")
    return True


def sweep1site(tci, f, sweepdirection="forward", reltol=1e-14, abstol=0.0, maxbonddim=I64MAX, updatetensors=True):
    """sweep1site! (tensorci2.jl:402-461)."""
    flushpivoterror(tci)
    invalidatesitetensors(tci)
    if sweepdirection not in ("forward", "backward"):
        raise ValueError(f"Unknown sweep direction {sweepdirection}: choose between :forward, :backward.")
    fwd = sweepdirection == "forward"
    n = len(tci)
    for b in (range(n - 1) if fwd else range(n - 1, 0, -1)):
        Is = kronecker_left(tci.Iset[b], tci.localdims[b]) if fwd else tci.Iset[b]
        Js = tci.Jset[b] if fwd else kronecker_right(tci.localdims[b], tci.Jset[b])
        Pi, mx = filltensor(f, tci.localdims, tci.Iset[b], tci.Jset[b], 1, device=True)
        updatemaxsample(tci, mx)
        if fwd:  # (|I|*d) x |J| is already the matrix shape
            luci = MatrixLUCI(Pi, reltol=reltol, abstol=abstol, maxrank=min(maxbonddim, I64MAX), leftorthogonal=True)
        else:  # |I| x (d*|J|): the same tensor folded differently -- refolded on the device (tci_dmat_refold)
            luci = MatrixLUCI(Pi.refold(len(Is), len(Js)), reltol=reltol, abstol=abstol,
                              maxrank=min(maxbonddim, I64MAX), leftorthogonal=False)
        nIb, nJb = tci.Iset[b].shape[0], tci.Jset[b].shape[0]
        if fwd:
            tci.Iset[b + 1] = Is[rowindices(luci) - 1]
            tci.Jset[b] = Js[colindices(luci) - 1]
        else:
            tci.Iset[b] = Is[rowindices(luci) - 1]
            tci.Jset[b - 1] = Js[colindices(luci) - 1]
        if updatetensors:
            T = luci.left() if fwd else luci.right()
            shape = (nIb, tci.localdims[b], luci.npivot) if fwd else (luci.npivot, tci.localdims[b], nJb)
            tci.sitetensors[b] = np.asfortranarray(T).reshape(shape, order="F")
            if np.isnan(tci.sitetensors[b]).any():
                raise RuntimeError(f"Error: NaN in tensor T[{b + 1}]")  # :439-441
        updateerrors(tci, b if fwd else b - 1, pivoterrors(luci))
    if updatetensors:  # :448-459
        last = n - 1 if fwd else 0
        T = f(tci.Iset[last], tci.Jset[last], 1)
        tci.sitetensors[last] = np.asfortranarray(T).reshape(
            (tci.Iset[last].shape[0], tci.localdims[last], tci.Jset[last].shape[0]), order="F")


class SubMatrix:
    """mutable struct SubMatrix (tensorci2.jl:476-503): Pi evaluated lazily on row / column subsets,
    with the running max |value| the rook branch feeds into maxsamplevalue (:569)."""

    def __init__(self, f, rows, cols):
        self.f, self.rows, self.cols = f, rows, cols
        self.maxsamplevalue = 0.0

    def __call__(self, irows, icols, device=True):
        I = np.ascontiguousarray(self.rows[np.asarray(irows) - 1])
        J = np.ascontiguousarray(self.cols[np.asarray(icols) - 1])
        host, dev, mx = self.f._pi(I, J, 0, not device, device)
        self.maxsamplevalue = jl_max(self.maxsamplevalue, mx)
        return dev if device else host.reshape((len(I), len(J)), order="F")


def _positions(subset, combined):
    """[findfirst(isequal(i), combined) for i in subset], dropping the ones that are absent (:554-555)."""
    from .util import _rowkeys
    if subset.shape[0] == 0 or combined.shape[0] == 0 or subset.shape[1] != combined.shape[1]:
        return []
    if combined.shape[1] == 0:
        return [1] * subset.shape[0]
    where = {}
    for i, k in enumerate(_rowkeys(combined)):
        where.setdefault(k, i + 1)
    return [where[k] for k in _rowkeys(subset) if k in where]


def updatepivots(tci, b, f, leftorthogonal, reltol=1e-14, abstol=0.0, maxbonddim=I64MAX, sweepdirection="forward",
                 pivotsearch="full", verbosity=0, extraIset=None, extraJset=None, set_sitetensors=False, rng=None):
    """updatepivots! (tensorci2.jl:510-607), pivotsearch = :full.  b is 0-based here."""
    invalidatesitetensors(tci)
    n = len(tci)
    if extraIset is None:
        extraIset = np.zeros((0, b + 1), dtype=np.int64)
    if extraJset is None:
        extraJset = np.zeros((0, n - 1 - b), dtype=np.int64)
    Icombined = union(kronecker_left(tci.Iset[b], tci.localdims[b]), extraIset)
    Jcombined = union(kronecker_right(tci.localdims[b + 1], tci.Jset[b + 1]), extraJset)
    if pivotsearch not in ("full", "rook"):
        raise ValueError(f"Unknown pivot search strategy {pivotsearch}. Choose from :rook, :full.")
    luci = res = None
    t1 = time.perf_counter()
    if pivotsearch == "rook":  # :552-595 (only rows / columns of Pi are ever evaluated)
        from .matrixlu import arrlu
        I0 = _positions(tci.Iset[b + 1], Icombined)
        J0 = _positions(tci.Jset[b], Jcombined)
        Pif = SubMatrix(f, Icombined, Jcombined)
        lu = arrlu(Pif, (len(Icombined), len(Jcombined)), I0, J0, reltol=reltol, abstol=abstol,
                   maxrank=min(maxbonddim, I64MAX), leftorthogonal=leftorthogonal, rng=rng)
        updatemaxsample(tci, Pif.maxsamplevalue)
        if lu.npivot > 0:
            res = lu
    t2 = time.perf_counter()
    if res is None:  # :full, or the fall back of :573-588 -- Pi evaluation and rrLU fused in the library
        want = set_sitetensors and len(extraIset) == 0 and len(extraJset) == 0
        res, mx = f.bond_update(Icombined, Jcombined, maxrank=min(maxbonddim, I64MAX), reltol=reltol, abstol=abstol,
                                leftorthogonal=leftorthogonal, want_factors=want)
        updatemaxsample(tci, mx)
        if want:
            luci = MatrixLUCI(res)
    t3 = time.perf_counter()
    if verbosity > 2:
        print(f"    Computing Pi ({len(Icombined)} x {len(Jcombined)}) at bond {b + 1} + LU: {t3 - t2} sec "
              f"(rook search {t2 - t1} sec)")
    tci.Iset[b + 1] = Icombined[rowindices(res) - 1]
    tci.Jset[b] = Jcombined[colindices(res) - 1]
    if luci is not None:  # :601-604
        setsitetensor(tci, b, luci.left())
        setsitetensor(tci, b + 1, luci.right())
    updateerrors(tci, b, pivoterrors(res))
    if hasattr(tci, "trace"):
        tci.trace.append((b + 1, len(Icombined), len(Jcombined), res.npivot))


def optfirstpivot(f, localdims, firstpivot=None, maxsweep=1000):
    """optfirstpivot (util.jl:78-109): coordinate ascent on |f| from `firstpivot`.  The reference evaluates one point
    at a time (and notes "TODO: use batch evaluation"); here every site is one M = 1 fill on the device.  Accepting
    `newval > valf` for d = 1, 2, ... in turn ends at the first index that attains the site's maximum, provided that
    maximum exceeds the current value."""
    n = len(localdims)
    pivot = [1] * n if firstpivot is None else [int(v) for v in firstpivot]
    valf = abs(float(f(pivot)))
    for _ in range(maxsweep):
        prev = valf
        for i in range(n):
            left = np.asarray([pivot[:i]], dtype=np.int64).reshape(1, i)
            right = np.asarray([pivot[i + 1:]], dtype=np.int64).reshape(1, n - i - 1)
            vals = np.abs(np.asarray(f(left, right, 1)).reshape(-1))
            best = float(np.max(vals))
            if best > valf:
                valf = best
                pivot[i] = int(np.argmax(vals)) + 1  # first maximum
        if prev == valf:
            break
    return pivot


def makecanonical(tci, f, reltol=1e-14, abstol=0.0, maxbonddim=I64MAX):
    """makecanonical! (tensorci2.jl:463-474): an exact forward half-sweep, then a truncating backward and forward
    one; only the last one sets the site tensors."""
    sweep1site(tci, f, "forward", reltol=0.0, abstol=0.0, maxbonddim=I64MAX, updatetensors=False)
    sweep1site(tci, f, "backward", reltol=reltol, abstol=abstol, maxbonddim=maxbonddim, updatetensors=False)
    sweep1site(tci, f, "forward", reltol=reltol, abstol=abstol, maxbonddim=maxbonddim, updatetensors=True)


def addglobalpivots1sitesweep(tci, f, pivots, reltol=1e-14, abstol=0.0, maxbonddim=I64MAX):
    """addglobalpivots1sitesweep! (tensorci2.jl:219-229)."""
    addglobalpivots(tci, as_indexset(pivots, len(tci)))
    makecanonical(tci, f, reltol=reltol, abstol=abstol, maxbonddim=maxbonddim)


def existaspivot(tci, indexset):
    """existaspivot (tensorci2.jl:232-236): per bond, is (indexset[1:b-1], indexset[b+1:end]) in Iset[b] x Jset[b]?"""
    x = [int(v) for v in indexset]
    n = len(tci)
    out = []
    for b in range(n):
        left = np.asarray(x[:b], dtype=np.int64)
        right = np.asarray(x[b + 1:], dtype=np.int64)
        inI = bool(np.any(np.all(tci.Iset[b] == left, axis=1))) if tci.Iset[b].shape[0] else False
        inJ = bool(np.any(np.all(tci.Jset[b] == right, axis=1))) if tci.Jset[b].shape[0] else False
        out.append(inI and inJ)
    return out


def addglobalpivots2sitesweep(tci, f, pivots, tolerance=1e-8, normalizeerror=True, maxbonddim=I64MAX,
                              pivotsearch="full", verbosity=0, ntry=10, strictlynested=False, rng=None):
    """addglobalpivots2sitesweep! (tensorci2.jl:243-288): add the pivots, sweep twice, and retry with the pivots
    that are still not interpolated, at most ntry times.  Returns the number of pivots left."""
    n = len(tci)
    pivots = as_indexset(pivots, n)
    if pivots.shape[1] != n:
        raise ValueError("DimensionMismatch: Please specify a pivot as one index per leg of the MPS.")
    pivots_ = pivots
    for _ in range(ntry):
        abstol = tolerance * (tci.maxsamplevalue if normalizeerror else 1.0)
        addglobalpivots(tci, pivots_)
        sweep2site(tci, f, 2, abstol=abstol, maxbonddim=maxbonddim, pivotsearch=pivotsearch,
                   strictlynested=strictlynested, verbosity=verbosity, rng=rng)
        vals = evaluate_points(TensorTrain(tci.sitetensors), pivots)
        exact = f.evaluate_points(pivots)
        newpivots = pivots[np.abs(vals - exact) > abstol]  # NB: tested on ALL requested pivots (:275)
        if verbosity > 0:
            print(f"Trying to add {len(pivots_)} global pivots, {len(newpivots)} still remain.")
        if len(newpivots) == 0 or {tuple(p) for p in newpivots.tolist()} == {tuple(p) for p in pivots_.tolist()}:
            return len(newpivots)
        pivots_ = newpivots
    return len(pivots_)


def sweep0site(tci, f, b, reltol=1e-14, abstol=0.0):
    """sweep0site! / rmbadpivots! (tensorci2.jl:341-363), b 0-based: drop the pivots of bond b whose diagonal
    entry of U is below the tolerances."""
    invalidatesitetensors(tci)
    P, mx = f.batchevaluate_device(tci.Iset[b + 1], tci.Jset[b], 0)
    updatemaxsample(tci, mx)
    F = MatrixLUCI(P, reltol=reltol, abstol=abstol, leftorthogonal=True)
    d = np.abs(np.diag(F.lu.U))
    ndiag = int(np.sum((d > abstol) & (d / d[0] > reltol))) if len(d) else 0
    tci.Iset[b + 1] = tci.Iset[b + 1][rowindices(F)[:ndiag] - 1]
    tci.Jset[b] = tci.Jset[b][colindices(F)[:ndiag] - 1]


rmbadpivots = sweep0site  # backward compatibility alias (:366)


def searchglobalpivots(tci, f, abstol, verbosity=0, nsearch=100, maxnglobalpivot=5, rng=None):
    """searchglobalpivots (tensorci2.jl:958-1000): floating-zone searches from random starts; keeps the points whose
    error exceeds abstol (keyed by the error, as the reference's Dict{Float64,MultiIndex})."""
    from .cachedtensortrain import TTCache
    from .globalsearch import _floatingzone
    if nsearch == 0 or maxnglobalpivot == 0:
        return []
    if not issitetensorsavailable(tci):
        fillsitetensors(tci, f)
    pivots = {}
    ttcache = TTCache(TensorTrain(tci.sitetensors), ctx=f.ctx)
    for _ in range(nsearch):
        pivot, error = _floatingzone(ttcache, f, earlystoptol=10 * abstol, nsweeps=100, rng=rng)
        if error > abstol:
            pivots[error] = pivot
        if len(pivots) == maxnglobalpivot:
            break
    if verbosity > 1:
        print("  No global pivot found" if not pivots
              else f"  Found {len(pivots)} global pivots: max error {max(pivots)}")
    return list(pivots.values())


def convergencecriterion(ranks, errors, nglobalpivots, tolerance, maxbonddim, ncheckhistory,
                         checkconvglobalpivot=True):  # :609-628
    if len(errors) < ncheckhistory:
        return False
    lastranks = ranks[-ncheckhistory:]
    lastng = nglobalpivots[-ncheckhistory:]
    return bool((all(e < tolerance for e in errors[-ncheckhistory:])
                 and (all(g == 0 for g in lastng) if checkconvglobalpivot else True)
                 and min(lastranks) == lastranks[-1])
                or all(r >= maxbonddim for r in lastranks))


def sweep2site(tci, f, niter, iter1=1, abstol=1e-8, maxbonddim=I64MAX, sweepstrategy="backandforth",
               pivotsearch="full", verbosity=0, strictlynested=False, fillsitetensors_=True, rng=None):
    """sweep2site! (tensorci2.jl:855-916)."""
    invalidatesitetensors(tci)
    n = len(tci)
    for it in range(iter1, iter1 + niter):
        extraI = [np.zeros((0, b), dtype=np.int64) for b in range(n)]
        extraJ = [np.zeros((0, n - 1 - b), dtype=np.int64) for b in range(n)]
        if not strictlynested and len(tci.Iset_history) > 0:
            extraI = tci.Iset_history[-1]
            extraJ = tci.Jset_history[-1]
        tci.Iset_history.append([s.copy() for s in tci.Iset])
        tci.Jset_history.append([s.copy() for s in tci.Jset])
        if len(tci.Iset_history) > 2:  # only history[end] is ever read (:874-877)
            tci.Iset_history = tci.Iset_history[-2:]
            tci.Jset_history = tci.Jset_history[-2:]
        flushpivoterror(tci)
        fwd = forwardsweep(sweepstrategy, it)
        if (pivotsearch == "full" and verbosity <= 2 and type(f).sweep2site_half is BatchEvaluator.sweep2site_half
                and not getattr(f, "is_complex", False) and not getattr(tci, "per_bond_calls", False)):
            # the bond loop of this half-sweep in ONE library call (tci_sweep2site_half): the same updatepivots!
            # sequence, with the index bookkeeping between two bonds done next to the launches
            tci.sitetensors = [np.zeros((0, 0, 0), order="F") for _ in range(n)]
            tci.device_tt = None
            Io, Jo, be, pe, mx, tr = f.sweep2site_half(tci.Iset, tci.Jset, extraI, extraJ, fwd, abstol=abstol,
                                                       maxbonddim=maxbonddim)
            tci.Iset, tci.Jset = Io, Jo
            updatemaxsample(tci, mx)
            tci.bonderrors[:] = be  # every bond is visited: updatebonderror for b = 1 .. n-1
            updatepivoterror(tci, pe)
            if hasattr(tci, "trace"):
                tci.trace.extend(tr)
            continue
        for b in (range(n - 1) if fwd else range(n - 2, -1, -1)):
            updatepivots(tci, b, f, fwd, abstol=abstol, maxbonddim=maxbonddim,
                         sweepdirection="forward" if fwd else "backward", pivotsearch=pivotsearch,
                         verbosity=verbosity, extraIset=extraI[b + 1], extraJset=extraJ[b], rng=rng)
    if fillsitetensors_:
        fillsitetensors(tci, f)


def optimize(tci, f, tolerance=None, pivottolerance=None, maxbonddim=I64MAX, maxiter=20,
             sweepstrategy="backandforth", pivotsearch="full", verbosity=0, loginterval=10, normalizeerror=True,
             ncheckhistory=3, globalpivotfinder=None, maxnglobalpivot=5, nsearchglobalpivot=5,
             tolmarginglobalsearch=10.0, strictlynested=False, checkbatchevaluatable=False,
             checkconvglobalpivot=True, rng=None):
    """optimize! (tensorci2.jl:700-850).  Returns (ranks, errors ./ normalisation)."""
    errors, ranks, nglobalpivots = [], [], []
    if checkbatchevaluatable and not isinstance(f, BatchEvaluator):
        raise RuntimeError("Function `f` is not batch evaluatable")
    if nsearchglobalpivot > 0 and nsearchglobalpivot < maxnglobalpivot:
        raise RuntimeError("nsearchglobalpivot < maxnglobalpivot!")
    if pivottolerance is not None:
        if tolerance is not None and tolerance != pivottolerance:
            raise ValueError("Got different values for pivottolerance and tolerance in optimize!(TCI2). For TCI2, "
                             "both of these options have the same meaning. Please assign only `tolerance`.")
        import warnings
        warnings.warn("The option `pivottolerance` of `optimize!(tci::TensorCI2, f)` is deprecated. Please update "
                      "your code to use `tolerance`, as `pivottolerance` will be removed in the future.")
        tol = pivottolerance
    elif tolerance is not None:
        tol = tolerance
    else:
        tol = 1e-8
    tstart = time.perf_counter()
    if maxbonddim >= I64MAX and tol <= 0:
        raise ValueError("Specify either tolerance > 0 or some maxbonddim; otherwise, the convergence criterion is "
                         "not reachable!")
    finder = globalpivotfinder if globalpivotfinder is not None else DefaultGlobalPivotFinder(
        nsearch=nsearchglobalpivot, maxnglobalpivot=maxnglobalpivot, tolmarginglobalsearch=tolmarginglobalsearch)
    if not isinstance(finder, AbstractGlobalPivotFinder) and not callable(finder):
        raise TypeError("globalpivotfinder must be an AbstractGlobalPivotFinder")
    rng = rng if rng is not None else CounterRNG(1)
    for it in range(1, maxiter + 1):
        errornormalization = tci.maxsamplevalue if normalizeerror else 1.0
        abstol = tol * errornormalization
        if verbosity > 1:
            print(f"  Walltime {time.perf_counter() - tstart} sec: starting 2site sweep", flush=True)
        sweep2site(tci, f, 2, iter1=1, abstol=abstol, maxbonddim=maxbonddim, pivotsearch=pivotsearch,
                   strictlynested=strictlynested, verbosity=verbosity, sweepstrategy=sweepstrategy,
                   fillsitetensors_=True, rng=rng)
        errors.append(pivoterror(tci))
        if verbosity > 1:
            print(f"  Walltime {time.perf_counter() - tstart} sec: start searching global pivots", flush=True)
        if isinstance(tci.sitetensors, DeviceSiteTensors):  # the cores are on the device already: nothing is copied
            current_tt = TensorTrain.__new__(TensorTrain)
            current_tt.sitetensors = tci.sitetensors
        else:
            current_tt = TensorTrain(tci.sitetensors)
        current_tt.device_handle = tci.device_tt  # the same cores, already on the device (tci_fill_sitetensors)
        inp = GlobalPivotSearchInput(tci.localdims, current_tt, tci.maxsamplevalue, tci.Iset, tci.Jset)
        globalpivots = finder(inp, f, abstol, verbosity=verbosity, rng=rng)
        addglobalpivots(tci, globalpivots)
        nglobalpivots.append(len(globalpivots))
        if verbosity > 1:
            print(f"  Walltime {time.perf_counter() - tstart} sec: done searching global pivots", flush=True)
        ranks.append(rank(tci))
        if verbosity > 0 and it % loginterval == 0:
            print(f"iteration = {it}, rank = {ranks[-1]}, error= {errors[-1]}, maxsamplevalue= "
                  f"{tci.maxsamplevalue}, nglobalpivot={len(globalpivots)}", flush=True)
        if convergencecriterion(ranks, errors, nglobalpivots, abstol, maxbonddim, ncheckhistory,
                                checkconvglobalpivot=checkconvglobalpivot):
            break
    errornormalization = tci.maxsamplevalue if normalizeerror else 1.0
    abstol = tol * errornormalization
    sweep1site(tci, f, abstol=abstol, maxbonddim=maxbonddim)
    _sanitycheck(tci)
    tci.nglobalpivots_history = nglobalpivots
    return ranks, [e / errornormalization for e in errors]


def crossinterpolate2(f, localdims, initialpivots=None, **kwargs):
    """crossinterpolate2(ValueType, f, localdims, initialpivots; kwargs...) (tensorci2.jl:943-953) for
    ValueType = Float64.  Returns (tci, ranks, errors)."""
    tci = TensorCI2(f, localdims, initialpivots)
    tci.trace = []
    ranks, errors = optimize(tci, f, **kwargs)
    return tci, ranks, errors


def evaluate(tci, indexset):
    """tci(indexset): evaluate of abstracttensortrain.jl:124-132 on the site tensors."""
    if len(indexset) != len(tci):
        raise ValueError(f"To evaluate a tt of length {len(tci)}, you have to provide {len(tci)} indices, but there "
                         f"were {len(indexset)}.")
    v = evaluate_points(TensorTrain(tci.sitetensors), [indexset])[0]
    return complex(v) if np.iscomplexobj(v) else float(v)


def tci_sum(tci):
    return tt_sum(TensorTrain(tci.sitetensors))
