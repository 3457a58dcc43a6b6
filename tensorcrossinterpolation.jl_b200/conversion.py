"""Host mirror of the TensorTrain -> TensorCI2 conversion (conversion.jl:73-176): a direct
consumer of K2/K3.  Every factorisation is a MatrixLUCI on the GPU (tci_rrlu + tci_luci_*), every
product a tci_dgemm_host call.  The TCI1 conversions (conversion.jl:1-71) are outside the hot path.
"""
import numpy as np

from . import _lib
from .matrixlu import MatrixLUCI, colindices, npivots, pivoterrors, rowindices
from .tensorci2 import TensorCI2
from .tensortrain import rank as ttrank
from .util import kronecker_left, kronecker_right

I64MAX = 2**63 - 1


def sweep1sitegetindices(tt, forwardsweep, spectatorindices=None, maxbonddim=I64MAX, tolerance=0.0, ctx=None):
    """sweep1sitegetindices!(tt, forwardsweep, spectatorindices; maxbonddim, tolerance) (conversion.jl:73-139).
    Mutates tt.sitetensors and, like the reference (:107-115), re-indexes `spectatorindices` in place.
    Returns (indexset, pivoterrorsarray)."""
    cores = tt.sitetensors
    L = len(cores)
    indexset = [np.zeros((1, 0), dtype=np.int64)]
    pivoterrorsarray = np.zeros(ttrank(tt) + 1)
    mr = None if maxbonddim >= I64MAX else int(maxbonddim)
    for i in range(1, L):
        ell = (i if forwardsweep else L - i + 1) - 1
        ellnext = (i + 1 if forwardsweep else L - i) - 1
        shape = cores[ell].shape
        shapenext = cores[ellnext].shape
        # groupindices(T, false) :81-88
        mat = cores[ell].reshape((-1, shape[-1]) if forwardsweep else (shape[0], -1), order="F")
        luci = MatrixLUCI(np.asfortranarray(mat), leftorthogonal=forwardsweep, abstol=tolerance, maxrank=mr, ctx=ctx)
        r = npivots(luci)
        if forwardsweep:  # :106-116
            indexset.append(kronecker_left(indexset[-1], shape[1])[rowindices(luci) - 1])
            if spectatorindices:
                spectatorindices[ell] = spectatorindices[ell][colindices(luci) - 1]
        else:
            indexset.append(kronecker_right(shape[1], indexset[-1])[colindices(luci) - 1])
            if spectatorindices:
                spectatorindices[ell] = spectatorindices[ell][rowindices(luci) - 1]
        # splitindices :90-97
        fac = luci.left() if forwardsweep else luci.right()
        cores[ell] = np.asfortranarray(fac).reshape((*shape[:-1], r) if forwardsweep else (r, *shape[1:]), order="F")
        if forwardsweep:  # :124-128
            nxt = _lib.gemm(luci.right(), cores[ellnext].reshape((shapenext[0], -1), order="F"), ctx)
            cores[ellnext] = nxt.reshape((r, *shapenext[1:]), order="F")
        else:
            nxt = _lib.gemm(cores[ellnext].reshape((-1, shapenext[-1]), order="F"), luci.left(), ctx)
            cores[ellnext] = nxt.reshape((*shapenext[:-1], r), order="F")
        pe = pivoterrors(luci)
        pivoterrorsarray[: r + 1] = np.maximum(pivoterrorsarray[: r + 1], pe)  # :131
    if forwardsweep:
        return indexset, pivoterrorsarray
    return indexset[::-1], pivoterrorsarray


def tensorci2_from_tensortrain(tt, tolerance=1e-12, maxbonddim=I64MAX, maxiter=3, ctx=None):
    """TensorCI2{ValueType}(tt::TensorTrain{ValueType,3}; tolerance, maxbonddim, maxiter) (conversion.jl:141-176).
    `tt` is modified in place and its site tensors become the TCI's, as in the reference (:170).  The
    refinement sweeps (:150-162) only compare the new index sets with the old ones; they do not replace
    them -- kept as is."""
    Iset, _ = sweep1sitegetindices(tt, True, maxbonddim=maxbonddim, tolerance=tolerance, ctx=ctx)
    Jset, pe = sweep1sitegetindices(tt, False, maxbonddim=maxbonddim, tolerance=tolerance, ctx=ctx)
    for it in range(3, maxiter + 1):
        if it % 2 == 1:
            Isetnew, pe = sweep1sitegetindices(tt, True, Jset, ctx=ctx)
            if len(Isetnew) == len(Iset) and all(np.array_equal(a, b) for a, b in zip(Isetnew, Iset)):
                break
        else:
            Jsetnew, pe = sweep1sitegetindices(tt, False, Iset, ctx=ctx)
            if len(Jsetnew) == len(Jset) and all(np.array_equal(a, b) for a, b in zip(Jsetnew, Jset)):
                break
    tci2 = TensorCI2([int(t.shape[1]) for t in tt.sitetensors])
    tci2.Iset = Iset
    tci2.Jset = Jset
    tci2.sitetensors = tt.sitetensors
    tci2.pivoterrors = pe
    tci2.maxsamplevalue = max(float(np.max(np.abs(t))) for t in tt.sitetensors)
    return tci2
