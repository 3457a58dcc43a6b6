"""tci_b200: B200 (sm_100a) drop-in for the TensorCI2 two-site update hot path of
TensorCrossInterpolation.jl.  `csrc/` holds the CUDA kernels and the C ABI
(include/tci_b200.h -> libtci_b200.so); the Python modules mirror the reference's host
interface for this path (same names and keyword arguments) and call the ABI via ctypes.
There is no CPU fallback anywhere in this package."""
from ._lib import Context, DeviceMatrix, TCIError, default_context, lib  # noqa: F401
from .batcheval import (GKCOSEXP, LORENTZ, QUANTICS1D, QUANTICS2D, SEPCOS, SUM, TABLE, BatchEvaluator,  # noqa: F401
                        BuiltinTarget, SourceTarget, makebatchevaluatable)
from .cachedfunction import CachedFunction  # noqa: F401
from .cachedtensortrain import TTCache, isbatchevaluable  # noqa: F401
from .contraction import (Contraction, _contractsitetensors, _factorize, compress, contract, contract_naive,  # noqa: F401
                          contract_TCI, contract_zipup)
from .complexf64 import ZContraction, ZDeviceMatrix, ZrrLU, ZTTCache, zgemm, zrrlu  # noqa: F401
from .conversion import sweep1sitegetindices, tensorci2_from_tensortrain  # noqa: F401
from .globalsearch import _floatingzone, estimatetrueerror  # noqa: F401
from .globalpivotfinder import (AbstractGlobalPivotFinder, DefaultGlobalPivotFinder,  # noqa: F401
                                GlobalPivotSearchInput)
from .matrixlu import (MatrixLUCI, RookLU, arrlu, colindices, lastpivoterror, left, npivots, pivoterrors, right,  # noqa: F401
                       rowindices, rrLU, rrlu, size)
from .tensorci2 import (TensorCI2, addglobalpivots, addglobalpivots1sitesweep, addglobalpivots2sitesweep,  # noqa: F401
                        convergencecriterion, crossinterpolate2, evaluate, existaspivot, fillsitetensors, filltensor,
                        linkdims, makecanonical, optfirstpivot, optimize, pivoterror, rank, rmbadpivots, searchglobalpivots, sweep0site,
                        sweep1site, sweep2site, tci_sum, updatepivots)
from .tensortrain import (TensorTrain, add, divide, evaluate_points, fulltensor, multiply, norm, norm2, reverse,  # noqa: F401
                          sitedims, subtract, sum_dims, tensortrain, tt_sum)
from .util import CounterRNG, forwardsweep, kronecker_left, kronecker_right  # noqa: F401
