"""Host mirror of the BatchEvaluator interface (src/cachedtensortrain.jl:1,
src/batcheval.jl:4-83, docs/src/index.md:174-241): objects that are callable on one
multi-index and on (leftindexset, rightindexset, M).  All of them evaluate on the GPU
through tci_pi_eval / tci_target_eval; none of them can run without the library."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DeviceMatrix, lib, pf, pi
from .util import as_indexset

LORENTZ, SUM, QUANTICS2D, SEPCOS, TABLE, QUANTICS1D, GKCOSEXP = 1, 2, 3, 4, 5, 6, 7


class BatchEvaluator:
    """abstract type BatchEvaluator{V} (cachedtensortrain.jl:1) bound to a device target id."""

    def __init__(self, ctx, target_id, localdims):
        self.ctx = ctx
        self.id = int(target_id)
        self.localdims = [int(d) for d in localdims]
        self.nevals = 0

    def __len__(self):
        return len(self.localdims)

    # f(indexset) -- batcheval.jl:11-13
    def __call__(self, *args):
        if len(args) == 3:
            return self.batchevaluate(*args)
        (indexset,) = args
        return float(self.evaluate_points([indexset])[0])

    def evaluate_points(self, points):
        pts = as_indexset(points, len(self.localdims))
        out = np.zeros(pts.shape[0], dtype=np.float64)
        if pts.shape[0]:
            self.ctx.check(lib().tci_target_eval(self.ctx.h, self.id, pi(pts), pts.shape[0], pf(out)))
        self.nevals += pts.shape[0]
        return out

    def _pi(self, Iset, Jset, M, want_host, want_dev):
        n = len(self.localdims)
        nI, nJ = len(Iset), len(Jset)
        if nI * nJ == 0:  # batcheval.jl:40-42
            return np.zeros((0,) * (M + 2), order="F"), None, 0.0
        I, J = as_indexset(Iset), as_indexset(Jset)
        nl, nr = I.shape[1], J.shape[1]
        if nl + M + nr != n:
            raise RuntimeError("Invalid number of central indices")  # tensorci2.jl:307
        cd = self.localdims[nl:nl + M]
        Csz = int(np.prod(cd, dtype=np.int64)) if M else 1
        host = np.zeros(nI * Csz * nJ, dtype=np.float64) if want_host else None
        dev = C.c_void_p()
        mx = C.c_double(0.0)
        self.ctx.check(lib().tci_pi_eval(self.ctx.h, self.id, pi(I), nl, nI, pi(J), nr, nJ, M, pf(host),
                                         C.byref(dev) if want_dev else None, C.byref(mx)))
        self.nevals += nI * Csz * nJ
        if want_host:
            host = host.reshape((nI, *cd, nJ), order="F")
        return host, (DeviceMatrix(self.ctx, dev) if want_dev else None), mx.value

    # f(leftindexset, rightindexset, Val(M)) -- batcheval.jl:15-24, cachedtensortrain.jl:219-225
    def batchevaluate(self, leftindexset, rightindexset, M):
        return self._pi(leftindexset, rightindexset, M, True, False)[0]

    def batchevaluate_device(self, leftindexset, rightindexset, M):
        """Pi stays in HBM: returns (DeviceMatrix of shape (nI*prod(d_centre)) x nJ, max|Pi|)."""
        _, dev, mx = self._pi(leftindexset, rightindexset, M, False, True)
        return dev, mx

    def batchevaluate_into(self, dst, col0, leftindexset, rightindexset, M):
        """Column-block form (tci_pi_eval_into): fills dst[:, col0:col0+len(J)], returns max|block|."""
        I, J = as_indexset(leftindexset), as_indexset(rightindexset)
        if len(I) * len(J) == 0:
            return 0.0
        mx = C.c_double(0.0)
        self.ctx.check(lib().tci_pi_eval_into(self.ctx.h, self.id, pi(I), I.shape[1], I.shape[0], pi(J), J.shape[1],
                                              J.shape[0], M, dst.h, int(col0), C.byref(mx)))
        return mx.value

    # ---- environments of TT / MPO-pair targets as objects of their own (tci_env_eval) ----
    has_environments = False  # TTCache and Contraction set this

    def env_dim(self, side, length):
        D = C.c_int64(0)
        self.ctx.check(lib().tci_env_dim(self.ctx.h, self.id, int(side), int(length), C.byref(D)))
        return D.value

    def env_eval_into(self, dst, col0, side, indexset):
        """Environments (side 0: evaluateleft, side 1: evaluateright) of the entries of `indexset` into
        dst[:, col0:col0+len(indexset)]."""
        idx = as_indexset(indexset)
        if len(idx):
            self.ctx.check(lib().tci_env_eval(self.ctx.h, self.id, int(side), pi(idx), idx.shape[1], idx.shape[0],
                                              dst.h, int(col0)))

    def pi_from_envs(self, left, l0, nI, right, r0, nJ, dst, col0):
        """dst[:, col0:col0+nJ] = left[:, l0:l0+nI]^T right[:, r0:r0+nJ]; returns max|block|."""
        mx = C.c_double(0.0)
        self.ctx.check(lib().tci_pi_from_envs(self.ctx.h, left.h, int(l0), int(nI), right.h, int(r0), int(nJ), dst.h,
                                              int(col0), C.byref(mx)))
        self.nevals += int(nI) * int(nJ)
        return mx.value

    # ---- fused entry points of the driver's inner loop (tci_bond_update / tci_fill_sitetensors) ----
    def bond_update(self, Icombined, Jcombined, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True,
                    want_factors=False, exact=True):
        """The `:full` branch of updatepivots! (tensorci2.jl:529-551) in one library call: Pi evaluation (sharded over
        the context's GPUs when that pays) -> rrLU on the owner, one host synchronisation.  Returns
        (rrLU object without factors unless want_factors, max|Pi|)."""
        from .matrixlu import rrLU
        I, J = as_indexset(Icombined), as_indexset(Jcombined)
        m, n = len(I), len(J)
        if m == 0 or n == 0:
            raise ValueError("rows must not be empty")  # matrixlu.jl:10
        if I.shape[1] + J.shape[1] != len(self.localdims):
            raise RuntimeError("Invalid number of central indices")
        rowperm = np.zeros(m, dtype=np.int64)
        colperm = np.zeros(n, dtype=np.int64)
        npiv, err, mx = C.c_int64(0), C.c_double(0.0), C.c_double(0.0)
        pe = np.zeros(min(m, n) + 1, dtype=np.float64)
        h = C.c_void_p()
        mr = 0 if maxrank is None else int(min(maxrank, 2**62))
        self.ctx.check(lib().tci_bond_update(self.ctx.h, self.id, pi(I), I.shape[1], m, pi(J), J.shape[1], n, mr,
                                             float(reltol), float(abstol), int(bool(leftorthogonal)), int(bool(exact)),
                                             pi(rowperm), pi(colperm), C.byref(npiv), C.byref(err), pf(pe), C.byref(mx),
                                             C.byref(h) if want_factors else None))
        self.nevals += m * n
        r = npiv.value
        return rrLU(self.ctx, h if want_factors else None, rowperm, colperm, r, err.value, pe[: r + 1].copy(),
                    bool(leftorthogonal), (m, n)), mx.value

    def sweep2site_half(self, Isets, Jsets, extraI, extraJ, forward, reltol=1e-14, abstol=0.0, maxbonddim=None,
                        exact=True):
        """The bond loop of one half-sweep of sweep2site! (tensorci2.jl:866-907) in one library call
        (tci_sweep2site_half + tci_sweep2site_fetch).  Returns (Isets, Jsets, bonderrors, pivoterrors, max|Pi|, trace)."""
        n = len(self.localdims)
        Is = [as_indexset(s, b) for b, s in enumerate(Isets)]
        Js = [as_indexset(s, n - 1 - b) for b, s in enumerate(Jsets)]
        nI = np.array([len(s) for s in Is], dtype=np.int64)
        nJ = np.array([len(s) for s in Js], dtype=np.int64)
        Ip = (_lib.P_i64 * n)(*[pi(s) for s in Is])
        Jp = (_lib.P_i64 * n)(*[pi(s) for s in Js])
        if extraI is not None:
            eI = [as_indexset(s, b) for b, s in enumerate(extraI)]
            eJ = [as_indexset(s, n - 1 - b) for b, s in enumerate(extraJ)]
            neI = np.array([len(s) for s in eI], dtype=np.int64)
            neJ = np.array([len(s) for s in eJ], dtype=np.int64)
            eIp = (_lib.P_i64 * n)(*[pi(s) for s in eI])
            eJp = (_lib.P_i64 * n)(*[pi(s) for s in eJ])
            extra = (eIp, pi(neI), eJp, pi(neJ))
        else:
            extra = (None, None, None, None)
        nIo = np.zeros(n, dtype=np.int64)
        nJo = np.zeros(n, dtype=np.int64)
        npe = C.c_int64(0)
        mb = 0 if maxbonddim is None or maxbonddim >= 2**62 else int(maxbonddim)
        rc = lib().tci_sweep2site_half(self.ctx.h, self.id, int(bool(forward)), Ip, pi(nI), Jp, pi(nJ), *extra, float(reltol),
                                       float(abstol), mb, int(bool(exact)), pi(nIo), pi(nJo), C.byref(npe))
        self.ctx.check(rc)
        Io = [np.zeros((int(nIo[b]), b), dtype=np.int64) for b in range(n)]
        Jo = [np.zeros((int(nJo[b]), n - 1 - b), dtype=np.int64) for b in range(n)]
        Iop = (_lib.P_i64 * n)(*[pi(s) for s in Io])
        Jop = (_lib.P_i64 * n)(*[pi(s) for s in Jo])
        be = np.zeros(max(n - 1, 0), dtype=np.float64)
        pe = np.zeros(npe.value, dtype=np.float64)
        mx = C.c_double(0.0)
        trace = np.zeros(4 * max(n - 1, 0), dtype=np.int64)
        self.ctx.check(lib().tci_sweep2site_fetch(self.ctx.h, Iop, Jop, pf(be), pf(pe), C.byref(mx), pi(trace)))
        tr = trace.reshape(-1, 4)
        self.nevals += int(np.sum(tr[:, 1] * tr[:, 2]))
        return Io, Jo, be, pe, mx.value, [tuple(int(v) for v in row) for row in tr]

    def fill_sitetensors(self, Isets, Jsets, want_handle=True, want_host=True):
        """fillsitetensors! (globalsearch.jl:97-103) for all sites in one library call (tci_fill_sitetensors).
        Returns (list of T_b as (nI_b, d_b, nJ_b) arrays, max over |Pi1|, device-resident TT handle or None).
        want_host=False leaves the tensors in HBM (the list is None): they are read through the handle on demand."""
        n = len(self.localdims)
        Is = [as_indexset(s, b) for b, s in enumerate(Isets)]
        Js = [as_indexset(s, n - 1 - b) for b, s in enumerate(Jsets)]
        nI = np.array([len(s) for s in Is], dtype=np.int64)
        nJ = np.array([len(s) for s in Js], dtype=np.int64)
        Ts = ([np.zeros((int(nI[b]), self.localdims[b], int(nJ[b])), dtype=np.float64, order="F") for b in range(n)]
              if want_host else None)
        Ip = (_lib.P_i64 * n)(*[pi(s) for s in Is])
        Jp = (_lib.P_i64 * n)(*[pi(s) for s in Js])
        Tp = (_lib.P_f64 * n)(*[pf(t) for t in Ts]) if want_host else None
        mx = C.c_double(0.0)
        tid = C.c_int64(0)
        rc = lib().tci_fill_sitetensors(self.ctx.h, self.id, n, Ip, pi(nI), Jp, pi(nJ), Tp, C.byref(mx),
                                        C.byref(tid) if want_handle else None)
        if rc == _lib.TCI_ERR_ARG or rc == _lib.TCI_ERR_SINGULAR:
            raise RuntimeError(lib().tci_last_error(self.ctx.h).decode())
        self.ctx.check(rc)
        self.nevals += int(sum(nI[b] * self.localdims[b] * nJ[b] for b in range(n)) + sum(nJ[b] ** 2 for b in range(n - 1)))
        handle = DeviceTT(self.ctx, tid.value, self.localdims) if want_handle else None
        return Ts, mx.value, handle

    def __del__(self):
        try:
            lib().tci_target_destroy(self.ctx.h, self.id)
        except Exception:
            pass


class DeviceTT(BatchEvaluator):
    """A tensor train whose cores live on the device (tci_fill_sitetensors' `tt_id`): the `current_tt` the global
    pivot finder probes (globalpivotfinder.jl:160-183) without any upload; also a TT target like TTCache."""
    has_environments = True

    def core(self, b):
        """Site tensor b as a host array (Dl, d, Dr) (tci_tt_fetch_core)."""
        d3 = np.zeros(3, dtype=np.int64)
        self.ctx.check(lib().tci_tt_fetch_core(self.ctx.h, self.id, int(b), pi(d3), None))
        out = np.zeros(tuple(int(x) for x in d3), dtype=np.float64, order="F")
        if out.size:
            self.ctx.check(lib().tci_tt_fetch_core(self.ctx.h, self.id, int(b), None, pf(out)))
        return out


def apply_projector(res, centre_sitedims, projector):
    """The projected batchevaluate of TTCache / Contraction (cachedtensortrain.jl:170-215, contraction.jl:247-335):
    `projector[n][k] == 0` keeps sub-index k of centre site n free, a value v > 0 fixes it to v
    (projector_to_slice, util.jl:124-126).  `res` is the unprojected result (nI, d_1, ..., d_M, nJ) from the
    device; the projected one is its slice, with the free sub-indices of every centre site fused (first fastest)
    into one dimension, as the reference's `reshape(s, size(s)[1], :, size(s)[end])` leaves it."""
    nI, nJ = res.shape[0], res.shape[-1]
    full = res.reshape([nI] + [int(d) for sd in centre_sitedims for d in sd] + [nJ], order="F")
    index = [slice(None)]
    outdims = []
    for sd, pr in zip(centre_sitedims, projector):
        free = 1
        for d, v in zip(sd, pr):
            index.append(slice(None) if v == 0 else int(v) - 1)
            free *= int(d) if v == 0 else 1
        outdims.append(free)
    index.append(slice(None))
    return np.asfortranarray(full[tuple(index)]).reshape([nI] + outdims + [nJ], order="F")


class BuiltinTarget(BatchEvaluator):
    """A device-resident analytic target registered by kind id (include/tci_targets.h)."""

    def __init__(self, kind, params, localdims, ctx=None):
        ctx = ctx or _lib.default_context()
        p = np.ascontiguousarray(params, dtype=np.float64)
        ld = np.ascontiguousarray(localdims, dtype=np.int64)
        tid = C.c_int64(0)
        ctx.check(lib().tci_target_builtin(ctx.h, int(kind), pf(p) if p.size else None, p.size, pi(ld), ld.size,
                                           C.byref(tid)))
        super().__init__(ctx, tid.value, ld.tolist())
        self.kind = kind


class SourceTarget(BatchEvaluator):
    """A user-defined target given as CUDA source (tci_target_source): the device route for an arbitrary `f`
    (batcheval.jl:32-61 evaluates a Julia closure; a kernel cannot call one).  `source` defines
    `__device__ double tci_user_f(const long long *x, int n, const double *params)`."""

    def __init__(self, source, params, localdims, ctx=None):
        ctx = ctx or _lib.default_context()
        p = np.ascontiguousarray(params, dtype=np.float64)
        ld = np.ascontiguousarray(localdims, dtype=np.int64)
        tid = C.c_int64(0)
        ctx.check(lib().tci_target_source(ctx.h, source.encode(), pf(p) if p.size else None, p.size, pi(ld), ld.size,
                                          C.byref(tid)))
        super().__init__(ctx, tid.value, ld.tolist())


def makebatchevaluatable(kind, params, localdims, ctx=None):  # batcheval.jl:9
    return BuiltinTarget(kind, params, localdims, ctx)


def _batchevaluate_dispatch(f, localdims, Iset, Jset, M):  # batcheval.jl:67-83
    if len(Iset) * len(Jset) == 0:
        return np.zeros((0,) * (M + 2), order="F")
    if not isinstance(f, BatchEvaluator):
        raise TypeError("Function `f` is not batch evaluatable")  # tensorci2.jl:726-728
    return f(Iset, Jset, M)
