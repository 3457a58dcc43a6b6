"""ctypes front-end of the CPU oracle (oracle/tci_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product never imports it.
All matrices are column-major float64, all indices 1-based (Julia conventions).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i64 = C.c_int64
f64 = C.c_double
P_i64 = C.POINTER(C.c_int64)
P_f64 = C.POINTER(C.c_double)
I64MAX = 2**63 - 1


def build(force=False):
    so = os.path.join(_HERE, "libtci_oracle.so")
    src = os.path.join(_HERE, "tci_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "tci_targets.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtci_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libtci_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_last_error.restype = C.c_char_p
        _LIB.orc_target_builtin.restype = C.c_void_p
        _LIB.orc_tt_create.restype = C.c_void_p
        _LIB.orc_mpo_pair_create.restype = C.c_void_p
        _LIB.orc_crossinterpolate2.restype = C.c_void_p
        _LIB.orc_eval_point.restype = f64
        _LIB.orc_tt_evaluate.restype = f64
        _LIB.orc_tt_sum.restype = f64
        _LIB.orc_tci_maxsamplevalue.restype = f64
        _LIB.orc_tci_evaluate.restype = f64
        _LIB.orc_tci_sum.restype = f64
        for name in ("orc_tci_niter", "orc_tci_npivoterrors", "orc_tci_indexset", "orc_tci_tracelen",
                     "orc_target_nevals"):
            getattr(_LIB, name).restype = i64
    return _LIB


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode())


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _pf(a):
    return a.ctypes.data_as(P_f64)


def _pi(a):
    return a.ctypes.data_as(P_i64)


def _flat_idx(lst, length=None):
    """list of multi-indices -> (len x count) column-major int64 array"""
    if len(lst) == 0:
        return np.zeros((length or 0, 0), dtype=np.int64, order="F")
    a = np.asarray(lst, dtype=np.int64).reshape(len(lst), -1)
    return np.asfortranarray(a.T)


class LU:
    pass


def rrlu(A, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True):
    A = _f(A)
    m, n = A.shape
    maxrank = I64MAX if maxrank is None else int(maxrank)
    mr = max(0, min(maxrank, m, n))
    rowperm = np.zeros(m, dtype=np.int64)
    colperm = np.zeros(n, dtype=np.int64)
    npiv = i64(0)
    err = f64(0.0)
    L = np.zeros(m * mr, dtype=np.float64)
    U = np.zeros(mr * n, dtype=np.float64)
    pe = np.zeros(min(m, n) + 1, dtype=np.float64)
    _check(lib().orc_rrlu(_pf(A), i64(m), i64(n), i64(maxrank), f64(reltol), f64(abstol), C.c_int(int(leftorthogonal)),
                          _pi(rowperm), _pi(colperm), C.byref(npiv), C.byref(err), _pf(L), _pf(U), _pf(pe)))
    r = npiv.value
    out = LU()
    out.rowpermutation, out.colpermutation = rowperm, colperm
    out.npivot, out.error = r, err.value
    out.L = L[: m * r].reshape((m, r), order="F")
    out.U = U[: r * n].reshape((r, n), order="F")
    out.pivoterrors = pe[: r + 1].copy()
    out.leftorthogonal = leftorthogonal
    return out


def luci(A, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True):
    A = _f(A)
    m, n = A.shape
    maxrank = I64MAX if maxrank is None else int(maxrank)
    mr = max(0, min(maxrank, m, n))
    rowperm = np.zeros(m, dtype=np.int64)
    colperm = np.zeros(n, dtype=np.int64)
    npiv = i64(0)
    left = np.zeros(m * mr, dtype=np.float64)
    right = np.zeros(mr * n, dtype=np.float64)
    pe = np.zeros(min(m, n) + 1, dtype=np.float64)
    _check(lib().orc_luci(_pf(A), i64(m), i64(n), i64(maxrank), f64(reltol), f64(abstol), C.c_int(int(leftorthogonal)),
                          _pi(rowperm), _pi(colperm), C.byref(npiv), _pf(pe), _pf(left), _pf(right)))
    r = npiv.value
    out = LU()
    out.rowpermutation, out.colpermutation, out.npivot = rowperm, colperm, r
    out.rowindices, out.colindices = rowperm[:r].copy(), colperm[:r].copy()
    out.left = left[: m * r].reshape((m, r), order="F")
    out.right = right[: r * n].reshape((r, n), order="F")
    out.pivoterrors = pe[: r + 1].copy()
    return out


def arrlu(A, I0=(), J0=(), maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True, seed=1):
    """arrlu(Float64, (i, j) -> A[i, j], size(A), I0, J0; ...) with injected random subsets."""
    A = _f(A)
    m, n = A.shape
    maxrank = I64MAX if maxrank is None else int(maxrank)
    mr = max(0, min(maxrank, m, n))
    I0 = np.ascontiguousarray(I0, dtype=np.int64)
    J0 = np.ascontiguousarray(J0, dtype=np.int64)
    rowperm = np.zeros(m, dtype=np.int64)
    colperm = np.zeros(n, dtype=np.int64)
    npiv, err = i64(0), f64(0.0)
    L = np.zeros(m * mr, dtype=np.float64)
    U = np.zeros(mr * n, dtype=np.float64)
    _check(lib().orc_arrlu(_pf(A), i64(m), i64(n), _pi(I0), i64(I0.size), _pi(J0), i64(J0.size), i64(maxrank),
                           f64(reltol), f64(abstol), C.c_int(int(leftorthogonal)), C.c_uint64(seed), _pi(rowperm),
                           _pi(colperm), C.byref(npiv), C.byref(err), _pf(L), _pf(U)))
    r = npiv.value
    out = LU()
    out.rowpermutation, out.colpermutation, out.npivot, out.error = rowperm, colperm, r, err.value
    out.L = L[: m * r].reshape((m, r), order="F")
    out.U = U[: r * n].reshape((r, n), order="F")
    out.leftorthogonal = leftorthogonal
    return out


def argmax_abs2(A, k=1):
    A = _f(A)
    r, c = i64(0), i64(0)
    lib().orc_argmax_abs2(_pf(A), i64(A.shape[0]), i64(A.shape[1]), i64(k), C.byref(r), C.byref(c))
    return r.value, c.value


def _core_ptrs(cores):
    keep = [np.asfortranarray(c, dtype=np.float64) for c in cores]
    arr = (P_f64 * len(keep))(*[_pf(c) for c in keep])
    return keep, arr


def _dims(cores, nd):
    return np.ascontiguousarray(np.array([c.shape for c in cores], dtype=np.int64).reshape(len(cores), nd))


class Target:
    def __init__(self, handle, localdims, keep=None):
        self.h = C.c_void_p(handle)
        self.localdims = [int(d) for d in localdims]
        self._keep = keep

    @classmethod
    def builtin(cls, kind, params, localdims):
        p = np.ascontiguousarray(params, dtype=np.float64)
        ld = np.ascontiguousarray(localdims, dtype=np.int64)
        h = lib().orc_target_builtin(C.c_int(kind), _pf(p), i64(p.size), _pi(ld), i64(ld.size))
        return cls(h, ld)

    @classmethod
    def tt(cls, cores):
        keep, arr = _core_ptrs(cores)
        d = _dims(keep, 3)
        h = lib().orc_tt_create(i64(len(keep)), _pi(d), arr)
        return cls(h, d[:, 1])

    @classmethod
    def mpo_pair(cls, A, B):
        ka, pa = _core_ptrs(A)
        kb, pb = _core_ptrs(B)
        da, db = _dims(ka, 4), _dims(kb, 4)
        h = lib().orc_mpo_pair_create(i64(len(ka)), _pi(da), pa, _pi(db), pb)
        return cls(h, da[:, 1] * db[:, 2])

    def __del__(self):
        try:
            lib().orc_target_destroy(self.h)
        except Exception:
            pass

    def __call__(self, idx):
        v = np.ascontiguousarray(idx, dtype=np.int64)
        return lib().orc_eval_point(self.h, _pi(v))

    def pi_eval(self, I, J, M, maxabs=None):
        """Returns the (nI, d..., nJ) array (Fortran order) and the updated max-abs."""
        n = len(self.localdims)
        if len(I) == 0 or len(J) == 0:
            return np.zeros((0,) * (M + 2), order="F"), maxabs
        Ia, Ja = _flat_idx(I), _flat_idx(J)
        nl, nI = Ia.shape
        nr, nJ = Ja.shape
        assert nl + M + nr == n
        cd = self.localdims[nl:nl + M]
        out = np.zeros(nI * int(np.prod(cd, dtype=np.int64)) * nJ, dtype=np.float64)
        ma = f64(0.0 if maxabs is None else maxabs)
        _check(lib().orc_pi_eval(self.h, _pi(Ia), i64(nl), i64(nI), _pi(Ja), i64(nr), i64(nJ), i64(M), _pf(out),
                                 C.byref(ma)))
        return out.reshape((nI, *cd, nJ), order="F"), ma.value

    @property
    def nevals(self):
        return lib().orc_target_nevals(self.h)


def tt_evaluate(cores, idx):
    keep, arr = _core_ptrs(cores)
    d = _dims(keep, 3)
    v = np.ascontiguousarray(idx, dtype=np.int64)
    return lib().orc_tt_evaluate(i64(len(keep)), _pi(d), arr, _pi(v))


def tt_sum(cores):
    keep, arr = _core_ptrs(cores)
    d = _dims(keep, 3)
    return lib().orc_tt_sum(i64(len(keep)), _pi(d), arr)


def start_points(seed, it, nsearch, localdims):
    ld = np.ascontiguousarray(localdims, dtype=np.int64)
    out = np.zeros((ld.size, nsearch), dtype=np.int64, order="F")
    lib().orc_start_points(C.c_uint64(seed), i64(it), i64(nsearch), _pi(ld), i64(ld.size), _pi(out))
    return out


def globalsearch(target, cores, starts, abstol, tolmargin=10.0, maxn=5):
    keep, arr = _core_ptrs(cores)
    d = _dims(keep, 3)
    starts = np.asfortranarray(starts, dtype=np.int64)
    n, nsearch = starts.shape
    piv = np.zeros((n, max(nsearch, 1)), dtype=np.int64, order="F")
    errs = np.zeros(max(nsearch, 1), dtype=np.float64)
    nf = i64(0)
    _check(lib().orc_globalsearch(target.h, i64(n), _pi(d), arr, _pi(starts), i64(nsearch), f64(abstol),
                                  f64(tolmargin), i64(maxn), _pi(piv), _pf(errs), C.byref(nf)))
    return [piv[:, q].tolist() for q in range(nf.value)], errs[: nf.value].copy()


def convergencecriterion(ranks, errors, nglobalpivots, tol, maxbonddim, ncheckhistory, checkconvglobalpivot=True):
    r = np.ascontiguousarray(ranks, dtype=np.int64)
    e = np.ascontiguousarray(errors, dtype=np.float64)
    g = np.ascontiguousarray(nglobalpivots, dtype=np.int64)
    return bool(lib().orc_convergencecriterion(_pi(r), _pf(e), _pi(g), i64(r.size), f64(tol), i64(maxbonddim),
                                               i64(ncheckhistory), C.c_int(int(checkconvglobalpivot))))


class _Options(C.Structure):
    _fields_ = [("tolerance", f64), ("maxbonddim", i64), ("maxiter", i64), ("sweepstrategy", C.c_int),
                ("normalizeerror", C.c_int), ("ncheckhistory", i64), ("maxnglobalpivot", i64),
                ("nsearchglobalpivot", i64), ("tolmarginglobalsearch", f64), ("strictlynested", C.c_int),
                ("checkconvglobalpivot", C.c_int), ("seed", C.c_uint64), ("pivotsearch", C.c_int)]


_STRATEGY = {"backandforth": 0, "forward": 1, "backward": 2}


class TCIResult:
    def __init__(self, h, n):
        self.h = C.c_void_p(h)
        self.n = n
        L = lib()
        k = L.orc_tci_niter(self.h)
        self.ranks = np.zeros(k, dtype=np.int64)
        self.errors = np.zeros(k, dtype=np.float64)
        self.nglobalpivots = np.zeros(k, dtype=np.int64)
        L.orc_tci_history(self.h, _pi(self.ranks), _pf(self.errors), _pi(self.nglobalpivots))
        self.maxsamplevalue = L.orc_tci_maxsamplevalue(self.h)
        self.pivoterrors = np.zeros(L.orc_tci_npivoterrors(self.h), dtype=np.float64)
        self.bonderrors = np.zeros(n - 1, dtype=np.float64)
        L.orc_tci_pivoterrors(self.h, _pf(self.pivoterrors), _pf(self.bonderrors))
        self.Iset, self.Jset = [], []
        for which, dst in ((0, self.Iset), (1, self.Jset)):
            for b in range(n):
                cnt = L.orc_tci_indexset(self.h, C.c_int(which), i64(b), None)
                ln = b if which == 0 else n - 1 - b
                buf = np.zeros((ln, cnt), dtype=np.int64, order="F")
                if ln * cnt:
                    L.orc_tci_indexset(self.h, C.c_int(which), i64(b), _pi(buf))
                dst.append([tuple(buf[:, q].tolist()) for q in range(cnt)])
        self.sitetensors = []
        for b in range(n):
            d3 = np.zeros(3, dtype=np.int64)
            L.orc_tci_coredims(self.h, i64(b), _pi(d3))
            core = np.zeros(int(np.prod(d3)), dtype=np.float64)
            L.orc_tci_core(self.h, i64(b), _pf(core))
            self.sitetensors.append(core.reshape(tuple(d3), order="F"))
        tl = L.orc_tci_tracelen(self.h)
        self.trace = np.zeros((tl, 5), dtype=np.int64)
        if tl:
            L.orc_tci_trace(self.h, _pi(self.trace))

    @property
    def linkdims(self):
        return [len(self.Iset[b + 1]) for b in range(self.n - 1)]

    def evaluate(self, idx):
        v = np.ascontiguousarray(idx, dtype=np.int64)
        return lib().orc_tci_evaluate(self.h, _pi(v))

    def sum(self):
        return lib().orc_tci_sum(self.h)

    def __del__(self):
        try:
            lib().orc_tci_destroy(self.h)
        except Exception:
            pass


def crossinterpolate2(target, localdims, initialpivots=None, tolerance=1e-8, maxbonddim=None, maxiter=20,
                      sweepstrategy="backandforth", normalizeerror=True, ncheckhistory=3, maxnglobalpivot=5,
                      nsearchglobalpivot=5, tolmarginglobalsearch=10.0, strictlynested=False,
                      checkconvglobalpivot=True, seed=1, pivotsearch="full"):
    ld = np.ascontiguousarray(localdims, dtype=np.int64)
    n = ld.size
    if initialpivots is None:
        initialpivots = [[1] * n]
    pv = _flat_idx(initialpivots, n)
    o = _Options(tolerance, I64MAX if maxbonddim is None else int(maxbonddim), maxiter, _STRATEGY[sweepstrategy],
                 int(normalizeerror), ncheckhistory, maxnglobalpivot, nsearchglobalpivot, tolmarginglobalsearch,
                 int(strictlynested), int(checkconvglobalpivot), seed, {"full": 0, "rook": 1}[pivotsearch])
    st = C.c_int(0)
    h = lib().orc_crossinterpolate2(target.h, _pi(ld), i64(n), _pi(pv), i64(pv.shape[1]), C.byref(o), C.byref(st))
    if st.value != 0:
        msg = lib().orc_last_error().decode()
        lib().orc_tci_destroy(C.c_void_p(h))
        raise OracleError(msg)
    return TCIResult(h, n)


# ---- TensorTrain -> TensorCI2 conversion (conversion.jl:73-176), restated on the C oracle's LUCI ----
def _kron_rows(Iset, d):  # kronecker(Iset, d) tensorci2.jl:315-320
    return np.array([list(i) + [s] for s in range(1, d + 1) for i in Iset], dtype=np.int64).reshape(len(Iset) * d, -1)


def _kron_cols(d, Jset):  # kronecker(d, Jset) tensorci2.jl:322-327
    return np.array([[s] + list(j) for j in Jset for s in range(1, d + 1)], dtype=np.int64).reshape(len(Jset) * d, -1)


def tt_sweep1sitegetindices(cores, forward, spectators=None, maxbonddim=None, tolerance=0.0):
    """conversion.jl:73-139 on a list of (chi, d, chi') Fortran-ordered arrays (modified in place)."""
    L = len(cores)
    sets = [np.zeros((1, 0), dtype=np.int64)]
    errs = np.zeros(max(c.shape[0] for c in cores[1:]) + 1)
    for step in range(1, L):
        here = step - 1 if forward else L - step
        nxt = step if forward else L - step - 1
        sh, shn = cores[here].shape, cores[nxt].shape
        if forward:
            f = luci(cores[here].reshape((sh[0] * sh[1], sh[2]), order="F"), maxrank=maxbonddim, abstol=tolerance,
                     leftorthogonal=True)
            sets.append(_kron_rows(sets[-1], sh[1])[f.rowindices - 1])
            if spectators:
                spectators[here] = spectators[here][f.colindices - 1]
            cores[here] = np.asfortranarray(f.left).reshape((sh[0], sh[1], f.npivot), order="F")
            cores[nxt] = np.asfortranarray(f.right @ cores[nxt].reshape((shn[0], -1), order="F")).reshape(
                (f.npivot, shn[1], shn[2]), order="F")
        else:
            f = luci(cores[here].reshape((sh[0], sh[1] * sh[2]), order="F"), maxrank=maxbonddim, abstol=tolerance,
                     leftorthogonal=False)
            sets.append(_kron_cols(sh[1], sets[-1])[f.colindices - 1])
            if spectators:
                spectators[here] = spectators[here][f.rowindices - 1]
            cores[here] = np.asfortranarray(f.right).reshape((f.npivot, sh[1], sh[2]), order="F")
            cores[nxt] = np.asfortranarray(cores[nxt].reshape((-1, shn[2]), order="F") @ f.left).reshape(
                (shn[0], shn[1], f.npivot), order="F")
        errs[: f.npivot + 1] = np.maximum(errs[: f.npivot + 1], f.pivoterrors)
    return (sets if forward else sets[::-1]), errs


def tensorci2_from_tt(cores, tolerance=1e-12, maxbonddim=None, maxiter=3):
    """conversion.jl:141-176.  Returns (Iset, Jset, cores, pivoterrors, maxsamplevalue)."""
    cores = [np.asfortranarray(c, dtype=np.float64).copy(order="F") for c in cores]
    Iset, _ = tt_sweep1sitegetindices(cores, True, None, maxbonddim, tolerance)
    Jset, pe = tt_sweep1sitegetindices(cores, False, None, maxbonddim, tolerance)
    same = lambda a, b: len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))  # noqa: E731
    for it in range(3, maxiter + 1):
        if it % 2:
            new, pe = tt_sweep1sitegetindices(cores, True, Jset)
            if same(new, Iset):
                break
        else:
            new, pe = tt_sweep1sitegetindices(cores, False, Iset)
            if same(new, Jset):
                break
    return Iset, Jset, cores, pe, max(float(np.abs(c).max()) for c in cores)


# ---- projected batchevaluate (numpy restatement; small cases) ---------------------------------------------------
def _projector_to_slice(p):  # util.jl:124-126
    return tuple(slice(None) if x == 0 else int(x) - 1 for x in p)


def tt_batchevaluate_projected(cores, sitedims, I, J, M, projector=None):
    """batchevaluate(tt::TTCache, leftindexset, rightindexset, Val(M), projector), cachedtensortrain.jl:151-215:
    cores are (Dl, d, Dr) with d = prod(sitedims[n]); the centre cores are sliced by the projector BEFORE they are
    contracted (:198-205), left/right environments are the chains of :77-128."""
    N = len(cores)
    nl, nr = len(I[0]), len(J[0])
    if N - nl - nr != M:
        raise OracleError(f"Invalid parameter M: {M}")  # :167-169
    if projector is None:
        projector = [[0] * len(sitedims[n]) for n in range(nl, N - nr)]  # :170-172
    if len(projector) != M:
        raise OracleError(f"Invalid length of projector: {projector}, correct length should be M={M}")
    lenv = np.ones((len(I), 1))
    if nl > 0:  # evaluateleft :77-100
        rows = []
        for idx in I:
            v = np.ones((1, 1))
            for s in range(nl):
                v = v @ cores[s][:, idx[s] - 1, :]
            rows.append(v[0])
        lenv = np.array(rows)
    renv = np.ones((1, len(J)))
    if nr > 0:  # evaluateright :102-128
        cols = []
        for idx in J:
            v = np.ones((1, 1))
            for s in range(nr - 1, -1, -1):
                v = cores[N - nr + s][:, idx[s] - 1, :] @ v
            cols.append(v[:, 0])
        renv = np.array(cols).T
    localdim = []
    for n in range(nl, N - nr):  # :196-208
        c = np.asarray(cores[n])
        c4 = c.reshape((c.shape[0], *sitedims[n], c.shape[-1]), order="F")
        sl = c4[(slice(None),) + _projector_to_slice(projector[n - nl]) + (slice(None),)]
        kept = [d for d, x in zip(sitedims[n], projector[n - nl]) if x == 0]
        T_ = sl.reshape((c.shape[0], int(np.prod(kept, dtype=np.int64)) if kept else 1, c.shape[-1]), order="F")
        localdim.append(T_.shape[1])
        lenv = lenv.reshape((-1, T_.shape[0]), order="F") @ T_.reshape((T_.shape[0], -1), order="F")
    lenv = lenv.reshape((-1, renv.shape[0]), order="F") @ renv  # :211-212
    return lenv.reshape((len(I), *localdim, len(J)), order="F")


def mpo_batchevaluate_projected(A, B, I, J, M, projector=None):
    """batchevaluate(obj::Contraction, ..., projector), contraction.jl:236-335 with f = nothing: the product MPO's
    site tensor (Da*Db, d1*d3, Da'*Db') with fused index i + d1*(k-1) (:95-101), its A / B factors sliced by the
    projector (:290-302) -- restated through the product core, which is what the slices of A and B multiply to."""
    prod, sitedims = [], []
    for a, b in zip(A, B):
        Da, d1, S, Dan = a.shape
        Db, _, d3, Dbn = b.shape
        # out[(la, lb), (x, z), (lan, lbn)] = sum_h a[la, x, h, lan] * b[lb, h, z, lbn]   contraction.jl:338-349
        c = np.einsum("axhc,bhzd->abxzcd", a, b)
        prod.append(np.asfortranarray(c.reshape((Da * Db, d1 * d3, Dan * Dbn), order="F")))
        sitedims.append([d1, d3])
    nl = len(I[0])
    if projector is not None:
        if len(projector) != M:
            raise OracleError(f"Length mismatch: length of projector (={len(projector)}) must be {M}")  # :250
        for k, pr in enumerate(projector):
            if len(pr) != 2:
                raise OracleError(f"Invalid projector at {nl + k + 1}: {list(pr)}, the length must be 2")  # :252
    return tt_batchevaluate_projected(prod, sitedims, I, J, M, projector)


# ---- ComplexF64 ------------------------------------------------------------------------------------
def zrrlu(A, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True):
    """rrlu on a Matrix{ComplexF64} through the C++ restatement (orc_zrrlu): Julia Base's complex arithmetic of
    include/tci_zarith.h, operation for operation.  Returns an LU object with complex L / U."""
    A = np.asfortranarray(A, dtype=np.complex128)
    m, n = A.shape
    maxrank = I64MAX if maxrank is None else int(maxrank)
    mr = max(0, min(maxrank, m, n))
    rowperm = np.zeros(m, dtype=np.int64)
    colperm = np.zeros(n, dtype=np.int64)
    npiv, err = i64(0), f64(0.0)
    L = np.zeros(m * mr, dtype=np.complex128)
    U = np.zeros(mr * n, dtype=np.complex128)
    pe = np.zeros(min(m, n) + 1, dtype=np.float64)
    fn = lib().orc_zrrlu
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, i64, i64, i64, f64, f64, C.c_int, P_i64, P_i64, C.POINTER(i64), C.POINTER(f64),
                   C.c_void_p, C.c_void_p, P_f64]
    _check(fn(A.ctypes.data, m, n, maxrank, reltol, abstol, int(leftorthogonal), _pi(rowperm), _pi(colperm),
              C.byref(npiv), C.byref(err), L.ctypes.data, U.ctypes.data, _pf(pe)))
    r = npiv.value
    out = LU()
    out.rowpermutation, out.colpermutation = rowperm, colperm
    out.npivot, out.error = r, err.value
    out.L = L[: m * r].reshape((m, r), order="F")
    out.U = U[: r * n].reshape((r, n), order="F")
    out.pivoterrors = pe[: r + 1].copy()
    out.leftorthogonal = leftorthogonal
    return out


def zluci(A, **kw):
    """MatrixLUCI on a complex matrix: left / right (matrixluci.jl:40-84) from the factors of zrrlu; the triangular
    solves and products are BLAS calls in the reference (ztrsm / zgemm, pinned to sqrt(eps) only), numpy here."""
    import scipy.linalg as sla
    lu = zrrlu(A, **kw)
    m, n = np.asarray(A).shape
    r = lu.npivot
    L, U = lu.L, lu.U
    if lu.leftorthogonal:  # colstimespivotinv :48-57, rowmatrix :44-46
        left_p = np.vstack([np.eye(r, dtype=np.complex128),
                            sla.solve_triangular(L[:r, :r].T, L[r:, :].T, lower=False, unit_diagonal=True).T]) if r else L
        right_p = L[:r, :r] @ U
    else:  # colmatrix :40-42, pivotinvtimesrows :59-68
        left_p = L @ U[:r, :r]
        right_p = np.hstack([np.eye(r, dtype=np.complex128),
                             sla.solve_triangular(U[:r, :r], U[:, r:], lower=False, unit_diagonal=True)]) if r else U
    lu.left = np.zeros((m, r), dtype=np.complex128)
    lu.left[lu.rowpermutation - 1, :] = left_p
    lu.right = np.zeros((r, n), dtype=np.complex128)
    lu.right[:, lu.colpermutation - 1] = right_p
    lu.rowindices, lu.colindices = lu.rowpermutation[:r].copy(), lu.colpermutation[:r].copy()
    return lu


# ---- ComplexF64, numpy restatement for small cases (pinned on the reference's complex arg-max literal) -----------
def _abs2c(z):  # abs2(z::Complex) = real(z)*real(z) + imag(z)*imag(z)   (Base complex.jl)
    return z.real * z.real + z.imag * z.imag


def submatrixargmax_abs2_complex(A, rows=None, cols=None):
    """submatrixargmax(abs2, A, rows, cols), matrixlu.jl:1-32, for a complex matrix: columns outer, rows inner,
    strict '>' from typemin, first maximum wins.  rows / cols are 1-based lists (None = all).  Returns 1-based (r, c)."""
    A = np.asarray(A, dtype=np.complex128)
    rows = list(range(1, A.shape[0] + 1)) if rows is None else list(rows)
    cols = list(range(1, A.shape[1] + 1)) if cols is None else list(cols)
    if not rows:
        raise OracleError("rows must not be empty")  # :10
    if not cols:
        raise OracleError("cols must not be empty")  # :11
    m, mr, mc = -np.inf, rows[0], cols[0]
    for c in cols:
        for r in rows:
            v = _abs2c(A[r - 1, c - 1])
            if v > m:
                m, mr, mc = v, r, c
    return mr, mc


def rrlu_complex(A, maxrank=None, reltol=1e-14, abstol=0.0, leftorthogonal=True):
    """_optimizerrlu! (matrixlu.jl:141-181) with swaprow!/swapcol!/addpivot! (:98-136) for a ComplexF64 matrix, element by
    element in numpy complex128: pivot metric abs2, stop rule on abs(pivot), physical swaps, scaling by true division,
    trailing update a - x*y (complex multiply then subtract).  Python's complex division is Smith's algorithm, Julia's
    `/` is a scaled variant of it: the quotients can differ in the last bit, so L/U are pinned to ~1e-15, the pivot
    ORDER exactly wherever candidates are not within rounding of each other.  Returns
    (rowperm, colperm, L, U, npivot, error), permutations 1-based."""
    A = np.array(A, dtype=np.complex128, order="F")
    m, n = A.shape
    mr = min(m, n) if maxrank is None else min(maxrank, m, n)
    rowperm, colperm = list(range(1, m + 1)), list(range(1, n + 1))
    npivot, maxerror, error = 0, 0.0, 0.0
    while npivot < mr:
        k = npivot
        pr, pc = submatrixargmax_abs2_complex(A, range(k + 1, m + 1), range(k + 1, n + 1))
        error = abs(A[pr - 1, pc - 1])
        if (error < reltol * maxerror or error < abstol) and npivot > 0:  # :155 (abs(rtol*maxerror) == rtol*maxerror)
            break
        maxerror = max(maxerror, error)
        A[[k, pr - 1], :] = A[[pr - 1, k], :]  # swaprow! :98-104
        rowperm[k], rowperm[pr - 1] = rowperm[pr - 1], rowperm[k]
        A[:, [k, pc - 1]] = A[:, [pc - 1, k]]  # swapcol! :106-112
        colperm[k], colperm[pc - 1] = colperm[pc - 1], colperm[k]
        if leftorthogonal:  # addpivot! :114-136
            for i in range(k + 1, m):
                A[i, k] = A[i, k] / A[k, k]
        else:
            for j in range(k + 1, n):
                A[k, j] = A[k, j] / A[k, k]
        for j in range(k + 1, n):
            for i in range(k + 1, m):
                A[i, j] = A[i, j] - A[i, k] * A[k, j]
        npivot += 1
    r = npivot
    L = np.tril(A[:, :r])
    U = np.triu(A[:r, :])
    if np.isnan(L).any():
        raise OracleError("lu.L contains NaNs")  # :164-166
    if np.isnan(U).any():
        raise OracleError("lu.U contains NaNs")
    if leftorthogonal:
        L[np.arange(r), np.arange(r)] = 1.0
    else:
        U[np.arange(r), np.arange(r)] = 1.0
    if r >= min(m, n):
        error = 0.0  # :176-178
    return np.array(rowperm), np.array(colperm), L, U, r, float(error)


# ---------------------------------------------------------------------------------------------------------------
# numpy / OpenBLAS restatement of Contraction.batchevaluate for M = 0 (contraction.jl:71-176, 236-335).  The
# reference's environment extensions and the final product are BLAS dgemm calls on permuted copies
# (`amat * bmat`, :92), multithreaded by default in Julia; this is the CPU baseline of the contraction stage
# (bench.py, `--impl reference`), with numpy's OpenBLAS standing in for Julia's.  Checked against the C++ oracle in
# tests/test_oracle_golden.py.
def _contract_np(a, b, idx_a, idx_b):  # _contract, contraction.jl:71-93 (0-based axes)
    rest_a = [k for k in range(a.ndim) if k not in idx_a]
    rest_b = [k for k in range(b.ndim) if k not in idx_b]
    amat = np.transpose(a, rest_a + list(idx_a)).reshape(int(np.prod([a.shape[k] for k in rest_a], dtype=np.int64)), -1)
    bmat = np.transpose(b, list(idx_b) + rest_b).reshape(-1, int(np.prod([b.shape[k] for k in rest_b], dtype=np.int64)))
    return (amat @ bmat).reshape([a.shape[k] for k in rest_a] + [b.shape[k] for k in rest_b])


def _extend_cache_np(old, a_ell, b_ell, i, j):  # _extend_cache, :103-109
    tmp1 = _contract_np(old, a_ell[:, i, :, :], (0,), (0,))
    return _contract_np(tmp1, b_ell[:, :, j, :], (0, 1), (0, 1))


class ContractionBLAS:
    """Contraction(a, b) with the reference's Dict memo of left / right environments (:112-176)."""

    def __init__(self, A, B):
        self.a = [np.asarray(x, dtype=np.float64) for x in A]
        self.b = [np.asarray(x, dtype=np.float64) for x in B]
        self.aperm = [np.transpose(x, (3, 1, 2, 0)) for x in self.a]
        self.bperm = [np.transpose(x, (3, 1, 2, 0)) for x in self.b]
        self.leftcache, self.rightcache = {}, {}
        self.extensions = 0

    def _unfuse(self, n, idx):  # _unfuse_idx, :95-97 (idx 1-based, returns 0-based (i, j))
        d1 = self.a[n].shape[1]
        return (idx - 1) % d1, (idx - 1) // d1

    def evaluateleft(self, key):  # key: tuple of 0-based (i, j) pairs for sites 0 .. len-1
        if len(key) == 0:
            return np.ones((1, 1))
        if len(key) == 1:
            i, j = key[0]
            return self.a[0][0, i, :, :].T @ self.b[0][0, :, j, :]
        if key not in self.leftcache:
            i, j = key[-1]
            ell = len(key) - 1
            self.leftcache[key] = _extend_cache_np(self.evaluateleft(key[:-1]), self.a[ell], self.b[ell], i, j)
            self.extensions += 1
        return self.leftcache[key]

    def evaluateright(self, key):  # key: (i, j) pairs for the last len(key) sites
        if len(key) == 0:
            return np.ones((1, 1))
        if len(key) == 1:
            i, j = key[0]
            return self.a[-1][:, i, :, 0] @ self.b[-1][:, :, j, 0].T
        if key not in self.rightcache:
            i, j = key[0]
            ell = len(self.a) - len(key)
            self.rightcache[key] = _extend_cache_np(self.evaluateright(key[1:]), self.aperm[ell], self.bperm[ell], i, j)
            self.extensions += 1
        return self.rightcache[key]

    def batchevaluate0(self, I, J):
        """M = 0: res[i, j] = sum_{a,b} left_[i, a, b] * right_[a, b, j]   (:268-284, :328)."""
        N = len(self.a)
        nr = len(J[0])
        left = np.stack([self.evaluateleft(tuple(self._unfuse(n, v) for n, v in enumerate(ix))) for ix in I])
        right = np.stack([self.evaluateright(tuple(self._unfuse(N - nr + n, v) for n, v in enumerate(ix))) for ix in J],
                         axis=-1)
        return _contract_np(left, right, (1, 2), (0, 1))
