"""ctypes binding of libtci_b200.so (include/tci_b200.h).

This is the same ABI a Julia `ccall` shim binds.  There is no CPU
fallback: if the shared library is missing, or no CUDA device is present when a context
is created, this raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtci_b200.so")

i64 = C.c_int64
f64 = C.c_double
P_i64 = C.POINTER(C.c_int64)
P_f64 = C.POINTER(C.c_double)
VP = C.c_void_p
PP_f64 = C.POINTER(P_f64)

SYMBOLS = {
    # name: (restype, argtypes)
    "tci_version": (C.c_int, []),
    "tci_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(VP)]),
    "tci_ctx_destroy": (None, [VP]),
    "tci_ctx_ngpu": (C.c_int, [VP]),
    "tci_ctx_member_launches": (i64, [VP, C.c_int]),
    "tci_last_error": (C.c_char_p, [VP]),
    "tci_ctx_launches": (i64, [VP]),
    "tci_ctx_stream": (VP, [VP]),
    "tci_timers": (C.c_int, [VP, P_f64, i64, C.c_int]),
    "tci_dmat_create": (C.c_int, [VP, i64, i64, P_f64, C.POINTER(VP)]),
    "tci_dmat_create_async": (C.c_int, [VP, i64, i64, P_f64, C.POINTER(VP)]),
    "tci_dmat_shape": (C.c_int, [VP, P_i64, P_i64, P_i64]),
    "tci_dmat_ptr": (VP, [VP]),
    "tci_dmat_fetch": (C.c_int, [VP, P_f64]),
    "tci_dmat_destroy": (C.c_int, [VP]),
    "tci_dmat_resize_cols": (C.c_int, [VP, i64]),
    "tci_dmat_refold": (C.c_int, [VP, i64, i64, C.POINTER(VP)]),
    "tci_dmat_wrap": (C.c_int, [VP, VP, i64, i64, i64, C.POINTER(VP)]),
    "tci_target_builtin": (C.c_int, [VP, C.c_int, P_f64, i64, P_i64, i64, P_i64]),
    "tci_target_source": (C.c_int, [VP, C.c_char_p, P_f64, i64, P_i64, i64, P_i64]),
    "tci_tt_create": (C.c_int, [VP, i64, P_i64, PP_f64, P_i64]),
    "tci_tt_fetch_core": (C.c_int, [VP, i64, i64, P_i64, P_f64]),
    "tci_mpo_pair_create": (C.c_int, [VP, i64, P_i64, PP_f64, P_i64, PP_f64, P_i64]),
    "tci_target_cached": (C.c_int, [VP, i64, C.c_int, P_i64]),
    "tci_target_cache_stats": (C.c_int, [VP, i64, P_i64]),
    "tci_target_destroy": (C.c_int, [VP, i64]),
    "tci_target_set_elementwise": (C.c_int, [VP, i64, C.c_int, f64, f64]),
    "tci_target_eval": (C.c_int, [VP, i64, P_i64, i64, P_f64]),
    "tci_pi_eval": (C.c_int, [VP, i64, P_i64, i64, i64, P_i64, i64, i64, i64, P_f64, C.POINTER(VP), P_f64]),
    "tci_pi_eval_into": (C.c_int, [VP, i64, P_i64, i64, i64, P_i64, i64, i64, i64, VP, i64, P_f64]),
    "tci_env_dim": (C.c_int, [VP, i64, C.c_int, i64, P_i64]),
    "tci_env_eval": (C.c_int, [VP, i64, C.c_int, P_i64, i64, i64, VP, i64]),
    "tci_pi_from_envs": (C.c_int, [VP, VP, i64, i64, VP, i64, i64, VP, i64, P_f64]),
    "tci_rrlu": (C.c_int, [VP, P_f64, VP, i64, i64, i64, f64, f64, C.c_int, C.c_int, P_i64, P_i64, P_i64, P_f64,
                           P_f64, C.POINTER(VP)]),
    "tci_lu_fetch": (C.c_int, [VP, P_f64, P_f64]),
    "tci_luci_left": (C.c_int, [VP, P_f64, C.POINTER(VP)]),
    "tci_luci_right": (C.c_int, [VP, P_f64, C.POINTER(VP)]),
    "tci_lu_rdiv": (C.c_int, [VP, VP, P_f64, C.POINTER(VP)]),
    "tci_lu_complete": (C.c_int, [VP, VP, VP, P_f64, P_f64]),
    "tci_lu_destroy": (C.c_int, [VP]),
    "tci_fp64_peak": (C.c_int, [VP, P_f64]),
    "tci_dgemm_host": (C.c_int, [VP, C.c_int, C.c_int, i64, i64, i64, f64, P_f64, P_f64, f64, P_f64]),
    "tci_contract_zipup_site": (C.c_int, [VP, P_f64, i64, i64, i64, P_f64, i64, i64, i64, P_f64, i64, i64, P_f64,
                                          C.POINTER(VP)]),
    "tci_contract_naive_site": (C.c_int, [VP, P_f64, i64, i64, i64, i64, P_f64, i64, i64, i64, P_f64]),
    "tci_globalsearch": (C.c_int, [VP, i64, i64, P_i64, i64, f64, i64, C.c_int, P_i64, P_f64, P_i64, P_i64]),
    "tci_globalsearch_counter": (C.c_int, [VP, i64, i64, C.c_uint64, C.c_uint64, i64, f64, i64, C.c_int, P_i64, P_f64, P_i64,
                                           P_i64]),
    "tci_shard_order": (C.c_int, [P_i64, i64, i64, C.c_int, P_i64]),
    "tci_shard_range": (C.c_int, [i64, C.c_int, C.c_int, i64, P_i64, P_i64]),
    "tci_globalsearch_select": (C.c_int, [P_f64, P_i64, i64, P_i64, i64, P_i64, f64, i64, P_i64, P_f64, P_i64, P_i64]),
    "tci_bond_update": (C.c_int, [VP, i64, P_i64, i64, i64, P_i64, i64, i64, i64, f64, f64, C.c_int, C.c_int, P_i64,
                                  P_i64, P_i64, P_f64, P_f64, P_f64, C.POINTER(VP)]),
    "tci_sweep2site_half": (C.c_int, [VP, i64, C.c_int, C.POINTER(P_i64), P_i64, C.POINTER(P_i64), P_i64, C.POINTER(P_i64),
                                      P_i64, C.POINTER(P_i64), P_i64, f64, f64, i64, C.c_int, P_i64, P_i64, P_i64]),
    "tci_sweep2site_fetch": (C.c_int, [VP, C.POINTER(P_i64), C.POINTER(P_i64), P_f64, P_f64, P_f64, P_i64]),
    "tci_fill_sitetensors": (C.c_int, [VP, i64, i64, C.POINTER(P_i64), P_i64, C.POINTER(P_i64), P_i64, PP_f64, P_f64,
                                       P_i64]),
    "tci_tt_evaluate": (C.c_int, [VP, i64, P_i64, PP_f64, P_i64, i64, P_f64]),
}

TCI_OK, TCI_ERR_CUDA, TCI_ERR_ARG, TCI_ERR_NAN_L, TCI_ERR_NAN_U, TCI_ERR_CENTRE, TCI_ERR_NO_DEVICE = range(7)
TCI_ERR_BUSY, TCI_ERR_UNSUPPORTED, TCI_ERR_SINGULAR = 7, 8, 9

_lib = None


class TCIError(RuntimeError):
    """ErrorException of the Julia shim: carries the library's status code."""

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libtci_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def pf(a):
    return a.ctypes.data_as(P_f64) if a is not None else None


def pi(a):
    return a.ctypes.data_as(P_i64) if a is not None else None


def core_ptrs(cores):
    keep = [np.asfortranarray(c, dtype=np.float64) for c in cores]
    arr = (P_f64 * len(keep))(*[pf(c) for c in keep])
    return keep, arr


class Context:
    """tci_ctx: one GPU, or several GPUs of one node driven from this process (devices[0] owns the per-bond rrLU,
    the stages that shard are split over all of them inside the library); one caller at a time."""

    def __init__(self, device=0, devices=None):
        self.h = VP()
        devs = [int(device)] if devices is None else [int(d) for d in devices]
        arr = (C.c_int * len(devs))(*devs)
        rc = lib().tci_ctx_create(len(devs), arr, C.byref(self.h))
        if rc != 0:
            raise TCIError(rc, lib().tci_last_error(None).decode())
        self.devices = devs
        self.device = devs[0]

    @property
    def ngpu(self):
        return int(lib().tci_ctx_ngpu(self.h))

    def member_launches(self, k):
        return int(lib().tci_ctx_member_launches(self.h, int(k)))

    def check(self, rc):
        if rc != 0:
            msg = lib().tci_last_error(self.h).decode()
            if rc == TCI_ERR_ARG:
                raise ValueError(msg)
            raise TCIError(rc, msg)

    @property
    def stream(self):
        """cudaStream_t of this context as an integer (wrap with torch.cuda.ExternalStream)."""
        return int(lib().tci_ctx_stream(self.h) or 0)

    @property
    def launches(self):
        return int(lib().tci_ctx_launches(self.h))

    def fp64_peak(self):
        """(DFMA TFLOP/s, DMMA TFLOP/s) measured by register-resident loops on this GPU (tci_fp64_peak)."""
        out = np.zeros(2, dtype=np.float64)
        self.check(lib().tci_fp64_peak(self.h, pf(out)))
        return float(out[0]), float(out[1])

    def timers(self, reset=False):
        out = np.zeros(9, dtype=np.float64)
        lib().tci_timers(self.h, pf(out), 9, int(reset))
        names = ["pi_eval", "rrlu", "luci", "env", "globalsearch", "gemm", "h2d", "d2h", "rrlu_kernel"]
        return dict(zip(names, out.tolist()))

    def close(self):
        if self.h:
            lib().tci_ctx_destroy(self.h)
            self.h = VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default = None


def default_context():
    """TCI_B200_DEVICES="0,1,2,3" makes the default context a multi-GPU one; otherwise LOCAL_RANK's GPU (or GPU 0)."""
    global _default
    if _default is None:
        devs = os.environ.get("TCI_B200_DEVICES")
        if devs:
            _default = Context(devices=[int(x) for x in devs.split(",")])
        else:
            _default = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default


def set_default_context(ctx):
    global _default
    _default = ctx


def shard_range(n, world, rank, align=1):
    """tci_shard_range: the contiguous block of `rank` (host-only, no GPU needed)."""
    lo, hi = i64(0), i64(0)
    rc = lib().tci_shard_range(int(n), int(world), int(rank), int(align), C.byref(lo), C.byref(hi))
    if rc != 0:
        raise ValueError("tci_shard_range: bad arguments")
    return lo.value, hi.value


def shard_order(indexset, side):
    """tci_shard_order: the prefix- (side 0) / suffix-sorted (side 1) order a sharded Pi deals its rows / columns in."""
    idx = np.ascontiguousarray(indexset, dtype=np.int64)
    perm = np.zeros(idx.shape[0], dtype=np.int64)
    rc = lib().tci_shard_order(pi(idx), idx.shape[1], idx.shape[0], int(side), pi(perm))
    if rc != 0:
        raise ValueError("tci_shard_order: bad arguments")
    return perm


def gemm(A, B, ctx=None):
    """A * B for host Float64 matrices through tci_dgemm_host (the library's DMMA GEMM); the host mirror
    never multiplies matrices itself."""
    ctx = ctx or default_context()
    A = np.asfortranarray(A, dtype=np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    if A.shape[1] != B.shape[0]:
        raise ValueError(f"DimensionMismatch: A has dimensions {A.shape}, B has dimensions {B.shape}")
    M, K = A.shape
    N = B.shape[1]
    out = np.zeros((M, N), dtype=np.float64, order="F")
    if M and N and K:
        ctx.check(lib().tci_dgemm_host(ctx.h, 0, 0, M, N, K, 1.0, pf(A), pf(B), 0.0, pf(out)))
    return out


class DeviceMatrix:
    """tci_dmat handle (column-major m x n with leading dimension ld on the GPU)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.h = VP(handle) if not isinstance(handle, VP) else handle
        self._owned = True

    @classmethod
    def from_host(cls, ctx, a):
        a = np.asfortranarray(a, dtype=np.float64)
        h = VP()
        ctx.check(lib().tci_dmat_create(ctx.h, a.shape[0], a.shape[1], pf(a), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_host_async(cls, ctx, a):
        """Upload on the context's copy stream (tci_dmat_create_async): returns once the copy is enqueued, so it
        overlaps with whatever the context's main stream is running.  `a` must be Fortran-ordered Float64 (ideally
        page-locked) and is kept alive by the returned object until it is released."""
        if not (a.dtype == np.float64 and a.flags.f_contiguous and a.ndim == 2):
            raise ValueError("from_host_async needs a Fortran-ordered Float64 matrix")
        h = VP()
        ctx.check(lib().tci_dmat_create_async(ctx.h, a.shape[0], a.shape[1], pf(a), C.byref(h)))
        out = cls(ctx, h)
        out._src = a
        return out

    @classmethod
    def wrap(cls, ctx, dptr, m, n, ld):
        """Non-owning view on caller-managed device memory (tci_dmat_wrap)."""
        h = VP()
        ctx.check(lib().tci_dmat_wrap(ctx.h, VP(dptr), m, n, ld, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def empty(cls, ctx, m, n):
        h = VP()
        ctx.check(lib().tci_dmat_create(ctx.h, m, n, None, C.byref(h)))
        return cls(ctx, h)

    @property
    def shape(self):
        m, n, ld = i64(), i64(), i64()
        lib().tci_dmat_shape(self.h, C.byref(m), C.byref(n), C.byref(ld))
        return m.value, n.value

    @property
    def ld(self):
        m, n, ld = i64(), i64(), i64()
        lib().tci_dmat_shape(self.h, C.byref(m), C.byref(n), C.byref(ld))
        return ld.value

    @property
    def ptr(self):
        return lib().tci_dmat_ptr(self.h)

    @property
    def __cuda_array_interface__(self):  # padded (ld x n) view for torch.as_tensor / collectives
        m, n = self.shape
        return {"shape": (n, self.ld), "typestr": "<f8", "data": (int(self.ptr or 0), False), "version": 2}

    def to_host(self):
        m, n = self.shape
        out = np.zeros((m, n), dtype=np.float64, order="F")
        if m * n:
            self.ctx.check(lib().tci_dmat_fetch(self.h, pf(out)))
        return out

    def refold(self, m2, n2):
        """reshape(A, m2, n2) on the device (tci_dmat_refold)."""
        h = VP()
        self.ctx.check(lib().tci_dmat_refold(self.h, int(m2), int(n2), C.byref(h)))
        return DeviceMatrix(self.ctx, h)

    def resize_cols(self, n):
        self.ctx.check(lib().tci_dmat_resize_cols(self.h, int(n)))

    def release(self):
        """Give up ownership (the handle was consumed by tci_rrlu)."""
        self._owned = False

    def __del__(self):
        try:
            if self._owned and self.h:
                lib().tci_dmat_destroy(self.h)
        except Exception:
            pass
