"""Multi-GPU: a thin host-side view.  The sharding itself lives behind the C ABI -- `Context(devices=[...])`
(tci_ctx_create(ngpu, device_ids)) drives all GPUs of a node from one process, and tci_pi_eval / tci_bond_update /
tci_globalsearch split their work inside the library:

* Pi evaluation of an analytic target: column blocks of Jcombined; every GPU's evaluation kernel stores its block
  straight into the rrLU owner's HBM (peer st.global over NVLink), one NCCL all-reduce(max) combines max|Pi| and orders
  the owner's rrLU behind the peer stores.  A cost model shards only when it pays (Pi must cross NVLink once).
* TT / MPO x MPO targets: row blocks of the PREFIX-SORTED Icombined (entries that share partial indices share
  environments, so they go to the same GPU) -- right environments of each GPU's block of suffix-sorted columns on a
  high-priority side stream, ONE ncclAllGather, left environments of the GPU's own rows concurrently on the main stream,
  block product scattered to the caller's rows / columns of the owner's Pi.
* global pivot search: contiguous blocks of the start points; fixed-size (error, probe) records, ONE ncclAllGather;
  the reference's selection (start order, truncation, globalpivotfinder.jl:180-188) replayed on the records.
* the per-bond rrLU stays on the owner.

What remains here are the two host-only pieces of that logic, exported by the library so that they can be tested
without a GPU (tests/test_parallel_gloo.py runs them under gloo with the CPU oracle as the per-rank evaluator).
"""
import ctypes as C

import numpy as np

from ._lib import Context, lib, pf, pi, shard_range


def multi_gpu_context(devices):
    """One process, several GPUs: devices[0] owns the per-bond rrLU."""
    return Context(devices=list(devices))


def column_blocks(ncols, world):
    """The library's partition of `ncols` columns over `world` GPUs (tci_shard_range, alignment 1)."""
    ranges = [shard_range(ncols, world, r, 1) for r in range(world)]
    blk = (ncols + world - 1) // world if ncols else 0
    return blk, ranges


def row_blocks(nrows, world, align=1):
    """The library's row partition of a sharded TT / contraction Pi: contiguous blocks (tci_shard_range) of the
    prefix-sorted order (tci_shard_order); a block product is scattered to the caller's rows, so no alignment is needed."""
    ranges = [shard_range(nrows, world, r, align) for r in range(world)]
    blk = (nrows + world - 1) // world if nrows else 0
    blk = (blk + align - 1) // align * align
    return blk, ranges


def prefix_partition(indexset, world, side=0):
    """Which caller-order entries every GPU takes: blocks of the prefix- (side 0: rows) or suffix-sorted (side 1: columns)
    order.  Returns a list of index arrays, one per rank."""
    from ._lib import shard_order
    perm = shard_order(indexset, side)
    return [perm[lo:hi] for lo, hi in (shard_range(len(perm), world, r, 1) for r in range(world))]


def select_global_pivots(rec_err, rec_idx, starts, localdims, threshold, maxn):
    """tci_globalsearch_select: the selection of globalpivotfinder.jl:180-188 on gathered per-start records
    (best error, probe index within the star or -1).  starts: (nsearch, n).  Returns (pivots, errors, start indices)."""
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    nsearch, n = starts.shape
    rec_err = np.ascontiguousarray(rec_err, dtype=np.float64)
    rec_idx = np.ascontiguousarray(rec_idx, dtype=np.int64)
    ld = np.ascontiguousarray(localdims, dtype=np.int64)
    piv = np.zeros((max(maxn, 1), n), dtype=np.int64)
    errs = np.zeros(max(maxn, 1), dtype=np.float64)
    sidx = np.zeros(max(maxn, 1), dtype=np.int64)
    nf = C.c_int64(0)
    rc = lib().tci_globalsearch_select(pf(rec_err), pi(rec_idx), nsearch, pi(starts), n, pi(ld), float(threshold),
                                       int(maxn), pi(piv), pf(errs), pi(sidx), C.byref(nf))
    if rc != 0:
        raise ValueError("tci_globalsearch_select: bad arguments")
    k = nf.value
    return piv[:k].copy(), errs[:k].copy(), sidx[:k].copy()
