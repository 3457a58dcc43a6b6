"""Multi-GPU sharding of the stages that shard: one process per GPU,
`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests).

* Pi evaluation: contiguous column blocks of Jcombined.  Every rank evaluates its block with
  K1/K4/K5 straight into its slice of the full buffer (tci_pi_eval_into), ONE in-place
  all-gather assembles Pi on every rank, and an 8-byte all-reduce(max) combines max|Pi|.
* MPO x MPO contraction (and TT) targets shard by ROW blocks of Icombined (`shard="rows"`): every
  rank extends the left environments of its own rows and the right environments of its own column block
  (tci_env_eval), ONE NCCL all-gather shares the right environments, and the rank's block Pi = left^T right
  (tci_pi_from_envs) is stored into the owner's Pi at its row offset -- so both chains, which carry nearly all the
  flops, scale with the number of GPUs.
* The per-bond rrLU stays on one GPU (rank `owner`), which broadcasts the chosen pivots
  (npivot, row/column indices, pivot errors) -- the only other collective of a bond update.
* Global pivot search: the independent start points are dealt round-robin; the accepted
  (start index, point, error) candidates are all-gathered and the reference's selection
  (start order, truncation to maxnglobalpivot, globalpivotfinder.jl:186-188) is replayed
  identically on every rank, so no further broadcast is needed.

The collectives move torch tensors that alias the library's device buffers
(`__cuda_array_interface__`); no matrix data passes through the host on the GPU path.
"""
import numpy as np

from .batcheval import BatchEvaluator


def column_blocks(ncols, world):
    """Equal-width contiguous column blocks (the last ones may be short or empty)."""
    blk = (ncols + world - 1) // world if ncols else 0
    return blk, [(min(r * blk, ncols), min((r + 1) * blk, ncols)) for r in range(world)]


def row_blocks(nrows, world, align=16):
    """Contiguous row blocks whose starts are multiples of `align` rows (128-byte lines of the column-major Pi, so
    the evaluation kernels' 16-byte stores stay aligned); the last ones may be short or empty."""
    blk = (nrows + world - 1) // world if nrows else 0
    blk = (blk + align - 1) // align * align
    return blk, [(min(r * blk, nrows), min((r + 1) * blk, nrows)) for r in range(world)]


def maxabs_allreduce(dist, torch, value, device, group=None):
    """NaN-propagating max of |x| over ranks: integer max on the bit pattern (NaN sorts above Inf),
    the same trick the evaluation kernel uses (util.jl:1-10 semantics)."""
    if value != value:
        bits = np.array([0x7FF8000000000000], dtype=np.int64)
    else:
        bits = np.array([abs(value)], dtype=np.float64).view(np.int64).copy()
    t = torch.from_numpy(bits).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.cpu().numpy().view(np.float64)[0])


def gather_column_blocks(dist, torch, full, blk, rank, group=None):
    """full: tensor (world*blk, ld), this rank's block already sits in rows [rank*blk, (rank+1)*blk).
    After the call every rank holds all blocks."""
    if blk == 0 or dist.get_world_size(group) == 1:
        return
    mine = full[rank * blk:(rank + 1) * blk]
    if full.device.type == "cpu":
        mine = mine.clone()  # gloo: no in-place aliasing
    dist.all_gather_into_tensor(full, mine, group=group)


def select_global_pivots(candidates, maxn):
    """candidates: iterable of (start_index, point, error) from all ranks -> reference order:
    found pivots are kept in start order and truncated to the first maxn (no dedup)."""
    ordered = sorted(candidates, key=lambda c: c[0])[:maxn]
    return [list(c[1]) for c in ordered], [c[2] for c in ordered]


def allgather_candidates(dist, local, group=None):
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return [c for part in out for c in part]


class PivotResult:
    """What the host driver needs from one bond factorisation (tensorci2.jl:599-605)."""

    def __init__(self, npivot, rowindices, colindices, pivoterrors):
        self.npivot = int(npivot)
        self.rowindices = np.asarray(rowindices, dtype=np.int64)
        self.colindices = np.asarray(colindices, dtype=np.int64)
        self.pivoterrors = np.asarray(pivoterrors, dtype=np.float64)


def broadcast_pivots(dist, result, owner, group=None):
    """owner -> everyone: the chosen pivots of one bond update."""
    obj = [result if dist.get_rank(group) == owner else None]
    dist.broadcast_object_list(obj, src=owner, group=group)
    return obj[0]


class ShardedEvaluator(BatchEvaluator):
    """Wraps a BatchEvaluator for world_size > 1: Pi is evaluated in column blocks (GPU path)."""

    def __init__(self, f, dist, torch, owner=0, group=None, mode="peer", shard="cols"):
        """shard "cols": column blocks of Jcombined; shard "rows" (peer mode, M = 0 calls): row blocks of Icombined,
        the partitioning of the MPO x MPO contraction -- other calls fall back to column blocks.
        mode "peer": every rank's evaluation kernel stores its column block straight into the rrLU
        owner's HBM through an IPC-mapped pointer (NVLink peer st.global; compute and transfer are the
        same kernel).  mode "allgather": per-rank blocks + one NCCL all-gather (every rank ends with Pi)."""
        self.f, self.dist, self.torch, self.owner, self.group = f, dist, torch, owner, group
        self.local = f  # unsharded evaluator for the small replicated stages (sweep1site)
        self.mode = mode
        if shard not in ("cols", "rows"):
            raise ValueError(f"Unknown sharding {shard}. Choose from cols, rows.")
        self.shard = shard
        self._ptr, self._cap = None, 0
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.ctx = f.ctx
        self.localdims = f.localdims
        self.id = f.id
        self.gather_ms = 0.0
        self.nevals = 0

    def __del__(self):  # the wrapped evaluator owns the device target
        pass

    def evaluate_points(self, pts):
        return self.f.evaluate_points(pts)

    def __call__(self, *args):
        return self.f(*args)

    def _pi(self, *a):
        return self.f._pi(*a)

    def _ensure_shared(self, nbytes):
        """(Re)allocate the peer-shared Pi arena on the owner and map it on the other ranks."""
        import ctypes as C
        from ._lib import lib
        if nbytes <= self._cap:
            return
        L, ctx, dist = lib(), self.ctx, self.dist
        self.release()
        cap = max(int(nbytes), 2 * self._cap, 1 << 22)
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        if self.rank == self.owner:
            ctx.check(L.tci_shared_alloc(ctx.h, cap, C.byref(ptr), handle))
        obj = [handle.raw if self.rank == self.owner else None]
        dist.broadcast_object_list(obj, src=self.owner, group=self.group)
        if self.rank != self.owner:
            ctx.check(L.tci_shared_open(ctx.h, obj[0], C.byref(ptr)))
        self._ptr, self._cap = ptr.value, cap

    def release(self):
        import ctypes as C
        from ._lib import lib
        if self._ptr:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            if self.rank == self.owner:
                lib().tci_shared_free(self.ctx.h, C.c_void_p(self._ptr))
            else:
                lib().tci_shared_close(self.ctx.h, C.c_void_p(self._ptr))
            self._ptr, self._cap = None, 0

    def _batch_peer(self, I, J, M, rows):
        from ._lib import DeviceMatrix
        torch, dist = self.torch, self.dist
        nJ = len(J)
        ld = (rows + 15) // 16 * 16
        self._ensure_shared(ld * nJ * 8)
        view = DeviceMatrix.wrap(self.ctx, self._ptr, rows, nJ, ld)
        if self.shard == "rows" and M == 0:
            blk, ranges = row_blocks(rows, self.world)
            lo, hi = ranges[self.rank]
            mx = 0.0
            # rows lo..hi of every column: a view that starts lo rows into the owner's buffer
            mine = DeviceMatrix.wrap(self.ctx, self._ptr + 8 * lo, hi - lo, nJ, ld) if hi > lo else None
            if getattr(self.f, "has_environments", False):
                # TT / MPO x MPO target: BOTH environment chains are sharded.  Every rank extends the right
                # environments of its column block, one NCCL all-gather (in place, over NVLink) gives every rank all
                # of them, then the rank extends the left environments of its own rows and its block of
                # Pi = left^T right goes to the owner by peer stores from the GEMM.
                D = self.f.env_dim(1, J.shape[1])
                cblk, cr = column_blocks(nJ, self.world)
                renv = DeviceMatrix.empty(self.ctx, D, cblk * self.world)
                self.f.env_eval_into(renv, self.rank * cblk, 1, J[cr[self.rank][0]:cr[self.rank][1]])
                rv = torch.as_tensor(renv, device=torch.device("cuda", self.ctx.device))
                gather_column_blocks(dist, torch, rv, cblk, self.rank, self.group)
                torch.cuda.current_stream().synchronize()
                if mine is not None:
                    lenv = DeviceMatrix.empty(self.ctx, self.f.env_dim(0, I.shape[1]), hi - lo)
                    self.f.env_eval_into(lenv, 0, 0, I[lo:hi])
                    mx = self.f.pi_from_envs(lenv, 0, hi - lo, renv, 0, nJ, mine, 0)
                    del lenv
                del rv, renv
            elif mine is not None:
                mx = self.f.batchevaluate_into(mine, 0, I[lo:hi], J, 0)
        else:
            blk, ranges = column_blocks(nJ, self.world)
            lo, hi = ranges[self.rank]
            mx = self.f.batchevaluate_into(view, lo, I, J[lo:hi], M) if hi > lo else 0.0
        torch.cuda.synchronize()  # the block is in the owner's HBM when the kernel has retired
        dist.barrier(group=self.group)
        mx = maxabs_allreduce(dist, torch, mx, torch.device("cuda", self.ctx.device), self.group)
        return view, mx

    def batchevaluate_device(self, Iset, Jset, M):
        from ._lib import DeviceMatrix
        from .util import as_indexset
        torch, dist = self.torch, self.dist
        I, J = as_indexset(Iset), as_indexset(Jset)
        nJ = len(J)
        nl = I.shape[1]
        rows = len(I) * int(np.prod(self.localdims[nl:nl + M], dtype=np.int64))
        if self.mode == "peer" and self.world > 1:
            return self._batch_peer(I, J, M, rows)
        blk, ranges = column_blocks(nJ, self.world)
        lo, hi = ranges[self.rank]
        full = DeviceMatrix.empty(self.ctx, rows, blk * self.world)
        mx = self.f.batchevaluate_into(full, self.rank * blk, I, J[lo:hi], M) if hi > lo else 0.0
        dev = torch.device("cuda", self.ctx.device)
        fv = torch.as_tensor(full, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gather_column_blocks(dist, torch, fv, blk, self.rank, self.group)
        e1.record()
        torch.cuda.current_stream().synchronize()
        self.gather_ms += e0.elapsed_time(e1)
        full.resize_cols(nJ)
        mx = maxabs_allreduce(dist, torch, mx, dev, self.group)
        return full, mx
