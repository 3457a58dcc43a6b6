"""world_size-2 gloo tests (CPU) of the host-side logic of the sharded stages.  The sharding itself lives in the
library (tci_ctx_create(ngpu, device_ids)); its two host-only pieces -- the partition (tci_shard_range) and the
selection of the global search on gathered per-start records (tci_globalsearch_select) -- are exported by the C ABI and
need no GPU.  Here two gloo ranks play the two GPUs: each evaluates ITS block (the CPU oracle is the per-rank
evaluator, checker role only), the blocks are gathered the way the library gathers them (in-place all-gather of
equal-sized blocks; disjoint row blocks into one buffer), and the assembled result must equal the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_blocks(full, blk, rank):
    """in-place all-gather of equal-sized blocks, as group_allgather does it (csrc/group.cu)"""
    if blk == 0 or dist.get_world_size() == 1:
        return
    dist.all_gather_into_tensor(full, full[rank * blk:(rank + 1) * blk].clone())


def _maxabs_allreduce(value):
    """NaN-propagating max of |x|: integer max on the bit pattern (NaN sorts above Inf), what the library's
    all-reduce(max) on the max|.| words does (csrc/pi_eval.cu)"""
    bits = np.array([0x7FF8000000000000], dtype=np.int64) if value != value else \
        np.array([abs(value)], dtype=np.float64).view(np.int64).copy()
    t = torch.from_numpy(bits)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.numpy().view(np.float64)[0])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        import tci_b200  # noqa: F401  (loads the package; no GPU needed for parallel.py)
        from tci_b200 import parallel as P
        from oracle import oracle as orc
        res = {}
        # ---- sharded Pi: column blocks + all-gather ----
        ld = [10] * 6
        o = orc.Target.builtin(1, [1.0], ld)
        rng = np.random.default_rng(0)
        I = np.stack([rng.integers(1, 11, 23) for _ in range(3)], axis=1)
        J = np.stack([rng.integers(1, 11, 17) for _ in range(3)], axis=1)
        ref, refmx = o.pi_eval(I.tolist(), J.tolist(), 0, 0.0)
        blk, ranges = P.column_blocks(len(J), world)
        lo, hi = ranges[rank]
        ldm = 32
        full = torch.zeros((blk * world, ldm), dtype=torch.float64)
        mine, mx = o.pi_eval(I.tolist(), J[lo:hi].tolist(), 0, 0.0)
        full[rank * blk:rank * blk + (hi - lo), :23] = torch.from_numpy(np.ascontiguousarray(mine.T))
        _gather_blocks(full, blk, rank)
        got = full[:len(J), :23].numpy().T
        res["pi_equal"] = bool(np.array_equal(got, ref))
        res["maxabs"] = _maxabs_allreduce(mx) == refmx
        nanmax = _maxabs_allreduce(float("nan") if rank == 1 else 1.0)
        res["nan"] = nanmax != nanmax
        # ---- sharded Pi: row blocks (the MPO x MPO partitioning), each rank's rows written at their offset ----
        rblk, rranges = P.row_blocks(len(I), world)
        rlo, rhi = rranges[rank]
        part = torch.zeros((len(J), ldm), dtype=torch.float64)
        if rhi > rlo:
            blockv, bmx = o.pi_eval(I[rlo:rhi].tolist(), J.tolist(), 0, 0.0)
            part[:, rlo:rhi] = torch.from_numpy(np.ascontiguousarray(blockv.T))
        else:
            bmx = 0.0
        dist.all_reduce(part)  # stands in for the peer stores into the owner's buffer (disjoint row ranges)
        res["pi_rows_equal"] = bool(np.array_equal(part[:, :23].numpy().T, ref))
        res["maxabs_rows"] = _maxabs_allreduce(bmx) == refmx
        # ---- contraction Pi sharded by rows with BOTH environment chains sharded (ShardedEvaluator shard="rows"):
        # right environments of the rank's column block -> all-gather -> left environments of its rows -> block
        # product at its row offset.  Environments from a numpy chain here (the device computes them on the GPU box).
        g2 = np.random.default_rng(5)
        bonds, dims = [1, 3, 4, 3, 2, 1], [2, 3, 2, 3, 2]
        cores = [np.asfortranarray(g2.random((bonds[k], dims[k], bonds[k + 1])) - 0.4) for k in range(5)]
        It = np.stack([g2.integers(1, dims[k] + 1, 37) for k in range(2)], axis=1)
        Jt = np.stack([g2.integers(1, dims[2 + k] + 1, 21) for k in range(3)], axis=1)
        reft, _ = orc.Target.tt(cores).pi_eval(It.tolist(), Jt.tolist(), 0)

        def lenv(idx):  # evaluateleft, cachedtensortrain.jl:77-100
            v = np.ones((1, 1))
            for k2, sgm in enumerate(idx):
                v = v @ cores[k2][:, sgm - 1, :]
            return v[0]

        def renv(idx):  # evaluateright, :102-128
            v = np.ones((1, 1))
            for k2 in range(len(idx) - 1, -1, -1):
                v = cores[2 + k2][:, idx[k2] - 1, :] @ v
            return v[:, 0]

        D = bonds[2]
        cblk, cranges = P.column_blocks(len(Jt), world)
        clo, chi = cranges[rank]
        rfull = torch.zeros((cblk * world, 8), dtype=torch.float64)  # (columns, padded D): one environment per row
        for jj in range(clo, chi):
            rfull[rank * cblk + (jj - clo), :D] = torch.from_numpy(renv(Jt[jj]))
        _gather_blocks(rfull, cblk, rank)
        R = rfull[:len(Jt), :D].numpy().T  # D x nJ on every rank
        rb, rr = P.row_blocks(len(It), world)
        rlo2, rhi2 = rr[rank]
        acc = torch.zeros((len(It), len(Jt)), dtype=torch.float64)
        if rhi2 > rlo2:
            Lb = np.array([lenv(It[ii]) for ii in range(rlo2, rhi2)])
            acc[rlo2:rhi2] = torch.from_numpy(Lb @ R)
        dist.all_reduce(acc)  # stands in for the peer stores into the owner's Pi (disjoint row ranges)
        res["tt_rows_two_chains"] = bool(np.max(np.abs(acc.numpy() - reft)) <= 1e-13 * np.max(np.abs(reft)))
        # ---- the same with the library's PREFIX-AWARE partition (tci_shard_order): rows in lexicographic order of their
        # multi-indices, columns with the last site most significant, contiguous blocks of those orders per rank; the
        # right environments are gathered in sorted column order and a rank's block product is scattered to the
        # caller's rows / columns (k_scatter_block on the device)
        from tci_b200._lib import shard_order
        myrows = P.prefix_partition(It, world, 0)[rank]
        cperm = shard_order(Jt, 1)
        mycols = cperm[rank * cblk:(rank + 1) * cblk]
        rfull2 = torch.zeros((cblk * world, 8), dtype=torch.float64)
        for qq, jj in enumerate(mycols):
            rfull2[rank * cblk + qq, :D] = torch.from_numpy(renv(Jt[jj]))
        _gather_blocks(rfull2, cblk, rank)
        R2 = rfull2[:len(Jt), :D].numpy().T  # D x nJ, columns in suffix-sorted order
        acc2 = torch.zeros((len(It), len(Jt)), dtype=torch.float64)
        if len(myrows):
            Lb2 = np.array([lenv(It[ii]) for ii in myrows])
            blk2 = np.zeros((len(It), len(Jt)))
            blk2[np.ix_(myrows, cperm)] = Lb2 @ R2
            acc2 += torch.from_numpy(blk2)
        dist.all_reduce(acc2)
        res["tt_prefix_partition"] = bool(np.max(np.abs(acc2.numpy() - reft)) <= 1e-13 * np.max(np.abs(reft)))
        rows_all = np.concatenate(P.prefix_partition(It, world, 0))
        res["tt_prefix_partition_covers"] = sorted(rows_all.tolist()) == list(range(len(It)))
        # ---- sharded global search ----
        R = 10
        t = orc.Target.builtin(6, [R, 1], [2] * R)
        tci = orc.crossinterpolate2(t, [2] * R, [[1] * R, [1] + [2] * (R - 1)], tolerance=1e-4, maxbonddim=1,
                                    normalizeerror=False)
        starts = orc.start_points(7, 1, 12, [2] * R)  # (n, nsearch)
        piv_ref, err_ref = orc.globalsearch(t, tci.sitetensors, starts, abstol=1e-9, tolmargin=1.0, maxn=5)
        # contiguous blocks of the starts (tci_shard_range); every rank produces the per-start record (best error,
        # probe index in the star) of ITS starts, the records are all-gathered in place, and the library's selection
        # (tci_globalsearch_select) is replayed on them
        nsearch = starts.shape[1]
        blk, ranges = P.column_blocks(nsearch, world)
        lo, hi = ranges[rank]
        rec = torch.zeros((blk * world, 2), dtype=torch.float64)
        off = np.concatenate([[0], np.cumsum([2] * R)])
        for s in range(lo, hi):  # threshold -1: every start is accepted, so its best probe is reported
            pv, er = orc.globalsearch(t, tci.sitetensors, starts[:, [s]], abstol=-1.0, tolmargin=1.0, maxn=1)
            st = starts[:, s]
            diff = [k for k in range(R) if pv[0][k] != st[k]]
            p = diff[0] if diff else 0  # the star contains the start itself once per site: the first one wins
            rec[s, 0] = float(er[0])
            rec[s, 1] = float(off[p] + pv[0][p] - 1)
        _gather_blocks(rec, blk, rank)
        piv, es, sidx = P.select_global_pivots(rec[:nsearch, 0].numpy(), rec[:nsearch, 1].numpy().astype(np.int64),
                                               starts.T, [2] * R, 1e-9, 5)
        res["gs_points"] = piv.tolist() == piv_ref
        res["gs_errs"] = bool(np.array_equal(es, err_ref))
        q.put((rank, res))
    except Exception as e:  # report instead of leaving the parent waiting on the queue
        import traceback
        q.put((rank, {"exception: " + repr(e) + "\n" + traceback.format_exc(): False}))
        raise
    finally:
        dist.destroy_process_group()


def test_sharding_logic_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        rank, res = q.get(timeout=300)
        out[rank] = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        assert all(out[rank].values()), (rank, out[rank])


def test_column_blocks():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import tci_b200  # noqa: F401
    from tci_b200.parallel import column_blocks, select_global_pivots
    assert column_blocks(10, 4) == (3, [(0, 3), (3, 6), (6, 9), (9, 10)])
    assert column_blocks(2, 4) == (1, [(0, 1), (1, 2), (2, 2), (2, 2)])
    assert column_blocks(0, 2) == (0, [(0, 0), (0, 0)])
    from tci_b200.parallel import row_blocks
    assert row_blocks(100, 2) == (50, [(0, 50), (50, 100)])
    assert row_blocks(100, 2, align=16) == (64, [(0, 64), (64, 100)])
    assert row_blocks(10, 4) == (3, [(0, 3), (3, 6), (6, 9), (9, 10)])
    assert row_blocks(4096, 8) == (512, [(512 * r, 512 * (r + 1)) for r in range(8)])
    assert row_blocks(0, 2) == (0, [(0, 0), (0, 0)])
    starts = np.array([[1, 1], [2, 2], [1, 2], [2, 1]], dtype=np.int64)
    # records (best error, probe index in the star of localdims [2, 2]): start 1 has no finite probe (index -1)
    piv, es, sidx = select_global_pivots([0.3, 0.05, 0.2, 0.4], [3, -1, 0, 1], starts, [2, 2], 0.1, 2)
    assert piv.tolist() == [[1, 2], [1, 2]] and es.tolist() == [0.3, 0.2] and sidx.tolist() == [0, 2]
    piv, es, sidx = select_global_pivots([0.3, 0.05, 0.2, 0.4], [3, -1, 0, 1], starts, [2, 2], 0.25, 5)
    assert piv.tolist() == [[1, 2], [2, 1]] and sidx.tolist() == [0, 3]  # probe 1 of start (2, 1) is x_1 = 2


def test_prefix_aware_partition():
    """tci_shard_order (host only): rows are dealt to the GPUs in lexicographic order of their multi-indices, columns
    with the last site most significant; the order is stable.  For kronecker(Iset, d) -- i fastest, then sigma
    (tensorci2.jl:315-320) -- every GPU then holds ALL sigma of its parents, so no prefix environment is extended on
    two GPUs; the caller-order blocks would put one sigma of many parents on each GPU."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import tci_b200 as T
    from tci_b200._lib import shard_order
    from tci_b200.parallel import prefix_partition
    rng = np.random.default_rng(0)
    S = rng.integers(1, 4, (50, 5)).astype(np.int64)
    perm = shard_order(S, 0)
    rows = [tuple(r) for r in S[perm].tolist()]
    assert rows == sorted(rows) and sorted(perm.tolist()) == list(range(50))
    assert all(perm[q] < perm[q + 1] for q in range(49) if rows[q] == rows[q + 1])  # stable
    perm = shard_order(S, 1)
    cols = [tuple(reversed(r)) for r in S[perm].tolist()]
    assert cols == sorted(cols)
    parents = np.unique(rng.integers(1, 5, (64, 6)), axis=0).astype(np.int64)
    K = T.kronecker_left(parents, 4)  # (i, sigma), i fastest
    world = 8
    per_gpu_parents = [set(map(tuple, K[idx][:, :-1].tolist())) for idx in prefix_partition(K, world, 0)]
    blk = (len(K) + world - 1) // world
    naive = [set(map(tuple, K[r * blk:(r + 1) * blk][:, :-1].tolist())) for r in range(world)]
    # parents touched summed over GPUs: the sorted partition splits at most one parent per boundary
    assert sum(len(p) for p in per_gpu_parents) <= len(parents) + world - 1
    assert sum(len(p) for p in naive) >= 3 * len(parents)
    assert sorted(np.concatenate(prefix_partition(K, world, 0)).tolist()) == list(range(len(K)))
    KJ = T.kronecker_right(4, parents)  # (sigma, j), sigma fastest: suffix = j
    per_gpu = [set(map(tuple, KJ[idx][:, 1:].tolist())) for idx in prefix_partition(KJ, world, 1)]
    assert sum(len(p) for p in per_gpu) <= len(parents) + world - 1
    with pytest.raises(ValueError):
        shard_order(S, 2)
