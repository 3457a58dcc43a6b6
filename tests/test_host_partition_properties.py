"""Property tests (hypothesis, CPU only) of the host-only pieces of the sharded stages exported by the C ABI:
tci_shard_range (contiguous, aligned blocks that cover [0, n)) and tci_shard_order (a stable permutation that sorts rows
by prefix / columns by suffix), and of the index-set helpers of the host mirror against their definitions in the
reference (kronecker, tensorci2.jl:315-327; Base.union keeps first occurrences in order)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import tci_b200  # noqa: F401
from tci_b200._lib import shard_order, shard_range
from tci_b200.util import kronecker_left, kronecker_right, union


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 5000), world=st.integers(1, 16), align=st.sampled_from([1, 2, 16, 64]))
def test_shard_range_covers_and_aligns(n, world, align):
    blocks = [shard_range(n, world, r, align) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for (lo, hi), (lo2, _) in zip(blocks, blocks[1:]):
        assert lo <= hi == lo2
    for lo, hi in blocks:
        assert lo % align == 0 or lo == n
    sizes = [hi - lo for lo, hi in blocks if hi > lo]
    assert len(set(sizes[:-1])) <= 1 and (not sizes or sizes[-1] <= sizes[0])  # equal blocks, a shorter tail at most


@settings(max_examples=150, deadline=None)
@given(data=st.data())
def test_shard_order_is_a_stable_sort(data):
    ln = data.draw(st.integers(1, 6))
    cnt = data.draw(st.integers(0, 60))
    d = data.draw(st.integers(1, 4))
    S = np.array(data.draw(st.lists(st.lists(st.integers(1, d), min_size=ln, max_size=ln), min_size=cnt, max_size=cnt)),
                 dtype=np.int64).reshape(cnt, ln)
    for side, key in ((0, lambda r: tuple(r)), (1, lambda r: tuple(reversed(r)))):
        perm = shard_order(S, side)
        assert sorted(perm.tolist()) == list(range(cnt))
        keys = [key(S[q].tolist()) for q in perm]
        assert keys == sorted(keys)
        assert all(perm[q] < perm[q + 1] for q in range(cnt - 1) if keys[q] == keys[q + 1])


@settings(max_examples=100, deadline=None)
@given(data=st.data())
def test_kronecker_and_union_definitions(data):
    ln = data.draw(st.integers(0, 4))
    cnt = data.draw(st.integers(1, 12))
    d = data.draw(st.integers(1, 5))
    S = np.array(data.draw(st.lists(st.lists(st.integers(1, 3), min_size=ln, max_size=ln), min_size=cnt, max_size=cnt)),
                 dtype=np.int64).reshape(cnt, ln)
    KL = kronecker_left(S, d)  # [[is..., j] for is in Iset, j in 1:d][:]  -> is fastest
    assert KL.tolist() == [S[i].tolist() + [j] for j in range(1, d + 1) for i in range(cnt)]
    KR = kronecker_right(d, S)  # [[i, js...] for i in 1:d, js in Jset][:]  -> i fastest
    assert KR.tolist() == [[i] + S[j].tolist() for j in range(cnt) for i in range(1, d + 1)]
    if ln:
        A = np.unique(S, axis=0)
        B = np.array(data.draw(st.lists(st.lists(st.integers(1, 3), min_size=ln, max_size=ln), min_size=0, max_size=10)),
                     dtype=np.int64).reshape(-1, ln)
        seen, ref = set(map(tuple, A.tolist())), [tuple(r) for r in A.tolist()]
        for r in map(tuple, B.tolist()):
            if r not in seen:
                seen.add(r)
                ref.append(r)
        assert [tuple(r) for r in union(A, B).tolist()] == ref
