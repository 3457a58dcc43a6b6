"""CPU tests of the host-side mirror: index-set conventions, sweep strategy, convergence
criterion and the injected start-point generator (shared with the oracle)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def T():
    import __graft_entry__ as g
    g.build()
    import tci_b200
    return tci_b200


def test_kronecker_order(T):  # test_tensorci2.jl:9-25 + tensorci2.jl:315-327
    I = np.array([[1, 2], [3, 4], [5, 6]])
    k = T.kronecker_left(I, 2)
    assert k.tolist() == [[1, 2, 1], [3, 4, 1], [5, 6, 1], [1, 2, 2], [3, 4, 2], [5, 6, 2]]
    k = T.kronecker_right(2, I)
    assert k.tolist() == [[1, 1, 2], [2, 1, 2], [1, 3, 4], [2, 3, 4], [1, 5, 6], [2, 5, 6]]
    multiset = np.tile(np.arange(1, 6), (5, 1))
    c = T.kronecker_left(multiset, 4)
    assert all(ci[:5].tolist() == [1, 2, 3, 4, 5] and 1 <= ci[5] <= 4 for ci in c)


def test_union_and_pushunique(T):
    from tci_b200.util import pushunique, union
    a = np.array([[1, 1], [2, 1], [1, 2]])
    b = np.array([[2, 1], [3, 3], [1, 1], [3, 3]])
    assert union(a, b).tolist() == [[1, 1], [2, 1], [1, 2], [3, 3]]
    assert union(a, np.zeros((0, 2), dtype=np.int64)).tolist() == a.tolist()
    e = np.zeros((0, 0), dtype=np.int64)
    e = pushunique(e, [])
    assert e.shape == (1, 0)
    assert pushunique(e, []).shape == (1, 0)
    s = pushunique(np.zeros((0, 2), dtype=np.int64), [1, 2])
    s = pushunique(s, [1, 2])
    s = pushunique(s, [2, 2])
    assert s.tolist() == [[1, 2], [2, 2]]


def test_forwardsweep(T):  # test_sweepstrategies.jl:4-9
    assert T.forwardsweep("forward", 1) and T.forwardsweep("forward", 2)
    assert not T.forwardsweep("backward", 1) and not T.forwardsweep("backward", 2)
    assert T.forwardsweep("backandforth", 1) and not T.forwardsweep("backandforth", 2)


def test_convergencecriterion(T):  # test_tensorci2.jl:504-554
    cc = T.convergencecriterion
    assert cc([1, 2], [1e-2, 1e-5], [0, 0], 1e-4, 4, 3) is False
    assert cc([1, 2, 2, 2], [1e-2, 1e-5, 1e-5, 1e-5], [0, 0, 0, 0], 1e-4, 4, 3) is True
    assert cc([1, 2, 2, 2], [1e-2, 1e-2, 1e-5, 1e-5], [0, 0, 0, 0], 1e-4, 4, 3) is False
    assert cc([1, 2, 2, 2], [1e-2] * 4, [0, 0, 0, 0], 1e-4, 2, 3) is True
    assert cc([1, 2, 2, 2], [1e-2] * 4, [0, 1, 1, 1], 1e-4, 2, 3) is True


def test_counter_rng_matches_oracle(T, oracle):
    ld = [10, 3, 64, 2, 7]
    rng = T.CounterRNG(5)
    for it in (1, 2, 3):
        got = rng.start_points(6, ld)
        ref = oracle.start_points(5, it, 6, ld)
        assert np.array_equal(got, ref.T)
    assert got.min() >= 1 and all(got[:, p].max() <= ld[p] for p in range(5))


def test_apply_projector(T):  # projector_to_slice util.jl:124-126; cachedtensortrain.jl:198-205; contraction.jl:290-302
    from tci_b200.batcheval import apply_projector
    rng = np.random.default_rng(0)
    full = np.asfortranarray(rng.random((3, 2, 3, 2, 2, 4)))  # (nI, [2,3], [2,2], nJ)
    res = full.reshape((3, 6, 4, 4), order="F")
    out = apply_projector(res, [[2, 3], [2, 2]], [[0, 0], [1, 0]])
    assert out.shape == (3, 6, 2, 4)
    assert np.array_equal(out, full[:, :, :, 0, :, :].reshape((3, 6, 2, 4), order="F"))
    out = apply_projector(res, [[2, 3], [2, 2]], [[0, 2], [1, 2]])
    assert out.shape == (3, 2, 1, 4) and np.array_equal(out[:, :, 0, :], full[:, :, 1, 0, 1, :])
    out = apply_projector(res, [[2, 3], [2, 2]], [[0, 0], [0, 0]])
    assert np.array_equal(out, res)
