"""CachedFunction as a device-resident memo (cachedfunction.jl, test_cachedfunction.jl restated).  The
values are those of the wrapped target bit for bit (a memo of a pure function), so the parity bar is array_equal."""
import numpy as np
import pytest


def test_cachekey_encode_decode_host():
    """"encode and decode cachekey" / "computekey boundary check" (test_cachedfunction.jl:137-160) on the host-side key
    arithmetic, which needs no GPU: key(x) = sum coeffs[n] (x_n - 1), decode inverts it."""
    import tci_b200 as T
    cf = T.CachedFunction.__new__(T.CachedFunction)
    cf.localdims, cf.coeffs = [2, 3, 4], [1, 2, 6]
    seen = set()
    for i1 in range(1, 3):
        for i2 in range(1, 4):
            for i3 in range(1, 5):
                k = cf.encodecachekey([i1, i2, i3])
                assert cf.decodecachekey(k) == [i1, i2, i3]
                seen.add(k)
    assert seen == set(range(24))
    with pytest.raises(RuntimeError, match="Invalid length of indexset"):
        cf._key([1] * 6)


@pytest.fixture(scope="module")
def T():
    import tci_b200
    return tci_b200


@pytest.mark.gpu
def test_cache_pointwise(T):  # "cache" testset, test_cachedfunction.jl:48-59
    f = T.BuiltinTarget(T.SUM, [], [4, 2])
    cf = T.CachedFunction(f, [4, 2], capacity_log2=8)
    assert cf.f is f
    n = 0
    for i in range(1, 5):
        for j in range(1, 3):
            x = [i, j]
            assert cf(x) == f(x)
            n += 1
            st = cf.stats()
            assert st["entries"] == n and st["misses"] == n  # TCI._key(cf, x) in keys(cf.cache)
            assert cf(x) == f(x)  # second access: answered from the memo
            assert cf.stats()["hits"] == n and cf.stats()["entries"] == n


@pytest.mark.gpu
def test_cache_batcheval(T, oracle):  # "cache(batcheval)" testset, :82-93: 100 identical left / right index sets
    ld = [2, 2, 2, 2, 2]
    f = T.BuiltinTarget(T.SUM, [], ld)
    cf = T.CachedFunction(f, ld, capacity_log2=10)
    left = [[1, 1]] * 100
    right = [[1, 1]] * 100
    res = cf(left, right, 1)
    ref = np.array([[[sum(l) + c + sum(r) for r in right] for c in (1, 2)] for l in left], dtype=np.float64)
    assert res.shape == (100, 2, 100) and np.array_equal(res, ref)
    st = cf.stats()
    assert st["entries"] == 2 and st["unstored"] == 0  # two distinct points, inserted once each
    assert np.array_equal(cf(left, right, 1), ref) and cf.stats()["misses"] == st["misses"]  # all hits now


@pytest.mark.gpu
def test_cache_many_keys_and_all_splits(T, oracle):
    """"key collision" testset (:117-135) in spirit: many distinct keys -> as many entries; and every (nl, M, nr) split of
    a cached analytic target returns the wrapped target's values bit for bit, before and after they are memoised."""
    rng = np.random.default_rng(5)
    ld = [2] * 36
    f = T.BuiltinTarget(T.LORENTZ, [1.0], ld)
    cf = T.CachedFunction(f, capacity_log2=18)
    pts = np.unique(np.stack([rng.integers(1, 3, 50000) for _ in ld], axis=1), axis=0)
    v1 = cf.evaluate_points(pts)
    assert np.array_equal(v1, f.evaluate_points(pts))
    st = cf.stats()
    assert st["entries"] + st["unstored"] == len(pts) and st["unstored"] < len(pts) // 100
    assert np.array_equal(cf.evaluate_points(pts), v1)
    assert cf.stats()["hits"] >= len(pts) - st["unstored"]
    ld = [3, 4, 2, 5, 3]
    g = T.BuiltinTarget(T.LORENTZ, [0.5], ld)
    cg = T.CachedFunction(g, capacity_log2=12)
    for rep in range(2):
        for nl in range(6):
            for nr in range(6 - nl):
                M = 5 - nl - nr
                I = np.stack([rng.integers(1, d + 1, 7 if nl else 1) for d in ld[:nl]], axis=1) if nl else np.zeros((1, 0), dtype=np.int64)
                J = np.stack([rng.integers(1, d + 1, 5 if nr else 1) for d in ld[5 - nr:]], axis=1) if nr else np.zeros((1, 0), dtype=np.int64)
                assert np.array_equal(cg(I, J, M), g(I, J, M))
    assert cg.stats()["entries"] <= int(np.prod(ld))
    with pytest.raises(RuntimeError, match="Overflow in CachedFunction"):
        T.CachedFunction(T.BuiltinTarget(T.SUM, [], [4] * 70))  # "many bits" (:95-104) needs keys wider than UInt128


@pytest.mark.gpu
def test_crossinterpolate2_through_cache(T):
    """crossinterpolate2 on CachedFunction(f): identical pivots, ranks and site tensors, and the wrapped target is
    evaluated once per distinct point (fewer evaluations than the uncached run, which re-evaluates Pi every sweep)."""
    ld = [10] * 6
    f1 = T.BuiltinTarget(T.LORENTZ, [1.0], ld)
    f2 = T.BuiltinTarget(T.LORENTZ, [1.0], ld)
    cf = T.CachedFunction(f2, capacity_log2=20)
    a, ra, ea = T.crossinterpolate2(f1, ld, tolerance=1e-8, rng=T.CounterRNG(1))
    b, rb, eb = T.crossinterpolate2(cf, ld, tolerance=1e-8, rng=T.CounterRNG(1))
    assert ra == rb and ea == eb
    for s in range(len(ld)):
        assert np.array_equal(a.Iset[s], b.Iset[s]) and np.array_equal(a.Jset[s], b.Jset[s])
        assert np.array_equal(a.sitetensors[s], b.sitetensors[s])
    st = cf.stats()
    assert st["unstored"] == 0 and st["hits"] > 0
    assert f2.nevals == 0  # the wrapped target is only reached through the library
    assert st["misses"] < f1.nevals  # distinct points vs every requested element
