"""Pins the CPU oracle against the fixtures of the reference's own test-suite.

Every test names the reference test (file:line under /root/reference/test) whose
data and assertion it re-states.  These run on CPU (no gpu marker).
"""
import itertools

import numpy as np
import pytest

from tests.golden import reference_fixtures as G

LORENTZ, SUM, Q2D, SEPCOS, TABLE, Q1D, GK = 1, 2, 3, 4, 5, 6, 7


# ---------------------------------------------------------------- argmax ----
def test_argmax_literal(oracle):  # test_matrixlu.jl:7-28
    A = G.ARGMAX_A
    # submatrixargmax(abs2, A, 1) == argmax(abs2.(A)) (column-major first maximum)
    B = A * A
    cm = np.argmax(B.flatten(order="F"))
    assert oracle.argmax_abs2(A, 1) == (cm % A.shape[0] + 1, cm // A.shape[0] + 1)
    k = min(A.shape)
    sub = B[k - 1:, k - 1:]
    cm = np.argmax(sub.flatten(order="F"))
    assert oracle.argmax_abs2(A, k) == (k + cm % sub.shape[0], k + cm // sub.shape[0])


def test_argmax_first_of_ties(oracle):  # matrixlu.jl:16-29 strict '>' keeps the first maximum
    A = np.ones((4, 5))
    assert oracle.argmax_abs2(A, 2) == (2, 2)
    A[3, 2] = -1.0
    A[1, 4] = np.nan
    assert oracle.argmax_abs2(A, 1) == (1, 1)


# ------------------------------------------------------------------ rrLU ----
def _check_lu_structure(lu, A):
    m, n = A.shape
    r = lu.npivot
    assert lu.L.shape == (m, r) and lu.U.shape == (r, n)
    assert np.all(lu.L == np.tril(lu.L)) and np.all(lu.U == np.triu(lu.U))
    left = np.zeros_like(lu.L)
    left[lu.rowpermutation - 1, :] = lu.L
    right = np.zeros_like(lu.U)
    right[:, lu.colpermutation - 1] = lu.U
    return left, right


def test_rrlu_4x4(oracle):  # test_matrixlu.jl:54-69
    A = G.RRLU_4x4
    lu = oracle.rrlu(A)
    left, right = _check_lu_structure(lu, A)
    assert np.all(np.diag(lu.L) == 1.0)
    np.testing.assert_allclose(left @ right, A, rtol=1.5e-8)
    assert sorted(lu.rowpermutation) == [1, 2, 3, 4] and sorted(lu.colpermutation) == [1, 2, 3, 4]


def test_rrlu_truncated(oracle):  # test_matrixlu.jl:88-97
    A = np.zeros((3, 3))
    A[0, 0] = 1.0
    assert oracle.rrlu(A).npivot == 1


def test_rrlu_maxrank_and_reltol(oracle):  # test_matrixlu.jl:99-138
    A = G.RRLU_8x6
    lu = oracle.rrlu(A, maxrank=4)
    assert lu.npivot == 4
    _check_lu_structure(lu, A)
    rng = np.random.default_rng(0)
    A2 = np.hstack([A, A + 1e-3 * rng.random((8, 6))])
    lu = oracle.rrlu(A2, reltol=1e-2)
    assert lu.npivot < 8
    left, right = _check_lu_structure(lu, A2)
    assert np.max(np.abs(left @ right - A2)) < 1e-2


def test_rrlu_exact_lowrank(oracle):  # test_matrixlu.jl:140-165
    A = G.LOWRANK_P @ G.LOWRANK_Q
    lu = oracle.rrlu(A)
    assert lu.npivot == 3
    left, right = _check_lu_structure(lu, A)
    np.testing.assert_allclose(left @ right, A, rtol=1.5e-8)


def test_lastpivoterror_fullrank(oracle):  # test_matrixlu.jl:167-175
    lu = oracle.rrlu(np.eye(2))
    assert lu.pivoterrors.tolist() == [1.0, 1.0, 0.0]
    assert lu.error == 0.0


def test_lastpivoterror_limited(oracle):  # test_matrixlu.jl:177-195
    A = G.RRLU_5x5
    lu = oracle.rrlu(A, maxrank=2)
    assert len(lu.pivoterrors) == 3 and lu.error > 0
    assert oracle.rrlu(A, abstol=0.5).error < 0.5
    assert oracle.rrlu(A, abstol=0.0).error == 0.0


def test_rrlu_tiny_values(oracle):  # test_matrixlu.jl:197-211
    A = G.RRLU_TINY
    lu = oracle.rrlu(A, abstol=1e-3)
    assert lu.npivot == 1 and lu.error > 0
    left, right = _check_lu_structure(lu, A)
    assert np.max(np.abs(left @ right - A)) < 1e-3


def test_rrlu_not_leftorthogonal(oracle):  # matrixlu.jl:117-123,170-174
    A = G.RRLU_5x5
    lu = oracle.rrlu(A, leftorthogonal=False)
    left, right = _check_lu_structure(lu, A)
    assert np.all(np.diag(lu.U) == 1.0)
    np.testing.assert_allclose(left @ right, A, rtol=1.5e-8)


def test_arrlu_4x4(oracle):  # test_matrixlu.jl:71-86 "Implementation of approximate rank-revealing LU"
    A = G.RRLU_4x4
    lu = oracle.arrlu(A, [1], [1])
    assert lu.L.shape == (4, lu.npivot) and lu.U.shape == (lu.npivot, 4)
    assert np.all(lu.L == np.tril(lu.L)) and np.all(np.diag(lu.L) == 1.0)
    assert np.all(lu.U == np.triu(lu.U))
    left = np.zeros_like(lu.L)
    left[lu.rowpermutation - 1, :] = lu.L
    right = np.zeros_like(lu.U)
    right[:, lu.colpermutation - 1] = lu.U
    np.testing.assert_allclose(left @ right, A, rtol=1.5e-8)


# ------------------------------------------------------------------ LUCI ----
@pytest.mark.parametrize("leftorth", [True, False])
def test_luci_matches_ci(oracle, leftorth):  # test_matrixluci.jl:6-38
    A = G.LUCI_8x6
    lc = oracle.luci(A, maxrank=4, leftorthogonal=leftorth)
    assert lc.npivot == 4
    I, J = lc.rowindices - 1, lc.colindices - 1
    P = A[np.ix_(I, J)]
    ci = A[:, J] @ np.linalg.solve(P, A[I, :])
    np.testing.assert_allclose(lc.left @ lc.right, ci, rtol=1.5e-8)
    if leftorth:  # left = colstimespivotinv == A[:, J] P^-1
        np.testing.assert_allclose(lc.left, A[:, J] @ np.linalg.inv(P), rtol=1e-8, atol=1e-12)
    else:  # right = pivotinvtimesrows == P^-1 A[I, :]
        np.testing.assert_allclose(lc.right, np.linalg.solve(P, A[I, :]), rtol=1e-8, atol=1e-12)


def test_luci_lowrank(oracle):  # test_matrixluci.jl:48-73
    A = G.LOWRANK_P @ G.LOWRANK_Q
    lc = oracle.luci(A)
    assert lc.npivot == 3
    np.testing.assert_allclose(lc.left @ lc.right, A, rtol=1.5e-8)
    P = A[np.ix_(lc.rowindices - 1, lc.colindices - 1)]
    assert np.linalg.cond(P) < 1e12


# ------------------------------------------------------------- batcheval ----
def test_batcheval_layout(oracle):  # test_batcheval.jl:13-35
    ld = [2, 2, 2, 2, 2]
    t = oracle.Target.builtin(SUM, [], ld)
    left = [[1, 1]] * 100
    right = [[1, 1]] * 100
    res, _ = t.pi_eval(left, right, 1)
    ref = np.array([[[sum(l) + c + sum(r) for r in right] for c in (1, 2)] for l in left], dtype=float)
    assert res.shape == (100, 2, 100)
    np.testing.assert_array_equal(res, ref)
    rng = np.random.default_rng(1)
    left = [[int(rng.integers(1, 3))] for _ in range(7)]
    right = [rng.integers(1, 3, 2).tolist() for _ in range(5)]
    res, _ = t.pi_eval(left, right, 2)
    ref = np.array([[[[sum(l) + c + cp + sum(r) for r in right] for cp in (1, 2)] for c in (1, 2)] for l in left],
                   dtype=float)
    np.testing.assert_array_equal(res, ref)


def test_batcheval_empty(oracle):  # batcheval.jl:40-42
    t = oracle.Target.builtin(SUM, [], [3, 3, 3, 3])
    res, _ = t.pi_eval([], [[1]], 1)
    assert res.size == 0


# ----------------------------------------------------------------- TCI2 -----
def test_pivoterrors_diag(oracle):  # test_tensorci2.jl:27-39
    table = np.diag(G.PIVOTERRORS_DIAGS).flatten(order="F")
    t = oracle.Target.builtin(TABLE, table, [3, 3])
    res = oracle.crossinterpolate2(t, [3, 3], [[1, 1]], tolerance=1e-8)
    assert res.pivoterrors.tolist() == G.PIVOTERRORS_DIAGS


def test_convergencecriterion_table(oracle):  # test_tensorci2.jl:504-554
    cc = oracle.convergencecriterion
    assert cc([1, 2], [1e-2, 1e-5], [0, 0], 1e-4, 4, 3) is False
    assert cc([1, 2, 2, 2], [1e-2, 1e-5, 1e-5, 1e-5], [0, 0, 0, 0], 1e-4, 4, 3) is True
    assert cc([1, 2, 2, 2], [1e-2, 1e-2, 1e-5, 1e-5], [0, 0, 0, 0], 1e-4, 4, 3) is False
    assert cc([1, 2, 2, 2], [1e-2] * 4, [0, 0, 0, 0], 1e-4, 2, 3) is True
    assert cc([1, 2, 2, 2], [1e-2] * 4, [0, 1, 1, 1], 1e-4, 2, 3) is True


def test_trivial_mps_exp(oracle):  # test_tensorci2.jl:55-102 (pivotsearch=:full)
    R = 8
    t = oracle.Target.builtin(Q1D, [R, 0], [2] * R)
    for nsearch in (0, 10):
        res = oracle.crossinterpolate2(t, [2] * R, [[1] * R, [1] + [2] * (R - 1)], tolerance=1e-4, maxbonddim=1,
                                       maxiter=2, normalizeerror=False, nsearchglobalpivot=nsearch,
                                       maxnglobalpivot=min(5, nsearch))
        assert res.linkdims == [1] * (R - 1)
        for x in (0.1, 0.3, 0.6, 0.9):
            q = int(x * 2**R)
            bits = [((q >> (R - 1 - b)) & 1) + 1 for b in range(R)]
            xq = q / 2**R
            assert abs(t(bits) - np.exp(-xq)) < 1e-14
            assert abs(res.evaluate(bits) - t(bits)) < 1e-4


def test_lorentz_5x10(oracle):  # test_tensorci2.jl:247-340 (ValueType Float64, :full)
    n = 5
    t = oracle.Target.builtin(LORENTZ, [1.0], [10] * n)
    res = oracle.crossinterpolate2(t, [10] * n, tolerance=1e-12, maxiter=200)
    assert max(res.bonderrors) <= 2e-12 * res.maxsamplevalue + 1e-300 or max(res.bonderrors) <= 2e-12
    assert max(res.linkdims) <= 200
    for v in itertools.product(range(1, 4), repeat=n):
        f = 1.0 / (1.0 + sum(x * x for x in v))
        assert t(v) == f
        assert abs(res.evaluate(v) - f) <= 1.5e-8 * abs(f)
    res8 = oracle.crossinterpolate2(t, [10] * n, tolerance=1e-8, maxiter=8, sweepstrategy="forward")
    assert max(res8.linkdims) >= 3


def test_lorentz_5x10_rook(oracle):  # test_tensorci2.jl:247-340 with pivotsearch = :rook
    n = 5
    t = oracle.Target.builtin(LORENTZ, [1.0], [10] * n)
    res = oracle.crossinterpolate2(t, [10] * n, tolerance=1e-12, maxiter=200, pivotsearch="rook")
    assert max(res.bonderrors) <= 2e-12 and max(res.linkdims) <= 200
    for v in itertools.product(range(1, 4), repeat=n):
        f = 1.0 / (1.0 + sum(x * x for x in v))
        assert abs(res.evaluate(v) - f) <= 1.5e-8 * abs(f)


def test_trivial_mps_exp_rook(oracle):  # test_tensorci2.jl:55-102 with pivotsearch = :rook
    R = 8
    t = oracle.Target.builtin(Q1D, [R, 0], [2] * R)
    res = oracle.crossinterpolate2(t, [2] * R, [[1] * R, [1] + [2] * (R - 1)], tolerance=1e-4, maxbonddim=1, maxiter=2,
                                   normalizeerror=False, nsearchglobalpivot=0, maxnglobalpivot=0, pivotsearch="rook")
    assert res.linkdims == [1] * (R - 1)
    for x in (0.1, 0.3, 0.6, 0.9):
        q = int(x * 2**R)
        bits = [((q >> (R - 1 - b)) & 1) + 1 for b in range(R)]
        assert abs(res.evaluate(bits) - t(bits)) < 1e-4


def test_lorentz_sum_docs_example(oracle):  # docs/src/index.md:15-43 (sum vs brute force)
    n = 5
    t = oracle.Target.builtin(LORENTZ, [1.0], [10] * n)
    res = oracle.crossinterpolate2(t, [10] * n, tolerance=1e-10)
    brute = sum(1.0 / (1.0 + sum(x * x for x in v)) for v in itertools.product(range(1, 11), repeat=n))
    assert abs(res.sum() - brute) <= 1e-8 * abs(brute)
    assert res.errors[-1] < 1e-10


def test_reinit_consistency(oracle):  # test_tensorci2.jl:461-475: Iset/Jset sizes are square after the run
    t = oracle.Target.builtin(LORENTZ, [1.0], [4] * 6)
    res = oracle.crossinterpolate2(t, [4] * 6, maxbonddim=5)
    for b in range(5):
        assert len(res.Iset[b + 1]) == len(res.Jset[b]) <= 5


def test_integration_10d_known_answer(oracle):  # test_integration.jl:61-70 ; integration.jl:20-58
    xgk = [0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
           0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
           0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
           0.207784955007898467600689403773245, 0.0]
    wgk = [0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
           0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
           0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
           0.204432940075298892414161999234649, 0.209482141084727828012999174891714]
    nodes = [-x for x in xgk] + xgk[-2::-1]
    weights = wgk + wgk[-2::-1]
    assert len(nodes) == 15 and abs(sum(weights) - 2.0) < 1e-14
    t = oracle.Target.builtin(GK, [15] + nodes + weights, [15] * 10)
    res = oracle.crossinterpolate2(t, [15] * 10, tolerance=1e-8, nsearchglobalpivot=10)
    integral = res.sum() / 15.0**10
    assert abs(integral - G.INTEGRAL_10D_REF) < 1e-3


# ------------------------------------------------------- TTCache / MPO ------
def _rand_tt(rng, bonds, dims):
    return [rng.random((bonds[i], dims[i], bonds[i + 1])) for i in range(len(dims))]


def _tt_full(cores):
    out = cores[0]
    for c in cores[1:]:
        out = np.tensordot(out, c, axes=([-1], [0]))
    return out[0, ..., 0]


def test_ttcache_all_splits(oracle):  # test_tensortrain.jl:113-140, test_cachedtensortrain.jl:8-65
    rng = np.random.default_rng(3)
    dims = [2, 3, 3, 2]
    cores = _rand_tt(rng, [1, 2, 3, 2, 1], dims)
    full = _tt_full(cores)
    t = oracle.Target.tt(cores)
    N = len(dims)
    for v in itertools.product(*[range(1, d + 1) for d in dims]):
        ref = full[tuple(x - 1 for x in v)]
        assert abs(t(v) - ref) < 1e-13
        assert abs(oracle.tt_evaluate(cores, v) - ref) < 1e-13
    assert abs(oracle.tt_sum(cores) - full.sum()) < 1e-12
    for nl in range(N + 1):
        for nr in range(N - nl + 1):
            M = N - nl - nr
            left = [list(v) for v in itertools.product(*[range(1, d + 1) for d in dims[:nl]])]
            right = [list(v) for v in itertools.product(*[range(1, d + 1) for d in dims[N - nr:]])]
            res, _ = t.pi_eval(left, right, M)
            ref = full.reshape((len(left), *dims[nl:nl + M], len(right)), order="C")
            # row index enumerates the left multi-indices in itertools order (last fastest) == C order
            np.testing.assert_allclose(res, ref, rtol=1e-12, atol=1e-14)


def test_tci_of_ttcache(oracle):  # test_tensorci2.jl:477-502
    rng = np.random.default_rng(4)
    dims = [2, 3, 3, 2]
    cores = _rand_tt(rng, [1, 2, 3, 2, 1], dims)
    full = _tt_full(cores)
    t = oracle.Target.tt(cores)
    res = oracle.crossinterpolate2(t, dims, tolerance=1e-10, maxbonddim=10)
    rec = _tt_full(res.sitetensors)
    np.testing.assert_allclose(rec, full, rtol=1.5e-8, atol=1e-12)


def _rand_mpo(rng, bonds, d1, d2):
    return [rng.random((bonds[i], d1[i], d2[i], bonds[i + 1])) - 0.5 for i in range(len(d1))]


def _mpo_dense(cores):
    out = cores[0]
    for c in cores[1:]:
        out = np.tensordot(out, c, axes=([-1], [0]))
    return out[0, ..., 0]  # (s1,s1',s2,s2',...)


def _contract_dense(A, B):
    N = len(A)
    a, b = _mpo_dense(A), _mpo_dense(B)
    # a[i1,j1,i2,j2,..] b[j1,k1,j2,k2,...] -> c[i1,k1,i2,k2,...]
    la = "".join(chr(ord("a") + 2 * s) + chr(ord("a") + 2 * s + 1) for s in range(N))
    lb = "".join(chr(ord("a") + 2 * s + 1) + chr(ord("A") + s) for s in range(N))
    lc = "".join(chr(ord("a") + 2 * s) + chr(ord("A") + s) for s in range(N))
    return np.einsum(f"{la},{lb}->{lc}", a, b)


def test_contraction_vs_dense(oracle):  # test_contraction.jl:68-146 (real-valued instance)
    rng = np.random.default_rng(5)
    N = 4
    d1, d2, d3 = [2, 2, 3, 2], [2, 3, 2, 2], [3, 2, 2, 2]
    A = _rand_mpo(rng, [1, 2, 3, 2, 1], d1, d2)
    B = _rand_mpo(rng, [1, 3, 2, 3, 1], d2, d3)
    dense = _contract_dense(A, B)
    t = oracle.Target.mpo_pair(A, B)
    ld = [d1[s] * d3[s] for s in range(N)]
    assert t.localdims == ld

    def unf(s, idx):
        return (idx - 1) % d1[s], (idx - 1) // d1[s]

    for v in itertools.product(*[range(1, d + 1) for d in ld]):
        key = tuple(x for s in range(N) for x in unf(s, v[s]))
        assert abs(t(v) - dense[key]) < 1e-13
    for nl, nr in ((1, 1), (0, 2), (2, 0), (1, 2), (2, 2), (0, 0)):
        M = N - nl - nr
        left = [list(v) for v in itertools.product(*[range(1, d + 1) for d in ld[:nl]])]
        right = [list(v) for v in itertools.product(*[range(1, d + 1) for d in ld[N - nr:]])]
        res, _ = t.pi_eval(left, right, M)
        for il, l in enumerate(left):
            for ir, r in enumerate(right):
                for c in itertools.product(*[range(1, d + 1) for d in ld[nl:nl + M]]):
                    v = list(l) + list(c) + list(r)
                    key = tuple(x for s in range(N) for x in unf(s, v[s]))
                    assert abs(res[(il, *[x - 1 for x in c], ir)] - dense[key]) < 1e-13
    res = oracle.crossinterpolate2(t, ld, tolerance=1e-12, maxbonddim=50)
    rec = _tt_full(res.sitetensors)
    for v in itertools.product(*[range(1, d + 1) for d in ld]):
        key = tuple(x for s in range(N) for x in unf(s, v[s]))
        assert abs(rec[tuple(x - 1 for x in v)] - dense[key]) < 1e-9


def test_globalsearch_errors(oracle):  # test_globalsearch.jl:7-36 (reported errors equal |f - tt|)
    R = 10
    t = oracle.Target.builtin(Q1D, [R, 1], [2] * R)
    res = oracle.crossinterpolate2(t, [2] * R, [[1] * R, [1] + [2] * (R - 1)], tolerance=1e-4, maxbonddim=1,
                                   normalizeerror=False)
    starts = oracle.start_points(7, 1, 20, [2] * R)
    piv, errs = oracle.globalsearch(t, res.sitetensors, starts, abstol=1e-9, tolmargin=1.0, maxn=20)
    assert len(piv) > 0
    for p, e in zip(piv, errs):
        assert abs(abs(t(p) - oracle.tt_evaluate(res.sitetensors, p)) - e) < 1e-15


def test_oracle_tt_to_tci2_conversion(oracle):  # test_conversion.jl:76-92 structure (real-valued)
    rng = np.random.default_rng(5)
    bonds, dims = [1, 3, 5, 3, 1], [4, 4, 4, 4]
    cores = [np.asfortranarray(rng.uniform(-1, 1, (bonds[i], dims[i], bonds[i + 1]))) for i in range(4)]

    def dense(cs):
        t = cs[0]
        for c in cs[1:]:
            t = np.tensordot(t, c, axes=([-1], [0]))
        return t.reshape(t.shape[1:-1])

    I, J, out, pe, mx = oracle.tensorci2_from_tt(cores, tolerance=1e-14)
    assert [len(i) for i in I[1:]] == bonds[1:-1] == [len(j) for j in J[:-1]]
    assert np.abs(dense(out) - dense(cores)).max() < 1e-13
    assert pe.shape == (max(bonds) + 1,) and pe[-1] == 0.0
    # the pivots are interpolation points of the converted train: T[I_{b+1}, J_b] is invertible
    D = dense(cores)
    for b in range(3):
        P = np.array([[D[tuple(np.concatenate([i, j]) - 1)] for j in J[b]] for i in I[b + 1]])
        assert np.linalg.matrix_rank(P) == bonds[b + 1]


def test_projected_batchevaluate(oracle):
    """test_contraction.jl:101-139 ("Contraction, batchevaluate", real-valued): the projected result is the slice of
    the unprojected one; the same for a TTCache with two-index sites (cachedtensortrain.jl:170-215).  The projected
    restatement slices the cores BEFORE contracting, as the reference does."""
    rng = np.random.default_rng(3)
    N = 4
    bd = [1, 2, 3, 2, 1]
    A = [np.asfortranarray(rng.random((bd[n], 2, 3, bd[n + 1]))) for n in range(N)]
    B = [np.asfortranarray(rng.random((bd[n], 3, 2, bd[n + 1]))) for n in range(N)]
    ab = oracle.Target.mpo_pair(A, B)
    ref, _ = ab.pi_eval([[1]], [[1]], 2)
    mi = ref.reshape((1, 2, 2, 2, 2, 1), order="F")
    cases = [([[0, 0], [1, 0]], mi[:, :, :, 0, :, :]), ([[0, 0], [1, 1]], mi[:, :, :, 0, 0, :]),
             ([[0, 1], [1, 0]], mi[:, :, 0, 0, :, :])]
    for proj, sl in cases:
        res = oracle.mpo_batchevaluate_projected(A, B, [[1]], [[1]], 2, proj)
        assert res.ndim == 4
        np.testing.assert_allclose(res.flatten(order="F"), sl.flatten(order="F"), rtol=1e-13)
    np.testing.assert_allclose(oracle.mpo_batchevaluate_projected(A, B, [[1]], [[1]], 2), ref, rtol=1e-13)
    with pytest.raises(oracle.OracleError, match="Length mismatch"):
        oracle.mpo_batchevaluate_projected(A, B, [[1]], [[1]], 2, [[0, 0]])
    cores = [np.asfortranarray(rng.random((b0, 6, b1)) - 0.4) for b0, b1 in ((1, 3), (3, 4), (4, 2), (2, 1))]
    I, J = [[1], [3], [6]], [[2], [5]]
    full, _ = oracle.Target.tt(cores).pi_eval(I, J, 2)
    mi = full.reshape((3, 2, 3, 2, 3, 2), order="F")
    res = oracle.tt_batchevaluate_projected(cores, [[2, 3]] * 4, I, J, 2, [[0, 2], [1, 0]])
    assert res.shape == (3, 2, 3, 2)
    np.testing.assert_allclose(res, mi[:, :, 1, 0, :, :], rtol=1e-13)
    with pytest.raises(oracle.OracleError, match="Invalid parameter M"):
        oracle.tt_batchevaluate_projected(cores, [[2, 3]] * 4, I, J, 1)


def test_complex_groundwork_argmax_and_rrlu(oracle):
    """ComplexF64 groundwork for SURVEY 8f-4 (no product path yet).  The reference's complex arg-max literal
    (test_matrixlu.jl:39-52) and the real-valued rrLU testsets re-used with complex data: L*U reproduces the permuted
    matrix, the first pivot is the largest element, rank-3 data stops at 3 pivots (test_matrixlu.jl:142-165 structure), and on
    real input the complex restatement picks exactly the pivots of the Float64 oracle."""
    A = np.array([[0, 1, 2, 3, 4, 5],
                  [1, 1 + 1j, 2 + 1j, 3 + 1j, 4 + 1j, 5 + 1j],
                  [1, 1 + 2j, 2 + 2j, 3 + 2j, 4 + 2j, 5 + 2j]], dtype=np.complex128)
    am = oracle.submatrixargmax_abs2_complex
    assert am(A, [3], [5]) == (3, 5)
    ref = np.unravel_index(np.argmax((np.abs(A) ** 2).flatten(order="F")), A.shape, order="F")  # Julia argmax: column-major
    assert am(A) == (ref[0] + 1, ref[1] + 1)
    assert am(A, [1], None) == (1, int(np.argmax(np.abs(A[0]) ** 2)) + 1)
    assert am(A, None, [1]) == (int(np.argmax(np.abs(A[:, 0]) ** 2)) + 1, 1)
    with pytest.raises(oracle.OracleError, match="rows must not be empty"):
        am(A, [], [1])
    rng = np.random.default_rng(12)
    for lo in (True, False):
        B = rng.standard_normal((9, 7)) + 1j * rng.standard_normal((9, 7))
        rp, cp, L, U, r, err = oracle.rrlu_complex(B, leftorthogonal=lo)
        assert r == 7 and err == 0.0
        np.testing.assert_allclose(L @ U, B[rp - 1][:, cp - 1], rtol=1e-13, atol=1e-13)
        assert abs((U if lo else L)[0, 0]) == np.max(np.abs(B))  # the first pivot is the largest element
        p = rng.standard_normal((10, 3)) + 1j * rng.standard_normal((10, 3))
        q = rng.standard_normal((3, 12)) + 1j * rng.standard_normal((3, 12))
        rp, cp, L, U, r, err = oracle.rrlu_complex(p @ q, reltol=1e-10, leftorthogonal=lo)
        assert r == 3 and err < 1e-10 * np.max(np.abs(p @ q)) * 100
        np.testing.assert_allclose(L @ U, (p @ q)[rp - 1][:, cp - 1], rtol=1e-10, atol=1e-12)
        R = rng.standard_normal((8, 6))
        ref = oracle.rrlu(R, leftorthogonal=lo)
        rp, cp, L, U, r, err = oracle.rrlu_complex(R, leftorthogonal=lo)
        assert r == ref.npivot and np.array_equal(rp, ref.rowpermutation) and np.array_equal(cp, ref.colpermutation)
        np.testing.assert_allclose(L.real, ref.L, rtol=1e-13, atol=1e-14)
        assert np.max(np.abs(L.imag)) == 0.0


def test_contraction_blas_restatement_matches_oracle(oracle):
    """oracle.ContractionBLAS (numpy / OpenBLAS restatement of contraction.jl:71-176, 236-335 for M = 0: the CPU
    baseline of the contraction stage) against the C++ oracle's batchevaluate."""
    rng = np.random.default_rng(3)
    ns, D = 7, 5
    bonds = [1] + [D] * (ns - 1) + [1]
    A = [np.asfortranarray(rng.random((bonds[i], 2, 3, bonds[i + 1])) - 0.5) for i in range(ns)]
    B = [np.asfortranarray(rng.random((bonds[i], 3, 2, bonds[i + 1])) - 0.5) for i in range(ns)]
    I = np.stack([rng.integers(1, 5, 9) for _ in range(3)], axis=1)
    J = np.stack([rng.integers(1, 5, 11) for _ in range(4)], axis=1)
    ref, _ = oracle.Target.mpo_pair(A, B).pi_eval(I.tolist(), J.tolist(), 0, 0.0)
    got = oracle.ContractionBLAS(A, B).batchevaluate0(I.tolist(), J.tolist())
    assert got.shape == (9, 11)
    assert np.max(np.abs(got - np.asarray(ref).reshape((9, 11), order="F"))) <= 1e-13 * np.max(np.abs(got))
