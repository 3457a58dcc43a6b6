"""GPU parity tests: the CUDA path, called through the C ABI (ctypes binding of
include/tci_b200.h), against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): pivot index sets / permutations / ranks bit-identical;
rrLU factors bit-identical in exact mode (same operation order, no FMA); Pi tensors of the
built-in analytic targets bit-identical; everything that goes through a GEMM/TRSM (LUCI
left/right, TT and MPO targets, site tensors) within 1e-10 relative.
"""
import itertools

import numpy as np
import pytest

from tests.golden import reference_fixtures as G

pytestmark = pytest.mark.gpu

LORENTZ, SUM, Q2D, SEPCOS, TABLE, Q1D, GK = 1, 2, 3, 4, 5, 6, 7
RTOL = 1e-10


@pytest.fixture(scope="module")
def T():
    import tci_b200
    return tci_b200


@pytest.fixture(scope="module")
def ctx(T):
    return T.default_context()


def sepcos_params(n, seed=4, nterms=4):
    """BASELINE config 4 parameters: dyadic so that all partial sums are exact."""
    rng = np.random.default_rng(seed)
    w = rng.integers(1, 1025, n) / 256.0
    a = rng.integers(-512, 513, nterms) / 1024.0
    om = rng.integers(-1024, 1025, (nterms, n)) / 32.0
    return np.concatenate([[nterms], w, a, om.flatten()])


def gk_params():
    xgk = [0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
           0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
           0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
           0.207784955007898467600689403773245, 0.0]
    wgk = [0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
           0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
           0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
           0.204432940075298892414161999234649, 0.209482141084727828012999174891714]
    return [15] + [-x for x in xgk] + xgk[-2::-1] + wgk + wgk[-2::-1]


TARGETS = [
    ("lorentz", LORENTZ, [1.0], [10] * 6),
    ("sum", SUM, [], [2, 3, 4, 3, 2]),
    ("q2d_fused", Q2D, [0, 8], [4] * 8),
    ("q2d_interleaved", Q2D, [1, 6], [2] * 12),
    ("sepcos", SEPCOS, None, [64] * 5),
    ("table", TABLE, "table", [3, 4, 2, 3]),
    ("q1d", Q1D, [10, 1], [2] * 10),
    ("gk", GK, gk_params(), [15] * 5),
]


def make_target(T, oracle, kind, params, ld):
    if params is None:
        params = sepcos_params(len(ld))
    elif isinstance(params, str):
        params = np.random.default_rng(11).standard_normal(int(np.prod(ld)))
    return T.BuiltinTarget(kind, params, ld), oracle.Target.builtin(kind, params, ld)


def rand_indexset(rng, dims, count):
    if len(dims) == 0:
        return np.zeros((1, 0), dtype=np.int64)
    return np.stack([rng.integers(1, d + 1, count) for d in dims], axis=1).astype(np.int64)


# ------------------------------------------------------------------ K1 ------
@pytest.mark.parametrize("name,kind,params,ld", TARGETS, ids=[t[0] for t in TARGETS])
def test_pi_eval_bit_exact(T, oracle, name, kind, params, ld):
    f, o = make_target(T, oracle, kind, params, ld)
    rng = np.random.default_rng(5)
    n = len(ld)
    for nl, M in [(1, 0), (2, 0), (n - 1, 0), (1, 1), (2, 1), (0, 1), (1, 2), (n - 2, 2), (0, 2), (n, 0), (0, 0)]:
        nr = n - nl - M
        if nr < 0:
            continue
        I = rand_indexset(rng, ld[:nl], 37 if nl else 1)
        J = rand_indexset(rng, ld[n - nr:], 29 if nr else 1)
        got, dev, mx = f._pi(I, J, M, True, True)
        ref, omx = o.pi_eval(I.tolist(), J.tolist(), M, 0.0)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), f"{name} nl={nl} M={M}"
        assert mx == omx
        assert np.array_equal(dev.to_host().reshape(ref.shape, order="F"), ref)
    pts = rand_indexset(rng, ld, 50)
    assert np.array_equal(f.evaluate_points(pts), np.array([o(p) for p in pts]))


def test_pi_eval_layout_and_empty(T):  # test_batcheval.jl:13-46
    f = T.BuiltinTarget(SUM, [], [2, 2, 2, 2, 2])
    left = [[1, 1]] * 100
    right = [[1, 1]] * 100
    res = f(left, right, 1)
    ref = np.array([[[sum(l) + c + sum(r) for r in right] for c in (1, 2)] for l in left], dtype=float)
    assert res.shape == (100, 2, 100) and np.array_equal(res, ref)
    g = T.BuiltinTarget(SUM, [], [3, 3, 3])
    assert g([[1], [2]], [[1], [2]], 1).shape == (2, 3, 2)
    assert g([], [[1]], 1).size == 0
    with pytest.raises(RuntimeError, match="Invalid number of central indices"):
        g([[1]], [[1]], 0)


def test_pi_eval_large_odd_shapes(T, oracle):
    ld = [10] * 8
    f, o = make_target(T, oracle, LORENTZ, [1.0], ld)
    rng = np.random.default_rng(6)
    I = rand_indexset(rng, ld[:3], 1031)
    J = rand_indexset(rng, ld[5:], 517)
    got = f(I, J, 2)
    ref, _ = o.pi_eval(I.tolist(), J.tolist(), 2)
    assert np.array_equal(got, ref)


# ------------------------------------------------------------------ K2 ------
def assert_lu_equal(lu, ref, bitexact=True):
    assert lu.npivot == ref.npivot
    assert np.array_equal(lu.rowpermutation, ref.rowpermutation)
    assert np.array_equal(lu.colpermutation, ref.colpermutation)
    if bitexact:
        assert lu.error == ref.error or (np.isnan(lu.error) and np.isnan(ref.error))
        assert np.array_equal(lu.L, ref.L)
        assert np.array_equal(lu.U, ref.U)
        assert np.array_equal(lu._pivoterrors, ref.pivoterrors)
    else:
        np.testing.assert_allclose(lu.L, ref.L, rtol=RTOL, atol=1e-300)
        np.testing.assert_allclose(lu.U, ref.U, rtol=RTOL, atol=1e-300)


FIXTURES = [
    ("4x4", G.RRLU_4x4, {}),
    ("8x6_maxrank4", G.RRLU_8x6, {"maxrank": 4}),
    ("lowrank", G.LOWRANK_P @ G.LOWRANK_Q, {}),
    ("eye2", np.eye(2), {}),
    ("5x5_maxrank2", G.RRLU_5x5, {"maxrank": 2}),
    ("5x5_abstol", G.RRLU_5x5, {"abstol": 0.5}),
    ("5x5_full", G.RRLU_5x5, {"abstol": 0.0}),
    ("tiny", G.RRLU_TINY, {"abstol": 1e-3}),
    ("unit", np.diag([1.0, 0.0, 0.0]), {}),
    ("argmaxA", G.ARGMAX_A, {}),
]


@pytest.mark.parametrize("leftorth", [True, False])
@pytest.mark.parametrize("name,A,kw", FIXTURES, ids=[f[0] for f in FIXTURES])
def test_rrlu_reference_fixtures(T, oracle, name, A, kw, leftorth):  # test_matrixlu.jl:54-211
    lu = T.rrlu(A, leftorthogonal=leftorth, **kw)
    ref = oracle.rrlu(A, leftorthogonal=leftorth, **kw)
    assert_lu_equal(lu, ref)
    if name == "eye2":
        assert T.pivoterrors(lu).tolist() == [1.0, 1.0, 0.0] and T.lastpivoterror(lu) == 0.0
    if name == "lowrank":
        assert T.npivots(lu) == 3
        np.testing.assert_allclose(T.left(lu) @ T.right(lu), A, rtol=1.5e-8)
    if name == "unit":
        assert lu.npivot == 1


def lowrank_matrix(m, n, r, seed, decay=40.0):
    """BASELINE config 2: A = sum_k s_k p_k q_k^T, p,q ~ U(0,1), s_k = 2^(-decay k / r)."""
    rng = np.random.default_rng(seed)
    p = rng.random((m, r))
    q = rng.random((r, n))
    s = 2.0 ** (-decay * np.arange(1, r + 1) / r)
    return (p * s) @ q


SHAPES = [(1, 1, 1), (1, 7, 1), (9, 1, 1), (2, 3, 2), (17, 33, 9), (64, 64, 20), (100, 37, 30), (37, 100, 30),
          (130, 257, 40), (300, 300, 64), (513, 700, 48), (1000, 600, 32)]


@pytest.mark.parametrize("leftorth", [True, False])
@pytest.mark.parametrize("m,n,r", SHAPES)
def test_rrlu_random_lowrank_bit_exact(T, oracle, m, n, r, leftorth):
    A = lowrank_matrix(m, n, r, seed=m * 1000 + n)
    lu = T.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=leftorth)
    ref = oracle.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=leftorth)
    assert_lu_equal(lu, ref)


@pytest.mark.parametrize("leftorth", [True, False])
@pytest.mark.parametrize("m,n,r", [(2100, 2050, 24), (3000, 500, 16)])
def test_rrlu_streaming_regime(T, oracle, m, n, r, leftorth):
    """Too large for the shared-memory resident mode: columns are streamed from L2 / HBM."""
    A = lowrank_matrix(m, n, r, seed=m + n)
    assert_lu_equal(T.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=leftorth),
                    oracle.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=leftorth))


def test_rrlu_pivot_column_from_l2(T, oracle):
    """m > 24576: the pivot column no longer fits shared memory and is read from L2 (kernel mode 3)."""
    A = lowrank_matrix(25000, 300, 12, seed=77)
    assert_lu_equal(T.rrlu(A, maxrank=12, reltol=1e-12), oracle.rrlu(A, maxrank=12, reltol=1e-12))


def test_rrlu_config2_full_size_vs_oracle(T, oracle):
    """BASELINE config 2 at its largest size, 8192 x 8192 (537 MB, HBM streamed), truncated at 32 pivots so
    that the single-threaded oracle finishes in seconds; permutations and factors are bit-identical."""
    A = lowrank_matrix(8192, 8192, 64, seed=2)
    lu = T.rrlu(A, maxrank=32, reltol=1e-12)
    ref = oracle.rrlu(A, maxrank=32, reltol=1e-12)
    assert_lu_equal(lu, ref)


def test_rrlu_config2_headline_vs_oracle(T, oracle):
    """BASELINE config 2 at the headline point of bench.py: 8192 x 8192, maxrank 1024, reltol 1e-12, the same generator
    and seed as the bench (bench.factors(m, n, r, 2)).  The single-threaded oracle needs ~100 s for it; permutations,
    pivot errors, lu.error and both factors must be bit-identical (matrixlu.jl:141-181)."""
    A = np.asfortranarray(lowrank_matrix(8192, 8192, 1024, seed=2))
    lu = T.rrlu(A, maxrank=1024, reltol=1e-12)
    ref = oracle.rrlu(A, maxrank=1024, reltol=1e-12)
    assert lu.npivot == ref.npivot == 1024
    assert_lu_equal(lu, ref)


def test_rrlu_config2_full_rank_properties(T):
    """Size-independent properties at 8192 x 8192, maxrank 1024 (no oracle run: ~100 s on a CPU core):
    L unit lower / U upper, permutations are permutations, full-pivoting bound |L| <= 1, sampled
    reconstruction error below the last pivot error."""
    m = n = 8192
    r = 1024
    rng = np.random.default_rng(2)
    p = rng.random((m, r)) * 2.0 ** (-40.0 * np.arange(1, r + 1) / r)
    q = rng.random((r, n))
    A = np.asfortranarray(p @ q)
    lu = T.rrlu(A, maxrank=r, reltol=1e-12)
    assert lu.npivot == r
    assert sorted(lu.rowpermutation.tolist()) == list(range(1, m + 1))
    assert sorted(lu.colpermutation.tolist()) == list(range(1, n + 1))
    L, U = lu.L, lu.U
    assert np.all(L == np.tril(L)) and np.all(U == np.triu(U)) and np.all(np.diag(L) == 1.0)
    assert np.max(np.abs(L)) <= 1.0  # every multiplier is a ratio to the largest entry
    pe = T.pivoterrors(lu)
    assert np.all(np.diff(np.abs(np.diag(U))) <= 1e-9 * pe[0] + 2.0 * np.abs(np.diag(U))[:-1])
    rows = rng.integers(0, m, 200)
    cols = rng.integers(0, n, 200)
    Ap = A[lu.rowpermutation - 1][:, lu.colpermutation - 1]
    rec = np.einsum("ik,ki->i", L[rows, :], U[:, cols])
    assert np.max(np.abs(rec - Ap[rows, cols])) <= 50 * max(lu.error, 1e-16 * pe[0])


@pytest.mark.parametrize("m,n,r", [(17, 33, 9), (130, 257, 40), (513, 700, 48)])
def test_rrlu_streaming_forced(T, oracle, m, n, r, monkeypatch):
    """Same kernels with residency switched off (every size takes the global-memory path)."""
    monkeypatch.setenv("TCI_RRLU_NO_RES", "1")
    A = lowrank_matrix(m, n, r, seed=m * 1000 + n)
    for lo in (True, False):
        assert_lu_equal(T.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo),
                        oracle.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo))


LAZY_SHAPES = [(2, 3, 2), (17, 33, 9), (64, 64, 20), (100, 37, 30), (130, 257, 41), (300, 300, 64), (513, 700, 47),
               (1000, 130, 33), (64, 3000, 21), (2100, 2050, 24)]


@pytest.mark.parametrize("m,n,r", LAZY_SHAPES)
def test_rrlu_deferred_update_kernel(T, oracle, m, n, r, monkeypatch):
    """rrlu_lazy.cu (Schur updates deferred, committed every 4 pivots; normally used from m*n >= 13e6, about 3600^2, up) forced at
    small sizes: block boundaries, ranks that are not multiples of the block, row swaps inside a block."""
    monkeypatch.setenv("TCI_RRLU_NO_RES", "1")
    monkeypatch.setenv("TCI_RRLU_LAZY_MIN", "0")
    A = lowrank_matrix(m, n, r, seed=m * 1000 + n)
    for lo in (True, False):
        assert_lu_equal(T.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo),
                        oracle.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo))
    for mr in (1, 2, 3, 4, 5):  # stop inside / at the end of the first block
        if mr <= min(m, n):
            assert_lu_equal(T.rrlu(A, maxrank=mr), oracle.rrlu(A, maxrank=mr))
    assert_lu_equal(T.rrlu(A, reltol=1e-6), oracle.rrlu(A, reltol=1e-6))  # stop rule with pending updates


@pytest.mark.parametrize("m,n,r", [(64, 3000, 21), (700, 1500, 37), (2100, 2050, 24)])
def test_rrlu_deferred_update_kernel_uneven_ownership(T, oracle, m, n, r, monkeypatch):
    """Speed-weighted column ownership (CTAs ranked by %smid, columns dealt by a quota table) with a strongly
    uneven artificial speed table: who owns a column must not change a single bit of the result."""
    monkeypatch.setenv("TCI_RRLU_NO_RES", "1")
    monkeypatch.setenv("TCI_RRLU_LAZY_MIN", "0")
    monkeypatch.setenv("TCI_RRLU_TEST_SPEEDS", "1")
    A = lowrank_matrix(m, n, r, seed=m * 1000 + n)
    for lo in (True, False):
        assert_lu_equal(T.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo),
                        oracle.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo))
    B = np.random.default_rng(5).integers(-2, 3, (40, 400)).astype(np.float64)  # ties across CTAs
    assert_lu_equal(T.rrlu(B), oracle.rrlu(B))


@pytest.mark.parametrize("m,n,r", [(120000, 300, 11), (300, 120000, 11), (6000, 5000, 13)])
def test_rrlu_deferred_update_kernel_natural_sizes(T, oracle, m, n, r):
    """Shapes that take the deferred-update kernel without forcing (m*n >= 13e6): very tall, very wide and square;
    the rank is not a multiple of the commit block."""
    A = lowrank_matrix(m, n, r + 3, seed=m + n)
    for lo in (True, False):
        assert_lu_equal(T.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo),
                        oracle.rrlu(A, maxrank=r, reltol=1e-12, leftorthogonal=lo))


def test_rrlu_deferred_update_kernel_special_cases(T, oracle, monkeypatch):
    monkeypatch.setenv("TCI_RRLU_NO_RES", "1")
    monkeypatch.setenv("TCI_RRLU_LAZY_MIN", "0")
    rng = np.random.default_rng(99)
    # full rank (every row and column becomes a pivot; the last block is partial)
    for m, n in ((50, 50), (121, 83), (83, 121)):
        A = rng.random((m, n))
        for lo in (True, False):
            assert_lu_equal(T.rrlu(A, leftorthogonal=lo), oracle.rrlu(A, leftorthogonal=lo))
    # exact ties and exact zeros: small-integer matrix (first maximum in column-major order must win)
    A = rng.integers(-3, 4, (90, 70)).astype(np.float64)
    assert_lu_equal(T.rrlu(A), oracle.rrlu(A))
    assert_lu_equal(T.rrlu(A, leftorthogonal=False), oracle.rrlu(A, leftorthogonal=False))
    # pivots already on the diagonal (no row swaps) and in reverse order (every pivot swaps)
    D = np.diag(np.arange(40, 0, -1.0))
    assert_lu_equal(T.rrlu(D), oracle.rrlu(D))
    assert_lu_equal(T.rrlu(D[::-1].copy()), oracle.rrlu(D[::-1].copy()))
    # NaN in the matrix: never selected while finite entries remain (matrixlu.jl:16-29), error at the end
    B = rng.random((60, 60))
    B[7, 9] = np.nan
    with pytest.raises(T.TCIError):
        T.rrlu(B)
    lu = T.rrlu(rng.random((6, 5000)), maxrank=4)  # more CTAs than rows
    assert lu.npivot == 4
    # all zeros: the first pivot is 0, 0/0 reaches L (matrixlu.jl:164-169)
    with pytest.raises(T.TCIError, match="contains NaNs"):
        T.rrlu(np.zeros((40, 30)))
    # huge entries: abs2 overflows to Inf, ties on Inf are broken by scan order; Inf * 0 must not leak anywhere
    H = rng.random((70, 90))
    H[3, 4], H[60, 2], H[33, 80] = 1e200, -1e200, 1e200
    for mr in (3, 6, 9):
        assert_lu_equal(T.rrlu(H, maxrank=mr), oracle.rrlu(H, maxrank=mr))
    assert_lu_equal(T.rrlu(H * 1e-300, maxrank=6), oracle.rrlu(H * 1e-300, maxrank=6))
    # a matrix with exactly representable entries and rank 5: the trailing block becomes exactly zero
    Z = (rng.integers(-4, 5, (64, 5)) @ rng.integers(-4, 5, (5, 48))).astype(np.float64)
    assert_lu_equal(T.rrlu(Z), oracle.rrlu(Z))


@pytest.mark.parametrize("m,n", [(50, 50), (120, 80), (257, 300)])
def test_rrlu_full_rank_random(T, oracle, m, n):  # benchmark/rrlu.jl:12-17 shape
    A = np.random.default_rng(m).random((m, n))
    lu = T.rrlu(A)
    ref = oracle.rrlu(A)
    assert_lu_equal(lu, ref)
    assert lu.npivot == min(m, n) and lu.error == 0.0


def test_rrlu_exact_ties(T, oracle):
    """Symmetric targets make many entries bit-equal; the winner is defined by scan order
    (columns outer, rows inner, strict >) on the swapped matrix."""
    rng = np.random.default_rng(3)
    A = rng.integers(-3, 4, (60, 45)).astype(float)
    assert_lu_equal(T.rrlu(A), oracle.rrlu(A))
    v = np.arange(1, 41, dtype=float)
    L = 1.0 / (1.0 + v[:, None] ** 2 + v[None, :] ** 2)  # symmetric Lorentzian slice
    assert_lu_equal(T.rrlu(L, reltol=1e-13), oracle.rrlu(L, reltol=1e-13))
    assert_lu_equal(T.rrlu(np.ones((33, 65))), oracle.rrlu(np.ones((33, 65))))
    assert_lu_equal(T.rrlu(np.zeros((5, 4)) + 2.0, maxrank=1), oracle.rrlu(np.zeros((5, 4)) + 2.0, maxrank=1))


def test_rrlu_nan_and_errors(T):
    A = np.zeros((3, 3))
    with pytest.raises(T.TCIError, match="contains NaNs"):  # matrixlu.jl:164-169 (0/0 in the first column)
        T.rrlu(A)
    B = np.random.default_rng(0).random((6, 6))
    B[2, 3] = np.nan
    with pytest.raises(T.TCIError, match="contains NaNs"):
        T.rrlu(B)


def test_rrlu_fast_mode_same_pivots(T, oracle):
    A = lowrank_matrix(200, 180, 30, seed=9)
    lu = T.rrlu(A, maxrank=30, reltol=1e-12, exact=False)
    ref = oracle.rrlu(A, maxrank=30, reltol=1e-12)
    # fused multiply-add changes last bits only: same pivots, same quality of the factorisation
    assert lu.npivot == ref.npivot
    assert np.array_equal(lu.rowpermutation, ref.rowpermutation)
    assert np.array_equal(lu.colpermutation, ref.colpermutation)
    np.testing.assert_allclose(lu._pivoterrors, ref.pivoterrors, rtol=1e-6, atol=1e-13 * ref.pivoterrors[0])
    assert np.max(np.abs(T.left(lu) @ T.right(lu) - A)) <= 10 * lu.error


def test_rrlu_async_upload_and_fetch_into(T, oracle):
    """tci_dmat_create_async (upload on the copy stream, consumers order themselves after it) + fetch_into."""
    import torch
    ctx = T.default_context()
    mats, refs = [], []
    for k in range(3):  # several uploads in flight before the first one is consumed
        m, n, r = 300 + 50 * k, 200 + 30 * k, 40
        pinned = torch.empty((n, m), dtype=torch.float64, pin_memory=True).numpy().T
        pinned[:] = lowrank_matrix(m, n, r, seed=70 + k)
        mats.append((T.DeviceMatrix.from_host_async(ctx, pinned), pinned, r))
    for dm, host, r in mats:
        ref = oracle.rrlu(host, maxrank=r, reltol=1e-12)
        lu = T.rrlu(dm, maxrank=r, reltol=1e-12)
        L = np.zeros((host.shape[0], lu.npivot), order="F")
        U = np.zeros((lu.npivot, host.shape[1]), order="F")
        lu.fetch_into(L, U)
        assert_lu_equal(lu, ref)
        assert np.array_equal(L, ref.L) and np.array_equal(U, ref.U)
    dm = T.DeviceMatrix.from_host_async(ctx, np.asfortranarray(np.arange(12.0).reshape(3, 4)))
    assert np.array_equal(dm.to_host(), np.arange(12.0).reshape(3, 4))  # fetch waits for the upload
    T.DeviceMatrix.from_host_async(ctx, np.asfortranarray(np.ones((5, 5))))  # destroyed before first use
    with pytest.raises(ValueError):
        T.DeviceMatrix.from_host_async(ctx, np.ones((3, 4)))  # not Fortran-ordered


def test_rrlu_device_input_from_pi_eval(T, oracle):
    ld = [10] * 6
    f, o = make_target(T, oracle, LORENTZ, [1.0], ld)
    rng = np.random.default_rng(8)
    I = rand_indexset(rng, ld[:3], 200)
    J = rand_indexset(rng, ld[3:], 150)
    dev, mx = f.batchevaluate_device(I, J, 0)
    Pi, omx = o.pi_eval(I.tolist(), J.tolist(), 0, 0.0)
    luci = T.MatrixLUCI(dev, reltol=1e-14, abstol=1e-10 * mx, maxrank=40, leftorthogonal=True)
    ref = oracle.luci(Pi, reltol=1e-14, abstol=1e-10 * omx, maxrank=40, leftorthogonal=True)
    assert luci.npivot == ref.npivot
    assert np.array_equal(T.rowindices(luci), ref.rowindices)
    assert np.array_equal(T.colindices(luci), ref.colindices)
    assert np.array_equal(T.pivoterrors(luci), ref.pivoterrors)


# ------------------------------------------------------------ rook search ---
@pytest.mark.parametrize("leftorth", [True, False])
def test_arrlu_matches_oracle(T, oracle, leftorth):  # test_matrixlu.jl:71-86 + matrixlu.jl:227-293
    ctx = T.default_context()
    cases = [(G.RRLU_4x4, [1], [1], {}), (lowrank_matrix(60, 45, 5, seed=1, decay=8.0), [3], [7], {"reltol": 1e-10}),
             (lowrank_matrix(200, 300, 12, seed=2, decay=20.0), [5, 9], [1, 2, 3], {"reltol": 1e-9, "maxrank": 10})]
    for A, I0, J0, kw in cases:
        A = np.asfortranarray(A)

        def fsub(ir, ic, device=True):
            blk = np.asfortranarray(A[np.ix_(np.asarray(ir) - 1, np.asarray(ic) - 1)])
            return T.DeviceMatrix.from_host(ctx, blk) if device else blk

        lu = T.arrlu(fsub, A.shape, I0, J0, leftorthogonal=leftorth, rng=T.CounterRNG(4), **kw)
        ref = oracle.arrlu(A, I0, J0, leftorthogonal=leftorth, seed=4, **kw)
        assert lu.npivot == ref.npivot
        assert np.array_equal(lu.rowpermutation, ref.rowpermutation)
        assert np.array_equal(lu.colpermutation, ref.colpermutation)
        assert lu.error == ref.error
        # the completed rows / columns go through a triangular solve with the (ill-conditioned) pivot block
        np.testing.assert_allclose(lu.L, ref.L, rtol=1e-8, atol=1e-10 * np.max(np.abs(ref.L)))
        np.testing.assert_allclose(lu.U, ref.U, rtol=1e-8, atol=1e-10 * np.max(np.abs(ref.U)))
        if A.shape == (4, 4):
            assert np.all(lu.L == np.tril(lu.L)) and np.all(lu.U == np.triu(lu.U))
            np.testing.assert_allclose(T.left(lu) @ T.right(lu), A, rtol=1.5e-8)


# ------------------------------------------------------------------ K3 ------
@pytest.mark.parametrize("leftorth", [True, False])
@pytest.mark.parametrize("m,n,r", [(8, 6, 4), (40, 70, 33), (300, 200, 64), (129, 515, 100), (700, 90, 90)])
def test_luci_left_right(T, oracle, m, n, r, leftorth):  # test_matrixluci.jl:6-74
    A = G.LUCI_8x6 if (m, n) == (8, 6) else lowrank_matrix(m, n, r, seed=m + n, decay=20.0)
    luci = T.MatrixLUCI(A, maxrank=r, leftorthogonal=leftorth)
    ref = oracle.luci(A, maxrank=r, leftorthogonal=leftorth)
    assert luci.npivot == ref.npivot
    Lg, Rg = T.left(luci), T.right(luci)
    scale_l, scale_r = np.max(np.abs(ref.left)), np.max(np.abs(ref.right))
    assert np.max(np.abs(Lg - ref.left)) <= RTOL * scale_l
    assert np.max(np.abs(Rg - ref.right)) <= RTOL * scale_r
    Ld = luci.left(device=True).to_host()
    assert np.array_equal(Ld, Lg)


def test_dgemm_kernel(T, ctx):
    import ctypes as C
    from tci_b200 import _lib
    rng = np.random.default_rng(1)
    # (the last three shapes take the bulk-copy / mbarrier kernel: M, N multiples of 128, K a multiple of 16, enough
    # CTAs for every SM; K = 16 and K = 48 are shorter than its prefetch distance)
    for (M, N, K), (ta, tb) in itertools.product([(1, 1, 1), (5, 130, 17), (257, 129, 300), (64, 64, 64), (300, 7, 513),
                                                  (1536, 1664, 16), (2048, 1280, 48), (1536, 1664, 400)],
                                                 [(0, 0), (1, 0), (0, 1), (1, 1)]):
        A = np.asfortranarray(rng.standard_normal((K, M) if ta else (M, K)))
        B = np.asfortranarray(rng.standard_normal((N, K) if tb else (K, N)))
        Cm = np.asfortranarray(rng.standard_normal((M, N)))
        ref = 0.5 * (A.T if ta else A) @ (B.T if tb else B) - 2.0 * Cm
        out = Cm.copy(order="F")
        ctx.check(_lib.lib().tci_dgemm_host(ctx.h, ta, tb, M, N, K, 0.5, _lib.pf(A), _lib.pf(B), -2.0, _lib.pf(out)))
        assert np.max(np.abs(out - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))) * K


# ------------------------------------------------------------- K4 / K5 ------
def _rand_tt(rng, bonds, dims):
    return [np.asfortranarray(rng.random((bonds[i], dims[i], bonds[i + 1])) - 0.3) for i in range(len(dims))]


def test_tt_target_all_splits(T, oracle):  # test_tensortrain.jl:113-140, test_cachedtensortrain.jl:8-65
    rng = np.random.default_rng(3)
    dims = [2, 3, 3, 2, 4]
    cores = _rand_tt(rng, [1, 2, 5, 3, 2, 1], dims)
    f = T.TTCache(T.TensorTrain(cores))
    o = oracle.Target.tt(cores)
    N = len(dims)
    pts = np.array(list(itertools.product(*[range(1, d + 1) for d in dims])), dtype=np.int64)
    got = f.evaluate_points(pts)
    ref = np.array([o(p) for p in pts])
    assert np.array_equal(got, ref)  # same operation order -> bit exact
    ev = T.evaluate_points(T.TensorTrain(cores), pts)
    assert np.array_equal(ev, np.array([oracle.tt_evaluate(cores, p) for p in pts]))
    for nl in range(N + 1):
        for nr in range(N - nl + 1):
            M = N - nl - nr
            I = rand_indexset(rng, dims[:nl], 11 if nl else 1)
            J = rand_indexset(rng, dims[N - nr:], 7 if nr else 1)
            res = f(I, J, M)
            oref, _ = o.pi_eval(I.tolist(), J.tolist(), M)
            np.testing.assert_allclose(res, oref, rtol=RTOL, atol=1e-14)


def test_tt_evaluate_tiled_chain_bit_exact(T, oracle, monkeypatch):
    """The tiled, bucket-ordered chain step (k_env_left_step_tiled) against the oracle's left-to-right product
    (abstracttensortrain.jl:124-132), bit for bit: ragged point counts, bond dimensions that are neither even nor
    multiples of the staging chunk, a site of dimension 1, star-shaped point sets (consecutive points that differ in
    one site only) and random ones; the untiled kernel must give the same bits."""
    rng = np.random.default_rng(21)
    dims = [3, 5, 1, 4, 6, 2]
    cores = _rand_tt(rng, [1, 9, 70, 131, 130, 17, 1], dims)
    tt = T.TensorTrain(cores)
    rnd = rand_indexset(rng, dims, 1003)
    start = rand_indexset(rng, dims, 40)
    star = []
    for x in start:
        for p, d in enumerate(dims):
            for v in range(1, d + 1):
                y = x.copy()
                y[p] = v
                star.append(y)
    star = np.array(star, dtype=np.int64)
    for pts in (rnd, star, rnd[:129], star[:257]):
        ref = np.array([oracle.tt_evaluate(cores, p) for p in pts])
        got = T.evaluate_points(tt, pts)
        assert np.array_equal(got, ref)
        monkeypatch.setenv("TCI_TT_NO_BUCKETS", "1")
        assert np.array_equal(T.evaluate_points(tt, pts), ref)
        monkeypatch.setenv("TCI_TT_NO_TILED", "1")
        assert np.array_equal(T.evaluate_points(tt, pts), ref)
        monkeypatch.delenv("TCI_TT_NO_BUCKETS")
        monkeypatch.delenv("TCI_TT_NO_TILED")
    f = T.TTCache(tt)  # TTCache.evaluate: left half through the same chain
    o = oracle.Target.tt(cores)
    assert np.array_equal(f.evaluate_points(rnd[:300]), np.array([o(p) for p in rnd[:300]]))


def _rand_mpo(rng, bonds, d1, d2):
    return [np.asfortranarray(rng.random((bonds[i], d1[i], d2[i], bonds[i + 1])) - 0.5) for i in range(len(d1))]


def test_mpo_target(T, oracle):  # test_contraction.jl:68-146 (real-valued)
    rng = np.random.default_rng(5)
    N = 5
    d1, d2, d3 = [2, 2, 3, 2, 2], [2, 3, 2, 2, 3], [3, 2, 2, 2, 2]
    A = _rand_mpo(rng, [1, 2, 3, 4, 2, 1], d1, d2)
    B = _rand_mpo(rng, [1, 3, 2, 3, 3, 1], d2, d3)
    f = T.Contraction(T.TensorTrain(A), T.TensorTrain(B))
    o = oracle.Target.mpo_pair(A, B)
    ld = f.localdims
    assert ld == o.localdims
    pts = rand_indexset(rng, ld, 64)
    np.testing.assert_allclose(f.evaluate_points(pts), np.array([o(p) for p in pts]), rtol=RTOL, atol=1e-14)
    for nl, nr in ((1, 1), (0, 2), (2, 0), (1, 2), (2, 2), (0, 0), (2, 3), (3, 2), (4, 1), (0, 5), (5, 0)):
        M = N - nl - nr
        I = rand_indexset(rng, ld[:nl], 9 if nl else 1)
        J = rand_indexset(rng, ld[N - nr:], 6 if nr else 1)
        res = f(I, J, M)
        oref, _ = o.pi_eval(I.tolist(), J.tolist(), M)
        np.testing.assert_allclose(res, oref, rtol=RTOL, atol=1e-13)


def _nested_sets(T, rng, N, d, keep):
    """left[k] / right[k]: nested sets of prefixes / suffixes of length k, grown site by site as TCI grows Iset / Jset."""
    left, right = {1: np.arange(1, d + 1, dtype=np.int64)[:, None]}, {1: np.arange(1, d + 1, dtype=np.int64)[:, None]}
    for k in range(2, N):
        cl, cr = T.kronecker_left(left[k - 1], d), T.kronecker_right(d, right[k - 1])
        left[k] = cl[np.sort(rng.choice(len(cl), min(keep, len(cl)), replace=False))]
        right[k] = cr[np.sort(rng.choice(len(cr), min(keep, len(cr)), replace=False))]
    return left, right


def test_mpo_shared_prefixes(T, oracle, monkeypatch):
    """Nested index sets (Icombined = kronecker(Iset, d) with Iset grown site by site, as a TCI run produces them): the
    chain evaluates every distinct prefix / suffix once (ChainPlan in csrc/mpo.cu -- the within-call effect of the
    reference's Dict memo, contraction.jl:112-176).  Same Pi as the oracle, and as the plain chain."""
    rng = np.random.default_rng(77)
    N = 10
    d1, d2, d3 = [2] * N, [2, 3] * (N // 2), [2] * N
    bonds = [1, 3, 5, 6, 6, 7, 6, 6, 5, 3, 1]
    A = _rand_mpo(rng, bonds, d1, d2)
    B = _rand_mpo(rng, bonds, d2, d3)
    f = T.Contraction(T.TensorTrain(A), T.TensorTrain(B))
    o = oracle.Target.mpo_pair(A, B)
    left, right = _nested_sets(T, rng, N, 4, 13)
    for nl in range(1, N):
        nr = N - nl
        I = T.kronecker_left(left[nl - 1], 4) if nl > 1 else left[1]
        J = T.kronecker_right(4, right[nr - 1]) if nr > 1 else right[1]
        assert I.shape[1] == nl and J.shape[1] == nr
        res = f(I, J, 0)
        oref, _ = o.pi_eval(I.tolist(), J.tolist(), 0)
        np.testing.assert_allclose(res, oref, rtol=RTOL, atol=1e-13)
        monkeypatch.setenv("TCI_MPO_NO_DEDUP", "1")
        plain = f(I, J, 0)
        monkeypatch.delenv("TCI_MPO_NO_DEDUP")
        np.testing.assert_allclose(res, plain, rtol=1e-12, atol=1e-15)
        # duplicated entries in the index sets (the gather after the last level)
        Id, Jd = np.concatenate([I, I[:3]]), np.concatenate([J[-2:], J])
        resd = f(Id, Jd, 0)
        tol = dict(rtol=1e-12, atol=1e-15)  # (the final GEMM may tile the larger matrix differently)
        np.testing.assert_allclose(resd[:len(I), 2:], res, **tol)
        np.testing.assert_allclose(resd[len(I):, 2:], res[:3], **tol)
        np.testing.assert_allclose(resd[:len(I), :2], res[:, -2:], **tol)
    # M = 1 and M = 2 calls go through the same chains
    I, J = T.kronecker_left(left[3], 4), T.kronecker_right(4, right[3])
    for M, Jm in ((2, J), (1, T.kronecker_right(4, right[4]))):
        res = f(I, Jm, M)
        oref, _ = o.pi_eval(I.tolist(), Jm.tolist(), M)
        np.testing.assert_allclose(res, oref, rtol=RTOL, atol=1e-13)


def test_batchevaluate_projector(T, oracle):  # test_contraction.jl:101-139 (real-valued), cachedtensortrain.jl:170-215
    rng = np.random.default_rng(31)
    N = 4
    A = _rand_mpo(rng, [1, 2, 3, 2, 1], [2] * N, [3] * N)
    B = _rand_mpo(rng, [1, 2, 3, 2, 1], [3] * N, [2] * N)
    ab = T.Contraction(T.TensorTrain(A), T.TensorTrain(B))
    o = oracle.Target.mpo_pair(A, B)
    left, right = np.array([[1]]), np.array([[1]])
    ref = ab(left, right, 2)
    oref, _ = o.pi_eval(left.tolist(), right.tolist(), 2)
    np.testing.assert_allclose(ref, oref, rtol=RTOL, atol=1e-14)
    mi = oref.reshape((1, 2, 2, 2, 2, 1), order="F")
    for proj, sl in (([[0, 0], [1, 0]], mi[:, :, :, 0, :, :]), ([[0, 0], [1, 1]], mi[:, :, :, 0, 0, :]),
                     ([[0, 1], [1, 0]], mi[:, :, 0, 0, :, :]), ([[2, 0], [0, 2]], mi[:, 1, :, :, 1, :])):
        res = ab.batchevaluate(left, right, 2, proj)
        np.testing.assert_allclose(res.flatten(order="F"), sl.flatten(order="F"), rtol=RTOL, atol=1e-14)
        assert res.ndim == 4
        # the oracle's restatement slices the cores before contracting, as contraction.jl:290-302 does
        oproj = oracle.mpo_batchevaluate_projected(A, B, left.tolist(), right.tolist(), 2, proj)
        assert res.shape == oproj.shape
        np.testing.assert_allclose(res, oproj, rtol=RTOL, atol=1e-14)
    with pytest.raises(RuntimeError, match="Length mismatch"):
        ab.batchevaluate(left, right, 2, [[0, 0]])
    with pytest.raises(RuntimeError, match="the length must be 2"):
        ab.batchevaluate(left, right, 2, [[0, 0], [1]])
    with pytest.raises(RuntimeError, match="Invalid projector"):
        ab.batchevaluate(left, right, 2, [[0, 0], [3, 0]])
    # TTCache with multi-dimensional sites
    cores = [np.asfortranarray(rng.random((b0, 2, 3, b1)) - 0.4) for b0, b1 in ((1, 3), (3, 4), (4, 2), (2, 1))]
    f = T.TTCache(T.TensorTrain(cores))
    assert f.sitedims == [[2, 3]] * 4 and f.localdims == [6] * 4
    ot = oracle.Target.tt([c.reshape((c.shape[0], 6, c.shape[-1]), order="F") for c in cores])
    I, J = rand_indexset(rng, [6], 5), rand_indexset(rng, [6], 4)
    full, _ = ot.pi_eval(I.tolist(), J.tolist(), 2)
    mi = full.reshape((5, 2, 3, 2, 3, 4), order="F")
    res = f.batchevaluate(I, J, 2, [[0, 2], [1, 0]])
    assert res.shape == (5, 2, 3, 4)
    np.testing.assert_allclose(res, mi[:, :, 1, 0, :, :], rtol=RTOL, atol=1e-14)
    res = f.batchevaluate(I, J, 2, [[1, 2], [2, 3]])
    assert res.shape == (5, 1, 1, 4)
    np.testing.assert_allclose(res[:, 0, 0, :], mi[:, 0, 1, 1, 2, :], rtol=RTOL, atol=1e-14)
    with pytest.raises(RuntimeError, match="Invalid parameter M"):
        f.batchevaluate(I, J, 1)
    with pytest.raises(RuntimeError, match="Invalid length of projector"):
        f.batchevaluate(I, J, 2, [[0, 0]])
    with pytest.raises(RuntimeError, match="Invalid projector"):
        f.batchevaluate(I, J, 2, [[0, 4], [0, 0]])


def test_environments_abi(T, oracle):
    """tci_env_eval / tci_pi_from_envs (the pieces the row-block sharding of the contraction is made of): the M = 0
    Pi assembled from separately evaluated left / right environment blocks equals the oracle's batchevaluate
    (cachedtensortrain.jl:151-215, contraction.jl:236-335) for every split, also block by block."""
    rng = np.random.default_rng(15)
    dims = [2, 3, 3, 2, 4]
    cores = _rand_tt(rng, [1, 2, 5, 3, 2, 1], dims)
    d1, d2, d3 = [2, 2, 3, 2, 2], [2, 3, 2, 2, 3], [3, 2, 2, 2, 2]
    A = _rand_mpo(rng, [1, 2, 3, 4, 2, 1], d1, d2)
    B = _rand_mpo(rng, [1, 3, 2, 3, 3, 1], d2, d3)
    for f, o in ((T.TTCache(T.TensorTrain(cores)), oracle.Target.tt(cores)),
                 (T.Contraction(T.TensorTrain(A), T.TensorTrain(B)), oracle.Target.mpo_pair(A, B))):
        ld, N = f.localdims, len(f.localdims)
        assert f.has_environments
        for nl in range(N + 1):
            nr = N - nl
            I = rand_indexset(rng, ld[:nl], 37 if nl else 1)
            J = rand_indexset(rng, ld[nl:], 21 if nr else 1)
            oref, omx = o.pi_eval(I.tolist(), J.tolist(), 0)
            Dl, Dr = f.env_dim(0, nl), f.env_dim(1, nr)
            assert Dl == Dr
            lenv = T.DeviceMatrix.empty(f.ctx, Dl, len(I))
            renv = T.DeviceMatrix.empty(f.ctx, Dr, len(J))
            h = len(J) // 2
            f.env_eval_into(lenv, 0, 0, I)
            f.env_eval_into(renv, h, 1, J[h:])  # two column blocks, as two ranks would fill them
            f.env_eval_into(renv, 0, 1, J[:h])
            out = T.DeviceMatrix.empty(f.ctx, len(I), len(J))
            mx = f.pi_from_envs(lenv, 0, len(I), renv, 0, len(J), out, 0)
            got = out.to_host()
            np.testing.assert_allclose(got, oref, rtol=RTOL, atol=1e-13)
            assert abs(mx - omx) <= RTOL * omx
            # a row block of Pi written into its place of a larger matrix through a wrapped view
            if len(I) > 16:
                big = T.DeviceMatrix.empty(f.ctx, len(I), len(J))
                view = T.DeviceMatrix.wrap(f.ctx, big.ptr + 8 * 16, len(I) - 16, len(J), big.ld)
                f.pi_from_envs(lenv, 16, len(I) - 16, renv, 0, len(J), view, 0)
                assert np.array_equal(big.to_host()[16:], got[16:])
    g = T.BuiltinTarget(T.LORENTZ, [1.0], [3, 3])
    with pytest.raises(ValueError, match="no environments"):
        g.env_dim(0, 1)


def test_zipup_and_naive_site(T):  # contraction.jl:338-349, 455-464
    import ctypes as C
    from tci_b200 import _lib
    rng = np.random.default_rng(7)
    chi, Da, Db, s1, s2, s3, Dan, Dbn = 5, 4, 3, 2, 3, 2, 6, 5
    R = np.asfortranarray(rng.standard_normal((chi, Da, Db)))
    A = np.asfortranarray(rng.standard_normal((Da, s1, s2, Dan)))
    B = np.asfortranarray(rng.standard_normal((Db, s2, s3, Dbn)))
    ctx = T.default_context()
    out = np.zeros(chi * s1 * s3 * Dan * Dbn)
    ctx.check(_lib.lib().tci_contract_zipup_site(ctx.h, _lib.pf(R), chi, Da, Db, _lib.pf(A), s1, s2, Dan, _lib.pf(B),
                                                 s3, Dbn, _lib.pf(out), None))
    ref = np.einsum("cab,axhn,bhzm->cxznm", R, A, B)
    np.testing.assert_allclose(out.reshape((chi, s1, s3, Dan, Dbn), order="F"), ref, rtol=1e-12, atol=1e-13)
    nv = T._contractsitetensors(A, B)
    refn = np.einsum("axhn,bhzm->abxznm", A, B).reshape((Da * Db, s1, s3, Dan * Dbn), order="F")
    np.testing.assert_allclose(nv, refn, rtol=1e-12, atol=1e-13)


def test_contract_zipup_and_tci_vs_dense(T):  # test_contraction.jl:68-99, 185-195
    rng = np.random.default_rng(9)
    N = 4
    d1, d2, d3 = [2, 2, 2, 2], [2, 3, 2, 2], [2, 2, 3, 2]
    A = _rand_mpo(rng, [1, 2, 3, 2, 1], d1, d2)
    B = _rand_mpo(rng, [1, 3, 2, 3, 1], d2, d3)

    def dense(cores):
        out = cores[0]
        for c in cores[1:]:
            out = np.tensordot(out, c, axes=([-1], [0]))
        return out[0, ..., 0]

    a, b = dense(A), dense(B)
    la = "".join(chr(97 + 2 * s) + chr(97 + 2 * s + 1) for s in range(N))
    lb = "".join(chr(97 + 2 * s + 1) + chr(65 + s) for s in range(N))
    lc = "".join(chr(97 + 2 * s) + chr(65 + s) for s in range(N))
    ref = np.einsum(f"{la},{lb}->{lc}", a, b)
    for method in ("LU", "SVD"):
        tt = T.contract_zipup(T.TensorTrain(A), T.TensorTrain(B), tolerance=1e-13, method=method)
        np.testing.assert_allclose(dense(tt.sitetensors), ref, rtol=1e-8, atol=1e-10)
    tt = T.contract(T.TensorTrain(A), T.TensorTrain(B), algorithm="TCI", tolerance=1e-12, maxbonddim=60)
    np.testing.assert_allclose(dense(tt.sitetensors), ref, rtol=1e-8, atol=1e-10)
    ttn = T.contract(T.TensorTrain(A), T.TensorTrain(B), algorithm="naive", tolerance=0.0)
    np.testing.assert_allclose(dense(ttn.sitetensors), ref, rtol=1e-10, atol=1e-12)


def test_contraction_elementwise_function(T, oracle):
    """Contraction with an elementwise f (contraction.jl:203-205, 330-332; the reference tests use f = x -> 2x,
    test_contraction.jl:148-181): registered by id on the device.  f(Pi) against the oracle's Pi with f applied on
    the host, pointwise evaluation, and contract(...; algorithm=:TCI, f) against the dense product."""
    rng = np.random.default_rng(41)
    N = 4
    d1, d2, d3 = [2, 2, 2, 2], [2, 3, 2, 2], [2, 2, 3, 2]
    A = _rand_mpo(rng, [1, 2, 3, 2, 1], d1, d2)
    B = _rand_mpo(rng, [1, 3, 2, 3, 1], d2, d3)
    o = oracle.Target.mpo_pair(A, B)
    ld = o.localdims
    pts = rand_indexset(rng, ld, 50)
    for spec, fn in ((("affine", 2.0, 0.0), lambda x: 2.0 * x), (("affine", -0.5, 0.25), lambda x: -0.5 * x + 0.25),
                     (("abs",), np.abs), (("square",), lambda x: x * x)):
        f = T.Contraction(T.TensorTrain(A), T.TensorTrain(B), f=spec)
        assert not f.has_environments
        for nl, nr in ((1, 1), (2, 2), (0, 3), (4, 0)):
            M = N - nl - nr
            I = rand_indexset(rng, ld[:nl], 7 if nl else 1)
            J = rand_indexset(rng, ld[N - nr:], 5 if nr else 1)
            oref, _ = o.pi_eval(I.tolist(), J.tolist(), M)
            np.testing.assert_allclose(f(I, J, M), fn(oref), rtol=RTOL, atol=1e-13)
            dev, mx = f.batchevaluate_device(I, J, M)
            assert abs(mx - np.max(np.abs(fn(oref)))) <= RTOL * max(mx, 1e-300)
        np.testing.assert_allclose(f.evaluate_points(pts), fn(np.array([o(p) for p in pts])), rtol=RTOL, atol=1e-13)
    with pytest.raises(NotImplementedError):
        T.Contraction(T.TensorTrain(A), T.TensorTrain(B), f=lambda x: 2 * x)

    def dense(cores):
        out = cores[0]
        for c in cores[1:]:
            out = np.tensordot(out, c, axes=([-1], [0]))
        return out[0, ..., 0]

    la = "".join(chr(97 + 2 * s) + chr(97 + 2 * s + 1) for s in range(N))
    lb = "".join(chr(97 + 2 * s + 1) + chr(65 + s) for s in range(N))
    lc = "".join(chr(97 + 2 * s) + chr(65 + s) for s in range(N))
    ref = np.einsum(f"{la},{lb}->{lc}", dense(A), dense(B))
    tt = T.contract(T.TensorTrain(A), T.TensorTrain(B), algorithm="TCI", tolerance=1e-12, maxbonddim=60,
                    f=("affine", 2.0, 0.0))
    np.testing.assert_allclose(dense(tt.sitetensors), 2.0 * ref, rtol=1e-8, atol=1e-10)
    with pytest.raises(RuntimeError, match="Naive contraction implementation cannot contract"):
        T.contract(T.TensorTrain(A), T.TensorTrain(B), algorithm="naive", f=("affine", 2.0, 0.0))


@pytest.mark.parametrize("algorithm", ["TCI", "naive", "zipup"])
def test_contract_mpo_mps(T, algorithm):  # test_contraction.jl:148-181, 190-194 (real-valued; f = nothing and x -> 2x)
    rng = np.random.default_rng(43)
    N = 4
    A = _rand_mpo(rng, [1, 2, 3, 2, 1], [3] * N, [3] * N)
    b = _rand_tt(rng, [1, 2, 3, 2, 1], [3] * N)

    def dense(cores):
        out = cores[0]
        for c in cores[1:]:
            out = np.tensordot(out, c, axes=([-1], [0]))
        return out[0, ..., 0]

    mat = dense(A).transpose([0, 2, 4, 6, 1, 3, 5, 7]).reshape(81, 81)  # rows: first site legs, columns: second site legs
    vec = dense(b).reshape(81)
    kw = dict(tolerance=1e-12, maxbonddim=40) if algorithm == "TCI" else {}
    ab = T.contract(T.TensorTrain(A), T.TensorTrain(b), algorithm=algorithm, **kw)
    ba = T.contract(T.TensorTrain(b), T.TensorTrain(A), algorithm=algorithm, **kw)
    assert [c.shape[1] for c in ab.sitetensors] == [3] * N and all(c.ndim == 3 for c in ab.sitetensors)
    np.testing.assert_allclose(dense(ab.sitetensors).reshape(81), mat @ vec, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(dense(ba.sitetensors).reshape(81), vec @ mat, rtol=1e-8, atol=1e-10)
    f2 = ("affine", 2.0, 0.0)
    if algorithm == "TCI":
        ab2 = T.contract(T.TensorTrain(A), T.TensorTrain(b), algorithm=algorithm, f=f2, **kw)
        np.testing.assert_allclose(dense(ab2.sitetensors).reshape(81), 2.0 * (mat @ vec), rtol=1e-8, atol=1e-10)
    else:
        with pytest.raises(RuntimeError, match="cannot contract matrix product with a function"):
            T.contract(T.TensorTrain(A), T.TensorTrain(b), algorithm=algorithm, f=f2)


@pytest.mark.parametrize("method", ["LU", "CI", "SVD"])
def test_compress(T, method):  # tensortrain.jl:149-183 ; test_tensortrain.jl compress tests (structure re-used)
    rng = np.random.default_rng(12)
    dims = [3, 4, 3, 4, 3]
    small = _rand_tt(rng, [1, 3, 4, 4, 3, 1], dims)
    # embed the rank-(3,4,4,3) train into bond dimension 9 so that exact compression must find it again
    big = []
    for i, c in enumerate(small):
        Dl, d, Dr = c.shape
        L = np.eye(Dl) if i == 0 else Q_prev
        Q = rng.standard_normal((Dr, 9)) if i < len(small) - 1 else np.eye(Dr)
        core = np.einsum("al,ldr,rb->adb", L, c, Q)
        Q_prev = np.linalg.pinv(Q) if i < len(small) - 1 else None
        big.append(np.asfortranarray(core))
    tt = T.TensorTrain(big)
    full = T.fulltensor(tt)
    T.compress(tt, method, tolerance=1e-11)
    assert max(c.shape[0] for c in tt.sitetensors[1:]) <= 4
    np.testing.assert_allclose(T.fulltensor(tt), full, rtol=1e-8, atol=1e-10 * np.max(np.abs(full)))
    T.compress(tt, method, tolerance=1e-11, maxbonddim=2)
    assert max(c.shape[0] for c in tt.sitetensors[1:]) <= 2


@pytest.mark.parametrize("k,rows,lo", [(1, 5, True), (37, 300, True), (70, 129, False), (300, 2000, True)])
def test_lu_rdiv_matches_host_solve(T, k, rows, lo):  # the `\\` of setsitetensor! tensorci2.jl:391
    rng = np.random.default_rng(k)
    P = np.asfortranarray(rng.standard_normal((k, k)))
    B = np.asfortranarray(rng.standard_normal((rows, k)))
    lu = T.rrlu(P, reltol=0.0, abstol=0.0, leftorthogonal=lo)
    assert lu.npivot == k
    X = lu.rdiv(B)
    ref = np.linalg.solve(P.T, B.T).T
    np.testing.assert_allclose(X, ref, rtol=1e-9, atol=1e-10 * np.abs(ref).max())
    np.testing.assert_allclose(X @ P, B, rtol=0, atol=1e-10 * np.abs(B).max() * np.linalg.cond(P))
    Xd = lu.rdiv(T.DeviceMatrix.from_host(lu.ctx, B), device=True).to_host()
    assert np.array_equal(Xd, X)
    with pytest.raises(ValueError):  # TCI_ERR_ARG
        T.rrlu(np.ones((4, 6), order="F"), maxrank=2).rdiv(np.ones((3, 4)))


def test_tt_to_tci2_conversion_matches_oracle(T, oracle):  # conversion.jl:73-176
    rng = np.random.default_rng(21)
    dims = [3, 4, 2, 4, 3]
    cores = _rand_tt(rng, [1, 3, 7, 6, 3, 1], dims)
    full = T.fulltensor(T.TensorTrain([c.copy(order="F") for c in cores]))
    for kw in (dict(tolerance=1e-13), dict(tolerance=1e-13, maxiter=4), dict(tolerance=1e-3, maxbonddim=4)):
        tt = T.TensorTrain([c.copy(order="F") for c in cores])
        tci = T.tensorci2_from_tensortrain(tt, **kw)
        I, J, ocores, pe, mx = oracle.tensorci2_from_tt(cores, **{k: v for k, v in kw.items()})
        for a, b in zip(tci.Iset, I):
            assert np.array_equal(a, b)
        for a, b in zip(tci.Jset, J):
            assert np.array_equal(a, b)
        np.testing.assert_allclose(tci.pivoterrors, pe, rtol=1e-10, atol=1e-13)
        assert abs(tci.maxsamplevalue - mx) <= 1e-10 * mx
        for a, b in zip(tci.sitetensors, ocores):
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-10 * mx)
        assert tci.sitetensors is tt.sitetensors  # conversion.jl:170
        if "maxbonddim" not in kw:
            np.testing.assert_allclose(T.fulltensor(T.TensorTrain(tci.sitetensors)), full, rtol=1e-9, atol=1e-11)
        else:
            assert max(T.linkdims(tci)) <= 4


def test_tt_to_tci2_then_optimize(T):  # test_conversion.jl:76-99 (real-valued)
    ld = [4] * 4
    f = T.BuiltinTarget(LORENTZ, [1.0], ld)
    tci, _, _ = T.crossinterpolate2(f, ld, tolerance=1e-14, maxbonddim=5)
    tt = T.TensorTrain([c.copy(order="F") for c in tci.sitetensors])
    tcib = T.tensorci2_from_tensortrain(tt, tolerance=1e-14)
    assert T.rank(tcib) == 5 and T.linkdims(tcib) == T.linkdims(tci)
    for v in itertools.product(range(1, 5), repeat=4):
        assert abs(tcib(list(v)) - tci(list(v))) < 1e-13
    T.optimize(tcib, f, tolerance=1e-14)
    for v in itertools.product(range(1, 5), repeat=4):
        assert abs(tcib(list(v)) - 1.0 / (1.0 + sum(x * x for x in v))) < 1e-13


def test_tensortrain_arithmetic(T):  # test_tensortrain.jl (add / subtract / multiply / divide / reverse / norm / sum)
    rng = np.random.default_rng(31)
    dims = [3, 2, 4, 3]
    a = T.TensorTrain(_rand_tt(rng, [1, 3, 4, 2, 1], dims))
    b = T.TensorTrain(_rand_tt(rng, [1, 2, 5, 3, 1], dims))
    A, B = T.fulltensor(a), T.fulltensor(b)
    c = T.add(a, b)
    # bond dimensions add up (abstracttensortrain.jl:238) before the lossless recompression caps them at 3, 6, 3
    assert [t.shape[0] for t in c.sitetensors[1:]] == [3, 6, 3]
    np.testing.assert_allclose(T.fulltensor(c), A + B, rtol=1e-12, atol=1e-13)
    c2 = T.add(a, b, factorlhs=2.0, factorrhs=-0.5, tolerance=1e-12)
    np.testing.assert_allclose(T.fulltensor(c2), 2.0 * A - 0.5 * B, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(T.fulltensor(a - b), A - B, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(T.fulltensor(a + a), 2.0 * A, rtol=1e-12, atol=1e-13)
    z = T.subtract(a, a, tolerance=1e-10)  # exact cancellation compresses to bond dimension 1
    assert max(t.shape[0] for t in z.sitetensors[1:]) <= 1 or np.max(np.abs(T.fulltensor(z))) < 1e-12
    np.testing.assert_allclose(T.fulltensor(3.0 * a), 3.0 * A, rtol=1e-14)
    np.testing.assert_allclose(T.fulltensor(a * 3.0), 3.0 * A, rtol=1e-14)
    np.testing.assert_allclose(T.fulltensor(a / 4.0), A / 4.0, rtol=1e-14)
    np.testing.assert_allclose(T.fulltensor(T.reverse(a)), np.transpose(A, (3, 2, 1, 0)), rtol=1e-14)
    assert abs(T.norm2(a) - np.sum(A * A)) <= 1e-12 * np.sum(A * A)
    assert abs(T.norm(b) - np.linalg.norm(B)) <= 1e-12 * np.linalg.norm(B)
    assert abs(T.sum_dims(a) - A.sum()) <= 1e-12 * abs(A).sum()
    assert abs(T.sum_dims(a) - T.tt_sum(a)) <= 1e-12 * abs(A).sum()
    s2 = T.sum_dims(a, dims=(2, 4))
    np.testing.assert_allclose(T.fulltensor(s2), A.sum(axis=(1, 3)), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(T.fulltensor(T.sum_dims(a, dims=1)), A.sum(axis=0), rtol=1e-12, atol=1e-13)
    assert abs(a([2, 1, 3, 2]) - A[1, 0, 2, 1]) <= 1e-14 * np.max(np.abs(A))
    tt2 = T.tensortrain(a)
    tt2.sitetensors[0][0, 0, 0] += 1.0  # tensortrain() copies
    assert a.sitetensors[0][0, 0, 0] != tt2.sitetensors[0][0, 0, 0]
    with pytest.raises(ValueError):
        T.add(a, T.TensorTrain(_rand_tt(rng, [1, 2, 1], [3, 2])))


# ------------------------------------------------------- K7 + the driver ----
def test_globalsearch_matches_oracle(T, oracle):  # test_globalsearch.jl:7-36
    R = 10
    f, o = make_target(T, oracle, Q1D, [R, 1], [2] * R)
    res = oracle.crossinterpolate2(o, [2] * R, [[1] * R, [1] + [2] * (R - 1)], tolerance=1e-4, maxbonddim=1,
                                   normalizeerror=False)
    starts = oracle.start_points(7, 1, 20, [2] * R)  # (n, nsearch)
    piv, errs = oracle.globalsearch(o, res.sitetensors, starts, abstol=1e-9, tolmargin=1.0, maxn=20)
    finder = T.DefaultGlobalPivotFinder(nsearch=20, maxnglobalpivot=20, tolmarginglobalsearch=1.0)
    inp = T.GlobalPivotSearchInput([2] * R, T.TensorTrain(res.sitetensors), res.maxsamplevalue, None, None)
    got = finder(inp, f, 1e-9, rng=np.ascontiguousarray(starts.T))
    assert got.tolist() == piv
    assert np.array_equal(finder.last_errors, errs)


def test_estimatetrueerror(T):  # test_globalsearch.jl:7-36
    R = 20
    f = T.BuiltinTarget(Q1D, [R, 1], [2] * R)  # exp(-x) + 1e-3 sin(1000 x)
    tci, ranks, errors = T.crossinterpolate2(f, [2] * R, [[1] * R, [1] + [2] * (R - 1)], tolerance=1e-4, maxbonddim=1,
                                             normalizeerror=False)
    pivoterrors = T.estimatetrueerror(T.TensorTrain(tci.sitetensors), f, nsearch=12, rng=T.CounterRNG(9))
    errs = [e for _, e in pivoterrors]
    for p, e in pivoterrors:
        assert abs(abs(f(p) - tci(p)) - e) <= 1e-12 * max(1.0, e) + 1e-15
    assert all(a >= b for a, b in zip(errs[:-1], errs[1:]))  # sorted in descending order
    assert len(set((tuple(p), e) for p, e in pivoterrors)) == len(pivoterrors)


def compare_tci(tci, ranks, errors, res, T):
    n = len(tci)
    assert [int(r) for r in ranks] == res.ranks.tolist()
    for b in range(n):
        assert [tuple(x) for x in tci.Iset[b].tolist()] == res.Iset[b], f"Iset[{b}]"
        assert [tuple(x) for x in tci.Jset[b].tolist()] == res.Jset[b], f"Jset[{b}]"
    np.testing.assert_allclose(errors, res.errors, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(tci.pivoterrors, res.pivoterrors, rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(tci.bonderrors, res.bonderrors, rtol=1e-9, atol=1e-300)
    assert tci.maxsamplevalue == res.maxsamplevalue
    for b in range(n):
        ref = res.sitetensors[b]
        assert tci.sitetensors[b].shape == ref.shape
        assert np.max(np.abs(tci.sitetensors[b] - ref)) <= RTOL * max(1.0, np.max(np.abs(ref)))
    s_ref = res.sum()
    assert abs(T.tci_sum(tci) - s_ref) <= RTOL * abs(s_ref)


def test_crossinterpolate2_pivoterrors_diag(T):  # test_tensorci2.jl:27-39
    table = np.diag(G.PIVOTERRORS_DIAGS).flatten(order="F")
    f = T.BuiltinTarget(TABLE, table, [3, 3])
    tci, ranks, errors = T.crossinterpolate2(f, [3, 3], [[1, 1]], tolerance=1e-8)
    assert tci.pivoterrors.tolist() == G.PIVOTERRORS_DIAGS


@pytest.mark.parametrize("pivotsearch", ["full", "rook"])
def test_lorentz_rank_sequence_and_1site_global_pivot(T, pivotsearch):  # test_tensorci2.jl:247-283
    n = 5
    ld = [10] * n
    f = T.BuiltinTarget(LORENTZ, [1.0], ld)
    tci = T.TensorCI2(f, ld)
    assert T.linkdims(tci) == [1] * (n - 1) and T.rank(tci) == 1
    assert len(tci.Iset[0]) == 1 and len(tci.Jset[-1]) == 1
    for b in range(n - 1):
        T.updatepivots(tci, b, f, True, reltol=1e-8, maxbonddim=2, pivotsearch=pivotsearch, rng=T.CounterRNG(2))
    assert T.linkdims(tci) == [2] * (n - 1) and T.rank(tci) == 2
    globalpivot = [2, 9, 10, 5, 7]
    T.addglobalpivots1sitesweep(tci, f, [globalpivot], reltol=1e-12)
    assert T.linkdims(tci) == [3] * (n - 1) and T.rank(tci) == 3
    assert len(tci.Iset[0]) == 1 and len(tci.Jset[-1]) == 1
    assert all(T.existaspivot(tci, globalpivot)[1:-1]) or T.rank(tci) == 3  # the pivot survives the 1-site sweeps
    for _ in range(4, 21):
        for b in range(n - 1):
            T.updatepivots(tci, b, f, True, reltol=1e-8, pivotsearch=pivotsearch, rng=T.CounterRNG(2))
    tci2, _, _ = T.crossinterpolate2(f, ld, [[1] * n], tolerance=1e-8, maxiter=8, sweepstrategy="forward",
                                     pivotsearch=pivotsearch, rng=T.CounterRNG(2))
    if pivotsearch == "full":
        assert T.rank(tci) == T.rank(tci2)


def test_optfirstpivot(T):  # util.jl:78-109, used at test_tensorci2.jl:444
    ld = [5, 4, 6, 3]
    table = np.random.default_rng(8).standard_normal(int(np.prod(ld)))
    f = T.BuiltinTarget(TABLE, table, ld)
    tab = table.reshape(ld, order="F")

    def reference(first, maxsweep=1000):  # the reference's point-by-point loop
        pivot = list(first)
        valf = abs(tab[tuple(p - 1 for p in pivot)])
        for _ in range(maxsweep):
            prev = valf
            for i in range(len(ld)):
                for d in range(1, ld[i] + 1):
                    bak = pivot[i]
                    pivot[i] = d
                    newval = abs(tab[tuple(p - 1 for p in pivot)])
                    if newval > valf:
                        valf = newval
                    else:
                        pivot[i] = bak
            if prev == valf:
                break
        return pivot

    for first in ([1, 1, 1, 1], [5, 4, 6, 3], [2, 3, 1, 2]):
        assert T.optfirstpivot(f, ld, first) == reference(first)
    assert T.optfirstpivot(f, ld) == reference([1, 1, 1, 1])
    assert T.optfirstpivot(f, ld, [2, 3, 1, 2], maxsweep=1) == reference([2, 3, 1, 2], maxsweep=1)


def test_insert_global_pivots_2site(T):  # test_tensorci2.jl:395-431
    R = 12  # (the reference uses R = 20; 2^R table entries here)
    ld = [2] * R
    table = np.zeros(2 ** R)
    table[0] = table[-1] = 1.0  # f(q) = 1 for q == all ones or q == all twos, else 0
    f = T.BuiltinTarget(TABLE, table, ld)
    tci, ranks, errors = T.crossinterpolate2(f, ld, [[1] * R], tolerance=1e-4, maxbonddim=1000, maxiter=20,
                                             normalizeerror=False, strictlynested=False, rng=T.CounterRNG(1234))
    r = [2] * R
    left = T.addglobalpivots2sitesweep(tci, f, [r], tolerance=1e-4, normalizeerror=False, maxbonddim=1000,
                                       strictlynested=False)
    assert left == 0
    assert abs(tci(r) - 1.0) <= 1e-8 and abs(tci([1] * R) - 1.0) <= 1e-8
    assert abs(tci([1, 2] * (R // 2))) <= 1e-8
    with pytest.raises(ValueError):
        T.addglobalpivots2sitesweep(tci, f, [[1, 2, 1]])


def test_sweep0site_and_searchglobalpivots(T):  # tensorci2.jl:341-366, 958-1000
    ld = [6] * 5
    f = T.BuiltinTarget(LORENTZ, [1.0], ld)
    tci, _, _ = T.crossinterpolate2(f, ld, tolerance=1e-6, rng=T.CounterRNG(1))
    dims = T.linkdims(tci)
    for b in range(len(ld) - 1):  # with loose tolerances the weakest pivots of every bond are removed ...
        T.sweep0site(tci, f, b, reltol=1e-3, abstol=0.0)
    after = T.linkdims(tci)
    assert all(a <= d for a, d in zip(after, dims)) and sum(after) < sum(dims) and min(after) >= 1
    for b in range(len(ld)):
        assert len(tci.Iset[b + 1] if b + 1 < len(ld) else tci.Jset[b]) >= 1
    found = T.searchglobalpivots(tci, f, 1e-9, nsearch=20, maxnglobalpivot=3, rng=np.random.default_rng(0))
    assert 1 <= len(found) <= 3  # ... so the truncated interpolation has points with error above 1e-9
    for p in found:
        assert abs(tci(p) - f(p)) > 1e-9
    tci2, _, _ = T.crossinterpolate2(f, ld, tolerance=1e-12, maxiter=30, rng=T.CounterRNG(1))
    assert T.searchglobalpivots(tci2, f, 1e-6, nsearch=5, rng=np.random.default_rng(0)) == []
    assert T.searchglobalpivots(tci2, f, 1e-6, nsearch=0) == []
    T.makecanonical(tci2, f, reltol=1e-12)
    for v in itertools.product(range(1, 4), repeat=5):
        fv = 1.0 / (1.0 + sum(x * x for x in v))
        assert abs(tci2(list(v)) - fv) <= 1e-8 * fv


def test_crossinterpolate2_lorentz_matches_oracle(T, oracle):  # README.md:21-29 (config 1, 6 sites here)
    ld = [10] * 6
    f, o = make_target(T, oracle, LORENTZ, [1.0], ld)
    tci, ranks, errors = T.crossinterpolate2(f, ld, tolerance=1e-8, rng=T.CounterRNG(1))
    res = oracle.crossinterpolate2(o, ld, tolerance=1e-8, seed=1)
    compare_tci(tci, ranks, errors, res, T)
    for v in itertools.product(range(1, 4), repeat=6):  # test_tensorci2.jl:335-339
        fv = 1.0 / (1.0 + sum(x * x for x in v))
        assert abs(tci(list(v)) - fv) <= 1e-6 * fv


@pytest.mark.parametrize("strategy", ["backandforth", "forward"])
def test_crossinterpolate2_quantics2d_matches_oracle(T, oracle, strategy):  # config 3, reduced to R=8
    ld = [4] * 8
    f, o = make_target(T, oracle, Q2D, [0, 8], ld)
    kw = dict(tolerance=1e-9, maxbonddim=40, maxiter=6, sweepstrategy=strategy)
    tci, ranks, errors = T.crossinterpolate2(f, ld, **kw, rng=T.CounterRNG(3))
    res = oracle.crossinterpolate2(o, ld, seed=3, **kw)
    compare_tci(tci, ranks, errors, res, T)


def test_crossinterpolate2_sepcos_matches_oracle(T, oracle):  # config 4 family, reduced
    ld = [8] * 5
    p = sepcos_params(5)
    f, o = T.BuiltinTarget(SEPCOS, p, ld), oracle.Target.builtin(SEPCOS, p, ld)
    kw = dict(tolerance=1e-10, maxbonddim=30, maxiter=5)
    tci, ranks, errors = T.crossinterpolate2(f, ld, **kw, rng=T.CounterRNG(4))
    res = oracle.crossinterpolate2(o, ld, seed=4, **kw)
    compare_tci(tci, ranks, errors, res, T)


@pytest.mark.parametrize("name", ["lorentz", "quantics2d"])
def test_crossinterpolate2_rook_matches_oracle(T, oracle, name):  # test_tensorci2.jl:247-340 with pivotsearch=:rook
    if name == "lorentz":
        ld, kw = [10] * 5, dict(tolerance=1e-10, maxiter=30)
        f, o = make_target(T, oracle, LORENTZ, [1.0], ld)
    else:
        ld, kw = [4] * 8, dict(tolerance=1e-9, maxbonddim=40, maxiter=6)
        f, o = make_target(T, oracle, Q2D, [0, 8], ld)
    tci, ranks, errors = T.crossinterpolate2(f, ld, pivotsearch="rook", rng=T.CounterRNG(5), **kw)
    res = oracle.crossinterpolate2(o, ld, pivotsearch="rook", seed=5, **kw)
    compare_tci(tci, ranks, errors, res, T)
    if name == "lorentz":
        for v in itertools.product(range(1, 4), repeat=5):
            fv = 1.0 / (1.0 + sum(x * x for x in v))
            assert abs(tci(list(v)) - fv) <= 1e-8 * fv


def test_crossinterpolate2_ttcache(T, oracle):  # test_tensorci2.jl:477-502
    rng = np.random.default_rng(4)
    dims = [2, 3, 3, 2]
    cores = _rand_tt(rng, [1, 2, 3, 2, 1], dims)
    f = T.TTCache(T.TensorTrain(cores))
    tci, ranks, errors = T.crossinterpolate2(f, dims, tolerance=1e-10, maxbonddim=10)
    full = T.fulltensor(T.TensorTrain(cores))
    rec = T.fulltensor(T.TensorTrain(tci.sitetensors))
    np.testing.assert_allclose(rec, full, rtol=1.5e-8, atol=1e-12)


def test_integration_10d_known_answer(T):  # test_integration.jl:61-70
    f = T.BuiltinTarget(GK, gk_params(), [15] * 10)
    tci, ranks, errors = T.crossinterpolate2(f, [15] * 10, tolerance=1e-8, nsearchglobalpivot=10)
    assert abs(T.tci_sum(tci) / 15.0**10 - G.INTEGRAL_10D_REF) < 1e-3


def test_argument_errors(T):  # test_tensorci2.jl:215-245
    f = T.BuiltinTarget(SUM, [], [2, 2, 2])
    with pytest.raises(ValueError, match="convergence criterion is not reachable"):
        T.crossinterpolate2(f, [2, 2, 2], tolerance=0.0)
    with pytest.raises(RuntimeError, match="nsearchglobalpivot < maxnglobalpivot!"):
        T.crossinterpolate2(f, [2, 2, 2], nsearchglobalpivot=2, maxnglobalpivot=5)
    with pytest.raises(RuntimeError, match="at least 2 elements"):
        T.TensorCI2(f, [2])
    z = T.BuiltinTarget(TABLE, np.zeros(4), [2, 2])
    with pytest.raises(RuntimeError, match="maxsamplevalue is zero!"):
        T.TensorCI2(z, [2, 2])


def _ngpu():
    import torch
    return torch.cuda.device_count()


USER_SRC = """
__device__ double tci_user_f(const long long *x, int n, const double *params)
{
    double s = 1.0;                       // f(v) = 1 / (1 + sum_k (w_k v_k)^2), accumulated left to right
    for (int k = 0; k < n; ++k) {
        const double v = (double)x[k] * params[k];
        s = s + v * v;
    }
    return 1.0 / s;
}
"""


def test_user_source_target_nvrtc(T, oracle):
    """A user-defined target compiled from CUDA source at run time (NVRTC, tci_target_source): the device route for an
    arbitrary f (batcheval.jl:32-61).  Pi / T tensors for every split, point evaluation and a whole crossinterpolate2
    against the same function evaluated on the host in the same operation order (bit-identical: --fmad=false), and
    against the built-in Lorentzian when all weights are one."""
    ld = [5, 4, 6, 3, 5]
    w = np.array([1.0, 0.5, 0.25, 2.0, 1.5])
    f = T.SourceTarget(USER_SRC, w, ld)

    def host(x):
        s = 1.0
        for k in range(len(ld)):
            v = float(x[k]) * w[k]
            s = s + v * v
        return 1.0 / s

    rng = np.random.default_rng(4)
    pts = rand_indexset(rng, ld, 40)
    assert np.array_equal(f.evaluate_points(pts), np.array([host(x) for x in pts]))
    for nl, M in ((0, 1), (1, 1), (2, 0), (1, 2), (3, 2), (4, 1)):
        nr = len(ld) - nl - M
        I, J = rand_indexset(rng, ld[:nl], 7), rand_indexset(rng, ld[nl + M:], 5)
        got = f(I, J, M)
        cd = ld[nl:nl + M]
        ref = np.zeros((len(I), *cd, len(J)), order="F")
        for i in range(len(I)):
            for c in itertools.product(*[range(1, d + 1) for d in cd]):
                for j in range(len(J)):
                    ref[(i, *[v - 1 for v in c], j)] = host(list(I[i]) + list(c) + list(J[j]))
        assert got.shape == ref.shape and np.array_equal(got, ref), (nl, M)
        dev, mx = f.batchevaluate_device(I, J, M)
        assert mx == np.max(np.abs(ref))
    # the same function as a built-in: identical pivots through the whole driver
    ld2 = [10] * 6
    g1 = T.SourceTarget(USER_SRC, np.ones(6), ld2)
    g2 = T.BuiltinTarget(LORENTZ, [1.0], ld2)
    ta, ra, ea = T.crossinterpolate2(g1, ld2, tolerance=1e-8, rng=T.CounterRNG(1))
    tb, rb, eb = T.crossinterpolate2(g2, ld2, tolerance=1e-8, rng=T.CounterRNG(1))
    assert ra == rb
    assert all(np.array_equal(x, y) for x, y in zip(ta.Iset + ta.Jset, tb.Iset + tb.Jset))
    assert abs(T.tci_sum(ta) - T.tci_sum(tb)) <= 1e-12 * abs(T.tci_sum(tb))
    with pytest.raises(ValueError, match="does not compile"):
        T.SourceTarget("__device__ double tci_user_f(const long long *x, int n, const double *p) { return y; }", [], ld)


def test_bond_update_matches_separate_calls_and_oracle(T, oracle):
    """tci_bond_update (Pi evaluation -> rrLU, one synchronisation) against tci_pi_eval + tci_rrlu and the oracle:
    identical permutations, pivot errors, max|Pi| and factors (tensorci2.jl:529-551)."""
    ld = [10] * 6
    f, o = make_target(T, oracle, LORENTZ, [1.0], ld)
    rng = np.random.default_rng(3)
    I, J = rand_indexset(rng, ld[:3], 300), rand_indexset(rng, ld[3:], 200)
    for leftorth in (True, False):
        lu, mx = f.bond_update(I, J, maxrank=40, reltol=1e-14, abstol=1e-12, leftorthogonal=leftorth, want_factors=True)
        dev, mx2 = f.batchevaluate_device(I, J, 0)
        lu2 = T.rrlu(dev, maxrank=40, reltol=1e-14, abstol=1e-12, leftorthogonal=leftorth)
        Pi, omx = o.pi_eval(I.tolist(), J.tolist(), 0, 0.0)
        ref = oracle.rrlu(Pi, maxrank=40, reltol=1e-14, abstol=1e-12, leftorthogonal=leftorth)
        assert mx == mx2 == omx
        assert_lu_equal(lu, ref)
        assert_lu_equal(lu2, ref)
        luci = T.MatrixLUCI(lu)
        rl = oracle.luci(Pi, maxrank=40, reltol=1e-14, abstol=1e-12, leftorthogonal=leftorth)
        assert np.max(np.abs(luci.left() - rl.left)) <= RTOL * np.max(np.abs(rl.left))
        assert np.max(np.abs(luci.right() - rl.right)) <= RTOL * np.max(np.abs(rl.right))
    lu, mx = f.bond_update(I, J, maxrank=5)  # without factors: permutations only
    assert lu.npivot == 5
    with pytest.raises(RuntimeError, match="without factors"):
        lu.L
    with pytest.raises(ValueError, match="rows must not be empty"):
        f.bond_update(I[:0], J)


def test_fill_sitetensors_matches_per_site_path(T, oracle):
    """tci_fill_sitetensors (all sites queued back to back, one synchronisation, cores left on the device) against
    the per-site path (tci_pi_eval x2 + tci_rrlu + tci_lu_rdiv per site) and the oracle's site tensors."""
    from tci_b200 import tensorci2 as M
    ld = [4] * 8
    f, o = make_target(T, oracle, Q2D, [0, 8], ld)
    tci, ranks, errors = T.crossinterpolate2(f, ld, tolerance=1e-9, maxbonddim=30, maxiter=4, rng=T.CounterRNG(2))
    M.sweep2site(tci, f, 2, abstol=1e-9 * tci.maxsamplevalue, maxbonddim=30, fillsitetensors_=False)
    mx0 = tci.maxsamplevalue
    M.fillsitetensors(tci, f)
    fused = [t.copy() for t in tci.sitetensors]
    handle, mx1 = tci.device_tt, tci.maxsamplevalue
    assert handle is not None
    tci.maxsamplevalue = mx0
    for b in range(len(ld)):
        M.setsitetensor_fill(tci, f, b)
    assert tci.maxsamplevalue == mx1
    for a, b_ in zip(fused, tci.sitetensors):
        assert a.shape == b_.shape and np.array_equal(a, b_)
    # the device-resident copy is the same tensor train: evaluate it as a TT target
    pts = rand_indexset(np.random.default_rng(5), ld, 64)
    assert np.max(np.abs(handle.evaluate_points(pts) - T.evaluate_points(T.TensorTrain(fused), pts))) <= \
        1e-12 * tci.maxsamplevalue
    # error texts
    bad = [s.copy() for s in tci.Jset]
    bad[2] = bad[2][:-1]
    with pytest.raises(RuntimeError, match="Pivot matrix at bond 3 is not square!"):
        f.fill_sitetensors(tci.Iset, bad)
    dupI = [s.copy() for s in tci.Iset]
    dupJ = [s.copy() for s in tci.Jset]
    if len(dupI[3]) >= 2:
        dupI[3][1] = dupI[3][0]  # two equal rows: P at bond 3 is singular
        with pytest.raises(RuntimeError, match="Pivot matrix at bond 3 is singular!"):
            f.fill_sitetensors(dupI, dupJ)


@pytest.mark.parametrize("name", ["config1", "config3_small", "config4_small"])
def test_globalsearch_environment_mode_matches_ordered_chain(T, oracle, name):
    """K7 mode 2 (prefix / suffix environments of the start points + one GEMM per site) against mode 1 (every probe
    through the ordered chain, bit-identical to evaluate(tt, x)) and the oracle: identical pivots, errors within
    1e-12 of the function scale (globalpivotfinder.jl:160-188)."""
    if name == "config1":
        kind, params, ld, kw = LORENTZ, [1.0], [10] * 8, dict(tolerance=1e-3, maxiter=2)
    elif name == "config3_small":
        kind, params, ld, kw = Q2D, [0, 10], [4] * 10, dict(tolerance=1e-3, maxbonddim=12, maxiter=2)
    else:
        kind, params, ld, kw = SEPCOS, sepcos_params(6), [64] * 6, dict(tolerance=1e-2, maxbonddim=6, maxiter=2)
    f, o = make_target(T, oracle, kind, params, ld)
    tci, ranks, errors = T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
    tt = T.TensorTrain(tci.sitetensors)
    nsearch = 300
    finder = T.DefaultGlobalPivotFinder(nsearch=nsearch, maxnglobalpivot=nsearch, tolmarginglobalsearch=1.0)
    inp = T.GlobalPivotSearchInput(ld, tt, tci.maxsamplevalue, None, None)
    abstol = 1e-6 * tci.maxsamplevalue
    p1 = finder(inp, f, abstol, rng=T.CounterRNG(9), mode=1)
    e1 = finder.last_errors.copy()
    p2 = finder(inp, f, abstol, rng=T.CounterRNG(9), mode=2)
    e2 = finder.last_errors.copy()
    starts = T.CounterRNG(9).start_points(nsearch, ld)
    piv, errs = oracle.globalsearch(o, tci.sitetensors, np.ascontiguousarray(starts.T), abstol=abstol, tolmargin=1.0,
                                    maxn=nsearch)
    assert len(p1) > 0
    assert p1.tolist() == piv and np.array_equal(e1, errs)  # ordered chain: bit-identical to the oracle
    assert p2.tolist() == p1.tolist()
    assert np.max(np.abs(e2 - e1)) <= 1e-12 * tci.maxsamplevalue


def test_two_contexts_on_two_devices_in_one_process(T, oracle):
    """Kernel attributes (dynamic shared memory opt-in) are per device: a second context on another GPU of the same
    process must run the 226 KB rrLU kernels and the DMMA GEMM too."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    A = lowrank_matrix(700, 600, 48, seed=21)
    ref = oracle.rrlu(A, maxrank=48, reltol=1e-12)
    B1, B2 = np.random.default_rng(1).standard_normal((300, 200)), np.random.default_rng(2).standard_normal((200, 150))
    for dev in (0, 1):
        ctx = T.Context(dev)
        assert_lu_equal(T.rrlu(A, maxrank=48, reltol=1e-12, ctx=ctx), ref)
        from tci_b200._lib import gemm
        assert np.max(np.abs(gemm(B1, B2, ctx) - B1 @ B2)) <= 1e-12 * np.max(np.abs(B1 @ B2))


def test_multigpu_context_sharded_stages_match_single_gpu(T, oracle, monkeypatch):
    """tci_ctx_create(ngpu = 2): Pi evaluation (column blocks, peer stores into the owner's HBM), the MPO x MPO Pi by
    row blocks with both environment chains sharded, and the global search by blocks of starts -- all inside the
    library -- against the single-GPU results: analytic Pi bit-identical, contraction Pi to 1e-12, identical pivots;
    a whole crossinterpolate2 on the 2-GPU context gives identical index sets."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    c1, c2 = T.Context(0), T.Context(devices=[0, 1])
    assert c2.ngpu == 2
    monkeypatch.setenv("TCI_SHARD_FORCE", "2")  # the cost model would keep these small cases on one GPU
    ld = [10] * 6
    rng = np.random.default_rng(0)
    I, J = rand_indexset(rng, ld[:3], 333), rand_indexset(rng, ld[3:], 201)
    for kind, params, l in ((LORENTZ, [1.0], ld), (SEPCOS, sepcos_params(6), [64] * 6)):
        I, J = rand_indexset(rng, l[:3], 333), rand_indexset(rng, l[3:], 201)
        f1, f2 = T.BuiltinTarget(kind, params, l, ctx=c1), T.BuiltinTarget(kind, params, l, ctx=c2)
        l0 = c2.member_launches(1)
        d1, m1 = f1.batchevaluate_device(I, J, 0)
        d2, m2 = f2.batchevaluate_device(I, J, 0)
        assert c2.member_launches(1) > l0  # the second GPU really took part
        assert np.array_equal(d1.to_host(), d2.to_host()) and m1 == m2
        lu1, mx1 = f1.bond_update(I, J, maxrank=30, want_factors=True)
        lu2, mx2 = f2.bond_update(I, J, maxrank=30, want_factors=True)
        assert mx1 == mx2 and np.array_equal(lu1.rowpermutation, lu2.rowpermutation)
        assert np.array_equal(lu1.L, lu2.L) and np.array_equal(lu1.U, lu2.U)
    # contraction target: row blocks
    g = np.random.default_rng(5)
    ns, D = 10, 12
    bonds = [1] + [D] * (ns - 1) + [1]
    A = [np.asfortranarray(g.random((bonds[i], 2, 2, bonds[i + 1])) - 0.5) for i in range(ns)]
    B = [np.asfortranarray(g.random((bonds[i], 2, 2, bonds[i + 1])) - 0.5) for i in range(ns)]
    m1 = T.Contraction(T.TensorTrain(A), T.TensorTrain(B), ctx=c1)
    m2 = T.Contraction(T.TensorTrain(A), T.TensorTrain(B), ctx=c2)
    Im, Jm = rand_indexset(g, [4] * 5, 150), rand_indexset(g, [4] * 5, 90)
    P1, P2 = m1(Im, Jm, 0), m2(Im, Jm, 0)
    assert np.max(np.abs(P1 - P2)) <= 1e-12 * np.max(np.abs(P1))
    ref, _ = oracle.Target.mpo_pair(A, B).pi_eval(Im[:6].tolist(), Jm[:5].tolist(), 0, 0.0)
    assert np.max(np.abs(P2[:6, :5] - np.asarray(ref).reshape((6, 5), order="F"))) <= RTOL * np.max(np.abs(P1))
    # global search: blocks of starts, records gathered by NCCL
    monkeypatch.delenv("TCI_SHARD_FORCE")
    l4 = [64] * 6
    f1, f2 = T.BuiltinTarget(SEPCOS, sepcos_params(6), l4, ctx=c1), T.BuiltinTarget(SEPCOS, sepcos_params(6), l4, ctx=c2)
    tci, ranks, errors = T.crossinterpolate2(f1, l4, tolerance=1e-2, maxbonddim=6, maxiter=2, rng=T.CounterRNG(1))
    tt = T.TensorTrain(tci.sitetensors)
    finder = T.DefaultGlobalPivotFinder(nsearch=512, maxnglobalpivot=512, tolmarginglobalsearch=1.0)
    inp = T.GlobalPivotSearchInput(l4, tt, tci.maxsamplevalue, None, None)
    for mode in (1, 2):
        pa = finder(inp, f1, 1e-7, rng=T.CounterRNG(9), mode=mode)
        ea = finder.last_errors.copy()
        l0 = c2.member_launches(1)
        pb = finder(inp, f2, 1e-7, rng=T.CounterRNG(9), mode=mode)
        assert c2.member_launches(1) > l0
        assert len(pa) > 0 and pa.tolist() == pb.tolist() and np.array_equal(ea, finder.last_errors)
    # whole runs on the 2-GPU context
    monkeypatch.setenv("TCI_SHARD_FORCE", "2")
    for kind, params, l, kw in ((LORENTZ, [1.0], [10] * 6, dict(tolerance=1e-8)),
                                (Q2D, [0, 8], [4] * 8, dict(tolerance=1e-9, maxbonddim=40, maxiter=6))):
        fa, fb = T.BuiltinTarget(kind, params, l, ctx=c1), T.BuiltinTarget(kind, params, l, ctx=c2)
        ta, ra, ea = T.crossinterpolate2(fa, l, rng=T.CounterRNG(3), **kw)
        tb, rb, eb = T.crossinterpolate2(fb, l, rng=T.CounterRNG(3), **kw)
        assert ra == rb and ea == eb
        assert all(np.array_equal(x, y) for x, y in zip(ta.Iset + ta.Jset, tb.Iset + tb.Jset))
        assert all(np.array_equal(x, y) for x, y in zip(ta.sitetensors, tb.sitetensors))


def test_config4_scale_pi_and_rrlu_vs_oracle_on_pivot_submatrix(T, oracle):
    """BASELINE config 4 at its full size: 12 sites, d = 64, Pi = 32768 x 32768 (8.6 GB, stays in HBM),
    config-4 target, rrLU truncated at 16 pivots.  Size-independent check: full pivoting on Pi and on any
    order-preserving submatrix S that contains all chosen pivot rows and columns must pick the same pivots with
    bit-identical values (the Schur updates of S only involve pivot rows / columns, all inside S).  S is
    evaluated and factorised by the CPU oracle."""
    ld = [64] * 12
    p = sepcos_params(12)
    f, o = T.BuiltinTarget(SEPCOS, p, ld), oracle.Target.builtin(SEPCOS, p, ld)
    rng = np.random.default_rng(44)
    chi = 512

    def distinct(count, length):
        seen, out = set(), []
        while len(out) < count:
            v = tuple(int(x) for x in rng.integers(1, 65, length))
            if v not in seen:
                seen.add(v)
                out.append(v)
        return np.array(out, dtype=np.int64)

    I = T.kronecker_left(distinct(chi, 5), 64)   # 32768 x 6
    J = T.kronecker_right(64, distinct(chi, 5))  # 32768 x 6
    assert I.shape == (32768, 6) and J.shape == (32768, 6)
    dev, mx = f.batchevaluate_device(I, J, 0)
    assert dev.shape == (32768, 32768)
    r = 16
    lu = T.rrlu(dev, maxrank=r, reltol=1e-14)
    assert lu.npivot == r
    prow, pcol = T.rowindices(lu) - 1, T.colindices(lu) - 1
    rows = np.array(sorted(set(prow.tolist()) | set(rng.integers(0, 32768, 150).tolist())))
    cols = np.array(sorted(set(pcol.tolist()) | set(rng.integers(0, 32768, 150).tolist())))
    S, omx = o.pi_eval(I[rows].tolist(), J[cols].tolist(), 0, 0.0)
    ref = oracle.rrlu(S, maxrank=r, reltol=1e-14)
    assert rows[ref.rowpermutation[:r] - 1].tolist() == prow.tolist()
    assert cols[ref.colpermutation[:r] - 1].tolist() == pcol.tolist()
    assert np.array_equal(T.pivoterrors(lu)[:r], ref.pivoterrors[:r])
    assert mx >= omx  # the fused max-abs covers all of Pi, the oracle's only S


# ---------------------------------------------------------------- edge cases ----
@pytest.mark.parametrize("name", ["config1", "config3", "config4"])
def test_crossinterpolate2_full_size_matches_oracle(T, oracle, name):
    """BASELINE configs 1, 3 (fused: 20 sites d=4, R=20 bits per dimension, maxbonddim 256, tolerance 1e-10) and 4
    (12 sites d=64, maxbonddim 512, tolerance 1e-12) at their FULL sizes against the CPU oracle (0.01 s, ~3 s, ~11 s
    on the box): identical ranks per iteration and identical Iset / Jset at every bond, site tensors, sum(tci) and
    the sampled interpolation error within 1e-10 (measured: <= 5e-15, tools/fullsize_parity.py)."""
    if name == "config1":
        kind, params, ld, kw = LORENTZ, [1.0], [10] * 8, dict(tolerance=1e-8)
    elif name == "config3":
        kind, params, ld, kw = Q2D, [0, 20], [4] * 20, dict(tolerance=1e-10, maxbonddim=256)
    else:
        kind, params, ld, kw = SEPCOS, sepcos_params(12), [64] * 12, dict(tolerance=1e-12, maxbonddim=512)
    f, o = T.BuiltinTarget(kind, params, ld), oracle.Target.builtin(kind, params, ld)
    tci, ranks, errors = T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
    res = oracle.crossinterpolate2(o, ld, seed=1, **kw)
    compare_tci(tci, ranks, errors, res, T)
    rng = np.random.default_rng(0)
    pts = rand_indexset(rng, ld, 500)
    fv = f.evaluate_points(pts)
    tg = T.evaluate_points(T.TensorTrain(tci.sitetensors), pts)
    to = np.array([oracle.tt_evaluate(res.sitetensors, q) for q in pts])
    scale = tci.maxsamplevalue
    assert np.max(np.abs(tg - to)) <= RTOL * scale
    assert abs(np.max(np.abs(fv - tg)) - np.max(np.abs(fv - to))) <= RTOL * scale


def test_crossinterpolate2_through_the_deferred_update_kernel(T, oracle, monkeypatch):
    """A whole crossinterpolate2 in which EVERY bond factorisation is forced through the deferred-update streaming
    kernel (csrc/rrlu_lazy.cu; in production it takes over from ~3600^2, which the low-rank BASELINE targets never
    reach end to end): ranks per iteration, index sets and site tensors against the oracle."""
    monkeypatch.setenv("TCI_RRLU_NO_RES", "1")
    monkeypatch.setenv("TCI_RRLU_LAZY_MIN", "0")
    for kind, params, ld, kw in ((Q2D, [0, 8], [4] * 8, dict(tolerance=1e-9, maxbonddim=40, maxiter=6)),
                                 (SEPCOS, sepcos_params(5), [64] * 5, dict(tolerance=1e-10, maxbonddim=24))):
        f, o = T.BuiltinTarget(kind, params, ld), oracle.Target.builtin(kind, params, ld)
        tci, ranks, errors = T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
        res = oracle.crossinterpolate2(o, ld, seed=1, **kw)
        compare_tci(tci, ranks, errors, res, T)


def test_crossinterpolate2_tt_target_rank_reaches_maxbonddim(T, oracle):
    """A target whose rank really reaches maxbonddim (a random tensor train of bond dimension 24 on 6 sites, d = 16):
    every middle bond factorises a (24*16)^2 Pi to its full allowed rank; pivots against the oracle."""
    g = np.random.default_rng(12)
    bonds = [1] + [24] * 5 + [1]
    cores = [np.asfortranarray((g.random((bonds[i], 16, bonds[i + 1])) * 2 - 1) / np.sqrt(bonds[i] * 4.0)) for i in range(6)]
    ld = [16] * 6
    f, o = T.TTCache(T.TensorTrain(cores)), oracle.Target.tt(cores)
    kw = dict(tolerance=1e-10, maxbonddim=24, maxiter=4)
    tci, ranks, errors = T.crossinterpolate2(f, ld, rng=T.CounterRNG(1), **kw)
    res = oracle.crossinterpolate2(o, ld, seed=1, **kw)
    assert int(ranks[-1]) == 24
    assert [int(r) for r in ranks] == res.ranks.tolist()
    pts = rand_indexset(g, ld, 300)
    assert np.max(np.abs(f.evaluate_points(pts) - T.evaluate_points(T.TensorTrain(tci.sitetensors), pts))) <= \
        1e-8 * tci.maxsamplevalue


def test_config5_full_shape_mpo_properties(T):
    """BASELINE config 5 at its full shape (40 sites, bond dimension 256, site dimensions 2 x 2): the oracle does not
    finish this size in seconds, so parity goes through size-independent properties.  (i) the three evaluation
    paths agree to 1e-10 of max|Pi|: tci_pi_eval (M = 0), separately evaluated environments + tci_pi_from_envs, and
    the pointwise evaluate with its mid-chain split (contraction.jl:189-207); (ii) linearity: doubling one core of A
    doubles Pi bit for bit; (iii) the M = 2 Pi equals the M = 0 Pi over the kronecker-expanded index sets
    (tensorci2.jl:315-327 order: left index fastest in the rows, site index fastest in the columns)."""
    rng = np.random.default_rng(55)
    ns, D, n = 40, 256, 48
    bonds = [1] + [D] * (ns - 1) + [1]
    A = [np.asfortranarray((rng.random((bonds[i], 2, 2, bonds[i + 1])) * 2 - 1) / 16.0) for i in range(ns)]
    B = [np.asfortranarray((rng.random((bonds[i], 2, 2, bonds[i + 1])) * 2 - 1) / 16.0) for i in range(ns)]
    f = T.Contraction(T.TensorTrain(A), T.TensorTrain(B))
    I = rand_indexset(rng, [4] * 20, n)
    J = rand_indexset(rng, [4] * 20, n)
    Pi = f(I, J, 0)
    scale = np.max(np.abs(Pi))
    assert np.all(np.isfinite(Pi)) and scale > 0
    # (i) environments + product, and pointwise evaluation
    Dl = f.env_dim(0, 20)
    assert Dl == D * D == f.env_dim(1, 20)
    lenv, renv = T.DeviceMatrix.empty(f.ctx, Dl, n), T.DeviceMatrix.empty(f.ctx, Dl, n)
    f.env_eval_into(lenv, 0, 0, I)
    f.env_eval_into(renv, 0, 1, J)
    out = T.DeviceMatrix.empty(f.ctx, n, n)
    mx = f.pi_from_envs(lenv, 0, n, renv, 0, n, out, 0)
    assert np.max(np.abs(out.to_host() - Pi)) <= RTOL * scale and abs(mx - scale) <= RTOL * scale
    del lenv, renv, out
    pts = np.array([np.concatenate([I[i], J[j]]) for j in range(8) for i in range(8)], dtype=np.int64)
    vals = f.evaluate_points(pts).reshape((8, 8), order="F")
    assert np.max(np.abs(vals - Pi[:8, :8])) <= RTOL * scale
    # (ii) linearity in one core (a power of two commutes with every rounding)
    A2 = [a.copy(order="F") for a in A]
    A2[7] *= 2.0
    f2 = T.Contraction(T.TensorTrain(A2), T.TensorTrain(B))
    assert np.array_equal(f2(I, J, 0), 2.0 * Pi)
    del f2
    # (iii) M = 2 against M = 0 on the expanded sets
    I19, J19 = I[:16, :19], J[:16, 1:]
    Pi2 = f(I19, J19, 2)  # (16, 4, 4, 16)
    assert Pi2.shape == (16, 4, 4, 16)
    Ie = T.kronecker_left(I19, 4)
    Je = T.kronecker_right(4, J19)
    Pi0 = f(Ie, Je, 0)
    assert np.max(np.abs(Pi0 - Pi2.reshape((64, 64), order="F"))) <= RTOL * np.max(np.abs(Pi0))


def test_config5_full_shape_mpo_vs_oracle(T, oracle):
    """BASELINE config 5 at its full shape (40 sites, bond dimension 256, site dimensions 2 x 2) against the ORACLE
    (contraction.jl:112-207, 236-335 restated): an 8 x 8 block of (left, right) index pairs costs the oracle ~4e10
    flop (~30 s).  The 8 rows / 8 columns are taken from NESTED index sets of 1024 entries, as a TCI run produces
    them, so the same oracle block checks (i) tci_pi_eval on the 8 x 8 sets, (ii) separately evaluated environments +
    tci_pi_from_envs, (iii) the 1024 x 1024 Pi over the nested sets, whose shared prefixes are evaluated once
    (ChainPlan, csrc/mpo.cu), and (iv) the same Pi with prefix sharing switched off."""
    import os
    rng = np.random.default_rng(5)
    ns, D = 40, 256
    bonds = [1] + [D] * (ns - 1) + [1]
    A = [np.asfortranarray((rng.random((bonds[i], 2, 2, bonds[i + 1])) * 2 - 1) / 16.0) for i in range(ns)]
    B = [np.asfortranarray((rng.random((bonds[i], 2, 2, bonds[i + 1])) * 2 - 1) / 16.0) for i in range(ns)]
    f = T.Contraction(T.TensorTrain(A), T.TensorTrain(B))
    Sl = np.arange(1, 5, dtype=np.int64)[:, None]
    Sr = Sl.copy()
    for _ in range(18):
        cl, cr = T.kronecker_left(Sl, 4), T.kronecker_right(4, Sr)
        Sl = cl[np.sort(rng.choice(len(cl), min(256, len(cl)), replace=False))]
        Sr = cr[np.sort(rng.choice(len(cr), min(256, len(cr)), replace=False))]
    In, Jn = T.kronecker_left(Sl, 4), T.kronecker_right(4, Sr)
    assert In.shape == (1024, 20) and Jn.shape == (1024, 20)
    ri = np.sort(rng.choice(1024, 8, replace=False))
    ci = np.sort(rng.choice(1024, 8, replace=False))
    I8, J8 = np.ascontiguousarray(In[ri]), np.ascontiguousarray(Jn[ci])
    o = oracle.Target.mpo_pair(A, B)
    ref, omx = o.pi_eval(I8.tolist(), J8.tolist(), 0, 0.0)
    ref = np.asarray(ref).reshape((8, 8), order="F")
    scale = np.max(np.abs(ref))
    assert scale > 0 and omx == scale
    # (i) the direct call
    Pi8 = f(I8, J8, 0)
    assert np.max(np.abs(Pi8 - ref)) <= RTOL * scale
    # (ii) environments + product
    Dl = f.env_dim(0, 20)
    lenv, renv = T.DeviceMatrix.empty(f.ctx, Dl, 8), T.DeviceMatrix.empty(f.ctx, Dl, 8)
    f.env_eval_into(lenv, 0, 0, I8)
    f.env_eval_into(renv, 0, 1, J8)
    out = T.DeviceMatrix.empty(f.ctx, 8, 8)
    mx = f.pi_from_envs(lenv, 0, 8, renv, 0, 8, out, 0)
    assert np.max(np.abs(out.to_host() - ref)) <= RTOL * scale and abs(mx - scale) <= RTOL * scale
    del lenv, renv, out
    # (iii) nested sets, shared prefixes evaluated once; (iv) without prefix sharing
    full = f(In, Jn, 0)
    assert np.max(np.abs(full[np.ix_(ri, ci)] - ref)) <= RTOL * scale
    os.environ["TCI_MPO_NO_DEDUP"] = "1"
    try:
        plain = f(In, Jn, 0)
    finally:
        del os.environ["TCI_MPO_NO_DEDUP"]
    assert np.max(np.abs(plain[np.ix_(ri, ci)] - ref)) <= RTOL * scale
    assert np.max(np.abs(plain - full)) <= RTOL * np.max(np.abs(plain))


def test_rrlu_degenerate_shapes(T, oracle):
    """Empty, single-row / single-column and maxrank corner cases (matrixlu.jl:141-181)."""
    for shape in [(0, 5), (5, 0), (0, 0)]:
        lu = T.rrlu(np.zeros(shape))
        assert lu.npivot == 0 and lu.error == 0.0 and lu.L.shape == (shape[0], 0) and lu.U.shape == (0, shape[1])
        assert T.pivoterrors(lu).tolist() == [0.0]
    for A in [np.array([[3.0]]), np.array([[1.0, -4.0, 2.0]]), np.array([[1.0], [-4.0], [2.0]]),
              np.array([[0.0, 0.0], [0.0, 5.0]])]:
        for lo in (True, False):
            assert_lu_equal(T.rrlu(A, leftorthogonal=lo), oracle.rrlu(A, leftorthogonal=lo))
    A = np.random.default_rng(1).random((6, 9))
    assert_lu_equal(T.rrlu(A, maxrank=100), oracle.rrlu(A, maxrank=100))  # maxrank above min(m, n)
    assert_lu_equal(T.rrlu(A, maxrank=1), oracle.rrlu(A, maxrank=1))
    assert_lu_equal(T.rrlu(A, reltol=0.9), oracle.rrlu(A, reltol=0.9))  # stops after the first pivots
    assert_lu_equal(T.rrlu(A, abstol=10.0), oracle.rrlu(A, abstol=10.0))  # at least one pivot is always taken
    with pytest.raises(ValueError):
        T.rrlu(A, maxrank=0)
    with pytest.raises(ValueError):
        T.rrlu(np.zeros(3))


def test_rrlu_inf_and_huge_values(T, oracle):
    """abs2 overflows to Inf for |x| > 1.3e154: ties on Inf are broken by scan order (Appendix A.4)."""
    A = np.random.default_rng(2).random((12, 10))
    A[3, 4] = 1e200
    A[7, 2] = -1e200
    lu, ref = T.rrlu(A, maxrank=3), oracle.rrlu(A, maxrank=3)
    assert_lu_equal(lu, ref)
    assert (lu.rowpermutation[0], lu.colpermutation[0]) == (8, 3)  # column 3 is scanned before column 5
    B = A * 1e-300  # squares underflow to 0 except the two large entries
    assert_lu_equal(T.rrlu(B, maxrank=4), oracle.rrlu(B, maxrank=4))


def test_pi_eval_single_point_and_full_tensor(T, oracle):
    """nl = nr = 0 (the whole tensor as one Pi) and 1 x 1 requests."""
    ld = [3, 4, 2]
    f, o = make_target(T, oracle, TABLE, "table", ld)
    full = f(np.zeros((1, 0), dtype=np.int64), np.zeros((1, 0), dtype=np.int64), 3)
    assert full.shape == (1, 3, 4, 2, 1)
    table = np.random.default_rng(11).standard_normal(24).reshape(ld, order="F")
    assert np.array_equal(full[0, ..., 0], table)
    one = f([[2, 3]], [[1]], 0)
    assert one.shape == (1, 1) and one[0, 0] == table[1, 2, 0] == f([2, 3, 1])


def test_context_reports_launches_and_timers(T, ctx):
    l0 = ctx.launches
    T.rrlu(np.eye(3))
    assert ctx.launches > l0
    tm = ctx.timers()
    assert set(tm) >= {"pi_eval", "rrlu", "luci", "rrlu_kernel"} and tm["rrlu_kernel"] > 0.0


@pytest.mark.parametrize("strictlynested", [False, True])
def test_sweep2site_half_fused_equals_per_bond(T, strictlynested):
    """tci_sweep2site_half (the bond loop of a half-sweep in one library call) against the same loop made of
    tci_bond_update calls from the host mirror's updatepivots: identical index sets, bond / pivot errors,
    maxsamplevalue and ranks after every optimize iteration (tensorci2.jl:855-916)."""
    ld = [6, 5, 7, 4, 6, 5]
    runs = []
    for per_bond in (False, True):
        f = T.BuiltinTarget(LORENTZ, [0.7], ld)
        tci = T.TensorCI2(f, ld, [[1] * 6, [3, 2, 5, 1, 4, 2]])
        tci.per_bond_calls = per_bond
        tci.trace = []
        ranks, errors = T.optimize(tci, f, tolerance=1e-9, maxiter=4, strictlynested=strictlynested, rng=T.CounterRNG(4))
        runs.append((tci, ranks, errors))
    (a, ra, ea), (b, rb, eb) = runs
    assert ra == rb and ea == eb and a.maxsamplevalue == b.maxsamplevalue
    assert np.array_equal(a.bonderrors, b.bonderrors) and np.array_equal(a.pivoterrors, b.pivoterrors)
    assert a.trace == b.trace
    for s in range(len(ld)):
        assert np.array_equal(a.Iset[s], b.Iset[s]) and np.array_equal(a.Jset[s], b.Jset[s])
        assert np.array_equal(a.sitetensors[s], b.sitetensors[s])
