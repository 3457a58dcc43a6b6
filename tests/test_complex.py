"""ComplexF64 value type.

CPU part: the oracle's C++ restatement of matrixlu.jl on Matrix{ComplexF64} (orc_zrrlu, Julia Base's complex arithmetic
of include/tci_zarith.h) pinned on the reference's complex arg-max literal (test_matrixlu.jl:39-52), on the structure of
its rrLU testsets (:54-211) and against the independent numpy restatement.  GPU part: tci_zrrlu bit-identical to that
oracle; left / right of MatrixLUCI, the complex Pi of a Contraction for every split, the complex GEMM and the
reference's complex contraction tests (test_contraction.jl:68-195) through the C ABI.  Tolerances: bit-exact for
permutations / pivot errors / L / U; 1e-10 relative for everything that is BLAS in the reference (pinned there to
sqrt(eps) only)."""
import numpy as np
import pytest

RTOL = 1e-10

ARGMAX_LITERAL = np.array([[0, 1, 2, 3, 4, 5],
                           [1, 1 + 1j, 2 + 1j, 3 + 1j, 4 + 1j, 5 + 1j],
                           [1, 1 + 2j, 2 + 2j, 3 + 2j, 4 + 2j, 5 + 2j]], dtype=np.complex128)  # test_matrixlu.jl:40-44


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def lowrank(rng, m, n, r, decay=30.0):
    s = 2.0 ** (-decay * np.arange(r) / max(r, 1))
    return (crand(rng, m, r) * s) @ crand(rng, r, n)


# ---- CPU: the oracle ---------------------------------------------------------------------------------------------
def test_oracle_zrrlu_first_pivot_is_reference_argmax(oracle):
    """submatrixargmax(abs2, A, :, :) == Tuple(argmax(abs2.(A))) (test_matrixlu.jl:47): the first pivot of the C++
    restatement is that element, for the reference's literal."""
    A = ARGMAX_LITERAL
    ref = np.unravel_index(np.argmax((A.real ** 2 + A.imag ** 2).flatten(order="F")), A.shape, order="F")
    for lo in (True, False):
        lu = oracle.zrrlu(A, leftorthogonal=lo)
        assert (lu.rowpermutation[0], lu.colpermutation[0]) == (ref[0] + 1, ref[1] + 1)
        assert oracle.submatrixargmax_abs2_complex(A) == (ref[0] + 1, ref[1] + 1)
        np.testing.assert_allclose(lu.L @ lu.U, A[lu.rowpermutation - 1][:, lu.colpermutation - 1], atol=1e-14)


@pytest.mark.parametrize("lo", [True, False])
def test_oracle_zrrlu_matches_numpy_restatement(oracle, lo):
    """Two independent restatements of matrixlu.jl:98-181 on complex data -- C++ with Julia's robust division, numpy
    with Python's Smith division -- pick identical pivots; factors agree to rounding; the testset structure of
    test_matrixlu.jl:54-211 (L*U reproduces the permuted matrix, rank-3 data stops at 3 pivots, pivoterrors)."""
    rng = np.random.default_rng(21)
    for (m, n) in ((9, 7), (6, 11), (8, 8)):
        B = crand(rng, m, n)
        lu = oracle.zrrlu(B, leftorthogonal=lo)
        rp, cp, L, U, r, err = oracle.rrlu_complex(B, leftorthogonal=lo)
        assert lu.npivot == r == min(m, n) and lu.error == err == 0.0
        assert np.array_equal(lu.rowpermutation, rp) and np.array_equal(lu.colpermutation, cp)
        np.testing.assert_allclose(lu.L, L, rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(lu.U, U, rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(lu.L @ lu.U, B[rp - 1][:, cp - 1], rtol=1e-13, atol=1e-13)
        diag = np.abs(np.diag(lu.U if lo else lu.L))
        assert np.array_equal(lu.pivoterrors[:-1], diag[:r]) or np.allclose(lu.pivoterrors[:-1], diag[:r], rtol=1e-15)
    A = crand(rng, 10, 3) @ crand(rng, 3, 12)  # test_matrixlu.jl:142-165 structure
    lu = oracle.zrrlu(A, reltol=1e-10, leftorthogonal=lo)
    assert lu.npivot == 3 and lu.error < 1e-8 * np.max(np.abs(A))
    lu = oracle.zrrlu(A, maxrank=2, leftorthogonal=lo)  # maxrank stop: error = last accepted pivot
    assert lu.npivot == 2 and lu.error == lu.pivoterrors[1]
    eye = np.eye(2, dtype=np.complex128) * 1j  # test_matrixlu.jl:167-175 with a phase: |pivots| = [1, 1], error 0
    assert oracle.zrrlu(eye, leftorthogonal=lo).pivoterrors.tolist() == [1.0, 1.0, 0.0]
    R = rng.standard_normal((8, 6))  # real data through the complex path: the Float64 oracle's pivots and factors
    ref = oracle.rrlu(R, leftorthogonal=lo)
    lu = oracle.zrrlu(R, leftorthogonal=lo)
    assert np.array_equal(lu.rowpermutation, ref.rowpermutation) and np.array_equal(lu.colpermutation, ref.colpermutation)
    # (not bit-equal: Julia's Complex(a, 0) / Complex(c, 0) evaluates a * (1 / c), one rounding more than a / c)
    np.testing.assert_allclose(lu.L.real, ref.L, rtol=1e-14, atol=0)
    assert np.max(np.abs(lu.L.imag)) == 0.0 and np.max(np.abs(lu.U.imag)) == 0.0


def test_oracle_zarith_division_and_hypot(oracle):
    """include/tci_zarith.h against numpy on ordinary operands (1 ulp) and on the operands the robust algorithm exists
    for (no spurious overflow / underflow: test cases of Baudin & Smith, arXiv:1210.4539, section 5)."""
    rng = np.random.default_rng(2)
    for _ in range(200):
        a, b = complex(*rng.standard_normal(2)), complex(*rng.standard_normal(2))
        lu = oracle.zrrlu(np.array([[b]], dtype=np.complex128))
        assert lu.pivoterrors[0] == pytest.approx(abs(b), rel=4e-16)
        lu = oracle.zrrlu(np.array([[b], [a]], dtype=np.complex128) if abs(b) >= abs(a) else
                          np.array([[a], [b]], dtype=np.complex128))
        q = lu.L[1, 0]
        ref = (a / b) if abs(b) >= abs(a) else (b / a)
        assert abs(q - ref) <= 4e-16 * abs(ref)
    big = np.array([[2.0 ** 1023 * (1 + 1j)], [2.0 ** 1023]], dtype=np.complex128)  # (1)/(1+i) scaled to the overflow edge
    lu = oracle.zrrlu(big)
    assert lu.L[1, 0] == pytest.approx(0.5 - 0.5j, rel=1e-15) and lu.pivoterrors[0] == pytest.approx(2.0 ** 1023 * np.sqrt(2))
    tiny = np.array([[2.0 ** -1074 * 3j], [2.0 ** -1074]], dtype=np.complex128)
    assert oracle.zrrlu(tiny).L[1, 0] == pytest.approx(-1j / 3, rel=1e-15)


@pytest.mark.parametrize("lo", [True, False])
def test_oracle_zluci_interpolation_properties(oracle, lo):
    """MatrixLUCI on complex data (test_matrixluci.jl:7-46 restated on the oracle): A ~ left * right, the pivot rows /
    columns are reproduced exactly (right[:, J] = A[I, J] when leftorthogonal, left[I, :] = A[I, J] otherwise), and
    left * right is exact when the rank is reached."""
    rng = np.random.default_rng(9)
    A = crand(rng, 12, 4) @ crand(rng, 4, 15)  # rank 4
    ref = oracle.zluci(A, reltol=1e-10, leftorthogonal=lo)
    assert ref.npivot == 4
    I, J = ref.rowindices - 1, ref.colindices - 1
    np.testing.assert_allclose(ref.left @ ref.right, A, rtol=1e-10, atol=1e-12)
    if lo:
        np.testing.assert_allclose(ref.right[:, J], A[I][:, J], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(ref.left[I, :], np.eye(4), atol=1e-12)
    else:
        np.testing.assert_allclose(ref.left[I, :], A[I][:, J], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(ref.right[:, J], np.eye(4), atol=1e-12)
    B = crand(rng, 9, 9)
    full = oracle.zluci(B, leftorthogonal=lo)
    assert full.npivot == 9 and full.error == 0.0
    np.testing.assert_allclose(full.left @ full.right, B, rtol=1e-10, atol=1e-11)


def test_complex_global_pivot_finder_host_logic():
    """DefaultGlobalPivotFinder for a ComplexF64 target (complexf64.zfind_global_pivots) against a direct restatement of
    globalpivotfinder.jl:143-195: star probes from every start, error = abs(f - tt), strict `>` keeps the first maximum,
    threshold abstol * tolmarginglobalsearch, truncation to the first maxnglobalpivot in start order.  The evaluators are
    numpy stand-ins (no GPU): only the host logic and the library's host-only selection are exercised."""
    import tci_b200 as T
    from tci_b200.complexf64 import zfind_global_pivots
    ld = [3, 4, 2, 5]
    rng = np.random.default_rng(77)
    F = crand(rng, *ld)
    G = F.copy()
    bumps = [(0, 1, 1, 2), (2, 3, 0, 4), (1, 0, 1, 1)]
    for k, b in enumerate(bumps):
        G[b] += (0.5 + 0.25 * k) * (1 + 1j)

    class Ev:
        is_complex = True
        ctx = None

        def __init__(self, arr):
            self.arr = arr

        def evaluate_points(self, pts):
            pts = np.asarray(pts)
            return self.arr[tuple((pts - 1).T)]

    class TT:
        device_handle = None

    tt = TT()
    tt.device_handle = Ev(F)
    finder = T.DefaultGlobalPivotFinder(nsearch=12, maxnglobalpivot=3, tolmarginglobalsearch=10.0)
    inp = T.GlobalPivotSearchInput(ld, tt, 1.0, None, None)
    abstol = 0.01
    got = zfind_global_pivots(finder, inp, Ev(G), abstol, rng=T.CounterRNG(5))
    starts = T.CounterRNG(5).start_points(12, ld)
    ref = []
    for s in starts:  # :156-188
        best, bestp = 0.0, list(s)
        for p, d in enumerate(ld):
            for v in range(1, d + 1):
                x = list(s)
                x[p] = v
                e = abs(G[tuple(np.array(x) - 1)] - F[tuple(np.array(x) - 1)])
                if e > best:
                    best, bestp = e, x
        if best > abstol * 10.0:
            ref.append([int(v) for v in bestp])
    assert got.tolist() == ref[:3] and len(ref) >= 1
    np.testing.assert_allclose(finder.last_errors, [abs(G[tuple(np.array(x) - 1)] - F[tuple(np.array(x) - 1)]) for x in ref[:3]],
                               rtol=1e-15)


# ---- GPU ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def T():
    import tci_b200
    return tci_b200


def assert_zlu_equal(lu, ref):
    assert lu.npivot == ref.npivot
    assert np.array_equal(lu.rowpermutation, ref.rowpermutation)
    assert np.array_equal(lu.colpermutation, ref.colpermutation)
    assert lu.error == ref.error or (np.isnan(lu.error) and np.isnan(ref.error))
    assert np.array_equal(lu._pivoterrors, ref.pivoterrors)
    assert np.array_equal(lu.L, ref.L) and np.array_equal(lu.U, ref.U)  # bit-identical complex factors


ZSHAPES = [(1, 1, 1), (3, 6, 3), (9, 7, 7), (40, 70, 12), (257, 300, 33), (513, 129, 40), (64, 1500, 21), (1200, 90, 17)]


@pytest.mark.gpu
@pytest.mark.parametrize("lo", [True, False])
@pytest.mark.parametrize("m,n,r", ZSHAPES)
def test_zrrlu_bit_exact(T, oracle, m, n, r, lo):
    rng = np.random.default_rng(1000 * m + n)
    A = lowrank(rng, m, n, r)
    for kw in (dict(maxrank=r, reltol=1e-12), dict(reltol=1e-9), dict(abstol=1e-6 * np.max(np.abs(A)))):
        assert_zlu_equal(T.rrlu(A, leftorthogonal=lo, **kw), oracle.zrrlu(A, leftorthogonal=lo, **kw))


@pytest.mark.gpu
def test_zrrlu_reference_literal_full_rank_and_ties(T, oracle):
    for lo in (True, False):
        assert_zlu_equal(T.rrlu(ARGMAX_LITERAL, leftorthogonal=lo), oracle.zrrlu(ARGMAX_LITERAL, leftorthogonal=lo))
        rng = np.random.default_rng(4)
        for (m, n) in ((50, 50), (120, 80), (33, 200)):
            A = crand(rng, m, n)
            assert_zlu_equal(T.rrlu(A, leftorthogonal=lo), oracle.zrrlu(A, leftorthogonal=lo))
        # exact ties: every |a|^2 equal (unit-modulus entries built from exactly representable parts), and a matrix of
        # equal entries -- the scan order (columns outer, rows inner, strict >) decides
        ph = np.array([1, 1j, -1, -1j])[rng.integers(0, 4, (12, 9))]
        assert_zlu_equal(T.rrlu(ph, leftorthogonal=lo, maxrank=5), oracle.zrrlu(ph, leftorthogonal=lo, maxrank=5))
        ones = np.full((7, 5), 1 + 1j)
        assert_zlu_equal(T.rrlu(ones, leftorthogonal=lo), oracle.zrrlu(ones, leftorthogonal=lo))
        # real data through the complex kernel: the Float64 path's pivots
        R = rng.standard_normal((40, 30))
        zr, rr = T.rrlu(R.astype(np.complex128), leftorthogonal=lo), T.rrlu(R, leftorthogonal=lo)
        assert np.array_equal(zr.rowpermutation, rr.rowpermutation) and np.array_equal(zr.colpermutation, rr.colpermutation)
        np.testing.assert_allclose(zr.L.real, rr.L, rtol=1e-12, atol=1e-14)  # complex division rounds once more
        assert np.max(np.abs(zr.L.imag)) == 0.0


@pytest.mark.gpu
def test_zrrlu_nan_errors(T, oracle):  # matrixlu.jl:164-169
    for (i, j), msg in (((2, 0), "lu.L contains NaNs"), ((0, 2), "lu.U contains NaNs")):
        A = np.ones((4, 4), dtype=np.complex128)
        A[0, 0] = 5.0
        A[i, j] = complex(np.nan, 0.0)
        with pytest.raises(oracle.OracleError, match=msg):
            oracle.zrrlu(A)
        with pytest.raises(RuntimeError, match=msg):
            T.rrlu(A)
    A = np.ones((4, 4), dtype=np.complex128)  # a NaN that never reaches L or U: no error, same result
    A[2, 1] = complex(np.nan, 0.0)
    assert_zlu_equal(T.rrlu(A), oracle.zrrlu(A))


@pytest.mark.gpu
@pytest.mark.parametrize("lo", [True, False])
@pytest.mark.parametrize("m,n,r", [(8, 6, 4), (40, 70, 33), (300, 200, 64), (129, 515, 100)])
def test_zluci_left_right(T, oracle, m, n, r, lo):  # test_matrixluci.jl:6-74 on complex data
    rng = np.random.default_rng(m + n)
    A = lowrank(rng, m, n, r, decay=8.0)
    luci = T.MatrixLUCI(A, maxrank=r, leftorthogonal=lo)
    ref = oracle.zluci(A, maxrank=r, leftorthogonal=lo)
    assert np.array_equal(T.rowindices(luci), ref.rowindices) and np.array_equal(T.colindices(luci), ref.colindices)
    L, Rt = luci.left(), luci.right()
    scale = max(1.0, np.max(np.abs(ref.left)))
    assert np.max(np.abs(L - ref.left)) <= RTOL * scale
    assert np.max(np.abs(Rt - ref.right)) <= RTOL * max(1.0, np.max(np.abs(ref.right)))
    I, J = T.rowindices(luci) - 1, T.colindices(luci) - 1  # A ~ A[:, J] P^-1 A[I, :]  (test_matrixluci.jl:29-37)
    np.testing.assert_allclose(L @ Rt, A, atol=1e-6 * np.max(np.abs(A)))
    if lo:
        np.testing.assert_allclose(Rt[:, J], A[I][:, J], rtol=1e-9, atol=1e-12)
    else:
        np.testing.assert_allclose(L[I, :], A[I][:, J], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_zlu_rdiv(T):
    rng = np.random.default_rng(8)
    for lo in (True, False):
        P, B = crand(rng, 37, 37), crand(rng, 90, 37)
        lu = T.rrlu(P, reltol=0.0, abstol=0.0, leftorthogonal=lo)
        assert lu.npivot == 37
        np.testing.assert_allclose(lu.rdiv(B), B @ np.linalg.inv(P), rtol=1e-9, atol=1e-9)


@pytest.mark.gpu
def test_zgemm_kernel(T):
    rng = np.random.default_rng(3)
    for (M, N, K) in ((1, 1, 1), (5, 7, 3), (64, 64, 8), (65, 130, 37), (200, 96, 256), (33, 1, 500)):
        A, B = crand(rng, M, K), crand(rng, K, N)
        ref = A @ B
        np.testing.assert_allclose(T.zgemm(A, B), ref, rtol=1e-12, atol=1e-12 * np.max(np.abs(ref)))
        np.testing.assert_allclose(T.zgemm(np.asfortranarray(A.T), B, transA=True), ref, rtol=1e-12, atol=1e-12 * np.max(np.abs(ref)))
        np.testing.assert_allclose(T.zgemm(A, np.asfortranarray(B.T), transB=True), ref, rtol=1e-12, atol=1e-12 * np.max(np.abs(ref)))


def _tto_tto(rng):  # _gen_testdata_TTO_TTO, test_contraction.jl:31-48
    ba = bb = [1, 2, 3, 2, 1]
    d1, d2, d3 = [2] * 4, [3] * 4, [2] * 4
    a = [np.asfortranarray(rng.random((ba[n], d1[n], d2[n], ba[n + 1])) + 1j * rng.random((ba[n], d1[n], d2[n], ba[n + 1])))
         for n in range(4)]
    b = [np.asfortranarray(rng.random((bb[n], d2[n], d3[n], bb[n + 1])) + 1j * rng.random((bb[n], d2[n], d3[n], bb[n + 1])))
         for n in range(4)]
    return a, b


@pytest.mark.gpu
def test_complex_contraction_target_all_splits(T, oracle):
    """Contraction(a, b) on TensorTrain{ComplexF64,4}: evaluate and batchevaluate for every split against the oracle's
    restatement of contraction.jl:236-335 (numpy complex128), incl. the layout of the reference's batchevaluate testset
    (test_contraction.jl:101-139: M = 2 with one-point index sets)."""
    rng = np.random.default_rng(17)
    a, b = _tto_tto(rng)
    f = T.ZContraction(a, b)
    N, ld = 4, f.localdims
    assert ld == [4, 4, 4, 4]
    full = oracle.mpo_batchevaluate_projected(a, b, [[]], [[]], 4).reshape(ld, order="F")
    pts = np.stack([rng.integers(1, d + 1, 50) for d in ld], axis=1)
    vals = f.evaluate_points(pts)
    np.testing.assert_allclose(vals, full[tuple((pts - 1).T)], rtol=RTOL, atol=1e-13)
    for nl in range(N + 1):
        for nr in range(N + 1 - nl):
            M = N - nl - nr
            I = np.stack([rng.integers(1, d + 1, 5 if nl else 1) for d in ld[:nl]], axis=1) if nl else np.zeros((1, 0), dtype=np.int64)
            J = np.stack([rng.integers(1, d + 1, 4 if nr else 1) for d in ld[N - nr:]], axis=1) if nr else np.zeros((1, 0), dtype=np.int64)
            res = f(I, J, M)
            ref = oracle.mpo_batchevaluate_projected(a, b, I.tolist(), J.tolist(), M)
            assert res.shape == ref.shape and res.dtype == np.complex128
            np.testing.assert_allclose(res, ref, rtol=RTOL, atol=1e-13)
            dev, mx = f.batchevaluate_device(I, J, M)
            assert np.array_equal(dev.to_host().reshape(res.shape, order="F"), res)
            assert mx == pytest.approx(np.max(np.abs(ref)), rel=1e-12)
    ref2 = f([[1]], [[1]], 2)  # test_contraction.jl:118-121
    assert ref2.shape == (1, 4, 4, 1)
    g = T.ZContraction(a, b, f=("affine", 2.0, 0.0))  # f = x -> 2x (test_contraction.jl:68)
    np.testing.assert_allclose(g([[1]], [[1]], 2), 2 * ref2, rtol=1e-15)
    np.testing.assert_allclose(g.evaluate_points(pts), 2 * vals, rtol=1e-15)


def _dense_product(a, b):
    va = np.ones((1, 1, 1), dtype=np.complex128)
    for c in a:
        va = np.einsum("rcb,bxyd->rxcyd", va, c).reshape((va.shape[0] * c.shape[1], va.shape[1] * c.shape[2], c.shape[3]))
    vb = np.ones((1, 1, 1), dtype=np.complex128)
    for c in b:
        vb = np.einsum("rcb,bxyd->rxcyd", vb, c).reshape((vb.shape[0] * c.shape[1], vb.shape[1] * c.shape[2], c.shape[3]))
    return va[:, :, 0], vb[:, :, 0]  # site-major (first site slowest) enumeration on both legs, consistently


@pytest.mark.gpu
@pytest.mark.parametrize("fspec", [None, ("affine", 2.0, 0.0)])
def test_complex_mpo_mpo_contraction_tci(T, fspec):
    """"MPO-MPO contraction" for algorithm = :TCI (test_contraction.jl:68-99) on ComplexF64 data: _tomat(ab) ~
    f.(_tomat(a) * _tomat(b)); every stage -- Pi evaluation, rrLU, site-tensor solves, global search -- runs the
    complex kernels."""
    rng = np.random.default_rng(23)
    a, b = _tto_tto(rng)
    ab = T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="TCI", tolerance=1e-12, f=fspec, rng=T.CounterRNG(3))
    assert T.sitedims(ab) == [[2, 2]] * 4 and ab.sitetensors[0].dtype == np.complex128
    ma, mb = _dense_product(a, b)
    mab, _ = _dense_product(ab.sitetensors, ab.sitetensors)
    ref = ma @ mb
    if fspec is not None:
        ref = 2 * ref
    np.testing.assert_allclose(mab, ref, rtol=1e-8, atol=1e-9 * np.max(np.abs(ref)))


@pytest.mark.gpu
def test_complex_mpo_mps_contraction_tci(T):
    """"MPO-MPS contraction" (test_contraction.jl:148-183), algorithm = :TCI, ComplexF64."""
    rng = np.random.default_rng(29)
    bonds = [1, 2, 3, 2, 1]
    a = [np.asfortranarray(crand(rng, bonds[n], 3, 3, bonds[n + 1])) for n in range(4)]
    b = [np.asfortranarray(crand(rng, bonds[n], 3, bonds[n + 1])) for n in range(4)]
    ab = T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="TCI", tolerance=1e-12, rng=T.CounterRNG(5))
    assert T.sitedims(ab) == [[3]] * 4
    ma, _ = _dense_product(a, a)
    vb = np.ones((1, 1), dtype=np.complex128)
    for c in b:
        vb = np.einsum("rb,bxd->rxd", vb, c).reshape((-1, c.shape[2]))
    vab = np.ones((1, 1), dtype=np.complex128)
    for c in ab.sitetensors:
        vab = np.einsum("rb,bxd->rxd", vab, c).reshape((-1, c.shape[2]))
    ref = ma @ vb[:, 0]
    np.testing.assert_allclose(vab[:, 0], ref, rtol=1e-8, atol=1e-9 * np.max(np.abs(ref)))
    with pytest.raises(RuntimeError, match="Naive contraction implementation cannot contract"):
        T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="naive", f=("affine", 2.0, 0.0))


@pytest.mark.gpu
def test_complex_crossinterpolate2_of_tt_target(T):
    """crossinterpolate2(ComplexF64, f, ...) with f a TTCache{ComplexF64} (test_conversion.jl:76-79 structure: a complex
    tensor train re-interpolated to the same tensor): rank and values recovered."""
    rng = np.random.default_rng(31)
    bonds = [1, 3, 4, 3, 1]
    cores = [np.asfortranarray(crand(rng, bonds[n], 4, bonds[n + 1])) for n in range(4)]
    f = T.ZTTCache(cores)
    tci, ranks, errors = T.crossinterpolate2(f, [4] * 4, tolerance=1e-12, rng=T.CounterRNG(2))
    assert T.linkdims(tci) == [3, 4, 3]
    pts = np.stack([rng.integers(1, 5, 100) for _ in range(4)], axis=1)
    exact = f.evaluate_points(pts)
    got = T.evaluate_points(T.TensorTrain(tci.sitetensors), pts)
    np.testing.assert_allclose(got, exact, rtol=1e-9, atol=1e-10 * np.max(np.abs(exact)))
    dense = np.ones((1, 1), dtype=np.complex128)
    for c in cores:
        dense = np.einsum("rb,bxd->rxd", dense, c).reshape((-1, c.shape[2]))
    assert T.tci_sum(tci) == pytest.approx(np.sum(dense), rel=1e-9)


@pytest.mark.gpu
def test_complex_mpo_mpo_contraction_naive_and_zipup(T):
    """"MPO-MPO contraction" for algorithm = :naive (test_contraction.jl:68-99) and "MPO-MPO contraction (zipup)" for
    method in [:SVD, :LU] (:185-189) on ComplexF64 data: site contractions, factorisations and recompression products all
    on the complex kernels; `f` with :naive raises as in the reference."""
    rng = np.random.default_rng(41)
    a, b = _tto_tto(rng)
    ma, mb = _dense_product(a, b)
    ref = ma @ mb
    ab = T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="naive")
    assert T.sitedims(ab) == [[2, 2]] * 4 and ab.sitetensors[0].dtype == np.complex128
    np.testing.assert_allclose(_dense_product(ab.sitetensors, ab.sitetensors)[0], ref, rtol=1e-10, atol=1e-12 * np.max(np.abs(ref)))
    ab = T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="naive", tolerance=1e-12)  # + compress(:SVD)
    np.testing.assert_allclose(_dense_product(ab.sitetensors, ab.sitetensors)[0], ref, rtol=1e-8, atol=1e-9 * np.max(np.abs(ref)))
    with pytest.raises(RuntimeError, match="Naive contraction implementation cannot contract"):
        T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="naive", f=("affine", 2.0, 0.0))
    for method in ("SVD", "LU"):
        ab = T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="zipup", method=method)
        np.testing.assert_allclose(_dense_product(ab.sitetensors, ab.sitetensors)[0], ref, rtol=1e-8,
                                   atol=1e-9 * np.max(np.abs(ref)))
    tt = T.TensorTrain([c.copy() for c in T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="naive").sitetensors])
    T.compress(tt, "LU", tolerance=1e-12)  # compress! (:LU) on complex cores: complex rrLU + complex GEMM
    np.testing.assert_allclose(_dense_product(tt.sitetensors, tt.sitetensors)[0], ref, rtol=1e-8, atol=1e-9 * np.max(np.abs(ref)))


@pytest.mark.gpu
def test_complex_mpo_mps_contraction_zipup(T):
    """"MPO-MPS contraction (zipup)" (test_contraction.jl:191-195), ComplexF64."""
    rng = np.random.default_rng(43)
    bonds = [1, 2, 3, 2, 1]
    a = [np.asfortranarray(crand(rng, bonds[n], 3, 3, bonds[n + 1])) for n in range(4)]
    b = [np.asfortranarray(crand(rng, bonds[n], 3, bonds[n + 1])) for n in range(4)]
    ma, _ = _dense_product(a, a)
    vb = np.ones((1, 1), dtype=np.complex128)
    for c in b:
        vb = np.einsum("rb,bxd->rxd", vb, c).reshape((-1, c.shape[2]))
    ref = ma @ vb[:, 0]
    for method in ("SVD", "LU"):
        ab = T.contract(T.TensorTrain(a), T.TensorTrain(b), algorithm="zipup", method=method)
        v = np.ones((1, 1), dtype=np.complex128)
        for c in ab.sitetensors:
            v = np.einsum("rb,bxd->rxd", v, c).reshape((-1, c.shape[2]))
        np.testing.assert_allclose(v[:, 0], ref, rtol=1e-8, atol=1e-9 * np.max(np.abs(ref)))
