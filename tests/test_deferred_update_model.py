"""CPU model of the idea behind csrc/rrlu_lazy.cu: the Schur updates of the full-pivot rrLU (matrixlu.jl:114-136) can
be DEFERRED -- the stored matrix is rewritten only every NB pivots and in between every trailing value is re-derived
as  v = a;  v = v - x_1*y_1;  v = v - x_2*y_2; ...  (each product and difference rounded on its own) -- without
changing a single bit of the factorisation.  Row swaps are kept virtual inside a block exactly as in the kernel:
a list of at most 2*NB "special" base rows with their current positions.  Checked against the oracle's in-place
restatement of the reference."""
import numpy as np
import pytest


def deferred_rrlu(A, maxrank, NB=4, leftorthogonal=True):
    C = np.array(A, dtype=np.float64, order="F")  # committed matrix, rows physically permuted at commits only
    m, n = C.shape
    rowperm, colperm = np.arange(m), np.arange(n)  # position -> original index
    colpos = np.arange(n)  # physical column -> position (columns are only ever permuted virtually)
    pivvals = []
    k0 = 0
    pend = []  # (base row of the pivot, physical column, x indexed by base row, y indexed by physical column, value)
    pos = np.arange(m)  # base row -> current position inside the block

    def current(rows, cols):  # values of the given base rows / physical columns with all pending updates applied
        V = C[np.ix_(rows, cols)].copy()
        for (_, _, x, y, _) in pend:
            V = V - np.multiply.outer(x[rows], y[cols])  # rounded product, then rounded difference
        return V

    def commit():
        nonlocal C, pos, k0, pend
        nd = len(pend)
        lo = k0 + nd
        act_rows = np.where(pos >= lo)[0]
        act_cols = np.where(colpos >= lo)[0]
        new = C.copy()
        new[np.ix_(act_rows, act_cols)] = current(act_rows, act_cols)
        for j, (p, c, x, y, val) in enumerate(pend):
            new[p, act_cols] = y[act_cols]  # U row
            later = [q for q in range(m) if pos[q] > k0 + j]  # rows that were active when pivot j was taken
            new[later, c] = x[later]  # L column (or the unscaled column when not left-orthogonal)
            new[p, c] = val
            for jj in range(j):  # U entries above the diagonal inside the block
                new[pend[jj][0], c] = pend[jj][3][c]
        C = new[np.argsort(pos)]  # every row to the position the reference's physical swaps give it
        pos = np.arange(m)
        k0 = lo
        pend = []

    maxerror, lasterr = 0.0, np.nan
    for s in range(maxrank):
        lo = k0 + len(pend)
        rows = np.where(pos >= lo)[0]
        rows = rows[np.argsort(pos[rows])]  # scan order = position order
        cols = np.where(colpos >= lo)[0]
        cols = cols[np.argsort(colpos[cols])]
        V = current(rows, cols)
        q = V * V
        q[np.isnan(q)] = -np.inf
        flat = np.argmax(q.T)  # columns outer, rows inner, first maximum (matrixlu.jl:16-29)
        ci, ri = divmod(int(flat), len(rows))
        p, c, val = rows[ri], cols[ci], V[ri, ci]
        lasterr = abs(val)
        if s > 0 and (lasterr < 1e-14 * maxerror):
            break
        maxerror = max(maxerror, lasterr)
        # virtual swaps: position s <-> position of the pivot
        qrow = np.where(pos == s)[0][0]
        pos[qrow], pos[p] = pos[p], s
        rowperm[s], rowperm[pos[qrow]] = rowperm[pos[qrow]], rowperm[s]
        ccol = np.where(colpos == s)[0][0]
        colpos[ccol], colpos[c] = colpos[c], s
        colperm[s], colperm[colpos[ccol]] = colperm[colpos[ccol]], colperm[s]
        full_rows, full_cols = np.arange(m), np.arange(n)
        x = current(full_rows, [c])[:, 0]
        y = current([p], full_cols)[0]
        if leftorthogonal:
            x = x / val
        else:
            y = y / val
        pend.append((p, c, x, y, val))
        pivvals.append(val)
        if len(pend) == NB:
            commit()
    if pend:
        commit()
    r = len(pivvals)
    F = C[:, np.argsort(colpos)]  # columns to their positions
    L = np.tril(F[:, :r], -1)
    U = np.triu(F[:r, :], 1)
    d = np.diag(F)[:r]
    if leftorthogonal:
        L[np.arange(r), np.arange(r)] = 1.0
        U[np.arange(r), np.arange(r)] = d
    else:
        L[np.arange(r), np.arange(r)] = d
        U[np.arange(r), np.arange(r)] = 1.0
    return rowperm + 1, colperm + 1, L, U


@pytest.mark.parametrize("leftorthogonal", [True, False])
@pytest.mark.parametrize("m,n,r,nb", [(9, 7, 5, 4), (20, 31, 13, 4), (40, 40, 40, 4), (33, 18, 18, 3), (25, 25, 10, 8)])
def test_deferred_updates_are_bit_identical(oracle, m, n, r, nb, leftorthogonal):
    rng = np.random.default_rng(m * 100 + n)
    A = rng.standard_normal((m, n))
    ref = oracle.rrlu(A, maxrank=r, reltol=1e-14, leftorthogonal=leftorthogonal)
    rp, cp, L, U = deferred_rrlu(A, r, NB=nb, leftorthogonal=leftorthogonal)
    k = ref.npivot
    assert np.array_equal(rp[:k], ref.rowpermutation[:k]) and np.array_equal(cp[:k], ref.colpermutation[:k])
    assert np.array_equal(L[:, :k], ref.L) and np.array_equal(U[:k], ref.U)


def test_deferred_updates_with_ties_and_exact_zeros(oracle):
    rng = np.random.default_rng(3)
    A = rng.integers(-2, 3, (24, 30)).astype(np.float64)  # many equal maxima: the first in column-major order wins
    ref = oracle.rrlu(A, maxrank=12)
    rp, cp, L, U = deferred_rrlu(A, 12)
    k = ref.npivot
    assert np.array_equal(rp[:k], ref.rowpermutation[:k]) and np.array_equal(cp[:k], ref.colpermutation[:k])
    assert np.array_equal(L[:, :k], ref.L) and np.array_equal(U[:k], ref.U)
