// tci_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A plain, single-threaded restatement of the reference algorithm for the
// TensorCI2 two-site hot path of tensor4all/TensorCrossInterpolation.jl v0.9.19.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may load this library; the product (libtci_b200.so and the Python host
// mirror) never does.
//
// PARITY STATUS: the reference is Julia and no Julia runtime exists in this
// image or on the GPU boxes, so this restatement cannot be run against the
// reference itself.  It is pinned against every literal fixture the reference's
// own tests hold for this path (tests/test_oracle_golden.py lists them with
// file:line).  Parts whose bit-level behaviour lives in OpenBLAS/LAPACK
// (TRSM/GEMM/getrf inside matrixluci.jl, cachedtensortrain.jl, contraction.jl,
// tensorci2.jl:391) or in Julia's RNG are "parity unpinned" below the
// reference's own test tolerance (rtol sqrt(eps)); random choices are injected.
//
// Build: g++ -O2 -ffp-contract=off (no FMA contraction: Julia does not contract
// `a - x*y` either), see oracle/Makefile.
//
// All file:line citations are relative to /root/reference/src/.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../include/tci_targets.h"
#include "../include/tci_zarith.h"

typedef int64_t i64;
typedef std::vector<i64> MultiIndex; // abstracttensortrain.jl:6-7 (1-based values)
typedef std::vector<MultiIndex> IndexList;

static thread_local std::string g_err;

// Julia's max(x, y) for Float64 propagates NaN (unlike fmax).
static inline double jl_max(double x, double y)
{
    if (std::isnan(x) || std::isnan(y)) return NAN;
    return x > y ? x : y;
}

// ============================================================== rrLU =======
// matrixlu.jl:1-32 -- first maximum of f(A[r,c]) over r,c >= k, columns outer,
// rows inner, strict '>' starting from -Inf (NaN never selected).
static void submatrixargmax_abs2(const double *A, i64 m, i64 n, i64 k0, i64 &mr, i64 &mc)
{
    double best = -INFINITY;
    mr = k0;
    mc = k0;
    for (i64 c = k0; c < n; ++c) {
        const double *col = A + c * m;
        for (i64 r = k0; r < m; ++r) {
            double v = col[r] * col[r]; // abs2, matrixlu.jl:152
            if (v > best) {
                best = v;
                mr = r;
                mc = c;
            }
        }
    }
}

struct RRLU { // matrixlu.jl:71-96
    i64 m = 0, n = 0;
    std::vector<i64> rowperm, colperm; // 1-based
    std::vector<double> L, U;          // m x r, r x n (column-major)
    bool leftorthogonal = true;
    i64 npivot = 0;
    double error = NAN;
};

// matrixlu.jl:98-136 (swaprow!/swapcol!/addpivot!), in place on A (m x n).
static void addpivot(RRLU &lu, double *A, i64 m, i64 n, i64 pr, i64 pc)
{
    i64 k = lu.npivot++; // 0-based position of the new pivot
    std::swap(lu.rowperm[k], lu.rowperm[pr]);
    for (i64 j = 0; j < n; ++j) std::swap(A[k + j * m], A[pr + j * m]);
    std::swap(lu.colperm[k], lu.colperm[pc]);
    for (i64 i = 0; i < m; ++i) std::swap(A[i + k * m], A[i + pc * m]);

    double piv = A[k + k * m];
    if (lu.leftorthogonal) {
        for (i64 i = k + 1; i < m; ++i) A[i + k * m] /= piv; // :120
    } else {
        for (i64 j = k + 1; j < n; ++j) A[k + j * m] /= piv; // :122
    }
    const double *x = A + k * m; // column k
    for (i64 j = k + 1; j < n; ++j) {
        double yj = A[k + j * m];
        double *col = A + j * m;
        for (i64 i = k + 1; i < m; ++i) col[i] -= x[i] * yj; // :132, mul then sub
    }
}

// matrixlu.jl:141-181.  Returns 0, or 1/2 for "lu.L/lu.U contains NaNs".
static int optimizerrlu(RRLU &lu, double *A, i64 m, i64 n, i64 maxrank, double reltol, double abstol)
{
    maxrank = std::min(maxrank, std::min(m, n));
    double maxerror = 0.0;
    while (lu.npivot < maxrank) {
        i64 k = lu.npivot;
        i64 pr, pc;
        submatrixargmax_abs2(A, m, n, k, pr, pc);
        lu.error = std::fabs(A[pr + pc * m]);
        if ((std::fabs(lu.error) < reltol * maxerror || std::fabs(lu.error) < abstol) && lu.npivot > 0) break;
        maxerror = jl_max(maxerror, lu.error);
        addpivot(lu, A, m, n, pr, pc);
    }
    i64 r = lu.npivot;
    lu.L.assign((size_t)(m * r), 0.0);
    lu.U.assign((size_t)(r * n), 0.0);
    for (i64 c = 0; c < r; ++c)
        for (i64 i = c; i < m; ++i) lu.L[i + c * m] = A[i + c * m]; // tril(A[:,1:r])
    for (i64 c = 0; c < n; ++c)
        for (i64 i = 0; i <= std::min(c, r - 1); ++i) lu.U[i + c * r] = A[i + c * m]; // triu(A[1:r,:])
    for (double v : lu.L)
        if (std::isnan(v)) {
            g_err = "lu.L contains NaNs";
            return 1;
        }
    for (double v : lu.U)
        if (std::isnan(v)) {
            g_err = "lu.U contains NaNs";
            return 2;
        }
    for (i64 c = 0; c < r; ++c) {
        if (lu.leftorthogonal)
            lu.L[c + c * m] = 1.0;
        else
            lu.U[c + c * r] = 1.0;
    }
    if (lu.npivot >= std::min(m, n)) lu.error = 0.0;
    return 0;
}

// rrlu(A; ...) = rrlu!(copy(A)) matrixlu.jl:194-225
static int rrlu(RRLU &lu, const double *Ain, i64 m, i64 n, i64 maxrank, double reltol, double abstol,
                bool leftorthogonal)
{
    lu = RRLU();
    lu.m = m;
    lu.n = n;
    lu.leftorthogonal = leftorthogonal;
    lu.rowperm.resize(m);
    lu.colperm.resize(n);
    for (i64 i = 0; i < m; ++i) lu.rowperm[i] = i + 1;
    for (i64 j = 0; j < n; ++j) lu.colperm[j] = j + 1;
    std::vector<double> A(Ain, Ain + m * n);
    return optimizerrlu(lu, A.data(), m, n, maxrank, reltol, abstol);
}

// pivoterrors = [abs.(diag(lu)); lu.error]  matrixlu.jl:394-416
static std::vector<double> pivoterrors(const RRLU &lu)
{
    std::vector<double> e;
    i64 r = lu.npivot;
    for (i64 i = 0; i < r; ++i)
        e.push_back(std::fabs(lu.leftorthogonal ? lu.U[i + i * r] : lu.L[i + i * lu.m]));
    e.push_back(lu.error);
    return e;
}

// C(MxN) = A(MxK) * B(KxN), column-major, k accumulated in order (mul, add).
static void gemm_seq(const double *A, i64 lda, const double *B, i64 ldb, double *C, i64 ldc, i64 M, i64 N, i64 K)
{
    for (i64 j = 0; j < N; ++j) {
        double *c = C + j * ldc;
        for (i64 i = 0; i < M; ++i) c[i] = 0.0;
        for (i64 k = 0; k < K; ++k) {
            double b = B[k + j * ldb];
            const double *a = A + k * lda;
            for (i64 i = 0; i < M; ++i) c[i] += a[i] * b;
        }
    }
}

// matrixluci.jl:70-84 -- left(luci): m x r, right(luci): r x n.
static void luci_left(const RRLU &lu, std::vector<double> &out)
{
    i64 m = lu.m, r = lu.npivot;
    out.assign((size_t)(m * r), 0.0);
    std::vector<double> res((size_t)(m * r), 0.0);
    if (lu.leftorthogonal) { // colstimespivotinv :48-57
        for (i64 c = 0; c < r; ++c) res[c + c * m] = 1.0;
        // X * L11 = L21, L11 unit lower: columns from last to first
        for (i64 j = r - 1; j >= 0; --j) {
            for (i64 i = r; i < m; ++i) {
                double acc = lu.L[i + j * m];
                for (i64 k = j + 1; k < r; ++k) acc -= res[i + k * m] * lu.L[k + j * m];
                res[i + j * m] = acc / lu.L[j + j * m];
            }
        }
    } else { // colmatrix :40-42 : L (m x r) * U[:, 1:r]
        gemm_seq(lu.L.data(), m, lu.U.data(), r, res.data(), m, m, r, r);
    }
    for (i64 c = 0; c < r; ++c)
        for (i64 i = 0; i < m; ++i) out[(lu.rowperm[i] - 1) + c * m] = res[i + c * m];
}

static void luci_right(const RRLU &lu, std::vector<double> &out)
{
    i64 m = lu.m, n = lu.n, r = lu.npivot;
    out.assign((size_t)(r * n), 0.0);
    std::vector<double> res((size_t)(r * n), 0.0);
    if (lu.leftorthogonal) { // rowmatrix :44-46 : L[1:r,:] * U
        gemm_seq(lu.L.data(), m, lu.U.data(), r, res.data(), r, r, n, r);
    } else { // pivotinvtimesrows :59-68 : U11 \ U12, U11 unit upper
        for (i64 c = 0; c < r; ++c) res[c + c * r] = 1.0;
        for (i64 j = r; j < n; ++j) {
            for (i64 i = r - 1; i >= 0; --i) {
                double acc = lu.U[i + j * r];
                for (i64 k = i + 1; k < r; ++k) acc -= lu.U[i + k * r] * res[k + j * r];
                res[i + j * r] = acc / lu.U[i + i * r];
            }
        }
    }
    for (i64 c = 0; c < n; ++c)
        for (i64 i = 0; i < r; ++i) out[i + (lu.colperm[c] - 1) * r] = res[i + c * r];
}

// ------------------------------------------------------------ rook search ---
// Injected replacement of Julia's rand(1:len) in randomsubset (util.jl:36-52): counter based.
struct RookRng {
    uint64_t seed = 0, counter = 0;
    i64 draw(i64 len)
    {
        i64 v = 1 + (i64)(tci_uniform01(seed ^ 0x726f6f6bull, counter++) * (double)len);
        return v > len ? len : v;
    }
};

// pushrandomsubset!(subset, 1:N, cnt)  util.jl:54-58
static void pushrandomsubset(std::vector<i64> &subset, i64 N, i64 cnt, RookRng &rng)
{
    std::vector<i64> c;
    for (i64 v = 1; v <= N; ++v)
        if (std::find(subset.begin(), subset.end(), v) == subset.end()) c.push_back(v);
    i64 nn = std::min<i64>(cnt, (i64)c.size());
    for (i64 q = 0; q < nn; ++q) {
        i64 index = rng.draw((i64)c.size());
        subset.push_back(c[index - 1]);
        c.erase(c.begin() + (index - 1));
    }
}

// cols2Lmatrix! / rows2Umatrix!  matrixlu.jl:314-358 (C: rows x r, R: r x cols, P: r x r, column-major)
static void cols2Lmatrix(std::vector<double> &C, i64 rows, const std::vector<double> &P, i64 r)
{
    for (i64 k = 0; k < r; ++k) {
        for (i64 i = 0; i < rows; ++i) C[i + k * rows] /= P[k + k * r];
        for (i64 j = k + 1; j < r; ++j) {
            double y = P[k + j * r];
            for (i64 i = 0; i < rows; ++i) C[i + j * rows] -= C[i + k * rows] * y;
        }
    }
}
static void rows2Umatrix(std::vector<double> &R, i64 cols, const std::vector<double> &P, i64 r)
{
    for (i64 k = 0; k < r; ++k) {
        for (i64 j = 0; j < cols; ++j) R[k + j * r] /= P[k + k * r];
        for (i64 j = 0; j < cols; ++j) {
            double y = R[k + j * r];
            for (i64 i = k + 1; i < r; ++i) R[i + j * r] -= P[i + k * r] * y;
        }
    }
}

// f(rows, cols) -> |rows| x |cols| matrix (1-based indices), the `_batchf` of matrixlu.jl:243
typedef void (*MatEval)(void *user, const std::vector<i64> &rows, const std::vector<i64> &cols,
                        std::vector<double> &out);

// arrlu  matrixlu.jl:227-293
static int arrlu(RRLU &lu, MatEval f, void *user, i64 m, i64 n, std::vector<i64> I0, std::vector<i64> J0, i64 maxrank,
                 double reltol, double abstol, bool leftorth, int numrookiter, RookRng &rng)
{
    lu = RRLU();
    lu.leftorthogonal = leftorth;
    lu.rowperm.resize(m);
    lu.colperm.resize(n);
    for (i64 i = 0; i < m; ++i) lu.rowperm[i] = i + 1;
    for (i64 j = 0; j < n; ++j) lu.colperm[j] = j + 1;
    bool islowrank = false;
    maxrank = std::min(maxrank, std::min(m, n));
    i64 Lrows = m, Ucols = n; // shape of the last factorised submatrix
    std::vector<double> sub;
    while (true) {
        if (leftorth)
            pushrandomsubset(J0, n, std::max<i64>(1, (i64)J0.size()), rng);
        else
            pushrandomsubset(I0, m, std::max<i64>(1, (i64)I0.size()), rng);
        for (int rookiter = 1; rookiter <= numrookiter; ++rookiter) {
            bool colmove = ((rookiter % 2 == 0) == leftorth);
            if (colmove)
                f(user, I0, lu.colperm, sub);
            else
                f(user, lu.rowperm, J0, sub);
            i64 sr = colmove ? (i64)I0.size() : m, sc = colmove ? n : (i64)J0.size();
            lu.npivot = 0;
            lu.m = sr;
            lu.n = sc;
            int rc = optimizerrlu(lu, sub.data(), sr, sc, maxrank, reltol, abstol);
            if (rc) return rc;
            Lrows = sr;
            Ucols = sc;
            islowrank = islowrank || lu.npivot < std::min(sr, sc);
            std::vector<i64> ri(lu.rowperm.begin(), lu.rowperm.begin() + lu.npivot);
            std::vector<i64> ci(lu.colperm.begin(), lu.colperm.begin() + lu.npivot);
            if (ri == I0 && ci == J0) break;
            J0 = ci;
            I0 = ri;
        }
        if (islowrank || (i64)I0.size() >= maxrank) break;
    }
    const i64 r = lu.npivot;
    std::vector<double> L11((size_t)(r * r)), U11((size_t)(r * r));
    for (i64 c = 0; c < r; ++c)
        for (i64 i = 0; i < r; ++i) {
            L11[i + c * r] = lu.L[i + c * Lrows];
            U11[i + c * r] = lu.U[i + c * r];
        }
    if (Lrows < m) { // :274-280
        std::vector<i64> I2;
        for (i64 v = 1; v <= m; ++v)
            if (std::find(I0.begin(), I0.end(), v) == I0.end()) I2.push_back(v);
        lu.rowperm = I0;
        lu.rowperm.insert(lu.rowperm.end(), I2.begin(), I2.end());
        std::vector<double> L2;
        if (!I2.empty() && !J0.empty()) f(user, I2, J0, L2);
        cols2Lmatrix(L2, (i64)I2.size(), U11, r);
        std::vector<double> L((size_t)(m * r));
        for (i64 c = 0; c < r; ++c) {
            for (i64 i = 0; i < r; ++i) L[i + c * m] = L11[i + c * r];
            for (i64 i = 0; i < (i64)I2.size(); ++i) L[r + i + c * m] = L2[i + c * (i64)I2.size()];
        }
        lu.L = L;
    }
    if (Ucols < n) { // :282-288
        std::vector<i64> J2;
        for (i64 v = 1; v <= n; ++v)
            if (std::find(J0.begin(), J0.end(), v) == J0.end()) J2.push_back(v);
        lu.colperm = J0;
        lu.colperm.insert(lu.colperm.end(), J2.begin(), J2.end());
        std::vector<double> U2;
        if (!J2.empty() && !I0.empty()) f(user, I0, J2, U2);
        rows2Umatrix(U2, (i64)J2.size(), L11, r);
        std::vector<double> U((size_t)(r * n));
        for (i64 c = 0; c < r; ++c)
            for (i64 i = 0; i < r; ++i) U[i + c * r] = U11[i + c * r];
        for (i64 c = 0; c < (i64)J2.size(); ++c)
            for (i64 i = 0; i < r; ++i) U[i + (r + c) * r] = U2[i + c * r];
        lu.U = U;
    }
    lu.m = m;
    lu.n = n;
    return 0;
}

// ============================================================ targets ======
struct TT3 { // TensorTrain{Float64,3}: cores (Dl, d, Dr) column-major
    std::vector<std::vector<double>> cores;
    std::vector<i64> dl, d, dr;
    i64 n() const { return (i64)cores.size(); }
};

// abstracttensortrain.jl:124-132 -- left-to-right product of T[:, i, :]
static double tt_evaluate(const TT3 &tt, const i64 *idx)
{
    std::vector<double> v(1, 1.0), w;
    for (i64 s = 0; s < tt.n(); ++s) {
        i64 Dl = tt.dl[s], d = tt.d[s], Dr = tt.dr[s];
        const double *T = tt.cores[s].data() + (idx[s] - 1) * Dl;
        w.assign(Dr, 0.0);
        for (i64 b = 0; b < Dr; ++b) {
            double acc = 0.0;
            for (i64 a = 0; a < Dl; ++a) acc += v[a] * T[a + b * Dl * d];
            w[b] = acc;
        }
        v.swap(w);
    }
    return v[0];
}

// abstracttensortrain.jl:164-199 (dims = all sites)
static double tt_sum(const TT3 &tt)
{
    std::vector<double> v(1, 1.0), w;
    for (i64 s = 0; s < tt.n(); ++s) {
        i64 Dl = tt.dl[s], d = tt.d[s], Dr = tt.dr[s];
        std::vector<double> S((size_t)(Dl * Dr), 0.0); // sum(T, dims=2)
        for (i64 b = 0; b < Dr; ++b)
            for (i64 a = 0; a < Dl; ++a) {
                double acc = 0.0;
                for (i64 x = 0; x < d; ++x) acc += tt.cores[s][a + x * Dl + b * Dl * d];
                S[a + b * Dl] = acc;
            }
        w.assign(Dr, 0.0);
        for (i64 b = 0; b < Dr; ++b) {
            double acc = 0.0;
            for (i64 a = 0; a < Dl; ++a) acc += v[a] * S[a + b * Dl];
            w[b] = acc;
        }
        v.swap(w);
    }
    return v[0];
}

struct MPO4 { // TensorTrain{Float64,4}: cores (Dl, s1, s2, Dr)
    std::vector<std::vector<double>> cores;
    std::vector<i64> dl, s1, s2, dr;
    i64 n() const { return (i64)cores.size(); }
};

struct Target {
    int kind = 0; // 0 analytic, 1 TTCache, 2 Contraction
    i64 nsites = 0;
    std::vector<i64> localdims;
    // analytic
    tci_analytic_t an{};
    std::vector<double> params;
    // TTCache (cachedtensortrain.jl:9-30)
    TT3 tt;
    std::vector<std::map<MultiIndex, std::vector<double>>> cacheleft, cacheright;
    // Contraction (contraction.jl:5-62)
    MPO4 A, B;
    std::map<MultiIndex, std::vector<double>> cl, cr; // keyed by fused indices
    i64 nevals = 0;
};

// ---- TTCache ----  cachedtensortrain.jl:77-98
static const std::vector<double> &ttc_left(Target &t, const i64 *idx, i64 ell)
{
    static thread_local std::vector<double> one;
    if (ell == 0) {
        one.assign(1, 1.0);
        return one;
    }
    MultiIndex key(idx, idx + ell);
    auto &cache = t.cacheleft[ell - 1];
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const TT3 &tt = t.tt;
    i64 s = ell - 1, Dl = tt.dl[s], d = tt.d[s], Dr = tt.dr[s];
    std::vector<double> prev = ttc_left(t, idx, ell - 1); // copy: recursion may rehash
    std::vector<double> out(Dr, 0.0);
    const double *T = tt.cores[s].data() + (idx[s] - 1) * Dl;
    for (i64 b = 0; b < Dr; ++b) {
        double acc = 0.0;
        for (i64 a = 0; a < Dl; ++a) acc += prev[a] * T[a + b * Dl * d];
        out[b] = acc;
    }
    return cache.emplace(key, std::move(out)).first->second;
}

// cachedtensortrain.jl:100-121; idx points at the first of `len` trailing indices
static const std::vector<double> &ttc_right(Target &t, const i64 *idx, i64 len)
{
    static thread_local std::vector<double> one;
    if (len == 0) {
        one.assign(1, 1.0);
        return one;
    }
    const TT3 &tt = t.tt;
    i64 s = tt.n() - len;
    MultiIndex key(idx, idx + len);
    auto &cache = t.cacheright[s];
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    i64 Dl = tt.dl[s], d = tt.d[s], Dr = tt.dr[s];
    std::vector<double> next = ttc_right(t, idx + 1, len - 1);
    std::vector<double> out(Dl, 0.0);
    const double *T = tt.cores[s].data() + (idx[0] - 1) * Dl;
    for (i64 a = 0; a < Dl; ++a) {
        double acc = 0.0;
        for (i64 b = 0; b < Dr; ++b) acc += T[a + b * Dl * d] * next[b];
        out[a] = acc;
    }
    return cache.emplace(key, std::move(out)).first->second;
}

// cachedtensortrain.jl:130-146 (usecache=true): dot at midpoint div(n,2)
static double ttc_evaluate(Target &t, const i64 *idx)
{
    i64 n = t.tt.n(), mid = n / 2;
    std::vector<double> l = ttc_left(t, idx, mid);
    const std::vector<double> &r = ttc_right(t, idx + mid, n - mid);
    double acc = 0.0;
    for (size_t a = 0; a < l.size(); ++a) acc += l[a] * r[a];
    return acc;
}

// ---- Contraction ---- contraction.jl:95-101
static inline void unfuse(const Target &t, i64 site, i64 idx, i64 &i, i64 &j)
{
    i64 d1 = t.A.s1[site];
    i = (idx - 1) % d1 + 1;
    j = (idx - 1) / d1 + 1;
}

// contraction.jl:103-109 with explicit strides so that the same routine serves
// the left (cores as stored) and right (cores with link indices exchanged,
// permutedims(a,(4,2,3,1)) :164-166) environments.
//   old: (La x Lb);  a(l, i, s, r) ; b(l, s, j, r)
static std::vector<double> extend_cache(const std::vector<double> &old, i64 La, i64 Lb, const double *a, i64 a_sl,
                                        i64 a_ss, i64 a_sr, i64 Lan, const double *b, i64 b_sl, i64 b_ss, i64 b_sr,
                                        i64 Lbn, i64 S)
{
    // tmp1[bq, s, an] = sum_l old[l, bq] * a[l, s, an]
    std::vector<double> tmp((size_t)(Lb * S * Lan), 0.0);
    for (i64 an = 0; an < Lan; ++an)
        for (i64 s = 0; s < S; ++s)
            for (i64 bq = 0; bq < Lb; ++bq) {
                double acc = 0.0;
                for (i64 l = 0; l < La; ++l) acc += old[l + bq * La] * a[l * a_sl + s * a_ss + an * a_sr];
                tmp[bq + Lb * (s + S * an)] = acc;
            }
    // out[an, bn] = sum_{bq,s} tmp1[bq, s, an] * b[bq, s, bn]
    std::vector<double> out((size_t)(Lan * Lbn), 0.0);
    for (i64 bn = 0; bn < Lbn; ++bn)
        for (i64 an = 0; an < Lan; ++an) {
            double acc = 0.0;
            for (i64 s = 0; s < S; ++s)
                for (i64 bq = 0; bq < Lb; ++bq)
                    acc += tmp[bq + Lb * (s + S * an)] * b[bq * b_sl + s * b_ss + bn * b_sr];
            out[an + Lan * bn] = acc;
        }
    return out;
}

// contraction.jl:112-139 ; fused indices idx[0..ell)
static std::vector<double> con_left(Target &t, const i64 *idx, i64 ell)
{
    if (ell == 0) return std::vector<double>(1, 1.0);
    MultiIndex key(idx, idx + ell);
    if (ell > 1) {
        auto it = t.cl.find(key);
        if (it != t.cl.end()) return it->second;
    }
    i64 s = ell - 1;
    i64 i, j;
    unfuse(t, s, idx[s], i, j);
    i64 La = t.A.dl[s], Lb = t.B.dl[s], Lan = t.A.dr[s], Lbn = t.B.dr[s], S = t.A.s2[s];
    i64 d1 = t.A.s1[s], d3 = t.B.s2[s];
    (void)d3;
    std::vector<double> prev = con_left(t, idx, ell - 1);
    const double *a = t.A.cores[s].data() + (i - 1) * La; // a[l,i,s,r] = a[l + La*(i + d1*(s + S*r))]
    const double *b = t.B.cores[s].data() + (j - 1) * Lb * S;
    std::vector<double> out = extend_cache(prev, La, Lb, a, 1, La * d1, La * d1 * S, Lan, b, 1, Lb, Lb * S * t.B.s2[s],
                                           Lbn, S);
    if (ell > 1) t.cl.emplace(key, out);
    return out;
}

// contraction.jl:144-176 ; idx points at the first of len trailing fused indices
static std::vector<double> con_right(Target &t, const i64 *idx, i64 len)
{
    if (len == 0) return std::vector<double>(1, 1.0);
    MultiIndex key(idx, idx + len);
    if (len > 1) {
        auto it = t.cr.find(key);
        if (it != t.cr.end()) return it->second;
    }
    i64 n = t.A.n(), s = n - len;
    i64 i, j;
    unfuse(t, s, idx[0], i, j);
    i64 La = t.A.dl[s], Lb = t.B.dl[s], Lan = t.A.dr[s], Lbn = t.B.dr[s], S = t.A.s2[s];
    i64 d1 = t.A.s1[s];
    std::vector<double> next = con_right(t, idx + 1, len - 1); // (Lan x Lbn)
    const double *a = t.A.cores[s].data() + (i - 1) * La;
    const double *b = t.B.cores[s].data() + (j - 1) * Lb * S;
    // roles of left/right links exchanged
    std::vector<double> out = extend_cache(next, Lan, Lbn, a, La * d1 * S, La * d1, 1, La, b, Lb * S * t.B.s2[s], Lb,
                                           1, Lb, S);
    if (len > 1) t.cr.emplace(key, out);
    return out;
}

// contraction.jl:189-207
static double con_evaluate(Target &t, const i64 *idx)
{
    i64 n = t.A.n(), mid = n / 2;
    std::vector<double> l = con_left(t, idx, mid);
    std::vector<double> r = con_right(t, idx + mid, n - mid);
    double acc = 0.0;
    for (size_t a = 0; a < l.size(); ++a) acc += l[a] * r[a];
    return acc;
}

static double target_eval(Target &t, const i64 *idx)
{
    t.nevals++;
    switch (t.kind) {
    case 0: return tci_target_eval(&t.an, idx);
    case 1: return ttc_evaluate(t, idx);
    default: return con_evaluate(t, idx);
    }
}

// ---------------------------------------------------------- batch eval ----
// batcheval.jl:32-61 -- generic triple loop; layout (left fastest, centre with
// first centre index fastest, right slowest).
static void batch_generic(Target &t, const IndexList &I, const IndexList &J, i64 M, std::vector<double> &out)
{
    i64 nI = (i64)I.size(), nJ = (i64)J.size();
    i64 nl = (i64)I[0].size(), nr = (i64)J[0].size();
    i64 C = 1;
    for (i64 c = 0; c < M; ++c) C *= t.localdims[nl + c];
    out.assign((size_t)(nI * C * nJ), 0.0);
    MultiIndex idx(nl + M + nr);
    for (i64 i = 0; i < nI; ++i)
        for (i64 c = 0; c < C; ++c) {
            i64 rem = c;
            for (i64 q = 0; q < M; ++q) {
                idx[nl + q] = rem % t.localdims[nl + q] + 1;
                rem /= t.localdims[nl + q];
            }
            for (i64 j = 0; j < nJ; ++j) {
                for (i64 q = 0; q < nl; ++q) idx[q] = I[i][q];
                for (i64 q = 0; q < nr; ++q) idx[nl + M + q] = J[j][q];
                out[i + nI * (c + C * j)] = target_eval(t, idx.data());
            }
        }
}

// cachedtensortrain.jl:151-215 (projector = nothing)
static void batch_ttcache(Target &t, const IndexList &I, const IndexList &J, i64 M, std::vector<double> &out)
{
    const TT3 &tt = t.tt;
    i64 N = tt.n(), nI = (i64)I.size(), nJ = (i64)J.size();
    i64 nl = (i64)I[0].size(), nr = (i64)J[0].size();
    i64 DL = (nl > 0 && nl < N) ? tt.dr[nl - 1] : 1;
    std::vector<double> lenv((size_t)(nI * DL), 1.0);
    if (nl > 0)
        for (i64 i = 0; i < nI; ++i) {
            const std::vector<double> &e = ttc_left(t, I[i].data(), nl);
            for (i64 a = 0; a < DL; ++a) lenv[i + nI * a] = e[a];
        }
    i64 DR = (nr > 0 && nr < N) ? tt.dr[nl + M - 1] : 1;
    std::vector<double> renv((size_t)(DR * nJ), 1.0);
    if (nr > 0)
        for (i64 j = 0; j < nJ; ++j) {
            const std::vector<double> &e = ttc_right(t, J[j].data(), nr);
            for (i64 a = 0; a < DR; ++a) renv[a + DR * j] = e[a];
        }
    i64 rows = nI;
    for (i64 s = nl; s < nl + M; ++s) { // :198-208
        i64 D = tt.dl[s], d = tt.d[s], Dn = tt.dr[s];
        std::vector<double> nxt((size_t)(rows * d * Dn));
        gemm_seq(lenv.data(), rows, tt.cores[s].data(), D, nxt.data(), rows, rows, d * Dn, D);
        lenv.swap(nxt);
        rows *= d;
    }
    out.assign((size_t)(rows * nJ), 0.0);
    gemm_seq(lenv.data(), rows, renv.data(), DR, out.data(), rows, rows, nJ, DR); // :211-212
}

// contraction.jl:236-335 (projector = nothing, f = nothing)
static void batch_contraction(Target &t, const IndexList &I, const IndexList &J, i64 M, std::vector<double> &out)
{
    i64 N = t.A.n(), nI = (i64)I.size(), nJ = (i64)J.size();
    i64 nl = (i64)I[0].size(), nr = (i64)J[0].size();
    i64 s0 = nl, e0 = N - nr; // centre sites [s0, e0)
    i64 La = (s0 == 0) ? 1 : t.A.dr[s0 - 1], Lb = (s0 == 0) ? 1 : t.B.dr[s0 - 1];
    std::vector<double> left((size_t)(nI * La * Lb));
    for (i64 i = 0; i < nI; ++i) { // :270-273
        std::vector<double> e = con_left(t, I[i].data(), nl);
        for (i64 q = 0; q < La * Lb; ++q) left[i + nI * q] = e[q];
    }
    i64 Ra = (e0 == N) ? 1 : t.A.dl[e0], Rb = (e0 == N) ? 1 : t.B.dl[e0];
    std::vector<double> right((size_t)(Ra * Rb * nJ));
    for (i64 j = 0; j < nJ; ++j) { // :276-284
        std::vector<double> e = con_right(t, J[j].data(), nr);
        for (i64 q = 0; q < Ra * Rb; ++q) right[q + Ra * Rb * j] = e[q];
    }
    // leftobj (nI, La, Lb, S)
    std::vector<double> lo = left;
    i64 S = 1;
    (void)M;
    for (i64 s = s0; s < e0; ++s) { // :290-317
        i64 d1 = t.A.s1[s], sh = t.A.s2[s], d3 = t.B.s2[s], Lan = t.A.dr[s], Lbn = t.B.dr[s];
        const std::vector<double> &a = t.A.cores[s];
        const std::vector<double> &b = t.B.cores[s];
        // tmp1 (nI, Lb, S, d1, sh, Lan) = sum_la lo[i, la, lb, S] a[la, d1, sh, lan]
        std::vector<double> tmp1((size_t)(nI * Lb * S * d1 * sh * Lan), 0.0);
        for (i64 lan = 0; lan < Lan; ++lan)
            for (i64 h = 0; h < sh; ++h)
                for (i64 x = 0; x < d1; ++x)
                    for (i64 q = 0; q < S; ++q)
                        for (i64 lb = 0; lb < Lb; ++lb)
                            for (i64 i = 0; i < nI; ++i) {
                                double acc = 0.0;
                                for (i64 la = 0; la < La; ++la)
                                    acc += lo[i + nI * (la + La * (lb + Lb * q))] *
                                           a[la + La * (x + d1 * (h + sh * lan))];
                                tmp1[i + nI * (lb + Lb * (q + S * (x + d1 * (h + sh * lan))))] = acc;
                            }
        // tmp2 (nI, S, d1, Lan, d3, Lbn) = sum_{lb,h} tmp1[i, lb, S, d1, h, lan] b[lb, h, d3, lbn]
        // then permuted to (nI, Lan, Lbn, S, d1, d3)
        std::vector<double> nxt((size_t)(nI * Lan * Lbn * S * d1 * d3), 0.0);
        for (i64 lbn = 0; lbn < Lbn; ++lbn)
            for (i64 z = 0; z < d3; ++z)
                for (i64 lan = 0; lan < Lan; ++lan)
                    for (i64 x = 0; x < d1; ++x)
                        for (i64 q = 0; q < S; ++q)
                            for (i64 i = 0; i < nI; ++i) {
                                double acc = 0.0;
                                for (i64 h = 0; h < sh; ++h)
                                    for (i64 lb = 0; lb < Lb; ++lb)
                                        acc += tmp1[i + nI * (lb + Lb * (q + S * (x + d1 * (h + sh * lan))))] *
                                               b[lb + Lb * (h + sh * (z + d3 * lbn))];
                                nxt[i + nI * (lan + Lan * (lbn + Lbn * (q + S * (x + d1 * z))))] = acc;
                            }
        lo.swap(nxt);
        La = Lan;
        Lb = Lbn;
        S *= d1 * d3;
    }
    // res[i, S, j] = sum_{a,b} lo[i,a,b,S] right[a,b,j]   :328
    out.assign((size_t)(nI * S * nJ), 0.0);
    for (i64 j = 0; j < nJ; ++j)
        for (i64 q = 0; q < S; ++q)
            for (i64 i = 0; i < nI; ++i) {
                double acc = 0.0;
                for (i64 ab = 0; ab < La * Lb; ++ab) acc += lo[i + nI * (ab + La * Lb * q)] * right[ab + La * Lb * j];
                out[i + nI * (q + S * j)] = acc;
            }
}

// filltensor / _batchevaluate_dispatch  tensorci2.jl:290-312, batcheval.jl:32-83
static void filltensor(Target &t, const IndexList &I, const IndexList &J, i64 M, std::vector<double> &out)
{
    out.clear();
    if (I.empty() || J.empty()) return;
    switch (t.kind) {
    case 0: batch_generic(t, I, J, M, out); break;
    case 1: batch_ttcache(t, I, J, M, out); break;
    default: batch_contraction(t, I, J, M, out); break;
    }
}

// ========================================================= TCI2 driver =====
struct TCI2 { // tensorci2.jl:6-40
    i64 n = 0;
    std::vector<i64> localdims;
    std::vector<IndexList> Iset, Jset;
    std::vector<std::vector<double>> sitetensors; // (|Iset[b]|, d_b, |Jset[b]| or |Iset[b+1]|)
    std::vector<i64> tdl, tdr;
    std::vector<double> pivoterrors, bonderrors;
    double maxsamplevalue = 0.0;
    std::vector<IndexList> Ihist_last, Jhist_last; // only history[end] is ever read (:874-877)
    bool has_hist = false;
    std::vector<i64> ranks, nglobalpivots;
    std::vector<double> errors; // already divided by the normalisation
    // trace of every 2-site bond update, for bond-by-bond parity checks
    std::vector<i64> trace; // (iter, bond, m, n, npivot) quintuples
    RookRng rook;           // injected random subsets of the rook search
};

static void pushunique(IndexList &c, const MultiIndex &x)
{
    if (std::find(c.begin(), c.end(), x) == c.end()) c.push_back(x);
}

// tensorci2.jl:193-213
static void addglobalpivots(TCI2 &tci, const IndexList &pivots)
{
    for (const MultiIndex &p : pivots)
        for (i64 b = 0; b < tci.n; ++b) {
            pushunique(tci.Iset[b], MultiIndex(p.begin(), p.begin() + b));
            pushunique(tci.Jset[b], MultiIndex(p.begin() + b + 1, p.end()));
        }
    if (!pivots.empty())
        for (auto &T : tci.sitetensors) T.clear();
}

// tensorci2.jl:315-327
static IndexList kron_left(const IndexList &I, i64 d)
{
    IndexList out;
    for (i64 j = 1; j <= d; ++j)
        for (const MultiIndex &is : I) {
            MultiIndex v = is;
            v.push_back(j);
            out.push_back(v);
        }
    return out;
}
static IndexList kron_right(i64 d, const IndexList &J)
{
    IndexList out;
    for (const MultiIndex &js : J)
        for (i64 i = 1; i <= d; ++i) {
            MultiIndex v;
            v.push_back(i);
            v.insert(v.end(), js.begin(), js.end());
            out.push_back(v);
        }
    return out;
}
// Base.union(a, b): order-preserving, duplicates dropped
static IndexList union_lists(const IndexList &a, const IndexList &b)
{
    std::map<MultiIndex, int> seen;
    IndexList out;
    for (const IndexList *l : {&a, &b})
        for (const MultiIndex &x : *l)
            if (seen.emplace(x, 1).second) out.push_back(x);
    return out;
}

static void updatemaxsample(TCI2 &tci, const std::vector<double> &v)
{ // util.jl:1-10
    double m = tci.maxsamplevalue;
    for (double x : v) m = jl_max(std::fabs(m), std::fabs(x));
    tci.maxsamplevalue = m;
}

static void updateerrors(TCI2 &tci, i64 b, const std::vector<double> &e)
{ // tensorci2.jl:143-169
    tci.bonderrors[b] = e.back();
    size_t L = std::max(tci.pivoterrors.size(), e.size());
    std::vector<double> out(L, 0.0);
    for (size_t i = 0; i < L; ++i) {
        double a = i < tci.pivoterrors.size() ? tci.pivoterrors[i] : 0.0;
        double c = i < e.size() ? e[i] : 0.0;
        out[i] = jl_max(a, c);
    }
    tci.pivoterrors = out;
}

// SubMatrix (tensorci2.jl:476-503): lazily evaluated Pi with a running max |value|
struct SubMatrixCtx {
    Target *f;
    const IndexList *rows, *cols;
    double maxsamplevalue = 0.0;
};
static void submatrix_eval(void *user, const std::vector<i64> &ir, const std::vector<i64> &ic, std::vector<double> &out)
{
    SubMatrixCtx *c = static_cast<SubMatrixCtx *>(user);
    IndexList I, J;
    for (i64 i : ir) I.push_back((*c->rows)[i - 1]);
    for (i64 j : ic) J.push_back((*c->cols)[j - 1]);
    filltensor(*c->f, I, J, 0, out);
    double mx = -INFINITY; // maximum(abs, res)
    for (double v : out) mx = jl_max(mx, std::fabs(v));
    if (!out.empty()) c->maxsamplevalue = jl_max(c->maxsamplevalue, mx);
}

// tensorci2.jl:510-607 ; pivotsearch 0 = :full, 1 = :rook
static int updatepivots(TCI2 &tci, Target &f, i64 b, bool leftorth, double reltol, double abstol, i64 maxbonddim,
                        const IndexList &extraI, const IndexList &extraJ, i64 iter, int pivotsearch = 0)
{
    for (auto &T : tci.sitetensors) T.clear();
    IndexList Ic = union_lists(kron_left(tci.Iset[b], tci.localdims[b]), extraI);
    IndexList Jc = union_lists(kron_right(tci.localdims[b + 1], tci.Jset[b + 1]), extraJ);
    RRLU lu;
    bool need_full = true;
    if (pivotsearch == 1) { // :552-595
        std::vector<i64> I0, J0;
        for (const MultiIndex &x : tci.Iset[b + 1]) {
            auto it = std::find(Ic.begin(), Ic.end(), x);
            if (it != Ic.end()) I0.push_back((i64)(it - Ic.begin()) + 1);
        }
        for (const MultiIndex &x : tci.Jset[b]) {
            auto it = std::find(Jc.begin(), Jc.end(), x);
            if (it != Jc.end()) J0.push_back((i64)(it - Jc.begin()) + 1);
        }
        SubMatrixCtx sm{&f, &Ic, &Jc};
        int rc = arrlu(lu, submatrix_eval, &sm, (i64)Ic.size(), (i64)Jc.size(), I0, J0, maxbonddim, reltol, abstol,
                       leftorth, 5, tci.rook);
        if (rc) return rc;
        tci.maxsamplevalue = jl_max(std::fabs(tci.maxsamplevalue), std::fabs(sm.maxsamplevalue)); // :569
        need_full = lu.npivot == 0; // fall back to the full search if the rook search fails (:573-588)
    }
    if (need_full) {
        std::vector<double> Pi;
        filltensor(f, Ic, Jc, 0, Pi);
        updatemaxsample(tci, Pi);
        int rc = rrlu(lu, Pi.data(), (i64)Ic.size(), (i64)Jc.size(), maxbonddim, reltol, abstol, leftorth);
        if (rc) return rc;
    }
    IndexList In, Jn;
    for (i64 k = 0; k < lu.npivot; ++k) In.push_back(Ic[lu.rowperm[k] - 1]);
    for (i64 k = 0; k < lu.npivot; ++k) Jn.push_back(Jc[lu.colperm[k] - 1]);
    tci.Iset[b + 1] = In;
    tci.Jset[b] = Jn;
    // left/right site tensors (:601-604) are invalidated again before anybody reads them
    updateerrors(tci, b, pivoterrors(lu));
    i64 tr[5] = {iter, b + 1, (i64)Ic.size(), (i64)Jc.size(), lu.npivot};
    tci.trace.insert(tci.trace.end(), tr, tr + 5);
    return 0;
}

// Solve X * P = B for X (B: rows x k, P: k x k): Tmat = transpose(transpose(P) \ transpose(Pi1))
// tensorci2.jl:391.  Partial-pivoting LU of P^T (LAPACK getrf stand-in), then two
// triangular solves per right-hand side.
static int solve_right(const std::vector<double> &B, i64 rows, const std::vector<double> &P, i64 k,
                       std::vector<double> &X)
{
    std::vector<double> Mx((size_t)(k * k));
    for (i64 i = 0; i < k; ++i)
        for (i64 j = 0; j < k; ++j) Mx[i + j * k] = P[j + i * k]; // M = P^T
    std::vector<i64> piv(k);
    for (i64 c = 0; c < k; ++c) {
        i64 p = c;
        double best = std::fabs(Mx[c + c * k]);
        for (i64 i = c + 1; i < k; ++i)
            if (std::fabs(Mx[i + c * k]) > best) {
                best = std::fabs(Mx[i + c * k]);
                p = i;
            }
        piv[c] = p;
        if (p != c)
            for (i64 j = 0; j < k; ++j) std::swap(Mx[c + j * k], Mx[p + j * k]);
        double d = Mx[c + c * k];
        for (i64 i = c + 1; i < k; ++i) Mx[i + c * k] /= d;
        for (i64 j = c + 1; j < k; ++j) {
            double u = Mx[c + j * k];
            for (i64 i = c + 1; i < k; ++i) Mx[i + j * k] -= Mx[i + c * k] * u;
        }
    }
    X.assign((size_t)(rows * k), 0.0);
    std::vector<double> y(k);
    for (i64 r = 0; r < rows; ++r) { // rhs = B[r, :]^T
        for (i64 i = 0; i < k; ++i) y[i] = B[r + i * rows];
        for (i64 c = 0; c < k; ++c) std::swap(y[c], y[piv[c]]);
        for (i64 i = 0; i < k; ++i) {
            double acc = y[i];
            for (i64 j = 0; j < i; ++j) acc -= Mx[i + j * k] * y[j];
            y[i] = acc;
        }
        for (i64 i = k - 1; i >= 0; --i) {
            double acc = y[i];
            for (i64 j = i + 1; j < k; ++j) acc -= Mx[i + j * k] * y[j];
            y[i] = acc / Mx[i + i * k];
        }
        for (i64 i = 0; i < k; ++i) X[r + i * rows] = y[i];
    }
    return 0;
}

// tensorci2.jl:367-394
static int setsitetensor(TCI2 &tci, Target &f, i64 b)
{
    std::vector<double> Pi1;
    filltensor(f, tci.Iset[b], tci.Jset[b], 1, Pi1);
    updatemaxsample(tci, Pi1);
    i64 nI = (i64)tci.Iset[b].size(), d = tci.localdims[b], nJ = (i64)tci.Jset[b].size();
    if (b == tci.n - 1) {
        tci.sitetensors[b] = Pi1;
        tci.tdl[b] = nI;
        tci.tdr[b] = nJ;
        return 0;
    }
    std::vector<double> P;
    filltensor(f, tci.Iset[b + 1], tci.Jset[b], 0, P);
    i64 k = (i64)tci.Iset[b + 1].size();
    if (k != nJ) {
        g_err = "Pivot matrix at bond " + std::to_string(b + 1) + " is not square!";
        return 3;
    }
    std::vector<double> X;
    solve_right(Pi1, nI * d, P, k, X);
    tci.sitetensors[b] = X;
    tci.tdl[b] = nI;
    tci.tdr[b] = k;
    return 0;
}

// sweepstrategies.jl:1-6 ; strategy 0 = :backandforth, 1 = :forward, 2 = :backward
static bool forwardsweep(int strategy, i64 iter) { return strategy == 1 || (strategy == 0 && (iter % 2 == 1)); }

// tensorci2.jl:855-916
static int sweep2site(TCI2 &tci, Target &f, i64 niter, i64 iter1, double abstol, i64 maxbonddim, int strategy,
                      bool strictlynested, i64 outer_iter, int pivotsearch = 0)
{
    for (auto &T : tci.sitetensors) T.clear();
    i64 n = tci.n;
    for (i64 iter = iter1; iter < iter1 + niter; ++iter) {
        std::vector<IndexList> extraI(n), extraJ(n);
        if (!strictlynested && tci.has_hist) {
            extraI = tci.Ihist_last;
            extraJ = tci.Jhist_last;
        }
        tci.Ihist_last = tci.Iset;
        tci.Jhist_last = tci.Jset;
        tci.has_hist = true;
        tci.pivoterrors.clear();
        if (forwardsweep(strategy, iter)) {
            for (i64 b = 0; b < n - 1; ++b) {
                int rc = updatepivots(tci, f, b, true, 1e-14, abstol, maxbonddim, extraI[b + 1], extraJ[b],
                                      outer_iter * 100 + iter, pivotsearch);
                if (rc) return rc;
            }
        } else {
            for (i64 b = n - 2; b >= 0; --b) {
                int rc = updatepivots(tci, f, b, false, 1e-14, abstol, maxbonddim, extraI[b + 1], extraJ[b],
                                      outer_iter * 100 + iter, pivotsearch);
                if (rc) return rc;
            }
        }
    }
    for (i64 b = 0; b < n; ++b) { // fillsitetensors! globalsearch.jl:97-103
        int rc = setsitetensor(tci, f, b);
        if (rc) return rc;
    }
    return 0;
}

// tensorci2.jl:402-461 (forward, updatetensors = true)
static int sweep1site(TCI2 &tci, Target &f, double reltol, double abstol, i64 maxbonddim)
{
    tci.pivoterrors.clear();
    for (auto &T : tci.sitetensors) T.clear();
    i64 n = tci.n;
    for (i64 b = 0; b < n - 1; ++b) {
        IndexList Is = kron_left(tci.Iset[b], tci.localdims[b]);
        const IndexList &Js = tci.Jset[b];
        std::vector<double> Pi;
        filltensor(f, tci.Iset[b], tci.Jset[b], 1, Pi);
        updatemaxsample(tci, Pi);
        RRLU lu;
        int rc = rrlu(lu, Pi.data(), (i64)Is.size(), (i64)Js.size(), maxbonddim, reltol, abstol, true);
        if (rc) return rc;
        IndexList In, Jn;
        for (i64 k = 0; k < lu.npivot; ++k) In.push_back(Is[lu.rowperm[k] - 1]);
        for (i64 k = 0; k < lu.npivot; ++k) Jn.push_back(Js[lu.colperm[k] - 1]);
        i64 nIb = (i64)tci.Iset[b].size();
        tci.Iset[b + 1] = In;
        tci.Jset[b] = Jn;
        luci_left(lu, tci.sitetensors[b]);
        tci.tdl[b] = nIb;
        tci.tdr[b] = lu.npivot;
        for (double v : tci.sitetensors[b])
            if (std::isnan(v)) {
                g_err = "Error: NaN in tensor T[" + std::to_string(b + 1) + "]";
                return 4;
            }
        updateerrors(tci, b, pivoterrors(lu));
    }
    std::vector<double> last;
    filltensor(f, tci.Iset[n - 1], tci.Jset[n - 1], 1, last);
    tci.sitetensors[n - 1] = last;
    tci.tdl[n - 1] = (i64)tci.Iset[n - 1].size();
    tci.tdr[n - 1] = 1;
    return 0;
}

static TT3 tt_from_tci(const TCI2 &tci)
{
    TT3 tt;
    for (i64 b = 0; b < tci.n; ++b) {
        tt.cores.push_back(tci.sitetensors[b]);
        tt.dl.push_back(tci.tdl[b]);
        tt.d.push_back(tci.localdims[b]);
        tt.dr.push_back(tci.tdr[b]);
    }
    return tt;
}

// Injected replacement of rand(rng, 1:d) (globalpivotfinder.jl:156): counter based.
static i64 start_point(uint64_t seed, i64 iter, i64 s, i64 p, i64 d)
{
    uint64_t idx = ((uint64_t)iter * 1000003ull + (uint64_t)s) * 1009ull + (uint64_t)p;
    i64 v = 1 + (i64)(tci_uniform01(seed, idx) * (double)d);
    return v > d ? d : v;
}

// globalpivotfinder.jl:143-195 with explicit start points (n x nsearch, column-major)
static IndexList default_finder(Target &f, const TT3 &tt, const std::vector<i64> &localdims, const i64 *starts,
                                i64 nsearch, double abstol, double tolmargin, i64 maxn, std::vector<double> *errs)
{
    i64 L = (i64)localdims.size();
    IndexList found;
    for (i64 s = 0; s < nsearch; ++s) {
        MultiIndex point(starts + s * L, starts + (s + 1) * L);
        MultiIndex cur = point, best = point;
        double best_error = 0.0;
        for (i64 p = 0; p < L; ++p) {
            for (i64 v = 1; v <= localdims[p]; ++v) {
                cur[p] = v;
                double e = std::fabs(target_eval(f, cur.data()) - tt_evaluate(tt, cur.data()));
                if (e > best_error) {
                    best_error = e;
                    best = cur;
                }
            }
            cur[p] = point[p];
        }
        if (best_error > abstol * tolmargin) {
            found.push_back(best);
            if (errs) errs->push_back(best_error);
        }
    }
    if ((i64)found.size() > maxn) {
        found.resize(maxn);
        if (errs) errs->resize(maxn);
    }
    return found;
}

// tensorci2.jl:609-628
static bool convergencecriterion(const std::vector<i64> &ranks, const std::vector<double> &errors,
                                 const std::vector<i64> &ngp, double tol, i64 maxbonddim, i64 nhist, bool checkgp)
{
    if ((i64)errors.size() < nhist) return false;
    size_t L = ranks.size();
    bool allerr = true, allgp = true, allmax = true;
    i64 minrank = INT64_MAX;
    for (size_t q = L - nhist; q < L; ++q) {
        allerr &= errors[q] < tol;
        allgp &= ngp[q] == 0;
        allmax &= ranks[q] >= maxbonddim;
        minrank = std::min(minrank, ranks[q]);
    }
    return (allerr && (checkgp ? allgp : true) && minrank == ranks[L - 1]) || allmax;
}

struct Options { // tensorci2.jl:700-720
    double tolerance = 1e-8;
    i64 maxbonddim = INT64_MAX;
    i64 maxiter = 20;
    int sweepstrategy = 0;
    int normalizeerror = 1;
    i64 ncheckhistory = 3;
    i64 maxnglobalpivot = 5;
    i64 nsearchglobalpivot = 5;
    double tolmarginglobalsearch = 10.0;
    int strictlynested = 0;
    int checkconvglobalpivot = 1;
    uint64_t seed = 1;
    int pivotsearch = 0; // 0 :full, 1 :rook
};

// tensorci2.jl:42-53 + 700-850
static int crossinterpolate2(TCI2 &tci, Target &f, const std::vector<i64> &localdims, const IndexList &initialpivots,
                             const Options &o)
{
    i64 n = (i64)localdims.size();
    if (n < 2) {
        g_err = "localdims should have at least 2 elements!";
        return 5;
    }
    tci = TCI2();
    tci.n = n;
    tci.localdims = localdims;
    tci.Iset.resize(n);
    tci.Jset.resize(n);
    tci.sitetensors.resize(n);
    tci.tdl.assign(n, 0);
    tci.tdr.assign(n, 0);
    tci.bonderrors.assign(n - 1, 0.0);
    tci.rook.seed = o.seed;
    addglobalpivots(tci, initialpivots);
    double ms = 0.0;
    for (const MultiIndex &p : initialpivots) ms = jl_max(ms, std::fabs(target_eval(f, p.data())));
    tci.maxsamplevalue = ms;
    if (!(std::fabs(ms) > 0.0)) {
        g_err = "maxsamplevalue is zero!";
        return 6;
    }
    if (o.nsearchglobalpivot > 0 && o.nsearchglobalpivot < o.maxnglobalpivot) {
        g_err = "nsearchglobalpivot < maxnglobalpivot!";
        return 7;
    }
    if (o.maxbonddim >= INT64_MAX && o.tolerance <= 0) {
        g_err = "Specify either tolerance > 0 or some maxbonddim; otherwise, the convergence criterion is not "
                "reachable!";
        return 8;
    }
    std::vector<double> errors_abs;
    for (i64 iter = 1; iter <= o.maxiter; ++iter) {
        double norm = o.normalizeerror ? tci.maxsamplevalue : 1.0;
        double abstol = o.tolerance * norm;
        int rc = sweep2site(tci, f, 2, 1, abstol, o.maxbonddim, o.sweepstrategy, o.strictlynested != 0, iter,
                            o.pivotsearch);
        if (rc) return rc;
        double pe = -INFINITY;
        for (double e : tci.bonderrors) pe = jl_max(pe, e);
        errors_abs.push_back(pe);
        TT3 tt = tt_from_tci(tci);
        std::vector<i64> starts((size_t)(n * o.nsearchglobalpivot));
        for (i64 s = 0; s < o.nsearchglobalpivot; ++s)
            for (i64 p = 0; p < n; ++p) starts[p + s * n] = start_point(o.seed, iter, s, p, localdims[p]);
        IndexList gp = default_finder(f, tt, localdims, starts.data(), o.nsearchglobalpivot, abstol,
                                      o.tolmarginglobalsearch, o.maxnglobalpivot, nullptr);
        addglobalpivots(tci, gp);
        tci.nglobalpivots.push_back((i64)gp.size());
        i64 rk = 0;
        for (i64 b = 0; b < n - 1; ++b) rk = std::max(rk, (i64)tci.Iset[b + 1].size());
        tci.ranks.push_back(rk);
        if (convergencecriterion(tci.ranks, errors_abs, tci.nglobalpivots, abstol, o.maxbonddim, o.ncheckhistory,
                                 o.checkconvglobalpivot != 0))
            break;
    }
    double norm = o.normalizeerror ? tci.maxsamplevalue : 1.0;
    double abstol = o.tolerance * norm;
    int rc = sweep1site(tci, f, 1e-14, abstol, o.maxbonddim);
    if (rc) return rc;
    for (i64 b = 0; b < n - 1; ++b)
        if (tci.Iset[b + 1].size() != tci.Jset[b].size()) {
            g_err = "Pivot matrix at bond " + std::to_string(b + 1) + " is not square!";
            return 3;
        }
    tci.errors.clear();
    for (double e : errors_abs) tci.errors.push_back(e / norm);
    return 0;
}

// ================================================================ C API ====
static IndexList unflatten(const i64 *flat, i64 len, i64 count)
{
    IndexList out((size_t)count);
    for (i64 i = 0; i < count; ++i) out[i].assign(flat + i * len, flat + (i + 1) * len);
    return out;
}

extern "C" {

const char *orc_last_error(void) { return g_err.c_str(); }

// A: m x n column-major (not modified).  Outputs: rowperm[m], colperm[n] (1-based),
// L (m x maxr), U (maxr x n) with maxr = min(maxrank,m,n) leading dimension = npivot on
// return (tightly packed m x r and r x n), pivoterrors[min(m,n)+1].
int orc_rrlu(const double *A, i64 m, i64 n, i64 maxrank, double reltol, double abstol, int leftorthogonal,
             i64 *rowperm, i64 *colperm, i64 *npivot, double *error, double *L, double *U, double *pe)
{
    RRLU lu;
    int rc = rrlu(lu, A, m, n, maxrank, reltol, abstol, leftorthogonal != 0);
    if (rc) return rc;
    std::copy(lu.rowperm.begin(), lu.rowperm.end(), rowperm);
    std::copy(lu.colperm.begin(), lu.colperm.end(), colperm);
    *npivot = lu.npivot;
    *error = lu.error;
    if (L) std::copy(lu.L.begin(), lu.L.end(), L);
    if (U) std::copy(lu.U.begin(), lu.U.end(), U);
    if (pe) {
        std::vector<double> e = pivoterrors(lu);
        std::copy(e.begin(), e.end(), pe);
    }
    return 0;
}

// MatrixLUCI(A; ...) then left/right (matrixluci.jl:5-7, 70-84); left m x r, right r x n.
int orc_luci(const double *A, i64 m, i64 n, i64 maxrank, double reltol, double abstol, int leftorthogonal,
             i64 *rowperm, i64 *colperm, i64 *npivot, double *pe, double *left, double *right)
{
    RRLU lu;
    int rc = rrlu(lu, A, m, n, maxrank, reltol, abstol, leftorthogonal != 0);
    if (rc) return rc;
    std::copy(lu.rowperm.begin(), lu.rowperm.end(), rowperm);
    std::copy(lu.colperm.begin(), lu.colperm.end(), colperm);
    *npivot = lu.npivot;
    std::vector<double> e = pivoterrors(lu);
    if (pe) std::copy(e.begin(), e.end(), pe);
    std::vector<double> l, r;
    if (left) {
        luci_left(lu, l);
        std::copy(l.begin(), l.end(), left);
    }
    if (right) {
        luci_right(lu, r);
        std::copy(r.begin(), r.end(), right);
    }
    return 0;
}

// ---- ComplexF64: matrixlu.jl:1-32, 98-181 on a Matrix{ComplexF64}, element by element with Julia
// Base's complex arithmetic (include/tci_zarith.h: abs2, abs = hypot, robust division, multiply-then-subtract).  A is
// m x n column-major, interleaved (re, im) -- the memory of a Julia Matrix{ComplexF64}; L (m x r) and U (r x n) come
// back the same way (2*m*maxr and 2*maxr*n doubles), pe = [abs.(diag); lu.error] (:394-416).
// Returns 0, 1 ("lu.L contains NaNs") or 2 ("lu.U contains NaNs").
int orc_zrrlu(const double *Ain, i64 m, i64 n, i64 maxrank, double reltol, double abstol, int leftorthogonal,
              i64 *rowperm, i64 *colperm, i64 *npivot, double *error, double *L, double *U, double *pe)
{
    std::vector<tci_z> A((size_t)(m * n));
    for (i64 e = 0; e < m * n; ++e) A[e] = tci_zmake(Ain[2 * e], Ain[2 * e + 1]);
    for (i64 i = 0; i < m; ++i) rowperm[i] = i + 1;
    for (i64 j = 0; j < n; ++j) colperm[j] = j + 1;
    maxrank = std::min(maxrank, std::min(m, n));
    double maxerror = 0.0, err = NAN;
    i64 np = 0;
    std::vector<double> diag;
    while (np < maxrank) { // _optimizerrlu! :150-160
        const i64 k = np;
        double best = -INFINITY; // submatrixargmax(abs2, A, k) :1-32: columns outer, rows inner, strict >
        i64 pr = k, pc = k;
        for (i64 c = k; c < n; ++c)
            for (i64 r = k; r < m; ++r) {
                const double v = tci_zabs2(A[r + c * m]);
                if (v > best) {
                    best = v;
                    pr = r;
                    pc = c;
                }
            }
        err = tci_zabs(A[pr + pc * m]);
        if ((std::fabs(err) < reltol * maxerror || std::fabs(err) < abstol) && np > 0) break;
        maxerror = jl_max(maxerror, err);
        np++; // addpivot! :114-136
        std::swap(rowperm[k], rowperm[pr]);
        for (i64 j = 0; j < n; ++j) std::swap(A[k + j * m], A[pr + j * m]);
        std::swap(colperm[k], colperm[pc]);
        for (i64 i = 0; i < m; ++i) std::swap(A[i + k * m], A[i + pc * m]);
        const tci_z piv = A[k + k * m];
        diag.push_back(tci_zabs(piv));
        if (leftorthogonal) {
            for (i64 i = k + 1; i < m; ++i) A[i + k * m] = tci_zdiv(A[i + k * m], piv);
        } else {
            for (i64 j = k + 1; j < n; ++j) A[k + j * m] = tci_zdiv(A[k + j * m], piv);
        }
        for (i64 j = k + 1; j < n; ++j) {
            const tci_z yj = A[k + j * m];
            for (i64 i = k + 1; i < m; ++i) A[i + j * m] = tci_zsub(A[i + j * m], tci_zmul(A[i + k * m], yj));
        }
    }
    const i64 r = np;
    std::vector<tci_z> Lz((size_t)(m * r), tci_zmake(0.0, 0.0)), Uz((size_t)(r * n), tci_zmake(0.0, 0.0));
    for (i64 c = 0; c < r; ++c)
        for (i64 i = c; i < m; ++i) Lz[i + c * m] = A[i + c * m]; // tril(A[:, 1:r])
    for (i64 c = 0; c < n; ++c)
        for (i64 i = 0; i <= std::min(c, r - 1); ++i) Uz[i + c * r] = A[i + c * m]; // triu(A[1:r, :])
    for (const tci_z &v : Lz)
        if (std::isnan(v.re) || std::isnan(v.im)) {
            g_err = "lu.L contains NaNs";
            return 1;
        }
    for (const tci_z &v : Uz)
        if (std::isnan(v.re) || std::isnan(v.im)) {
            g_err = "lu.U contains NaNs";
            return 2;
        }
    for (i64 c = 0; c < r; ++c) {
        if (leftorthogonal)
            Lz[c + c * m] = tci_zmake(1.0, 0.0);
        else
            Uz[c + c * r] = tci_zmake(1.0, 0.0);
    }
    if (r >= std::min(m, n)) err = 0.0;
    *npivot = r;
    *error = err;
    if (L) std::memcpy(L, Lz.data(), Lz.size() * sizeof(tci_z));
    if (U) std::memcpy(U, Uz.data(), Uz.size() * sizeof(tci_z));
    if (pe) {
        for (i64 i = 0; i < r; ++i) pe[i] = diag[i];
        pe[r] = err;
    }
    return 0;
}

static void dense_eval(void *user, const std::vector<i64> &ir, const std::vector<i64> &ic, std::vector<double> &out)
{
    const double *A = static_cast<const double *const *>(user)[0];
    i64 m = (i64)(intptr_t) static_cast<const double *const *>(user)[1];
    out.resize(ir.size() * ic.size());
    for (size_t j = 0; j < ic.size(); ++j)
        for (size_t i = 0; i < ir.size(); ++i) out[i + j * ir.size()] = A[(ir[i] - 1) + (ic[j] - 1) * m];
}

// arrlu on a dense matrix (test_matrixlu.jl:71-86); I0/J0 1-based, nI0/nJ0 may be 0
int orc_arrlu(const double *A, i64 m, i64 n, const i64 *I0, i64 nI0, const i64 *J0, i64 nJ0, i64 maxrank, double reltol,
              double abstol, int leftorthogonal, uint64_t seed, i64 *rowperm, i64 *colperm, i64 *npivot, double *error,
              double *L, double *U)
{
    RRLU lu;
    RookRng rng;
    rng.seed = seed;
    const void *user[2] = {A, (const void *)(intptr_t)m};
    int rc = arrlu(lu, dense_eval, (void *)user, m, n, std::vector<i64>(I0, I0 + nI0), std::vector<i64>(J0, J0 + nJ0),
                   maxrank, reltol, abstol, leftorthogonal != 0, 5, rng);
    if (rc) return rc;
    std::copy(lu.rowperm.begin(), lu.rowperm.end(), rowperm);
    std::copy(lu.colperm.begin(), lu.colperm.end(), colperm);
    *npivot = lu.npivot;
    *error = lu.error;
    if (L) std::copy(lu.L.begin(), lu.L.end(), L);
    if (U) std::copy(lu.U.begin(), lu.U.end(), U);
    return 0;
}

void orc_argmax_abs2(const double *A, i64 m, i64 n, i64 k1, i64 *row, i64 *col)
{
    i64 r, c;
    submatrixargmax_abs2(A, m, n, k1 - 1, r, c);
    *row = r + 1;
    *col = c + 1;
}

Target *orc_target_builtin(int kind, const double *params, i64 nparams, const i64 *localdims, i64 nsites)
{
    Target *t = new Target();
    t->kind = 0;
    t->nsites = nsites;
    t->localdims.assign(localdims, localdims + nsites);
    if (kind == TCI_TARGET_TABLE) { // prepend strides
        double st = 1.0;
        for (i64 k = 0; k < nsites; ++k) {
            t->params.push_back(st);
            st *= (double)localdims[k];
        }
    }
    t->params.insert(t->params.end(), params, params + nparams);
    t->an.kind = kind;
    t->an.nsites = (int)nsites;
    t->an.nparams = (i64)t->params.size();
    t->an.params = t->params.data();
    t->an.localdims = t->localdims.data();
    t->an.nstate = tci_target_nstate(kind, t->params.data());
    return t;
}

Target *orc_tt_create(i64 nsites, const i64 *dims3, const double *const *cores)
{
    Target *t = new Target();
    t->kind = 1;
    t->nsites = nsites;
    for (i64 s = 0; s < nsites; ++s) {
        i64 Dl = dims3[3 * s], d = dims3[3 * s + 1], Dr = dims3[3 * s + 2];
        t->tt.dl.push_back(Dl);
        t->tt.d.push_back(d);
        t->tt.dr.push_back(Dr);
        t->tt.cores.emplace_back(cores[s], cores[s] + Dl * d * Dr);
        t->localdims.push_back(d);
    }
    t->cacheleft.resize(nsites);
    t->cacheright.resize(nsites);
    return t;
}

Target *orc_mpo_pair_create(i64 nsites, const i64 *dimsA4, const double *const *A, const i64 *dimsB4,
                            const double *const *B)
{
    Target *t = new Target();
    t->kind = 2;
    t->nsites = nsites;
    for (i64 s = 0; s < nsites; ++s) {
        const i64 *da = dimsA4 + 4 * s, *db = dimsB4 + 4 * s;
        t->A.dl.push_back(da[0]);
        t->A.s1.push_back(da[1]);
        t->A.s2.push_back(da[2]);
        t->A.dr.push_back(da[3]);
        t->A.cores.emplace_back(A[s], A[s] + da[0] * da[1] * da[2] * da[3]);
        t->B.dl.push_back(db[0]);
        t->B.s1.push_back(db[1]);
        t->B.s2.push_back(db[2]);
        t->B.dr.push_back(db[3]);
        t->B.cores.emplace_back(B[s], B[s] + db[0] * db[1] * db[2] * db[3]);
        t->localdims.push_back(da[1] * db[2]);
    }
    return t;
}

void orc_target_destroy(Target *t) { delete t; }

double orc_eval_point(Target *t, const i64 *idx) { return target_eval(*t, idx); }

// I: nl x nI column-major (multi-index i is I[nl*i .. nl*i+nl)); out sized nI*C*nJ.
int orc_pi_eval(Target *t, const i64 *I, i64 nl, i64 nI, const i64 *J, i64 nr, i64 nJ, i64 M, double *out,
                double *maxabs)
{
    if (nI * nJ == 0) return 0;
    if (nl + M + nr != t->nsites) {
        g_err = "Invalid number of central indices";
        return 9;
    }
    std::vector<double> o;
    filltensor(*t, unflatten(I, nl, nI), unflatten(J, nr, nJ), M, o);
    std::copy(o.begin(), o.end(), out);
    if (maxabs) {
        double m = *maxabs;
        for (double x : o) m = jl_max(std::fabs(m), std::fabs(x));
        *maxabs = m;
    }
    return 0;
}

double orc_tt_evaluate(i64 nsites, const i64 *dims3, const double *const *cores, const i64 *idx)
{
    std::unique_ptr<Target> t(orc_tt_create(nsites, dims3, cores));
    return tt_evaluate(t->tt, idx);
}
double orc_tt_sum(i64 nsites, const i64 *dims3, const double *const *cores)
{
    std::unique_ptr<Target> t(orc_tt_create(nsites, dims3, cores));
    return tt_sum(t->tt);
}

// Default global pivot finder with injected start points (n x nsearch).
int orc_globalsearch(Target *f, i64 nsites, const i64 *dims3, const double *const *cores, const i64 *starts,
                     i64 nsearch, double abstol, double tolmargin, i64 maxn, i64 *pivots_out, double *errs_out,
                     i64 *nfound)
{
    std::unique_ptr<Target> t(orc_tt_create(nsites, dims3, cores));
    std::vector<double> errs;
    IndexList found = default_finder(*f, t->tt, t->localdims, starts, nsearch, abstol, tolmargin, maxn, &errs);
    *nfound = (i64)found.size();
    for (size_t q = 0; q < found.size(); ++q) {
        std::copy(found[q].begin(), found[q].end(), pivots_out + q * nsites);
        errs_out[q] = errs[q];
    }
    return 0;
}

void orc_start_points(uint64_t seed, i64 iter, i64 nsearch, const i64 *localdims, i64 n, i64 *out)
{
    for (i64 s = 0; s < nsearch; ++s)
        for (i64 p = 0; p < n; ++p) out[p + s * n] = start_point(seed, iter, s, p, localdims[p]);
}

int orc_convergencecriterion(const i64 *ranks, const double *errors, const i64 *ngp, i64 len, double tol,
                             i64 maxbonddim, i64 nhist, int checkgp)
{
    return convergencecriterion(std::vector<i64>(ranks, ranks + len), std::vector<double>(errors, errors + len),
                                std::vector<i64>(ngp, ngp + len), tol, maxbonddim, nhist, checkgp != 0)
               ? 1
               : 0;
}

// ---- driver ----
struct orc_options {
    double tolerance;
    i64 maxbonddim;
    i64 maxiter;
    int sweepstrategy;
    int normalizeerror;
    i64 ncheckhistory;
    i64 maxnglobalpivot;
    i64 nsearchglobalpivot;
    double tolmarginglobalsearch;
    int strictlynested;
    int checkconvglobalpivot;
    uint64_t seed;
    int pivotsearch;
};

TCI2 *orc_crossinterpolate2(Target *f, const i64 *localdims, i64 n, const i64 *pivots, i64 npivots,
                            const orc_options *opt, int *status)
{
    Options o;
    o.tolerance = opt->tolerance;
    o.maxbonddim = opt->maxbonddim;
    o.maxiter = opt->maxiter;
    o.sweepstrategy = opt->sweepstrategy;
    o.normalizeerror = opt->normalizeerror;
    o.ncheckhistory = opt->ncheckhistory;
    o.maxnglobalpivot = opt->maxnglobalpivot;
    o.nsearchglobalpivot = opt->nsearchglobalpivot;
    o.tolmarginglobalsearch = opt->tolmarginglobalsearch;
    o.strictlynested = opt->strictlynested;
    o.checkconvglobalpivot = opt->checkconvglobalpivot;
    o.seed = opt->seed;
    o.pivotsearch = opt->pivotsearch;
    TCI2 *tci = new TCI2();
    *status = crossinterpolate2(*tci, *f, std::vector<i64>(localdims, localdims + n), unflatten(pivots, n, npivots),
                                o);
    return tci;
}
void orc_tci_destroy(TCI2 *t) { delete t; }
i64 orc_tci_niter(TCI2 *t) { return (i64)t->ranks.size(); }
void orc_tci_history(TCI2 *t, i64 *ranks, double *errors, i64 *ngp)
{
    std::copy(t->ranks.begin(), t->ranks.end(), ranks);
    std::copy(t->errors.begin(), t->errors.end(), errors);
    std::copy(t->nglobalpivots.begin(), t->nglobalpivots.end(), ngp);
}
double orc_tci_maxsamplevalue(TCI2 *t) { return t->maxsamplevalue; }
i64 orc_tci_npivoterrors(TCI2 *t) { return (i64)t->pivoterrors.size(); }
void orc_tci_pivoterrors(TCI2 *t, double *pe, double *be)
{
    std::copy(t->pivoterrors.begin(), t->pivoterrors.end(), pe);
    std::copy(t->bonderrors.begin(), t->bonderrors.end(), be);
}
// which = 0: Iset[b], 1: Jset[b] (b 0-based).  Returns count; if out != NULL writes len x count.
i64 orc_tci_indexset(TCI2 *t, int which, i64 b, i64 *out)
{
    const IndexList &l = which ? t->Jset[b] : t->Iset[b];
    if (out)
        for (size_t q = 0; q < l.size(); ++q) std::copy(l[q].begin(), l[q].end(), out + q * l[q].size());
    return (i64)l.size();
}
void orc_tci_coredims(TCI2 *t, i64 b, i64 *dims3)
{
    dims3[0] = t->tdl[b];
    dims3[1] = t->localdims[b];
    dims3[2] = t->tdr[b];
}
void orc_tci_core(TCI2 *t, i64 b, double *out)
{
    std::copy(t->sitetensors[b].begin(), t->sitetensors[b].end(), out);
}
i64 orc_tci_tracelen(TCI2 *t) { return (i64)t->trace.size() / 5; }
void orc_tci_trace(TCI2 *t, i64 *out) { std::copy(t->trace.begin(), t->trace.end(), out); }
double orc_tci_evaluate(TCI2 *t, const i64 *idx) { return tt_evaluate(tt_from_tci(*t), idx); }
double orc_tci_sum(TCI2 *t) { return tt_sum(tt_from_tci(*t)); }
i64 orc_target_nevals(Target *t) { return t->nevals; }

} // extern "C"
