// zpath.cu -- the ComplexF64 value type of the TensorCI2 path.
//
// The reference is generic in the value type; its contraction and conversion tests run on ComplexF64
// (test_contraction.jl:39-46, test_matrixlu.jl:39-52).  A Matrix{ComplexF64} lives in a tci_dmat as interleaved
// (re, im) pairs -- 2m rows of doubles, exactly Julia's memory layout -- and a factorisation handle carries
// is_complex.  Pieces:
//   * k_zrrlu<LEFT>: rrlu (matrixlu.jl:1-32, 98-181) on a complex matrix, one persistent cooperative kernel.  CTA g
//     owns the physical columns g, g+G, ...; rows are swapped physically inside the owner's columns, columns are
//     permuted virtually.  Every CTA posts, next to its candidate record, the scaled column of its candidate, so ONE
//     grid synchronisation per pivot is enough.  The arithmetic is Julia Base's (include/tci_zarith.h: abs2 metric,
//     hypot for the stop rule, the robust complex division, multiply-then-subtract, nothing fused), the same header
//     the CPU oracle compiles, so permutations, pivot errors, L and U are bit-identical to it.
//   * left / right of MatrixLUCI (matrixluci.jl:40-84), the solve of setsitetensor! (tensorci2.jl:391) and lu.L / lu.U
//     on complex factors: triangular solves with a thread per row / column, products through the complex DMMA GEMM.
//   * Pi / point evaluation of a complex MPO pair (contraction.jl:189-335) and of a complex tensor train, which is
//     stored as the pair (core as (Dl, d, 1, Dr), 1 x 1 x 1 x 1 unit cores) and runs through the same chains.
#include <cooperative_groups.h>

#include <algorithm>

#include "tci_internal.h"
#include "../../include/tci_zarith.h"

namespace cg = cooperative_groups;

#define ZR_THREADS 512
#define ZR_XCAP 8192 // pivot-column entries kept in shared memory (16 B each)
#define ZR_CH 512    // rows of a work item of the trailing update

struct __align__(16) ZCand {
    double val; // abs2 of the candidate (-inf: none)
    int row, pos;
    double pre, pim; // the candidate's value
    int col, pad;
    double pad2;
};

struct ZArgs {
    double2 *A;
    i64 ld; // in (re, im) pairs
    int m, n, maxrank;
    double reltol, abstol;
    ZCand *cand;   // [2][G]
    double2 *post; // [2][G][mpad]
    i64 mpad;
    i64 *rowperm, *colperm;
    int *colpos;
    int *result; // [0] npivot, [1] flags (1: L has NaNs, 2: U has NaNs), bytes 16..23: lu.error
    double *piv; // |pivot| per accepted pivot
};

__device__ __forceinline__ bool zbetter(double v, int pos, int row, double bv, int bpos, int brow)
{ // "columns outer, rows inner, strict >" of submatrixargmax as an order: value down, position up, row up
    return v > bv || (v == bv && (pos < bpos || (pos == bpos && row < brow)));
}
__device__ __forceinline__ tci_z zld(const double2 *p)
{
    const double2 v = *p;
    return tci_zmake(v.x, v.y);
}
__device__ __forceinline__ double jlmax(double a, double b) { return (isnan(a) || isnan(b)) ? NAN : (a > b ? a : b); }

template <bool LEFT> __global__ void __launch_bounds__(ZR_THREADS, 1) k_zrrlu(ZArgs a)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char zsm[];
    double2 *xs = reinterpret_cast<double2 *>(zsm);
    __shared__ double s_val[ZR_THREADS / 32];
    __shared__ int s_row[ZR_THREADS / 32], s_pos[ZR_THREADS / 32], s_col[ZR_THREADS / 32];
    __shared__ ZCand s_win;
    __shared__ int s_wg;
    const int G = gridDim.x, g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = ZR_THREADS / 32;
    const int m = a.m, n = a.n;
    const i64 ld = a.ld;
    double2 *A = a.A;

    for (int c = g + G * tid; c < n; c += G * ZR_THREADS) a.colpos[c] = c;
    if (g == 0)
        for (int i = tid; i < m; i += ZR_THREADS) a.rowperm[i] = i;
    __syncthreads();

    double maxerror = 0.0, err = NAN;
    int k = 0; // pivots accepted so far
    // bv.. : this thread's best candidate of the trailing block (rows >= k, positions >= k)
    double bv;
    int bpos, brow, bcol;
    auto reset_best = [&]() {
        bv = -INFINITY;
        bpos = brow = bcol = 0x7fffffff;
    };
    auto consider_default = [&](int pos, int c, int kk) { // an all-NaN block selects (kk, kk): matrixlu.jl:14-16
        if (pos == kk && lane == 0 && zbetter(-INFINITY, kk, kk, bv, bpos, brow)) {
            bv = -INFINITY;
            bpos = kk;
            brow = kk;
            bcol = c;
        }
    };
    // posts this CTA's candidate for step kk: record + its column over rows >= kk (scaled by the candidate if LEFT)
    auto post_candidate = [&](int kk) {
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int op = __shfl_xor_sync(0xffffffffu, bpos, o), orow = __shfl_xor_sync(0xffffffffu, brow, o),
                      oc = __shfl_xor_sync(0xffffffffu, bcol, o);
            if (zbetter(ov, op, orow, bv, bpos, brow) || (ov == bv && op == bpos && orow == brow && oc < bcol)) {
                bv = ov;
                bpos = op;
                brow = orow;
                bcol = oc;
            }
        }
        if (lane == 0) {
            s_val[warp] = bv;
            s_row[warp] = brow;
            s_pos[warp] = bpos;
            s_col[warp] = bcol;
        }
        __syncthreads();
        if (tid == 0) {
            double v = s_val[0];
            int r = s_row[0], p = s_pos[0], c = s_col[0];
            for (int w = 1; w < nwarps; ++w)
                if (zbetter(s_val[w], s_pos[w], s_row[w], v, p, r)) {
                    v = s_val[w];
                    r = s_row[w];
                    p = s_pos[w];
                    c = s_col[w];
                }
            ZCand cd;
            cd.val = v;
            cd.row = r;
            cd.pos = p;
            cd.col = c;
            cd.pad = 0;
            cd.pad2 = 0.0;
            cd.pre = cd.pim = 0.0;
            if (c != 0x7fffffff) {
                const double2 pv = A[r + ld * c];
                cd.pre = pv.x;
                cd.pim = pv.y;
            }
            s_win = cd;
        }
        __syncthreads();
        const ZCand cd = s_win;
        double2 *dst = a.post + ((i64)(kk & 1) * G + g) * a.mpad;
        if (cd.col != 0x7fffffff) {
            const tci_z pv = tci_zmake(cd.pre, cd.pim);
            const double2 *src = A + ld * cd.col;
            for (int i = kk + tid; i < m; i += ZR_THREADS) {
                tci_z v = zld(src + i);
                if (LEFT) v = tci_zdiv(v, pv);
                __stcg(dst + i, make_double2(v.re, v.im));
            }
        }
        if (tid == 0) {
            double4 *rec = reinterpret_cast<double4 *>(a.cand + (i64)(kk & 1) * G + g);
            const ZCand *s = &s_win;
            __stcg(reinterpret_cast<double2 *>(rec), *reinterpret_cast<const double2 *>(s));
            __stcg(reinterpret_cast<double2 *>(rec) + 1, *(reinterpret_cast<const double2 *>(s) + 1));
            __stcg(reinterpret_cast<double2 *>(rec) + 2, *(reinterpret_cast<const double2 *>(s) + 2));
        }
    };

    // first search: the whole matrix
    reset_best();
    {
        const int nown = (n - g + G - 1) / G, nch = (m + ZR_CH - 1) / ZR_CH;
        for (int item = warp; item < nown * nch; item += nwarps) {
            const int oc = item / nch, ch = item - oc * nch;
            const int c = g + G * oc, pos = c;
            if (ch == 0) consider_default(pos, c, 0);
            const double2 *col = A + ld * c;
            const int r1 = min(m, (ch + 1) * ZR_CH);
            for (int i = ch * ZR_CH + lane; i < r1; i += 32) {
                const double v = tci_zabs2(zld(col + i));
                if (v >= bv && zbetter(v, pos, i, bv, bpos, brow)) {
                    bv = v;
                    bpos = pos;
                    brow = i;
                    bcol = c;
                }
            }
        }
    }
    post_candidate(0);

    int stop = 0;
    while (k < a.maxrank) {
        grid.sync();
        // reduce the G candidates (every CTA redundantly, identical result)
        if (warp == 0) {
            double v = -INFINITY;
            int r = 0x7fffffff, p = 0x7fffffff, wg = -1;
            const ZCand *recs = a.cand + (i64)(k & 1) * G;
            for (int q = lane; q < G; q += 32) {
                const double2 h = __ldcg(reinterpret_cast<const double2 *>(recs + q));
                const double2 rp = __ldcg(reinterpret_cast<const double2 *>(recs + q) + 2);
                const int qr = (int)(__double_as_longlong(h.y) & 0xffffffffll), qp = (int)(__double_as_longlong(h.y) >> 32);
                const int qc = (int)(__double_as_longlong(rp.x) & 0xffffffffll);
                if (qc != 0x7fffffff && zbetter(h.x, qp, qr, v, p, r)) {
                    v = h.x;
                    r = qr;
                    p = qp;
                    wg = q;
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int orow = __shfl_xor_sync(0xffffffffu, r, o), op = __shfl_xor_sync(0xffffffffu, p, o),
                          og = __shfl_xor_sync(0xffffffffu, wg, o);
                if (og >= 0 && (wg < 0 || zbetter(ov, op, orow, v, p, r))) {
                    v = ov;
                    r = orow;
                    p = op;
                    wg = og;
                }
            }
            if (lane == 0) {
                s_wg = wg;
                const double2 *rec = reinterpret_cast<const double2 *>(recs + wg);
                ZCand w;
                const double2 h0 = __ldcg(rec), h1 = __ldcg(rec + 1), h2 = __ldcg(rec + 2);
                w.val = h0.x;
                w.row = (int)(__double_as_longlong(h0.y) & 0xffffffffll);
                w.pos = (int)(__double_as_longlong(h0.y) >> 32);
                w.col = (int)(__double_as_longlong(h2.x) & 0xffffffffll);
                w.pad = 0;
                w.pad2 = 0.0;
                w.pre = h1.x;
                w.pim = h1.y;
                s_win = w;
            }
        }
        __syncthreads();
        const ZCand w = s_win;
        const int wg = s_wg;
        const tci_z piv = tci_zmake(w.pre, w.pim);
        err = tci_zabs(piv); // matrixlu.jl:153
        if (k > 0 && (fabs(err) < a.reltol * maxerror || fabs(err) < a.abstol)) { // :155
            stop = 1;
            break;
        }
        maxerror = jlmax(maxerror, err);
        const int pr = w.row, pw = w.pos, cw = w.col;
        if (g == 0 && tid == 0) {
            a.piv[k] = err;
            const i64 t = a.rowperm[k];
            a.rowperm[k] = a.rowperm[pr];
            a.rowperm[pr] = t;
        }
        // the pivot column for the rows below k (after the row swap, the old row k sits at row pr)
        const double2 *px = a.post + ((i64)(k & 1) * G + wg) * a.mpad;
        for (int i = k + 1 + tid; i < m && i < ZR_XCAP; i += ZR_THREADS) xs[i] = __ldcg(px + (i == pr ? k : i));
        // physical row swap inside the own columns (full width, matrixlu.jl:98-104), new positions, the pivot row
        for (int c = g + G * tid; c < n; c += G * ZR_THREADS) {
            double2 *col = A + ld * c;
            if (pr != k) {
                const double2 t = col[k];
                col[k] = col[pr];
                col[pr] = t;
            }
            int pos = a.colpos[c];
            if (c == cw)
                pos = k;
            else if (pos == k)
                pos = pw;
            a.colpos[c] = pos;
            if (!LEFT && pos > k) { // the row of U is scaled instead of the column of L (:122-124)
                const tci_z y = tci_zdiv(zld(col + k), piv);
                col[k] = make_double2(y.re, y.im);
            }
        }
        __syncthreads();
        if (LEFT && g == cw % G) { // the owner stores the scaled column of L (:120-121)
            double2 *col = A + ld * cw;
            for (int i = k + 1 + tid; i < m; i += ZR_THREADS) col[i] = i < ZR_XCAP ? xs[i] : __ldcg(px + (i == pr ? k : i));
        }
        k++;
        if (k >= a.maxrank) break;
        // trailing update a - x*y (:132) fused with the search of the next pivot.  Work items = (own column, chunk of
        // ZR_CH rows), dealt round-robin to the warps: a CTA owns only ~n/G columns, so whole columns per warp would
        // leave most warps idle.  zbetter is a total order, so the visiting order does not matter.
        reset_best();
        {
            const int nown = (n - g + G - 1) / G, nch = (m - k + ZR_CH - 1) / ZR_CH;
            for (int item = warp; item < nown * nch; item += nwarps) {
                const int oc = item / nch, ch = item - oc * nch;
                const int c = g + G * oc;
                const int pos = a.colpos[c];
                if (pos < k) continue;
                if (ch == 0) consider_default(pos, c, k);
                double2 *col = A + ld * c;
                const tci_z y = zld(col + (k - 1));
                const int r1 = min(m, k + (ch + 1) * ZR_CH);
                int i = k + ch * ZR_CH + lane;
                for (; i + 96 < r1; i += 128) { // four independent 16-byte loads in flight per lane
                    double2 a4[4], x4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) a4[u] = col[i + 32 * u];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = i + 32 * u;
                        x4[u] = r < ZR_XCAP ? xs[r] : __ldcg(px + (r == pr ? k - 1 : r));
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = i + 32 * u;
                        const tci_z v = tci_zsub(tci_zmake(a4[u].x, a4[u].y), tci_zmul(tci_zmake(x4[u].x, x4[u].y), y));
                        col[r] = make_double2(v.re, v.im);
                        const double v2 = tci_zabs2(v);
                        if (v2 >= bv && zbetter(v2, pos, r, bv, bpos, brow)) {
                            bv = v2;
                            bpos = pos;
                            brow = r;
                            bcol = c;
                        }
                    }
                }
                for (; i < r1; i += 32) {
                    const double2 xv = i < ZR_XCAP ? xs[i] : __ldcg(px + (i == pr ? k - 1 : i));
                    const tci_z v = tci_zsub(zld(col + i), tci_zmul(tci_zmake(xv.x, xv.y), y));
                    col[i] = make_double2(v.re, v.im);
                    const double v2 = tci_zabs2(v);
                    if (v2 >= bv && zbetter(v2, pos, i, bv, bpos, brow)) {
                        bv = v2;
                        bpos = pos;
                        brow = i;
                        bcol = c;
                    }
                }
            }
        }
        __syncthreads(); // the candidate's column is read back by other warps
        post_candidate(k);
    }
    (void)stop;
    const int r = k;
    if (r >= (m < n ? m : n)) err = 0.0; // :176-178
    // NaN checks on L = tril(A[:, 1:r]) and U = triu(A[1:r, :]) before the unit diagonal is written (:162-169)
    int flags = 0;
    for (int c = g + G * warp; c < n; c += G * nwarps) {
        const int pos = a.colpos[c];
        const double2 *col = A + ld * c;
        if (lane == 0) a.colperm[pos] = c;
        if (pos < r)
            for (int i = pos + lane; i < m; i += 32) {
                const double2 v = col[i];
                if (isnan(v.x) || isnan(v.y)) flags |= 1;
            }
        const int top = pos < r - 1 ? pos : r - 1;
        for (int i = lane; i <= top; i += 32) {
            const double2 v = col[i];
            if (isnan(v.x) || isnan(v.y)) flags |= 2;
        }
    }
    if (flags) atomicOr(a.result + 1, flags);
    if (g == 0 && tid == 0) {
        a.result[0] = r;
        *reinterpret_cast<double *>(a.result + 4) = err;
    }
}

// lu.L (m x r) / lu.U (r x n) in position order with the unit diagonal (matrixlu.jl:162-174)
__global__ void k_zextract_L(const double2 *__restrict__ A, i64 m, i64 ld, const i64 *__restrict__ colperm, int r, int leftorth,
                             double2 *__restrict__ L, i64 ldl)
{
    const i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= m * r) return;
    const i64 i = e % m, c = e / m;
    double2 v = make_double2(0.0, 0.0);
    if (i > c)
        v = A[i + ld * colperm[c]];
    else if (i == c)
        v = leftorth ? make_double2(1.0, 0.0) : A[i + ld * colperm[c]];
    L[i + ldl * c] = v;
}
__global__ void k_zextract_U(const double2 *__restrict__ A, i64 n, i64 ld, const i64 *__restrict__ colperm, int r, int leftorth,
                             double2 *__restrict__ U, i64 ldu)
{
    const i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= (i64)r * n) return;
    const i64 i = e % r, q = e / r;
    double2 v = make_double2(0.0, 0.0);
    if (q > i)
        v = A[i + ld * colperm[q]];
    else if (q == i)
        v = leftorth ? A[i + ld * colperm[q]] : make_double2(1.0, 0.0);
    U[i + ldu * q] = v;
}

__device__ __forceinline__ double2 zfms(double2 acc, double2 x, double2 y)
{ // acc - x*y   (BLAS-level arithmetic: the reference pins these solves to sqrt(eps) only)
    acc.x = fma(-x.x, y.x, fma(x.y, y.y, acc.x));
    acc.y = fma(-x.x, y.y, fma(-x.y, y.x, acc.y));
    return acc;
}
// W (rows x k) <- W * T^-1, T (k x k) upper (columns solved left to right) or lower (right to left) triangular;
// a thread per row, T read as warp-wide broadcasts
template <bool UPPER, bool UNIT>
__global__ void k_ztrsm_right(double2 *__restrict__ W, i64 rows, i64 ldw, const double2 *__restrict__ T, i64 ldt, int k)
{
    const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (i >= rows) return;
    for (int q = 0; q < k; ++q) {
        const int c = UPPER ? q : k - 1 - q;
        double2 acc = W[i + ldw * c];
        if (UPPER)
            for (int j = 0; j < c; ++j) acc = zfms(acc, W[i + ldw * j], T[j + ldt * c]);
        else
            for (int j = c + 1; j < k; ++j) acc = zfms(acc, W[i + ldw * j], T[j + ldt * c]);
        if (!UNIT) {
            const double2 d = T[c + ldt * c];
            const tci_z qv = tci_zdiv(tci_zmake(acc.x, acc.y), tci_zmake(d.x, d.y));
            acc = make_double2(qv.re, qv.im);
        }
        W[i + ldw * c] = acc;
    }
}
// X (k x cols) <- U^-1 X for a unit upper-triangular U; a thread per column
__global__ void k_ztrsm_left_upper_unit(double2 *__restrict__ X, i64 cols, i64 ldx, const double2 *__restrict__ U, i64 ldu, int k)
{
    const i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (j >= cols) return;
    double2 *x = X + ldx * j;
    for (int i = k - 1; i >= 0; --i) {
        double2 acc = x[i];
        for (int q = i + 1; q < k; ++q) acc = zfms(acc, U[i + ldu * q], x[q]);
        x[i] = acc;
    }
}
// dst[perm[i], :] = src[i, :]; `unit_top`: the first r rows of src are the identity (colstimespivotinv)
__global__ void k_zscatter_rows(const double2 *__restrict__ src, i64 lds, i64 m, i64 r, const i64 *__restrict__ perm,
                                double2 *__restrict__ dst, i64 ldd, int unit_top)
{
    const i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= m * r) return;
    const i64 i = e % m, c = e / m;
    double2 v = src[i + lds * c];
    if (unit_top && i < r) v = make_double2(i == c ? 1.0 : 0.0, 0.0);
    dst[perm[i] + ldd * c] = v;
}
__global__ void k_zscatter_cols(const double2 *__restrict__ src, i64 lds, i64 r, i64 n, const i64 *__restrict__ perm,
                                double2 *__restrict__ dst, i64 ldd, int unit_left)
{
    const i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= r * n) return;
    const i64 i = e % r, q = e / r;
    double2 v = src[i + lds * q];
    if (unit_left && q < r) v = make_double2(i == q ? 1.0 : 0.0, 0.0);
    dst[i + ldd * perm[q]] = v;
}
__global__ void k_zgather_cols(const double2 *__restrict__ src, i64 lds, i64 rows, i64 n, const i64 *__restrict__ perm,
                               double2 *__restrict__ dst, i64 ldd)
{
    const i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= rows * n) return;
    const i64 i = e % rows, q = e / rows;
    dst[i + ldd * q] = src[i + lds * perm[q]];
}

static int zlu_extract(tci_lu *lu, double2 *dL, i64 ldl, double2 *dU, i64 ldu)
{
    tci_ctx *ctx = lu->ctx;
    const i64 m = lu->m, n = lu->n;
    const int r = (int)lu->r;
    if (r == 0) return TCI_OK;
    const double2 *A = reinterpret_cast<const double2 *>(lu->A->p);
    const i64 ld = lu->A->ld / 2;
    if (dL) {
        k_zextract_L<<<(unsigned)((m * r + 255) / 256), 256, 0, ctx->stream>>>(A, m, ld, lu->d_colperm, r, lu->leftorthogonal, dL, ldl);
        ctx->launches++;
    }
    if (dU) {
        k_zextract_U<<<(unsigned)(((i64)r * n + 255) / 256), 256, 0, ctx->stream>>>(A, n, ld, lu->d_colperm, r, lu->leftorthogonal, dU, ldu);
        ctx->launches++;
    }
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

// rrlu(A::Matrix{ComplexF64}; ...) -- A (2m x n doubles) is factorised in place
static int zrrlu_core(tci_ctx *ctx, tci_dmat *A, i64 m, i64 n, i64 maxrank, double reltol, double abstol, int leftorthogonal,
                      i64 *rowperm, i64 *colperm, i64 *npivot, double *error, double *pivoterrors, tci_lu **factors)
{
    dmat_wait_ready(ctx, A);
    const i64 mn = std::min(m, n);
    const i64 mr = (maxrank <= 0 || maxrank > mn) ? mn : maxrank;
    for (i64 i = 0; i < m; ++i) rowperm[i] = i + 1;
    for (i64 j = 0; j < n; ++j) colperm[j] = j + 1;
    *npivot = 0;
    if (mr == 0) {
        *error = 0.0;
        if (pivoterrors) pivoterrors[0] = 0.0;
        return TCI_OK;
    }
    const void *fn = leftorthogonal ? (const void *)k_zrrlu<true> : (const void *)k_zrrlu<false>;
    const size_t smem = (size_t)std::min<i64>(m, ZR_XCAP) * sizeof(double2) + 16;
    TCI_CUDA(ctx, ctx_func_smem(ctx, fn, (int)(ZR_XCAP * sizeof(double2) + 16)));
    int per_sm = 0;
    TCI_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, ZR_THREADS, smem));
    if (per_sm < 1) return tci_fail(ctx, TCI_ERR_CUDA, "k_zrrlu does not fit an SM");
    const int G = (int)std::max<i64>(1, std::min<i64>(ctx->sm_count, n));
    const i64 mpad = round_up(m, 8);
    // arena: result words | pivots | rowperm | colperm | colpos | records | posted columns
    const size_t o_piv = 64, o_rp = o_piv + round_up(mr, 2) * 8, o_cp = o_rp + (size_t)m * 8, o_cpos = o_cp + (size_t)n * 8,
                 o_back = o_cpos, o_cand = round_up((i64)(o_cpos + (size_t)n * 4), 64),
                 o_post = o_cand + sizeof(ZCand) * 2 * (size_t)G, total = o_post + sizeof(double2) * 2 * (size_t)G * (size_t)mpad;
    DevBuf<char> arena(ctx);
    TCI_CUDA(ctx, arena.alloc(total));
    TCI_CUDA(ctx, cudaMemsetAsync(arena.p, 0, 64, ctx->stream));
    ZArgs args;
    args.A = reinterpret_cast<double2 *>(A->p);
    args.ld = A->ld / 2;
    args.m = (int)m;
    args.n = (int)n;
    args.maxrank = (int)mr;
    args.reltol = reltol;
    args.abstol = abstol;
    args.cand = reinterpret_cast<ZCand *>(arena.p + o_cand);
    args.post = reinterpret_cast<double2 *>(arena.p + o_post);
    args.mpad = mpad;
    args.rowperm = reinterpret_cast<i64 *>(arena.p + o_rp);
    args.colperm = reinterpret_cast<i64 *>(arena.p + o_cp);
    args.colpos = reinterpret_cast<int *>(arena.p + o_cpos);
    args.result = reinterpret_cast<int *>(arena.p);
    args.piv = reinterpret_cast<double *>(arena.p + o_piv);
    char *back = static_cast<char *>(ctx_pinned(ctx, o_back + 64));
    if (!back) return tci_fail(ctx, TCI_ERR_CUDA, "page-locked staging buffer");
    cudaEventRecord(ctx->ev0, ctx->stream);
    void *kargs[] = {&args};
    cudaEventRecord(ctx->ev2, ctx->stream);
    TCI_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(ZR_THREADS), kargs, smem, ctx->stream));
    cudaEventRecord(ctx->ev3, ctx->stream);
    ctx->launches++;
    TCI_CUDA(ctx, cudaMemcpyAsync(back, arena.p, o_back, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev1, ctx->stream);
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->stage_ms[ST_RRLU_KERNEL] += ms;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stage_ms[ST_RRLU] += ms;
    }
    const int *res = reinterpret_cast<const int *>(back);
    const double lu_error = *reinterpret_cast<const double *>(back + 16);
    const int r = res[0];
    if (res[1] & 1) return tci_fail(ctx, TCI_ERR_NAN_L, "lu.L contains NaNs");
    if (res[1] & 2) return tci_fail(ctx, TCI_ERR_NAN_U, "lu.U contains NaNs");
    const double *pv = reinterpret_cast<const double *>(back + o_piv);
    const i64 *rp = reinterpret_cast<const i64 *>(back + o_rp);
    const i64 *cp = reinterpret_cast<const i64 *>(back + o_cp);
    for (i64 i = 0; i < m; ++i) rowperm[i] = rp[i] + 1;
    for (i64 j = 0; j < n; ++j) colperm[j] = cp[j] + 1;
    *npivot = r;
    *error = lu_error;
    if (pivoterrors) {
        for (int q = 0; q < r; ++q) pivoterrors[q] = pv[q];
        pivoterrors[r] = lu_error;
    }
    if (factors) {
        tci_lu *lu = new tci_lu();
        lu->ctx = ctx;
        lu->A = A;
        lu->m = m;
        lu->n = n;
        lu->r = r;
        lu->leftorthogonal = leftorthogonal != 0;
        lu->is_complex = true;
        lu->arena = arena.p;
        lu->d_rowperm = args.rowperm;
        lu->d_colperm = args.colperm;
        lu->d_colpos = args.colpos;
        arena.p = nullptr;
        ctx->live_handles++;
        *factors = lu;
    }
    return TCI_OK;
}

extern "C" int tci_zrrlu(tci_ctx *ctx, const double *A_host, tci_dmat *A_dev, int64_t m, int64_t n, int64_t maxrank,
                         double reltol, double abstol, int leftorthogonal, int64_t *rowperm, int64_t *colperm,
                         int64_t *npivot, double *error, double *pivoterrors, tci_lu **factors)
{
    if (factors) *factors = nullptr;
    tci_dmat *A = A_dev;
    if (!ctx) return TCI_ERR_ARG;
    if ((A_host == nullptr) == (A_dev == nullptr) && m * n > 0)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_zrrlu: pass exactly one of A_host / A_dev");
    if (m < 0 || n < 0 || m > 0x3ffffff0 || n > 0x3ffffff0) return tci_fail(ctx, TCI_ERR_ARG, "tci_zrrlu: bad shape");
    if (!rowperm || !colperm || !npivot || !error) return tci_fail(ctx, TCI_ERR_ARG, "tci_zrrlu: output missing");
    if (A_dev && (A_dev->m != 2 * m || A_dev->n != n))
        return tci_fail(ctx, TCI_ERR_ARG, "tci_zrrlu: shape mismatch (a complex m x n matrix is a 2m x n device matrix)");
    if (A_dev && A_dev->ctx != ctx) return tci_fail(ctx, TCI_ERR_ARG, "tci_zrrlu: the matrix belongs to another context");
    if (!A_dev) {
        int rc = tci_dmat_create(ctx, 2 * m, n, A_host, &A);
        if (rc) return rc;
    }
    int rc;
    {
        TCI_ENTER(ctx);
        rc = zrrlu_core(ctx, A, m, n, maxrank, reltol, abstol, leftorthogonal, rowperm, colperm, npivot, error, pivoterrors,
                        factors);
    }
    if (!A_dev && !(factors && *factors)) tci_dmat_destroy(A);
    return rc;
}

extern "C" int tci_zlu_fetch(tci_lu *lu, double *L, double *U)
{
    if (!lu) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    if (!lu->is_complex) return tci_fail(ctx, TCI_ERR_ARG, "tci_zlu_fetch: not a ComplexF64 factorisation");
    const i64 m = lu->m, n = lu->n, r = lu->r;
    if (r == 0) return TCI_OK;
    DevBuf<double2> dL(ctx), dU(ctx);
    if (L) TCI_CUDA(ctx, dL.alloc((size_t)(m * r)));
    if (U) TCI_CUDA(ctx, dU.alloc((size_t)(r * n)));
    int rc = zlu_extract(lu, L ? dL.p : nullptr, m, U ? dU.p : nullptr, r);
    if (rc) return rc;
    StageTimer tm(ctx, ST_D2H);
    if (L) TCI_CUDA(ctx, cudaMemcpyAsync(L, dL.p, m * r * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    if (U) TCI_CUDA(ctx, cudaMemcpyAsync(U, dU.p, r * n * sizeof(double2), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

// res: complex (rows x cols) stored as a (2 rows) x cols device matrix
static int zfinish(tci_ctx *ctx, tci_dmat *res, double *out_host, tci_dmat **out_dev)
{
    if (out_host && res->m * res->n > 0) {
        StageTimer tm(ctx, ST_D2H);
        TCI_CUDA(ctx, cudaMemcpy2DAsync(out_host, res->m * sizeof(double), res->p, res->ld * sizeof(double),
                                        res->m * sizeof(double), res->n, cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (out_dev)
        *out_dev = res;
    else
        tci_dmat_destroy(res);
    return TCI_OK;
}
static inline unsigned nblk(i64 total) { return (unsigned)((total + 255) / 256); }

// left(luci) for complex factors: matrixluci.jl:40-42 / 48-57
extern "C" int tci_zluci_left(tci_lu *lu, double *out_host, tci_dmat **out_dev)
{
    if (!lu) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    if (out_dev) *out_dev = nullptr;
    if (!lu->is_complex) return tci_fail(ctx, TCI_ERR_ARG, "tci_zluci_left: not a ComplexF64 factorisation");
    const i64 m = lu->m, r = lu->r;
    tci_dmat *res = nullptr;
    int rc = dmat_alloc(ctx, 2 * m, r, &res);
    if (rc) return rc;
    if (r == 0) return zfinish(ctx, res, out_host, out_dev);
    {
        StageTimer tm(ctx, ST_LUCI);
        DevBuf<double2> L(ctx), U(ctx), Y(ctx);
        double2 *out = reinterpret_cast<double2 *>(res->p);
        TCI_CUDA(ctx, L.alloc((size_t)(m * r)));
        if (lu->leftorthogonal) { // [I; L21 L11^-1]
            rc = zlu_extract(lu, L.p, m, nullptr, 0);
            if (!rc && m > r) {
                k_ztrsm_right<false, true><<<nblk(m - r), 256, 0, ctx->stream>>>(L.p + r, m - r, m, L.p, m, (int)r);
                ctx->launches++;
            }
            if (!rc) {
                k_zscatter_rows<<<nblk(m * r), 256, 0, ctx->stream>>>(L.p, m, m, r, lu->d_rowperm, out, res->ld / 2, 1);
                ctx->launches++;
            }
        } else { // L * U11
            TCI_CUDA(ctx, U.alloc((size_t)(r * lu->n)));
            TCI_CUDA(ctx, Y.alloc((size_t)(m * r)));
            rc = zlu_extract(lu, L.p, m, U.p, r);
            if (!rc) rc = zgemm_dev(ctx, false, false, m, r, r, 1.0, L.p, m, U.p, r, 0.0, Y.p, m);
            if (!rc) {
                k_zscatter_rows<<<nblk(m * r), 256, 0, ctx->stream>>>(Y.p, m, m, r, lu->d_rowperm, out, res->ld / 2, 0);
                ctx->launches++;
            }
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, "zluci_left launch failed");
    }
    if (rc) {
        tci_dmat_destroy(res);
        return rc;
    }
    return zfinish(ctx, res, out_host, out_dev);
}

// right(luci) for complex factors: matrixluci.jl:44-46 / 59-68
extern "C" int tci_zluci_right(tci_lu *lu, double *out_host, tci_dmat **out_dev)
{
    if (!lu) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    if (out_dev) *out_dev = nullptr;
    if (!lu->is_complex) return tci_fail(ctx, TCI_ERR_ARG, "tci_zluci_right: not a ComplexF64 factorisation");
    const i64 m = lu->m, n = lu->n, r = lu->r;
    tci_dmat *res = nullptr;
    int rc = dmat_alloc(ctx, 2 * r, n, &res);
    if (rc) return rc;
    if (r == 0) return zfinish(ctx, res, out_host, out_dev);
    {
        StageTimer tm(ctx, ST_LUCI);
        DevBuf<double2> L(ctx), U(ctx), Y(ctx);
        double2 *out = reinterpret_cast<double2 *>(res->p);
        TCI_CUDA(ctx, U.alloc((size_t)(r * n)));
        if (lu->leftorthogonal) { // L11 * U
            TCI_CUDA(ctx, L.alloc((size_t)(m * r)));
            TCI_CUDA(ctx, Y.alloc((size_t)(r * n)));
            rc = zlu_extract(lu, L.p, m, U.p, r);
            if (!rc) rc = zgemm_dev(ctx, false, false, r, n, r, 1.0, L.p, m, U.p, r, 0.0, Y.p, r);
            if (!rc) {
                k_zscatter_cols<<<nblk(r * n), 256, 0, ctx->stream>>>(Y.p, r, r, n, lu->d_colperm, out, res->ld / 2, 0);
                ctx->launches++;
            }
        } else { // [I, U11^-1 U12]
            rc = zlu_extract(lu, nullptr, 0, U.p, r);
            if (!rc && n > r) {
                k_ztrsm_left_upper_unit<<<nblk(n - r), 256, 0, ctx->stream>>>(U.p + r * r, n - r, r, U.p, r, (int)r);
                ctx->launches++;
            }
            if (!rc) {
                k_zscatter_cols<<<nblk(r * n), 256, 0, ctx->stream>>>(U.p, r, r, n, lu->d_colperm, out, res->ld / 2, 1);
                ctx->launches++;
            }
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, "zluci_right launch failed");
    }
    if (rc) {
        tci_dmat_destroy(res);
        return rc;
    }
    return zfinish(ctx, res, out_host, out_dev);
}

// B * A^-1 (setsitetensor!, tensorci2.jl:391) with complex full-rank factors: X[:, rowperm] = B[:, colperm] U^-1 L^-1
extern "C" int tci_zlu_rdiv(tci_lu *lu, tci_dmat *B, double *out_host, tci_dmat **out_dev)
{
    if (!lu || !B) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    if (out_dev) *out_dev = nullptr;
    if (!lu->is_complex) return tci_fail(ctx, TCI_ERR_ARG, "tci_zlu_rdiv: not a ComplexF64 factorisation");
    const i64 k = lu->r, rows = B->m / 2;
    if (lu->m != lu->n || k != lu->m)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_zlu_rdiv: the factorised matrix must be square and of full rank");
    if (B->n != k || B->m % 2) return tci_fail(ctx, TCI_ERR_ARG, "tci_zlu_rdiv: DimensionMismatch between B and the factorised matrix");
    dmat_wait_ready(ctx, B);
    tci_dmat *res = nullptr;
    int rc = dmat_alloc(ctx, 2 * rows, k, &res);
    if (rc) return rc;
    if (rows == 0 || k == 0) return zfinish(ctx, res, out_host, out_dev);
    {
        StageTimer tm(ctx, ST_LUCI);
        DevBuf<double2> L(ctx), U(ctx), W(ctx);
        TCI_CUDA(ctx, L.alloc((size_t)(k * k)));
        TCI_CUDA(ctx, U.alloc((size_t)(k * k)));
        TCI_CUDA(ctx, W.alloc((size_t)(rows * k)));
        rc = zlu_extract(lu, L.p, k, U.p, k);
        if (!rc) {
            k_zgather_cols<<<nblk(rows * k), 256, 0, ctx->stream>>>(reinterpret_cast<const double2 *>(B->p), B->ld / 2, rows, k,
                                                                   lu->d_colperm, W.p, rows);
            if (lu->leftorthogonal) {
                k_ztrsm_right<true, false><<<nblk(rows), 256, 0, ctx->stream>>>(W.p, rows, rows, U.p, k, (int)k);
                k_ztrsm_right<false, true><<<nblk(rows), 256, 0, ctx->stream>>>(W.p, rows, rows, L.p, k, (int)k);
            } else {
                k_ztrsm_right<true, true><<<nblk(rows), 256, 0, ctx->stream>>>(W.p, rows, rows, U.p, k, (int)k);
                k_ztrsm_right<false, false><<<nblk(rows), 256, 0, ctx->stream>>>(W.p, rows, rows, L.p, k, (int)k);
            }
            k_zscatter_cols<<<nblk(rows * k), 256, 0, ctx->stream>>>(W.p, rows, rows, k, lu->d_rowperm,
                                                                    reinterpret_cast<double2 *>(res->p), res->ld / 2, 0);
            ctx->launches += 4;
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, "zlu_rdiv launch failed");
    }
    if (rc) {
        tci_dmat_destroy(res);
        return rc;
    }
    return zfinish(ctx, res, out_host, out_dev);
}

// ---- targets -------------------------------------------------------------------------------------------------
static int zmpo_create(tci_ctx *ctx, i64 nsites, const i64 *dimsA4, const double *const *A, const i64 *dimsB4,
                       const double *const *B, i64 *target_id)
{
    std::unique_ptr<TargetDev> t(new TargetDev());
    t->kind = 2;
    t->is_complex = true;
    t->nsites = nsites;
    const double one[2] = {1.0, 0.0};
    for (i64 s = 0; s < nsites; ++s) {
        const i64 unit4[4] = {1, 1, 1, 1};
        const i64 *da = dimsA4 + 4 * s, *db = dimsB4 ? dimsB4 + 4 * s : unit4;
        if (da[2] != db[1])
            return tci_fail(ctx, TCI_ERR_ARG, "Tensor trains must share the identical index at n=" + std::to_string(s + 1) + "!");
        t->adl.push_back(da[0]);
        t->as1.push_back(da[1]);
        t->as2.push_back(da[2]);
        t->adr.push_back(da[3]);
        t->bdl.push_back(db[0]);
        t->bs1.push_back(db[1]);
        t->bs2.push_back(db[2]);
        t->bdr.push_back(db[3]);
        t->localdims.push_back(da[1] * db[2]);
        const i64 na = 2 * da[0] * da[1] * da[2] * da[3], nb = 2 * db[0] * db[1] * db[2] * db[3];
        double *pa = nullptr, *pb = nullptr;
        cudaError_t e = cudaMalloc(&pa, std::max<i64>(na, 2) * sizeof(double));
        if (e == cudaSuccess) t->A.push_back(pa);
        if (e == cudaSuccess) e = cudaMalloc(&pb, std::max<i64>(nb, 2) * sizeof(double));
        if (e == cudaSuccess) t->B.push_back(pb);
        if (e == cudaSuccess) e = cudaMemcpy(pa, A[s], na * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(pb, B ? B[s] : one, nb * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { // nothing of a half-built target is kept
            target_free(ctx, *t);
            return tci_fail(ctx, TCI_ERR_CUDA, std::string("complex target upload: ") + cudaGetErrorString(e));
        }
    }
    const i64 id = ctx->next_target++;
    ctx->targets[id] = std::move(t);
    *target_id = id;
    return target_replicate(ctx, id);
}

extern "C" int tci_zmpo_pair_create(tci_ctx *ctx, int64_t nsites, const int64_t *dimsA4, const double *const *A,
                                    const int64_t *dimsB4, const double *const *B, int64_t *target_id)
{
    TCI_ENTER(ctx);
    if (!target_id || nsites < 1 || !dimsA4 || !dimsB4 || !A || !B)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_zmpo_pair_create: bad arguments");
    return zmpo_create(ctx, nsites, dimsA4, A, dimsB4, B, target_id);
}

extern "C" int tci_ztt_create(tci_ctx *ctx, int64_t nsites, const int64_t *dims3, const double *const *cores,
                              int64_t *target_id)
{
    TCI_ENTER(ctx);
    if (!target_id || nsites < 1 || !dims3 || !cores) return tci_fail(ctx, TCI_ERR_ARG, "tci_ztt_create: bad arguments");
    std::vector<i64> d4((size_t)(4 * nsites));
    for (i64 s = 0; s < nsites; ++s) {
        d4[4 * s] = dims3[3 * s];
        d4[4 * s + 1] = dims3[3 * s + 1];
        d4[4 * s + 2] = 1;
        d4[4 * s + 3] = dims3[3 * s + 2];
    }
    return zmpo_create(ctx, nsites, d4.data(), cores, nullptr, nullptr, target_id);
}

// f(z) = a*z + b on ComplexF64 (a, b real: Julia's Real * Complex and Complex + Real act component-wise / on re)
__global__ void k_zapply_affine(double2 *p, i64 m, i64 n, i64 ld, double fa, double fb)
{
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < m * n; e += (i64)gridDim.x * blockDim.x) {
        double2 *q = p + e % m + ld * (e / m);
        double2 v = *q;
        v.x = __dadd_rn(__dmul_rn(fa, v.x), fb);
        v.y = __dmul_rn(fa, v.y);
        *q = v;
    }
}
// maxabs(0, out) of util.jl:1-10 on complex data: abs = hypot; the ordered-bits max propagates NaN like Julia's max
__global__ void k_zmaxabs(const double2 *__restrict__ p, i64 m, i64 n, i64 ld, unsigned long long *gmax)
{
    unsigned long long mx = 0ull;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < m * n; e += (i64)gridDim.x * blockDim.x) {
        const double2 v = p[e % m + ld * (e / m)];
        const unsigned long long b = (unsigned long long)__double_as_longlong(tci_hypot(v.x, v.y)) & 0x7fffffffffffffffull;
        mx = b > mx ? b : mx;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = other > mx ? other : mx;
    }
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(gmax, mx);
}

static int zapply_f(tci_ctx *ctx, const TargetDev &t, double2 *p, i64 m, i64 n, i64 ld)
{
    if (t.fkind == TCI_F_NONE || m * n == 0) return TCI_OK;
    if (t.fkind != TCI_F_AFFINE)
        return tci_fail(ctx, TCI_ERR_UNSUPPORTED, "ComplexF64 contraction: only the affine elementwise function is available");
    k_zapply_affine<<<(unsigned)std::min<i64>((m * n + 255) / 256, (i64)ctx->sm_count * 8), 256, 0, ctx->stream>>>(p, m, n, ld, t.fa, t.fb);
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

static TargetDev *ztarget(tci_ctx *ctx, i64 id)
{
    auto it = ctx->targets.find(id);
    if (it == ctx->targets.end() || !it->second->is_complex) return nullptr;
    return it->second.get();
}

// batchevaluate(::Contraction{ComplexF64} / ::TTCache{ComplexF64}) -- tci_pi_eval for the complex value type
extern "C" int tci_zpi_eval(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                            int64_t nr, int64_t nJ, int64_t M, double *out_host, tci_dmat **out_dev, double *maxabs)
{
    TCI_ENTER(ctx);
    if (out_dev) *out_dev = nullptr;
    if (maxabs) *maxabs = 0.0;
    TargetDev *t = ztarget(ctx, target_id);
    if (!t) return tci_fail(ctx, TCI_ERR_ARG, "tci_zpi_eval: not a ComplexF64 target");
    if (nl < 0 || nr < 0 || M < 0 || nI < 0 || nJ < 0) return tci_fail(ctx, TCI_ERR_ARG, "tci_zpi_eval: negative size");
    if (nl + M + nr != t->nsites) return tci_fail(ctx, TCI_ERR_CENTRE, "Invalid number of central indices");
    if (nI * nJ == 0) return TCI_OK; // batcheval.jl:40-42
    i64 C = 1;
    for (i64 s = nl; s < nl + M; ++s) C *= t->localdims[s];
    const i64 rows = nI * C;
    tci_dmat *out = nullptr;
    int rc = dmat_alloc(ctx, 2 * rows, nJ, &out);
    if (rc) return rc;
    DevBuf<i64> idx(ctx);
    {
        StageTimer tm(ctx, ST_H2D);
        TCI_CUDA(ctx, idx.alloc(2 + (size_t)(nl * nI) + (size_t)(nr * nJ)));
        TCI_CUDA(ctx, cudaMemsetAsync(idx.p, 0, 16, ctx->stream));
        if (nl * nI > 0)
            TCI_CUDA(ctx, cudaMemcpyAsync(idx.p + 2, I, (size_t)(nl * nI) * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
        if (nr * nJ > 0)
            TCI_CUDA(ctx, cudaMemcpyAsync(idx.p + 2 + nl * nI, J, (size_t)(nr * nJ) * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
    }
    double mx = 0.0;
    {
        StageTimer tm(ctx, ST_PI);
        rc = pi_eval_mpo_z(ctx, *t, idx.p + 2, nl, nI, idx.p + 2 + nl * nI, nr, nJ, M, out, I, J);
        double2 *p = reinterpret_cast<double2 *>(out->p);
        if (!rc) rc = zapply_f(ctx, *t, p, rows, nJ, out->ld / 2);
        if (!rc && maxabs) {
            unsigned long long *w = reinterpret_cast<unsigned long long *>(idx.p);
            k_zmaxabs<<<(unsigned)std::min<i64>((rows * nJ + 255) / 256, (i64)ctx->sm_count * 8), 256, 0, ctx->stream>>>(p, rows, nJ, out->ld / 2, w);
            ctx->launches++;
            TCI_CUDA(ctx, cudaMemcpyAsync(&mx, w, 8, cudaMemcpyDeviceToHost, ctx->stream));
            TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    if (rc) {
        tci_dmat_destroy(out);
        return rc;
    }
    if (maxabs) *maxabs = mx;
    return zfinish(ctx, out, out_host, out_dev);
}

// evaluate(::Contraction{ComplexF64}, indexset) contraction.jl:189-207 for `count` points; out: 2*count doubles
extern "C" int tci_ztarget_eval(tci_ctx *ctx, int64_t target_id, const int64_t *idx, int64_t count, double *out)
{
    TCI_ENTER(ctx);
    TargetDev *t = ztarget(ctx, target_id);
    if (!t) return tci_fail(ctx, TCI_ERR_ARG, "tci_ztarget_eval: not a ComplexF64 target");
    if (count <= 0) return TCI_OK;
    DevBuf<i64> d_idx(ctx);
    DevBuf<double> d_out(ctx);
    {
        StageTimer tm(ctx, ST_H2D);
        TCI_CUDA(ctx, d_idx.upload(idx, (size_t)(t->nsites * count)));
        TCI_CUDA(ctx, d_out.alloc((size_t)(2 * count)));
    }
    int rc;
    {
        StageTimer tm(ctx, ST_PI);
        rc = target_eval_mpo_z(ctx, *t, d_idx.p, count, d_out.p);
        if (!rc) rc = zapply_f(ctx, *t, reinterpret_cast<double2 *>(d_out.p), count, 1, count);
    }
    if (rc) return rc;
    StageTimer tm(ctx, ST_D2H);
    TCI_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, (size_t)(2 * count) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

// The `:full` branch of updatepivots! (tensorci2.jl:529-551) on a ComplexF64 target: Pi stays in HBM between the
// evaluation and the factorisation; *factors (nullable) owns it afterwards.
extern "C" int tci_zbond_update(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                                int64_t nr, int64_t nJ, int64_t maxrank, double reltol, double abstol, int leftorthogonal,
                                int64_t *rowperm, int64_t *colperm, int64_t *npivot, double *error, double *pivoterrors,
                                double *maxabs, tci_lu **factors)
{
    if (factors) *factors = nullptr;
    if (!ctx) return TCI_ERR_ARG;
    if (nI * nJ == 0) return tci_fail(ctx, TCI_ERR_ARG, "tci_zbond_update: empty index set");
    tci_dmat *Pi = nullptr;
    int rc = tci_zpi_eval(ctx, target_id, I, nl, nI, J, nr, nJ, 0, nullptr, &Pi, maxabs);
    if (rc) return rc;
    rc = tci_zrrlu(ctx, nullptr, Pi, nI, nJ, maxrank, reltol, abstol, leftorthogonal, rowperm, colperm, npivot, error,
                   pivoterrors, factors);
    if (!(factors && *factors)) tci_dmat_destroy(Pi);
    return rc;
}
