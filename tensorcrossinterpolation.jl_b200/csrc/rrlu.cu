// rrlu.cu -- K2: full-pivot rank-revealing LU as ONE persistent cooperative kernel.
//
// Replaces the Julia loops of matrixlu.jl: submatrixargmax (:1-32), swaprow!/swapcol!
// (:98-112), addpivot! (:114-136) and _optimizerrlu! (:141-181).
//
// Work decomposition.  CTA g owns the physical columns j = g, g+G, g+2G, ... of the
// column-major matrix for the whole factorisation.  Rows are swapped physically (each
// CTA swaps inside its own columns), columns are permuted virtually (colpos[j] is the
// position the reference's physically swapped matrix would hold column j at), so no
// CTA ever writes another CTA's data.  One pivot step is:
//
//   1. every CTA reads the G posted candidates (|a|^2, value, row, column position)
//      and reduces them with the reference's tie-break: larger abs2, then smaller
//      column position, then smaller row  == "columns outer, rows inner, strict >"
//      of matrixlu.jl:16-29 on the physically swapped matrix;
//   2. stop rule (matrixlu.jl:153-158), evaluated redundantly and identically;
//   3. row swap s <-> pr in the own columns; colpos update; y_j = A[s, j];
//   4. x = the winner's *posted* pivot column (already divided by the pivot when
//      leftorthogonal) -- every CTA posts the column of its own candidate together
//      with the candidate, so that after ONE grid barrier all CTAs have the pivot
//      column without a second synchronisation;
//   5. trailing update a -= x_i * y_j (rounded multiply, rounded subtract in exact
//      mode: Julia does not contract) fused with the arg-max search for
//      the next pivot;
//   6. post candidate + its column for the next step; grid barrier.
//
// HBM traffic is the algorithmic 16 B per trailing element per pivot (one read, one
// write); for matrices up to ~100 MB the trailing matrix is L2 resident.
#include <cooperative_groups.h>

#include "tci_internal.h"

#define RR_MAX_THREADS 1024
#define RR_XS_CAP 24576 // doubles of shared memory for the pivot column
#define RR_U 4          // 16-byte loads in flight per lane in the trailing update

struct __align__(16) RRCand {
    double v;   // abs2 of the candidate
    double val; // its value
    int row;
    int colpos;
    int physcol;
    int valid;
};

struct RRArgs {
    double *A;
    i64 m, n, ld;
    int maxrank;
    double reltol, abstol;
    int leftorth;
    int *colpos;     // [n]
    i64 *rowperm;    // [m] 0-based
    i64 *colperm;    // [n] position -> physical column
    double *pivvals; // [maxrank]
    RRCand *cand;    // [2][G]
    double *xbuf;    // [2][G][ldx]
    i64 ldx;
    int *result;        // [0] npivot, [1] flags (1: no finite candidate left)
    double *result_err; // lu.error
    unsigned *barrier;  // [0] arrivals, [1] generation
    int xs_in_smem;
    int maxown;
};

__device__ __forceinline__ bool cand_better(double v, int cp, int row, double bv, int bcp, int brow)
{
    return v > bv || (v == bv && (cp < bcp || (cp == bcp && row < brow)));
}

__device__ __forceinline__ void grid_barrier(unsigned *bar, int G, unsigned &gen)
{
    __syncthreads();
    if (G > 1 && threadIdx.x == 0) {
        __threadfence();
        unsigned target = gen + 1;
        if (atomicAdd(&bar[0], 1u) == (unsigned)(G - 1)) {
            atomicExch(&bar[0], 0u);
            __threadfence();
            atomicExch(&bar[1], target);
        } else {
            while (*((volatile unsigned *)&bar[1]) != target) {
            }
        }
        __threadfence();
    }
    gen++;
    __syncthreads();
}

template <bool EXACT> __device__ __forceinline__ double schur(double a, double x, double y)
{
    if (EXACT) return __dsub_rn(a, __dmul_rn(x, y)); // matrixlu.jl:132
    return fma(-x, y, a);
}

template <bool EXACT> __global__ void __launch_bounds__(RR_MAX_THREADS, 1) k_rrlu(RRArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = gridDim.x, g = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const i64 m = a.m, n = a.n, ld = a.ld;
    double *const A = a.A;

    double *xs = reinterpret_cast<double *>(smem_raw);
    double *ys = xs + (a.xs_in_smem ? ((m + 1) & ~(i64)1) : 0);
    int *actc = reinterpret_cast<int *>(ys + a.maxown); // physical column of active entry
    int *actp = actc + a.maxown;                        // its position

    __shared__ double red_v[32];
    __shared__ int red_cp[32], red_row[32], red_col[32];
    __shared__ RRCand win;
    __shared__ int sh_nact;

    const int nown = (g < n) ? (int)((n - g + G - 1) / G) : 0;
    for (int e = tid; e < nown; e += T) {
        int j = g + e * G;
        actc[e] = j;
        actp[e] = j;
        a.colpos[j] = j;
    }
    if (tid == 0) sh_nact = nown;
    if (g == 0)
        for (i64 i = tid; i < m; i += T) a.rowperm[i] = i;
    unsigned gen = *((volatile unsigned *)&a.barrier[1]);
    __syncthreads();

    double maxerror = 0.0;
    double lasterr = nan("");
    int npiv = 0;
    int flags = 0;

    // s = -1 is the initial arg-max scan (no update); s >= 0 are pivot steps.
    for (int s = -1; s < a.maxrank; ++s) {
        int nact = sh_nact;
        const double *xw = nullptr; // winner's posted column (global)
        int pr = 0;
        bool do_update = false;
        if (s >= 0) {
            // ---- 1. reduce the posted candidates --------------------------------
            if (warp == 0) {
                const RRCand *cd = a.cand + (size_t)(s & 1) * G;
                double bv = -INFINITY, bval = 0.0;
                int bcp = 0x7fffffff, brow = 0x7fffffff, bcol = -1, bcta = -1;
                for (int q = lane; q < G; q += 32) {
                    const double2 d0 = __ldcg(reinterpret_cast<const double2 *>(cd + q));
                    const int4 d1 = __ldcg(reinterpret_cast<const int4 *>(cd + q) + 1);
                    if (d1.w && cand_better(d0.x, d1.y, d1.x, bv, bcp, brow)) {
                        bv = d0.x;
                        bval = d0.y;
                        brow = d1.x;
                        bcp = d1.y;
                        bcol = d1.z;
                        bcta = q;
                    }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    double oval = __shfl_xor_sync(0xffffffffu, bval, o);
                    int ocp = __shfl_xor_sync(0xffffffffu, bcp, o);
                    int orow = __shfl_xor_sync(0xffffffffu, brow, o);
                    int ocol = __shfl_xor_sync(0xffffffffu, bcol, o);
                    int octa = __shfl_xor_sync(0xffffffffu, bcta, o);
                    if (octa >= 0 && (bcta < 0 || cand_better(ov, ocp, orow, bv, bcp, brow))) {
                        bv = ov;
                        bval = oval;
                        bcp = ocp;
                        brow = orow;
                        bcol = ocol;
                        bcta = octa;
                    }
                }
                if (lane == 0) {
                    win.v = bv;
                    win.val = bval;
                    win.row = brow;
                    win.colpos = bcp;
                    win.physcol = bcol;
                    win.valid = bcta; // CTA index of the winner, -1 if none
                }
            }
            __syncthreads();
            const int wcta = win.valid;
            if (wcta < 0) { // nothing but NaNs left in the trailing block
                flags |= 1;
                break;
            }
            const double val = win.val;
            pr = win.row;
            const int jp = win.physcol, pcpos = win.colpos;
            // ---- 2. stop rule  matrixlu.jl:153-158 -----------------------------
            const double err = fabs(val);
            lasterr = err;
            if (s > 0 && (err < a.reltol * maxerror || err < a.abstol)) break;
            maxerror = (isnan(maxerror) || isnan(err)) ? nan("") : (err > maxerror ? err : maxerror);
            npiv = s + 1;
            if (g == 0 && tid == 0) {
                i64 t0 = a.rowperm[s];
                a.rowperm[s] = a.rowperm[pr];
                a.rowperm[pr] = t0;
                a.pivvals[s] = val;
            }
            do_update = (s + 1 < a.maxrank);
            xw = a.xbuf + ((size_t)(s & 1) * G + wcta) * a.ldx;

            // ---- 3. row swap in own columns, column bookkeeping, y ---------------
            if (pr != s)
                for (int e = tid; e < nown; e += T) {
                    double *col = A + (size_t)ld * (g + e * G);
                    double t0 = col[s];
                    col[s] = col[pr];
                    col[pr] = t0;
                }
            // position swap: the column sitting at position s moves to the pivot's old position
            int removed = -1;
            for (int e = tid; e < nact; e += T) {
                if (actc[e] == jp) {
                    removed = e;
                } else if (actp[e] == s) {
                    actp[e] = pcpos;
                    a.colpos[actc[e]] = pcpos;
                }
            }
            if (removed >= 0) { // exactly one thread of the owning CTA
                a.colpos[jp] = s;
                a.colperm[s] = jp;
                sh_nact = -(removed + 1); // signal, resolved after the barrier below
            }
            __syncthreads();
            if (sh_nact < 0) { // owner: drop the pivot column from the active list (swap with last)
                if (tid == 0) {
                    int e = -sh_nact - 1;
                    int last = nact - 1;
                    actc[e] = actc[last];
                    actp[e] = actp[last];
                    sh_nact = last;
                }
                __syncthreads();
            }
            const bool owner = (jp % G) == g;
            nact = sh_nact;
            if (do_update || !a.leftorth) {
                for (int e = tid; e < nact; e += T) {
                    double *p = A + (size_t)ld * actc[e] + s;
                    double y = *p;
                    if (!a.leftorth) { // matrixlu.jl:122
                        y = __ddiv_rn(y, val);
                        *p = y;
                    }
                    ys[e] = y;
                }
            }
            // ---- 4. pivot column ------------------------------------------------
            if (a.xs_in_smem && (do_update || (owner && a.leftorth))) {
                for (i64 i = s + 1 + tid; i < m; i += T) xs[i] = __ldcg(xw + (i == pr ? s : i));
            }
            __syncthreads();
            if (owner && a.leftorth) { // L column: A[k+1:end, k] ./= A[k, k]  matrixlu.jl:120
                double *col = A + (size_t)ld * jp;
                if (a.xs_in_smem)
                    for (i64 i = s + 1 + tid; i < m; i += T) col[i] = xs[i];
                else
                    for (i64 i = s + 1 + tid; i < m; i += T) col[i] = __ldcg(xw + (i == pr ? s : i));
            }
            if (!do_update) break; // the last Schur update never reaches L or U
        }

        // ---- 5. trailing update fused with the arg-max for the next pivot --------
        const int lo = s + 1; // first trailing row
        double bv = -INFINITY;
        int bcp = 0x7fffffff, brow = 0x7fffffff, bcol = -1;
        if (nact > 0) {
            // Tiles of RR_U*64 rows x 1 column are dealt round-robin to the warps; a lane issues its
            // RR_U 16-byte loads back to back so that enough bytes are in flight to cover HBM latency.
            const i64 i0 = lo & ~1;
            const int ntr = (int)((m - i0 + RR_U * 64 - 1) / (RR_U * 64));
            const i64 ntiles = (i64)nact * ntr;
            for (i64 t = warp; t < ntiles; t += nwarps) {
                const int e = (int)(t / ntr);
                const int rt = (int)(t - (i64)e * ntr);
                const int col = actc[e], cp = actp[e];
                const double y = do_update ? ys[e] : 0.0;
                double *const cptr = A + (size_t)ld * col;
                const i64 base = i0 + (i64)rt * (RR_U * 64) + 2 * lane;
                double2 d[RR_U];
#pragma unroll
                for (int u = 0; u < RR_U; ++u) {
                    const i64 i = base + u * 64;
                    if (i < m) d[u] = *reinterpret_cast<const double2 *>(cptr + i);
                }
#pragma unroll
                for (int u = 0; u < RR_U; ++u) {
                    const i64 i = base + u * 64;
                    if (i < m) {
                        const bool v0ok = i >= lo, v1ok = i + 1 < m;
                        if (do_update) {
                            double x0, x1;
                            if (a.xs_in_smem) {
                                const double2 xx = *reinterpret_cast<const double2 *>(xs + i);
                                x0 = xx.x;
                                x1 = xx.y;
                            } else {
                                x0 = v0ok ? __ldcg(xw + (i == pr ? s : i)) : 0.0;
                                x1 = v1ok ? __ldcg(xw + (i + 1 == pr ? s : i + 1)) : 0.0;
                            }
                            if (v0ok) d[u].x = schur<EXACT>(d[u].x, x0, y);
                            if (v1ok) d[u].y = schur<EXACT>(d[u].y, x1, y);
                            *reinterpret_cast<double2 *>(cptr + i) = d[u];
                        }
                        const double q0 = d[u].x * d[u].x, q1 = d[u].y * d[u].y;
                        if (v0ok && q0 >= bv && cand_better(q0, cp, (int)i, bv, bcp, brow)) {
                            bv = q0;
                            bcp = cp;
                            brow = (int)i;
                            bcol = col;
                        }
                        if (v1ok && q1 >= bv && cand_better(q1, cp, (int)i + 1, bv, bcp, brow)) {
                            bv = q1;
                            bcp = cp;
                            brow = (int)i + 1;
                            bcol = col;
                        }
                    }
                }
            }
        }
        // block reduction of the candidate
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int ocp = __shfl_xor_sync(0xffffffffu, bcp, o);
            int orow = __shfl_xor_sync(0xffffffffu, brow, o);
            int ocol = __shfl_xor_sync(0xffffffffu, bcol, o);
            if (ocol >= 0 && (bcol < 0 || cand_better(ov, ocp, orow, bv, bcp, brow))) {
                bv = ov;
                bcp = ocp;
                brow = orow;
                bcol = ocol;
            }
        }
        if (lane == 0) {
            red_v[warp] = bv;
            red_cp[warp] = bcp;
            red_row[warp] = brow;
            red_col[warp] = bcol;
        }
        __syncthreads();
        if (warp == 0) {
            bv = lane < nwarps ? red_v[lane] : -INFINITY;
            bcp = lane < nwarps ? red_cp[lane] : 0x7fffffff;
            brow = lane < nwarps ? red_row[lane] : 0x7fffffff;
            bcol = lane < nwarps ? red_col[lane] : -1;
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                int ocp = __shfl_xor_sync(0xffffffffu, bcp, o);
                int orow = __shfl_xor_sync(0xffffffffu, brow, o);
                int ocol = __shfl_xor_sync(0xffffffffu, bcol, o);
                if (ocol >= 0 && (bcol < 0 || cand_better(ov, ocp, orow, bv, bcp, brow))) {
                    bv = ov;
                    bcp = ocp;
                    brow = orow;
                    bcol = ocol;
                }
            }
            if (lane == 0) {
                red_v[0] = bv;
                red_cp[0] = bcp;
                red_row[0] = brow;
                red_col[0] = bcol;
            }
        }
        __syncthreads();
        // ---- 6. post the candidate and its column for step s+1 -------------------
        {
            const int nxt = (s + 1) & 1;
            bv = red_v[0];
            bcp = red_cp[0];
            brow = red_row[0];
            bcol = red_col[0];
            RRCand *slot = a.cand + (size_t)nxt * G + g;
            if (bcol >= 0) {
                const double *col = A + (size_t)ld * bcol;
                const double cval = col[brow];
                double *xo = a.xbuf + ((size_t)nxt * G + g) * a.ldx;
                if (a.leftorth)
                    for (i64 i = lo + tid; i < m; i += T) xo[i] = __ddiv_rn(col[i], cval);
                else
                    for (i64 i = lo + tid; i < m; i += T) xo[i] = col[i];
                if (tid == 0) {
                    slot->v = bv;
                    slot->val = cval;
                    slot->row = brow;
                    slot->colpos = bcp;
                    slot->physcol = bcol;
                    slot->valid = 1;
                }
            } else if (tid == 0) {
                slot->v = -INFINITY;
                slot->val = 0.0;
                slot->row = 0;
                slot->colpos = 0;
                slot->physcol = -1;
                slot->valid = 0;
            }
        }
        grid_barrier(a.barrier, G, gen);
    }

    // remaining (unpicked) columns keep the positions the reference's swaps gave them
    __syncthreads();
    {
        const int nact = sh_nact;
        for (int e = tid; e < nact; e += T) a.colperm[actp[e]] = actc[e];
    }
    if (g == 0 && tid == 0) {
        a.result[0] = npiv;
        a.result[1] = flags;
        const i64 mn = m < n ? m : n;
        *a.result_err = (npiv >= mn) ? 0.0 : lasterr; // matrixlu.jl:176-178
    }
}

// any NaN in L = tril(A[:,1:r]) or U = triu(A[1:r,:]) (matrixlu.jl:164-169); flags bit0 L, bit1 U
__global__ void k_nancheck(const double *__restrict__ A, i64 m, i64 n, i64 ld, const int *__restrict__ colpos, int r,
                           int *flags)
{
    int f = 0;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < m * n; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % m, j = e / m;
        int cp = colpos[j];
        bool inL = cp < r && i >= cp, inU = i < r && cp >= i;
        if ((inL || inU) && isnan(A[i + ld * j])) f |= (inL ? 1 : 0) | (inU ? 2 : 0);
    }
    if (f) atomicOr(flags, f);
}

// L (m x r, ldl) / U (r x n, ldu) in position order, as lu.L / lu.U of matrixlu.jl:162-174
__global__ void k_extract_L(const double *__restrict__ A, i64 m, i64 ld, const i64 *__restrict__ colperm, int r,
                            int leftorth, double *__restrict__ L, i64 ldl)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= m * r) return;
    i64 i = e % m, c = e / m;
    double v = 0.0;
    if (i > c)
        v = A[i + ld * colperm[c]];
    else if (i == c)
        v = leftorth ? 1.0 : A[i + ld * colperm[c]];
    L[i + ldl * c] = v;
}
__global__ void k_extract_U(const double *__restrict__ A, i64 n, i64 ld, const i64 *__restrict__ colperm, int r,
                            int leftorth, double *__restrict__ U, i64 ldu)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= (i64)r * n) return;
    i64 i = e % r, q = e / r;
    double v = 0.0;
    if (q > i)
        v = A[i + ld * colperm[q]];
    else if (q == i)
        v = leftorth ? A[i + ld * colperm[q]] : 1.0;
    U[i + ldu * q] = v;
}

static int rrlu_launch(tci_ctx *ctx, RRArgs &args, int G, int T, size_t smem, bool exact)
{
    void *kargs[] = {&args};
    const void *fn = exact ? (const void *)k_rrlu<true> : (const void *)k_rrlu<false>;
    TCI_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEventRecord(ctx->ev2, ctx->stream);
    TCI_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(T), kargs, smem, ctx->stream));
    cudaEventRecord(ctx->ev3, ctx->stream);
    ctx->launches++;
    return TCI_OK;
}

extern "C" int tci_rrlu(tci_ctx *ctx, const double *A_host, tci_dmat *A_dev, int64_t m, int64_t n, int64_t maxrank,
                        double reltol, double abstol, int leftorthogonal, int exact_mode, int64_t *rowperm,
                        int64_t *colperm, int64_t *npivot, double *error, double *pivoterrors, tci_lu **factors)
{
    if (factors) *factors = nullptr;
    tci_dmat *A = A_dev;
    if (!ctx) return TCI_ERR_ARG;
    if ((A_host == nullptr) == (A_dev == nullptr) && m * n > 0)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_rrlu: pass exactly one of A_host / A_dev");
    if (m < 0 || n < 0 || m > 0x7ffffff0 || n > 0x7ffffff0) return tci_fail(ctx, TCI_ERR_ARG, "tci_rrlu: bad shape");
    if (A_dev && (A_dev->m != m || A_dev->n != n)) return tci_fail(ctx, TCI_ERR_ARG, "tci_rrlu: shape mismatch");
    if (!A_dev) {
        int rc = tci_dmat_create(ctx, m, n, A_host, &A);
        if (rc) return rc;
    }
    TCI_ENTER(ctx);
    struct Cleanup {
        tci_dmat *a;
        bool armed;
        ~Cleanup()
        {
            if (armed) tci_dmat_destroy(a);
        }
    } cleanup{A, A_dev == nullptr};

    const i64 mn = std::min(m, n);
    i64 mr = (maxrank <= 0 || maxrank > mn) ? mn : maxrank;
    for (i64 i = 0; i < m; ++i) rowperm[i] = i + 1;
    for (i64 j = 0; j < n; ++j) colperm[j] = j + 1;
    *npivot = 0;
    if (mr == 0) { // nothing to do: lu.error = 0 because npivot >= min(m,n) = 0 (matrixlu.jl:176-178)
        *error = 0.0;
        if (pivoterrors) pivoterrors[0] = 0.0;
        return TCI_OK;
    }

    int G = (int)std::min<i64>(ctx->sm_count, std::max<i64>(1, n / 4));
    int T = m >= 1024 ? 1024 : (m >= 384 ? 512 : 256);
    const int maxown = (int)((n + G - 1) / G);
    const int xs_in_smem = m <= RR_XS_CAP;
    size_t smem = (xs_in_smem ? (size_t)((m + 1) & ~(i64)1) : 0) * sizeof(double) + (size_t)maxown * sizeof(double) +
                  2 * (size_t)maxown * sizeof(int) + 16;

    RRArgs args{};
    args.A = A->p;
    args.m = m;
    args.n = n;
    args.ld = A->ld;
    args.maxrank = (int)mr;
    args.reltol = reltol;
    args.abstol = abstol;
    args.leftorth = leftorthogonal ? 1 : 0;
    args.ldx = round_up(m, 16);
    args.xs_in_smem = xs_in_smem;
    args.maxown = maxown;
    args.barrier = ctx->rr_barrier;

    DevBuf<int> colpos(ctx), result(ctx);
    DevBuf<i64> d_rowperm(ctx), d_colperm(ctx);
    DevBuf<double> pivvals(ctx), xbuf(ctx), d_err(ctx);
    DevBuf<RRCand> cand(ctx);
    TCI_CUDA(ctx, colpos.alloc(n));
    TCI_CUDA(ctx, result.alloc(4));
    TCI_CUDA(ctx, d_rowperm.alloc(m));
    TCI_CUDA(ctx, d_colperm.alloc(n));
    TCI_CUDA(ctx, pivvals.alloc(mr));
    TCI_CUDA(ctx, xbuf.alloc((size_t)2 * G * args.ldx));
    TCI_CUDA(ctx, d_err.alloc(1));
    TCI_CUDA(ctx, cand.alloc((size_t)2 * G));
    TCI_CUDA(ctx, cudaMemsetAsync(result.p, 0, 4 * sizeof(int), ctx->stream));
    args.colpos = colpos.p;
    args.rowperm = d_rowperm.p;
    args.colperm = d_colperm.p;
    args.pivvals = pivvals.p;
    args.cand = cand.p;
    args.xbuf = xbuf.p;
    args.result = result.p;
    args.result_err = d_err.p;

    int res[4] = {0, 0, 0, 0};
    double lu_error = 0.0;
    std::vector<double> pv;
    {
        StageTimer tm(ctx, ST_RRLU);
        int rc = rrlu_launch(ctx, args, G, T, smem, exact_mode != 0);
        if (rc) return rc;
        TCI_CUDA(ctx, cudaMemcpyAsync(res, result.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaMemcpyAsync(&lu_error, d_err.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        {
            float kms = 0.f;
            if (cudaEventElapsedTime(&kms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->stage_ms[ST_RRLU_KERNEL] += kms;
        }
        const int r = res[0];
        unsigned blocks = (unsigned)std::min<i64>((m * n + 255) / 256, (i64)ctx->sm_count * 8);
        k_nancheck<<<blocks, 256, 0, ctx->stream>>>(A->p, m, n, A->ld, colpos.p, r, result.p + 2);
        ctx->launches++;
        pv.resize(r);
        TCI_CUDA(ctx, cudaMemcpyAsync(res + 2, result.p + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        if (r > 0)
            TCI_CUDA(ctx, cudaMemcpyAsync(pv.data(), pivvals.p, r * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaMemcpyAsync(rowperm, d_rowperm.p, m * sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaMemcpyAsync(colperm, d_colperm.p, n * sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const int r = res[0];
    if ((res[1] & 1) || (res[2] & 1)) return tci_fail(ctx, TCI_ERR_NAN_L, "lu.L contains NaNs");
    if (res[2] & 2) return tci_fail(ctx, TCI_ERR_NAN_U, "lu.U contains NaNs");
    for (i64 i = 0; i < m; ++i) rowperm[i] += 1;
    for (i64 j = 0; j < n; ++j) colperm[j] += 1;
    *npivot = r;
    *error = lu_error;
    if (pivoterrors) {
        for (int q = 0; q < r; ++q) pivoterrors[q] = std::fabs(pv[q]);
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
        lu->d_rowperm = d_rowperm.p;
        lu->d_colperm = d_colperm.p;
        lu->d_colpos = colpos.p;
        d_rowperm.p = nullptr; // ownership moves to the handle
        d_colperm.p = nullptr;
        colpos.p = nullptr;
        cleanup.armed = false;
        *factors = lu;
    }
    return TCI_OK;
}

int lu_extract(tci_lu *lu, double *dL, i64 ldl, double *dU, i64 ldu)
{
    tci_ctx *ctx = lu->ctx;
    const i64 m = lu->m, n = lu->n;
    const int r = (int)lu->r;
    if (r == 0) return TCI_OK;
    if (dL) {
        k_extract_L<<<(unsigned)((m * r + 255) / 256), 256, 0, ctx->stream>>>(lu->A->p, m, lu->A->ld, lu->d_colperm, r,
                                                                              lu->leftorthogonal, dL, ldl);
        ctx->launches++;
    }
    if (dU) {
        k_extract_U<<<(unsigned)(((i64)r * n + 255) / 256), 256, 0, ctx->stream>>>(lu->A->p, n, lu->A->ld, lu->d_colperm,
                                                                                  r, lu->leftorthogonal, dU, ldu);
        ctx->launches++;
    }
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

extern "C" int tci_lu_fetch(tci_lu *lu, double *L, double *U)
{
    if (!lu) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    const i64 m = lu->m, n = lu->n, r = lu->r;
    if (r == 0) return TCI_OK;
    DevBuf<double> dL(ctx), dU(ctx);
    if (L) TCI_CUDA(ctx, dL.alloc((size_t)(m * r)));
    if (U) TCI_CUDA(ctx, dU.alloc((size_t)(r * n)));
    int rc = lu_extract(lu, L ? dL.p : nullptr, m, U ? dU.p : nullptr, r);
    if (rc) return rc;
    StageTimer tm(ctx, ST_D2H);
    if (L) TCI_CUDA(ctx, cudaMemcpyAsync(L, dL.p, m * r * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (U) TCI_CUDA(ctx, cudaMemcpyAsync(U, dU.p, r * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

extern "C" int tci_lu_destroy(tci_lu *lu)
{
    if (!lu) return TCI_OK;
    tci_ctx *ctx = lu->ctx;
    cudaSetDevice(ctx->device);
    dev_free(ctx, lu->d_rowperm);
    dev_free(ctx, lu->d_colperm);
    dev_free(ctx, lu->d_colpos);
    tci_dmat_destroy(lu->A);
    delete lu;
    return TCI_OK;
}
