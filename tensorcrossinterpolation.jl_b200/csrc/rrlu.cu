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
//   1. warp 0 of every CTA polls the G posted candidate records (one aligned 16-byte word each:
//      value, row | phase bit, column position); the phase bit flips every second step, so the
//      polling loop IS the grid barrier and returns the candidates in the same round trip.  The
//      winner is reduced with the reference's tie-break: larger abs2, then smaller column position,
//      then smaller row == "columns outer, rows inner, strict >" of matrixlu.jl:16-29 on the
//      physically swapped matrix;
//   2. stop rule (matrixlu.jl:153-158), evaluated redundantly and identically;
//   3. row swap s <-> pr in the own columns; position update; y_j = A[s, j];
//   4. x = the winner's *posted* pivot column (already divided by the pivot when
//      leftorthogonal) -- every CTA posts the column of its own candidate before its record,
//      so after ONE synchronisation all CTAs have the pivot column;
//   5. trailing update a -= x_i * y_j (rounded multiply, rounded subtract in exact
//      mode: Julia does not contract) fused with the arg-max search for
//      the next pivot;
//   6. post the column, fence, then the record.
//
// HBM traffic is the algorithmic 16 B per trailing element per pivot (one read, one
// write); matrices up to ~30 MB live in shared memory for the whole factorisation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <set>

#include "tci_internal.h"

#include "rrlu_common.cuh"

#define RR_MARK(ph)                                                      \
    do {                                                                 \
        if (a.dbg && g == a.dbg_cta && tid == 0) {                       \
            long long now__ = clock64();                                 \
            dbg_acc[ph] += now__ - tmark;                                \
            tmark = now__;                                               \
        }                                                                \
    } while (0)

// MODE 0: single CTA, matrix resident in its shared memory, no global synchronisation at all
// MODE 1: columns resident in the shared memory of G CTAs (matrices up to ~30 MB over 148 SMs)
// MODE 2: columns streamed from L2 / HBM, pivot column staged in shared memory
// MODE 3: as 2 for m > RR_XS_CAP, pivot column read from L2
// The variants are separate instantiations so that the per-pivot loop stays inside the 32 KB
// instruction cache (a single generic kernel was 61 KB of SASS and refetched itself every step).
template <bool EXACT, int MODE, bool LEFT> __device__ __forceinline__ void rrlu_body(RRArgs &a, const int Gq, const int gq)
{
    constexpr bool RES = MODE <= 1, SINGLE = MODE == 0, XS = MODE != 3;
    constexpr int U = RES ? 2 : RR_U; // 16-byte accesses in flight per lane (deep only when streaming)
    long long tmark = clock64();
    __shared__ long long dbg_acc[16];
    if (threadIdx.x < 16) dbg_acc[threadIdx.x] = 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = SINGLE ? 1 : Gq, g = SINGLE ? 0 : gq, T = blockDim.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int m = (int)a.m, n = (int)a.n;
    const i64 cld = RES ? a.lds : a.ld; // column stride of the working copy

    double *xs = reinterpret_cast<double *>(smem_raw);
    double *ys = xs + (XS ? ((m + 1) & ~1) : 0);
    double *cols = ys + ((a.maxown + 1) & ~1);                                          // RES: maxown * lds doubles
    int *acto = reinterpret_cast<int *>(cols + (RES ? (size_t)a.maxown * a.lds : 0)); // own slot of active entry
    int *actp = acto + a.maxown;                                                       // its position
    // own slot o <-> physical column g + o*G ; working copy of that column:
    double *const W = RES ? cols : a.A + (size_t)a.ld * g;
    const i64 wstride = RES ? cld : cld * G;
#define RR_COL(o) (W + (size_t)wstride * (o))

    __shared__ unsigned long long red_v[32], red_key[32];
    __shared__ int red_slot;
    __shared__ double win_val;
    __shared__ int win_row, win_colpos, win_cta;
    __shared__ int sh_nact, sh_removed;

    const int nown = (g < n) ? (n - g + G - 1) / G : 0;
    _Pragma("unroll 1") for (int e = tid; e < nown; e += T) {
        const int j = g + e * G;
        acto[e] = e;
        actp[e] = j;
        a.colpos[j] = j;
    }
    if (tid == 0) {
        sh_nact = nown;
        sh_removed = -1;
    }
    if (g == 0)
        _Pragma("unroll 1") for (int i = tid; i < m; i += T) a.rowperm[i] = i;
    if (RES) { // stage the own columns
#pragma unroll 1
        for (int o = warp; o < nown; o += nwarps) {
            const double *src = a.A + (size_t)a.ld * (g + o * G);
            double *dst = cols + (size_t)a.lds * o;
            for (int i = 2 * lane; i < m; i += 64)
                *reinterpret_cast<double2 *>(dst + i) = *reinterpret_cast<const double2 *>(src + i);
        }
    }
    __syncthreads();

    double maxerror = 0.0;
    double lasterr = nan("");
    int npiv = 0;
    int flags = 0;

    // s = -1 is the initial arg-max scan (no update); s >= 0 are pivot steps.
#pragma unroll 1
    for (int s = -1; s < a.maxrank; ++s) {
        int nact = sh_nact;
        const double *xw = nullptr; // winner's posted column (global)
        int pr = 0;
        bool do_update = false;
        if (s >= 0) {
            // ---- 1. wait for / reduce the posted candidates ----------------------
            if (!SINGLE) {
                if (warp == 0) {
                    const RRCand *cd = a.cand + (size_t)(s & 1) * G;
                    const unsigned phase = (((unsigned)s >> 1) & 1u) ^ 1u;
                    double cv[RR_MAXQ];
                    unsigned crp[RR_MAXQ];
                    int ccp[RR_MAXQ];
                    for (;;) { // all records of this lane in flight together; polling them is the grid barrier
                        bool ok = true;
#pragma unroll
                        for (int k = 0; k < RR_MAXQ; ++k) {
                            const int q = lane + 32 * k;
                            if (q < G) ld_relaxed_16(cd + q, cv[k], crp[k], ccp[k]);
                        }
#pragma unroll
                        for (int k = 0; k < RR_MAXQ; ++k) {
                            const int q = lane + 32 * k;
                            if (q < G) ok = ok && ((crp[k] >> 31) == phase);
                        }
                        if (__all_sync(0xffffffffu, ok)) break;
                    }
                    RR_MARK(7); // poll until all records are fresh
#ifdef RR_ACQUIRE_FENCE
                    __threadfence(); // formal acquire; the posted column is read with ld.cg at an address that
#endif                               // depends on these records, after the writer's release fence
                    unsigned long long vb = 0ull, key = ~0ull;
#pragma unroll
                    for (int k = 0; k < RR_MAXQ; ++k) {
                        const int q = lane + 32 * k;
                        if (q < G && ccp[k] >= 0) {
                            const unsigned long long v = vbits(cv[k] * cv[k]);
                            const unsigned long long kk =
                                ((unsigned long long)(unsigned)ccp[k] << 32) | (crp[k] & 0x7fffffffu);
                            if (v > vb || (v == vb && kk < key)) {
                                vb = v;
                                key = kk;
                            }
                        }
                    }
                    warp_argmax(vb, key);
                    if (vb == 0ull) {
                        if (lane == 0) win_cta = -1; // no CTA has a finite candidate
                    } else {
#pragma unroll
                        for (int k = 0; k < RR_MAXQ; ++k) {
                            const int q = lane + 32 * k;
                            if (q < G && ccp[k] == (int)(key >> 32) && (crp[k] & 0x7fffffffu) == (unsigned)key) {
                                win_val = cv[k];
                                win_row = (int)(unsigned)key;
                                win_colpos = ccp[k];
                                win_cta = q;
                            }
                        }
                    }
                    RR_MARK(9);
                }
                __syncthreads();
            }
            RR_MARK(0);
            const int wcta = win_cta;
            if (wcta < 0) { // nothing but NaNs left in the trailing block
                flags |= 1;
                break;
            }
            const double val = win_val;
            pr = win_row;
            const int pcpos = win_colpos;
            // ---- 2. stop rule  matrixlu.jl:153-158 -----------------------------
            const double err = fabs(val);
            lasterr = err;
            if (s > 0 && (err < a.reltol * maxerror || err < a.abstol)) break;
            maxerror = (isnan(maxerror) || isnan(err)) ? nan("") : (err > maxerror ? err : maxerror);
            npiv = s + 1;
            if (g == 0 && tid == 0) { // row swaps are replayed into rowperm after the loop
                a.pivrows[s] = pr;
                a.pivvals[s] = val;
            }
            do_update = (s + 1 < a.maxrank);
            xw = a.xbuf + ((size_t)(s & 1) * G + wcta) * a.ldx;

            // ---- 3. row swap in own columns, column bookkeeping, pivot column ------
            if (pr != s)
                _Pragma("unroll 1") for (int o = tid; o < nown; o += T) {
                    double *col = RR_COL(o);
                    const double t0 = col[s];
                    col[s] = col[pr];
                    col[pr] = t0;
                }
            // the pivot column sits at position pcpos; the column at position s takes that position
            _Pragma("unroll 1") for (int e = tid; e < nact; e += T) {
                if (actp[e] == pcpos) {
                    sh_removed = e; // exactly one thread of the owning CTA
                } else if (actp[e] == s) {
                    actp[e] = pcpos;
                    a.colpos[g + acto[e] * G] = pcpos;
                }
            }
            // pivot column x (posted by the winner; SINGLE: already in xs from the previous step)
            if (!SINGLE && XS) {
                if (do_update || LEFT)
                    _Pragma("unroll 1") for (int i = s + 1 + tid; i < m; i += T) xs[i] = __ldcg(xw + (i == pr ? s : i));
            } else if (SINGLE && pr > s) {
                // xs was formed before the row swap: entry pr takes the value of entry s (never read again)
                if (tid == 0) xs[pr] = xs[s];
            }
            __syncthreads();
            RR_MARK(1);
            const int rem = sh_removed;
            int oslot = -1; // own slot of the pivot column if this CTA owns it
            if (rem >= 0) {
                oslot = acto[rem];
                __syncthreads();
                if (tid == 0) { // drop the pivot column from the active list (swap with last)
                    const int jp = g + oslot * G;
                    a.colpos[jp] = s;
                    a.colperm[s] = jp;
                    const int last = nact - 1;
                    acto[rem] = acto[last];
                    actp[rem] = actp[last];
                    sh_nact = last;
                    sh_removed = -1;
                }
                __syncthreads();
            }
            nact = sh_nact;
            if (do_update || !LEFT) {
                _Pragma("unroll 1") for (int e = tid; e < nact; e += T) {
                    double *p = RR_COL(acto[e]) + s;
                    double y = *p;
                    if (!LEFT) { // matrixlu.jl:122
                        y = __ddiv_rn(y, val);
                        *p = y;
                    }
                    ys[e] = y;
                }
            }
            if (LEFT && oslot >= 0) { // L column: A[k+1:end, k] ./= A[k, k]  matrixlu.jl:120
                double *col = RR_COL(oslot);
                if (XS)
                    _Pragma("unroll 1") for (int i = s + 1 + tid; i < m; i += T) col[i] = xs[i];
                else
                    _Pragma("unroll 1") for (int i = s + 1 + tid; i < m; i += T) col[i] = __ldcg(xw + (i == pr ? s : i));
            }
            __syncthreads();
            RR_MARK(2);
            if (!do_update) break; // the last Schur update never reaches L or U
        }

        // ---- 5. trailing update fused with the arg-max for the next pivot --------
        const int lo = s + 1; // first trailing row
        // This thread's best candidate: ordered value bits (0 = none), column position, row.  Order: larger
        // abs2, then smaller column position, then smaller row.  All compares are on integers (the squares
        // are non-negative, so their bit patterns are ordered) and the eight elements of a tile are reduced
        // as a tree, which keeps the dependent chain short.
        double bq = -1.0; // this thread's best square (-1: none)
        int bcpv = 0x7fffffff, browv = 0x7fffffff;
        if (nact > 0) {
            // Tiles of RR_U*64 rows x 1 column are dealt round-robin to the warps; a lane issues its
            // RR_U 16-byte loads back to back so that enough bytes are in flight to cover HBM latency.
            const int i0 = lo & ~1;
            const int ntr = (m - i0 + U * 64 - 1) / (U * 64);
            // tile t = e*ntr + rt, t = warp, warp + nwarps, ...  Odd steps walk the same tiles backwards: what
            // was touched last in the previous step is touched first now and is still in L2 (the trailing
            // matrix is swept once per pivot, so a fixed order would evict everything before it is reused).
            const int de = nwarps / ntr, dr = nwarps - de * ntr;
            const int ntiles = nact * ntr;
            const bool backward = !RES && (s & 1);
            int t0 = warp;
            if (backward && ntiles > warp) t0 = warp + ((ntiles - 1 - warp) / nwarps) * nwarps;
            int e = t0 / ntr, rt = t0 - e * ntr;
#pragma unroll 1
            while (e < nact && e >= 0) {
                const int cp = actp[e];
                const double y = do_update ? ys[e] : 0.0;
                double *const cptr = RR_COL(acto[e]);
                const int base = i0 + rt * (U * 64) + 2 * lane;
                double2 d[U];
                double tq = -1.0; // best square of this tile (-1: none)
                int tj = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * 64;
                    if (i < m) d[u] = *reinterpret_cast<const double2 *>(cptr + i);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * 64;
                    const bool v0ok = i >= lo && i < m, v1ok = i + 1 < m;
                    if (do_update && i < m) {
                        double x0, x1;
                        if (XS) {
                            const double2 xx = *reinterpret_cast<const double2 *>(xs + i);
                            x0 = xx.x;
                            x1 = xx.y;
                        } else {
                            x0 = v0ok ? __ldcg(xw + (i == pr ? s : i)) : 0.0;
                            x1 = v1ok ? __ldcg(xw + (i + 1 == pr ? s : i + 1)) : 0.0;
                        }
                        if (v0ok) d[u].x = schur<EXACT>(d[u].x, x0, y);
                        if (v1ok) d[u].y = schur<EXACT>(d[u].y, x1, y);
                        *reinterpret_cast<double2 *>(cptr + i) = d[u]; // masked entries are written back unchanged
                    }
                    // arg-max over the tile on the squares themselves: a strict > in row order keeps the first
                    // maximum and never selects a NaN (matrixlu.jl:16-29); masked entries do not take part
                    const double q0 = d[u].x * d[u].x, q1 = d[u].y * d[u].y;
                    if (v0ok && q0 > tq) {
                        tq = q0;
                        tj = 2 * u;
                    }
                    if (v1ok && q1 > tq) {
                        tq = q1;
                        tj = 2 * u + 1;
                    }
                }
                const int trow = base + (tj >> 1) * 64 + (tj & 1);
                if (tq > bq || (tq == bq && tq >= 0.0 && (cp < bcpv || (cp == bcpv && trow < browv)))) {
                    bq = tq;
                    bcpv = cp;
                    browv = trow;
                }
                if (backward) {
                    e -= de;
                    rt -= dr;
                    if (rt < 0) {
                        rt += ntr;
                        --e;
                    }
                } else {
                    e += de;
                    rt += dr;
                    if (rt >= ntr) {
                        rt -= ntr;
                        ++e;
                    }
                }
            }
        }
        unsigned long long bkey = ((unsigned long long)(unsigned)bcpv << 32) | (unsigned)browv;
        RR_MARK(3);
        // block reduction of the candidate (ordered value bits: the squares are non-negative; key)
        unsigned long long vb = bq < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(bq) + 1ull;
        warp_argmax(vb, bkey);
        if (lane == 0) {
            red_v[warp] = vb;
            red_key[warp] = bkey;
        }
        if (tid == 0) red_slot = -1;
        __syncthreads();
        vb = lane < nwarps ? red_v[lane] : 0ull; // every warp reduces the per-warp results itself
        bkey = lane < nwarps ? red_key[lane] : ~0ull;
        warp_argmax(vb, bkey);
        const int bcp = (int)(bkey >> 32), brow = (int)(bkey & 0xffffffffull);
        if (vb != 0ull) // own slot of the column that sits at position bcp
            _Pragma("unroll 1") for (int e = tid; e < nact; e += T)
                if (actp[e] == bcp) red_slot = acto[e];
        __syncthreads();
        const int bslot = red_slot;
        RR_MARK(4);
        // ---- 6. post the candidate and its column for step s+1 -------------------
        if (SINGLE) { // the winner is the own candidate, its column goes straight to xs
            if (bslot >= 0) {
                const double *col = RR_COL(bslot);
                const double cval = col[brow];
                _Pragma("unroll 1") for (int i = lo + tid; i < m; i += T) xs[i] = LEFT ? __ddiv_rn(col[i], cval) : col[i];
                if (tid == 0) {
                    win_val = cval;
                    win_row = brow;
                    win_colpos = bcp;
                    win_cta = 0;
                }
            } else if (tid == 0)
                win_cta = -1;
            __syncthreads();
        } else {
            const int nxt = (s + 1) & 1;
            RRCand *slot = a.cand + (size_t)nxt * G + g;
            double cval = 0.0;
            if (bslot >= 0) {
                const double *col = RR_COL(bslot);
                cval = col[brow];
                double *xo = a.xbuf + ((size_t)nxt * G + g) * a.ldx;
                _Pragma("unroll 1") for (int i = lo + tid; i < m; i += T) xo[i] = LEFT ? __ddiv_rn(col[i], cval) : col[i];
            }
            __syncthreads(); // all column stores of the CTA are ordered before the release below
            RR_MARK(5);
            if (tid == 0) {
                const unsigned phase = ((((unsigned)(s + 1)) >> 1) & 1u) ^ 1u;
                __threadfence(); // fence + relaxed store = release
                st_relaxed_16(slot, cval, (unsigned)(bslot >= 0 ? brow : 0) | (phase << 31), bslot >= 0 ? bcp : -1);
            }
            RR_MARK(6);
        }
    }

    __syncthreads();
    { // NaN scan of the L and U parts of the own columns (matrixlu.jl:164-169): bit 0 L, bit 1 U
        int f = 0;
#pragma unroll 1
        for (int o = warp; o < nown; o += nwarps) {
            const int cp = a.colpos[g + o * G];
            const double *col = RR_COL(o);
            const int hi = cp < npiv ? m : npiv; // picked column: all rows belong to L or U; otherwise rows < r (U)
            for (int i = lane; i < hi; i += 32)
                if (isnan(col[i])) f |= ((cp < npiv && i >= cp) ? 1 : 0) | ((i < npiv && cp >= i) ? 2 : 0);
        }
        if (f) atomicOr(&a.result[2], f);
    }
    if (RES) { // write the factors back
#pragma unroll 1
        for (int o = warp; o < nown; o += nwarps) {
            double *dst = a.A + (size_t)a.ld * (g + o * G);
            const double *src = cols + (size_t)a.lds * o;
            for (int i = 2 * lane; i < m; i += 64)
                *reinterpret_cast<double2 *>(dst + i) = *reinterpret_cast<const double2 *>(src + i);
        }
    }
    if (a.dbg && g == a.dbg_cta && tid < 16) a.dbg[tid] = dbg_acc[tid];
    // remaining (unpicked) columns keep the positions the reference's swaps gave them
    {
        const int nact = sh_nact;
        _Pragma("unroll 1") for (int e = tid; e < nact; e += T) a.colperm[actp[e]] = g + acto[e] * G;
    }
    if (g == 0 && tid == 0) {
        for (int q = 0; q < npiv; ++q) { // swaprow! bookkeeping, matrixlu.jl:99-100
            const int pq = a.pivrows[q];
            const i64 t0 = a.rowperm[q];
            a.rowperm[q] = a.rowperm[pq];
            a.rowperm[pq] = t0;
        }
        a.result[0] = npiv;
        a.result[1] = flags;
        const int mn = m < n ? m : n;
        *a.result_err = (npiv >= mn) ? 0.0 : lasterr; // matrixlu.jl:176-178
    }
#undef RR_COL
}

template <bool EXACT, int MODE, bool LEFT> __global__ void __launch_bounds__(RR_MAX_THREADS, 1) k_rrlu(RRArgs a)
{
    rrlu_body<EXACT, MODE, LEFT>(a, (int)gridDim.x, (int)blockIdx.x);
}

// Several independent small factorisations in ONE cooperative launch: matrix q is factorised by the Gq consecutive
// CTAs [q*Gq, (q+1)*Gq) with its columns resident in their shared memory (MODE 1); the record polling that
// synchronises a group never looks outside it.  On-chip factorisations are bound by the per-pivot latency chain, not by
// throughput, so the ~20 pivot matrices of fillsitetensors! cost the time of the largest one instead of their sum.
template <bool EXACT, bool LEFT> __global__ void __launch_bounds__(RR_MAX_THREADS, 1) k_rrlu_batch(const RRArgs *batch, int Gq)
{
    RRArgs a = batch[blockIdx.x / Gq];
    rrlu_body<EXACT, 1, LEFT>(a, Gq, (int)(blockIdx.x % Gq));
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

template <bool EXACT, int MODE> static const void *rrlu_fn(bool left)
{
    return left ? (const void *)k_rrlu<EXACT, MODE, true> : (const void *)k_rrlu<EXACT, MODE, false>;
}
template <bool EXACT> static const void *rrlu_fn_mode(int mode, bool left)
{
    switch (mode) {
    case 0: return rrlu_fn<EXACT, 0>(left);
    case 1: return rrlu_fn<EXACT, 1>(left);
    case 2: return rrlu_fn<EXACT, 2>(left);
    default: return rrlu_fn<EXACT, 3>(left);
    }
}

static int rrlu_launch(tci_ctx *ctx, RRArgs &args, int G, int T, size_t smem, bool exact, int mode)
{
    void *kargs[] = {&args};
    const void *fn = exact ? rrlu_fn_mode<true>(mode, args.leftorth != 0) : rrlu_fn_mode<false>(mode, args.leftorth != 0);
    TCI_CUDA(ctx, ctx_func_smem(ctx, fn, 227 * 1024 - 1024));
    cudaEventRecord(ctx->ev2, ctx->stream);
    TCI_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(T), kargs, smem, ctx->stream));
    cudaEventRecord(ctx->ev3, ctx->stream);
    ctx->launches++;
    return TCI_OK;
}

// The factorisation proper: A (on the context's device) is factorised in place; nothing but the final result copy
// synchronises.  extra_dev / extra_host / extra_bytes: a small device-to-host copy that rides on the same
// synchronisation (max|Pi| of the evaluation queued in front, tci_bond_update).
int rrlu_core(tci_ctx *ctx, tci_dmat *A, i64 m, i64 n, i64 maxrank, double reltol, double abstol, int leftorthogonal,
              int exact_mode, i64 *rowperm, i64 *colperm, i64 *npivot, double *error, double *pivoterrors,
              tci_lu **factors, const void *extra_dev, void *extra_host, size_t extra_bytes, int *deferred_result)
{
    // deferred_result != nullptr (batched callers, tci_fill_sitetensors): nothing synchronises; the 8 result words
    // (npivot, flags, flags, -, lu.error) and the maxrank pivot values are copied to deferred_result (page-locked,
    // 32 + 8 * maxrank bytes) in stream order and checked by the caller after ITS synchronisation; the handle is
    // returned under the assumption npivot == maxrank.
    dmat_wait_ready(ctx, A);
    const i64 mn = std::min(m, n);
    i64 mr = (maxrank <= 0 || maxrank > mn) ? mn : maxrank;
    if (rowperm)
        for (i64 i = 0; i < m; ++i) rowperm[i] = i + 1;
    if (colperm)
        for (i64 j = 0; j < n; ++j) colperm[j] = j + 1;
    *npivot = 0;
    if (mr == 0) { // nothing to do: lu.error = 0 because npivot >= min(m,n) = 0 (matrixlu.jl:176-178)
        *error = 0.0;
        if (pivoterrors) pivoterrors[0] = 0.0;
        if (extra_bytes) {
            TCI_CUDA(ctx, cudaMemcpyAsync(extra_host, extra_dev, extra_bytes, cudaMemcpyDeviceToHost, ctx->stream));
            TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        return TCI_OK;
    }

    // Grid / residency policy.  A matrix that fits one CTA's shared memory is factorised by a single
    // CTA without any global synchronisation; up to ~30 MB the columns are spread over the SMs'
    // shared memory (RES); beyond that they are streamed from L2 / HBM.
    const size_t SMEM_BUDGET = 220 * 1024;
    const i64 lds = (m + 1) & ~(i64)1;
    const int xs_in_smem = m <= RR_XS_CAP;
    auto smem_need = [&](int Gq, bool resq) {
        const i64 mo = (n + Gq - 1) / Gq;
        return (size_t)(xs_in_smem ? lds : 0) * 8 + (size_t)((mo + 1) & ~(i64)1) * 8 + (size_t)mo * 8 +
               (resq ? (size_t)mo * lds * 8 : 0) + 64;
    };
    int G;
    bool resident = false;
    if (smem_need(1, true) <= SMEM_BUDGET) {
        G = 1;
        resident = true;
    } else {
        // (measured: the full grid is best from 1024^2 up; a persisting-L2 access-policy window over the
        //  matrix was tried for the streaming regime and halved the bandwidth, so it is not used)
        G = (int)std::min<i64>(std::min<i64>(ctx->sm_count, 32 * RR_MAXQ), std::max<i64>(1, n / 2));
        resident = xs_in_smem && smem_need(G, true) <= SMEM_BUDGET;
    }
    if (getenv("TCI_RRLU_NO_RES")) resident = false;
    if (const char *ge = getenv("TCI_RRLU_G")) // experiments only
        if (!resident && atoi(ge) >= 1 && atoi(ge) <= G) G = atoi(ge);
    // streaming regime: deferred updates (rrlu_lazy.cu) unless disabled
    const bool no_lazy = getenv("TCI_RRLU_NO_LAZY") != nullptr;
    // (its tiles are moved by cp.async.bulk: columns must start on 16-byte boundaries)
    // (measured, ms deferred / in place: 2048^2 4.7 / 3.6, 3072^2 6.8 / 5.9, 3584^2 7.9 / 8.0, 4096^2 9.5 / 10.6,
    //  5120^2 15.5 / 17.2, 6144^2 18.8 / 25.1; below ~3600^2 the matrix stays in the 126 MB L2 and the in-place
    //  kernel's smaller fixed cost per pivot wins)
    const char *lazy_min_env = getenv("TCI_RRLU_LAZY_MIN");
    const double lazy_min = lazy_min_env ? atof(lazy_min_env) : 13e6;
    const bool lazy = !resident && !no_lazy && (double)m * (double)n >= lazy_min &&
                      rrlu_lazy_smem((int)((n + G - 1) / G) + 8, RRLU_LAZY_NB) <= SMEM_BUDGET && (A->ld % 2 == 0) && (reinterpret_cast<size_t>(A->p) % 16 == 0);
    const i64 per_cta = m * ((n + G - 1) / G);
    int T = (m >= 1024 || per_cta >= 32768) ? 1024 : ((m >= 384 || per_cta >= 8192) ? 512 : 256);
    int maxown = (int)((n + G - 1) / G);
    // Deferred-update kernel on the full grid: the SMs do not stream from HBM equally fast (measured spread
    // of the pass time per SM: -13% .. +10%, stable per %smid), and every pivot waits for the slowest CTA.
    // Columns are therefore dealt in proportion to the speeds learned from earlier factorisations on this
    // context.  The result does not depend on who owns which column.
    const bool ranked = lazy && G == ctx->sm_count && G <= 256 && !getenv("TCI_RRLU_NO_BALANCE");
    const bool use_speeds = getenv("TCI_RRLU_BALANCE") != nullptr; // (see the note on read / write speeds below)
    std::vector<int> h_colmap, h_colcnt;
    if (ranked && getenv("TCI_RRLU_TEST_SPEEDS")) { // tests: a strongly uneven, reproducible speed table
        for (int q = 0; q < G; ++q) ctx->sm_speed[q] = 0.7 + 0.6 * ((q * 37) % G) / (double)G;
        ctx->sm_speed_valid = true;
    }
    if (ranked && ctx->sm_speed_valid && (use_speeds || getenv("TCI_RRLU_TEST_SPEEDS"))) {
        double tot = 0.0;
        for (int q = 0; q < G; ++q) tot += ctx->sm_speed[q];
        h_colcnt.assign(G, 0);
        std::vector<std::pair<double, int>> rem(G);
        i64 given = 0;
        for (int q = 0; q < G; ++q) {
            const double share = (double)n * ctx->sm_speed[q] / tot;
            h_colcnt[q] = (int)share;
            rem[q] = {share - h_colcnt[q], q};
            given += h_colcnt[q];
        }
        std::sort(rem.begin(), rem.end(), [](const std::pair<double, int> &x, const std::pair<double, int> &y) {
            return x.first > y.first || (x.first == y.first && x.second < y.second);
        });
        for (i64 k = 0; k < n - given; ++k) h_colcnt[rem[k % G].second]++;
        maxown = *std::max_element(h_colcnt.begin(), h_colcnt.end());
        h_colmap.assign((size_t)G * maxown, 0);
        std::vector<int> fill(G, 0);
        int q = 0;
        for (i64 j = 0; j < n; ++j) { // round-robin over the CTAs that still have room
            while (fill[q] >= h_colcnt[q]) q = (q + 1) % G;
            h_colmap[(size_t)q * maxown + fill[q]++] = (int)j;
            q = (q + 1) % G;
        }
    }
    size_t smem = lazy ? rrlu_lazy_smem(maxown, RRLU_LAZY_NB) : smem_need(G, resident);

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
    args.lds = lds;

    // One scratch arena: [result 4 int | err | pad][cand 2G x 16 B][pivvals][rowperm][colperm] (this prefix is
    // what comes back in ONE device-to-host copy) [colpos][pivrows]; the posted columns are separate.
    const size_t o_cand = 32, o_piv = o_cand + (size_t)2 * G * sizeof(RRCand), o_rp = o_piv + (size_t)mr * 8,
                 o_cp = o_rp + (size_t)m * 8, o_back = o_cp + (size_t)n * 8, o_pos = o_back,
                 o_prow = o_pos + (((size_t)n * 4 + 15) & ~(size_t)15),
                 o_rank = (o_prow + (size_t)mr * 4 + 31) & ~(size_t)15, // [startbar 16 B][smids G][passstats 2G]
                 o_stat = (o_rank + 16 + (size_t)G * 4 + 15) & ~(size_t)15, o_cmap = o_stat + (size_t)2 * G * 8,
                 o_end = o_cmap + (h_colmap.size() + h_colcnt.size()) * 4 + 16;
    DevBuf<char> arena(ctx);
    DevBuf<double> xbuf(ctx);
    TCI_CUDA(ctx, arena.alloc(o_end));
    args.nxslots = lazy ? RRLU_LAZY_NB + 1 : 2;
    if (!(G == 1 && resident)) TCI_CUDA(ctx, xbuf.alloc((size_t)args.nxslots * G * args.ldx)); // mode 0 posts nothing
    TCI_CUDA(ctx, cudaMemsetAsync(arena.p, 0, o_piv, ctx->stream));
    args.result = reinterpret_cast<int *>(arena.p);
    args.result_err = reinterpret_cast<double *>(arena.p + 16);
    args.cand = reinterpret_cast<RRCand *>(arena.p + o_cand);
    args.pivvals = reinterpret_cast<double *>(arena.p + o_piv);
    args.rowperm = reinterpret_cast<i64 *>(arena.p + o_rp);
    args.colperm = reinterpret_cast<i64 *>(arena.p + o_cp);
    args.colpos = reinterpret_cast<int *>(arena.p + o_pos);
    args.pivrows = reinterpret_cast<int *>(arena.p + o_prow);
    args.xbuf = xbuf.p;
    if (ranked) {
        TCI_CUDA(ctx, cudaMemsetAsync(arena.p + o_rank, 0, o_cmap - o_rank, ctx->stream));
        args.startbar = reinterpret_cast<unsigned *>(arena.p + o_rank);
        args.smids = reinterpret_cast<int *>(arena.p + o_rank + 16);
        args.passstats = reinterpret_cast<long long *>(arena.p + o_stat);
        if (!h_colmap.empty()) {
            TCI_CUDA(ctx, cudaMemcpyAsync(arena.p + o_cmap, h_colmap.data(), h_colmap.size() * 4, cudaMemcpyHostToDevice,
                                          ctx->stream));
            TCI_CUDA(ctx, cudaMemcpyAsync(arena.p + o_cmap + h_colmap.size() * 4, h_colcnt.data(), h_colcnt.size() * 4,
                                          cudaMemcpyHostToDevice, ctx->stream));
            args.colmap = reinterpret_cast<const int *>(arena.p + o_cmap);
            args.colcnt = reinterpret_cast<const int *>(arena.p + o_cmap + h_colmap.size() * 4);
        }
    }

    DevBuf<long long> dbg(ctx);
    const char *dbgenv = getenv("TCI_RRLU_DEBUG");
    if (dbgenv) {
        TCI_CUDA(ctx, dbg.alloc(16 + 41 * (size_t)G));
        TCI_CUDA(ctx, cudaMemsetAsync(dbg.p, 0, (16 + 41 * (size_t)G) * sizeof(long long), ctx->stream));
        args.dbg = dbg.p;
        args.dbg_cta = atoi(dbgenv) % G;
    }
    // page-locked: the result copy is pure latency (a deferred caller owns the staging buffer: do not touch it)
    char *back = deferred_result ? nullptr : static_cast<char *>(ctx_pinned(ctx, o_back + 64));
    if (!back && !deferred_result) return tci_fail(ctx, TCI_ERR_CUDA, "page-locked staging buffer");
    cudaEventRecord(ctx->ev0, ctx->stream);
    {
        if (getenv("TCI_RRLU_DEBUG"))
            fprintf(stderr, "[rrlu launch] m=%lld n=%lld G=%d lazy=%d ranked=%d maxown=%d smem=%zu\n", (long long)m,
                    (long long)n, G, (int)lazy, (int)ranked, maxown, smem);
        int rc = lazy ? rrlu_lazy_launch(ctx, args, G, smem, exact_mode != 0)
                      : rrlu_launch(ctx, args, G, T, smem, exact_mode != 0,
                                    G == 1 && resident ? 0 : (resident ? 1 : (xs_in_smem ? 2 : 3)));
        if (rc) return rc;
    }
    if (deferred_result) {
        TCI_CUDA(ctx, cudaMemcpyAsync(deferred_result, arena.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaMemcpyAsync(reinterpret_cast<char *>(deferred_result) + 32, arena.p + o_piv, (size_t)mr * 8,
                                      cudaMemcpyDeviceToHost, ctx->stream)); // the pivot values follow the 8 words
        *npivot = mr;
        tci_lu *lu = new tci_lu();
        lu->ctx = ctx;
        lu->A = A;
        lu->m = m;
        lu->n = n;
        lu->r = mr;
        lu->leftorthogonal = leftorthogonal != 0;
        lu->arena = arena.p;
        lu->d_rowperm = args.rowperm;
        lu->d_colperm = args.colperm;
        lu->d_colpos = args.colpos;
        arena.p = nullptr;
        ctx->live_handles++;
        *factors = lu;
        return TCI_OK;
    }
    TCI_CUDA(ctx, cudaMemcpyAsync(back, arena.p, o_back, cudaMemcpyDeviceToHost, ctx->stream));
    if (extra_bytes)
        TCI_CUDA(ctx, cudaMemcpyAsync(extra_host, extra_dev, extra_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev1, ctx->stream);
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev3) == cudaSuccess) ctx->stage_ms[ST_RRLU_KERNEL] += ms;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stage_ms[ST_RRLU] += ms;
    }
    if (ranked) { // learn the per-SM streaming speed from this run (tiles per cycle, normalised to mean 1)
        std::vector<long long> st(2 * (size_t)G);
        TCI_CUDA(ctx, cudaMemcpy(st.data(), arena.p + o_stat, st.size() * 8, cudaMemcpyDeviceToHost));
        bool ok = true;
        double mean = 0.0;
        std::vector<double> sp(G);
        for (int q = 0; q < G; ++q) {
            if (st[q] <= 0 || st[G + q] < 4096) ok = false; // too little work to say anything
            sp[q] = ok ? (double)st[G + q] / (double)st[q] : 1.0;
            mean += sp[q];
        }
        mean /= G;
        if (ok) {
            for (int q = 0; q < G; ++q) {
                double v = sp[q] / mean;
                v = v < 0.7 ? 0.7 : (v > 1.3 ? 1.3 : v);
                ctx->sm_speed[q] = ctx->sm_speed_valid ? 0.5 * (ctx->sm_speed[q] + v) : v;
            }
            ctx->sm_speed_valid = true;
        }
    }
    const int *res = reinterpret_cast<const int *>(back);
    const double lu_error = *reinterpret_cast<const double *>(back + 16);
    const int r = res[0];
    if (dbgenv) {
        long long h[16];
        cudaMemcpy(h, dbg.p, sizeof(h), cudaMemcpyDeviceToHost);
        if (lazy && getenv("TCI_RRLU_DEBUG_CTAS")) {
            std::vector<long long> pc(2 * (size_t)G);
            cudaMemcpy(pc.data(), dbg.p + 16, pc.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            fprintf(stderr, "[rrlu lazy pass cycles/pivot per CTA (cta:smid:cycles)]");
            for (int q = 0; q < G; ++q) fprintf(stderr, " %d:%lld:%lld", q, pc[G + q], pc[q] / (r + 1));
            fprintf(stderr, "\n");
            {
                std::vector<long long> ts(32 * (size_t)G);
                cudaMemcpy(ts.data(), dbg.p + 16 + 9 * G, ts.size() * sizeof(long long), cudaMemcpyDeviceToHost);
                for (int pz = 0; pz < 16 && pz + 16 < r; ++pz) {
                    const long long *b = &ts[(size_t)pz * 2 * G], *e = b + G;
                    long long b0 = *std::min_element(b, b + G), b1 = *std::max_element(b, b + G);
                    long long e0 = *std::min_element(e, e + G), e1 = *std::max_element(e, e + G);
                    long long dmin = 1LL << 60, dmax = 0;
                    for (int q = 0; q < G; ++q) {
                        dmin = std::min(dmin, e[q] - b[q]);
                        dmax = std::max(dmax, e[q] - b[q]);
                    }
                    fprintf(stderr, "[rrlu lazy pivot %d] pass start spread %lld ns, end spread %lld ns, first start -> last end %lld ns, "
                                    "per-CTA duration min %lld max %lld ns, last-ending cta %d\n",
                            pz + 16, b1 - b0, e1 - e0, e1 - b0, dmin, dmax, (int)(std::max_element(e, e + G) - e));
                }
            }
            std::vector<long long> ph(7 * (size_t)G);
            cudaMemcpy(ph.data(), dbg.p + 16 + 2 * G, ph.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            const char *nm[7] = {"wait", "book", "pass", "imbal", "special", "reduce", "post"};
            for (int k = 0; k < 7; ++k) {
                long long mn = ph[(size_t)k * G], mx = mn, sum = 0;
                int amx = 0;
                for (int q = 0; q < G; ++q) {
                    const long long v = ph[(size_t)k * G + q];
                    sum += v;
                    if (v < mn) mn = v;
                    if (v > mx) mx = v, amx = q;
                }
                fprintf(stderr, "[rrlu lazy phase %-8s cycles/pivot over CTAs] min %lld mean %lld max %lld (cta %d)\n", nm[k],
                        mn / (r + 1), sum / G / (r + 1), mx / (r + 1), amx);
            }
        }
        if (lazy)
            fprintf(stderr,
                    "[rrlu lazy dbg] m=%lld n=%lld r=%d G=%d cycles/pivot: wait+reduce %lld | bookkeeping+y %lld | pass %lld | "
                    "pass imbalance %lld | special+commit %lld | blockreduce %lld | post %lld\n",
                    (long long)m, (long long)n, r, G, h[0] / (r + 1), h[1] / (r + 1), h[2] / (r + 1), h[3] / (r + 1),
                    h[4] / (r + 1), h[5] / (r + 1), h[6] / (r + 1));
        else
        fprintf(stderr,
                "[rrlu dbg] m=%lld n=%lld r=%d G=%d T=%d res=%d cycles/pivot: wait+reduce %lld (poll %lld acqfence %lld "
                "payload %lld) | swap+x %lld | ys+L %lld | update %lld | blockreduce %lld | post %lld | fence %lld\n",
                (long long)m, (long long)n, r, G, T, (int)resident, h[0] / (r + 1), h[7] / (r + 1), h[8] / (r + 1),
                h[9] / (r + 1), h[1] / (r + 1), h[2] / (r + 1), h[3] / (r + 1), h[4] / (r + 1), h[5] / (r + 1),
                h[6] / (r + 1));
    }
    if ((res[1] & 1) || (res[2] & 1)) return tci_fail(ctx, TCI_ERR_NAN_L, "lu.L contains NaNs");
    if (res[2] & 2) return tci_fail(ctx, TCI_ERR_NAN_U, "lu.U contains NaNs");
    const double *pv = reinterpret_cast<const double *>(back + o_piv);
    const i64 *rp = reinterpret_cast<const i64 *>(back + o_rp);
    const i64 *cp = reinterpret_cast<const i64 *>(back + o_cp);
    for (i64 i = 0; i < m; ++i) rowperm[i] = rp[i] + 1;
    for (i64 j = 0; j < n; ++j) colperm[j] = cp[j] + 1;
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
        lu->arena = arena.p; // ownership of the arena moves to the handle
        lu->d_rowperm = args.rowperm;
        lu->d_colperm = args.colperm;
        lu->d_colpos = args.colpos;
        arena.p = nullptr;
        ctx->live_handles++;
        *factors = lu;
    }
    return TCI_OK;
}

// Full-rank factorisations (reltol = abstol = 0, maxrank = k, leftorthogonal, exact) of `nmat` small square matrices
// in as few cooperative launches as possible (k_rrlu_batch); matrices that do not fit the shared memory of their CTA
// group take the one-at-a-time deferred path of rrlu_core.  Nothing synchronises: results[q] (page-locked, 32 + 8 k_q
// bytes: the 8 result words, then the pivot values) is filled in stream order and checked by the caller.
int rrlu_batch_fullrank(tci_ctx *ctx, int nmat, tci_dmat *const *P, tci_lu **lus, char *const *results)
{
    const size_t SMEM_BUDGET = 220 * 1024;
    std::vector<int> todo;
    for (int q = 0; q < nmat; ++q) {
        lus[q] = nullptr;
        todo.push_back(q);
    }
    const bool no_batch = getenv("TCI_RRLU_NO_BATCH") != nullptr;
    while (!todo.empty()) {
        const int nb = (int)std::min<size_t>(todo.size(), (size_t)ctx->sm_count);
        i64 kmax = 1;
        for (int t = 0; t < nb; ++t) kmax = std::max<i64>(kmax, P[todo[t]]->m);
        int Gq = (int)std::max<i64>(1, std::min<i64>(std::min<i64>(16, ctx->sm_count / nb), kmax / 2));
        auto smem_of = [&](i64 k) {
            const i64 lds = (k + 1) & ~(i64)1, mo = (k + Gq - 1) / Gq;
            return (size_t)lds * 8 + (size_t)((mo + 1) & ~(i64)1) * 8 + (size_t)mo * 8 + (size_t)mo * lds * 8 + 64;
        };
        std::vector<int> batch, rest;
        for (int t = 0; t < nb; ++t)
            (!no_batch && P[todo[t]]->m <= RR_XS_CAP && smem_of(P[todo[t]]->m) <= SMEM_BUDGET ? batch : rest).push_back(todo[t]);
        for (int q : rest) { // too large for a CTA group: one at a time, still without synchronising
            i64 np = 0;
            double err = 0.0;
            const i64 k = P[q]->m;
            int rc = rrlu_core(ctx, P[q], k, k, k, 0.0, 0.0, 1, 1, nullptr, nullptr, &np, &err, nullptr, &lus[q], nullptr,
                               nullptr, 0, reinterpret_cast<int *>(results[q]));
            if (rc) return rc;
        }
        if (!batch.empty()) {
            const int B = (int)batch.size();
            std::vector<RRArgs> args((size_t)B);
            std::vector<double *> xbufs; // posted pivot columns: released (in stream order) AFTER the launch
            size_t smem = 0;
            int T = 256;
            for (int t = 0; t < B; ++t) {
                tci_dmat *A = P[batch[t]];
                const i64 k = A->m, lds = (k + 1) & ~(i64)1, mo = (k + Gq - 1) / Gq;
                dmat_wait_ready(ctx, A);
                smem = std::max(smem, smem_of(k));
                const i64 per_cta = k * mo;
                T = std::max(T, (k >= 1024 || per_cta >= 32768) ? 1024 : ((k >= 384 || per_cta >= 8192) ? 512 : 256));
                RRArgs &a = args[t];
                a = RRArgs{};
                a.A = A->p;
                a.m = a.n = k;
                a.ld = A->ld;
                a.maxrank = (int)k;
                a.reltol = a.abstol = 0.0;
                a.leftorth = 1;
                a.ldx = round_up(k, 16);
                a.xs_in_smem = 1;
                a.maxown = (int)mo;
                a.lds = lds;
                a.nxslots = 2;
                const size_t o_cand = 32, o_piv = o_cand + (size_t)2 * Gq * sizeof(RRCand), o_rp = o_piv + (size_t)k * 8,
                             o_cp = o_rp + (size_t)k * 8, o_pos = o_cp + (size_t)k * 8,
                             o_prow = o_pos + (((size_t)k * 4 + 15) & ~(size_t)15), o_end = o_prow + (size_t)k * 4 + 32;
                char *arena = nullptr;
                double *xbuf = nullptr;
                TCI_CUDA(ctx, dev_alloc(ctx, (void **)&arena, o_end));
                TCI_CUDA(ctx, dev_alloc(ctx, (void **)&xbuf, (size_t)2 * Gq * a.ldx * sizeof(double)));
                TCI_CUDA(ctx, cudaMemsetAsync(arena, 0, o_piv, ctx->stream));
                a.result = reinterpret_cast<int *>(arena);
                a.result_err = reinterpret_cast<double *>(arena + 16);
                a.cand = reinterpret_cast<RRCand *>(arena + o_cand);
                a.pivvals = reinterpret_cast<double *>(arena + o_piv);
                a.rowperm = reinterpret_cast<i64 *>(arena + o_rp);
                a.colperm = reinterpret_cast<i64 *>(arena + o_cp);
                a.colpos = reinterpret_cast<int *>(arena + o_pos);
                a.pivrows = reinterpret_cast<int *>(arena + o_prow);
                a.xbuf = xbuf;
                tci_lu *lu = new tci_lu();
                lu->ctx = ctx;
                lu->A = A;
                lu->m = lu->n = lu->r = k;
                lu->leftorthogonal = true;
                lu->arena = arena;
                lu->d_rowperm = a.rowperm;
                lu->d_colperm = a.colperm;
                lu->d_colpos = a.colpos;
                ctx->live_handles++;
                lus[batch[t]] = lu;
                xbufs.push_back(xbuf);
            }
            DevBuf<RRArgs> dargs(ctx);
            TCI_CUDA(ctx, dargs.upload(args.data(), args.size()));
            const void *fn = (const void *)k_rrlu_batch<true, true>;
            TCI_CUDA(ctx, ctx_func_smem(ctx, fn, 227 * 1024 - 1024));
            const RRArgs *dp = dargs.p;
            void *kargs[] = {(void *)&dp, (void *)&Gq};
            cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(B * Gq), dim3(T), kargs, smem, ctx->stream);
            for (double *xb : xbufs) dev_free(ctx, xb);
            TCI_CUDA(ctx, le);
            ctx->launches++;
            for (int t = 0; t < B; ++t) {
                const RRArgs &a = args[t];
                TCI_CUDA(ctx, cudaMemcpyAsync(results[batch[t]], a.result, 32, cudaMemcpyDeviceToHost, ctx->stream));
                TCI_CUDA(ctx, cudaMemcpyAsync(results[batch[t]] + 32, a.pivvals, (size_t)a.m * 8, cudaMemcpyDeviceToHost,
                                              ctx->stream));
            }
        }
        todo.erase(todo.begin(), todo.begin() + nb);
    }
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
    if (A_dev && A_dev->ctx != ctx) return tci_fail(ctx, TCI_ERR_ARG, "tci_rrlu: the matrix belongs to another context");
    if (!A_dev) {
        int rc = tci_dmat_create(ctx, m, n, A_host, &A);
        if (rc) return rc;
    }
    int rc;
    {
        TCI_ENTER(ctx);
        rc = rrlu_core(ctx, A, m, n, maxrank, reltol, abstol, leftorthogonal, exact_mode, rowperm, colperm, npivot,
                       error, pivoterrors, factors, nullptr, nullptr, 0, nullptr);
    }
    if (!A_dev && !(factors && *factors)) tci_dmat_destroy(A); // our own upload, not handed to a factors handle
    return rc;
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
    dev_free(ctx, lu->arena);
    ctx->live_handles--;
    tci_dmat_destroy(lu->A); // releases the context if it was destroyed before its handles
    delete lu;
    return TCI_OK;
}
