// rrlu_lazy.cu -- K2 for matrices that do not fit in shared memory: the same full-pivot rrLU as
// rrlu.cu (same pivots, same bits), but the Schur updates are DEFERRED.
//
// The in-place kernel reads and writes the whole trailing matrix once per pivot (16 B per element
// per pivot).  Here the matrix in HBM is only rewritten every NB pivots ("commit"); in between, a
// pass re-reads the committed matrix and applies the nd <= NB pending rank-1 updates in registers,
//      v = a;  v = v - x_1*y_1;  v = v - x_2*y_2; ...        (each product and difference rounded)
// which is exactly the sequence of roundings the reference performs (matrixlu.jl:132), so every
// trailing value -- and therefore every pivot decision -- is bit-identical.  Traffic drops to
// 8 + 8/NB bytes per element per pivot; the extra multiplies/subtracts (NB+1)/2 per element on
// average) are far below the FP64 pipe's capacity at HBM speed.
//
// Rows: the reference swaps rows physically at every pivot.  Inside a block of NB pivots the swaps
// are virtual (a list of at most 2*NB "special" base rows with their current positions); the commit
// pass writes every row to the position the reference would hold it at.  Special rows (pivot rows
// and the rows they displaced) are handled separately from the streaming loop, which only ever
// touches rows whose position equals their base row.
//
// Columns, the candidate records that double as the grid barrier, the posted pivot columns and the
// stop rule are as in rrlu.cu.  A posted pivot column is indexed by BASE row and stays valid for its
// whole block (xbuf has NB+1 slots per CTA instead of 2).
#include <set>

#include "rrlu_common.cuh"

#define RL_THREADS 512
#define RL_R 2               // 16-byte loads per lane per column tile
#define RL_RG (RL_R * 64)    // rows of a column tile
// ring stages (1 KB tiles in flight) per warp.  Measured at 8192^2 r=1024 / 16384^2 r=128 (ms): 4: 140.5 / 66.6,
// 6: 140.3 / 65.7, 8: 142.3 / 69.8, 12: 149.1 / 82.1 -- deeper queues cost bandwidth instead of hiding latency.
#ifndef RL_S
#define RL_S 4 // slots of two tiles (2 KB)
#endif



// One active column as the streaming pass sees it (rebuilt in entry order before every pass).
struct __align__(16) RLEnt {
    double *ptr; // column in HBM
    int cp;      // its position
    int slot;
};

__device__ __forceinline__ int hi32(double q) { return __double2hiint(q); }
// predicated 16-byte store (kept opaque so that the compiler does not clone the arithmetic around a branch)
__device__ __forceinline__ void st_pred_f64x2(double *p, double a, double b, bool pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\t@p st.global.v2.f64 [%0], {%1, %2};\n\t}" ::"l"(p), "d"(a),
                 "d"(b), "r"((int)pred)
                 : "memory");
}

// ---- per-warp TMA ring -------------------------------------------------------------------------
// Every warp streams its tiles (RL_RG rows of one column = 1 KB contiguous in HBM) through a private ring of
// RL_S shared-memory stages filled by cp.async.bulk (one instruction per tile, issued by lane 0, completion
// on an mbarrier).  The bytes in flight no longer live in registers: 16 warps x (RL_S-1) KB per SM cover the
// HBM latency-bandwidth product, which 3 register-resident tiles per warp did not (measured: 3.3 TB/s).
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ double2 lds_f64x2(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}

// The tile stream of one warp in one pass: for every work item (RL_RG rows x a chunk of active columns) first the
// nd pending pivot columns x_i restricted to those rows, then the chunk's column tiles.  The producer side
// of the ring walks this stream RL_S tiles ahead of the consumer.  All source pointers come from one shared
// table: entries [-nd, 0) are the pivot columns, entries [0, nact) the active columns.
struct RLStream {
    // pass constants
    unsigned tab;  // shared address of table entry 0 (16-byte entries, pointer first)
    int nd, m, i0, ntr, CH, nact, nitems, step, estep; // estep = +1 / -1: direction of the sweep
    // ring: shared addresses of this warp's stage 0 / barrier 0, consumer and producer stage, consumer parity
    unsigned ring0, bar0, cs, cpar, ps;
    // producer cursor: item, tile within the item, item length, current table entry, row offset, bytes
    int pt, pj, p_len, p_efirst, p_row;
    unsigned p_bytes;

    __device__ __forceinline__ void item(int t, int &row, int &c0, int &cnt) const
    {
        const int ch = t / ntr, rg = t - ch * ntr;
        c0 = ch * CH;
        cnt = ((c0 + CH < nact) ? c0 + CH : nact) - c0;
        row = i0 + rg * RL_RG;
    }
    __device__ __forceinline__ void p_set()
    {
        if (pt >= 0 && pt < nitems) {
            int c0, cnt;
            item(pt, p_row, c0, cnt);
            const int rows = (m - p_row < RL_RG) ? m - p_row : RL_RG;
            p_bytes = ((unsigned)rows * 8u + 15u) & ~15u;
            p_len = nd + cnt;
            p_efirst = estep > 0 ? c0 : c0 + cnt - 1;
            pj = 0;
        } else
            p_len = 0; // end of the stream
    }
    // entry of tile number j of the item the producer is in: pivot columns first, then the chunk's columns
    __device__ __forceinline__ int p_entry(int j) const { return j < nd ? j - nd : p_efirst + (j - nd) * estep; }
    // A ring slot holds TWO consecutive tiles of the stream (the last slot of an item only one when the item has
    // an odd number of tiles): one mbarrier phase, one acquire / release and one cursor step per pair.
    __device__ __forceinline__ void produce(int lane)
    {
        if (p_len == 0) return;
        const bool two = pj + 1 < p_len;
        if (lane == 0) {
            const unsigned bar = bar0 + (ps << 3), dst = ring0 + (ps << 11);
            unsigned long long b0, b1 = 0ull;
            asm volatile("ld.shared.u64 %0, [%1];" : "=l"(b0) : "r"(tab + 16u * (unsigned)p_entry(pj)) : "memory");
            if (two) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(b1) : "r"(tab + 16u * (unsigned)p_entry(pj + 1)) : "memory");
            mbar_expect_tx(bar, two ? 2u * p_bytes : p_bytes);
            bulk_g2s(dst, reinterpret_cast<const double *>(b0) + p_row, p_bytes, bar);
            if (two) bulk_g2s(dst + 1024u, reinterpret_cast<const double *>(b1) + p_row, p_bytes, bar);
        }
        ps = (ps + 1u == RL_S) ? 0u : ps + 1u;
        pj += 2;
        if (pj >= p_len) {
            pt += step;
            p_set();
        }
    }
    // wait for the next slot; returns the shared address of its first byte
    __device__ __forceinline__ unsigned acquire()
    {
        const unsigned bar = bar0 + (cs << 3);
        while (!mbar_try_wait(bar, cpar)) {}
        return ring0 + (cs << 11);
    }
    // the lanes have their values in registers: hand the slot back and keep the ring full
    __device__ __forceinline__ void release(int lane)
    {
        __syncwarp();
        cs = (cs + 1u == RL_S) ? 0u : cs + 1u;
        cpar ^= (cs == 0u) ? 1u : 0u;
        produce(lane);
    }
};
static_assert(RL_RG * 8 == 1024, "ring addressing uses shifts");

// The column tiles of one work item.  The lane's 2*RL_R rows are fixed, so the pending pivot columns x_i live
// in registers; per column only the tile itself and the y_i move.  NDV >= nd updates are always applied:
// the missing ones (and every masked row) have x = y = 0, and v - 0*0 == v exactly, so there is no
// per-element predication and a masked entry is written back unchanged by a commit.
// Arg-max: the squares are non-negative, so their high words order them coarsely; the exact (value, column
// position, row) comparison of matrixlu.jl:16-29 only runs when a tile reaches the lane's current best.
template <bool EXACT, int NB, int NDV, bool COMMIT>
__device__ __forceinline__ void rl_item(RLStream &st, unsigned yEs, int efirst, int cnt, int rbase, int m, unsigned negm,
                                        int lane, unsigned open_slot, const double2 (&xr)[NB][RL_R],
                                        unsigned long long &bvb, int &bhi, int &bcpv, int &browv)
{
    static_assert(NB == 4 && NDV >= 1 && NDV <= 4, "y_i are fetched as one or two 16-byte pairs");
    int nmw[2 * RL_R];
    bool inb[RL_R];
#pragma unroll
    for (int w = 0; w < 2 * RL_R; ++w) {
        nmw[w] = ((negm >> w) & 1u) ? -1 : 0;
        asm volatile("" : "+r"(nmw[w])); // keep the masks in registers instead of re-deriving them per column
    }
#pragma unroll
    for (int u = 0; u < RL_R; ++u) {
        int ib = rbase + u * 64 < m;
        asm volatile("" : "+r"(ib));
        inb[u] = ib != 0;
    }
    unsigned lane16 = 16u * (unsigned)lane;
    asm volatile("" : "+r"(lane16));
    // one column tile whose 1 KB sits at shared address `sa`
    auto tile = [&](int e, unsigned sa) {
        unsigned long long cbase = 0ull;
        int cp;
        if (COMMIT) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(cbase) : "r"(st.tab + 16u * (unsigned)e) : "memory");
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(cp) : "r"(st.tab + 16u * (unsigned)e + 8u) : "memory");
        double y[NDV];
        {
            const double2 y01 = lds_f64x2(yEs + 32u * (unsigned)e);
            y[0] = y01.x;
            if (NDV >= 2) y[1] = y01.y;
            if (NDV >= 3) {
                const double2 y23 = lds_f64x2(yEs + 32u * (unsigned)e + 16u);
                y[2] = y23.x;
                if (NDV >= 4) y[3] = y23.y;
            }
        }
        double2 d[RL_R];
#pragma unroll
        for (int u = 0; u < RL_R; ++u) d[u] = lds_f64x2(sa + lane16 + 512u * u);
        double *const cptr = reinterpret_cast<double *>(cbase) + rbase;
        double q[2 * RL_R];
        int tmax = -1;
#pragma unroll
        for (int u = 0; u < RL_R; ++u) {
            double v0 = d[u].x, v1 = d[u].y;
#pragma unroll
            for (int i = 0; i < NDV; ++i) {
                v0 = schur<EXACT>(v0, xr[i][u].x, y[i]);
                v1 = schur<EXACT>(v1, xr[i][u].y, y[i]);
            }
            if (COMMIT) st_pred_f64x2(cptr + u * 64, v0, v1, inb[u]);
            q[2 * u] = v0 * v0;
            q[2 * u + 1] = v1 * v1;
            tmax = max(tmax, max(hi32(q[2 * u]) | nmw[2 * u], hi32(q[2 * u + 1]) | nmw[2 * u + 1]));
        }
        // bhi is warp-uniform: the high word of the best square any lane of this warp has seen.  An element below it
        // cannot win, so the exact path runs about once per new warp-wide maximum instead of once per lane maximum
        // (a lane-local threshold sent 21 % of the tiles through it, because one lane is enough to divert the warp).
        if (__any_sync(0xffffffffu, tmax >= bhi)) {
            if (tmax >= bhi) { // exact comparison
                unsigned long long tb = 0ull;
                int tj = 0;
#pragma unroll
                for (int w = 0; w < 2 * RL_R; ++w) {
                    const bool ok = nmw[w] == 0 && q[w] == q[w];
                    const unsigned long long ob = ok ? (unsigned long long)__double_as_longlong(q[w]) + 1ull : 0ull;
                    if (ob > tb) {
                        tb = ob;
                        tj = w;
                    }
                }
                const int trow = rbase + (tj >> 1) * 64 + (tj & 1);
                if (tb > bvb || (tb == bvb && tb != 0ull && (cp < bcpv || (cp == bcpv && trow < browv)))) {
                    bvb = tb;
                    bcpv = cp;
                    browv = trow;
                }
            }
            const int mine = bvb ? (int)((bvb - 1ull) >> 32) : -1;
            bhi = max(bhi, (int)__reduce_max_sync(0xffffffffu, mine));
        }
    };
    int e = efirst, idx = 0;
    if (open_slot) { // the slot's first half was the last pivot-column tile of this item
        tile(e, open_slot + 1024u);
        st.release(lane);
        e += st.estep;
        idx = 1;
    }
#pragma unroll 1
    for (; idx + 1 < cnt; idx += 2, e += 2 * st.estep) {
        const unsigned sa = st.acquire();
        tile(e, sa);
        tile(e + st.estep, sa + 1024u);
        st.release(lane);
    }
    if (idx < cnt) { // an odd tile closes the item: it has a slot of its own
        const unsigned sa = st.acquire();
        tile(e, sa);
        st.release(lane);
    }
}

template <bool EXACT, bool LEFT, int NB> __global__ void __launch_bounds__(RL_THREADS, 1) k_rrlu_lazy(RRArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int G = gridDim.x, T = blockDim.x, tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const int m = (int)a.m, n = (int)a.n;
    const int MO = a.maxown;
    // Logical CTA index.  With speed-weighted ownership the CTAs are ranked by the SM they run on (one CTA per
    // SM), so that a table indexed by the rank always addresses the same SM; otherwise it is blockIdx.x.
    int g = blockIdx.x;
    if (a.smids) {
        __shared__ int sh_rank;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (tid == 0) {
            sh_rank = 0;
            a.smids[blockIdx.x] = (int)smid + 1;
            __threadfence();
            atomicAdd(a.startbar, 1u);
            while (*reinterpret_cast<volatile unsigned *>(a.startbar) < (unsigned)G) {}
            __threadfence();
        }
        __syncthreads();
        int below = 0;
        for (int b = tid; b < G; b += T) {
            const int sb = __ldcg(a.smids + b);
            below += (sb < (int)smid + 1 || (sb == (int)smid + 1 && b < (int)blockIdx.x)) ? 1 : 0;
        }
        if (below) atomicAdd(&sh_rank, below);
        __syncthreads();
        g = sh_rank;
    }

    double *ys = reinterpret_cast<double *>(smem_raw); // [NB][MO]: pivot row j in the own column slot o
    int *acto = reinterpret_cast<int *>(ys + (size_t)NB * MO);
    int *actp = acto + MO;    // position of the active entry
    int *slotpos = actp + MO; // position of every own slot
    int *pcol = slotpos + MO; // physical column of every own slot
    // table of streamed columns: entries [-NB, 0) pending pivot columns, [0, MO) active columns in entry order
    RLEnt *ent = reinterpret_cast<RLEnt *>(smem_raw + (((size_t)NB * MO * 8 + (size_t)4 * MO * 4 + 15) & ~(size_t)15)) + NB;
    double *yE = reinterpret_cast<double *>(ent + MO); // [MO][NB] pivot-row values in entry order
    unsigned char *ringp = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<size_t>(yE + (size_t)NB * MO) + 127) & ~(size_t)127); // [nwarps][RL_S][RL_RG] doubles
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(ringp + (size_t)nwarps * RL_S * RL_RG * 16);
#define RL_COL(o) (a.A + (size_t)a.ld * pcol[o])

    long long tmark = clock64();
    long long pass_total = 0, pass_t0 = 0, pass_work = 0;
    __shared__ long long dbg_acc[8];
    if (tid < 8) dbg_acc[tid] = 0;
#define RL_MARK(ph)                                    \
    do {                                               \
        if (a.dbg && tid == 0) {                       \
            long long now__ = clock64();               \
            dbg_acc[ph] += now__ - tmark;              \
            tmark = now__;                             \
        }                                              \
    } while (0)
    __shared__ unsigned long long red_v[32], red_key[32];
    __shared__ int red_slot;
    __shared__ double win_val;
    __shared__ int win_row, win_colpos, win_cta;
    __shared__ int sh_nact, sh_removed;
    __shared__ int sp_base[2 * NB], sp_pos[2 * NB], sh_nsp; // special rows: base row -> current position
    __shared__ int prow[NB], pslot[NB];                     // pending pivots: base row, own slot of the column (-1)
    __shared__ double pval[NB];
    __shared__ const double *xptr[NB]; // posted pivot columns of the pending pivots (indexed by base row)

    const int nown = a.colmap ? a.colcnt[g] : ((g < n) ? (n - g + G - 1) / G : 0);
    _Pragma("unroll 1") for (int e = tid; e < nown; e += T) {
        const int j = a.colmap ? a.colmap[(size_t)g * MO + e] : g + e * G;
        pcol[e] = j;
        acto[e] = e;
        actp[e] = j;
        slotpos[e] = j;
        a.colpos[j] = j;
    }
    if (tid == 0) {
        sh_nact = nown;
        sh_removed = -1;
        sh_nsp = 0;
    }
    if (g == 0)
        _Pragma("unroll 1") for (int i = tid; i < m; i += T) a.rowperm[i] = i;
    RLStream st;
    st.ring0 = smem_u32(ringp + (size_t)warp * RL_S * RL_RG * 16);
    st.bar0 = smem_u32(mbar + warp * RL_S);
    st.cs = st.cpar = st.ps = 0u;
    if (lane == 0) {
        for (int k = 0; k < RL_S; ++k) mbar_init(st.bar0 + 8u * k, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();

    double maxerror = 0.0;
    double lasterr = nan("");
    int npiv = 0;
    int flags = 0;
    int k0 = 0; // pivots committed to HBM
    int nd = 0; // pending pivots (k0 + nd == number of pivots made)

#pragma unroll 1
    for (int s = -1;; ++s) {
        int nact = sh_nact;
        bool fin = false;
        if (s >= 0) {
            // ---- 1. wait for / reduce the posted candidates (as rrlu.cu) ------------------
            if (warp == 0) {
                const RRCand *cd = a.cand + (size_t)(s & 1) * G;
                const unsigned phase = (((unsigned)s >> 1) & 1u) ^ 1u;
                double cv[RR_MAXQ];
                unsigned crp[RR_MAXQ];
                int ccp[RR_MAXQ];
                for (;;) {
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
                    if (lane == 0) win_cta = -1;
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
            }
            asm volatile("fence.proxy.async;" ::: "memory"); // posted columns are read through the async proxy
            __syncthreads();
            RL_MARK(0); // wait + reduce
            const int wcta = win_cta;
            const double val = win_val;
            const int pr = win_row; // POSITION of the pivot row
            const int pcpos = win_colpos;
            const double err = fabs(val);
            if (wcta < 0) { // nothing but NaNs left in the trailing block
                flags |= 1;
                fin = true;
            } else {
                lasterr = err;
                // ---- 2. stop rule  matrixlu.jl:153-158 ---------------------------------
                if (s > 0 && (err < a.reltol * maxerror || err < a.abstol)) fin = true;
            }
            if (!fin) {
                maxerror = (isnan(maxerror) || isnan(err)) ? nan("") : (err > maxerror ? err : maxerror);
                npiv = s + 1;
                if (g == 0 && tid == 0) {
                    a.pivrows[s] = pr;
                    a.pivvals[s] = val;
                }
                // ---- 3. virtual row swap s <-> pr, column bookkeeping ---------------------
                const int nsp0 = sh_nsp;
                int p = pr, q = s; // base rows sitting at positions pr and s
                for (int k = 0; k < nsp0; ++k) {
                    if (sp_pos[k] == pr) p = sp_base[k];
                    if (sp_pos[k] == s) q = sp_base[k];
                }
                __syncthreads(); // every thread has read the list before it changes
                const int j = nd;
                if (tid == 0) {
                    int kp = -1, kq = -1, cnt = nsp0;
                    for (int k = 0; k < cnt; ++k) {
                        if (sp_base[k] == p) kp = k;
                        if (sp_base[k] == q) kq = k;
                    }
                    if (kp < 0) {
                        kp = cnt++;
                        sp_base[kp] = p;
                    }
                    sp_pos[kp] = s;
                    if (q != p) {
                        if (kq < 0) {
                            kq = cnt++;
                            sp_base[kq] = q;
                        }
                        sp_pos[kq] = pr;
                    }
                    sh_nsp = cnt;
                    prow[j] = p;
                    pval[j] = val;
                    pslot[j] = -1;
                    xptr[j] = a.xbuf + ((size_t)(s % a.nxslots) * G + wcta) * a.ldx;
                }
                _Pragma("unroll 1") for (int e = tid; e < nact; e += T) {
                    if (actp[e] == pcpos) {
                        sh_removed = e;
                    } else if (actp[e] == s) {
                        actp[e] = pcpos;
                        slotpos[acto[e]] = pcpos;
                        a.colpos[pcol[acto[e]]] = pcpos;
                    }
                }
                __syncthreads();
                const int rem = sh_removed;
                if (rem >= 0) {
                    const int oslot = acto[rem];
                    __syncthreads();
                    if (tid == 0) {
                        const int jp = pcol[oslot];
                        a.colpos[jp] = s;
                        a.colperm[s] = jp;
                        slotpos[oslot] = s;
                        pslot[j] = oslot;
                        const int last = nact - 1;
                        acto[rem] = acto[last];
                        actp[rem] = actp[last];
                        sh_nact = last;
                        sh_removed = -1;
                    }
                    __syncthreads();
                }
                nact = sh_nact;
                // ---- 4. pivot row in the own active columns: y_j = A_j[p, :] -----------------
                {
                    double xpj[NB];
#pragma unroll
                    for (int i = 0; i < NB; ++i) xpj[i] = (i < j) ? __ldcg(xptr[i] + p) : 0.0;
                    _Pragma("unroll 1") for (int e = tid; e < nact; e += T) {
                        const int o = acto[e];
                        double y = RL_COL(o)[p];
#pragma unroll
                        for (int i = 0; i < NB; ++i)
                            if (i < j) y = schur<EXACT>(y, xpj[i], ys[i * MO + o]);
                        if (!LEFT) y = __ddiv_rn(y, val); // matrixlu.jl:122
                        ys[j * MO + o] = y;
                    }
                }
                nd = j + 1;
                __syncthreads();
                if (s + 1 >= a.maxrank) fin = true; // the last Schur update never reaches L or U
            }
        }

        RL_MARK(1); // bookkeeping + pivot row
        if (tid == 0) pass_t0 = clock64();
        if (a.dbg && tid == 0 && s >= 16 && s < 32) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            a.dbg[16 + 9 * G + (size_t)(s - 16) * 2 * G + g] = (long long)gt;
        }
        const int lo = k0 + nd; // first trailing position
        const bool commit = fin ? (nd > 0) : (nd == NB);
        unsigned long long bvb = 0ull;
        int bcpv = 0x7fffffff, browv = 0x7fffffff;

        // ---- 5. streaming pass over the normal rows: pending updates in registers, arg-max, (write) --
        if (!fin && nact > 0) {
            _Pragma("unroll 1") for (int e = tid; e < nact; e += T) {
                const int o = acto[e];
                RLEnt en;
                en.ptr = RL_COL(o);
                en.cp = actp[e];
                en.slot = o;
                ent[e] = en;
                for (int i = 0; i < NB; ++i) yE[e * NB + i] = (i < nd) ? ys[i * MO + o] : 0.0;
            }
            if (tid < nd) { // the pending pivot columns precede the active columns in the stream table
                RLEnt en;
                en.ptr = const_cast<double *>(xptr[tid]);
                en.cp = en.slot = -1;
                ent[tid - nd] = en;
            }
            __syncthreads();
            int pr_[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) pr_[i] = (i < nd) ? prow[i] : -1;
            int bhi = -1;
            const int i0 = lo & ~1;
            const int ntr = (m - i0 + RL_RG - 1) / RL_RG;
            // Split the active columns into nch chunks so that the ntr*nch items deal evenly over the warps: every
            // item costs its chunk plus the nd pivot-column tiles it has to pull through the ring first.
            int nch = 1;
            {
                int best = 0x7fffffff;
                const int nchmax = (nact + 5) / 6;
                constexpr int NW = RL_THREADS / 32; // == nwarps (the kernel is always launched with RL_THREADS)
#pragma unroll
                for (int c = 1; c <= 8; ++c) { // constant divisors: this runs once per pivot in every thread
                    const int cost = ((ntr * c + NW - 1) / NW) * ((nact + c - 1) / c + nd + 1);
                    if (c <= nchmax && cost < best) {
                        best = cost;
                        nch = c;
                    }
                }
            }
            const int CH = (nact + nch - 1) / nch;
            nch = (nact + CH - 1) / CH;
            const int nitems = ntr * nch;
            const bool backward = (s & 1);
            int t = warp;
            if (backward && nitems > warp) t = warp + ((nitems - 1 - warp) / nwarps) * nwarps;
            st.tab = smem_u32(ent);
            st.nd = nd;
            st.m = m;
            st.i0 = i0;
            st.ntr = ntr;
            st.CH = CH;
            st.nact = nact;
            st.nitems = nitems;
            st.step = backward ? -nwarps : nwarps;
            st.estep = backward ? -1 : 1;
            st.pt = t;
            st.p_set();
            unsigned yEs = smem_u32(yE);
            // opaque copies: otherwise the shared-memory addresses are re-derived from the carve-up in every
            // iteration of the tile loop (a dozen integer instructions per tile)
            asm volatile("" : "+r"(yEs), "+r"(st.tab), "+r"(st.ring0), "+r"(st.bar0));
#pragma unroll 1
            for (int k = 0; k < RL_S; ++k) st.produce(lane); // fill the ring
#pragma unroll 1
            for (; t >= 0 && t < nitems; t += st.step) {
                int row0, c0, cnt;
                st.item(t, row0, c0, cnt);
                const int rbase = row0 + 2 * lane;
                unsigned negm = 0u;
                bool okr[2 * RL_R];
#pragma unroll
                for (int u = 0; u < RL_R; ++u) {
                    const int r = rbase + u * 64;
                    bool ok0 = r >= lo && r < m, ok1 = r + 1 < m;
#pragma unroll
                    for (int i = 0; i < NB; ++i) {
                        ok0 = ok0 && r != pr_[i];
                        ok1 = ok1 && r + 1 != pr_[i];
                    }
                    okr[2 * u] = ok0;
                    okr[2 * u + 1] = ok1;
                    negm |= (ok0 ? 0u : 1u) << (2 * u) | (ok1 ? 0u : 2u) << (2 * u);
                }
                double2 xr[NB][RL_R];
                unsigned open_slot = 0u; // shared address of a slot whose second half is still to be consumed
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    if (i < nd) { // the x_i tiles of this item come through the ring first, two per slot
                        if ((i & 1) == 0) open_slot = st.acquire();
                        const unsigned sa = open_slot + ((i & 1) ? 1024u : 0u) + 16u * (unsigned)lane;
#pragma unroll
                        for (int u = 0; u < RL_R; ++u) {
                            const double2 xx = lds_f64x2(sa + 512u * u);
                            xr[i][u].x = okr[2 * u] ? xx.x : 0.0;
                            xr[i][u].y = okr[2 * u + 1] ? xx.y : 0.0;
                        }
                        if (i & 1) {
                            st.release(lane);
                            open_slot = 0u;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < RL_R; ++u) xr[i][u] = make_double2(0.0, 0.0);
                    }
                }
                const int efirst = backward ? c0 + cnt - 1 : c0;
                // one instantiation per number of pending updates (nd == NB is always the commit)
                if (nd <= 1)
                    rl_item<EXACT, NB, 1, false>(st, yEs, efirst, cnt, rbase, m, negm, lane, open_slot, xr, bvb, bhi, bcpv, browv);
                else if (nd == 2)
                    rl_item<EXACT, NB, 2, false>(st, yEs, efirst, cnt, rbase, m, negm, lane, open_slot, xr, bvb, bhi, bcpv, browv);
                else if (nd == 3)
                    rl_item<EXACT, NB, 3, false>(st, yEs, efirst, cnt, rbase, m, negm, lane, open_slot, xr, bvb, bhi, bcpv, browv);
                else if (commit)
                    rl_item<EXACT, NB, NB, true>(st, yEs, efirst, cnt, rbase, m, negm, lane, open_slot, xr, bvb, bhi, bcpv, browv);
                else
                    rl_item<EXACT, NB, NB, false>(st, yEs, efirst, cnt, rbase, m, negm, lane, open_slot, xr, bvb, bhi, bcpv, browv);
            }
        }
        RL_MARK(2); // streaming pass (this warp)
        __syncthreads();
        if (a.dbg && tid == 0 && s >= 16 && s < 32) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            a.dbg[16 + 9 * G + (size_t)(s - 16) * 2 * G + G + g] = (long long)gt;
        }
        if (tid == 0 && !fin) {
            pass_total += clock64() - pass_t0;
            pass_work += (long long)nact * ((m - (lo & ~1) + RL_RG - 1) / RL_RG) * (commit ? 2 : 1);
        }
        RL_MARK(3); // wait for the other warps

        // ---- 6. special rows: displaced rows take part in the arg-max; at a commit every special row of
        //         every own column moves to its position with the value the reference holds there ------
        {
            const int nsp = sh_nsp;
            if (nsp > 0) {
                const int ncol = commit ? nown : nact;
                const int cpr = T / nsp; // whole columns per round
                const int ci_l = tid / nsp, k = tid - ci_l * nsp;
#pragma unroll 1
                for (int c0 = 0; c0 < ncol; c0 += cpr) {
                    const int ci = c0 + ci_l;
                    const bool act = ci_l < cpr && ci < ncol;
                    double v = 0.0;
                    double *col = nullptr;
                    int np = 0;
                    if (act) {
                        const int o = commit ? ci : acto[ci];
                        const int r = sp_base[k];
                        np = sp_pos[k];
                        col = RL_COL(o);
                        const int cp = slotpos[o];
                        int jj = -1;
                        for (int i = 0; i < nd; ++i)
                            if (prow[i] == r) jj = i;
                        if (cp < k0) { // picked in an earlier block: the row only moves
                            v = col[r];
                        } else if (cp < lo) { // picked in this block as pivot i
                            const int i = cp - k0;
                            if (jj < 0 || jj > i)
                                v = __ldcg(xptr[i] + r); // L entry
                            else if (jj == i)
                                v = pval[i];
                            else
                                v = ys[jj * MO + o]; // U entry
                        } else if (jj >= 0) { // active column, pivot row: U entry
                            v = ys[jj * MO + o];
                        } else { // active column, displaced row: all pending updates
                            v = col[r];
                            for (int i = 0; i < nd; ++i) v = schur<EXACT>(v, __ldcg(xptr[i] + r), ys[i * MO + o]);
                            const double q = v * v;
                            if (!fin && q == q) {
                                const unsigned long long tb = (unsigned long long)__double_as_longlong(q) + 1ull;
                                if (tb > bvb || (tb == bvb && (cp < bcpv || (cp == bcpv && np < browv)))) {
                                    bvb = tb;
                                    bcpv = cp;
                                    browv = np;
                                }
                            }
                        }
                    }
                    if (commit) {
                        __syncthreads(); // all special rows of these columns are read before any is written
                        if (act) col[np] = v;
                    }
                }
            }
            if (commit) {
                // L columns of the pivots picked in this block (normal rows): A[r, c_j] = x_j[r]
                for (int j = 0; j < nd; ++j) {
                    const int os = pslot[j];
                    if (os < 0) continue;
                    double *col = RL_COL(os);
                    const double *xj = xptr[j];
                    _Pragma("unroll 1") for (int r = lo + tid; r < m; r += T) {
                        bool sp = false;
                        for (int i = 0; i < nd; ++i) sp = sp || (r == prow[i]);
                        if (!sp) col[r] = __ldcg(xj + r);
                    }
                }
                __threadfence(); // committed values are read back by cp.async.bulk in the next pass
                asm volatile("fence.proxy.async;" ::: "memory");
                __syncthreads();
                k0 = lo;
                nd = 0;
                if (tid == 0) sh_nsp = 0;
                __syncthreads();
            }
        }
        RL_MARK(4); // special rows, commit extras
        if (fin) break;

        // ---- 7. block reduction of the candidate ---------------------------------------
        unsigned long long bkey = ((unsigned long long)(unsigned)bcpv << 32) | (unsigned)browv;
        unsigned long long vb = bvb;
        warp_argmax(vb, bkey);
        if (lane == 0) {
            red_v[warp] = vb;
            red_key[warp] = bkey;
        }
        if (tid == 0) red_slot = -1;
        __syncthreads();
        vb = lane < nwarps ? red_v[lane] : 0ull;
        bkey = lane < nwarps ? red_key[lane] : ~0ull;
        warp_argmax(vb, bkey);
        const int bcp = (int)(bkey >> 32), brow = (int)(bkey & 0xffffffffull);
        if (vb != 0ull)
            _Pragma("unroll 1") for (int e = tid; e < nact; e += T)
                if (actp[e] == bcp) red_slot = acto[e];
        __syncthreads();
        const int bslot = red_slot;
        RL_MARK(5); // block reduce
        // ---- 8. post the candidate's column (all pending updates applied) and its record ----
        {
            const int nxt = (s + 1) & 1;
            RRCand *slot = a.cand + (size_t)nxt * G + g;
            double cval = 0.0;
            if (bslot >= 0) {
                const int nsp = sh_nsp;
                int b = brow; // base row of the candidate
                for (int k = 0; k < nsp; ++k)
                    if (sp_pos[k] == brow) b = sp_base[k];
                const double *col = RL_COL(bslot);
                double yb[NB];
                const double *xq[NB];
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    yb[i] = (i < nd) ? ys[i * MO + bslot] : 0.0;
                    xq[i] = (i < nd) ? xptr[i] : nullptr;
                }
                double *xo = a.xbuf + ((size_t)((s + 1) % a.nxslots) * G + g) * a.ldx;
                constexpr int PB = 4; // rows in flight per thread: the loop is latency bound (64 KB per CTA)
                int r0 = (k0 & ~1) + 2 * tid;
                double2 d[PB], xx[NB][PB];
                // the first batch of column loads is issued before the candidate's own value is waited for
#pragma unroll
                for (int k = 0; k < PB; ++k)
                    if (r0 + 2 * T * k < m) d[k] = *reinterpret_cast<const double2 *>(col + r0 + 2 * T * k);
                cval = col[b];
#pragma unroll
                for (int i = 0; i < NB; ++i)
                    if (i < nd) cval = schur<EXACT>(cval, __ldcg(xq[i] + b), yb[i]);
                bool preloaded = true;
                _Pragma("unroll 1") for (; r0 < m; r0 += 2 * T * PB) {
#pragma unroll
                    for (int k = 0; k < PB; ++k) {
                        const int r = r0 + 2 * T * k;
                        if (!preloaded && r < m) d[k] = *reinterpret_cast<const double2 *>(col + r);
#pragma unroll
                        for (int i = 0; i < NB; ++i)
                            xx[i][k] = (i < nd && r < m) ? __ldcg(reinterpret_cast<const double2 *>(xq[i] + r))
                                                         : make_double2(0.0, 0.0);
                    }
                    preloaded = false;
#pragma unroll
                    for (int k = 0; k < PB; ++k) {
                        const int r = r0 + 2 * T * k;
#pragma unroll
                        for (int i = 0; i < NB; ++i) { // yb = x = 0 beyond nd: exact no-op
                            d[k].x = schur<EXACT>(d[k].x, xx[i][k].x, yb[i]);
                            d[k].y = schur<EXACT>(d[k].y, xx[i][k].y, yb[i]);
                        }
                        if (LEFT) {
                            d[k].x = __ddiv_rn(d[k].x, cval);
                            d[k].y = __ddiv_rn(d[k].y, cval);
                        }
                        if (r < m) *reinterpret_cast<double2 *>(xo + r) = d[k];
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                const unsigned phase = ((((unsigned)(s + 1)) >> 1) & 1u) ^ 1u;
                __threadfence();
                st_relaxed_16(slot, cval, (unsigned)(bslot >= 0 ? brow : 0) | (phase << 31), bslot >= 0 ? bcp : -1);
            }
            RL_MARK(6); // post
        }
    }
    if (a.dbg && g == a.dbg_cta && tid < 8) a.dbg[tid] = dbg_acc[tid];
    if (a.passstats && tid == 0) {
        a.passstats[g] = pass_total;
        a.passstats[G + g] = pass_work;
    }
    if (a.dbg && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.dbg[16 + g] = pass_total;
        a.dbg[16 + G + g] = smid;
        for (int k = 0; k < 7; ++k) a.dbg[16 + (2 + k) * G + g] = dbg_acc[k];
    }

    __syncthreads();
    { // NaN scan of the L and U parts of the own columns (matrixlu.jl:164-169): bit 0 L, bit 1 U
        int f = 0;
#pragma unroll 1
        for (int o = warp; o < nown; o += nwarps) {
            const int cp = slotpos[o];
            const double *col = RL_COL(o);
            const int hi = cp < npiv ? m : npiv;
            for (int i = lane; i < hi; i += 32)
                if (isnan(col[i])) f |= ((cp < npiv && i >= cp) ? 1 : 0) | ((i < npiv && cp >= i) ? 2 : 0);
        }
        if (f) atomicOr(&a.result[2], f);
    }
    {
        const int nact = sh_nact;
        _Pragma("unroll 1") for (int e = tid; e < nact; e += T) a.colperm[actp[e]] = pcol[acto[e]];
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
        *a.result_err = (npiv >= mn) ? 0.0 : lasterr;
    }
#undef RL_COL
}

size_t rrlu_lazy_smem(int maxown, int nb)
{
    return (size_t)nb * maxown * 8 + (size_t)4 * maxown * 4 + 16 + (size_t)(maxown + nb) * sizeof(RLEnt) + (size_t)nb * maxown * 8 +
           128 + (size_t)(RL_THREADS / 32) * RL_S * (RL_RG * 16 + 8) + 64;
}

template <bool EXACT, int NB> static const void *lazy_fn(bool left)
{
    return left ? (const void *)k_rrlu_lazy<EXACT, true, NB> : (const void *)k_rrlu_lazy<EXACT, false, NB>;
}

int rrlu_lazy_launch(tci_ctx *ctx, RRArgs &args, int G, size_t smem, bool exact)
{
    void *kargs[] = {&args};
    const void *fn = exact ? lazy_fn<true, RRLU_LAZY_NB>(args.leftorth != 0) : lazy_fn<false, RRLU_LAZY_NB>(args.leftorth != 0);
    TCI_CUDA(ctx, ctx_func_smem(ctx, fn, 227 * 1024 - 1024));
    cudaEventRecord(ctx->ev2, ctx->stream);
    TCI_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(RL_THREADS), kargs, smem, ctx->stream));
    cudaEventRecord(ctx->ev3, ctx->stream);
    ctx->launches++;
    return TCI_OK;
}
