// pi_eval.cu -- K1: batched evaluation of the Pi / T tensors for analytic targets
// (replaces the triple loop of batcheval.jl:50-58 and the max-abs pass of
// util.jl:1-10, fused), plus the tci_pi_eval / tci_target_eval entry points.
//
// Layout (Appendix A.1 of SURVEY.md): out[i + nI*c + ld*j] = f(I_i ++ c ++ J_j),
// centre multi-index c enumerated first-index-fastest.  The kernel is bound by
// the 8 B/evaluation HBM write for cheap targets: every thread owns two
// consecutive rows (one 16 B store) and walks a tile of columns, so a warp writes
// 512 contiguous bytes per column.
#include <cstring>

#include "tci_internal.h"

#define PI_THREADS 256
#define PI_TCOLS 32

// Row, centre and column states in ONE launch: thread q < nI handles left multi-index q (prefix state
// from tci_target_init), the next C threads the centre combinations (first centre index fastest), the last
// nJ threads the right multi-indices.  st[k * count + q] = state component k of entry q.
__global__ void k_states_all(tci_analytic_t t, const i64 *__restrict__ I, int nl, i64 nI, int M, i64 C,
                             const i64 *__restrict__ J, int nr, i64 nJ, double *__restrict__ rs,
                             double *__restrict__ cs, double *__restrict__ js, int *__restrict__ csig)
{
    i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    double s[TCI_MAX_STATE];
    if (q < nI) {
        tci_target_init(&t, s);
        for (int k = 0; k < nl; ++k) tci_target_accum(&t, k, I[(i64)nl * q + k], s);
        for (int k = 0; k < t.nstate; ++k) rs[(i64)k * nI + q] = s[k];
        return;
    }
    q -= nI;
    for (int k = 0; k < TCI_MAX_STATE; ++k) s[k] = 0.0;
    if (q < C) {
        i64 rem = q;
        for (int k = 0; k < M; ++k) {
            const i64 d = t.localdims[nl + k];
            const i64 sig = rem % d + 1;
            rem /= d;
            tci_target_accum(&t, nl + k, sig, s);
            if (csig) csig[(i64)k * C + q] = (int)sig;
        }
        for (int k = 0; k < t.nstate; ++k) cs[(i64)k * C + q] = s[k];
        return;
    }
    q -= C;
    if (q < nJ && js) {
        for (int k = 0; k < nr; ++k) tci_target_accum(&t, nl + M + k, J[(i64)nr * q + k], s);
        for (int k = 0; k < t.nstate; ++k) js[(i64)k * nJ + q] = s[k];
    }
}

__device__ __forceinline__ unsigned long long absbits(double v)
{ // |v| as an ordered integer; NaN sorts above +Inf, so an integer max propagates NaN like Julia's max
    return (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
}

__device__ __forceinline__ void block_max_commit(unsigned long long mx, unsigned long long *gmax)
{
    __shared__ unsigned long long wm[32];
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = other > mx ? other : mx;
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) wm[w] = mx;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        mx = l < nw ? wm[l] : 0ull;
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, mx, o);
            mx = other > mx ? other : mx;
        }
        if (l == 0 && mx) atomicMax(gmax, mx);
    }
}

// Exact targets: state(I ++ c ++ J) = rowstate + centrestate + colstate, bit-identical to
// the sequential definition because all partial sums are exact.
template <int NS, int KIND>
__global__ void __launch_bounds__(PI_THREADS)
    k_pi_exact(tci_analytic_t targ, const double *__restrict__ rs, i64 nI, const double *__restrict__ cs, i64 C,
               const double *__restrict__ js, i64 nJ, double *__restrict__ out, i64 ld, unsigned long long *gmax,
               int tcols)
{
    __shared__ double colst[NS][PI_TCOLS];
    __shared__ i64 coloff[PI_TCOLS];
    tci_analytic_t t = targ;
    t.kind = KIND; // compile-time kind: the switch in tci_target_finalize folds away
    const i64 ncols = C * nJ;
    const i64 q0 = (i64)blockIdx.y * tcols;
    if (threadIdx.x < tcols) {
        i64 q = q0 + threadIdx.x;
        if (q < ncols) {
            i64 c = q % C, j = q / C;
#pragma unroll
            for (int k = 0; k < NS; ++k) colst[k][threadIdx.x] = TCI_ADD(cs[(i64)k * C + c], js[(i64)k * nJ + j]);
            coloff[threadIdx.x] = nI * c + ld * j;
        }
    }
    __syncthreads();
    const i64 r0 = ((i64)blockIdx.x * PI_THREADS + threadIdx.x) * 2;
    unsigned long long mx = 0ull;
    if (r0 < nI) {
        const bool two = r0 + 1 < nI;
        double a0[NS], a1[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            a0[k] = rs[(i64)k * nI + r0];
            a1[k] = two ? rs[(i64)k * nI + r0 + 1] : 0.0;
        }
        const int nq = (int)(ncols - q0 < tcols ? ncols - q0 : tcols);
#pragma unroll 4
        for (int qq = 0; qq < nq; ++qq) {
            double s0[NS], s1[NS];
#pragma unroll
            for (int k = 0; k < NS; ++k) {
                double cst = colst[k][qq];
                s0[k] = TCI_ADD(a0[k], cst);
                s1[k] = TCI_ADD(a1[k], cst);
            }
            double v0 = tci_target_finalize(&t, s0);
            unsigned long long b0 = absbits(v0);
            mx = b0 > mx ? b0 : mx;
            double *dst = out + coloff[qq] + r0;
            if (two) {
                double v1 = tci_target_finalize(&t, s1);
                unsigned long long b1 = absbits(v1);
                mx = b1 > mx ? b1 : mx;
                if ((((size_t)dst) & 15) == 0)
                    __stcs(reinterpret_cast<double2 *>(dst), make_double2(v0, v1));
                else {
                    dst[0] = v0;
                    dst[1] = v1;
                }
            } else
                dst[0] = v0;
        }
    }
    block_max_commit(mx, gmax);
}

// Non-exact targets: the state of row i is the sequential prefix over the left sites;
// every element continues the accumulation over the centre and right sites in site order,
// which is exactly the scalar definition tci_target_eval.
__global__ void __launch_bounds__(PI_THREADS)
    k_pi_sequential(tci_analytic_t t, const double *__restrict__ rs, i64 nI, int nl, int M, i64 C,
                    const int *__restrict__ csig, const i64 *__restrict__ J, int nr, i64 nJ, double *__restrict__ out,
                    i64 ld, unsigned long long *gmax, int tcols)
{
    const i64 ncols = C * nJ;
    const i64 q0 = (i64)blockIdx.y * tcols;
    const i64 r = (i64)blockIdx.x * PI_THREADS + threadIdx.x;
    unsigned long long mx = 0ull;
    if (r < nI) {
        double a[TCI_MAX_STATE];
        for (int k = 0; k < TCI_MAX_STATE; ++k) a[k] = k < t.nstate ? rs[(i64)k * nI + r] : 0.0;
        const int nq = (int)(ncols - q0 < tcols ? ncols - q0 : tcols);
        for (int qq = 0; qq < nq; ++qq) {
            i64 q = q0 + qq, c = q % C, j = q / C;
            double s[TCI_MAX_STATE];
            for (int k = 0; k < TCI_MAX_STATE; ++k) s[k] = a[k];
            for (int k = 0; k < M; ++k) tci_target_accum(&t, nl + k, csig[(i64)k * C + c], s);
            for (int k = 0; k < nr; ++k) tci_target_accum(&t, nl + M + k, J[(i64)nr * j + k], s);
            double v = tci_target_finalize(&t, s);
            unsigned long long b = absbits(v);
            mx = b > mx ? b : mx;
            out[nI * c + ld * j + r] = v;
        }
    }
    block_max_commit(mx, gmax);
}

// Columns per CTA: PI_TCOLS for large Pi (the column states are staged once per CTA); small matrices -- what a TCI run
// at chi <= 256 produces -- get fewer columns per thread so that the grid still covers the SMs a few times over
// (a 1126 x 524 Pi was 51 CTAs with 64 serial evaluations per thread).
static int pick_tcols(tci_ctx *ctx, i64 rowblocks, i64 ncols)
{
    int t = PI_TCOLS;
    while (t > 2 && rowblocks * ((ncols + t - 1) / t) < 4 * (i64)ctx->sm_count) t >>= 1;
    return t;
}

template <int NS, int KIND>
static void launch_exact(tci_ctx *ctx, const tci_analytic_t &an, const double *rs, i64 nI, const double *cs, i64 C,
                         const double *js, i64 nJ, double *out, i64 ld, unsigned long long *gmax)
{
    const i64 rb = (nI + 2 * PI_THREADS - 1) / (2 * PI_THREADS);
    const int tc = pick_tcols(ctx, rb, C * nJ);
    dim3 grid((unsigned)rb, (unsigned)((C * nJ + tc - 1) / tc));
    k_pi_exact<NS, KIND><<<grid, PI_THREADS, 0, ctx->stream>>>(an, rs, nI, cs, C, js, nJ, out, ld, gmax, tc);
    ctx->launches++;
}

int pi_eval_analytic(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                     tci_dmat *out, unsigned long long *d_maxbits)
{
    const tci_analytic_t &an = t.an;
    const int NS = an.nstate;
    i64 C = 1;
    for (i64 k = 0; k < M; ++k) C *= t.localdims[nl + k];
    const bool exact = tci_target_exact(an.kind) != 0;
    // one scratch allocation: [rs NS*nI][cs NS*C][js NS*nJ][csig M*C ints]
    DevBuf<double> st(ctx);
    const size_t n_rs = (size_t)NS * nI, n_cs = (size_t)NS * C, n_js = exact ? (size_t)NS * nJ : 0;
    TCI_CUDA(ctx, st.alloc(n_rs + n_cs + n_js + (exact ? 0 : ((size_t)(M > 0 ? M : 1) * C + 1) / 2 + 1)));
    struct {
        double *p;
    } rs{st.p}, cs{st.p + n_rs}, js{st.p + n_rs + n_cs};
    struct {
        int *p;
    } csig{exact ? nullptr : reinterpret_cast<int *>(st.p + n_rs + n_cs + n_js)};
    const int TB = 128;
    const i64 nthreads = nI + C + (exact ? nJ : 0);
    k_states_all<<<(unsigned)((nthreads + TB - 1) / TB), TB, 0, ctx->stream>>>(
        an, dI, (int)nl, nI, (int)M, C, dJ, (int)nr, nJ, rs.p, cs.p, exact ? js.p : nullptr, csig.p);
    ctx->launches++;
    if (exact) {
#define PI_LAUNCH(NSV, KINDV) launch_exact<NSV, KINDV>(ctx, an, rs.p, nI, cs.p, C, js.p, nJ, out->p, out->ld, d_maxbits)
        switch (an.kind) {
        case TCI_TARGET_LORENTZ: PI_LAUNCH(1, TCI_TARGET_LORENTZ); break;
        case TCI_TARGET_SUM: PI_LAUNCH(1, TCI_TARGET_SUM); break;
        case TCI_TARGET_QUANTICS2D: PI_LAUNCH(2, TCI_TARGET_QUANTICS2D); break;
        case TCI_TARGET_TABLE: PI_LAUNCH(1, TCI_TARGET_TABLE); break;
        case TCI_TARGET_QUANTICS1D: PI_LAUNCH(1, TCI_TARGET_QUANTICS1D); break;
        case TCI_TARGET_SEPCOS:
            switch (NS) {
            case 1: PI_LAUNCH(1, TCI_TARGET_SEPCOS); break;
            case 2: PI_LAUNCH(2, TCI_TARGET_SEPCOS); break;
            case 3: PI_LAUNCH(3, TCI_TARGET_SEPCOS); break;
            case 4: PI_LAUNCH(4, TCI_TARGET_SEPCOS); break;
            case 5: PI_LAUNCH(5, TCI_TARGET_SEPCOS); break;
            case 6: PI_LAUNCH(6, TCI_TARGET_SEPCOS); break;
            default: return tci_fail(ctx, TCI_ERR_ARG, "target has an unsupported number of state sums");
            }
            break;
        default: return tci_fail(ctx, TCI_ERR_ARG, "unknown exact target kind");
        }
#undef PI_LAUNCH
    } else {
        const i64 rb = (nI + PI_THREADS - 1) / PI_THREADS;
        const int tc = pick_tcols(ctx, rb, C * nJ);
        dim3 grid((unsigned)rb, (unsigned)((C * nJ + tc - 1) / tc));
        k_pi_sequential<<<grid, PI_THREADS, 0, ctx->stream>>>(an, rs.p, nI, (int)nl, (int)M, C, csig.p, dJ, (int)nr,
                                                              nJ, out->p, out->ld, d_maxbits, tc);
        ctx->launches++;
    }
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

// Contraction.f: the elementwise function of the product (contraction.jl:203-205, 330-332), by id
__global__ void k_apply_f(double *__restrict__ p, i64 m, i64 n, i64 ld, int kind, double a, double b)
{
    const i64 total = m * n;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        double *x = p + (e % m) + ld * (e / m);
        const double v = *x;
        *x = kind == TCI_F_AFFINE ? __dadd_rn(__dmul_rn(a, v), b) : (kind == TCI_F_ABS ? fabs(v) : __dmul_rn(v, v));
    }
}

int apply_elementwise(tci_ctx *ctx, const TargetDev &t, double *p, i64 m, i64 n, i64 ld)
{
    if (t.fkind == TCI_F_NONE || m * n == 0) return TCI_OK;
    const unsigned blocks = (unsigned)std::min<i64>((m * n + 255) / 256, (i64)ctx->sm_count * 8);
    k_apply_f<<<blocks, 256, 0, ctx->stream>>>(p, m, n, ld, t.fkind, t.fa, t.fb);
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

extern "C" int tci_target_set_elementwise(tci_ctx *ctx, int64_t target_id, int kind, double a, double b)
{
    TCI_ENTER(ctx);
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &t = *it->second;
    if (t.kind != 2) return tci_fail(ctx, TCI_ERR_ARG, "tci_target_set_elementwise: only a Contraction carries a function f");
    if (kind < TCI_F_NONE || kind > TCI_F_SQUARE)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_target_set_elementwise: unknown function id");
    if (t.is_complex && kind != TCI_F_NONE && kind != TCI_F_AFFINE)
        return tci_fail(ctx, TCI_ERR_UNSUPPORTED, "tci_target_set_elementwise: a ComplexF64 contraction takes TCI_F_AFFINE only");
    t.fkind = kind;
    t.fa = a;
    t.fb = b;
    if (ctx->grp)
        for (int k = 1; k < ctx->grp->nlocal; ++k) {
            auto jt = ctx->grp->m[k]->targets.find(target_id);
            if (jt == ctx->grp->m[k]->targets.end()) continue;
            jt->second->fkind = kind;
            jt->second->fa = a;
            jt->second->fb = b;
        }
    return TCI_OK;
}

// max |x| over an m x n matrix (for targets whose Pi comes out of a GEMM)
__global__ void k_maxabs(const double *__restrict__ p, i64 m, i64 n, i64 ld, unsigned long long *gmax)
{
    unsigned long long mx = 0ull;
    i64 total = m * n;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % m, j = e / m;
        unsigned long long b = absbits(p[i + ld * j]);
        mx = b > mx ? b : mx;
    }
    block_max_commit(mx, gmax);
}

int maxabs_dev(tci_ctx *ctx, const double *p, i64 m, i64 n, i64 ld, unsigned long long *d_maxbits)
{
    i64 total = m * n;
    unsigned blocks = (unsigned)std::min<i64>((total + 255) / 256, (i64)ctx->sm_count * 8);
    k_maxabs<<<blocks ? blocks : 1, 256, 0, ctx->stream>>>(p, m, n, ld, d_maxbits);
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Enqueue / finish split.  pi_enqueue uploads the index sets and launches the evaluation on ctx->stream WITHOUT
// synchronising, so that a consumer (the rrLU of tci_bond_update, the solves of tci_fill_sitetensors) can be
// queued right behind it; max|.| is accumulated into *d_maxbits (device word, zeroed by the caller; nullable).
int pi_enqueue(tci_ctx *ctx, TargetDev &t, const i64 *I, i64 nl, i64 nI, const i64 *J, i64 nr, i64 nJ, i64 M,
               tci_dmat *out, unsigned long long *d_maxbits)
{
    DevBuf<i64> idx(ctx); // released in stream order after the kernels that read it
    TCI_CUDA(ctx, idx.alloc(2 + (size_t)(nl * nI) + (size_t)(nr * nJ)));
    i64 *dI = idx.p + 2, *dJ = idx.p + 2 + nl * nI;
    unsigned long long *dmax = d_maxbits ? d_maxbits : reinterpret_cast<unsigned long long *>(idx.p);
    if (!d_maxbits) TCI_CUDA(ctx, cudaMemsetAsync(idx.p, 0, 16, ctx->stream));
    if (nl * nI > 0)
        TCI_CUDA(ctx, cudaMemcpyAsync(dI, I, (size_t)(nl * nI) * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
    if (nr * nJ > 0)
        TCI_CUDA(ctx, cudaMemcpyAsync(dJ, J, (size_t)(nr * nJ) * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
    if (!ctx->nosync) cudaEventRecord(ctx->ev0, ctx->stream); // index upload is accounted as H2D, the kernels as Pi
    int rc = 0;
    switch (t.kind) {
    case 0: rc = pi_eval_analytic(ctx, t, dI, nl, nI, dJ, nr, nJ, M, out, dmax); break;
    case 1:
        rc = pi_eval_tt(ctx, t, dI, nl, nI, dJ, nr, nJ, M, out);
        if (!rc && d_maxbits) rc = maxabs_dev(ctx, out->p, out->m, out->n, out->ld, dmax);
        break;
    case 3: rc = pi_eval_user(ctx, t, dI, nl, nI, dJ, nr, nJ, M, out, dmax); break;
    case 4: rc = pi_eval_cached(ctx, t.cache_id, t, dI, nl, nI, dJ, nr, nJ, M, out, d_maxbits ? dmax : nullptr); break;
    default:
        rc = pi_eval_mpo(ctx, t, dI, nl, nI, dJ, nr, nJ, M, out, I, J);
        if (!rc) rc = apply_elementwise(ctx, t, out->p, out->m, out->n, out->ld);
        if (!rc && d_maxbits) rc = maxabs_dev(ctx, out->p, out->m, out->n, out->ld, dmax);
        break;
    }
    return rc;
}

unsigned long long *ctx_words(tci_ctx *ctx); // 64 persistent device words per context (below)

// ---- multi-GPU sharding of one evaluation -----------------------------------------------------
// Cost model.  Pi has to end up in the HBM of the rrLU owner, so every element evaluated elsewhere crosses NVLink
// once (peer stores from the evaluation kernel): ~8 B / 750 GB/s = 10.7 ps per remote element into ONE GPU.  With n
// shards the stage takes max(t_eval / n, 10.7 ps * (n-1)/n) per element + a fixed cost (worker wake-up, one NCCL
// all-reduce); the number of shards is the n <= world that minimises it, and 1 (no sharding) unless it wins by 10 %.
static double analytic_ps_per_eval(const TargetDev &t)
{
    switch (t.an.kind) {
    case TCI_TARGET_SEPCOS: return 1.5 + 3.6 * (double)(t.an.nstate - 1);
    case TCI_TARGET_QUANTICS2D: return 12.0;
    case TCI_TARGET_QUANTICS1D: return 6.0;
    case TCI_TARGET_GKCOSEXP: return 10.0;
    default: return 1.3; // HBM-write bound (8 B at ~6.2 TB/s)
    }
}
static int choose_shards(tci_ctx *ctx, const TargetDev &t, i64 nl, i64 nI, i64 nr, i64 nJ, i64 M, i64 C)
{
    const int world = ctx_world(ctx);
    if (world == 1) return 1;
    if (const char *f = getenv("TCI_SHARD_FORCE")) return std::max(1, std::min(world, atoi(f)));
    const double elems = (double)nI * (double)C * (double)nJ;
    const double fixed_ps = 60e6; // ~60 us
    if (t.kind == 0 || t.kind == 3) {
        const double te = t.kind == 3 ? 10.0 : analytic_ps_per_eval(t); // user source: assume a few transcendentals
        double best = te * elems;
        int bestn = 1;
        for (int n = 2; n <= world; ++n) {
            if (nJ < 2 * (i64)n) break;
            const double tn = std::max(te / n, 10.7 * (n - 1) / n) * elems + fixed_ps;
            if (tn < 0.9 * best && tn < (bestn == 1 ? 0.9 * te * elems : best)) {
                best = tn;
                bestn = n;
            }
        }
        return bestn;
    }
    if (M != 0) return 1; // T tensors (M = 1) are small; the row-block form below needs contiguous rows
    // TT / MPO pair: nearly all the flops sit in the two environment chains, which shard by rows / columns
    double D2 = 0.0;
    const i64 n = t.nsites;
    for (i64 s = 0; s < n; ++s) {
        if (s >= nl && s < n - nr) continue;
        const double cnt = s < nl ? (double)nI : (double)nJ;
        D2 += cnt * (t.kind == 1 ? 2.0 * t.dl[s] * t.dr[s]
                                 : 2.0 * t.adl[s] * t.bdl[s] * t.as2[s] * t.adr[s] + 2.0 * t.bdl[s] * t.as2[s] * t.adr[s] * t.bdr[s]);
    }
    const double ps = D2 / 20.0; // ~20 TFLOP/s = 20 flop per ps
    if (ps < 4.0 * fixed_ps || nI < 2 * (i64)world || nJ < 2 * (i64)world) return 1;
    return world;
}

// sharded M = 0 evaluation of an analytic target: column blocks, every member's kernel stores straight into the
// owner's HBM (peer st.global over NVLink: compute and transfer are one kernel); one NCCL all-reduce(max) of the
// max|.| bits follows on every member's stream and orders the owner's consumer behind all the peer stores.
static int pi_enqueue_sharded_analytic(tci_ctx *ctx, i64 target_id, int nshard, const i64 *I, i64 nl, i64 nI,
                                       const i64 *J, i64 nr, i64 nJ, i64 M, double *dst, i64 ld, i64 rows,
                                       unsigned long long *d_maxbits)
{
    tci_group *g = ctx->grp;
    const i64 blk = (nJ + nshard - 1) / nshard;
    group_follow_owner(g); // dst was allocated in the owner's stream order
    int rc = group_run(g, [&](int k) -> int {
        tci_ctx *c = g->m[k];
        unsigned long long *w = ctx_words(c);
        if (!w) return tci_fail(c, TCI_ERR_CUDA, "scratch words");
        if (k != 0) cudaMemsetAsync(w, 0, 8, c->stream); // the owner accumulates into the caller's word
        const int rank = c->rank;
        const i64 lo = std::min(nJ, rank * blk), hi = rank < nshard ? std::min(nJ, (rank + 1) * blk) : lo;
        if (hi <= lo) return TCI_OK;
        tci_dmat view;
        view.ctx = c;
        view.p = dst + ld * lo;
        view.m = rows;
        view.n = view.ncap = hi - lo;
        view.ld = ld;
        view.owned = false;
        TargetDev &t = *c->targets.at(target_id);
        return pi_enqueue(c, t, I, nl, nI, J + nr * lo, nr, hi - lo, M, &view,
                          rank == 0 ? d_maxbits : w);
    });
    if (rc) return rc;
    return group_allreduce_max_u64(g, [&](int k) { return g->m[k]->rank == 0 ? d_maxbits : ctx_words(g->m[k]); }, 1);
}

// sharded M = 0 evaluation of a TT / MPO-pair target by ROW blocks (SURVEY 8e "MPO x MPO contraction: row blocks"):
// every member extends the right environments of its column block, ONE NCCL all-gather shares them, the member
// extends the left environments of its own rows, and the GEMM epilogue of its block Pi = left^T right stores into
// the owner's Pi at the row offset.
// The order in which a sharded TT / contraction Pi deals its rows (side 0: lexicographic in the multi-index, first site
// most significant -- entries that share a prefix are neighbours) or columns (side 1: last site most significant --
// shared suffixes) to the GPUs; stable, so equal entries keep the caller's order.  perm[q] = caller's index of the q-th
// entry of that order.  Host only.
extern "C" int tci_shard_order(const int64_t *idx, int64_t len, int64_t count, int side, int64_t *perm)
{
    if (count < 0 || len < 0 || !perm || (len > 0 && count > 0 && !idx) || (side != 0 && side != 1)) return TCI_ERR_ARG;
    for (i64 q = 0; q < count; ++q) perm[q] = q;
    if (side == 0)
        std::stable_sort(perm, perm + count, [&](i64 a, i64 b) {
            return std::lexicographical_compare(idx + len * a, idx + len * (a + 1), idx + len * b, idx + len * (b + 1));
        });
    else
        std::stable_sort(perm, perm + count, [&](i64 a, i64 b) {
            for (i64 s = len - 1; s >= 0; --s)
                if (idx[len * a + s] != idx[len * b + s]) return idx[len * a + s] < idx[len * b + s];
            return false;
        });
    return TCI_OK;
}

// dst[rowmap[r] + ld * colmap[c]] = blk[r + rows * c]: a GPU's block of Pi, computed in prefix-sorted order, to the
// caller's row / column positions of the owner's matrix (8-byte stores; peer stores over NVLink when dst is remote)
__global__ void k_scatter_block(const double *__restrict__ blk, i64 rows, i64 ncols, const i64 *__restrict__ rowmap,
                                const i64 *__restrict__ colmap, double *__restrict__ dst, i64 ld)
{
    const i64 total = rows * ncols;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 r = e % rows, c = e / rows;
        dst[rowmap[r] + ld * colmap[c]] = blk[e];
    }
}

static int pi_enqueue_sharded_env(tci_ctx *ctx, i64 target_id, const i64 *I, i64 nl, i64 nI, const i64 *J, i64 nr,
                                  i64 nJ, double *dst, i64 ld, unsigned long long *d_maxbits)
{
    tci_group *g = ctx->grp;
    const int world = g->world;
    const i64 cblk = (nJ + world - 1) / world;
    const i64 rblk = (nI + world - 1) / world;
    // Prefix-aware partition.  An environment is shared by all entries with the same partial index (the reference's Dict
    // memo, contraction.jl:112-176), so the rows a GPU receives should share prefixes with EACH OTHER: the rows are put
    // in lexicographic order of their multi-indices (first site most significant; the columns by suffix: last site most
    // significant) and every GPU takes a contiguous block of that order.  For kronecker(Iset, d) -- i fastest, then sigma
    // -- a block of the caller's order holds one sigma of many different i, and every GPU would extend the whole prefix
    // chain of the i it touches once per sigma; with the sorted order a GPU holds all sigma of its i.  The block product is
    // computed in sorted order and scattered to the caller's rows / columns of the owner's Pi (peer stores).
    std::vector<i64> rperm((size_t)nI), cperm((size_t)nJ), Is((size_t)(nl * nI)), Js((size_t)(nr * nJ));
    tci_shard_order(I, nl, nI, 0, rperm.data());
    tci_shard_order(J, nr, nJ, 1, cperm.data());
    for (i64 q = 0; q < nI; ++q) std::copy(I + nl * rperm[q], I + nl * (rperm[q] + 1), Is.begin() + nl * q);
    for (i64 q = 0; q < nJ; ++q) std::copy(J + nr * cperm[q], J + nr * (cperm[q] + 1), Js.begin() + nr * q);
    I = Is.data();
    J = Js.data();
    std::vector<double *> renv(g->nlocal, nullptr);
    // TCI_SHARD_DEBUG: per-member timeline of the phases (events on the members' main streams)
    static const bool dbg = getenv("TCI_SHARD_DEBUG") != nullptr;
    std::vector<cudaEvent_t> dev_ev;
    auto mark = [&](int k, int phase) {
        if (!dbg) return;
        cudaEventRecord(dev_ev[(size_t)k * 5 + phase], g->m[k]->stream);
    };
    // Two streams per GPU.  The right-environment chain and the all-gather that follows it run on the member's
    // high-priority side stream, the left-environment chain on its main stream, concurrently: at 8 GPUs a chain level
    // is a batch of 128 small GEMMs = 1024 thread blocks = 3.46 waves, and the other chain's blocks fill the idle tail
    // of every launch (measured per GPU, one after the other: right chain 12.0 ms + left chain 10.5 ms).
    struct StreamSwap {
        tci_ctx *c;
        cudaStream_t main;
        explicit StreamSwap(tci_ctx *ctx_) : c(ctx_), main(ctx_->stream) { c->stream = c->copy_stream; }
        ~StreamSwap() { c->stream = main; }
    };
    if (dbg) {
        dev_ev.resize((size_t)g->nlocal * 5);
        for (int k = 0; k < g->nlocal; ++k) {
            cudaSetDevice(g->m[k]->device);
            for (int q = 0; q < 5; ++q) cudaEventCreate(&dev_ev[(size_t)k * 5 + q]);
        }
        cudaSetDevice(ctx->device);
    }
    i64 D = 1;
    {
        const TargetDev &t0 = *ctx->targets.at(target_id);
        const i64 n = t0.nsites;
        D = nr == 0 ? 1 : (t0.kind == 1 ? t0.dl[n - nr] : t0.adl[n - nr] * t0.bdl[n - nr]);
    }
    int rc = group_run(g, [&](int k) -> int {
        tci_ctx *c = g->m[k];
        TargetDev &t = *c->targets.at(target_id);
        const int rank = c->rank;
        TCI_CUDA(c, dev_alloc(c, (void **)&renv[k], (size_t)D * cblk * world * sizeof(double)));
        mark(k, 0);
        const i64 lo = std::min(nJ, rank * cblk), hi = std::min(nJ, (rank + 1) * cblk);
        if (hi <= lo) return TCI_OK;
        cudaEventRecord(c->ev_g0, c->stream); // the side stream starts behind what the main stream holds so far
        cudaStreamWaitEvent(c->copy_stream, c->ev_g0, 0);
        StreamSwap side(c);
        DevBuf<i64> dJ(c);
        TCI_CUDA(c, dJ.upload(J + nr * lo, (size_t)(nr * (hi - lo))));
        double *env = nullptr;
        i64 Dk = 1;
        int r = t.kind == 1 ? env_eval_tt(c, t, 1, dJ.p, (int)nr, hi - lo, &env, &Dk)
                            : env_eval_mpo(c, t, 1, dJ.p, (int)nr, hi - lo, &env, &Dk, J + nr * lo);
        if (r) return r;
        cudaError_t e = cudaMemcpyAsync(renv[k] + D * lo, env, (size_t)D * (hi - lo) * sizeof(double),
                                        cudaMemcpyDeviceToDevice, c->stream);
        dev_free(c, env);
        TCI_CUDA(c, e);
        mark(k, 1);
        return TCI_OK;
    });
    // the all-gather runs on the members' copy streams while their main streams extend the left environments, which do
    // not need it; the block products wait for it
    if (!rc) rc = group_allgather_side(g, [&](int k) { return (void *)renv[k]; }, (size_t)D * cblk * sizeof(double));
    if (!rc) group_follow_owner(g); // dst was allocated in the owner's stream order
    std::vector<double *> lenvs(g->nlocal, nullptr);
    if (!rc)
        rc = group_run(g, [&](int k) -> int {
            tci_ctx *c = g->m[k];
            TargetDev &t = *c->targets.at(target_id);
            const int rank = c->rank;
            const i64 lo = std::min(nI, rank * rblk), hi = std::min(nI, (rank + 1) * rblk);
            if (hi <= lo) return TCI_OK;
            DevBuf<i64> dI(c);
            TCI_CUDA(c, dI.upload(I + nl * lo, (size_t)(nl * (hi - lo))));
            i64 Dk = 1;
            int r = t.kind == 1 ? env_eval_tt(c, t, 0, dI.p, (int)nl, hi - lo, &lenvs[k], &Dk)
                                : env_eval_mpo(c, t, 0, dI.p, (int)nl, hi - lo, &lenvs[k], &Dk, I + nl * lo);
            mark(k, 2);
            return r;
        });
    group_allgather_join(g);
    for (int k = 0; k < g->nlocal; ++k) {
        if (dbg) cudaSetDevice(g->m[k]->device);
        mark(k, 3);
    }
    if (dbg) cudaSetDevice(ctx->device);
    if (!rc)
        rc = group_run(g, [&](int k) -> int {
            tci_ctx *c = g->m[k];
            TargetDev &t = *c->targets.at(target_id);
            const int rank = c->rank;
            unsigned long long *w = rank == 0 ? d_maxbits : ctx_words(c);
            if (rank != 0) cudaMemsetAsync(w, 0, 8, c->stream);
            const i64 lo = std::min(nI, rank * rblk), hi = std::min(nI, (rank + 1) * rblk);
            if (hi <= lo) return TCI_OK;
            // block (sorted rows lo:hi) x (sorted columns) = lenv^T renv    cachedtensortrain.jl:211-212, contraction.jl:328
            const i64 rows = hi - lo;
            DevBuf<double> blk(c);
            DevBuf<i64> maps(c);
            TCI_CUDA(c, blk.alloc((size_t)(rows * nJ)));
            TCI_CUDA(c, maps.alloc((size_t)(rows + nJ)));
            TCI_CUDA(c, cudaMemcpyAsync(maps.p, rperm.data() + lo, (size_t)rows * sizeof(i64), cudaMemcpyHostToDevice, c->stream));
            TCI_CUDA(c, cudaMemcpyAsync(maps.p + rows, cperm.data(), (size_t)nJ * sizeof(i64), cudaMemcpyHostToDevice, c->stream));
            int r = dgemm_dev(c, true, false, rows, nJ, D, 1.0, lenvs[k], D, renv[k], D, 0.0, blk.p, rows);
            if (!r) r = apply_elementwise(c, t, blk.p, rows, nJ, rows);
            if (!r && d_maxbits) r = maxabs_dev(c, blk.p, rows, nJ, rows, w);
            if (!r) { // to the caller's rows / columns of the owner's Pi
                k_scatter_block<<<(unsigned)std::min<i64>((rows * nJ + 255) / 256, (i64)c->sm_count * 8), 256, 0, c->stream>>>(
                    blk.p, rows, nJ, maps.p, maps.p + rows, dst, ld);
                c->launches++;
                if (cudaGetLastError() != cudaSuccess) r = tci_fail(c, TCI_ERR_CUDA, "k_scatter_block launch failed");
            }
            mark(k, 4);
            return r;
        });
    if (dbg) {
        for (int k = 0; k < g->nlocal; ++k) {
            cudaSetDevice(g->m[k]->device);
            cudaStreamSynchronize(g->m[k]->stream);
            float a = 0, b = 0, c2 = 0, d2 = 0;
            cudaEventElapsedTime(&a, dev_ev[(size_t)k * 5], dev_ev[(size_t)k * 5 + 1]);
            cudaEventElapsedTime(&b, dev_ev[(size_t)k * 5 + 1], dev_ev[(size_t)k * 5 + 2]);
            cudaEventElapsedTime(&c2, dev_ev[(size_t)k * 5 + 2], dev_ev[(size_t)k * 5 + 3]);
            cudaEventElapsedTime(&d2, dev_ev[(size_t)k * 5 + 3], dev_ev[(size_t)k * 5 + 4]);
            fprintf(stderr, "[shard dbg] member %d: right chain %.2f ms | left chain %.2f | wait for all-gather %.2f | product %.2f\n",
                    k, a, b, c2, d2);
            for (int q = 0; q < 5; ++q) cudaEventDestroy(dev_ev[(size_t)k * 5 + q]);
        }
        cudaSetDevice(ctx->device);
    }
    for (int k = 0; k < g->nlocal; ++k) dev_free(g->m[k], lenvs[k]);
    for (int k = 0; k < g->nlocal; ++k) dev_free(g->m[k], renv[k]);
    if (rc) return rc;
    return group_allreduce_max_u64(g, [&](int k) { return g->m[k]->rank == 0 ? d_maxbits : ctx_words(g->m[k]); }, 1);
}

unsigned long long *ctx_words(tci_ctx *ctx)
{
    static std::mutex mu;
    static std::map<tci_ctx *, unsigned long long *> words; // freed with the process; 512 B per context
    std::lock_guard<std::mutex> lk(mu);
    auto it = words.find(ctx);
    if (it != words.end()) return it->second;
    unsigned long long *p = nullptr;
    if (cudaMalloc(&p, 64 * sizeof(unsigned long long)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, 64 * sizeof(unsigned long long));
    words[ctx] = p;
    return p;
}

// Evaluation into `out` (on the owner), sharded over the group when the cost model says so; nothing is synchronised.
// *d_maxbits: zeroed device word on the owner.
int pi_enqueue_auto(tci_ctx *ctx, i64 target_id, const i64 *I, i64 nl, i64 nI, const i64 *J, i64 nr, i64 nJ, i64 M,
                    tci_dmat *out, unsigned long long *d_maxbits, int *nshard_out)
{
    TargetDev &t = *ctx->targets.at(target_id);
    i64 C = 1;
    for (i64 k = 0; k < M; ++k) C *= t.localdims[nl + k];
    const int ns = choose_shards(ctx, t, nl, nI, nr, nJ, M, C);
    if (nshard_out) *nshard_out = ns;
    if (ns == 1) return pi_enqueue(ctx, t, I, nl, nI, J, nr, nJ, M, out, d_maxbits);
    // `out` is stream-ordered pool memory of the owner; the pool is mapped on every member (group.cu), so the
    // members' kernels store into it directly
    return t.kind == 0 || t.kind == 3
               ? pi_enqueue_sharded_analytic(ctx, target_id, ns, I, nl, nI, J, nr, nJ, M, out->p, out->ld, nI * C, d_maxbits)
                       : pi_enqueue_sharded_env(ctx, target_id, I, nl, nI, J, nr, nJ, out->p, out->ld, d_maxbits);
}

static int pi_eval_core(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                        int64_t nr, int64_t nJ, int64_t M, double *out_host, tci_dmat **out_dev, tci_dmat *dst,
                        int64_t col0, double *maxabs)
{
    if (out_dev) *out_dev = nullptr;
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &t = *it->second;
    if (t.is_complex) return tci_fail(ctx, TCI_ERR_ARG, "ComplexF64 target: use the tci_z* entry points");
    if (nI < 0 || nJ < 0 || nl < 0 || nr < 0 || M < 0) return tci_fail(ctx, TCI_ERR_ARG, "negative size");
    if (nI * nJ == 0) { // batcheval.jl:40-42: empty result, not an error
        if (maxabs) *maxabs = 0.0;
        if (out_dev) return dmat_alloc(ctx, 0, 0, out_dev);
        return TCI_OK;
    }
    if (nl + M + nr != t.nsites) return tci_fail(ctx, TCI_ERR_CENTRE, "Invalid number of central indices");
    if ((nl > 0 && !I) || (nr > 0 && !J)) return tci_fail(ctx, TCI_ERR_ARG, "index sets missing");
    i64 C = 1;
    for (i64 k = 0; k < M; ++k) C *= t.localdims[nl + k];
    if (dst && (dst->m != nI * C || col0 < 0 || col0 + nJ > dst->ncap))
        return tci_fail(ctx, TCI_ERR_ARG, "tci_pi_eval_into: destination block does not fit");

    unsigned long long *dmax = ctx_words(ctx);
    if (!dmax) return tci_fail(ctx, TCI_ERR_CUDA, "scratch words");
    cudaEventRecord(ctx->ev2, ctx->stream);
    cudaEventRecord(ctx->ev0, ctx->stream); // (re-recorded after the index upload when the owner evaluates itself)
    TCI_CUDA(ctx, cudaMemsetAsync(dmax, 0, 8, ctx->stream));
    tci_dmat *out = nullptr;
    tci_dmat view; // column block of dst
    int rc = 0;
    if (dst) {
        view = *dst;
        view.p = dst->p + dst->ld * col0;
        view.n = nJ;
        view.owned = false;
        out = &view;
    } else {
        rc = dmat_alloc(ctx, nI * C, nJ, &out);
        if (rc) return rc;
    }
    // an evaluation into a caller-provided block is never sharded (the caller is doing the sharding)
    rc = dst ? pi_enqueue(ctx, t, I, nl, nI, J, nr, nJ, M, out, dmax)
             : pi_enqueue_auto(ctx, target_id, I, nl, nI, J, nr, nJ, M, out, dmax, nullptr);
    if (rc) {
        if (!dst) tci_dmat_destroy(out);
        return rc;
    }
    {
        unsigned long long bits = 0;
        cudaEventRecord(ctx->ev1, ctx->stream); // evaluation done (kernel stage), copies follow
        if (maxabs)
            TCI_CUDA(ctx, cudaMemcpyAsync(&bits, dmax, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
        if (out_host)
            TCI_CUDA(ctx, cudaMemcpy2DAsync(out_host, out->m * sizeof(double), out->p, out->ld * sizeof(double),
                                            out->m * sizeof(double), out->n, cudaMemcpyDeviceToHost, ctx->stream));
        cudaEventRecord(ctx->ev3, ctx->stream);
        TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev2, ctx->ev0) == cudaSuccess) ctx->stage_ms[ST_H2D] += ms;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stage_ms[ST_PI] += ms;
        if (cudaEventElapsedTime(&ms, ctx->ev1, ctx->ev3) == cudaSuccess) ctx->stage_ms[ST_D2H] += ms;
        if (maxabs) {
            double v;
            memcpy(&v, &bits, sizeof(v));
            *maxabs = v;
        }
    }
    if (dst) return TCI_OK;
    if (out_dev)
        *out_dev = out;
    else
        tci_dmat_destroy(out);
    return TCI_OK;
}

extern "C" int tci_pi_eval(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI, const int64_t *J,
                           int64_t nr, int64_t nJ, int64_t M, double *out_host, tci_dmat **out_dev, double *maxabs)
{
    TCI_ENTER(ctx);
    return pi_eval_core(ctx, target_id, I, nl, nI, J, nr, nJ, M, out_host, out_dev, nullptr, 0, maxabs);
}

extern "C" int tci_pi_eval_into(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI,
                                const int64_t *J, int64_t nr, int64_t nJ, int64_t M, tci_dmat *dst, int64_t col0,
                                double *maxabs)
{
    TCI_ENTER(ctx);
    if (!dst) return tci_fail(ctx, TCI_ERR_ARG, "tci_pi_eval_into: dst missing");
    return pi_eval_core(ctx, target_id, I, nl, nI, J, nr, nJ, M, nullptr, nullptr, dst, col0, maxabs);
}

// ---- environments of TT / MPO-pair targets as objects of their own (sharded contraction) ----
static i64 env_dim_of(const TargetDev &t, int side, i64 len)
{
    if (len == 0) return 1;
    const i64 n = t.nsites;
    if (t.kind == 1) return side == 0 ? t.dr[len - 1] : t.dl[n - len];
    return side == 0 ? t.adr[len - 1] * t.bdr[len - 1] : t.adl[n - len] * t.bdl[n - len];
}

extern "C" int tci_env_dim(tci_ctx *ctx, int64_t target_id, int side, int64_t len, int64_t *D)
{
    TCI_ENTER(ctx);
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &t = *it->second;
    if (t.is_complex) return tci_fail(ctx, TCI_ERR_ARG, "ComplexF64 target: use the tci_z* entry points");
    if (t.kind == 0 || t.kind == 3 || t.kind == 4) return tci_fail(ctx, TCI_ERR_ARG, "tci_env_dim: an analytic target has no environments");
    if (!D || (side != 0 && side != 1) || len < 0 || len > t.nsites)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_env_dim: bad arguments");
    *D = env_dim_of(t, side, len);
    return TCI_OK;
}

extern "C" int tci_env_eval(tci_ctx *ctx, int64_t target_id, int side, const int64_t *idx, int64_t len, int64_t count,
                            tci_dmat *dst, int64_t col0)
{
    TCI_ENTER(ctx);
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &t = *it->second;
    if (t.is_complex) return tci_fail(ctx, TCI_ERR_ARG, "ComplexF64 target: use the tci_z* entry points");
    if (t.kind == 0 || t.kind == 3) return tci_fail(ctx, TCI_ERR_ARG, "tci_env_eval: an analytic target has no environments");
    if (!dst || (side != 0 && side != 1) || len < 0 || len > t.nsites || count < 0 || (len > 0 && count > 0 && !idx))
        return tci_fail(ctx, TCI_ERR_ARG, "tci_env_eval: bad arguments");
    if (dst->m != env_dim_of(t, side, len) || col0 < 0 || col0 + count > dst->ncap)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_env_eval: destination block does not fit");
    if (count == 0) return TCI_OK;
    dmat_wait_ready(ctx, dst);
    DevBuf<i64> d_idx(ctx);
    TCI_CUDA(ctx, d_idx.upload(idx, (size_t)(len * count)));
    StageTimer tm(ctx, ST_ENV);
    double *env = nullptr;
    i64 D = 1;
    int rc = t.kind == 1 ? env_eval_tt(ctx, t, side, d_idx.p, (int)len, count, &env, &D)
                         : env_eval_mpo(ctx, t, side, d_idx.p, (int)len, count, &env, &D, idx);
    if (rc) return rc;
    cudaError_t e = cudaMemcpy2DAsync(dst->p + dst->ld * col0, dst->ld * sizeof(double), env, D * sizeof(double),
                                      D * sizeof(double), count, cudaMemcpyDeviceToDevice, ctx->stream);
    dev_free(ctx, env);
    TCI_CUDA(ctx, e);
    tm.stop();
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

extern "C" int tci_pi_from_envs(tci_ctx *ctx, tci_dmat *left, int64_t l0, int64_t nI, tci_dmat *right, int64_t r0,
                                int64_t nJ, tci_dmat *dst, int64_t col0, double *maxabs)
{
    TCI_ENTER(ctx);
    if (!left || !right || !dst || nI < 0 || nJ < 0 || l0 < 0 || r0 < 0 || col0 < 0)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_pi_from_envs: bad arguments");
    if (left->m != right->m || l0 + nI > left->ncap || r0 + nJ > right->ncap || dst->m != nI || col0 + nJ > dst->ncap)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_pi_from_envs: shapes do not fit");
    if (maxabs) *maxabs = 0.0;
    if (nI * nJ == 0) return TCI_OK;
    DevBuf<unsigned long long> dmax(ctx);
    TCI_CUDA(ctx, dmax.alloc(2));
    TCI_CUDA(ctx, cudaMemsetAsync(dmax.p, 0, 16, ctx->stream));
    StageTimer tm(ctx, ST_PI);
    double *out = dst->p + dst->ld * col0;
    // Pi[i, j] = sum_a left[a, i] * right[a, j]     cachedtensortrain.jl:211-212, contraction.jl:328
    int rc = dgemm_dev(ctx, true, false, nI, nJ, left->m, 1.0, left->p + left->ld * l0, left->ld,
                       right->p + right->ld * r0, right->ld, 0.0, out, dst->ld);
    if (!rc && maxabs) rc = maxabs_dev(ctx, out, nI, nJ, dst->ld, dmax.p);
    if (rc) return rc;
    tm.stop();
    unsigned long long bits = 0;
    if (maxabs) TCI_CUDA(ctx, cudaMemcpyAsync(&bits, dmax.p, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (maxabs) memcpy(maxabs, &bits, sizeof(double));
    return TCI_OK;
}

// ---- scalar evaluation f(x) for a batch of full multi-indices ---------------
__global__ void k_eval_points(tci_analytic_t t, const i64 *__restrict__ idx, i64 count, double *__restrict__ out)
{
    i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    out[q] = tci_target_eval(&t, idx + (i64)t.nsites * q);
}

int target_eval_tt(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out);  // tt.cu
int target_eval_mpo(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out); // mpo.cu

int target_eval_dev(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out)
{
    if (count == 0) return TCI_OK;
    switch (t.kind) {
    case 0:
        k_eval_points<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(t.an, d_idx, count, d_out);
        ctx->launches++;
        TCI_CUDA(ctx, cudaGetLastError());
        return TCI_OK;
    case 1: return target_eval_tt(ctx, t, d_idx, count, d_out);
    case 3: return target_eval_user(ctx, t, d_idx, count, d_out);
    case 4: return target_eval_cached(ctx, t.cache_id, t, d_idx, count, d_out);
    default: {
        int rc = target_eval_mpo(ctx, t, d_idx, count, d_out);
        return rc ? rc : apply_elementwise(ctx, t, d_out, count, 1, count);
    }
    }
}

extern "C" int tci_target_eval(tci_ctx *ctx, int64_t target_id, const int64_t *idx, int64_t count, double *out)
{
    TCI_ENTER(ctx);
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    if (count <= 0) return TCI_OK;
    TargetDev &t = *it->second;
    if (t.is_complex) return tci_fail(ctx, TCI_ERR_ARG, "ComplexF64 target: use the tci_z* entry points");
    DevBuf<i64> d_idx(ctx);
    DevBuf<double> d_out(ctx);
    TCI_CUDA(ctx, d_idx.upload(idx, (size_t)(t.nsites * count)));
    TCI_CUDA(ctx, d_out.alloc((size_t)count));
    int rc = target_eval_dev(ctx, t, d_idx.p, count, d_out.p);
    if (rc) return rc;
    TCI_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}
