// bond.cu -- fused entry points of the TCI2 driver's inner loop.
//
// tci_bond_update   = the `:full` branch of updatepivots! (tensorci2.jl:529-551): Pi evaluation (sharded over the
//                     group's GPUs when the cost model says so) -> rrLU on the owner, Pi never leaves HBM, ONE host
//                     synchronisation for the whole bond.
// tci_fill_sitetensors = fillsitetensors! (globalsearch.jl:97-103; setsitetensor! tensorci2.jl:367-394) for ALL sites:
//                     Pi1, P, the full-rank factorisation of P and the solve T = Pi1 P^-1 of every site are queued back
//                     to back, one synchronisation at the end; the cores can stay on the device as a TT target (the
//                     `current_tt` the global pivot finder probes, globalpivotfinder.jl:160-183).
#include <chrono>
#include <cmath>
#include <cstring>
#include <unordered_set>

#include "tci_internal.h"

int pi_enqueue_auto(tci_ctx *ctx, i64 target_id, const i64 *I, i64 nl, i64 nI, const i64 *J, i64 nr, i64 nJ, i64 M,
                    tci_dmat *out, unsigned long long *d_maxbits, int *nshard_out);
int pi_enqueue(tci_ctx *ctx, TargetDev &t, const i64 *I, i64 nl, i64 nI, const i64 *J, i64 nr, i64 nJ, i64 M,
               tci_dmat *out, unsigned long long *d_maxbits);
unsigned long long *ctx_words(tci_ctx *ctx);
int rrlu_core(tci_ctx *ctx, tci_dmat *A, i64 m, i64 n, i64 maxrank, double reltol, double abstol, int leftorthogonal,
              int exact_mode, i64 *rowperm, i64 *colperm, i64 *npivot, double *error, double *pivoterrors,
              tci_lu **factors, const void *extra_dev, void *extra_host, size_t extra_bytes, int *deferred_result);
int lu_rdiv_enqueue(tci_lu *lu, const double *B, i64 ldb, i64 rows, double *X, i64 ldx);
int rrlu_batch_fullrank(tci_ctx *ctx, int nmat, tci_dmat *const *P, tci_lu **lus, char *const *results);
int lu_rdiv_batched_small(tci_ctx *ctx, int n, tci_lu *const *lus, const double *const *B, const i64 *ldb, const i64 *rows,
                          double *const *X, const i64 *ldx, std::vector<char> &done);

extern "C" int tci_bond_update(tci_ctx *ctx, int64_t target_id, const int64_t *I, int64_t nl, int64_t nI,
                               const int64_t *J, int64_t nr, int64_t nJ, int64_t maxrank, double reltol, double abstol,
                               int leftorthogonal, int exact_mode, int64_t *rowperm, int64_t *colperm,
                               int64_t *npivot, double *error, double *pivoterrors, double *maxabs, tci_lu **factors)
{
    TCI_ENTER(ctx);
    if (factors) *factors = nullptr;
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &t = *it->second;
    if (t.is_complex) return tci_fail(ctx, TCI_ERR_ARG, "ComplexF64 target: use the tci_z* entry points");
    if (!rowperm || !colperm || !npivot || !error) return tci_fail(ctx, TCI_ERR_ARG, "tci_bond_update: outputs missing");
    if (nI <= 0 || nJ <= 0) return tci_fail(ctx, TCI_ERR_ARG, "rows must not be empty"); // matrixlu.jl:10
    if (nI > 0x7ffffff0 || nJ > 0x7ffffff0) return tci_fail(ctx, TCI_ERR_ARG, "tci_bond_update: bad shape");
    if (nl + nr != t.nsites) return tci_fail(ctx, TCI_ERR_CENTRE, "Invalid number of central indices");
    if ((nl > 0 && !I) || (nr > 0 && !J)) return tci_fail(ctx, TCI_ERR_ARG, "index sets missing");
    unsigned long long *dmax = ctx_words(ctx);
    if (!dmax) return tci_fail(ctx, TCI_ERR_CUDA, "scratch words");
    static const bool bdbg = getenv("TCI_BOND_DEBUG") != nullptr; // host-side timeline of the call on stderr
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = bdbg ? now() : 0.0;
    tci_dmat *Pi = nullptr;
    int rc = dmat_alloc(ctx, nI, nJ, &Pi);
    if (rc) return rc;
    TCI_CUDA(ctx, cudaMemsetAsync(dmax, 0, 8, ctx->stream));
    cudaEvent_t e0 = ctx->ev4, e1 = ctx->ev5;
    cudaEventRecord(e0, ctx->stream);
    const bool saved = ctx->nosync;
    ctx->nosync = true; // nothing inside the evaluation may synchronise
    rc = pi_enqueue_auto(ctx, target_id, I, nl, nI, J, nr, nJ, 0, Pi, dmax, nullptr);
    ctx->nosync = saved;
    cudaEventRecord(e1, ctx->stream);
    const double t1 = bdbg ? now() : 0.0;
    unsigned long long bits = 0;
    if (!rc)
        rc = rrlu_core(ctx, Pi, nI, nJ, maxrank, reltol, abstol, leftorthogonal, exact_mode, rowperm, colperm, npivot,
                       error, pivoterrors, factors, dmax, &bits, sizeof(bits), nullptr);
    if (bdbg)
        fprintf(stderr, "[bond dbg] %lld x %lld r=%lld: Pi enqueue %.1f us | rrLU launch + sync + results %.1f us\n", (long long)nI,
                (long long)nJ, (long long)*npivot, t1 - t0, now() - t1);
    float ms = 0.f;
    if (!rc && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) ctx->stage_ms[ST_PI] += ms;
    if (maxabs) memcpy(maxabs, &bits, sizeof(double));
    if (rc || !(factors && *factors)) tci_dmat_destroy(Pi);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void k_compact(const double *__restrict__ src, i64 lds, i64 m, i64 n, double *__restrict__ dst)
{
    const i64 total = m * n;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x)
        dst[e] = src[(e % m) + lds * (e / m)];
}

extern "C" int tci_fill_sitetensors(tci_ctx *ctx, int64_t target_id, int64_t nsites, const int64_t *const *Iset,
                                    const int64_t *nI, const int64_t *const *Jset, const int64_t *nJ,
                                    double *const *T_out, double *maxabs, int64_t *tt_id)
{
    TCI_ENTER(ctx);
    if (tt_id) *tt_id = 0;
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &t = *it->second;
    if (t.is_complex) return tci_fail(ctx, TCI_ERR_ARG, "ComplexF64 target: use the tci_z* entry points");
    if (nsites != t.nsites || !Iset || !Jset || !nI || !nJ)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_fill_sitetensors: bad arguments");
    const i64 n = nsites;
    for (i64 b = 0; b + 1 < n; ++b)
        if (nI[b + 1] != nJ[b]) // tensorci2.jl:388
            return tci_fail(ctx, TCI_ERR_ARG, "Pivot matrix at bond " + std::to_string(b + 1) + " is not square!");
    for (i64 b = 0; b < n; ++b)
        if (nI[b] <= 0 || nJ[b] <= 0) return tci_fail(ctx, TCI_ERR_ARG, "tci_fill_sitetensors: empty index set");
    unsigned long long *dmax = ctx_words(ctx);
    if (!dmax) return tci_fail(ctx, TCI_ERR_CUDA, "scratch words");

    // page-locked staging: [per site: 8 result words + the pivot values][max bits][T_0 | T_1 | ...] -- results leave
    // in ONE synchronisation
    std::vector<size_t> toff((size_t)n + 1, 0), roff((size_t)n + 1, 0);
    for (i64 b = 0; b < n; ++b) {
        toff[b + 1] = toff[b] + (size_t)(nI[b] * t.localdims[b] * nJ[b]);
        roff[b + 1] = roff[b] + 32 + (size_t)nJ[b] * 8;
    }
    const size_t head = roff[n] + 64;
    const bool stage_T = T_out && toff[n] * sizeof(double) <= ((size_t)256 << 20);
    char *pin = static_cast<char *>(ctx_pinned(ctx, head + (stage_T ? toff[n] * sizeof(double) : 0)));
    if (!pin) return tci_fail(ctx, TCI_ERR_CUDA, "page-locked staging buffer");
    unsigned long long *hbits = reinterpret_cast<unsigned long long *>(pin + roff[n]);
    double *hT = reinterpret_cast<double *>(pin + head);

    std::unique_ptr<TargetDev> tt(new TargetDev());
    tt->kind = 1;
    tt->nsites = n;
    tt->pooled = true;
    std::vector<tci_lu *> lus;
    std::vector<tci_dmat *> tmp;
    auto cleanup = [&](bool keep_tt) {
        for (tci_lu *l : lus) { // (tci_lu_destroy enters the context: release by hand)
            dev_free(ctx, l->arena);
            ctx->live_handles--;
            dev_free(ctx, l->A->p);
            delete l->A;
            ctx->live_handles--;
            delete l;
        }
        for (tci_dmat *a : tmp) {
            dev_free(ctx, a->p);
            delete a;
            ctx->live_handles--;
        }
        if (!keep_tt) target_free(ctx, *tt);
    };
    const bool saved = ctx->nosync;
    ctx->nosync = true;
    cudaEventRecord(ctx->ev4, ctx->stream);
    static const bool fdbg = getenv("TCI_FILL_DEBUG") != nullptr; // phase timeline of this call on stderr
    cudaEvent_t fe[3] = {nullptr, nullptr, nullptr};
    if (fdbg)
        for (auto &e : fe) cudaEventCreate(&e);
    int rc = 0;
    cudaError_t ce = cudaMemsetAsync(dmax, 0, 8, ctx->stream);
    // phase 1: every Pi1 (with the running max|Pi1|, :372-375) and every P (:383-385), back to back
    std::vector<tci_dmat *> Pi1s((size_t)n, nullptr), Ps((size_t)n, nullptr);
    for (i64 b = 0; b < n && !rc && ce == cudaSuccess; ++b) {
        const i64 d = t.localdims[b], rows = nI[b] * d, k = nJ[b];
        double *core = nullptr;
        ce = dev_alloc(ctx, (void **)&core, (size_t)(rows * k) * sizeof(double));
        if (ce != cudaSuccess) break;
        tt->cores.push_back(core);
        tt->dl.push_back(nI[b]);
        tt->d.push_back(d);
        tt->dr.push_back(k);
        tt->localdims.push_back(d);
        rc = dmat_alloc(ctx, rows, k, &Pi1s[b]);
        if (rc) break;
        tmp.push_back(Pi1s[b]);
        rc = pi_enqueue(ctx, t, Iset[b], b, nI[b], Jset[b], n - 1 - b, k, 1, Pi1s[b], dmax);
        if (rc || b == n - 1) continue;
        rc = dmat_alloc(ctx, k, k, &Ps[b]);
        if (rc) break;
        rc = pi_enqueue(ctx, t, Iset[b + 1], b + 1, k, Jset[b], n - 1 - b, k, 0, Ps[b], nullptr);
    }
    if (fdbg) cudaEventRecord(fe[0], ctx->stream);
    // phase 2: the n-1 pivot matrices are independent: factorised to full rank in ONE cooperative launch, each by its
    // own group of CTAs (reltol = abstol = 0 never truncates; a singular P shows up in the result words / pivot values)
    std::vector<tci_lu *> lub((size_t)n, nullptr);
    if (!rc && ce == cudaSuccess && n > 1) {
        std::vector<char *> res((size_t)n - 1);
        for (i64 b = 0; b + 1 < n; ++b) res[b] = pin + roff[b];
        rc = rrlu_batch_fullrank(ctx, (int)(n - 1), Ps.data(), lub.data(), res.data());
    }
    for (i64 b = 0; b + 1 < n; ++b) { // a P is released through its factorisation handle, or directly
        if (lub[b])
            lus.push_back(lub[b]);
        else if (Ps[b])
            tmp.push_back(Ps[b]);
    }
    if (fdbg) cudaEventRecord(fe[1], ctx->stream);
    // phase 3: T = Pi1 P^-1 (:391) from the full-pivot factors; the last tensor is Pi1 itself (:377-381)
    std::vector<char> solved;
    if (!rc && ce == cudaSuccess && n > 1) { // all sites with a small pivot matrix: one launch
        std::vector<const double *> Bp((size_t)n - 1);
        std::vector<double *> Xp((size_t)n - 1);
        std::vector<i64> ldb((size_t)n - 1), rws((size_t)n - 1), ldx((size_t)n - 1);
        for (i64 b = 0; b + 1 < n; ++b) {
            Bp[b] = Pi1s[b] ? Pi1s[b]->p : nullptr;
            ldb[b] = Pi1s[b] ? Pi1s[b]->ld : 0;
            rws[b] = nI[b] * t.localdims[b];
            Xp[b] = tt->cores[b];
            ldx[b] = rws[b];
        }
        rc = lu_rdiv_batched_small(ctx, (int)(n - 1), lub.data(), Bp.data(), ldb.data(), rws.data(), Xp.data(), ldx.data(), solved);
    }
    for (i64 b = 0; b < n && !rc && ce == cudaSuccess; ++b) {
        const i64 d = t.localdims[b], rows = nI[b] * d, k = nJ[b];
        double *core = tt->cores[b];
        if (b + 1 < n && b < (i64)solved.size() && solved[b]) {
            // T_b was written by the batched kernel
        } else if (b == n - 1) {
            k_compact<<<(unsigned)std::min<i64>((rows * k + 255) / 256, 4096), 256, 0, ctx->stream>>>(Pi1s[b]->p, Pi1s[b]->ld,
                                                                                                  rows, k, core);
            ctx->launches++;
        } else
            rc = lu_rdiv_enqueue(lub[b], Pi1s[b]->p, Pi1s[b]->ld, rows, core, rows);
        if (!rc && T_out && T_out[b] && stage_T)
            ce = cudaMemcpyAsync(hT + toff[b], core, (size_t)(rows * k) * sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream);
    }
    ctx->nosync = saved;
    if (!rc && ce != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, std::string("tci_fill_sitetensors: ") + cudaGetErrorString(ce));
    if (!rc) {
        ce = cudaMemcpyAsync(hbits, dmax, 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(ctx->ev5, ctx->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
        if (ce != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, std::string("tci_fill_sitetensors: ") + cudaGetErrorString(ce));
        float ms = 0.f;
        if (!rc && cudaEventElapsedTime(&ms, ctx->ev4, ctx->ev5) == cudaSuccess) ctx->stage_ms[ST_LUCI] += ms;
        if (fdbg && !rc) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, ctx->ev4, fe[0]);
            cudaEventElapsedTime(&b, fe[0], fe[1]);
            cudaEventElapsedTime(&c, fe[1], ctx->ev5);
            fprintf(stderr, "[fill dbg] n=%lld: Pi1 / P evaluations %.3f ms | batched rrLU %.3f | solves + D2H %.3f | launches so far %lld\n",
                    (long long)n, a, b, c, (long long)ctx->launches);
        }
    } else
        cudaStreamSynchronize(ctx->stream);
    for (i64 b = 0; b + 1 < n && !rc; ++b) {
        const int *res = reinterpret_cast<const int *>(pin + roff[b]);
        const double *pv = reinterpret_cast<const double *>(pin + roff[b] + 32);
        bool singular = res[0] != (int)nJ[b] || (res[1] & 1) || (res[2] & 3);
        for (i64 q = 0; q < nJ[b] && !singular; ++q) singular = !(std::fabs(pv[q]) > 0.0); // an exact zero (or NaN) pivot
        if (singular) // LAPACK's `\` (tensorci2.jl:391) throws SingularException on an exactly singular P
            rc = tci_fail(ctx, TCI_ERR_SINGULAR, "Pivot matrix at bond " + std::to_string(b + 1) + " is singular!");
    }
    if (!rc && T_out)
        for (i64 b = 0; b < n && !rc; ++b) {
            if (!T_out[b]) continue;
            const size_t cnt = toff[b + 1] - toff[b];
            if (stage_T)
                memcpy(T_out[b], hT + toff[b], cnt * sizeof(double));
            else if (cudaMemcpy(T_out[b], tt->cores[b], cnt * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
                rc = tci_fail(ctx, TCI_ERR_CUDA, "tci_fill_sitetensors: D2H of a site tensor failed");
        }
    if (maxabs) memcpy(maxabs, hbits, sizeof(double));
    if (fdbg)
        for (auto &e : fe) cudaEventDestroy(e);
    const bool keep = !rc && tt_id != nullptr;
    cleanup(keep);
    if (rc) return rc;
    if (keep) {
        const i64 id = ctx->next_target++;
        ctx->targets[id] = std::move(tt);
        *tt_id = id;
        return target_replicate(ctx, id);
    }
    return TCI_OK;
}

// ---------------------------------------------------------------------------------------------------------
// tci_sweep2site_half: the bond loop of one half-sweep of sweep2site! (tensorci2.jl:866-907) -- updatepivots!
// (:510-607, `:full` search) for b = 1 .. n-1 (forward) or n-1 .. 1 (backward) -- in ONE call.  Between two bonds the
// reference only does index bookkeeping (kronecker, union with the history sets, the rows / columns the rrLU picked),
// which here runs in C++ next to the launches instead of in the caller's interpreter; every bond is still one
// evaluation -> rrLU -> synchronisation, because the next bond's index sets are this bond's pivots.
struct SweepResult {
    std::vector<std::vector<i64>> I, J; // flattened (len x count)
    std::vector<i64> nI, nJ;
    std::vector<double> bonderrors, pivoterrors;
    std::vector<i64> trace; // per bond: bond (1-based), rows, columns, npivot
    double maxsample = 0.0;
};
static std::map<tci_ctx *, SweepResult> g_sweeps;
static std::mutex g_sweeps_mu;

static void kron_left(const std::vector<i64> &S, i64 len, i64 cnt, i64 d, std::vector<i64> &out) // kronecker(Iset, d) :315-320
{
    out.resize((size_t)((len + 1) * cnt * d));
    for (i64 s = 0; s < d; ++s)
        for (i64 i = 0; i < cnt; ++i) {
            i64 *o = out.data() + (len + 1) * (i + cnt * s);
            for (i64 k = 0; k < len; ++k) o[k] = S[(size_t)(len * i + k)];
            o[len] = s + 1;
        }
}
static void kron_right(i64 d, const std::vector<i64> &S, i64 len, i64 cnt, std::vector<i64> &out) // kronecker(d, Jset) :322-327
{
    out.resize((size_t)((len + 1) * cnt * d));
    for (i64 j = 0; j < cnt; ++j)
        for (i64 s = 0; s < d; ++s) {
            i64 *o = out.data() + (len + 1) * (s + d * j);
            o[0] = s + 1;
            for (i64 k = 0; k < len; ++k) o[1 + k] = S[(size_t)(len * j + k)];
        }
}
// Base.union(a, b) on vectors of multi-indices of length len: order preserving, duplicates dropped (:526-527)
static void union_into(std::vector<i64> &a, i64 len, const i64 *b, i64 nb)
{
    if (nb == 0 || len == 0) return;
    const i64 na = (i64)a.size() / len;
    a.reserve((size_t)((na + nb) * len)); // rows are referenced by pointer below: no reallocation while inserting
    struct Hash {
        i64 len;
        size_t operator()(const i64 *p) const
        {
            unsigned long long h = 0x9e3779b97f4a7c15ull;
            for (i64 k = 0; k < len; ++k) {
                h ^= (unsigned long long)p[k] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
                h *= 0xbf58476d1ce4e5b9ull;
            }
            return (size_t)(h ^ (h >> 31));
        }
    };
    struct Eq {
        i64 len;
        bool operator()(const i64 *x, const i64 *y) const { return std::memcmp(x, y, (size_t)len * sizeof(i64)) == 0; }
    };
    std::unordered_set<const i64 *, Hash, Eq> seen((size_t)(2 * (na + nb)), Hash{len}, Eq{len});
    for (i64 i = 0; i < na; ++i) seen.insert(a.data() + len * i);
    for (i64 i = 0; i < nb; ++i) {
        const i64 *row = b + len * i;
        if (seen.find(row) != seen.end()) continue;
        a.insert(a.end(), row, row + len);
        seen.insert(a.data() + a.size() - len);
    }
}

void sweep_results_drop(tci_ctx *ctx) // ctx.cu: a context that goes away takes its unfetched result with it
{
    std::lock_guard<std::mutex> lk(g_sweeps_mu);
    g_sweeps.erase(ctx);
}

extern "C" int tci_sweep2site_half(tci_ctx *ctx, int64_t target_id, int forward, const int64_t *const *Iset,
                                   const int64_t *nI, const int64_t *const *Jset, const int64_t *nJ,
                                   const int64_t *const *extraI, const int64_t *nextraI, const int64_t *const *extraJ,
                                   const int64_t *nextraJ, double reltol, double abstol, int64_t maxbonddim, int exact_mode,
                                   int64_t *nI_out, int64_t *nJ_out, int64_t *npivoterrors)
{
    if (!ctx) return TCI_ERR_ARG;
    i64 n = 0;
    std::vector<i64> localdims;
    {
        TCI_ENTER(ctx);
        auto it = ctx->targets.find(target_id);
        if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
        if (it->second->is_complex) return tci_fail(ctx, TCI_ERR_ARG, "ComplexF64 target: use the tci_z* entry points");
        if (!Iset || !Jset || !nI || !nJ || !nI_out || !nJ_out) return tci_fail(ctx, TCI_ERR_ARG, "tci_sweep2site_half: bad arguments");
        n = it->second->nsites;
        localdims = it->second->localdims;
    }
    SweepResult res;
    res.I.resize((size_t)n);
    res.J.resize((size_t)n);
    for (i64 b = 0; b < n; ++b) {
        res.I[b].assign(Iset[b], Iset[b] + b * nI[b]);
        res.J[b].assign(Jset[b], Jset[b] + (n - 1 - b) * nJ[b]);
    }
    res.nI.assign(nI, nI + n);
    res.nJ.assign(nJ, nJ + n);
    res.bonderrors.assign((size_t)std::max<i64>(n - 1, 0), 0.0);
    std::vector<i64> Ic, Jc, rowperm, colperm;
    std::vector<double> pe;
    const i64 maxrank = maxbonddim <= 0 ? 0 : maxbonddim;
    for (i64 step = 0; step + 1 < n; ++step) {
        const i64 b = forward ? step : n - 2 - step;
        kron_left(res.I[b], b, res.nI[b], localdims[b], Ic);                       // Icombined, :526
        kron_right(localdims[b + 1], res.J[b + 1], n - 2 - b, res.nJ[b + 1], Jc); // Jcombined, :527
        if (extraI && nextraI && extraI[b + 1]) union_into(Ic, b + 1, extraI[b + 1], nextraI[b + 1]);
        if (extraJ && nextraJ && extraJ[b]) union_into(Jc, n - 1 - b, extraJ[b], nextraJ[b]);
        const i64 m = (i64)Ic.size() / (b + 1), k = (i64)Jc.size() / (n - 1 - b);
        rowperm.resize((size_t)std::max<i64>(m, 1));
        colperm.resize((size_t)std::max<i64>(k, 1));
        pe.assign((size_t)(std::min(m, k) + 1), 0.0);
        i64 np = 0;
        double err = 0.0, mx = 0.0;
        int rc = tci_bond_update(ctx, target_id, Ic.data(), b + 1, m, Jc.data(), n - 1 - b, k, maxrank, reltol, abstol,
                                 forward ? 1 : 0, exact_mode, rowperm.data(), colperm.data(), &np, &err, pe.data(), &mx,
                                 nullptr);
        if (rc) return rc;
        res.maxsample = (std::isnan(res.maxsample) || std::isnan(mx)) ? NAN : std::max(res.maxsample, mx); // :538
        std::vector<i64> &In = res.I[b + 1], &Jn = res.J[b];
        In.resize((size_t)((b + 1) * np)); // Iset[b+1] = Icombined[rowindices], Jset[b] = Jcombined[colindices]  :597-598
        for (i64 q = 0; q < np; ++q)
            std::copy(Ic.begin() + (b + 1) * (rowperm[q] - 1), Ic.begin() + (b + 1) * rowperm[q], In.begin() + (b + 1) * q);
        Jn.resize((size_t)((n - 1 - b) * np));
        for (i64 q = 0; q < np; ++q)
            std::copy(Jc.begin() + (n - 1 - b) * (colperm[q] - 1), Jc.begin() + (n - 1 - b) * colperm[q],
                      Jn.begin() + (n - 1 - b) * q);
        res.nI[b + 1] = np;
        res.nJ[b] = np;
        // updateerrors (:161-169): bonderrors[b] = last pivot error; pivoterrors = elementwise max with zero padding
        res.bonderrors[b] = pe[np];
        if ((i64)res.pivoterrors.size() < np + 1) res.pivoterrors.resize((size_t)(np + 1), 0.0);
        for (i64 q = 0; q <= np; ++q) {
            double &a = res.pivoterrors[q];
            a = (std::isnan(a) || std::isnan(pe[q])) ? NAN : std::max(a, pe[q]);
        }
        res.trace.insert(res.trace.end(), {b + 1, m, k, np});
    }
    for (i64 b = 0; b < n; ++b) {
        nI_out[b] = res.nI[b];
        nJ_out[b] = res.nJ[b];
    }
    if (npivoterrors) *npivoterrors = (i64)res.pivoterrors.size();
    std::lock_guard<std::mutex> lk(g_sweeps_mu);
    g_sweeps[ctx] = std::move(res);
    return TCI_OK;
}

// The results of the last tci_sweep2site_half on this context: index sets (Iout[b]: b x nI_out[b], Jout[b]:
// (n-1-b) x nJ_out[b]; nullable entries are skipped), bonderrors (n-1), pivoterrors (npivoterrors), max |Pi| over the
// bonds, and per bond (bond, rows, columns, npivot) in the order visited (trace: 4 (n-1), nullable).
extern "C" int tci_sweep2site_fetch(tci_ctx *ctx, int64_t *const *Iout, int64_t *const *Jout, double *bonderrors,
                                    double *pivoterrors, double *maxsample, int64_t *trace)
{
    if (!ctx) return TCI_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_sweeps_mu);
    auto it = g_sweeps.find(ctx);
    if (it == g_sweeps.end()) return tci_fail(ctx, TCI_ERR_ARG, "tci_sweep2site_fetch: no sweep result on this context");
    const SweepResult &r = it->second;
    const size_t n = r.I.size();
    for (size_t b = 0; b < n; ++b) {
        if (Iout && Iout[b]) std::copy(r.I[b].begin(), r.I[b].end(), Iout[b]);
        if (Jout && Jout[b]) std::copy(r.J[b].begin(), r.J[b].end(), Jout[b]);
    }
    if (bonderrors) std::copy(r.bonderrors.begin(), r.bonderrors.end(), bonderrors);
    if (pivoterrors) std::copy(r.pivoterrors.begin(), r.pivoterrors.end(), pivoterrors);
    if (maxsample) *maxsample = r.maxsample;
    if (trace) std::copy(r.trace.begin(), r.trace.end(), trace);
    g_sweeps.erase(it);
    return TCI_OK;
}
