// bond.cu -- fused entry points of the TCI2 driver's inner loop.
//
// tci_bond_update   = the `:full` branch of updatepivots! (tensorci2.jl:529-551): Pi evaluation (sharded over the
//                     group's GPUs when the cost model says so) -> rrLU on the owner, Pi never leaves HBM, ONE host
//                     synchronisation for the whole bond.
// tci_fill_sitetensors = fillsitetensors! (globalsearch.jl:97-103; setsitetensor! tensorci2.jl:367-394) for ALL sites:
//                     Pi1, P, the full-rank factorisation of P and the solve T = Pi1 P^-1 of every site are queued back
//                     to back, one synchronisation at the end; the cores can stay on the device as a TT target (the
//                     `current_tt` the global pivot finder probes, globalpivotfinder.jl:160-183).
#include <cmath>
#include <cstring>

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
    unsigned long long bits = 0;
    if (!rc)
        rc = rrlu_core(ctx, Pi, nI, nJ, maxrank, reltol, abstol, leftorthogonal, exact_mode, rowperm, colperm, npivot,
                       error, pivoterrors, factors, dmax, &bits, sizeof(bits), nullptr);
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
    for (i64 b = 0; b < n && !rc && ce == cudaSuccess; ++b) {
        const i64 d = t.localdims[b], rows = nI[b] * d, k = nJ[b];
        double *core = tt->cores[b];
        if (b == n - 1) {
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
