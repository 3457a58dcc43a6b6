// ctx.cu -- context, device matrices, target registry, timers.
#include <cstring>

#include "tci_internal.h"

static thread_local std::string g_create_error;

int tci_fail(tci_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) {
        ctx->err = msg;
        ctx->failed = true;
    } else
        g_create_error = msg;
    return code;
}

extern "C" int tci_version(void) { return 100; }

int ctx_create_one(int device_id, tci_ctx **out)
{
    if (!out) return TCI_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return tci_fail(nullptr, TCI_ERR_NO_DEVICE,
                        std::string("no CUDA device available (libtci_b200 has no CPU fallback): ") +
                            cudaGetErrorString(e));
    if (device_id < 0 || device_id >= ndev) return tci_fail(nullptr, TCI_ERR_ARG, "invalid device id");
    tci_ctx *c = new tci_ctx();
    c->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess) {
        delete c;
        return tci_fail(nullptr, TCI_ERR_CUDA, "cudaSetDevice failed");
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device_id);
    c->sm_count = prop.multiProcessorCount;
    // the side stream (uploads that overlap with kernels; the right-environment chain + all-gather of a sharded
    // contraction Pi) has the highest priority: its thread blocks are scheduled first, the main stream's kernels fill
    // whatever it leaves idle
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        cudaEventCreate(&c->ev2) != cudaSuccess || cudaEventCreate(&c->ev3) != cudaSuccess ||
        cudaEventCreate(&c->ev4) != cudaSuccess || cudaEventCreate(&c->ev5) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_g0, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_g1, cudaEventDisableTiming) != cudaSuccess) {
        std::string m = cudaGetErrorString(cudaGetLastError());
        delete c;
        return tci_fail(nullptr, TCI_ERR_CUDA, "context setup failed: " + m);
    }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
        unsigned long long keep = ~0ull; // keep freed scratch in the pool
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = c;
    return TCI_OK;
}

int group_create_local(const std::vector<tci_ctx *> &members, tci_group **out); // group.cu
void sweep_results_drop(tci_ctx *ctx);                                             // bond.cu

// tci_ctx_create(ngpu, device_ids): ONE process drives ngpu GPUs (the reference's caller is a single Julia process,
// tensorci2.jl:805-809).  device_ids[0] owns the per-bond rrLU; the stages that shard are split over all
// of them inside the library.  device_ids == NULL means 0 .. ngpu-1.
extern "C" int tci_ctx_create(int ngpu, const int *device_ids, tci_ctx **out)
{
    if (!out) return TCI_ERR_ARG;
    *out = nullptr;
    if (ngpu < 1) return tci_fail(nullptr, TCI_ERR_ARG, "tci_ctx_create: ngpu must be >= 1");
    std::vector<tci_ctx *> members;
    for (int k = 0; k < ngpu; ++k) {
        const int dev = device_ids ? device_ids[k] : k;
        for (tci_ctx *m : members)
            if (m->device == dev) {
                for (tci_ctx *q : members) tci_ctx_destroy(q);
                return tci_fail(nullptr, TCI_ERR_ARG, "tci_ctx_create: device ids must be distinct");
            }
        tci_ctx *c = nullptr;
        int rc = ctx_create_one(dev, &c);
        if (rc) {
            for (tci_ctx *q : members) tci_ctx_destroy(q);
            return rc;
        }
        members.push_back(c);
    }
    if (ngpu > 1) {
        tci_group *g = nullptr;
        int rc = group_create_local(members, &g);
        if (rc) {
            for (tci_ctx *q : members) tci_ctx_destroy(q);
            return rc;
        }
    }
    cudaSetDevice(members[0]->device);
    *out = members[0];
    return TCI_OK;
}

void target_free(tci_ctx *ctx, TargetDev &t)
{
    if (t.kind == 4) {
        cache_target_free(ctx, t.cache_id);
        return;
    }
    if (t.pooled) {
        for (double *p : t.cores) dev_free(ctx, p);
        return;
    }
    user_target_unload(t);
    cudaFree(t.d_params);
    cudaFree(t.d_localdims);
    for (double *p : t.cores) cudaFree(p);
    for (double *p : t.A) cudaFree(p);
    for (double *p : t.B) cudaFree(p);
    for (double *p : t.Bp) cudaFree(p);
}

// frees a context that has been destroyed by its owner once the last dmat / lu handle that points at it is gone
// (Julia finalizers run in arbitrary order; Python drops a Context before the matrices that were created on it)
void ctx_release(tci_ctx *ctx)
{
    if (!ctx || !ctx->destroyed || ctx->live_handles > 0) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->targets) target_free(ctx, *kv.second);
    sweep_results_drop(ctx);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->ev2);
    cudaEventDestroy(ctx->ev3);
    cudaEventDestroy(ctx->ev4);
    cudaEventDestroy(ctx->ev5);
    cudaEventDestroy(ctx->ev_g0);
    cudaEventDestroy(ctx->ev_g1);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->shared_pool) cudaMemPoolDestroy(ctx->shared_pool);
    delete ctx;
}

extern "C" void tci_ctx_destroy(tci_ctx *ctx)
{
    if (!ctx || ctx->destroyed) return;
    if (ctx->grp) {
        tci_group *g = ctx->grp;
        ctx->grp = nullptr;
        if (ctx->member == 0) group_destroy(g);
    }
    ctx->destroyed = true;
    ctx_release(ctx);
}

// small device-to-host results go through page-locked staging (the copies are latency, not bandwidth)
void *ctx_pinned(tci_ctx *ctx, size_t bytes)
{
    if (bytes > ctx->pinned_cap) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        size_t cap = std::max(bytes, std::max((size_t)1 << 16, 2 * ctx->pinned_cap));
        if (cudaMallocHost(&ctx->pinned, cap) != cudaSuccess) {
            ctx->pinned_cap = 0;
            return nullptr;
        }
        ctx->pinned_cap = cap;
    }
    return ctx->pinned;
}

extern "C" const char *tci_last_error(tci_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int64_t tci_ctx_launches(tci_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int tci_ctx_ngpu(tci_ctx *ctx) { return ctx ? ctx_world(ctx) : 0; }

extern "C" int64_t tci_ctx_member_launches(tci_ctx *ctx, int k)
{
    if (!ctx) return 0;
    if (k == 0 || !ctx->grp) return k == 0 ? ctx->launches : 0;
    return k > 0 && k < ctx->grp->nlocal ? ctx->grp->m[k]->launches : 0;
}

extern "C" void *tci_ctx_stream(tci_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int tci_timers(tci_ctx *ctx, double *out, int64_t n, int reset)
{
    if (!ctx) return TCI_ERR_ARG;
    for (i64 q = 0; q < n && q < ST_COUNT; ++q) out[q] = ctx->stage_ms[q];
    if (reset)
        for (double &v : ctx->stage_ms) v = 0.0;
    return TCI_OK;
}

// ------------------------------------------------------------- dmat --------
int dmat_alloc(tci_ctx *ctx, i64 m, i64 n, tci_dmat **out)
{
    tci_dmat *a = new tci_dmat();
    a->ctx = ctx;
    a->m = m;
    a->n = n;
    a->ncap = n;
    a->ld = round_up(m > 0 ? m : 1, 16); // every column starts on a 128 B boundary
    size_t bytes = (size_t)a->ld * (size_t)(n > 0 ? n : 1) * sizeof(double);
    cudaError_t e = dev_alloc_shared(ctx, (void **)&a->p, bytes);
    if (e != cudaSuccess) {
        delete a;
        return tci_fail(ctx, TCI_ERR_CUDA, std::string("cudaMallocAsync dmat: ") + cudaGetErrorString(e));
    }
    ctx->live_handles++;
    *out = a;
    return TCI_OK;
}

extern "C" int tci_dmat_create(tci_ctx *ctx, int64_t m, int64_t n, const double *host, tci_dmat **out)
{
    TCI_ENTER(ctx);
    if (!out || m < 0 || n < 0) return tci_fail(ctx, TCI_ERR_ARG, "tci_dmat_create: bad arguments");
    tci_dmat *a = nullptr;
    int rc = dmat_alloc(ctx, m, n, &a);
    if (rc) return rc;
    if (host && m * n > 0) {
        StageTimer tm(ctx, ST_H2D);
        cudaError_t e = cudaMemcpy2DAsync(a->p, a->ld * sizeof(double), host, m * sizeof(double), m * sizeof(double),
                                          n, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            tci_dmat_destroy(a);
            return tci_fail(ctx, TCI_ERR_CUDA, std::string("H2D: ") + cudaGetErrorString(e));
        }
    }
    *out = a;
    return TCI_OK;
}

void dmat_wait_ready(tci_ctx *ctx, tci_dmat *a)
{
    if (a && a->ready) {
        cudaStreamWaitEvent(ctx->stream, a->ready, 0);
        cudaEventDestroy(a->ready); // released once the wait has been satisfied
        a->ready = nullptr;
    }
}

// Upload on the context's copy stream; returns as soon as the copy is enqueued.  `host` must stay valid (and
// should be pinned) until the matrix is first used; every consumer orders itself after the copy.
extern "C" int tci_dmat_create_async(tci_ctx *ctx, int64_t m, int64_t n, const double *host, tci_dmat **out)
{
    TCI_ENTER(ctx);
    if (!out || !host || m < 0 || n < 0) return tci_fail(ctx, TCI_ERR_ARG, "tci_dmat_create_async: bad arguments");
    tci_dmat *a = nullptr;
    int rc = dmat_alloc(ctx, m, n, &a);
    if (rc) return rc;
    if (m * n > 0) {
        cudaEvent_t alloc_done = nullptr;
        cudaError_t e = cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&a->ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(alloc_done, ctx->stream); // the allocation is stream ordered
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_stream, alloc_done, 0);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(a->p, a->ld * sizeof(double), host, m * sizeof(double), m * sizeof(double), n,
                                  cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(a->ready, ctx->copy_stream);
        if (alloc_done) cudaEventDestroy(alloc_done);
        if (e != cudaSuccess) {
            tci_dmat_destroy(a);
            return tci_fail(ctx, TCI_ERR_CUDA, std::string("async H2D: ") + cudaGetErrorString(e));
        }
    }
    *out = a;
    return TCI_OK;
}

extern "C" int tci_dmat_shape(tci_dmat *a, int64_t *m, int64_t *n, int64_t *ld)
{
    if (!a) return TCI_ERR_ARG;
    if (m) *m = a->m;
    if (n) *n = a->n;
    if (ld) *ld = a->ld;
    return TCI_OK;
}

extern "C" int tci_dmat_resize_cols(tci_dmat *a, int64_t n)
{
    if (!a || n < 0 || n > a->ncap) return TCI_ERR_ARG;
    a->n = n;
    return TCI_OK;
}

extern "C" int tci_dmat_wrap(tci_ctx *ctx, void *dptr, int64_t m, int64_t n, int64_t ld, tci_dmat **out)
{
    if (!ctx || !out || !dptr || m < 0 || n < 0 || ld < m || (ld & 1) || ((size_t)dptr & 15))
        return tci_fail(ctx, TCI_ERR_ARG, "tci_dmat_wrap: bad arguments");
    tci_dmat *a = new tci_dmat();
    a->ctx = ctx;
    a->p = static_cast<double *>(dptr);
    a->m = m;
    a->n = n;
    a->ncap = n;
    a->ld = ld;
    a->owned = false;
    ctx->live_handles++;
    *out = a;
    return TCI_OK;
}

// the same column-major data under another fold (m * n == m2 * n2): reshape(A, m2, n2) of a Julia array
__global__ void k_refold(const double *__restrict__ src, i64 lds, i64 m, i64 total, double *__restrict__ dst, i64 ldd,
                         i64 m2)
{
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x)
        dst[(e % m2) + ldd * (e / m2)] = src[(e % m) + lds * (e / m)];
}

extern "C" int tci_dmat_refold(tci_dmat *a, int64_t m2, int64_t n2, tci_dmat **out)
{
    if (!a || !out) return TCI_ERR_ARG;
    tci_ctx *ctx = a->ctx;
    TCI_ENTER(ctx);
    *out = nullptr;
    if (m2 < 0 || n2 < 0 || m2 * n2 != a->m * a->n)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_dmat_refold: DimensionMismatch: new dimensions must be consistent with array size");
    dmat_wait_ready(ctx, a);
    tci_dmat *r = nullptr;
    int rc = dmat_alloc(ctx, m2, n2, &r);
    if (rc) return rc;
    const i64 total = m2 * n2;
    if (total > 0) {
        k_refold<<<(unsigned)std::min<i64>((total + 255) / 256, (i64)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
            a->p, a->ld, a->m, total, r->p, r->ld, m2);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            tci_dmat_destroy(r);
            return tci_fail(ctx, TCI_ERR_CUDA, std::string("tci_dmat_refold: ") + cudaGetErrorString(e));
        }
    }
    *out = r;
    return TCI_OK;
}

extern "C" void *tci_dmat_ptr(tci_dmat *a) { return a ? a->p : nullptr; }

extern "C" int tci_dmat_fetch(tci_dmat *a, double *host)
{
    if (!a || !host) return TCI_ERR_ARG;
    tci_ctx *ctx = a->ctx;
    TCI_ENTER(ctx);
    if (a->m * a->n == 0) return TCI_OK;
    dmat_wait_ready(ctx, a);
    StageTimer tm(ctx, ST_D2H);
    TCI_CUDA(ctx, cudaMemcpy2DAsync(host, a->m * sizeof(double), a->p, a->ld * sizeof(double), a->m * sizeof(double),
                                    a->n, cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

extern "C" int tci_dmat_destroy(tci_dmat *a)
{
    if (!a) return TCI_OK;
    tci_ctx *ctx = a->ctx;
    cudaSetDevice(ctx->device);
    dmat_wait_ready(ctx, a);
    if (a->owned) dev_free(ctx, a->p);
    delete a;
    ctx->live_handles--;
    ctx_release(ctx); // no-op unless the context was destroyed before its handles
    return TCI_OK;
}

// ----------------------------------------------------------- targets -------
extern "C" int tci_target_builtin(tci_ctx *ctx, int kind_id, const double *params, int64_t nparams,
                                  const int64_t *localdims, int64_t nsites, int64_t *target_id)
{
    TCI_ENTER(ctx);
    if (!target_id || nsites < 1 || !localdims || (nparams > 0 && !params))
        return tci_fail(ctx, TCI_ERR_ARG, "tci_target_builtin: bad arguments");
    if (kind_id < TCI_TARGET_LORENTZ || kind_id > TCI_TARGET_GKCOSEXP)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_target_builtin: unknown kind");
    std::unique_ptr<TargetDev> t(new TargetDev());
    t->kind = 0;
    t->nsites = nsites;
    t->localdims.assign(localdims, localdims + nsites);
    std::vector<double> p;
    if (kind_id == TCI_TARGET_TABLE) { // strides in front of the table
        double st = 1.0;
        for (i64 k = 0; k < nsites; ++k) {
            p.push_back(st);
            st *= (double)localdims[k];
        }
        if ((double)nparams != st) return tci_fail(ctx, TCI_ERR_ARG, "table size must equal prod(localdims)");
    }
    p.insert(p.end(), params, params + nparams);
    if (p.empty()) p.push_back(0.0);
    if (kind_id == TCI_TARGET_SEPCOS) {
        i64 nt = (i64)p[0];
        if (nt < 0 || 1 + nt > TCI_MAX_STATE || (i64)p.size() < 1 + nsites + nt + nt * nsites)
            return tci_fail(ctx, TCI_ERR_ARG, "sepcos: parameter blob has the wrong size");
    }
    if (kind_id == TCI_TARGET_GKCOSEXP) {
        i64 q = (i64)p[0];
        if ((i64)p.size() < 1 + 2 * q) return tci_fail(ctx, TCI_ERR_ARG, "gkcosexp: parameter blob too short");
    }
    if ((kind_id == TCI_TARGET_QUANTICS2D || kind_id == TCI_TARGET_QUANTICS1D) && p.size() < 2)
        return tci_fail(ctx, TCI_ERR_ARG, "quantics: parameters are [layout|R, R|fid]");
    t->nparams_alloc = (i64)p.size();
    TCI_CUDA(ctx, cudaMalloc(&t->d_params, p.size() * sizeof(double)));
    TCI_CUDA(ctx, cudaMemcpy(t->d_params, p.data(), p.size() * sizeof(double), cudaMemcpyHostToDevice));
    TCI_CUDA(ctx, cudaMalloc(&t->d_localdims, nsites * sizeof(i64)));
    TCI_CUDA(ctx, cudaMemcpy(t->d_localdims, localdims, nsites * sizeof(i64), cudaMemcpyHostToDevice));
    t->an.kind = kind_id;
    t->an.nsites = (int)nsites;
    t->an.nparams = (i64)(kind_id == TCI_TARGET_LORENTZ ? nparams : (i64)p.size());
    t->an.params = t->d_params;
    t->an.localdims = t->d_localdims;
    t->an.nstate = tci_target_nstate(kind_id, p.data());
    i64 id = ctx->next_target++;
    ctx->targets[id] = std::move(t);
    *target_id = id;
    return target_replicate(ctx, id);
}

extern "C" int tci_tt_create(tci_ctx *ctx, int64_t nsites, const int64_t *dims3, const double *const *cores,
                             int64_t *target_id)
{
    TCI_ENTER(ctx);
    if (!target_id || nsites < 1 || !dims3 || !cores) return tci_fail(ctx, TCI_ERR_ARG, "tci_tt_create: bad arguments");
    std::unique_ptr<TargetDev> t(new TargetDev());
    t->kind = 1;
    t->nsites = nsites;
    for (i64 s = 0; s < nsites; ++s) {
        i64 Dl = dims3[3 * s], d = dims3[3 * s + 1], Dr = dims3[3 * s + 2];
        if (Dl < 1 || d < 1 || Dr < 1 || (s > 0 && Dl != dims3[3 * s - 1]))
            return tci_fail(ctx, TCI_ERR_ARG,
                            "The tensors must have consistent dimensions for a tensor train."); // tensortrain.jl:22-26
        t->dl.push_back(Dl);
        t->d.push_back(d);
        t->dr.push_back(Dr);
        t->localdims.push_back(d);
        double *p = nullptr;
        TCI_CUDA(ctx, cudaMalloc(&p, Dl * d * Dr * sizeof(double)));
        t->cores.push_back(p);
        TCI_CUDA(ctx, cudaMemcpy(p, cores[s], Dl * d * Dr * sizeof(double), cudaMemcpyHostToDevice));
    }
    i64 id = ctx->next_target++;
    ctx->targets[id] = std::move(t);
    *target_id = id;
    return target_replicate(ctx, id);
}

// One core of a device-resident tensor train (tci_tt_create, or the handle tci_fill_sitetensors returns): the site
// tensors stay in HBM between fillsitetensors! and the global pivot search and are fetched only when the host reads them.
extern "C" int tci_tt_fetch_core(tci_ctx *ctx, int64_t tt_id, int64_t site, int64_t *dims3, double *out)
{
    TCI_ENTER(ctx);
    auto it = ctx->targets.find(tt_id);
    if (it == ctx->targets.end() || it->second->kind != 1 || it->second->is_complex)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_tt_fetch_core: not a tensor-train handle");
    const TargetDev &t = *it->second;
    if (site < 0 || site >= t.nsites) return tci_fail(ctx, TCI_ERR_ARG, "tci_tt_fetch_core: site out of range");
    if (dims3) {
        dims3[0] = t.dl[site];
        dims3[1] = t.d[site];
        dims3[2] = t.dr[site];
    }
    if (!out) return TCI_OK;
    StageTimer tm(ctx, ST_D2H);
    TCI_CUDA(ctx, cudaMemcpyAsync(out, t.cores[site], (size_t)(t.dl[site] * t.d[site] * t.dr[site]) * sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

extern "C" int tci_mpo_pair_create(tci_ctx *ctx, int64_t nsites, const int64_t *dimsA4, const double *const *A,
                                   const int64_t *dimsB4, const double *const *B, int64_t *target_id)
{
    TCI_ENTER(ctx);
    if (!target_id || nsites < 1 || !dimsA4 || !dimsB4 || !A || !B)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_mpo_pair_create: bad arguments");
    std::unique_ptr<TargetDev> t(new TargetDev());
    t->kind = 2;
    t->nsites = nsites;
    for (i64 s = 0; s < nsites; ++s) {
        const i64 *da = dimsA4 + 4 * s, *db = dimsB4 + 4 * s;
        if (da[2] != db[1])
            return tci_fail(ctx, TCI_ERR_ARG, "Tensor trains must share the identical index at n=" +
                                                  std::to_string(s + 1) + "!"); // contraction.jl:44-48
        t->adl.push_back(da[0]);
        t->as1.push_back(da[1]);
        t->as2.push_back(da[2]);
        t->adr.push_back(da[3]);
        t->bdl.push_back(db[0]);
        t->bs1.push_back(db[1]);
        t->bs2.push_back(db[2]);
        t->bdr.push_back(db[3]);
        t->localdims.push_back(da[1] * db[2]);
        i64 na = da[0] * da[1] * da[2] * da[3], nb = db[0] * db[1] * db[2] * db[3];
        double *pa = nullptr, *pb = nullptr;
        TCI_CUDA(ctx, cudaMalloc(&pa, na * sizeof(double)));
        t->A.push_back(pa);
        TCI_CUDA(ctx, cudaMalloc(&pb, nb * sizeof(double)));
        t->B.push_back(pb);
        TCI_CUDA(ctx, cudaMemcpy(pa, A[s], na * sizeof(double), cudaMemcpyHostToDevice));
        TCI_CUDA(ctx, cudaMemcpy(pb, B[s], nb * sizeof(double), cudaMemcpyHostToDevice));
    }
    i64 id = ctx->next_target++;
    ctx->targets[id] = std::move(t);
    *target_id = id;
    return target_replicate(ctx, id);
}

extern "C" int tci_target_destroy(tci_ctx *ctx, int64_t target_id)
{
    TCI_ENTER(ctx);
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    target_free(ctx, *it->second);
    ctx->targets.erase(it);
    if (ctx->grp)
        for (int k = 1; k < ctx->grp->nlocal; ++k) {
            tci_ctx *c = ctx->grp->m[k];
            auto jt = c->targets.find(target_id);
            if (jt == c->targets.end()) continue;
            cudaSetDevice(c->device);
            cudaStreamSynchronize(c->stream);
            target_free(c, *jt->second);
            c->targets.erase(jt);
        }
    cudaSetDevice(ctx->device);
    return TCI_OK;
}

static cudaError_t peer_dup(double **dst, int dst_dev, const double *src, int src_dev, size_t count)
{
    cudaError_t e = cudaMalloc(dst, (count ? count : 1) * sizeof(double));
    if (e != cudaSuccess || count == 0) return e;
    return cudaMemcpyPeer(*dst, dst_dev, src, src_dev, count * sizeof(double));
}

int target_replicate(tci_ctx *ctx, i64 id)
{
    tci_group *g = ctx->grp;
    if (!g || g->nlocal == 1) return TCI_OK;
    const TargetDev &src = *ctx->targets.at(id);
    cudaStreamSynchronize(ctx->stream);
    for (int k = 1; k < g->nlocal; ++k) {
        tci_ctx *c = g->m[k];
        cudaSetDevice(c->device);
        std::unique_ptr<TargetDev> t(new TargetDev(src)); // sizes, dims, elementwise function; pointers re-made below
        t->pooled = false;
        t->d_params = nullptr;
        t->d_localdims = nullptr;
        std::fill(t->cores.begin(), t->cores.end(), nullptr);
        std::fill(t->A.begin(), t->A.end(), nullptr);
        std::fill(t->B.begin(), t->B.end(), nullptr);
        t->Bp.clear(); // made again on this device on first use
        cudaError_t e = cudaSuccess;
        t->user_lib = t->user_pi = t->user_points = nullptr;
        if (src.kind == 0 || src.kind == 3) {
            e = peer_dup(&t->d_params, c->device, src.d_params, ctx->device, (size_t)src.nparams_alloc);
            if (e == cudaSuccess) {
                double *ld = nullptr;
                static_assert(sizeof(i64) == sizeof(double), "localdims are copied as 8-byte words");
                e = peer_dup(&ld, c->device, reinterpret_cast<const double *>(src.d_localdims), ctx->device,
                             (size_t)src.nsites);
                t->d_localdims = reinterpret_cast<i64 *>(ld);
            }
            t->an.params = t->d_params;
            t->an.localdims = t->d_localdims;
            if (e == cudaSuccess && src.kind == 3 && user_target_load(c, *t) != TCI_OK) e = cudaErrorUnknown;
        } else if (src.kind == 1) {
            for (i64 s = 0; s < src.nsites && e == cudaSuccess; ++s)
                e = peer_dup(&t->cores[s], c->device, src.cores[s], ctx->device,
                             (size_t)(src.dl[s] * src.d[s] * src.dr[s]));
        } else {
            const size_t w = src.is_complex ? 2 : 1; // doubles per element
            for (i64 s = 0; s < src.nsites && e == cudaSuccess; ++s) {
                e = peer_dup(&t->A[s], c->device, src.A[s], ctx->device,
                             w * (size_t)(src.adl[s] * src.as1[s] * src.as2[s] * src.adr[s]));
                if (e == cudaSuccess)
                    e = peer_dup(&t->B[s], c->device, src.B[s], ctx->device,
                                 w * (size_t)(src.bdl[s] * src.bs1[s] * src.bs2[s] * src.bdr[s]));
            }
        }
        if (e != cudaSuccess) {
            target_free(c, *t);
            cudaSetDevice(ctx->device);
            return tci_fail(ctx, TCI_ERR_CUDA, std::string("replicating the target on GPU ") +
                                                   std::to_string(c->device) + ": " + cudaGetErrorString(e));
        }
        c->targets[id] = std::move(t);
        if (c->next_target <= id) c->next_target = id + 1;
    }
    cudaSetDevice(ctx->device);
    return TCI_OK;
}
