// ctx.cu -- context, device matrices, target registry, timers.
#include <cstring>

#include "tci_internal.h"

static thread_local std::string g_create_error;

int tci_fail(tci_ctx *ctx, int code, const std::string &msg)
{
    if (ctx)
        ctx->err = msg;
    else
        g_create_error = msg;
    return code;
}

extern "C" int tci_version(void) { return 100; }

extern "C" int tci_ctx_create(int device_id, tci_ctx **out)
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
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
        cudaEventCreate(&c->ev2) != cudaSuccess || cudaEventCreate(&c->ev3) != cudaSuccess) {
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

static void target_free(TargetDev &t)
{
    cudaFree(t.d_params);
    cudaFree(t.d_localdims);
    for (double *p : t.cores) cudaFree(p);
    for (double *p : t.A) cudaFree(p);
    for (double *p : t.B) cudaFree(p);
}

extern "C" void tci_ctx_destroy(tci_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->targets) target_free(*kv.second);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->ev2);
    cudaEventDestroy(ctx->ev3);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *tci_last_error(tci_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int64_t tci_ctx_launches(tci_ctx *ctx) { return ctx ? ctx->launches : 0; }

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
    cudaError_t e = dev_alloc(ctx, (void **)&a->p, bytes);
    if (e != cudaSuccess) {
        delete a;
        return tci_fail(ctx, TCI_ERR_CUDA, std::string("cudaMallocAsync dmat: ") + cudaGetErrorString(e));
    }
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
    *out = a;
    return TCI_OK;
}

extern "C" int tci_shared_alloc(tci_ctx *ctx, int64_t bytes, void **dptr, char handle[64])
{
    TCI_ENTER(ctx);
    if (!dptr || !handle || bytes <= 0) return tci_fail(ctx, TCI_ERR_ARG, "tci_shared_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    TCI_CUDA(ctx, cudaMalloc(dptr, (size_t)bytes)); // pool (async) allocations cannot be exported
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *dptr);
    if (e != cudaSuccess) {
        cudaFree(*dptr);
        *dptr = nullptr;
        return tci_fail(ctx, TCI_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle, &h, 64);
    return TCI_OK;
}

extern "C" int tci_shared_open(tci_ctx *ctx, const char handle[64], void **dptr)
{
    TCI_ENTER(ctx);
    if (!dptr || !handle) return tci_fail(ctx, TCI_ERR_ARG, "tci_shared_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    TCI_CUDA(ctx, cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TCI_OK;
}

extern "C" int tci_shared_close(tci_ctx *ctx, void *dptr)
{
    TCI_ENTER(ctx);
    if (dptr) TCI_CUDA(ctx, cudaIpcCloseMemHandle(dptr));
    return TCI_OK;
}

extern "C" int tci_shared_free(tci_ctx *ctx, void *dptr)
{
    TCI_ENTER(ctx);
    if (dptr) {
        cudaStreamSynchronize(ctx->stream);
        TCI_CUDA(ctx, cudaFree(dptr));
    }
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
    cudaSetDevice(a->ctx->device);
    dmat_wait_ready(a->ctx, a);
    if (a->owned) dev_free(a->ctx, a->p);
    delete a;
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
    return TCI_OK;
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
    return TCI_OK;
}

extern "C" int tci_target_destroy(tci_ctx *ctx, int64_t target_id)
{
    TCI_ENTER(ctx);
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    target_free(*it->second);
    ctx->targets.erase(it);
    return TCI_OK;
}
