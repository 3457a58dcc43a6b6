// tci_internal.h -- shared declarations of libtci_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tci_b200.h"
#include "../../include/tci_targets.h"

typedef int64_t i64;

enum { ST_PI = 0, ST_RRLU = 1, ST_LUCI = 2, ST_ENV = 3, ST_GSEARCH = 4, ST_GEMM = 5, ST_H2D = 6, ST_D2H = 7, ST_RRLU_KERNEL = 8, ST_COUNT = 9 };

struct tci_dmat {
    tci_ctx *ctx = nullptr;
    double *p = nullptr;
    i64 m = 0, n = 0, ld = 0;
    i64 ncap = 0; // allocated columns
    bool owned = true;
    cudaEvent_t ready = nullptr; // set by tci_dmat_create_async: the upload on the copy stream has finished
};
// make ctx->stream wait for a pending asynchronous upload of `a` (no-op otherwise)
void dmat_wait_ready(tci_ctx *ctx, tci_dmat *a);

struct TargetDev {
    int kind = 0; // 0 analytic, 1 TT, 2 MPO pair, 3 user source (NVRTC), 4 memoised wrapper of another target (cache.cu)
    i64 cache_id = 0; // kind 4: this target's own id, the key of its table
    i64 nsites = 0;
    std::vector<i64> localdims;
    // analytic
    tci_analytic_t an{};     // params/localdims are DEVICE pointers
    double *d_params = nullptr;
    i64 *d_localdims = nullptr;
    i64 nparams_alloc = 0; // doubles behind d_params
    bool is_complex = false; // ComplexF64 MPO pair: A / B hold interleaved (re, im) pairs (zpath.cu)
    bool pooled = false;   // device buffers come from the context's stream-ordered pool (tci_fill_sitetensors)
    // TT: cores on device, dims (Dl, d, Dr)
    std::vector<double *> cores;
    std::vector<i64> dl, d, dr;
    // MPO pair
    std::vector<double *> A, B;
    // B permuted to (Lb, Lbn, S, d3) per site: the right-environment chain contracts (br, h) as ONE inner dimension
    // (made on first use, mpo.cu)
    std::vector<double *> Bp;
    std::vector<i64> adl, as1, as2, adr, bdl, bs1, bs2, bdr;
    // user source: the compiled module (kept so that it can be loaded on every GPU of a group) and its kernels
    std::vector<char> cubin;
    void *user_lib = nullptr, *user_pi = nullptr, *user_points = nullptr;
    // elementwise function applied to the product (Contraction.f, contraction.jl:330-332): TCI_F_* id + parameters
    int fkind = 0;
    double fa = 1.0, fb = 0.0;
};
// applies the target's elementwise function to an m x n block (no-op for fkind == 0)   pi_eval.cu
int apply_elementwise(tci_ctx *ctx, const TargetDev &t, double *p, i64 m, i64 n, i64 ld);

struct tci_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr; // uploads that overlap with kernels on `stream` (tci_dmat_create_async)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    cudaEvent_t ev4 = nullptr, ev5 = nullptr; // stage boundaries inside the fused entry points (bond.cu)
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr; // ordering between `stream` and a collective on `copy_stream`
    std::string err;
    std::mutex mu;
    bool busy = false;
    i64 launches = 0;
    double stage_ms[ST_COUNT] = {0};
    // relative HBM streaming speed of the SMs in %smid order, learned from the rrLU streaming passes (rrlu.cu)
    double sm_speed[256];
    bool sm_speed_valid = false;
    std::map<i64, std::unique_ptr<TargetDev>> targets;
    i64 next_target = 1;
    // kernels whose dynamic shared-memory limit has been raised ON THIS CONTEXT'S DEVICE (the attribute is per device)
    std::set<const void *> smem_configured;
    // multi-GPU: the group this context belongs to (nullptr: a plain single-GPU context) and its place in it
    struct tci_group *grp = nullptr;
    int member = 0;      // index among the group's LOCAL members (0 = the context the caller holds)
    int rank = 0;        // global rank in the group (0 = owner of the per-bond rrLU)
    bool nosync = false; // inside a batched entry point: stage timers must not synchronise the stream
    // group members: an explicit stream-ordered pool mapped on every other member (Pi buffers that peers store into);
    // the device's default pool stays private -- growing two peer-mapped default pools from two threads at once
    // failed with cudaErrorMemoryAllocation on the 2-GPU box with 189 GB free
    cudaMemPool_t shared_pool = nullptr;
    int live_handles = 0; // dmat / lu handles that still point at this context (tci_ctx_destroy defers to the last)
    bool destroyed = false;
    // stream-ordered allocations made during the current API call and not yet released: when the call fails
    // (tci_fail was reached) whatever is still listed here was leaked by an early return and is freed by the guard
    std::vector<void *> call_allocs;
    bool failed = false;
    // pinned staging for small device-to-host results (latency, not bandwidth)
    void *pinned = nullptr;
    size_t pinned_cap = 0;
};

// ---- multi-GPU group (group.cu) -----------------------------------------------------------------------------
// tci_ctx_create with ngpu > 1 (one process, as the reference's caller is): every GPU of the group is a member
// context of its own (stream, targets replicated under the same ids), peer access is enabled both ways (also for the
// stream-ordered pools), one worker thread per member issues that member's share of a sharded stage, and the
// collectives are NCCL group calls over the members' streams (ncclCommInitAll).
typedef struct ncclComm *ncclComm_t;
struct tci_group {
    int world = 1;  // ranks in the group
    int nlocal = 1; // members that live in this process
    std::vector<tci_ctx *> m;   // local members; m[0] is the caller's context
    std::vector<ncclComm_t> comm;
    cudaEvent_t ev_owner = nullptr; // recorded on the owner's stream; members wait on it before touching owner memory
    // worker threads (members 1..nlocal-1)
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::function<int(int)> job;
    unsigned long long job_gen = 0;
    int pending = 0;
    std::vector<int> job_rc;
    bool stop = false;
};
// runs f(k) for every local member k (k = 0 on the calling thread), each with its device current; returns the first
// non-zero status
int group_run(tci_group *g, const std::function<int(int)> &f);
// orders every member's stream behind what the owner's stream has enqueued so far (buffers the owner allocated in
// stream order must not be touched earlier)
void group_follow_owner(tci_group *g);
// collectives over all ranks (in place on every local member; ptr(k) = member k's buffer)
int group_allreduce_max_u64(tci_group *g, const std::function<unsigned long long *(int)> &ptr, size_t count);
int group_allgather(tci_group *g, const std::function<void *(int)> &ptr, size_t bytes_per_rank);
// the same all-gather on every member's copy stream, ordered behind what its main stream has enqueued so far; the main
// streams go on without it until group_allgather_join makes them wait for it (a chain that does not need the gathered
// data overlaps with the transfer)
int group_allgather_side(tci_group *g, const std::function<void *(int)> &ptr, size_t bytes_per_rank);
void group_allgather_join(tci_group *g);
int group_broadcast(tci_group *g, const std::function<void *(int)> &ptr, size_t bytes, int root);
void group_destroy(tci_group *g);
void target_free(tci_ctx *ctx, TargetDev &t);  // ctx.cu
int user_target_load(tci_ctx *ctx, TargetDev &t); // user_target.cu
void user_target_unload(TargetDev &t);
int pi_eval_user(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                 tci_dmat *out, unsigned long long *d_maxbits);
int target_eval_user(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out);
void *ctx_pinned(tci_ctx *ctx, size_t bytes); // page-locked staging of at least `bytes` (ctx.cu)
// copies target `id` of the caller's context to the other local members under the same id (peer copies)   ctx.cu
int target_replicate(tci_ctx *ctx, i64 id);
void ctx_release(tci_ctx *c); // frees the context once it is destroyed and no handle points at it any more
static inline int ctx_world(const tci_ctx *c) { return c->grp ? c->grp->world : 1; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember it per context, not per process
static inline cudaError_t ctx_func_smem(tci_ctx *ctx, const void *fn, int bytes)
{
    if (ctx->smem_configured.count(fn)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) ctx->smem_configured.insert(fn);
    return e;
}

struct tci_lu {
    tci_ctx *ctx = nullptr;
    tci_dmat *A = nullptr; // factorised in place (rows physically permuted, columns virtually)
    i64 m = 0, n = 0, r = 0;
    bool leftorthogonal = true;
    bool is_complex = false;  // A holds interleaved (re, im) pairs: 2m rows of doubles (zpath.cu)
    void *arena = nullptr;    // owns the three arrays below
    i64 *d_rowperm = nullptr; // 0-based, device
    i64 *d_colperm = nullptr; // position -> physical column, 0-based, device
    int *d_colpos = nullptr;  // physical column -> position
};

int tci_fail(tci_ctx *ctx, int code, const std::string &msg);

#define TCI_CUDA(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return tci_fail((ctx), TCI_ERR_CUDA,                                                        \
                            std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                                std::to_string(__LINE__) + ")");                                        \
    } while (0)

struct CtxGuard { // single-caller contract (SURVEY 8b "Threading")
    tci_ctx *c;
    bool ok;
    explicit CtxGuard(tci_ctx *ctx) : c(ctx), ok(false)
    {
        std::lock_guard<std::mutex> g(c->mu);
        if (!c->busy) {
            c->busy = true;
            ok = true;
            c->call_allocs.clear();
            c->failed = false;
        }
    }
    ~CtxGuard()
    {
        if (ok) {
            // an error return (typically out of memory on a large Pi) must not keep the biggest buffers: every
            // scratch allocation of this call that no RAII owner released is freed here, in stream order
            if (c->failed)
                for (void *p : c->call_allocs) cudaFreeAsync(p, c->stream);
            c->call_allocs.clear();
            c->failed = false;
            std::lock_guard<std::mutex> g(c->mu);
            c->busy = false;
        }
    }
};
#define TCI_ENTER(ctx)                                                          \
    if (!(ctx)) return TCI_ERR_ARG;                                             \
    CtxGuard guard__(ctx);                                                      \
    if (!guard__.ok) return tci_fail((ctx), TCI_ERR_BUSY, "context is in use"); \
    cudaSetDevice((ctx)->device)

struct StageTimer {
    tci_ctx *c;
    int stage;
    StageTimer(tci_ctx *ctx, int st) : c(ctx->nosync ? nullptr : ctx), stage(st)
    {
        if (c) cudaEventRecord(c->ev0, c->stream);
    }
    void stop()
    {
        if (!c) return;
        cudaEventRecord(c->ev1, c->stream);
        cudaEventSynchronize(c->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->stage_ms[stage] += ms;
        c = nullptr;
    }
    ~StageTimer() { stop(); }
};

static inline i64 round_up(i64 x, i64 a) { return (x + a - 1) / a * a; }

// stream-ordered allocations on the context stream (pools never trimmed while they work).  A pool that has to grow
// while it holds cached free blocks can fail with cudaErrorMemoryAllocation although the device is almost empty --
// reproducibly so for peer-mapped pools on the 2-GPU B200 box (tools/pool_peer_probe.cu) -- and succeeds once the
// cached blocks have been returned, so a failed allocation is retried once after a synchronise + trim.
static inline cudaError_t pool_alloc(tci_ctx *ctx, cudaMemPool_t pool, void **p, size_t bytes)
{
    bytes = bytes ? bytes : 8;
    cudaError_t e = pool ? cudaMallocFromPoolAsync(p, bytes, pool, ctx->stream) : cudaMallocAsync(p, bytes, ctx->stream);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        cudaMemPool_t pl = pool;
        if (!pl) cudaDeviceGetDefaultMemPool(&pl, ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (pl) cudaMemPoolTrimTo(pl, 0);
        e = pool ? cudaMallocFromPoolAsync(p, bytes, pool, ctx->stream) : cudaMallocAsync(p, bytes, ctx->stream);
        if (getenv("TCI_DEBUG_ALLOC"))
            fprintf(stderr, "[tci alloc] %zu bytes on device %d: retried after trim: %s\n", bytes, ctx->device,
                    cudaGetErrorString(e));
    }
    if (e == cudaSuccess && ctx->busy && ctx->member == 0) ctx->call_allocs.push_back(*p);
    return e;
}
static inline cudaError_t dev_alloc(tci_ctx *ctx, void **p, size_t bytes) { return pool_alloc(ctx, nullptr, p, bytes); }
// matrices other group members may store into (tci_dmat): from the peer-mapped pool when the context has one
static inline cudaError_t dev_alloc_shared(tci_ctx *ctx, void **p, size_t bytes)
{
    return pool_alloc(ctx, ctx->shared_pool, p, bytes);
}
static inline void dev_free(tci_ctx *ctx, void *p)
{
    if (!p) return;
    for (size_t q = ctx->call_allocs.size(); q-- > 0;)
        if (ctx->call_allocs[q] == p) {
            ctx->call_allocs[q] = ctx->call_allocs.back();
            ctx->call_allocs.pop_back();
            break;
        }
    cudaFreeAsync(p, ctx->stream);
}
template <typename T> struct DevBuf { // RAII scratch buffer
    tci_ctx *ctx;
    T *p = nullptr;
    DevBuf(tci_ctx *c) : ctx(c) {}
    cudaError_t alloc(size_t count) { return dev_alloc(ctx, (void **)&p, count * sizeof(T)); }
    cudaError_t upload(const T *host, size_t count)
    {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    }
    ~DevBuf() { dev_free(ctx, p); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

// device allocation helpers (dmat.cu)
int dmat_alloc(tci_ctx *ctx, i64 m, i64 n, tci_dmat **out);

// dgemm.cu: C(MxN, ldc) = alpha * op(A) * op(B) + beta * C on the context stream
int dgemm_dev(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
              const double *B, i64 ldb, double beta, double *C, i64 ldc);
// batched variant: pointers advance by strideA/B/C per batch entry
int dgemm_dev_batched(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
                      i64 strideA, const double *B, i64 ldb, i64 strideB, double beta, double *C, i64 ldc,
                      i64 strideC, i64 batch);

// per-batch element offsets added to A / B (device arrays, nullable)
int dgemm_dev_batched_off(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
                          i64 strideA, const double *B, i64 ldb, i64 strideB, double beta, double *C, i64 ldc,
                          i64 strideC, i64 batch, const i64 *offA, const i64 *offB, bool offsets_even = false);

// pi_eval.cu / tt.cu / mpo.cu
int pi_eval_analytic(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                     tci_dmat *out, unsigned long long *d_maxbits);
int pi_eval_tt(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
               tci_dmat *out);
// hI / hJ: host copies of the index sets (nullable) -- shared prefixes / suffixes are evaluated once (mpo.cu)
int pi_eval_mpo(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                tci_dmat *out, const i64 *hI = nullptr, const i64 *hJ = nullptr);
int env_eval_tt(tci_ctx *ctx, TargetDev &t, int side, const i64 *d_idx, int len, i64 count, double **out, i64 *D);
int env_eval_mpo(tci_ctx *ctx, TargetDev &t, int side, const i64 *d_idx, int len, i64 count, double **out, i64 *D,
                 const i64 *h_idx = nullptr);
// ComplexF64 building blocks (zgemm.cu, mpo.cu): element counts / leading dimensions in (re, im) pairs
int zgemm_dev(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double2 *A, i64 lda,
              const double2 *B, i64 ldb, double beta, double2 *C, i64 ldc);
int zgemm_dev_batched_off(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double2 *A, i64 lda,
                          i64 strideA, const double2 *B, i64 ldb, i64 strideB, double beta, double2 *C, i64 ldc,
                          i64 strideC, i64 batch, const i64 *offA, const i64 *offB);
int pi_eval_mpo_z(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                  tci_dmat *out, const i64 *hI, const i64 *hJ);
int target_eval_mpo_z(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out);
int maxabs_dev(tci_ctx *ctx, const double *p, i64 m, i64 n, i64 ld, unsigned long long *d_maxbits);
// cache.cu: CachedFunction as a device-resident memo
int pi_eval_cached(tci_ctx *ctx, i64 target_id, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ,
                   i64 M, tci_dmat *out, unsigned long long *d_maxbits);
int target_eval_cached(tci_ctx *ctx, i64 target_id, TargetDev &t, const i64 *d_idx, i64 count, double *d_out);
void cache_target_free(tci_ctx *ctx, i64 id);
