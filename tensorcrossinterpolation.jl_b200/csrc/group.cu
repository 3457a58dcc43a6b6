// group.cu -- multi-GPU groups behind the C ABI (tci_ctx_create(ngpu, device_ids), 8e).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process this resolves to the library torch
// already loaded, for a Julia caller to the system one), so libtci_b200.so has no link-time dependency on it and a
// single-GPU context never touches it.  Only the handful of entry points below are used; their signatures have been
// stable since NCCL 2.0.
#include <dlfcn.h>

#include <cstring>

#include "tci_internal.h"

enum { NCCL_U8 = 1, NCCL_U64 = 5, NCCL_MAX = 2 };

struct NcclApi {
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

static NcclApi &nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
            return;
        }
#define NCCL_SYM(field, name)                                      \
    *(void **)(&api.field) = dlsym(h, name);                       \
    if (!api.field) {                                              \
        api.why = std::string("NCCL symbol missing: ") + name;     \
        return;                                                    \
    }
        NCCL_SYM(CommInitAll, "ncclCommInitAll");
        NCCL_SYM(CommDestroy, "ncclCommDestroy");
        NCCL_SYM(AllGather, "ncclAllGather");
        NCCL_SYM(AllReduce, "ncclAllReduce");
        NCCL_SYM(Broadcast, "ncclBroadcast");
        NCCL_SYM(GroupStart, "ncclGroupStart");
        NCCL_SYM(GroupEnd, "ncclGroupEnd");
        NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef NCCL_SYM
        api.ok = true;
    });
    return api;
}

#define TCI_NCCL(ctx, call)                                                                                      \
    do {                                                                                                         \
        int r__ = (call);                                                                                        \
        if (r__ != 0)                                                                                            \
            return tci_fail((ctx), TCI_ERR_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r__));        \
    } while (0)

// ------------------------------------------------------------------ worker threads --------------------------
static void worker_loop(tci_group *g, int k)
{
    cudaSetDevice(g->m[k]->device);
    unsigned long long seen = 0;
    for (;;) {
        std::function<int(int)> job;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv_job.wait(lk, [&] { return g->stop || g->job_gen != seen; });
            if (g->stop) return;
            seen = g->job_gen;
            job = g->job;
        }
        int rc = job(k);
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->job_rc[k] = rc;
            if (--g->pending == 0) g->cv_done.notify_all();
        }
    }
}

int group_run(tci_group *g, const std::function<int(int)> &f)
{
    if (!g || g->nlocal == 1) return f(0);
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->job = f;
        g->pending = g->nlocal - 1;
        std::fill(g->job_rc.begin(), g->job_rc.end(), 0);
        g->job_gen++;
    }
    g->cv_job.notify_all();
    cudaSetDevice(g->m[0]->device);
    const int rc0 = f(0);
    {
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv_done.wait(lk, [&] { return g->pending == 0; });
    }
    cudaSetDevice(g->m[0]->device);
    if (rc0) return rc0;
    for (int k = 1; k < g->nlocal; ++k)
        if (g->job_rc[k]) { // the message sits in the member's context: surface it on the caller's
            g->m[0]->err = g->m[k]->err;
            return g->job_rc[k];
        }
    return TCI_OK;
}

void group_follow_owner(tci_group *g)
{
    tci_ctx *c0 = g->m[0];
    cudaSetDevice(c0->device);
    cudaEventRecord(g->ev_owner, c0->stream);
    for (int k = 1; k < g->nlocal; ++k) cudaStreamWaitEvent(g->m[k]->stream, g->ev_owner, 0);
}

// ------------------------------------------------------------------ collectives -----------------------------
int group_allreduce_max_u64(tci_group *g, const std::function<unsigned long long *(int)> &ptr, size_t count)
{
    tci_ctx *c0 = g->m[0];
    TCI_NCCL(c0, nccl().GroupStart());
    for (int k = 0; k < g->nlocal; ++k) {
        unsigned long long *p = ptr(k);
        int r = nccl().AllReduce(p, p, count, NCCL_U64, NCCL_MAX, g->comm[k], g->m[k]->stream);
        if (r) {
            nccl().GroupEnd();
            return tci_fail(c0, TCI_ERR_CUDA, std::string("ncclAllReduce: ") + nccl().GetErrorString(r));
        }
    }
    TCI_NCCL(c0, nccl().GroupEnd());
    for (int k = 0; k < g->nlocal; ++k) g->m[k]->launches++;
    return TCI_OK;
}

// in place: rank q's block sits at ptr + q * bytes_per_rank on every member before the call
int group_allgather(tci_group *g, const std::function<void *(int)> &ptr, size_t bytes_per_rank)
{
    tci_ctx *c0 = g->m[0];
    if (bytes_per_rank == 0) return TCI_OK;
    TCI_NCCL(c0, nccl().GroupStart());
    for (int k = 0; k < g->nlocal; ++k) {
        char *base = static_cast<char *>(ptr(k));
        const int rank = g->m[k]->rank;
        int r = nccl().AllGather(base + (size_t)rank * bytes_per_rank, base, bytes_per_rank, NCCL_U8, g->comm[k],
                                 g->m[k]->stream);
        if (r) {
            nccl().GroupEnd();
            return tci_fail(c0, TCI_ERR_CUDA, std::string("ncclAllGather: ") + nccl().GetErrorString(r));
        }
    }
    TCI_NCCL(c0, nccl().GroupEnd());
    for (int k = 0; k < g->nlocal; ++k) g->m[k]->launches++;
    return TCI_OK;
}

int group_allgather_side(tci_group *g, const std::function<void *(int)> &ptr, size_t bytes_per_rank)
{
    tci_ctx *c0 = g->m[0];
    if (bytes_per_rank == 0) return TCI_OK;
    for (int k = 0; k < g->nlocal; ++k) {
        tci_ctx *c = g->m[k];
        cudaSetDevice(c->device);
        cudaEventRecord(c->ev_g0, c->stream);
        cudaStreamWaitEvent(c->copy_stream, c->ev_g0, 0);
    }
    TCI_NCCL(c0, nccl().GroupStart());
    for (int k = 0; k < g->nlocal; ++k) {
        char *base = static_cast<char *>(ptr(k));
        const int rank = g->m[k]->rank;
        int r = nccl().AllGather(base + (size_t)rank * bytes_per_rank, base, bytes_per_rank, NCCL_U8, g->comm[k],
                                 g->m[k]->copy_stream);
        if (r) {
            nccl().GroupEnd();
            return tci_fail(c0, TCI_ERR_CUDA, std::string("ncclAllGather: ") + nccl().GetErrorString(r));
        }
    }
    TCI_NCCL(c0, nccl().GroupEnd());
    for (int k = 0; k < g->nlocal; ++k) {
        tci_ctx *c = g->m[k];
        cudaSetDevice(c->device);
        cudaEventRecord(c->ev_g1, c->copy_stream);
        c->launches++;
    }
    cudaSetDevice(c0->device);
    return TCI_OK;
}

void group_allgather_join(tci_group *g)
{
    for (int k = 0; k < g->nlocal; ++k) {
        tci_ctx *c = g->m[k];
        cudaSetDevice(c->device);
        cudaStreamWaitEvent(c->stream, c->ev_g1, 0);
    }
    cudaSetDevice(g->m[0]->device);
}

int group_broadcast(tci_group *g, const std::function<void *(int)> &ptr, size_t bytes, int root)
{
    tci_ctx *c0 = g->m[0];
    if (bytes == 0) return TCI_OK;
    TCI_NCCL(c0, nccl().GroupStart());
    for (int k = 0; k < g->nlocal; ++k) {
        void *p = ptr(k);
        int r = nccl().Broadcast(p, p, bytes, NCCL_U8, root, g->comm[k], g->m[k]->stream);
        if (r) {
            nccl().GroupEnd();
            return tci_fail(c0, TCI_ERR_CUDA, std::string("ncclBroadcast: ") + nccl().GetErrorString(r));
        }
    }
    TCI_NCCL(c0, nccl().GroupEnd());
    for (int k = 0; k < g->nlocal; ++k) g->m[k]->launches++;
    return TCI_OK;
}

// ------------------------------------------------------------------ creation / destruction ------------------
int ctx_create_one(int device_id, tci_ctx **out); // ctx.cu

void group_destroy(tci_group *g)
{
    if (!g) return;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
    }
    g->cv_job.notify_all();
    for (std::thread &t : g->workers) t.join();
    for (size_t k = 0; k < g->m.size(); ++k) {
        cudaSetDevice(g->m[k]->device);
        cudaStreamSynchronize(g->m[k]->stream);
    }
    for (ncclComm_t cm : g->comm)
        if (cm) nccl().CommDestroy(cm);
    if (g->ev_owner) {
        cudaSetDevice(g->m[0]->device);
        cudaEventDestroy(g->ev_owner);
    }
    for (size_t k = 1; k < g->m.size(); ++k) { // member 0 is released by its own tci_ctx_destroy
        g->m[k]->grp = nullptr;
        g->m[k]->destroyed = true;
        ctx_release(g->m[k]);
    }
    delete g;
}

// single process, ngpu GPUs: peer access both ways (also for the stream-ordered pools), NCCL communicators, workers
int group_create_local(const std::vector<tci_ctx *> &members, tci_group **out)
{
    tci_ctx *c0 = members[0];
    const int n = (int)members.size();
    if (!nccl().ok) return tci_fail(nullptr, TCI_ERR_UNSUPPORTED, "multi-GPU context needs NCCL: " + nccl().why);
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, members[a]->device, members[b]->device);
            if (!can)
                return tci_fail(nullptr, TCI_ERR_UNSUPPORTED,
                                "GPUs " + std::to_string(members[a]->device) + " and " +
                                    std::to_string(members[b]->device) + " have no peer access (NVLink / NVSwitch)");
            cudaSetDevice(members[a]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(members[b]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        }
    // one explicit pool per member, mapped read/write on every other member: the matrices peers store into (Pi).
    // Scratch (environments, index lists, LU arenas) stays in the device's private default pool.
    for (int b = 0; b < n; ++b) {
        if (members[b]->shared_pool) continue;
        cudaSetDevice(members[b]->device);
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = members[b]->device;
        cudaMemPool_t pool = nullptr;
        cudaError_t e = cudaMemPoolCreate(&pool, &props);
        if (e != cudaSuccess)
            return tci_fail(nullptr, TCI_ERR_CUDA, std::string("cudaMemPoolCreate: ") + cudaGetErrorString(e));
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        std::vector<cudaMemAccessDesc> desc;
        for (int a = 0; a < n; ++a) {
            if (a == b) continue;
            cudaMemAccessDesc d{};
            d.location.type = cudaMemLocationTypeDevice;
            d.location.id = members[a]->device;
            d.flags = cudaMemAccessFlagsProtReadWrite;
            desc.push_back(d);
        }
        e = cudaMemPoolSetAccess(pool, desc.data(), desc.size());
        if (e != cudaSuccess) {
            cudaMemPoolDestroy(pool);
            return tci_fail(nullptr, TCI_ERR_CUDA, std::string("cudaMemPoolSetAccess: ") + cudaGetErrorString(e));
        }
        members[b]->shared_pool = pool;
    }
    tci_group *g = new tci_group();
    g->world = g->nlocal = n;
    g->m = members;
    g->comm.assign(n, nullptr);
    g->job_rc.assign(n, 0);
    std::vector<int> devs(n);
    for (int k = 0; k < n; ++k) devs[k] = members[k]->device;
    int r = nccl().CommInitAll(g->comm.data(), n, devs.data());
    if (r) {
        std::string msg = std::string("ncclCommInitAll: ") + nccl().GetErrorString(r);
        delete g;
        return tci_fail(nullptr, TCI_ERR_CUDA, msg);
    }
    for (int k = 0; k < n; ++k) {
        members[k]->grp = g;
        members[k]->member = k;
        members[k]->rank = k;
    }
    for (int k = 1; k < n; ++k) g->workers.emplace_back(worker_loop, g, k);
    cudaSetDevice(c0->device);
    cudaEventCreateWithFlags(&g->ev_owner, cudaEventDisableTiming);
    *out = g;
    return TCI_OK;
}

