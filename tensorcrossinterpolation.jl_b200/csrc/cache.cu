// cache.cu -- CachedFunction (cachedfunction.jl:1-230) as a device-resident memo in HBM.
//
// The reference wraps an expensive f in a Dict{UInt128, V} keyed by key(x) = sum_n coeffs[n] * (x_n - 1),
// coeffs[n] = prod_{m<n} localdims[m] (:14-17, :177-184).  Here the Dict is an open-addressing hash table in HBM with
// the same 128-bit keys, and a cached target is a target like any other (kind 4, wrapping a real-valued inner target):
// tci_pi_eval / tci_target_eval / tci_bond_update / tci_globalsearch on it follow the structure of
// _batcheval_imp_for_batchevaluator (:117-171) --
//   1. look every requested element up (keys of rows, centre combinations and columns are summed, one thread per
//      element); hits are written, misses are compacted into a list (warp-aggregated atomics),
//   2. the inner target is evaluated on the missing points only, in one batch,
//   3. the new values are written to the result and inserted into the table.
// A slot is claimed by a 64-bit fingerprint CAS, filled, and published by setting the fingerprint's READY bit, so that
// two threads that insert the same key in the same launch (the reference's test evaluates 100 identical index pairs)
// end up with ONE entry.  A table that is full keeps answering correctly: values that cannot be inserted are simply
// not memoised.  f must be pure, as it must be for the reference's Dict.
#include "tci_internal.h"

struct CacheSlot {
    unsigned long long tag;      // 0: empty; fingerprint (bit 63 clear): claimed; fingerprint | READY: published
    unsigned long long klo, khi; // the UInt128 key
    double val;
};
static const unsigned long long CACHE_READY = 1ull << 63;

struct CacheDev {
    CacheSlot *slots = nullptr;
    unsigned long long mask = 0;           // capacity - 1 (capacity is a power of two)
    unsigned long long *counters = nullptr; // [0] entries, [1] hits, [2] misses, [3] failed inserts, [4] miss cursor
    unsigned long long *coeff = nullptr;    // (lo, hi) per site
    i64 inner = 0;
};
static std::map<std::pair<tci_ctx *, i64>, CacheDev> g_caches;
static std::mutex g_caches_mu;

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{ // splitmix64 finaliser
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
__device__ __forceinline__ void add128(unsigned long long &lo, unsigned long long &hi, unsigned long long blo, unsigned long long bhi)
{
    const unsigned long long s = lo + blo;
    hi += bhi + (s < lo ? 1ull : 0ull);
    lo = s;
}
// coeff * v for a 128-bit coeff and a small v (< 2^63): (lo, hi) of the low 128 bits
__device__ __forceinline__ void mul128(unsigned long long clo, unsigned long long chi, unsigned long long v, unsigned long long &lo,
                                       unsigned long long &hi)
{
    lo = clo * v;
    hi = __umul64hi(clo, v) + chi * v;
}

// partial keys of `count` partial multi-indices (len sites starting at site `first`); idx is (len x count)
__global__ void k_cache_partial_keys(const i64 *__restrict__ idx, int len, int first, i64 count,
                                     const unsigned long long *__restrict__ coeff, unsigned long long *__restrict__ out)
{
    const i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    unsigned long long lo = 0, hi = 0;
    for (int s = 0; s < len; ++s) {
        unsigned long long a, b;
        mul128(coeff[2 * (first + s)], coeff[2 * (first + s) + 1], (unsigned long long)(idx[s + (i64)len * q] - 1), a, b);
        add128(lo, hi, a, b);
    }
    out[2 * q] = lo;
    out[2 * q + 1] = hi;
}
// partial keys of the C centre combinations (first centre index fastest, batcheval.jl:49-60); cdims: M local dimensions
__global__ void k_cache_centre_keys(const i64 *__restrict__ cdims, int M, int first, i64 C,
                                    const unsigned long long *__restrict__ coeff, unsigned long long *__restrict__ out)
{
    const i64 c = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (c >= C) return;
    unsigned long long lo = 0, hi = 0;
    i64 rem = c;
    for (int s = 0; s < M; ++s) {
        const i64 v = rem % cdims[s];
        rem /= cdims[s];
        unsigned long long a, b;
        mul128(coeff[2 * (first + s)], coeff[2 * (first + s) + 1], (unsigned long long)v, a, b);
        add128(lo, hi, a, b);
    }
    out[2 * c] = lo;
    out[2 * c + 1] = hi;
}

__device__ __forceinline__ bool cache_lookup(const CacheSlot *slots, unsigned long long mask, unsigned long long klo,
                                             unsigned long long khi, double &val)
{
    const unsigned long long h = mix64(klo ^ mix64(khi + 0x9e3779b97f4a7c15ull));
    const unsigned long long fp = ((h >> 1) | 1ull) & ~CACHE_READY;
    unsigned long long pos = mix64(h) & mask;
    for (int probe = 0; probe < 128; ++probe) {
        const CacheSlot *s = slots + pos;
        const unsigned long long tag = __ldcg(&s->tag);
        if (tag == 0) return false;
        if (tag == (fp | CACHE_READY) && __ldcg(&s->klo) == klo && __ldcg(&s->khi) == khi) {
            val = __ldcg(&s->val);
            return true;
        }
        pos = (pos + 1) & mask;
    }
    return false;
}
// returns 1 if a new entry was made, 0 if the key was already there, -1 if it could not be stored
__device__ __forceinline__ int cache_insert(CacheSlot *slots, unsigned long long mask, unsigned long long klo, unsigned long long khi,
                                            double val)
{
    const unsigned long long h = mix64(klo ^ mix64(khi + 0x9e3779b97f4a7c15ull));
    const unsigned long long fp = ((h >> 1) | 1ull) & ~CACHE_READY;
    unsigned long long pos = mix64(h) & mask;
    for (int probe = 0; probe < 128; ++probe) {
        CacheSlot *s = slots + pos;
        unsigned long long old = atomicCAS(&s->tag, 0ull, fp);
        if (old == 0) {
            s->klo = klo;
            s->khi = khi;
            s->val = val;
            __threadfence();
            atomicExch(&s->tag, fp | CACHE_READY);
            return 1;
        }
        if ((old & ~CACHE_READY) == fp) { // same fingerprint: the same key unless 2^-63 says otherwise
            for (int spin = 0; spin < 4096 && !(old & CACHE_READY); ++spin) {
                __nanosleep(32); // the owner may be a lane of this warp
                old = *((volatile unsigned long long *)&s->tag);
            }
            if (!(old & CACHE_READY)) return -1;
            __threadfence();
            if (__ldcg(&s->klo) == klo && __ldcg(&s->khi) == khi) return 0;
        }
        pos = (pos + 1) & mask;
    }
    return -1;
}

// pass 1 over the (nI*C) x nJ block: hits go to out, misses to the list (element ids)
__global__ void k_cache_lookup_pi(const CacheSlot *__restrict__ slots, unsigned long long mask,
                                  const unsigned long long *__restrict__ rowkey, const unsigned long long *__restrict__ ckey,
                                  const unsigned long long *__restrict__ colkey, i64 nI, i64 C, i64 nJ, double *__restrict__ out,
                                  i64 ld, i64 *__restrict__ misslist, unsigned long long *__restrict__ counters)
{
    const i64 rows = nI * C, total = rows * nJ;
    const i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    bool miss = false;
    if (e < total) {
        const i64 r = e % rows, j = e / rows, i = r % nI, c = r / nI;
        unsigned long long lo = rowkey[2 * i], hi = rowkey[2 * i + 1];
        add128(lo, hi, ckey[2 * c], ckey[2 * c + 1]);
        add128(lo, hi, colkey[2 * j], colkey[2 * j + 1]);
        double v;
        if (cache_lookup(slots, mask, lo, hi, v))
            out[r + ld * j] = v;
        else
            miss = true;
    }
    const unsigned ball = __ballot_sync(0xffffffffu, miss);
    if (ball) {
        const int lane = threadIdx.x & 31, leader = __ffs(ball) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(&counters[4], (unsigned long long)__popc(ball));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (miss) misslist[base + __popc(ball & ((1u << lane) - 1))] = e;
    }
    if ((threadIdx.x & 31) == 0) {
        const int nm = __popc(ball);
        const i64 first = e, lastp1 = first + 32 < total ? first + 32 : total;
        const int nvalid = first < total ? (int)(lastp1 - first) : 0;
        if (nm) atomicAdd(&counters[2], (unsigned long long)nm);
        if (nvalid - nm > 0) atomicAdd(&counters[1], (unsigned long long)(nvalid - nm));
    }
}
// the full multi-indices of the missing elements, (n x nmiss)
__global__ void k_cache_miss_points(const i64 *__restrict__ misslist, i64 nmiss, const i64 *__restrict__ I, int nl, i64 nI,
                                    const i64 *__restrict__ cdims, int M, i64 C, const i64 *__restrict__ J, int nr,
                                    i64 *__restrict__ pts)
{
    const i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= nmiss) return;
    const i64 rows = nI * C, e = misslist[q];
    const i64 r = e % rows, j = e / rows, i = r % nI;
    i64 c = r / nI;
    const int n = nl + M + nr;
    i64 *p = pts + (i64)n * q;
    for (int s = 0; s < nl; ++s) p[s] = I[s + (i64)nl * i];
    for (int s = 0; s < M; ++s) {
        p[nl + s] = c % cdims[s] + 1;
        c /= cdims[s];
    }
    for (int s = 0; s < nr; ++s) p[nl + M + s] = J[s + (i64)nr * j];
}
// pass 3: the new values into the result and into the table
__global__ void k_cache_store_pi(CacheSlot *__restrict__ slots, unsigned long long mask, const i64 *__restrict__ misslist,
                                 const double *__restrict__ vals, i64 nmiss, const unsigned long long *__restrict__ rowkey,
                                 const unsigned long long *__restrict__ ckey, const unsigned long long *__restrict__ colkey, i64 nI,
                                 i64 C, double *__restrict__ out, i64 ld, unsigned long long *__restrict__ counters)
{
    const i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= nmiss) return;
    const i64 rows = nI * C, e = misslist[q];
    const i64 r = e % rows, j = e / rows, i = r % nI, c = r / nI;
    const double v = vals[q];
    out[r + ld * j] = v;
    unsigned long long lo = rowkey[2 * i], hi = rowkey[2 * i + 1];
    add128(lo, hi, ckey[2 * c], ckey[2 * c + 1]);
    add128(lo, hi, colkey[2 * j], colkey[2 * j + 1]);
    const int st = cache_insert(slots, mask, lo, hi, v);
    if (st == 1)
        atomicAdd(&counters[0], 1ull);
    else if (st < 0)
        atomicAdd(&counters[3], 1ull);
}

static CacheDev *cache_of(tci_ctx *ctx, i64 id)
{
    std::lock_guard<std::mutex> lk(g_caches_mu);
    auto it = g_caches.find({ctx, id});
    return it == g_caches.end() ? nullptr : &it->second;
}
void cache_target_free(tci_ctx *ctx, i64 id)
{
    std::lock_guard<std::mutex> lk(g_caches_mu);
    auto it = g_caches.find({ctx, id});
    if (it == g_caches.end()) return;
    cudaFree(it->second.slots);
    cudaFree(it->second.counters);
    cudaFree(it->second.coeff);
    g_caches.erase(it);
}

int target_eval_dev(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out); // pi_eval.cu

// The three passes over a block whose element e = r + rows*j has the key rowkey[i] + ckey[c] + colkey[j]; `points`
// builds the (n x nmiss) multi-indices of the missing elements.
static int cached_block(tci_ctx *ctx, CacheDev &cd, TargetDev &inner, const i64 *dI, int nl, i64 nI, const i64 *d_cdims, int M,
                        i64 C, const i64 *dJ, int nr, i64 nJ, double *out, i64 ld)
{
    const int n = nl + M + nr;
    const i64 rows = nI * C, total = rows * nJ;
    DevBuf<unsigned long long> keys(ctx);
    TCI_CUDA(ctx, keys.alloc((size_t)(2 * (nI + C + nJ))));
    unsigned long long *rowkey = keys.p, *ckey = keys.p + 2 * nI, *colkey = ckey + 2 * C;
    k_cache_partial_keys<<<(unsigned)((nI + 127) / 128), 128, 0, ctx->stream>>>(dI, nl, 0, nI, cd.coeff, rowkey);
    k_cache_centre_keys<<<(unsigned)((C + 127) / 128), 128, 0, ctx->stream>>>(d_cdims, M, nl, C, cd.coeff, ckey);
    k_cache_partial_keys<<<(unsigned)((nJ + 127) / 128), 128, 0, ctx->stream>>>(dJ, nr, nl + M, nJ, cd.coeff, colkey);
    DevBuf<i64> misslist(ctx);
    TCI_CUDA(ctx, misslist.alloc((size_t)total));
    TCI_CUDA(ctx, cudaMemsetAsync(cd.counters + 4, 0, 8, ctx->stream));
    k_cache_lookup_pi<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(cd.slots, cd.mask, rowkey, ckey, colkey, nI, C, nJ,
                                                                               out, ld, misslist.p, cd.counters);
    ctx->launches += 4;
    unsigned long long nmiss = 0;
    TCI_CUDA(ctx, cudaMemcpyAsync(&nmiss, cd.counters + 4, 8, cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the size of the second batch (:150-152 builds the same lists)
    if (nmiss == 0) return TCI_OK;
    DevBuf<i64> pts(ctx);
    DevBuf<double> vals(ctx);
    TCI_CUDA(ctx, pts.alloc((size_t)n * nmiss));
    TCI_CUDA(ctx, vals.alloc((size_t)nmiss));
    k_cache_miss_points<<<(unsigned)((nmiss + 127) / 128), 128, 0, ctx->stream>>>(misslist.p, (i64)nmiss, dI, nl, nI, d_cdims, M, C,
                                                                                 dJ, nr, pts.p);
    ctx->launches++;
    int rc = target_eval_dev(ctx, inner, pts.p, (i64)nmiss, vals.p);
    if (rc) return rc;
    k_cache_store_pi<<<(unsigned)((nmiss + 127) / 128), 128, 0, ctx->stream>>>(cd.slots, cd.mask, misslist.p, vals.p, (i64)nmiss,
                                                                              rowkey, ckey, colkey, nI, C, out, ld, cd.counters);
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

// Pi of a cached target (pi_enqueue, kind 4)
int pi_eval_cached(tci_ctx *ctx, i64 target_id, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ,
                   i64 M, tci_dmat *out, unsigned long long *d_maxbits)
{
    CacheDev *cd = cache_of(ctx, target_id);
    if (!cd) return tci_fail(ctx, TCI_ERR_ARG, "cached target: no table");
    auto it = ctx->targets.find(cd->inner);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "cached target: the wrapped target was destroyed");
    i64 C = 1;
    std::vector<i64> cdims((size_t)std::max<i64>(M, 1), 1);
    for (i64 s = 0; s < M; ++s) {
        cdims[s] = t.localdims[nl + s];
        C *= cdims[s];
    }
    DevBuf<i64> d_cdims(ctx);
    TCI_CUDA(ctx, d_cdims.upload(cdims.data(), cdims.size()));
    int rc = cached_block(ctx, *cd, *it->second, dI, (int)nl, nI, d_cdims.p, (int)M, C, dJ, (int)nr, nJ, out->p, out->ld);
    if (!rc && d_maxbits) rc = maxabs_dev(ctx, out->p, out->m, out->n, out->ld, d_maxbits);
    return rc;
}

// f(x) for `count` full multi-indices: the same passes with every point a "row" (nl = n, C = nJ = 1)
int target_eval_cached(tci_ctx *ctx, i64 target_id, TargetDev &t, const i64 *d_idx, i64 count, double *d_out)
{
    CacheDev *cd = cache_of(ctx, target_id);
    if (!cd) return tci_fail(ctx, TCI_ERR_ARG, "cached target: no table");
    auto it = ctx->targets.find(cd->inner);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "cached target: the wrapped target was destroyed");
    const i64 one = 1;
    DevBuf<i64> d_cdims(ctx);
    TCI_CUDA(ctx, d_cdims.upload(&one, 1));
    return cached_block(ctx, *cd, *it->second, d_idx, (int)t.nsites, count, d_cdims.p, 0, 1, nullptr, 0, 1, d_out, count);
}

extern "C" int tci_target_cached(tci_ctx *ctx, int64_t inner_id, int capacity_log2, int64_t *target_id)
{
    TCI_ENTER(ctx);
    if (!target_id) return tci_fail(ctx, TCI_ERR_ARG, "tci_target_cached: target_id missing");
    if (ctx->grp && ctx->grp->world > 1)
        return tci_fail(ctx, TCI_ERR_UNSUPPORTED, "tci_target_cached: the memo lives in one GPU's HBM; use a single-GPU context");
    auto it = ctx->targets.find(inner_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &in = *it->second;
    if (in.is_complex || in.kind == 4) return tci_fail(ctx, TCI_ERR_ARG, "tci_target_cached: the wrapped target must be a real-valued, uncached target");
    if (capacity_log2 < 4 || capacity_log2 > 34) return tci_fail(ctx, TCI_ERR_ARG, "tci_target_cached: capacity_log2 must be in 4 .. 34");
    // coeffs[n] = prod_{m<n} localdims[m] as UInt128; the largest key must fit (cachedfunction.jl:22-24)
    const i64 n = in.nsites;
    std::vector<unsigned long long> coeff((size_t)(2 * n));
    unsigned __int128 c = 1, maxkey = 0;
    bool overflow = false;
    for (i64 s = 0; s < n; ++s) {
        coeff[2 * s] = (unsigned long long)c;
        coeff[2 * s + 1] = (unsigned long long)(c >> 64);
        const unsigned __int128 d = (unsigned __int128)in.localdims[s], term = c * (d - 1);
        if (d > 1 && term / (d - 1) != c) overflow = true;
        if (maxkey + term < maxkey) overflow = true;
        maxkey += term;
        if (s + 1 < n) {
            const unsigned __int128 next = c * d;
            if (next / d != c) overflow = true;
            c = next;
        }
    }
    if (overflow)
        return tci_fail(ctx, TCI_ERR_ARG, "Overflow in CachedFunction. Use ValueType = a bigger type with fixed size, e.g., BitIntegers.UInt256");
    CacheDev cd;
    cd.inner = inner_id;
    const size_t cap = (size_t)1 << capacity_log2;
    cd.mask = cap - 1;
    cudaError_t e = cudaMalloc(&cd.slots, cap * sizeof(CacheSlot));
    if (e == cudaSuccess) e = cudaMemset(cd.slots, 0, cap * sizeof(CacheSlot));
    if (e == cudaSuccess) e = cudaMalloc(&cd.counters, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(cd.counters, 0, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(&cd.coeff, coeff.size() * sizeof(unsigned long long));
    if (e == cudaSuccess)
        e = cudaMemcpy(cd.coeff, coeff.data(), coeff.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { // typically out of memory for a large table: nothing is kept
        cudaFree(cd.slots);
        cudaFree(cd.counters);
        cudaFree(cd.coeff);
        return tci_fail(ctx, TCI_ERR_CUDA, std::string("tci_target_cached: ") + cudaGetErrorString(e));
    }
    std::unique_ptr<TargetDev> t(new TargetDev());
    t->kind = 4;
    t->nsites = n;
    t->localdims = in.localdims;
    const i64 id = ctx->next_target++;
    t->cache_id = id;
    ctx->targets[id] = std::move(t);
    {
        std::lock_guard<std::mutex> lk(g_caches_mu);
        g_caches[{ctx, id}] = cd;
    }
    *target_id = id;
    return TCI_OK;
}

// length(cf.cache) and the traffic of the memo: out[0] entries, out[1] hits, out[2] misses, out[3] values that found no
// free slot within the probe limit (returned correctly, not memoised)
extern "C" int tci_target_cache_stats(tci_ctx *ctx, int64_t target_id, int64_t *out)
{
    TCI_ENTER(ctx);
    CacheDev *cd = cache_of(ctx, target_id);
    if (!cd || !out) return tci_fail(ctx, TCI_ERR_ARG, "tci_target_cache_stats: not a cached target");
    unsigned long long h[4];
    TCI_CUDA(ctx, cudaMemcpyAsync(h, cd->counters, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int q = 0; q < 4; ++q) out[q] = (int64_t)h[q];
    return TCI_OK;
}
