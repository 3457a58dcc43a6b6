// tt.cu -- K4: tensor-train targets (TTCache, cachedtensortrain.jl:9-225) and K7: the
// default global pivot search (globalpivotfinder.jl:143-195).
//
// Environments are stored as (D x count): the vector of point q is contiguous, so a warp
// that walks the output bond index writes coalesced and reads the previous vector as a
// broadcast.  The chain steps accumulate over the bond index IN ORDER with a rounded
// multiply and a rounded add, which makes evaluate(tt, x) bit-identical to the
// left-to-right product of abstracttensortrain.jl:124-132 as restated by the oracle
// (needed so that global-search accept/reject decisions cannot flip).
#include "tci_internal.h"

// out[b + Dr*q] = sum_a prev[a + Dl*q] * T[a, sig_q, b]   (prev == nullptr: Dl == 1, prev = 1)
__global__ void k_env_left_step(const double *__restrict__ prev, const double *__restrict__ T, int Dl, int d, int Dr,
                                const i64 *__restrict__ idx, int len, int pos, i64 count, double *__restrict__ out)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= (i64)Dr * count) return;
    const int b = (int)(e % Dr);
    const i64 q = e / Dr;
    const i64 sig = idx[(i64)len * q + pos] - 1;
    const double *t = T + (i64)Dl * (sig + (i64)d * b);
    double acc = 0.0;
    if (prev) {
        const double *p = prev + (i64)Dl * q;
        for (int a = 0; a < Dl; ++a) acc = __dadd_rn(acc, __dmul_rn(p[a], t[a]));
    } else
        acc = t[0];
    out[e] = acc;
}

// The same ordered step for a tile of ENV_PT consecutive points per CTA (thread = output bond index b).  The probes of
// a global-search star, and the rows of any index set that was expanded site by site, are consecutive points that
// mostly share sigma at a given site: then the slice value T[a, sig, b] is loaded once and applied to all ENV_PT
// vectors, which sit in shared memory and are read as broadcasts (two bond indices per 16-byte read).  Every output
// is still acc = acc + p[a]*t[a] for a = 0, 1, ... with a rounded multiply and a rounded add, so the result is
// bit-identical to k_env_left_step.  Tiles with mixed sigma take the point-by-point path.
#define ENV_PT 32
#define ENV_AC 64
__global__ void __launch_bounds__(128, 5)
    k_env_left_step_tiled(const double *__restrict__ prev, const double *__restrict__ T, int Dl, int d, int Dr,
                          const i64 *__restrict__ idx, int len, int pos, i64 count, double *__restrict__ out,
                          const int *__restrict__ perm)
{
    __shared__ __align__(16) double ps[ENV_PT][ENV_AC];
    __shared__ int sg[ENV_PT];
    __shared__ i64 pq[ENV_PT]; // the points of this tile (through the bucket order when there is one)
    __shared__ int mixed;
    const i64 q0 = (i64)blockIdx.x * ENV_PT;
    const int np = (int)min((i64)ENV_PT, count - q0);
    const int b = blockIdx.y * 128 + threadIdx.x;
    if (threadIdx.x == 0) mixed = 0;
    if (threadIdx.x < ENV_PT) {
        const i64 q = threadIdx.x < np ? (perm ? (i64)perm[q0 + threadIdx.x] : q0 + threadIdx.x) : 0;
        pq[threadIdx.x] = q;
        sg[threadIdx.x] = threadIdx.x < np ? (int)(idx[(i64)len * q + pos] - 1) : 0;
    }
    __syncthreads();
    if (threadIdx.x < np && sg[threadIdx.x] != sg[0]) mixed = 1;
    __syncthreads();
    const bool uniform = !mixed;
    double acc[ENV_PT];
#pragma unroll
    for (int j = 0; j < ENV_PT; ++j) acc[j] = 0.0;
    const bool live = b < Dr;
    const double *tcol = T + (i64)Dl * (sg[0] + (i64)d * (live ? b : 0));
    for (int a0 = 0; a0 < Dl; a0 += ENV_AC) {
        const int alen = min(ENV_AC, Dl - a0);
        __syncthreads();
        for (int t = threadIdx.x; t < ENV_PT * ENV_AC; t += 128) {
            const int j = t / ENV_AC, a = t % ENV_AC;
            ps[j][a] = (j < np && a < alen) ? prev[(a0 + a) + (i64)Dl * pq[j]] : 0.0;
        }
        __syncthreads();
        if (!live) continue;
        if (uniform) {
            int a = 0;
            // the slice values of the NEXT pair of bond indices are requested before this pair is used
            double n0 = alen > 1 ? __ldg(tcol + a0) : 0.0, n1 = alen > 1 ? __ldg(tcol + a0 + 1) : 0.0;
            for (; a + 1 < alen; a += 2) {
                const double t0 = n0, t1 = n1;
                if (a + 3 < alen) {
                    n0 = __ldg(tcol + a0 + a + 2);
                    n1 = __ldg(tcol + a0 + a + 3);
                }
#pragma unroll
                for (int j = 0; j < ENV_PT; ++j) {
                    const double2 pv = *reinterpret_cast<const double2 *>(&ps[j][a]);
                    acc[j] = __dadd_rn(acc[j], __dmul_rn(pv.x, t0));
                    acc[j] = __dadd_rn(acc[j], __dmul_rn(pv.y, t1));
                }
            }
            if (a < alen) {
                const double t0 = __ldg(tcol + a0 + a);
#pragma unroll
                for (int j = 0; j < ENV_PT; ++j) acc[j] = __dadd_rn(acc[j], __dmul_rn(ps[j][a], t0));
            }
        } else {
#pragma unroll
            for (int j = 0; j < ENV_PT; ++j) {
                const double *tj = T + (i64)Dl * (sg[j] + (i64)d * b) + a0;
                double v = acc[j];
                for (int a = 0; a < alen; ++a) v = __dadd_rn(v, __dmul_rn(ps[j][a], __ldg(tj + a)));
                acc[j] = v;
            }
        }
    }
    if (live) {
#pragma unroll
        for (int j = 0; j < ENV_PT; ++j)
            if (j < np) out[b + (i64)Dr * pq[j]] = acc[j];
    }
}

// Bucket order of the points by their index at one site (counting sort, d <= ENV_MAXD buckets): tiles of the tiled
// step then hold points with ONE sigma whatever the order of the point set (the arm of a global-search star that
// varies this very site is the worst case: consecutive probes differ in nothing but sigma).  The order inside a
// bucket is arbitrary and does not matter: every point's result is computed independently.
#define ENV_MAXD 1024
__global__ void k_bucket_count(const i64 *__restrict__ idx, int len, int pos, i64 count, int d, int *__restrict__ hist)
{
    __shared__ int lh[ENV_MAXD];
    for (int t = threadIdx.x; t < d; t += blockDim.x) lh[t] = 0;
    __syncthreads();
    const i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q < count) atomicAdd(&lh[(int)(idx[(i64)len * q + pos] - 1)], 1);
    __syncthreads();
    for (int t = threadIdx.x; t < d; t += blockDim.x)
        if (lh[t]) atomicAdd(&hist[t], lh[t]);
}
__global__ void k_bucket_scan(int *hist, int d) // hist -> exclusive prefix sums (the scatter cursors)
{
    int run = 0;
    for (int t = 0; t < d; ++t) {
        const int c = hist[t];
        hist[t] = run;
        run += c;
    }
}
__global__ void k_bucket_scatter(const i64 *__restrict__ idx, int len, int pos, i64 count, int d,
                                 int *__restrict__ cursor, int *__restrict__ perm)
{
    __shared__ int lh[ENV_MAXD], base[ENV_MAXD];
    for (int t = threadIdx.x; t < d; t += blockDim.x) lh[t] = 0;
    __syncthreads();
    const i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    int sig = 0, r = 0;
    if (q < count) {
        sig = (int)(idx[(i64)len * q + pos] - 1);
        r = atomicAdd(&lh[sig], 1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < d; t += blockDim.x)
        if (lh[t]) base[t] = atomicAdd(&cursor[t], lh[t]);
    __syncthreads();
    if (q < count) perm[base[sig] + r] = (int)q;
}

// out[a + Dl*q] = sum_b T[a, sig_q, b] * prev[b + Dr*q]   (prev == nullptr: Dr == 1, prev = 1)
__global__ void k_env_right_step(const double *__restrict__ prev, const double *__restrict__ T, int Dl, int d, int Dr,
                                 const i64 *__restrict__ idx, int len, int pos, i64 count, double *__restrict__ out)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= (i64)Dl * count) return;
    const int a = (int)(e % Dl);
    const i64 q = e / Dl;
    const i64 sig = idx[(i64)len * q + pos] - 1;
    const double *t = T + a + (i64)Dl * sig;
    double acc = 0.0;
    if (prev) {
        const double *p = prev + (i64)Dr * q;
        for (int b = 0; b < Dr; ++b) acc = __dadd_rn(acc, __dmul_rn(t[(i64)Dl * d * b], p[b]));
    } else
        acc = t[0];
    out[e] = acc;
}

__global__ void k_dot_points(const double *__restrict__ l, const double *__restrict__ r, int D, i64 count,
                             double *__restrict__ out)
{
    i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    double acc = 0.0;
    for (int a = 0; a < D; ++a) acc = __dadd_rn(acc, __dmul_rn(l[a + (i64)D * q], r[a + (i64)D * q]));
    out[q] = acc;
}

__global__ void k_fill(double *p, i64 n, double v)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e < n) p[e] = v;
}

// GEMM form of a chain step for small site dimension d: Y = T^T * prev for ALL d slices at once
// (d x the flops of the gather form, but on the DMMA GEMM), then the slice sigma_q of point q is picked.
// left:  Y[(sig + d*b) + d*Dr*q]  -> out[b + Dr*q]      right: Y[(a + Dl*sig) + Dl*d*q] -> out[a + Dl*q]
__global__ void k_env_select_left(const double *__restrict__ Y, int d, int Dr, const i64 *__restrict__ idx, int len,
                                  int pos, i64 count, double *__restrict__ out)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= (i64)Dr * count) return;
    const int b = (int)(e % Dr);
    const i64 q = e / Dr;
    const i64 sig = idx[(i64)len * q + pos] - 1;
    out[e] = Y[(sig + (i64)d * b) + (i64)d * Dr * q];
}
__global__ void k_env_select_right(const double *__restrict__ Y, int Dl, int d, const i64 *__restrict__ idx, int len,
                                   int pos, i64 count, double *__restrict__ out)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= (i64)Dl * count) return;
    const int a = (int)(e % Dl);
    const i64 q = e / Dl;
    const i64 sig = idx[(i64)len * q + pos] - 1;
    out[e] = Y[(a + (i64)Dl * sig) + (i64)Dl * d * q];
}

#define TT_GEMM_MAX_D 8

struct CoreView {
    const double *p;
    int Dl, d, Dr;
};

// Left environment over sites [0, nsteps) of a chain; idx is (len x count), site s reads idx[s + off].
// Result (D x count) in *out (allocated here, caller frees with dev_free).
static int env_left_chain(tci_ctx *ctx, const std::vector<CoreView> &cores, int nsteps, const i64 *d_idx, int len,
                          int off, i64 count, double **out, int *Dout, bool ordered = true)
{
    double *prev = nullptr;
    int D = 1;
    DevBuf<int> permbuf(ctx), histbuf(ctx); // bucket order of the tiled step (allocated on first use)
    int *perm = nullptr, *hist = nullptr;
    for (int s = 0; s < nsteps; ++s) {
        const CoreView &c = cores[s];
        double *nxt = nullptr;
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)c.Dr * count * sizeof(double)));
        i64 total = (i64)c.Dr * count;
        if (!ordered && prev && c.d <= TT_GEMM_MAX_D && c.Dl >= 32 && count >= 64) {
            double *Y = nullptr;
            TCI_CUDA(ctx, dev_alloc(ctx, (void **)&Y, (size_t)c.d * c.Dr * count * sizeof(double)));
            int rc = dgemm_dev(ctx, true, false, (i64)c.d * c.Dr, count, c.Dl, 1.0, c.p, c.Dl, prev, c.Dl, 0.0, Y,
                               (i64)c.d * c.Dr);
            if (rc) return rc;
            k_env_select_left<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(Y, c.d, c.Dr, d_idx, len,
                                                                                       s + off, count, nxt);
            dev_free(ctx, Y);
        } else if (prev && count >= 4 * ENV_PT && c.Dl >= 8 && !getenv("TCI_TT_NO_TILED")) {
            dim3 grid((unsigned)((count + ENV_PT - 1) / ENV_PT), (unsigned)((c.Dr + 127) / 128));
            const bool bucket = c.d > 1 && c.d <= ENV_MAXD && count < 0x7fffffff && !getenv("TCI_TT_NO_BUCKETS");
            if (bucket) {
                if (!perm) {
                    TCI_CUDA(ctx, permbuf.alloc((size_t)count));
                    TCI_CUDA(ctx, histbuf.alloc(ENV_MAXD));
                    perm = permbuf.p;
                    hist = histbuf.p;
                }
                const unsigned nb = (unsigned)((count + 255) / 256);
                TCI_CUDA(ctx, cudaMemsetAsync(hist, 0, (size_t)c.d * sizeof(int), ctx->stream));
                k_bucket_count<<<nb, 256, 0, ctx->stream>>>(d_idx, len, s + off, count, c.d, hist);
                k_bucket_scan<<<1, 1, 0, ctx->stream>>>(hist, c.d);
                k_bucket_scatter<<<nb, 256, 0, ctx->stream>>>(d_idx, len, s + off, count, c.d, hist, perm);
                ctx->launches += 3;
            }
            k_env_left_step_tiled<<<grid, 128, 0, ctx->stream>>>(prev, c.p, c.Dl, c.d, c.Dr, d_idx, len, s + off,
                                                                count, nxt, bucket ? perm : nullptr);
        } else
            k_env_left_step<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(prev, c.p, c.Dl, c.d, c.Dr,
                                                                                     d_idx, len, s + off, count, nxt);
        ctx->launches++;
        dev_free(ctx, prev);
        prev = nxt;
        D = c.Dr;
    }
    if (!prev) { // no sites: ones(1 x count)
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&prev, (size_t)count * sizeof(double)));
        k_fill<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(prev, count, 1.0);
        ctx->launches++;
    }
    *out = prev;
    *Dout = D;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

// Right environment over the last nsteps sites; site s (global) reads idx[s - (N - nsteps) + off].
static int env_right_chain(tci_ctx *ctx, const std::vector<CoreView> &cores, int nsteps, const i64 *d_idx, int len,
                           int off, i64 count, double **out, int *Dout, bool ordered = true)
{
    const int N = (int)cores.size();
    double *prev = nullptr;
    int D = 1;
    for (int s = N - 1; s >= N - nsteps; --s) {
        const CoreView &c = cores[s];
        double *nxt = nullptr;
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)c.Dl * count * sizeof(double)));
        i64 total = (i64)c.Dl * count;
        if (!ordered && prev && c.d <= TT_GEMM_MAX_D && c.Dr >= 32 && count >= 64) {
            double *Y = nullptr;
            TCI_CUDA(ctx, dev_alloc(ctx, (void **)&Y, (size_t)c.Dl * c.d * count * sizeof(double)));
            int rc = dgemm_dev(ctx, false, false, (i64)c.Dl * c.d, count, c.Dr, 1.0, c.p, (i64)c.Dl * c.d, prev, c.Dr,
                               0.0, Y, (i64)c.Dl * c.d);
            if (rc) return rc;
            k_env_select_right<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(
                Y, c.Dl, c.d, d_idx, len, s - (N - nsteps) + off, count, nxt);
            dev_free(ctx, Y);
        } else
            k_env_right_step<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(
                prev, c.p, c.Dl, c.d, c.Dr, d_idx, len, s - (N - nsteps) + off, count, nxt);
        ctx->launches++;
        dev_free(ctx, prev);
        prev = nxt;
        D = c.Dl;
    }
    if (!prev) {
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&prev, (size_t)count * sizeof(double)));
        k_fill<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(prev, count, 1.0);
        ctx->launches++;
    }
    *out = prev;
    *Dout = D;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

static std::vector<CoreView> views(const TargetDev &t)
{
    std::vector<CoreView> v;
    for (i64 s = 0; s < t.nsites; ++s) v.push_back({t.cores[s], (int)t.dl[s], (int)t.d[s], (int)t.dr[s]});
    return v;
}

// environments alone (tci_env_eval): side 0 = evaluateleft over the first len sites, side 1 = evaluateright over the
// last len sites (cachedtensortrain.jl:77-128), GEMM form as in the batched Pi
int env_eval_tt(tci_ctx *ctx, TargetDev &t, int side, const i64 *d_idx, int len, i64 count, double **out, i64 *D)
{
    std::vector<CoreView> cv = views(t);
    int d = 1;
    int rc = side == 0 ? env_left_chain(ctx, cv, len, d_idx, len, 0, count, out, &d, false)
                       : env_right_chain(ctx, cv, len, d_idx, len, 0, count, out, &d, false);
    *D = d;
    return rc;
}

// batchevaluate(::TTCache) cachedtensortrain.jl:151-215 (projector = nothing)
int pi_eval_tt(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
               tci_dmat *out)
{
    std::vector<CoreView> cv = views(t);
    double *lenv = nullptr, *renv = nullptr;
    int DL = 1, DR = 1;
    // the batched Pi is a GEMM product anyway (1e-10 bar): small-d chain steps go through the GEMM as well
    int rc = env_left_chain(ctx, cv, (int)nl, dI, (int)nl, 0, nI, &lenv, &DL, false);
    if (rc) return rc;
    rc = env_right_chain(ctx, cv, (int)nr, dJ, (int)nr, 0, nJ, &renv, &DR, false);
    if (rc) {
        dev_free(ctx, lenv);
        return rc;
    }
    // centre sites: (rows x D) * (D x d*D') -> (rows*d x D')   :198-208
    double *cur = lenv; // first operand is stored transposed (D x nI)
    bool cur_T = true;
    i64 rows = nI;
    int D = DL;
    for (i64 s = nl; s < nl + M && !rc; ++s) {
        const CoreView &c = cv[s];
        double *nxt = nullptr;
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)rows * c.d * c.Dr * sizeof(double)));
        rc = dgemm_dev(ctx, cur_T, false, rows, (i64)c.d * c.Dr, D, 1.0, cur, cur_T ? D : rows, c.p, c.Dl, 0.0, nxt,
                       rows);
        dev_free(ctx, cur);
        cur = nxt;
        cur_T = false;
        rows *= c.d;
        D = c.Dr;
    }
    if (!rc) // (rows x D) * (D x nJ)   :211-212
        rc = dgemm_dev(ctx, cur_T, false, rows, nJ, D, 1.0, cur, cur_T ? D : rows, renv, DR, 0.0, out->p, out->ld);
    dev_free(ctx, cur);
    dev_free(ctx, renv);
    return rc;
}

// (tt::TTCache)(indexset): dot of the two half environments, cachedtensortrain.jl:130-146
int target_eval_tt(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out)
{
    std::vector<CoreView> cv = views(t);
    const int N = (int)t.nsites, mid = N / 2;
    double *l = nullptr, *r = nullptr;
    int Dl = 1, Dr = 1;
    int rc = env_left_chain(ctx, cv, mid, d_idx, N, 0, count, &l, &Dl);
    if (rc) return rc;
    rc = env_right_chain(ctx, cv, N - mid, d_idx, N, mid, count, &r, &Dr);
    if (!rc) {
        k_dot_points<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(l, r, Dl, count, d_out);
        ctx->launches++;
    }
    dev_free(ctx, l);
    dev_free(ctx, r);
    return rc;
}

struct HostTT { // a tensor train uploaded for one call
    tci_ctx *ctx;
    std::vector<double *> dev;
    std::vector<CoreView> cv;
    explicit HostTT(tci_ctx *c) : ctx(c) {}
    int upload(i64 nsites, const i64 *dims3, const double *const *cores)
    {
        for (i64 s = 0; s < nsites; ++s) {
            i64 Dl = dims3[3 * s], d = dims3[3 * s + 1], Dr = dims3[3 * s + 2];
            double *p = nullptr;
            TCI_CUDA(ctx, dev_alloc(ctx, (void **)&p, (size_t)(Dl * d * Dr) * sizeof(double)));
            dev.push_back(p);
            TCI_CUDA(ctx, cudaMemcpyAsync(p, cores[s], Dl * d * Dr * sizeof(double), cudaMemcpyHostToDevice,
                                          ctx->stream));
            cv.push_back({p, (int)Dl, (int)d, (int)Dr});
        }
        return TCI_OK;
    }
    ~HostTT()
    {
        for (double *p : dev) dev_free(ctx, p);
    }
};

// evaluate(tt, x): ordered product, abstracttensortrain.jl:124-132
static int tt_eval_points(tci_ctx *ctx, const std::vector<CoreView> &cv, const i64 *d_idx, i64 count, double *d_out)
{
    double *env = nullptr;
    int D = 1;
    int rc = env_left_chain(ctx, cv, (int)cv.size(), d_idx, (int)cv.size(), 0, count, &env, &D);
    if (rc) return rc;
    TCI_CUDA(ctx, cudaMemcpyAsync(d_out, env, count * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    dev_free(ctx, env);
    return TCI_OK;
}

extern "C" int tci_tt_evaluate(tci_ctx *ctx, int64_t nsites, const int64_t *dims3, const double *const *cores,
                               const int64_t *idx, int64_t count, double *out)
{
    TCI_ENTER(ctx);
    if (count <= 0) return TCI_OK;
    HostTT tt(ctx);
    int rc = tt.upload(nsites, dims3, cores);
    if (rc) return rc;
    if (tt.cv.back().Dr != 1 || tt.cv.front().Dl != 1) return tci_fail(ctx, TCI_ERR_ARG, "boundary bonds must be 1");
    DevBuf<i64> d_idx(ctx);
    DevBuf<double> d_out(ctx);
    TCI_CUDA(ctx, d_idx.upload(idx, (size_t)(nsites * count)));
    TCI_CUDA(ctx, d_out.alloc((size_t)count));
    rc = tt_eval_points(ctx, tt.cv, d_idx.p, count, d_out.p);
    if (rc) return rc;
    TCI_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

int target_eval_dev(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out); // pi_eval.cu

__global__ void k_abs_diff(const double *__restrict__ f, const double *__restrict__ g, i64 n, double *__restrict__ out)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e < n) out[e] = fabs(__dsub_rn(f[e], g[e]));
}

// the star of probes around every start point (globalpivotfinder.jl:167-177), expanded on the device:
// probe q = s*star + off[p] + (v-1) is start s with x_p replaced by v
__global__ void k_star_points(const i64 *__restrict__ starts, const i64 *__restrict__ off, int nsites, i64 star,
                              i64 count, i64 *__restrict__ pts)
{
    i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    const i64 s = q / star, r = q % star;
    int p = 0;
    while (p + 1 < nsites && off[p + 1] <= r) ++p;
    i64 *x = pts + q * nsites;
    for (int k = 0; k < nsites; ++k) x[k] = starts[k + s * nsites];
    x[p] = r - off[p] + 1;
}

// the same star evaluated on the fly for an analytic target: no index array (150 MB at config-4 shape), same order of
// operations as tci_target_eval, so the values are bit-identical to evaluating the expanded points
__global__ void k_eval_star(tci_analytic_t t, const i64 *__restrict__ starts, const i64 *__restrict__ off, i64 star,
                            i64 count, double *__restrict__ out)
{
    i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    const i64 s = q / star, r = q % star;
    int p = 0;
    while (p + 1 < t.nsites && off[p + 1] <= r) ++p;
    const i64 v = r - off[p] + 1;
    const i64 *x = starts + s * t.nsites;
    double st[TCI_MAX_STATE];
    tci_target_init(&t, st);
    for (int k = 0; k < t.nsites; ++k) tci_target_accum(&t, k, k == p ? v : x[k], st);
    out[q] = tci_target_finalize(&t, st);
}

// tt value of every probe of the arm of site p from the prefix / suffix environments of the START points:
//   g[s*star + off_p + v] = sum_b W[s + ns*(v + d*b)] * R[b + Dr*s],  W = L_p^T T_p  (ns x d*Dr, from the DMMA GEMM)
__global__ void k_star_contract(const double *__restrict__ W, const double *__restrict__ R, i64 ns, int d, int Dr,
                                i64 star, i64 offp, double *__restrict__ g)
{
    const i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= ns * d) return;
    const i64 s = e % ns;
    const int v = (int)(e / ns);
    const double *w = W + s + ns * (i64)v;
    const double *r = R + (i64)Dr * s;
    double acc = 0.0;
    for (int b = 0; b < Dr; ++b) acc = fma(w[ns * (i64)d * b], r[b], acc);
    g[s * star + offp + v] = acc;
}

// per start: first maximum of |f - g| over its star with strict '>' (globalpivotfinder.jl:170-175): the record is
// (best error, probe index within the star or -1)
struct __align__(16) StarRec {
    double err;
    i64 idx;
};
__global__ void __launch_bounds__(256)
    k_star_argmax(const double *__restrict__ f, const double *__restrict__ g, i64 star, StarRec *__restrict__ rec)
{
    const i64 s = blockIdx.x;
    double best = 0.0;
    i64 bi = -1;
    for (i64 r = threadIdx.x; r < star; r += blockDim.x) {
        const double e = fabs(__dsub_rn(f[s * star + r], g[s * star + r]));
        if (e > best) { // never true for a NaN
            best = e;
            bi = r;
        }
    }
    __shared__ double sb[256];
    __shared__ i64 si[256];
    sb[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            const double ob = sb[threadIdx.x + w];
            const i64 oi = si[threadIdx.x + w];
            // larger error wins; equal errors: the earlier probe (the reference's scan keeps the first maximum)
            if (oi >= 0 && (ob > sb[threadIdx.x] || (ob == sb[threadIdx.x] && (si[threadIdx.x] < 0 || oi < si[threadIdx.x])))) {
                sb[threadIdx.x] = ob;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rec[s].err = sb[0];
        rec[s].idx = si[0];
    }
}

// One member's share of a search: starts [s0, s1) of d_starts (nsites x nsearch, on this member's device); records
// go to rec[s0 .. s1).  mode 1: ordered chain for every probe (bit-identical to evaluate(tt, x)); mode 2: prefix /
// suffix environments of the start points + one GEMM per site (~n x fewer flops, values within rounding of mode 1).
static int gsearch_member(tci_ctx *ctx, TargetDev &t, TargetDev &tt, const i64 *d_starts, i64 s0, i64 s1, int mode,
                          const i64 *d_off, const std::vector<i64> &off, StarRec *rec)
{
    const i64 ns = s1 - s0;
    if (ns <= 0) return TCI_OK;
    const int n = (int)t.nsites;
    const i64 star = off[n], count = star * ns;
    const i64 *starts = d_starts; // this member's block of starts (entry 0 = start s0)
    std::vector<CoreView> cv = views(tt);
    DevBuf<double> d_f(ctx), d_g(ctx);
    DevBuf<i64> d_idx(ctx);
    TCI_CUDA(ctx, d_f.alloc((size_t)count));
    TCI_CUDA(ctx, d_g.alloc((size_t)count));
    int rc = 0;
    const bool need_points = t.kind != 0 || mode == 1;
    if (need_points) {
        TCI_CUDA(ctx, d_idx.alloc((size_t)(count * n)));
        k_star_points<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(starts, d_off, n, star, count, d_idx.p);
        ctx->launches++;
    }
    if (t.kind == 0) {
        k_eval_star<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(t.an, starts, d_off, star, count, d_f.p);
        ctx->launches++;
    } else
        rc = target_eval_dev(ctx, t, d_idx.p, count, d_f.p);
    if (rc) return rc;
    if (mode == 1)
        rc = tt_eval_points(ctx, cv, d_idx.p, count, d_g.p);
    else {
        // prefix environments L[p] (D_p x ns) after sites 0..p-1 and suffix environments R[p] (D_p x ns) over p..n-1
        std::vector<double *> L((size_t)n + 1, nullptr), R((size_t)n + 1, nullptr);
        auto freeall = [&] {
            for (double *q : L) dev_free(ctx, q);
            for (double *q : R) dev_free(ctx, q);
        };
        cudaError_t e = cudaSuccess;
        for (int p = 1; p < n && e == cudaSuccess; ++p) {
            const CoreView &c = cv[p - 1];
            e = dev_alloc(ctx, (void **)&L[p], (size_t)c.Dr * ns * sizeof(double));
            if (e != cudaSuccess) break;
            const i64 total = (i64)c.Dr * ns;
            k_env_left_step<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(L[p - 1], c.p, c.Dl, c.d, c.Dr,
                                                                                     starts, n, p - 1, ns, L[p]);
            ctx->launches++;
        }
        for (int p = n - 1; p >= 1 && e == cudaSuccess; --p) {
            const CoreView &c = cv[p];
            e = dev_alloc(ctx, (void **)&R[p], (size_t)c.Dl * ns * sizeof(double));
            if (e != cudaSuccess) break;
            const i64 total = (i64)c.Dl * ns;
            k_env_right_step<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(p + 1 < n ? R[p + 1] : nullptr, c.p,
                                                                                      c.Dl, c.d, c.Dr, starts, n, p, ns,
                                                                                      R[p]);
            ctx->launches++;
        }
        if (e == cudaSuccess) {
            e = dev_alloc(ctx, (void **)&L[0], (size_t)ns * sizeof(double));
            if (e == cudaSuccess) e = dev_alloc(ctx, (void **)&R[n], (size_t)ns * sizeof(double));
        }
        if (e != cudaSuccess) {
            freeall();
            return tci_fail(ctx, TCI_ERR_CUDA, std::string("global search environments: ") + cudaGetErrorString(e));
        }
        k_fill<<<(unsigned)((ns + 127) / 128), 128, 0, ctx->stream>>>(L[0], ns, 1.0);
        k_fill<<<(unsigned)((ns + 127) / 128), 128, 0, ctx->stream>>>(R[n], ns, 1.0);
        ctx->launches += 2;
        i64 wmax = 0;
        for (int p = 0; p < n; ++p) wmax = std::max<i64>(wmax, (i64)cv[p].d * cv[p].Dr);
        DevBuf<double> W(ctx);
        e = W.alloc((size_t)(ns * wmax));
        if (e != cudaSuccess) {
            freeall();
            return tci_fail(ctx, TCI_ERR_CUDA, std::string("global search workspace: ") + cudaGetErrorString(e));
        }
        for (int p = 0; p < n && !rc; ++p) {
            const CoreView &c = cv[p];
            // W (ns x d*Dr) = L[p]^T (ns x Dl) * T_p (Dl x d*Dr)
            rc = dgemm_dev(ctx, true, false, ns, (i64)c.d * c.Dr, c.Dl, 1.0, L[p], c.Dl, c.p, c.Dl, 0.0, W.p, ns);
            if (rc) break;
            const i64 total = ns * c.d;
            k_star_contract<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(W.p, R[p + 1], ns, c.d, c.Dr, star,
                                                                                     off[p], d_g.p);
            ctx->launches++;
        }
        freeall();
    }
    if (rc) return rc;
    k_star_argmax<<<(unsigned)ns, 256, 0, ctx->stream>>>(d_f.p, d_g.p, star, rec + s0);
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

// start point s, site p of the injected counter generator (util.CounterRNG.start_points, oracle start_points):
// 1 + floor(u * d) with u = tci_uniform01(seed, (call * 1000003 + s) * 1009 + p), clamped to d
TCI_HD i64 counter_start(unsigned long long seed, unsigned long long call, i64 s, i64 p, i64 d)
{
    const unsigned long long idx = (call * 1000003ull + (unsigned long long)s) * 1009ull + (unsigned long long)p;
    const i64 v = 1 + (i64)(tci_uniform01(seed, idx) * (double)d);
    return v < d ? v : d;
}
__global__ void k_counter_starts(unsigned long long seed, unsigned long long call, i64 s0, i64 ns, int n,
                                 const i64 *__restrict__ off, i64 *__restrict__ out)
{
    const i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= ns * n) return;
    const i64 s = e / n, p = e % n;
    out[e] = counter_start(seed, call, s0 + s, p, off[p + 1] - off[p]);
}

static int globalsearch_core(tci_ctx *ctx, int64_t target_id, int64_t tt_id, const int64_t *starts, bool counter,
                             unsigned long long seed, unsigned long long call, int64_t nsearch, double threshold,
                             int64_t maxn, int mode, int64_t *pivots_out, double *errs_out, int64_t *start_idx_out,
                             int64_t *nfound)
{
    TCI_ENTER(ctx);
    if (!nfound) return tci_fail(ctx, TCI_ERR_ARG, "tci_globalsearch: nfound missing");
    *nfound = 0;
    auto it = ctx->targets.find(target_id), jt = ctx->targets.find(tt_id);
    if (it == ctx->targets.end() || jt == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &t = *it->second;
    if (t.is_complex) return tci_fail(ctx, TCI_ERR_ARG, "ComplexF64 target: use the tci_z* entry points");
    TargetDev &tt = *jt->second;
    if (tt.kind != 1) return tci_fail(ctx, TCI_ERR_ARG, "tci_globalsearch: tt_id must name a tensor train (tci_tt_create)");
    if (t.nsites != tt.nsites) return tci_fail(ctx, TCI_ERR_ARG, "tci_globalsearch: tensor train length mismatch");
    if (tt.dl.front() != 1 || tt.dr.back() != 1) return tci_fail(ctx, TCI_ERR_ARG, "boundary bonds must be 1");
    if (mode < 0 || mode > 2) return tci_fail(ctx, TCI_ERR_ARG, "tci_globalsearch: mode is 0 (auto), 1 (ordered chain) or 2 (environments)");
    if (nsearch <= 0 || maxn <= 0) return TCI_OK;
    if ((!starts && !counter) || !pivots_out || !errs_out) return tci_fail(ctx, TCI_ERR_ARG, "tci_globalsearch: buffers missing");
    const i64 nsites = t.nsites;
    std::vector<i64> off((size_t)nsites + 1, 0);
    for (i64 p = 0; p < nsites; ++p) off[p + 1] = off[p] + tt.d[p];
    const i64 star = off[nsites];
    // auto: a handful of starts (the reference's default nsearch = 5) keeps the ordered chain, whose values are
    // bit-identical to evaluate(tt, x); large searches use the prefix / suffix environments (~nsites x fewer flops)
    if (mode == 0) mode = star * nsearch >= 32768 ? 2 : 1;
    if (const char *e = getenv("TCI_GSEARCH_MODE"))
        if (atoi(e) == 1 || atoi(e) == 2) mode = atoi(e);
    tci_group *g = ctx->grp;
    const int world = (g && nsearch >= 4 * (i64)g->world && star * nsearch >= 65536) ? g->world : 1;
    const i64 blk = (nsearch + world - 1) / world;
    const size_t rec_bytes = (size_t)blk * world * sizeof(StarRec);
    StarRec *hrec = static_cast<StarRec *>(ctx_pinned(ctx, rec_bytes));
    if (!hrec) return tci_fail(ctx, TCI_ERR_CUDA, "page-locked staging buffer");
    cudaEventRecord(ctx->ev4, ctx->stream);
    const bool saved = ctx->nosync;
    ctx->nosync = true;
    int rc = 0;
    std::vector<StarRec *> drec(world, nullptr);
    std::vector<i64 *> dst(world, nullptr), doff(world, nullptr);
    auto member_job = [&](int k) -> int {
        tci_ctx *c = world == 1 ? ctx : g->m[k];
        TargetDev &tk = *c->targets.at(target_id);
        TargetDev &ttk = *c->targets.at(tt_id);
        TCI_CUDA(c, dev_alloc(c, (void **)&drec[k], rec_bytes));
        const i64 s0 = std::min(nsearch, c->rank * blk), s1 = std::min(nsearch, (c->rank + 1) * blk);
        // only this member's block of starts: uploaded, or drawn on the device by the counter generator
        TCI_CUDA(c, dev_alloc(c, (void **)&dst[k], (size_t)std::max<i64>(nsites * (s1 - s0), 1) * sizeof(i64)));
        TCI_CUDA(c, dev_alloc(c, (void **)&doff[k], off.size() * sizeof(i64)));
        TCI_CUDA(c, cudaMemcpyAsync(doff[k], off.data(), off.size() * sizeof(i64), cudaMemcpyHostToDevice, c->stream));
        if (s1 > s0) {
            if (counter) {
                const i64 tot = nsites * (s1 - s0);
                k_counter_starts<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(seed, call, s0, s1 - s0, (int)nsites, doff[k],
                                                                                 dst[k]);
                c->launches++;
            } else
                TCI_CUDA(c, cudaMemcpyAsync(dst[k], starts + nsites * s0, (size_t)(nsites * (s1 - s0)) * sizeof(i64),
                                            cudaMemcpyHostToDevice, c->stream));
        }
        return gsearch_member(c, tk, ttk, dst[k], s0, s1, mode, doff[k], off, drec[k]);
    };
    if (world == 1)
        rc = member_job(0);
    else {
        rc = group_run(g, member_job);
        // fixed-size (error, index) records of every rank's starts: one all-gather (globalpivotfinder.jl:186-188 is
        // replayed on the gathered records below)
        if (!rc) rc = group_allgather(g, [&](int k) { return (void *)drec[k]; }, (size_t)blk * sizeof(StarRec));
    }
    ctx->nosync = saved;
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(hrec, drec[0], (size_t)nsearch * sizeof(StarRec), cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(ctx->ev5, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, std::string("tci_globalsearch: ") + cudaGetErrorString(e));
        float ms = 0.f;
        if (!rc && cudaEventElapsedTime(&ms, ctx->ev4, ctx->ev5) == cudaSuccess) ctx->stage_ms[ST_GSEARCH] += ms;
    }
    for (int k = 0; k < world; ++k) {
        tci_ctx *c = world == 1 ? ctx : g->m[k];
        dev_free(c, drec[k]);
        dev_free(c, dst[k]);
        dev_free(c, doff[k]);
    }
    if (rc) return rc;
    std::vector<double> rerr((size_t)nsearch);
    std::vector<i64> ridx((size_t)nsearch), ld((size_t)nsites);
    for (i64 q = 0; q < nsearch; ++q) {
        rerr[q] = hrec[q].err;
        ridx[q] = hrec[q].idx;
    }
    for (i64 p = 0; p < nsites; ++p) ld[p] = tt.d[p];
    std::vector<i64> drawn;
    if (counter) { // the selection needs the start points themselves: the same generator on the host
        drawn.resize((size_t)(nsites * nsearch));
        for (i64 q = 0; q < nsearch; ++q)
            if (rerr[q] > threshold) // only accepted starts are read (tci_globalsearch_select)
                for (i64 p = 0; p < nsites; ++p) drawn[(size_t)(nsites * q + p)] = counter_start(seed, call, q, p, ld[p]);
        starts = drawn.data();
    }
    return tci_globalsearch_select(rerr.data(), ridx.data(), nsearch, starts, nsites, ld.data(), threshold, maxn,
                                   pivots_out, errs_out, start_idx_out, nfound);
}

extern "C" int tci_globalsearch(tci_ctx *ctx, int64_t target_id, int64_t tt_id, const int64_t *starts,
                                int64_t nsearch, double threshold, int64_t maxn, int mode, int64_t *pivots_out,
                                double *errs_out, int64_t *start_idx_out, int64_t *nfound)
{
    return globalsearch_core(ctx, target_id, tt_id, starts, false, 0ull, 0ull, nsearch, threshold, maxn, mode, pivots_out,
                             errs_out, start_idx_out, nfound);
}

// tci_globalsearch with the start points drawn INSIDE the library by the injected counter generator (seed, call): every
// GPU draws its own block on the device, nothing is generated or uploaded by the host.  Start s, site p is
// 1 + floor(tci_uniform01(seed, (call * 1000003 + s) * 1009 + p) * d_p), the sequence util.CounterRNG / the oracle use.
extern "C" int tci_globalsearch_counter(tci_ctx *ctx, int64_t target_id, int64_t tt_id, uint64_t seed, uint64_t call,
                                        int64_t nsearch, double threshold, int64_t maxn, int mode, int64_t *pivots_out,
                                        double *errs_out, int64_t *start_idx_out, int64_t *nfound)
{
    return globalsearch_core(ctx, target_id, tt_id, nullptr, true, seed, call, nsearch, threshold, maxn, mode, pivots_out,
                             errs_out, start_idx_out, nfound);
}

// selection of globalpivotfinder.jl:180-188 on the per-start (error, probe index) records: keep a start if its best
// error exceeds the threshold, in start order, truncated to the first maxn; the pivot is the start point with the
// coordinate of the best probe replaced.  Host only.
extern "C" int tci_globalsearch_select(const double *rec_err, const int64_t *rec_idx, int64_t nsearch,
                                       const int64_t *starts, int64_t nsites, const int64_t *localdims,
                                       double threshold, int64_t maxn, int64_t *pivots_out, double *errs_out,
                                       int64_t *start_idx_out, int64_t *nfound)
{
    if (!nfound || !rec_err || !rec_idx || !starts || !localdims || !pivots_out || !errs_out) return TCI_ERR_ARG;
    std::vector<i64> off((size_t)nsites + 1, 0);
    for (i64 p = 0; p < nsites; ++p) off[p + 1] = off[p] + localdims[p];
    i64 found = 0;
    for (i64 s = 0; s < nsearch && found < maxn; ++s) {
        const double best = rec_err[s];
        if (best > threshold) {
            for (i64 k = 0; k < nsites; ++k) pivots_out[k + found * nsites] = starts[k + s * nsites];
            if (rec_idx[s] >= 0) {
                const i64 r = rec_idx[s];
                i64 p = 0;
                while (p + 1 < nsites && off[p + 1] <= r) ++p;
                pivots_out[p + found * nsites] = r - off[p] + 1;
            }
            errs_out[found] = best;
            if (start_idx_out) start_idx_out[found] = s;
            ++found;
        }
    }
    *nfound = found;
    return TCI_OK;
}

// contiguous blocks whose starts are multiples of `align` (the partition of every sharded stage).  Host only.
extern "C" int tci_shard_range(int64_t n, int world, int rank, int64_t align, int64_t *lo, int64_t *hi)
{
    if (n < 0 || world < 1 || rank < 0 || rank >= world || align < 1 || !lo || !hi) return TCI_ERR_ARG;
    const i64 blk = round_up((n + world - 1) / world, align);
    *lo = std::min<i64>(n, rank * blk);
    *hi = std::min<i64>(n, (rank + 1) * blk);
    return TCI_OK;
}
