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

extern "C" int tci_globalsearch(tci_ctx *ctx, int64_t target_id, int64_t nsites, const int64_t *dims3,
                                const double *const *cores, const int64_t *starts, int64_t nsearch, double threshold,
                                int64_t maxn, int64_t *pivots_out, double *errs_out, int64_t *start_idx_out,
                                int64_t *nfound)
{
    TCI_ENTER(ctx);
    if (!nfound) return tci_fail(ctx, TCI_ERR_ARG, "tci_globalsearch: nfound missing");
    *nfound = 0;
    auto it = ctx->targets.find(target_id);
    if (it == ctx->targets.end()) return tci_fail(ctx, TCI_ERR_ARG, "unknown target id");
    TargetDev &t = *it->second;
    if (t.nsites != nsites) return tci_fail(ctx, TCI_ERR_ARG, "tci_globalsearch: tensor train length mismatch");
    if (nsearch <= 0 || maxn <= 0) return TCI_OK;
    StageTimer tm(ctx, ST_GSEARCH);
    // the star of probes around every start point (globalpivotfinder.jl:167-177)
    i64 star = 0;
    for (i64 p = 0; p < nsites; ++p) star += dims3[3 * p + 1];
    const i64 count = star * nsearch;
    std::vector<i64> pts((size_t)(count * nsites));
    i64 q = 0;
    for (i64 s = 0; s < nsearch; ++s)
        for (i64 p = 0; p < nsites; ++p)
            for (i64 v = 1; v <= dims3[3 * p + 1]; ++v, ++q) {
                i64 *x = pts.data() + q * nsites;
                for (i64 k = 0; k < nsites; ++k) x[k] = starts[k + s * nsites];
                x[p] = v;
            }
    HostTT tt(ctx);
    int rc = tt.upload(nsites, dims3, cores);
    if (rc) return rc;
    DevBuf<i64> d_idx(ctx);
    DevBuf<double> d_f(ctx), d_g(ctx), d_e(ctx);
    TCI_CUDA(ctx, d_idx.upload(pts.data(), pts.size()));
    TCI_CUDA(ctx, d_f.alloc((size_t)count));
    TCI_CUDA(ctx, d_g.alloc((size_t)count));
    TCI_CUDA(ctx, d_e.alloc((size_t)count));
    rc = target_eval_dev(ctx, t, d_idx.p, count, d_f.p);
    if (rc) return rc;
    rc = tt_eval_points(ctx, tt.cv, d_idx.p, count, d_g.p);
    if (rc) return rc;
    k_abs_diff<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(d_f.p, d_g.p, count, d_e.p);
    ctx->launches++;
    std::vector<double> err((size_t)count);
    TCI_CUDA(ctx, cudaMemcpyAsync(err.data(), d_e.p, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // selection: strict '>' keeps the first maximum, threshold, truncation in start order (:170-188)
    i64 found = 0;
    for (i64 s = 0; s < nsearch && found < maxn; ++s) {
        double best = 0.0;
        i64 bestq = -1;
        for (i64 e = s * star; e < (s + 1) * star; ++e)
            if (err[e] > best) {
                best = err[e];
                bestq = e;
            }
        if (best > threshold) {
            const i64 *x = bestq >= 0 ? pts.data() + bestq * nsites : starts + s * nsites;
            for (i64 k = 0; k < nsites; ++k) pivots_out[k + found * nsites] = x[k];
            errs_out[found] = best;
            if (start_idx_out) start_idx_out[found] = s;
            ++found;
        }
    }
    *nfound = found;
    return TCI_OK;
}
