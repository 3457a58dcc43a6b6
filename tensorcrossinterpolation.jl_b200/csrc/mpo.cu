// mpo.cu -- K5: MPO x MPO product as a target (Contraction, contraction.jl:5-335) and
// K6: site contractions of contract_zipup / contract_naive (contraction.jl:338-349,455-464).
//
// Everything is expressed as strided batched DGEMMs on views of the cores as stored
// (no permutedims copies as in _contract, contraction.jl:71-93):
//   A[:, i, :, :]  = (La x S*La')  matrix at A + La*i,      ld = La*d1
//   B[:, :, j, :]  = (Lb*S x Lb')  matrix at B + Lb*S*j,    ld = Lb*S*d3
// A per-point environment is the (La x Lb) matrix env[a + La*(b + Lb*q)].
#include "tci_internal.h"

struct MpoSite {
    const double *A, *B;
    i64 La, d1, S, Lan; // A: (La, d1, S, Lan)
    i64 Lb, d3, Lbn;    // B: (Lb, S, d3, Lbn)
};

static MpoSite site_of(const TargetDev &t, i64 s)
{
    return MpoSite{t.A[s], t.B[s], t.adl[s], t.as1[s], t.as2[s], t.adr[s], t.bdl[s], t.bs2[s], t.bdr[s]};
}

// unfuse idx = i + d1*(j-1) (contraction.jl:95-101) into element offsets of the two views
__global__ void k_mpo_offsets(const i64 *__restrict__ idx, int len, int pos, i64 count, i64 d1, i64 mulA, i64 mulB,
                              i64 *__restrict__ offA, i64 *__restrict__ offB)
{
    i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    i64 f = idx[(i64)len * q + pos] - 1;
    offA[q] = mulA * (f % d1);
    offB[q] = mulB * (f / d1);
}

__global__ void k_fill1(double *p, i64 n, double v)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e < n) p[e] = v;
}

// evaluateleft (contraction.jl:112-139) for `count` points over sites [0, nsteps)
static int mpo_left_chain(tci_ctx *ctx, const TargetDev &t, int nsteps, const i64 *d_idx, int len, int off, i64 count,
                          double **out, i64 *La_out, i64 *Lb_out)
{
    double *env = nullptr;
    TCI_CUDA(ctx, dev_alloc(ctx, (void **)&env, (size_t)count * sizeof(double)));
    k_fill1<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(env, count, 1.0);
    ctx->launches++;
    i64 La = 1, Lb = 1;
    DevBuf<i64> offA(ctx), offB(ctx);
    TCI_CUDA(ctx, offA.alloc((size_t)count));
    TCI_CUDA(ctx, offB.alloc((size_t)count));
    for (int s = 0; s < nsteps; ++s) {
        MpoSite m = site_of(t, s);
        k_mpo_offsets<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(d_idx, len, s + off, count, m.d1, m.La,
                                                                               m.Lb * m.S, offA.p, offB.p);
        ctx->launches++;
        double *tmp = nullptr, *nxt = nullptr;
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&tmp, (size_t)(m.Lb * m.S * m.Lan) * count * sizeof(double)));
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)(m.Lan * m.Lbn) * count * sizeof(double)));
        // tmp[q] (Lb x S*Lan) = env[q]^T (Lb x La) * A_i (La x S*Lan)       contraction.jl:105
        int rc = dgemm_dev_batched_off(ctx, true, false, m.Lb, m.S * m.Lan, m.La, 1.0, env, m.La, m.La * m.Lb, m.A,
                                       m.La * m.d1, 0, 0.0, tmp, m.Lb, m.Lb * m.S * m.Lan, count, nullptr, offA.p,
                                       m.La % 2 == 0);
        // nxt[q] (Lan x Lbn) = tmp[q]^T (Lan x Lb*S) * B_j (Lb*S x Lbn)       contraction.jl:108
        if (!rc)
            rc = dgemm_dev_batched_off(ctx, true, false, m.Lan, m.Lbn, m.Lb * m.S, 1.0, tmp, m.Lb * m.S,
                                       m.Lb * m.S * m.Lan, m.B, m.Lb * m.S * m.d3, 0, 0.0, nxt, m.Lan, m.Lan * m.Lbn,
                                       count, nullptr, offB.p, (m.Lb * m.S) % 2 == 0);
        dev_free(ctx, tmp);
        dev_free(ctx, env);
        env = nxt;
        La = m.Lan;
        Lb = m.Lbn;
        if (rc) {
            dev_free(ctx, env);
            return rc;
        }
    }
    *out = env;
    *La_out = La;
    *Lb_out = Lb;
    return TCI_OK;
}

// evaluateright (contraction.jl:144-176) over the last nsteps sites
static int mpo_right_chain(tci_ctx *ctx, const TargetDev &t, int nsteps, const i64 *d_idx, int len, int off, i64 count,
                           double **out, i64 *La_out, i64 *Lb_out)
{
    const int N = (int)t.nsites;
    double *env = nullptr;
    TCI_CUDA(ctx, dev_alloc(ctx, (void **)&env, (size_t)count * sizeof(double)));
    k_fill1<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(env, count, 1.0);
    ctx->launches++;
    i64 Ra = 1, Rb = 1;
    DevBuf<i64> offA(ctx), offB(ctx);
    TCI_CUDA(ctx, offA.alloc((size_t)count));
    TCI_CUDA(ctx, offB.alloc((size_t)count));
    for (int s = N - 1; s >= N - nsteps; --s) {
        MpoSite m = site_of(t, s);
        k_mpo_offsets<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
            d_idx, len, s - (N - nsteps) + off, count, m.d1, m.La, m.Lb * m.S, offA.p, offB.p);
        ctx->launches++;
        double *tmp = nullptr, *nxt = nullptr;
        // tmp[q][br + Lbn*(h + S*al)] = sum_ar env[ar, br] * A[al, i, h, ar]
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&tmp, (size_t)(m.Lbn * m.S * m.La) * count * sizeof(double)));
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)(m.La * m.Lb) * count * sizeof(double)));
        int rc = 0;
        for (i64 h = 0; h < m.S && !rc; ++h) // (Lbn x La) = env^T (Lbn x Lan) * A_{i,h}^T (Lan x La)
            rc = dgemm_dev_batched_off(ctx, true, true, m.Lbn, m.La, m.Lan, 1.0, env, m.Lan, m.Lan * m.Lbn,
                                       m.A + m.La * m.d1 * h, m.La * m.d1 * m.S, 0, 0.0, tmp + m.Lbn * h, m.Lbn * m.S,
                                       m.Lbn * m.S * m.La, count, nullptr, offA.p, m.La % 2 == 0);
        // nxt[q][al, bl] = sum_{br,h} tmp[br, h, al] * B[bl, h, j, br]
        for (i64 h = 0; h < m.S && !rc; ++h) // (La x Lb) += tmp_h^T (La x Lbn) * B_{j,h}^T (Lbn x Lb)
            rc = dgemm_dev_batched_off(ctx, true, true, m.La, m.Lb, m.Lbn, 1.0, tmp + m.Lbn * h, m.Lbn * m.S,
                                       m.Lbn * m.S * m.La, m.B + m.Lb * h, m.Lb * m.S * m.d3, 0, h ? 1.0 : 0.0, nxt,
                                       m.La, m.La * m.Lb, count, nullptr, offB.p, (m.Lb * m.S) % 2 == 0);
        dev_free(ctx, tmp);
        dev_free(ctx, env);
        env = nxt;
        Ra = m.La;
        Rb = m.Lb;
        if (rc) {
            dev_free(ctx, env);
            return rc;
        }
    }
    *out = env;
    *La_out = Ra;
    *Lb_out = Rb;
    return TCI_OK;
}

// environments alone (tci_env_eval): evaluateleft / evaluateright of contraction.jl:112-176 for `count` entries
int env_eval_mpo(tci_ctx *ctx, TargetDev &t, int side, const i64 *d_idx, int len, i64 count, double **out, i64 *D)
{
    i64 a = 1, b = 1;
    int rc = side == 0 ? mpo_left_chain(ctx, t, len, d_idx, len, 0, count, out, &a, &b)
                       : mpo_right_chain(ctx, t, len, d_idx, len, 0, count, out, &a, &b);
    *D = a * b;
    return rc;
}

// batchevaluate(::Contraction) contraction.jl:236-335 (projector = nothing, f = nothing)
int pi_eval_mpo(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                tci_dmat *out)
{
    double *X = nullptr, *right = nullptr;
    i64 La = 1, Lb = 1, Ra = 1, Rb = 1;
    int rc = mpo_left_chain(ctx, t, (int)nl, dI, (int)nl, 0, nI, &X, &La, &Lb);
    if (rc) return rc;
    rc = mpo_right_chain(ctx, t, (int)nr, dJ, (int)nr, 0, nJ, &right, &Ra, &Rb);
    if (rc) {
        dev_free(ctx, X);
        return rc;
    }
    // X: (La x Lb x R), r = i + nI*sacc ; centre sites :290-317
    i64 R = nI;
    for (i64 s = nl; s < nl + M && !rc; ++s) {
        MpoSite m = site_of(t, s);
        const i64 N1 = m.d1 * m.S * m.Lan;
        double *tmp = nullptr, *nxt = nullptr;
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&tmp, (size_t)(m.Lb * R * N1) * sizeof(double)));
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)(m.Lan * m.Lbn * R * m.d1 * m.d3) * sizeof(double)));
        // tmp[(b + Lb*r) + Lb*R*(x + d1*(h + S*an))] = sum_a X[a, b, r] * A[a, x, h, an]
        rc = dgemm_dev(ctx, true, false, m.Lb * R, N1, m.La, 1.0, X, m.La, m.A, m.La, 0.0, tmp, m.Lb * R);
        // nxt[an + Lan*(bn + Lbn*(r + R*(x + d1*z)))] = sum_{b,h} tmp[b, r, x, h, an] * B[b, h, z, bn]
        for (i64 z = 0; z < m.d3 && !rc; ++z)
            for (i64 x = 0; x < m.d1 && !rc; ++x)
                for (i64 h = 0; h < m.S && !rc; ++h)
                    rc = dgemm_dev_batched(ctx, true, false, m.Lan, m.Lbn, m.Lb, 1.0,
                                           tmp + m.Lb * R * (x + m.d1 * h), m.Lb * R * m.d1 * m.S, m.Lb,
                                           m.B + m.Lb * (h + m.S * z), m.Lb * m.S * m.d3, 0, h ? 1.0 : 0.0,
                                           nxt + m.Lan * m.Lbn * R * (x + m.d1 * z), m.Lan, m.Lan * m.Lbn, R);
        dev_free(ctx, tmp);
        dev_free(ctx, X);
        X = nxt;
        La = m.Lan;
        Lb = m.Lbn;
        R *= m.d1 * m.d3;
    }
    // res[r, j] = sum_{a,b} X[a, b, r] * right[a, b, j]   :328
    if (!rc) rc = dgemm_dev(ctx, true, false, R, nJ, La * Lb, 1.0, X, La * Lb, right, Ra * Rb, 0.0, out->p, out->ld);
    dev_free(ctx, X);
    dev_free(ctx, right);
    return rc;
}

__global__ void k_dot_env(const double *__restrict__ l, const double *__restrict__ r, i64 D, i64 count,
                          double *__restrict__ out)
{
    i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    double acc = 0.0;
    for (i64 a = 0; a < D; ++a) acc = __dadd_rn(acc, __dmul_rn(l[a + D * q], r[a + D * q]));
    out[q] = acc;
}

// evaluate(::Contraction, indexset)  contraction.jl:189-207
int target_eval_mpo(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out)
{
    const int N = (int)t.nsites, mid = N / 2;
    double *l = nullptr, *r = nullptr;
    i64 La, Lb, Ra, Rb;
    int rc = mpo_left_chain(ctx, t, mid, d_idx, N, 0, count, &l, &La, &Lb);
    if (rc) return rc;
    rc = mpo_right_chain(ctx, t, N - mid, d_idx, N, mid, count, &r, &Ra, &Rb);
    if (!rc) {
        k_dot_env<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(l, r, La * Lb, count, d_out);
        ctx->launches++;
    }
    dev_free(ctx, l);
    dev_free(ctx, r);
    return rc;
}

// ---- K6 -------------------------------------------------------------------------
extern "C" int tci_contract_zipup_site(tci_ctx *ctx, const double *R, int64_t chi, int64_t Da, int64_t Db,
                                       const double *A, int64_t s1, int64_t s2, int64_t Dan, const double *B,
                                       int64_t s3, int64_t Dbn, double *C_host, tci_dmat **C_dev)
{
    TCI_ENTER(ctx);
    if (C_dev) *C_dev = nullptr;
    if (chi < 1 || Da < 1 || Db < 1 || s1 < 1 || s2 < 1 || s3 < 1 || Dan < 1 || Dbn < 1 || !R || !A || !B)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_contract_zipup_site: bad arguments");
    DevBuf<double> dR(ctx), dA(ctx), dB(ctx), RA(ctx);
    {
        StageTimer tm(ctx, ST_H2D);
        TCI_CUDA(ctx, dR.upload(R, (size_t)(chi * Da * Db)));
        TCI_CUDA(ctx, dA.upload(A, (size_t)(Da * s1 * s2 * Dan)));
        TCI_CUDA(ctx, dB.upload(B, (size_t)(Db * s2 * s3 * Dbn)));
    }
    tci_dmat *C = nullptr;
    int rc = dmat_alloc(ctx, chi * s1 * s3, Dan * Dbn, &C);
    if (rc) return rc;
    {
        StageTimer tm(ctx, ST_GEMM);
        // RA2[(c + chi*x) + chi*s1*((b + Db*h) + Db*s2*an)] = sum_a R[c, a, b] * A[a, x, h, an]   :458
        TCI_CUDA(ctx, RA.alloc((size_t)(chi * s1 * Db * s2 * Dan)));
        for (i64 h = 0; h < s2 && !rc; ++h)
            for (i64 x = 0; x < s1 && !rc; ++x)
                rc = dgemm_dev_batched(ctx, false, false, chi, Dan, Da, 1.0, dR.p, chi, chi * Da,
                                       dA.p + Da * (x + s1 * h), Da * s1 * s2, 0, 0.0,
                                       RA.p + chi * x + chi * s1 * Db * h, chi * s1 * Db * s2, chi * s1, Db);
        // C[(c + chi*(x + s1*z)) + chi*s1*s3*(an + Dan*bn)] = sum_{b,h} RA2[(c,x),(b,h),an] * B[b, h, z, bn]  :464
        for (i64 z = 0; z < s3 && !rc; ++z)
            rc = dgemm_dev_batched(ctx, false, false, chi * s1, Dbn, Db * s2, 1.0, RA.p, chi * s1, chi * s1 * Db * s2,
                                   dB.p + Db * s2 * z, Db * s2 * s3, 0, 0.0, C->p + chi * s1 * z, C->ld * Dan,
                                   C->ld, Dan);
    }
    if (rc) {
        tci_dmat_destroy(C);
        return rc;
    }
    if (C_host) {
        StageTimer tm(ctx, ST_D2H);
        TCI_CUDA(ctx, cudaMemcpy2DAsync(C_host, C->m * sizeof(double), C->p, C->ld * sizeof(double),
                                        C->m * sizeof(double), C->n, cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (C_dev)
        *C_dev = C;
    else
        tci_dmat_destroy(C);
    return TCI_OK;
}

// out[(la + Da*lb), x, z, (lan + Dan*lbn)] = sum_h A[la, x, h, lan] * B[lb, h, z, lbn]   contraction.jl:338-349
__global__ void k_naive_site(const double *__restrict__ A, const double *__restrict__ B, i64 Da, i64 s1, i64 s2,
                             i64 Dan, i64 Db, i64 s3, i64 Dbn, double *__restrict__ out)
{
    i64 total = Da * Db * s1 * s3 * Dan * Dbn;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 rem = e;
        i64 la = rem % Da;
        rem /= Da;
        i64 lb = rem % Db;
        rem /= Db;
        i64 x = rem % s1;
        rem /= s1;
        i64 z = rem % s3;
        rem /= s3;
        i64 lan = rem % Dan, lbn = rem / Dan;
        double acc = 0.0;
        for (i64 h = 0; h < s2; ++h)
            acc = fma(A[la + Da * (x + s1 * (h + s2 * lan))], B[lb + Db * (h + s2 * (z + s3 * lbn))], acc);
        out[e] = acc;
    }
}

extern "C" int tci_contract_naive_site(tci_ctx *ctx, const double *A, int64_t Da, int64_t s1, int64_t s2, int64_t Dan,
                                       const double *B, int64_t Db, int64_t s3, int64_t Dbn, double *out_host)
{
    TCI_ENTER(ctx);
    if (!A || !B || !out_host) return tci_fail(ctx, TCI_ERR_ARG, "tci_contract_naive_site: bad arguments");
    const i64 total = Da * Db * s1 * s3 * Dan * Dbn;
    DevBuf<double> dA(ctx), dB(ctx), dO(ctx);
    TCI_CUDA(ctx, dA.upload(A, (size_t)(Da * s1 * s2 * Dan)));
    TCI_CUDA(ctx, dB.upload(B, (size_t)(Db * s2 * s3 * Dbn)));
    TCI_CUDA(ctx, dO.alloc((size_t)total));
    unsigned blocks = (unsigned)std::min<i64>((total + 255) / 256, (i64)ctx->sm_count * 16);
    k_naive_site<<<blocks, 256, 0, ctx->stream>>>(dA.p, dB.p, Da, s1, s2, Dan, Db, s3, Dbn, dO.p);
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    TCI_CUDA(ctx, cudaMemcpyAsync(out_host, dO.p, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}
