// mpo.cu -- K5: MPO x MPO product as a target (Contraction, contraction.jl:5-335) and
// K6: site contractions of contract_zipup / contract_naive (contraction.jl:338-349,455-464).
//
// Everything is expressed as strided batched DGEMMs on views of the cores as stored
// (no permutedims copies as in _contract, contraction.jl:71-93):
//   A[:, i, :, :]  = (La x S*La')  matrix at A + La*i,      ld = La*d1
//   B[:, :, j, :]  = (Lb*S x Lb')  matrix at B + Lb*S*j,    ld = Lb*S*d3
// A per-point environment is the (La x Lb) matrix env[a + La*(b + Lb*q)].
#include "tci_internal.h"

#include <algorithm>

// The chains are templates over the value type V: double (Float64 targets) or double2 = interleaved (re, im) pairs
// (ComplexF64 targets; the reference's contraction tests are complex, test_contraction.jl:39-46).
// Leading dimensions, strides and offsets are counted in elements of V.
template <class V> struct MpoSiteT {
    const V *A, *B;
    i64 La, d1, S, Lan; // A: (La, d1, S, Lan)
    i64 Lb, d3, Lbn;    // B: (Lb, S, d3, Lbn)
};

template <class V> static MpoSiteT<V> site_of(const TargetDev &t, i64 s)
{
    return MpoSiteT<V>{(const V *)t.A[s], (const V *)t.B[s], t.adl[s], t.as1[s], t.as2[s], t.adr[s], t.bdl[s], t.bs2[s], t.bdr[s]};
}

int zgemm_dev_batched_off(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double2 *A, i64 lda,
                          i64 strideA, const double2 *B, i64 ldb, i64 strideB, double beta, double2 *C, i64 ldc,
                          i64 strideC, i64 batch, const i64 *offA, const i64 *offB); // zgemm.cu

// Bp[bl, br, h, j] = B[bl, h, j, br]: with this copy the second product of a right-environment step,
//   nxt[al, bl] = sum_{br, h} tmp[(br, h), al] * B[bl, h, j, br]      (contraction.jl:144-176),
// is ONE GEMM with inner dimension Lbn * S instead of S accumulating GEMMs (the reference permutes both cores too, with
// permutedims on every extension: contraction.jl:170-171).
template <class V>
__global__ void k_permute_b(const V *__restrict__ B, i64 Lb, i64 S, i64 d3, i64 Lbn, V *__restrict__ Bp)
{
    const i64 total = Lb * S * d3 * Lbn;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 bl = e % Lb, br = (e / Lb) % Lbn, h = (e / (Lb * Lbn)) % S, j = e / (Lb * Lbn * S);
        Bp[e] = B[bl + Lb * (h + S * (j + d3 * br))];
    }
}
template <class V> static int ensure_bperm(tci_ctx *ctx, TargetDev &t)
{
    if (!t.Bp.empty()) return TCI_OK;
    std::vector<double *> bp((size_t)t.nsites, nullptr);
    for (i64 s = 0; s < t.nsites; ++s) {
        const i64 total = t.bdl[s] * t.bs1[s] * t.bs2[s] * t.bdr[s];
        cudaError_t e = cudaMalloc(&bp[s], (size_t)std::max<i64>(total, 1) * sizeof(V));
        if (e != cudaSuccess) {
            for (double *p : bp) cudaFree(p);
            return tci_fail(ctx, TCI_ERR_CUDA, std::string("permuted MPO cores: ") + cudaGetErrorString(e));
        }
        k_permute_b<V><<<(unsigned)std::min<i64>((total + 255) / 256, 1024), 256, 0, ctx->stream>>>(
            (const V *)t.B[s], t.bdl[s], t.bs1[s], t.bs2[s], t.bdr[s], (V *)bp[s]);
        ctx->launches++;
    }
    t.Bp = bp;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}
static inline int gemm_off(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
                           i64 sA, const double *B, i64 ldb, i64 sB, double beta, double *C, i64 ldc, i64 sC, i64 batch,
                           const i64 *offA, const i64 *offB, bool even)
{
    return dgemm_dev_batched_off(ctx, tA, tB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, offA, offB, even);
}
static inline int gemm_off(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double2 *A, i64 lda,
                           i64 sA, const double2 *B, i64 ldb, i64 sB, double beta, double2 *C, i64 ldc, i64 sC, i64 batch,
                           const i64 *offA, const i64 *offB, bool)
{
    return zgemm_dev_batched_off(ctx, tA, tB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, offA, offB);
}
template <class V> __host__ __device__ inline V v_one();
template <> __host__ __device__ inline double v_one<double>() { return 1.0; }
template <> __host__ __device__ inline double2 v_one<double2>() { return make_double2(1.0, 0.0); }

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

template <class V> __global__ void k_fill1(V *p, i64 n, V v)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e < n) p[e] = v;
}

// ---- shared prefixes -----------------------------------------------------------------------------------------
// The reference memoises environments in a Dict keyed by the partial index (contraction.jl:112-176): entries of an
// index set that share a prefix (left) or suffix (right) share the environment.  The index sets of a TCI run are
// nested (Icombined = kronecker(Iset, d), Iset itself grown site by site), so at chi = 256, d = 4 only a quarter of
// the chain steps are distinct.  ChainPlan finds, level by level, the distinct partial indices of ONE call on the
// host (direct-address table on (parent id, sigma)); the chain then runs one batched GEMM pair per level over the
// distinct entries only, reading the parent's environment through a per-batch offset, and a gather expands the last
// level to the entries of the index set.  (The Dict also persists across calls; this does not.)
struct ChainPlan {
    int nsteps = 0;
    std::vector<i64> cnt;                 // distinct entries per level
    std::vector<std::vector<i64>> parent; // per level: id of the parent entry at the previous level
    std::vector<std::vector<i64>> sigma;  // per level: fused site index (0-based)
    std::vector<i64> last_id;             // entry of the index set -> id at the last level
    bool identity = true;                 // last_id[q] == q: no gather needed after the last level
    bool shared = false;                  // some level has fewer distinct entries than the index set
};

// pos(k) = position inside an index entry read at level k: left chain off + k, right chain off + nsteps - 1 - k
static bool plan_chain(const i64 *h_idx, int len, int off, int nsteps, i64 count, bool right, const i64 *dims,
                       ChainPlan &P)
{
    if (!h_idx || nsteps <= 0 || count < 2 || count > (i64)1 << 22) return false;
    P.nsteps = nsteps;
    std::vector<i64> id((size_t)count, 0), nid((size_t)count);
    i64 nprev = 1;
    std::vector<i64> table;
    for (int k = 0; k < nsteps; ++k) {
        const int pos = right ? off + nsteps - 1 - k : off + k;
        const i64 d = dims[k];
        if (d <= 0 || nprev > ((i64)1 << 22) / d) return false; // table too large: fall back to the plain chain
        table.assign((size_t)(nprev * d), -1);
        std::vector<i64> par, sig;
        for (i64 q = 0; q < count; ++q) {
            const i64 f = h_idx[(i64)len * q + pos] - 1;
            if (f < 0 || f >= d) return false; // out-of-range index: let the plain path deal with it
            i64 &slot = table[(size_t)(id[q] * d + f)];
            if (slot < 0) {
                slot = (i64)par.size();
                par.push_back(id[q]);
                sig.push_back(f);
            }
            nid[q] = slot;
        }
        id.swap(nid);
        nprev = (i64)par.size();
        if (nprev < count) P.shared = true;
        P.cnt.push_back(nprev);
        P.parent.push_back(std::move(par));
        P.sigma.push_back(std::move(sig));
    }
    P.identity = nprev == count;
    if (P.identity)
        for (i64 q = 0; q < count; ++q)
            if (id[q] != q) {
                P.identity = false;
                break;
            }
    P.last_id.swap(id);
    return true;
}

template <class V>
__global__ void k_gather_cols(const V *__restrict__ src, i64 D, const i64 *__restrict__ ids, i64 count,
                              V *__restrict__ dst)
{
    const i64 total = D * count;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x)
        dst[e] = src[(e % D) + D * ids[e / D]];
}

// chain over the distinct entries of a plan; `right` selects the step formulas of mpo_right_chain
template <class V>
static int mpo_chain_planned(tci_ctx *ctx, const TargetDev &t, const ChainPlan &P, bool right, i64 count, V **out,
                             i64 *Da_out, i64 *Db_out)
{
    const int N = (int)t.nsites, nsteps = P.nsteps;
    // one upload: per level [env offset of the parent][offset into A][offset into B]
    std::vector<i64> host;
    std::vector<size_t> base((size_t)nsteps);
    {
        i64 ea = 1, eb = 1; // dimensions of the environment that enters the level
        for (int k = 0; k < nsteps; ++k) {
            MpoSiteT<V> m = site_of<V>(t, right ? N - 1 - k : k);
            const i64 c = P.cnt[k];
            base[k] = host.size();
            host.resize(host.size() + 3 * (size_t)c);
            i64 *g = host.data() + base[k], *oa = g + c, *ob = oa + c;
            for (i64 u = 0; u < c; ++u) {
                const i64 f = P.sigma[k][u];
                g[u] = P.parent[k][u] * ea * eb;
                oa[u] = m.La * (f % m.d1);
                ob[u] = (right ? m.Lb * m.Lbn * m.S : m.Lb * m.S) * (f / m.d1); // right chains read the permuted copy
            }
            ea = right ? m.La : m.Lan;
            eb = right ? m.Lb : m.Lbn;
        }
    }
    DevBuf<i64> dev(ctx), ids(ctx);
    TCI_CUDA(ctx, dev.upload(host.data(), host.size()));
    V *env = nullptr;
    TCI_CUDA(ctx, dev_alloc(ctx, (void **)&env, sizeof(V)));
    k_fill1<V><<<1, 32, 0, ctx->stream>>>(env, 1, v_one<V>());
    ctx->launches++;
    i64 Ea = 1, Eb = 1;
    for (int k = 0; k < nsteps; ++k) {
        MpoSiteT<V> m = site_of<V>(t, right ? N - 1 - k : k);
        const i64 c = P.cnt[k];
        const i64 *g = dev.p + base[k], *oa = g + c, *ob = oa + c;
        V *tmp = nullptr, *nxt = nullptr;
        int rc = 0;
        if (!right) {
            TCI_CUDA(ctx, dev_alloc(ctx, (void **)&tmp, (size_t)(m.Lb * m.S * m.Lan) * c * sizeof(V)));
            TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)(m.Lan * m.Lbn) * c * sizeof(V)));
            // the two products of mpo_left_chain, the environment read at the parent's offset
            rc = gemm_off(ctx, true, false, m.Lb, m.S * m.Lan, m.La, 1.0, env, m.La, 0, m.A, m.La * m.d1, 0,
                                       0.0, tmp, m.Lb, m.Lb * m.S * m.Lan, c, g, oa, m.La % 2 == 0);
            if (!rc)
                rc = gemm_off(ctx, true, false, m.Lan, m.Lbn, m.Lb * m.S, 1.0, tmp, m.Lb * m.S,
                                           m.Lb * m.S * m.Lan, m.B, m.Lb * m.S * m.d3, 0, 0.0, nxt, m.Lan,
                                           m.Lan * m.Lbn, c, nullptr, ob, (m.Lb * m.S) % 2 == 0);
            Ea = m.Lan;
            Eb = m.Lbn;
        } else {
            TCI_CUDA(ctx, dev_alloc(ctx, (void **)&tmp, (size_t)(m.Lbn * m.S * m.La) * c * sizeof(V)));
            TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)(m.La * m.Lb) * c * sizeof(V)));
            for (i64 h = 0; h < m.S && !rc; ++h)
                rc = gemm_off(ctx, true, true, m.Lbn, m.La, m.Lan, 1.0, env, m.Lan, 0,
                                           m.A + m.La * m.d1 * h, m.La * m.d1 * m.S, 0, 0.0, tmp + m.Lbn * h,
                                           m.Lbn * m.S, m.Lbn * m.S * m.La, c, g, oa,
                                           m.La % 2 == 0 && (m.Lan * m.Lbn) % 2 == 0);
            if (!rc) // (La x Lb) = tmp^T (La x Lbn*S) * Bp_j^T (Lbn*S x Lb)
                rc = gemm_off(ctx, true, true, m.La, m.Lb, m.Lbn * m.S, 1.0, tmp, m.Lbn * m.S, m.Lbn * m.S * m.La,
                              (const V *)t.Bp[right ? N - 1 - k : k], m.Lb, 0, 0.0, nxt, m.La, m.La * m.Lb, c, nullptr, ob,
                              (m.Lb * m.Lbn * m.S) % 2 == 0);
            Ea = m.La;
            Eb = m.Lb;
        }
        dev_free(ctx, tmp);
        dev_free(ctx, env);
        env = nxt;
        if (rc) {
            dev_free(ctx, env);
            return rc;
        }
    }
    if (!P.identity) { // expand the distinct environments of the last level to the entries of the index set
        V *full = nullptr;
        const i64 D = Ea * Eb;
        TCI_CUDA(ctx, ids.upload(P.last_id.data(), P.last_id.size()));
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&full, (size_t)(D * count) * sizeof(V)));
        const i64 total = D * count;
        k_gather_cols<V><<<(unsigned)std::min<i64>((total + 255) / 256, (i64)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
            env, D, ids.p, count, full);
        ctx->launches++;
        dev_free(ctx, env);
        env = full;
    }
    TCI_CUDA(ctx, cudaGetLastError());
    *out = env;
    *Da_out = Ea;
    *Db_out = Eb;
    return TCI_OK;
}

// evaluateleft (contraction.jl:112-139) for `count` points over sites [0, nsteps)
template <class V>
static int mpo_left_chain(tci_ctx *ctx, const TargetDev &t, int nsteps, const i64 *d_idx, int len, int off, i64 count,
                          V **out, i64 *La_out, i64 *Lb_out, const i64 *h_idx = nullptr)
{
    if (h_idx && !getenv("TCI_MPO_NO_DEDUP")) {
        ChainPlan P;
        std::vector<i64> dims((size_t)std::max(nsteps, 0));
        for (int k = 0; k < nsteps; ++k) dims[k] = t.localdims[k];
        if (plan_chain(h_idx, len, off, nsteps, count, false, dims.data(), P) && P.shared)
            return mpo_chain_planned<V>(ctx, t, P, false, count, out, La_out, Lb_out);
    }
    V *env = nullptr;
    TCI_CUDA(ctx, dev_alloc(ctx, (void **)&env, (size_t)count * sizeof(V)));
    k_fill1<V><<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(env, count, v_one<V>());
    ctx->launches++;
    i64 La = 1, Lb = 1;
    DevBuf<i64> offA(ctx), offB(ctx);
    TCI_CUDA(ctx, offA.alloc((size_t)count));
    TCI_CUDA(ctx, offB.alloc((size_t)count));
    for (int s = 0; s < nsteps; ++s) {
        MpoSiteT<V> m = site_of<V>(t, s);
        k_mpo_offsets<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(d_idx, len, s + off, count, m.d1, m.La,
                                                                               m.Lb * m.S, offA.p, offB.p);
        ctx->launches++;
        V *tmp = nullptr, *nxt = nullptr;
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&tmp, (size_t)(m.Lb * m.S * m.Lan) * count * sizeof(V)));
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)(m.Lan * m.Lbn) * count * sizeof(V)));
        // tmp[q] (Lb x S*Lan) = env[q]^T (Lb x La) * A_i (La x S*Lan)       contraction.jl:105
        int rc = gemm_off(ctx, true, false, m.Lb, m.S * m.Lan, m.La, 1.0, env, m.La, m.La * m.Lb, m.A,
                                       m.La * m.d1, 0, 0.0, tmp, m.Lb, m.Lb * m.S * m.Lan, count, nullptr, offA.p,
                                       m.La % 2 == 0);
        // nxt[q] (Lan x Lbn) = tmp[q]^T (Lan x Lb*S) * B_j (Lb*S x Lbn)       contraction.jl:108
        if (!rc)
            rc = gemm_off(ctx, true, false, m.Lan, m.Lbn, m.Lb * m.S, 1.0, tmp, m.Lb * m.S,
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
template <class V>
static int mpo_right_chain(tci_ctx *ctx, const TargetDev &t, int nsteps, const i64 *d_idx, int len, int off, i64 count,
                           V **out, i64 *La_out, i64 *Lb_out, const i64 *h_idx = nullptr)
{
    const int N = (int)t.nsites;
    if (int rcp = ensure_bperm<V>(ctx, const_cast<TargetDev &>(t))) return rcp; // a cache of the target, made once
    if (h_idx && !getenv("TCI_MPO_NO_DEDUP")) {
        ChainPlan P;
        std::vector<i64> dims((size_t)std::max(nsteps, 0));
        for (int k = 0; k < nsteps; ++k) dims[k] = t.localdims[N - 1 - k];
        if (plan_chain(h_idx, len, off, nsteps, count, true, dims.data(), P) && P.shared)
            return mpo_chain_planned<V>(ctx, t, P, true, count, out, La_out, Lb_out);
    }
    V *env = nullptr;
    TCI_CUDA(ctx, dev_alloc(ctx, (void **)&env, (size_t)count * sizeof(V)));
    k_fill1<V><<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(env, count, v_one<V>());
    ctx->launches++;
    i64 Ra = 1, Rb = 1;
    DevBuf<i64> offA(ctx), offB(ctx);
    TCI_CUDA(ctx, offA.alloc((size_t)count));
    TCI_CUDA(ctx, offB.alloc((size_t)count));
    for (int s = N - 1; s >= N - nsteps; --s) {
        MpoSiteT<V> m = site_of<V>(t, s);
        k_mpo_offsets<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(
            d_idx, len, s - (N - nsteps) + off, count, m.d1, m.La, m.Lb * m.Lbn * m.S, offA.p, offB.p);
        ctx->launches++;
        V *tmp = nullptr, *nxt = nullptr;
        // tmp[q][br + Lbn*(h + S*al)] = sum_ar env[ar, br] * A[al, i, h, ar]
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&tmp, (size_t)(m.Lbn * m.S * m.La) * count * sizeof(V)));
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)(m.La * m.Lb) * count * sizeof(V)));
        int rc = 0;
        for (i64 h = 0; h < m.S && !rc; ++h) // (Lbn x La) = env^T (Lbn x Lan) * A_{i,h}^T (Lan x La)
            rc = gemm_off(ctx, true, true, m.Lbn, m.La, m.Lan, 1.0, env, m.Lan, m.Lan * m.Lbn,
                                       m.A + m.La * m.d1 * h, m.La * m.d1 * m.S, 0, 0.0, tmp + m.Lbn * h, m.Lbn * m.S,
                                       m.Lbn * m.S * m.La, count, nullptr, offA.p, m.La % 2 == 0);
        // nxt[q][al, bl] = sum_{br,h} tmp[br, h, al] * B[bl, h, j, br]: (La x Lb) = tmp^T (La x Lbn*S) * Bp_j^T (Lbn*S x Lb)
        if (!rc)
            rc = gemm_off(ctx, true, true, m.La, m.Lb, m.Lbn * m.S, 1.0, tmp, m.Lbn * m.S, m.Lbn * m.S * m.La,
                          (const V *)t.Bp[s], m.Lb, 0, 0.0, nxt, m.La, m.La * m.Lb, count, nullptr, offB.p,
                          (m.Lb * m.Lbn * m.S) % 2 == 0);
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
int env_eval_mpo(tci_ctx *ctx, TargetDev &t, int side, const i64 *d_idx, int len, i64 count, double **out, i64 *D,
                 const i64 *h_idx)
{
    i64 a = 1, b = 1;
    int rc = side == 0 ? mpo_left_chain<double>(ctx, t, len, d_idx, len, 0, count, out, &a, &b, h_idx)
                       : mpo_right_chain<double>(ctx, t, len, d_idx, len, 0, count, out, &a, &b, h_idx);
    *D = a * b;
    return rc;
}

// batchevaluate(::Contraction) contraction.jl:236-335 (projector = nothing, f = nothing); outp: (nI*C) x nJ, ld outld
template <class V>
static int pi_eval_mpo_t(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                         V *outp, i64 outld, const i64 *hI, const i64 *hJ)
{
    V *X = nullptr, *right = nullptr;
    i64 La = 1, Lb = 1, Ra = 1, Rb = 1;
    int rc = mpo_left_chain<V>(ctx, t, (int)nl, dI, (int)nl, 0, nI, &X, &La, &Lb, hI);
    if (rc) return rc;
    rc = mpo_right_chain<V>(ctx, t, (int)nr, dJ, (int)nr, 0, nJ, &right, &Ra, &Rb, hJ);
    if (rc) {
        dev_free(ctx, X);
        return rc;
    }
    // X: (La x Lb x R), r = i + nI*sacc ; centre sites :290-317
    i64 R = nI;
    for (i64 s = nl; s < nl + M && !rc; ++s) {
        MpoSiteT<V> m = site_of<V>(t, s);
        const i64 N1 = m.d1 * m.S * m.Lan;
        V *tmp = nullptr, *nxt = nullptr;
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&tmp, (size_t)(m.Lb * R * N1) * sizeof(V)));
        TCI_CUDA(ctx, dev_alloc(ctx, (void **)&nxt, (size_t)(m.Lan * m.Lbn * R * m.d1 * m.d3) * sizeof(V)));
        // tmp[(b + Lb*r) + Lb*R*(x + d1*(h + S*an))] = sum_a X[a, b, r] * A[a, x, h, an]
        rc = gemm_off(ctx, true, false, m.Lb * R, N1, m.La, 1.0, X, m.La, 0, m.A, m.La, 0, 0.0, tmp, m.Lb * R, 0, 1, nullptr,
                      nullptr, false);
        // nxt[an + Lan*(bn + Lbn*(r + R*(x + d1*z)))] = sum_{b,h} tmp[b, r, x, h, an] * B[b, h, z, bn]
        for (i64 z = 0; z < m.d3 && !rc; ++z)
            for (i64 x = 0; x < m.d1 && !rc; ++x)
                for (i64 h = 0; h < m.S && !rc; ++h)
                    rc = gemm_off(ctx, true, false, m.Lan, m.Lbn, m.Lb, 1.0, tmp + m.Lb * R * (x + m.d1 * h),
                                  m.Lb * R * m.d1 * m.S, m.Lb, m.B + m.Lb * (h + m.S * z), m.Lb * m.S * m.d3, 0,
                                  h ? 1.0 : 0.0, nxt + m.Lan * m.Lbn * R * (x + m.d1 * z), m.Lan, m.Lan * m.Lbn, R, nullptr,
                                  nullptr, false);
        dev_free(ctx, tmp);
        dev_free(ctx, X);
        X = nxt;
        La = m.Lan;
        Lb = m.Lbn;
        R *= m.d1 * m.d3;
    }
    // res[r, j] = sum_{a,b} X[a, b, r] * right[a, b, j]   :328
    if (!rc)
        rc = gemm_off(ctx, true, false, R, nJ, La * Lb, 1.0, X, La * Lb, 0, right, Ra * Rb, 0, 0.0, outp, outld, 0, 1, nullptr,
                      nullptr, false);
    dev_free(ctx, X);
    dev_free(ctx, right);
    return rc;
}

int pi_eval_mpo(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                tci_dmat *out, const i64 *hI, const i64 *hJ)
{
    return pi_eval_mpo_t<double>(ctx, t, dI, nl, nI, dJ, nr, nJ, M, out->p, out->ld, hI, hJ);
}
// ComplexF64 MPO pair: `out` holds interleaved (re, im) pairs, 2 * rows doubles per column
int pi_eval_mpo_z(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                  tci_dmat *out, const i64 *hI, const i64 *hJ)
{
    return pi_eval_mpo_t<double2>(ctx, t, dI, nl, nI, dJ, nr, nJ, M, (double2 *)out->p, out->ld / 2, hI, hJ);
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
// sum(left .* right) of contraction.jl:196-200 on ComplexF64 (no conjugation), accumulated in order
__global__ void k_dot_env_z(const double2 *__restrict__ l, const double2 *__restrict__ r, i64 D, i64 count,
                            double2 *__restrict__ out)
{
    i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    double re = 0.0, im = 0.0;
    for (i64 a = 0; a < D; ++a) {
        const double2 x = l[a + D * q], y = r[a + D * q];
        re = __dadd_rn(re, __dsub_rn(__dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y)));
        im = __dadd_rn(im, __dadd_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x)));
    }
    out[q] = make_double2(re, im);
}

// evaluate(::Contraction, indexset)  contraction.jl:189-207
int target_eval_mpo(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out)
{
    const int N = (int)t.nsites, mid = N / 2;
    double *l = nullptr, *r = nullptr;
    i64 La, Lb, Ra, Rb;
    int rc = mpo_left_chain<double>(ctx, t, mid, d_idx, N, 0, count, &l, &La, &Lb);
    if (rc) return rc;
    rc = mpo_right_chain<double>(ctx, t, N - mid, d_idx, N, mid, count, &r, &Ra, &Rb);
    if (!rc) {
        k_dot_env<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(l, r, La * Lb, count, d_out);
        ctx->launches++;
    }
    dev_free(ctx, l);
    dev_free(ctx, r);
    return rc;
}
int target_eval_mpo_z(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out)
{
    const int N = (int)t.nsites, mid = N / 2;
    double2 *l = nullptr, *r = nullptr;
    i64 La, Lb, Ra, Rb;
    int rc = mpo_left_chain<double2>(ctx, t, mid, d_idx, N, 0, count, &l, &La, &Lb);
    if (rc) return rc;
    rc = mpo_right_chain<double2>(ctx, t, N - mid, d_idx, N, mid, count, &r, &Ra, &Rb);
    if (!rc) {
        k_dot_env_z<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(l, r, La * Lb, count, (double2 *)d_out);
        ctx->launches++;
    }
    dev_free(ctx, l);
    dev_free(ctx, r);
    return rc;
}

// ---- K6 -------------------------------------------------------------------------
// one zip-up step (contraction.jl:455-464) for value type V (double, or double2 = interleaved ComplexF64); sizes in
// elements of V; C is a ((W * chi*s1*s3) x Dan*Dbn) device matrix of doubles, W = doubles per element
template <class V>
static int zipup_site_t(tci_ctx *ctx, const double *R, i64 chi, i64 Da, i64 Db, const double *A, i64 s1, i64 s2, i64 Dan,
                        const double *B, i64 s3, i64 Dbn, double *C_host, tci_dmat **C_dev)
{
    constexpr i64 W = sizeof(V) / sizeof(double);
    if (C_dev) *C_dev = nullptr;
    if (chi < 1 || Da < 1 || Db < 1 || s1 < 1 || s2 < 1 || s3 < 1 || Dan < 1 || Dbn < 1 || !R || !A || !B)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_contract_zipup_site: bad arguments");
    DevBuf<double> dR(ctx), dA(ctx), dB(ctx), RA(ctx);
    {
        StageTimer tm(ctx, ST_H2D);
        TCI_CUDA(ctx, dR.upload(R, (size_t)(W * chi * Da * Db)));
        TCI_CUDA(ctx, dA.upload(A, (size_t)(W * Da * s1 * s2 * Dan)));
        TCI_CUDA(ctx, dB.upload(B, (size_t)(W * Db * s2 * s3 * Dbn)));
    }
    tci_dmat *C = nullptr;
    int rc = dmat_alloc(ctx, W * chi * s1 * s3, Dan * Dbn, &C);
    if (rc) return rc;
    {
        StageTimer tm(ctx, ST_GEMM);
        const V *pR = (const V *)dR.p, *pA = (const V *)dA.p, *pB = (const V *)dB.p;
        V *pC = (V *)C->p;
        const i64 ldC = C->ld / W;
        // RA2[(c + chi*x) + chi*s1*((b + Db*h) + Db*s2*an)] = sum_a R[c, a, b] * A[a, x, h, an]   :458
        TCI_CUDA(ctx, RA.alloc((size_t)(W * chi * s1 * Db * s2 * Dan)));
        V *pRA = (V *)RA.p;
        for (i64 h = 0; h < s2 && !rc; ++h)
            for (i64 x = 0; x < s1 && !rc; ++x)
                rc = gemm_off(ctx, false, false, chi, Dan, Da, 1.0, pR, chi, chi * Da, pA + Da * (x + s1 * h), Da * s1 * s2, 0,
                              0.0, pRA + chi * x + chi * s1 * Db * h, chi * s1 * Db * s2, chi * s1, Db, nullptr, nullptr, false);
        // C[(c + chi*(x + s1*z)) + chi*s1*s3*(an + Dan*bn)] = sum_{b,h} RA2[(c,x),(b,h),an] * B[b, h, z, bn]  :464
        for (i64 z = 0; z < s3 && !rc; ++z)
            rc = gemm_off(ctx, false, false, chi * s1, Dbn, Db * s2, 1.0, pRA, chi * s1, chi * s1 * Db * s2, pB + Db * s2 * z,
                          Db * s2 * s3, 0, 0.0, pC + chi * s1 * z, ldC * Dan, ldC, Dan, nullptr, nullptr, false);
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

extern "C" int tci_contract_zipup_site(tci_ctx *ctx, const double *R, int64_t chi, int64_t Da, int64_t Db,
                                       const double *A, int64_t s1, int64_t s2, int64_t Dan, const double *B,
                                       int64_t s3, int64_t Dbn, double *C_host, tci_dmat **C_dev)
{
    TCI_ENTER(ctx);
    return zipup_site_t<double>(ctx, R, chi, Da, Db, A, s1, s2, Dan, B, s3, Dbn, C_host, C_dev);
}
extern "C" int tci_zcontract_zipup_site(tci_ctx *ctx, const double *R, int64_t chi, int64_t Da, int64_t Db,
                                        const double *A, int64_t s1, int64_t s2, int64_t Dan, const double *B,
                                        int64_t s3, int64_t Dbn, double *C_host, tci_dmat **C_dev)
{
    TCI_ENTER(ctx);
    return zipup_site_t<double2>(ctx, R, chi, Da, Db, A, s1, s2, Dan, B, s3, Dbn, C_host, C_dev);
}

__device__ __forceinline__ double v_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ double2 v_fma(double2 a, double2 b, double2 c)
{ // c + a*b, complex (BLAS-level arithmetic: the reference's _contract is a zgemm)
    c.x = fma(a.x, b.x, fma(-a.y, b.y, c.x));
    c.y = fma(a.x, b.y, fma(a.y, b.x, c.y));
    return c;
}
template <class V> __device__ __forceinline__ V v_zero();
template <> __device__ __forceinline__ double v_zero<double>() { return 0.0; }
template <> __device__ __forceinline__ double2 v_zero<double2>() { return make_double2(0.0, 0.0); }

// out[(la + Da*lb), x, z, (lan + Dan*lbn)] = sum_h A[la, x, h, lan] * B[lb, h, z, lbn]   contraction.jl:338-349
template <class V>
__global__ void k_naive_site(const V *__restrict__ A, const V *__restrict__ B, i64 Da, i64 s1, i64 s2, i64 Dan, i64 Db, i64 s3,
                             i64 Dbn, V *__restrict__ out)
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
        V acc = v_zero<V>();
        for (i64 h = 0; h < s2; ++h)
            acc = v_fma(A[la + Da * (x + s1 * (h + s2 * lan))], B[lb + Db * (h + s2 * (z + s3 * lbn))], acc);
        out[e] = acc;
    }
}

template <class V>
static int naive_site_t(tci_ctx *ctx, const double *A, i64 Da, i64 s1, i64 s2, i64 Dan, const double *B, i64 Db, i64 s3,
                        i64 Dbn, double *out_host)
{
    constexpr i64 W = sizeof(V) / sizeof(double);
    if (!A || !B || !out_host) return tci_fail(ctx, TCI_ERR_ARG, "tci_contract_naive_site: bad arguments");
    const i64 total = Da * Db * s1 * s3 * Dan * Dbn;
    DevBuf<double> dA(ctx), dB(ctx), dO(ctx);
    TCI_CUDA(ctx, dA.upload(A, (size_t)(W * Da * s1 * s2 * Dan)));
    TCI_CUDA(ctx, dB.upload(B, (size_t)(W * Db * s2 * s3 * Dbn)));
    TCI_CUDA(ctx, dO.alloc((size_t)(W * total)));
    unsigned blocks = (unsigned)std::min<i64>((total + 255) / 256, (i64)ctx->sm_count * 16);
    k_naive_site<V><<<blocks, 256, 0, ctx->stream>>>((const V *)dA.p, (const V *)dB.p, Da, s1, s2, Dan, Db, s3, Dbn, (V *)dO.p);
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    TCI_CUDA(ctx, cudaMemcpyAsync(out_host, dO.p, W * total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

extern "C" int tci_contract_naive_site(tci_ctx *ctx, const double *A, int64_t Da, int64_t s1, int64_t s2, int64_t Dan,
                                       const double *B, int64_t Db, int64_t s3, int64_t Dbn, double *out_host)
{
    TCI_ENTER(ctx);
    return naive_site_t<double>(ctx, A, Da, s1, s2, Dan, B, Db, s3, Dbn, out_host);
}
extern "C" int tci_zcontract_naive_site(tci_ctx *ctx, const double *A, int64_t Da, int64_t s1, int64_t s2, int64_t Dan,
                                        const double *B, int64_t Db, int64_t s3, int64_t Dbn, double *out_host)
{
    TCI_ENTER(ctx);
    return naive_site_t<double2>(ctx, A, Da, s1, s2, Dan, B, Db, s3, Dbn, out_host);
}
