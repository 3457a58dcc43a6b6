// luci.cu -- K3: left(luci) / right(luci) of MatrixLUCI (matrixluci.jl:40-84) from the
// device-resident factorisation: unit-triangular TRSM (blocked: 32-wide diagonal solves
// in shared memory + DGEMM updates), triangular products as DGEMM, and the final row /
// column un-permutation (matrixluci.jl:55,66; matrixlu.jl:374-392).
#include "tci_internal.h"

int lu_extract(tci_lu *lu, double *dL, i64 ldl, double *dU, i64 ldu); // rrlu.cu

#define TB_NB 32
#define TB_THREADS 128

// X[:, 0:nb] * T = X[:, 0:nb] with T (nb x nb) lower triangular (unit diagonal if UNIT):
// x_j = (x_j - sum_{k>j} x_k T[k,j]) / T[j,j], j = nb-1 .. 0.  One thread per row of X.
template <bool UNIT>
__global__ void __launch_bounds__(TB_THREADS)
    k_trsm_rl_block(double *__restrict__ X, i64 rows, i64 ldx, const double *__restrict__ T, i64 ldt, int nb)
{
    __shared__ double Ts[TB_NB][TB_NB + 1];
    __shared__ double xs[TB_NB][TB_THREADS];
    for (int e = threadIdx.x; e < nb * nb; e += TB_THREADS) Ts[e % nb][e / nb] = T[(e % nb) + ldt * (e / nb)];
    __syncthreads();
    i64 row = blockIdx.x * (i64)TB_THREADS + threadIdx.x;
    if (row >= rows) return;
    for (int j = 0; j < nb; ++j) xs[j][threadIdx.x] = X[row + ldx * j];
    for (int j = nb - 1; j >= 0; --j) {
        double acc = xs[j][threadIdx.x];
        for (int k = j + 1; k < nb; ++k) acc = fma(-xs[k][threadIdx.x], Ts[k][j], acc);
        xs[j][threadIdx.x] = UNIT ? acc : acc / Ts[j][j];
    }
    for (int j = 0; j < nb; ++j) X[row + ldx * j] = xs[j][threadIdx.x];
}

// X[:, 0:nb] * T = X[:, 0:nb] with T (nb x nb) upper triangular (unit diagonal if UNIT):
// x_j = (x_j - sum_{k<j} x_k T[k,j]) / T[j,j], j = 0 .. nb-1.  One thread per row of X.
template <bool UNIT>
__global__ void __launch_bounds__(TB_THREADS)
    k_trsm_ru_block(double *__restrict__ X, i64 rows, i64 ldx, const double *__restrict__ T, i64 ldt, int nb)
{
    __shared__ double Ts[TB_NB][TB_NB + 1];
    __shared__ double xs[TB_NB][TB_THREADS];
    for (int e = threadIdx.x; e < nb * nb; e += TB_THREADS) Ts[e % nb][e / nb] = T[(e % nb) + ldt * (e / nb)];
    __syncthreads();
    i64 row = blockIdx.x * (i64)TB_THREADS + threadIdx.x;
    if (row >= rows) return;
    for (int j = 0; j < nb; ++j) xs[j][threadIdx.x] = X[row + ldx * j];
    for (int j = 0; j < nb; ++j) {
        double acc = xs[j][threadIdx.x];
        for (int k = 0; k < j; ++k) acc = fma(-xs[k][threadIdx.x], Ts[k][j], acc);
        xs[j][threadIdx.x] = UNIT ? acc : acc / Ts[j][j];
    }
    for (int j = 0; j < nb; ++j) X[row + ldx * j] = xs[j][threadIdx.x];
}

// dst[:, q] = src[:, perm[q]]
__global__ void k_gather_cols(const double *__restrict__ src, i64 lds, i64 rows, i64 n, const i64 *__restrict__ perm,
                              double *__restrict__ dst, i64 ldd)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= rows * n) return;
    i64 i = e % rows, q = e / rows;
    dst[i + ldd * q] = src[i + lds * perm[q]];
}

// T * X[0:nb, :] = X[0:nb, :] with T (nb x nb) unit upper triangular:
// x_i -= sum_{k>i} T[i,k] x_k, i = nb-1 .. 0.  One thread per column of X.
__global__ void __launch_bounds__(TB_THREADS)
    k_trsm_lu_block(double *__restrict__ X, i64 cols, i64 ldx, const double *__restrict__ T, i64 ldt, int nb)
{
    __shared__ double Ts[TB_NB][TB_NB + 1];
    __shared__ double xs[TB_NB][TB_THREADS];
    for (int e = threadIdx.x; e < nb * nb; e += TB_THREADS) Ts[e % nb][e / nb] = T[(e % nb) + ldt * (e / nb)];
    __syncthreads();
    i64 col = blockIdx.x * (i64)TB_THREADS + threadIdx.x;
    if (col >= cols) return;
    for (int i = 0; i < nb; ++i) xs[i][threadIdx.x] = X[i + ldx * col];
    for (int i = nb - 1; i >= 0; --i) {
        double acc = xs[i][threadIdx.x];
        for (int k = i + 1; k < nb; ++k) acc = fma(-Ts[i][k], xs[k][threadIdx.x], acc);
        xs[i][threadIdx.x] = acc;
    }
    for (int i = 0; i < nb; ++i) X[i + ldx * col] = xs[i][threadIdx.x];
}

// out[perm[i], :] = src[i, :]  (rows)   or   out[:, perm[q]] = src[:, q]  (cols)
__global__ void k_scatter_rows(const double *__restrict__ src, i64 lds, i64 m, i64 r, const i64 *__restrict__ perm,
                               double *__restrict__ out, i64 ldo, int identity_top)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= m * r) return;
    i64 i = e % m, c = e / m;
    double v = (identity_top && i < r) ? (i == c ? 1.0 : 0.0) : src[i + lds * c];
    out[perm[i] + ldo * c] = v;
}
__global__ void k_scatter_cols(const double *__restrict__ src, i64 lds, i64 r, i64 n, const i64 *__restrict__ perm,
                               double *__restrict__ out, i64 ldo, int identity_left)
{
    i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (e >= r * n) return;
    i64 i = e % r, q = e / r;
    double v = (identity_left && q < r) ? (i == q ? 1.0 : 0.0) : src[i + lds * q];
    out[i + ldo * perm[q]] = v;
}

// Recursive splitting (the updates between the halves are large GEMMs that run on the DMMA kernel; the leaves are
// 32-wide triangular solves in shared memory).  A loop over 32-wide blocks would make every update a GEMM with N = 32.
static i64 trsm_split(i64 k) { return ((k / 2 + TB_NB - 1) / TB_NB) * TB_NB; }

// W (rows x k, ldw) <- W * U^-1 for an upper-triangular U (k x k, ldu; unit diagonal if `unit`)
static int trsm_right_upper(tci_ctx *ctx, double *W, i64 rows, i64 ldw, const double *U, i64 ldu, i64 k, bool unit)
{
    if (k <= 0 || rows <= 0) return TCI_OK;
    if (k <= TB_NB) {
        const unsigned gb = (unsigned)((rows + TB_THREADS - 1) / TB_THREADS);
        if (unit)
            k_trsm_ru_block<true><<<gb, TB_THREADS, 0, ctx->stream>>>(W, rows, ldw, U, ldu, (int)k);
        else
            k_trsm_ru_block<false><<<gb, TB_THREADS, 0, ctx->stream>>>(W, rows, ldw, U, ldu, (int)k);
        ctx->launches++;
        return TCI_OK;
    }
    const i64 k1 = trsm_split(k), k2 = k - k1;
    int rc = trsm_right_upper(ctx, W, rows, ldw, U, ldu, k1, unit);                       // X1 U11 = W1
    if (!rc) rc = dgemm_dev(ctx, false, false, rows, k2, k1, -1.0, W, ldw, U + ldu * k1, ldu, 1.0, W + ldw * k1, ldw); // W2 -= X1 U12
    if (!rc) rc = trsm_right_upper(ctx, W + ldw * k1, rows, ldw, U + k1 + ldu * k1, ldu, k2, unit); // X2 U22 = W2
    return rc;
}

// W (rows x k, ldw) <- W * L^-1 for a lower-triangular L (k x k, ldl; unit diagonal if `unit`)
static int trsm_right_lower(tci_ctx *ctx, double *W, i64 rows, i64 ldw, const double *L, i64 ldl, i64 k, bool unit)
{
    if (k <= 0 || rows <= 0) return TCI_OK;
    if (k <= TB_NB) {
        const unsigned gb = (unsigned)((rows + TB_THREADS - 1) / TB_THREADS);
        if (unit)
            k_trsm_rl_block<true><<<gb, TB_THREADS, 0, ctx->stream>>>(W, rows, ldw, L, ldl, (int)k);
        else
            k_trsm_rl_block<false><<<gb, TB_THREADS, 0, ctx->stream>>>(W, rows, ldw, L, ldl, (int)k);
        ctx->launches++;
        return TCI_OK;
    }
    const i64 k1 = trsm_split(k), k2 = k - k1;
    int rc = trsm_right_lower(ctx, W + ldw * k1, rows, ldw, L + k1 + ldl * k1, ldl, k2, unit);   // X2 L22 = W2
    if (!rc) rc = dgemm_dev(ctx, false, false, rows, k1, k2, -1.0, W + ldw * k1, ldw, L + k1, ldl, 1.0, W, ldw); // W1 -= X2 L21
    if (!rc) rc = trsm_right_lower(ctx, W, rows, ldw, L, ldl, k1, unit);                           // X1 L11 = W1
    return rc;
}

// X (k x cols, ldx) <- U^-1 X for a UNIT upper-triangular U (k x k, ldu)
static int trsm_left_upper_unit(tci_ctx *ctx, double *X, i64 cols, i64 ldx, const double *U, i64 ldu, i64 k)
{
    if (k <= 0 || cols <= 0) return TCI_OK;
    if (k <= TB_NB) {
        k_trsm_lu_block<<<(unsigned)((cols + TB_THREADS - 1) / TB_THREADS), TB_THREADS, 0, ctx->stream>>>(X, cols, ldx, U,
                                                                                                  ldu, (int)k);
        ctx->launches++;
        return TCI_OK;
    }
    const i64 k1 = trsm_split(k), k2 = k - k1;
    int rc = trsm_left_upper_unit(ctx, X + k1, cols, ldx, U + k1 + ldu * k1, ldu, k2);            // U22 X2 = B2
    if (!rc) rc = dgemm_dev(ctx, false, false, k1, cols, k2, -1.0, U + ldu * k1, ldu, X + k1, ldx, 1.0, X, ldx); // B1 -= U12 X2
    if (!rc) rc = trsm_left_upper_unit(ctx, X, cols, ldx, U, ldu, k1);                            // U11 X1 = B1
    return rc;
}

static int finish(tci_ctx *ctx, tci_dmat *res, double *out_host, tci_dmat **out_dev)
{
    if (out_host && res->m * res->n > 0) {
        StageTimer tm(ctx, ST_D2H);
        TCI_CUDA(ctx, cudaMemcpy2DAsync(out_host, res->m * sizeof(double), res->p, res->ld * sizeof(double),
                                        res->m * sizeof(double), res->n, cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (out_dev)
        *out_dev = res;
    else
        tci_dmat_destroy(res);
    return TCI_OK;
}

extern "C" int tci_luci_left(tci_lu *lu, double *out_host, tci_dmat **out_dev)
{
    if (!lu) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    if (out_dev) *out_dev = nullptr;
    const i64 m = lu->m, r = lu->r;
    tci_dmat *res = nullptr;
    int rc = dmat_alloc(ctx, m, r, &res);
    if (rc) return rc;
    if (r == 0) return finish(ctx, res, out_host, out_dev);
    {
        StageTimer tm(ctx, ST_LUCI);
        DevBuf<double> L(ctx), U(ctx), Y(ctx);
        TCI_CUDA(ctx, L.alloc((size_t)(m * r)));
        if (lu->leftorthogonal) { // colstimespivotinv  matrixluci.jl:48-57
            rc = lu_extract(lu, L.p, m, nullptr, 0);
            const i64 rows = m - r;
            double *X = L.p + r;
            if (!rc) rc = trsm_right_lower(ctx, X, rows, m, L.p, m, r, true); // X L11 = L21
            if (!rc) {
                k_scatter_rows<<<(unsigned)((m * r + 255) / 256), 256, 0, ctx->stream>>>(L.p, m, m, r, lu->d_rowperm,
                                                                                        res->p, res->ld, 1);
                ctx->launches++;
            }
        } else { // colmatrix  matrixluci.jl:40-42 : left(lu) * U[:, 1:r]
            TCI_CUDA(ctx, U.alloc((size_t)(r * lu->n)));
            TCI_CUDA(ctx, Y.alloc((size_t)(m * r)));
            rc = lu_extract(lu, L.p, m, U.p, r);
            if (!rc) rc = dgemm_dev(ctx, false, false, m, r, r, 1.0, L.p, m, U.p, r, 0.0, Y.p, m);
            if (!rc) {
                k_scatter_rows<<<(unsigned)((m * r + 255) / 256), 256, 0, ctx->stream>>>(Y.p, m, m, r, lu->d_rowperm,
                                                                                        res->p, res->ld, 0);
                ctx->launches++;
            }
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, "luci_left launch failed");
    }
    if (rc) {
        tci_dmat_destroy(res);
        return rc;
    }
    return finish(ctx, res, out_host, out_dev);
}

extern "C" int tci_luci_right(tci_lu *lu, double *out_host, tci_dmat **out_dev)
{
    if (!lu) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    if (out_dev) *out_dev = nullptr;
    const i64 m = lu->m, n = lu->n, r = lu->r;
    tci_dmat *res = nullptr;
    int rc = dmat_alloc(ctx, r, n, &res);
    if (rc) return rc;
    if (r == 0) return finish(ctx, res, out_host, out_dev);
    {
        StageTimer tm(ctx, ST_LUCI);
        DevBuf<double> L(ctx), U(ctx), Y(ctx);
        TCI_CUDA(ctx, U.alloc((size_t)(r * n)));
        if (lu->leftorthogonal) { // rowmatrix  matrixluci.jl:44-46 : L[1:r, :] * right(lu)
            TCI_CUDA(ctx, L.alloc((size_t)(m * r)));
            TCI_CUDA(ctx, Y.alloc((size_t)(r * n)));
            rc = lu_extract(lu, L.p, m, U.p, r);
            if (!rc) rc = dgemm_dev(ctx, false, false, r, n, r, 1.0, L.p, m, U.p, r, 0.0, Y.p, r);
            if (!rc) {
                k_scatter_cols<<<(unsigned)((r * n + 255) / 256), 256, 0, ctx->stream>>>(Y.p, r, r, n, lu->d_colperm,
                                                                                        res->p, res->ld, 0);
                ctx->launches++;
            }
        } else { // pivotinvtimesrows  matrixluci.jl:59-68 : U11 \ U12
            rc = lu_extract(lu, nullptr, 0, U.p, r);
            const i64 cols = n - r;
            double *X = U.p + r * r;
            if (!rc) rc = trsm_left_upper_unit(ctx, X, cols, r, U.p, r, r); // U11 X = U12
            if (!rc) {
                k_scatter_cols<<<(unsigned)((r * n + 255) / 256), 256, 0, ctx->stream>>>(U.p, r, r, n, lu->d_colperm,
                                                                                        res->p, res->ld, 1);
                ctx->launches++;
            }
        }
        if (!rc && cudaGetLastError() != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, "luci_right launch failed");
    }
    if (rc) {
        tci_dmat_destroy(res);
        return rc;
    }
    return finish(ctx, res, out_host, out_dev);
}

// B * A^-1 for the square, fully factorised A = lu (the `\` of setsitetensor!, tensorci2.jl:391, which the
// reference leaves to LAPACK gesv; here the full-pivot factors of K2 are reused):
//   A[rowperm, colperm] = L U  =>  X[:, rowperm] = B[:, colperm] U^-1 L^-1
// W (rows x k, ldw) <- W * U^-1 for an upper-triangular U (k x k, ldu; unit diagonal if `unit`), blocked: 32-wide
// diagonal solves in shared memory + DMMA GEMM updates of the columns to the right
// dst (n x m, ldd) = src (m x n, lds)^T
__global__ void k_transpose(const double *__restrict__ src, i64 lds, i64 m, i64 n, double *__restrict__ dst, i64 ldd)
{
    __shared__ double tile[32][33];
    const i64 i0 = (i64)blockIdx.x * 32, j0 = (i64)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const i64 i = i0 + threadIdx.x, j = j0 + r;
        if (i < m && j < n) tile[r][threadIdx.x] = src[i + lds * j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const i64 j = j0 + threadIdx.x, i = i0 + r;
        if (i < m && j < n) dst[j + ldd * i] = tile[threadIdx.x][r];
    }
}

// The completion of a rook-search factorisation (arrlu, matrixlu.jl:274-288): with L11 / U11 the r x r pivot blocks of
// `lu`, L2 = A21 U11^-1 (cols2Lmatrix!, :314-335) for the rows the search never visited and U2 = L11^-1 A12
// (rows2Umatrix!, :337-358) for the columns -- two triangular solves, blocked TRSM + DMMA GEMM on the device (the
// left-lower solve runs as a right-upper solve on the transposed system).  The reference divides by the diagonal in
// both (it is 1 on the unit-diagonal factor).
extern "C" int tci_lu_complete(tci_lu *lu, tci_dmat *A21, tci_dmat *A12, double *L2_host, double *U2_host)
{
    if (!lu) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    const i64 r = lu->r;
    if ((A21 && A21->n != r) || (A12 && A12->m != r))
        return tci_fail(ctx, TCI_ERR_ARG, A21 && A21->n != r
                                              ? "C and P matrices must have same number of columns in `cols2Lmatrix!`."
                                              : "R and P matrices must have same number of rows in `rows2Umatrix!`.");
    if ((A21 && !L2_host) || (A12 && !U2_host)) return tci_fail(ctx, TCI_ERR_ARG, "tci_lu_complete: output missing");
    if (r == 0 || lu->m < r || lu->n < r) return TCI_OK;
    DevBuf<double> L(ctx), U(ctx), W(ctx), Wt(ctx), Lt(ctx);
    TCI_CUDA(ctx, L.alloc((size_t)(lu->m * r)));
    TCI_CUDA(ctx, U.alloc((size_t)(r * lu->n)));
    int rc = lu_extract(lu, L.p, lu->m, U.p, r);
    if (rc) return rc;
    StageTimer tm(ctx, ST_LUCI);
    if (A21 && A21->m > 0) { // X U11 = A21
        const i64 rows = A21->m;
        dmat_wait_ready(ctx, A21);
        TCI_CUDA(ctx, W.alloc((size_t)(rows * r)));
        TCI_CUDA(ctx, cudaMemcpy2DAsync(W.p, rows * sizeof(double), A21->p, A21->ld * sizeof(double), rows * sizeof(double),
                                        r, cudaMemcpyDeviceToDevice, ctx->stream));
        rc = trsm_right_upper(ctx, W.p, rows, rows, U.p, r, r, !lu->leftorthogonal);
        if (rc) return rc;
        TCI_CUDA(ctx, cudaMemcpyAsync(L2_host, W.p, (size_t)(rows * r) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (A12 && A12->n > 0) { // L11 X = A12  <=>  X^T L11^T = A12^T
        const i64 cols = A12->n;
        dmat_wait_ready(ctx, A12);
        TCI_CUDA(ctx, Wt.alloc((size_t)(cols * r)));
        TCI_CUDA(ctx, Lt.alloc((size_t)(r * r)));
        const dim3 tb(32, 8);
        k_transpose<<<dim3((unsigned)((r + 31) / 32), (unsigned)((cols + 31) / 32)), tb, 0, ctx->stream>>>(A12->p, A12->ld, r,
                                                                                                     cols, Wt.p, cols);
        k_transpose<<<dim3((unsigned)((r + 31) / 32), (unsigned)((r + 31) / 32)), tb, 0, ctx->stream>>>(L.p, lu->m, r, r,
                                                                                                  Lt.p, r);
        ctx->launches += 2;
        rc = trsm_right_upper(ctx, Wt.p, cols, cols, Lt.p, r, r, lu->leftorthogonal);
        if (rc) return rc;
        // back to r x cols, reusing the buffer of A12's transpose is not possible in place: go through U's storage
        DevBuf<double> X(ctx);
        TCI_CUDA(ctx, X.alloc((size_t)(r * cols)));
        k_transpose<<<dim3((unsigned)((cols + 31) / 32), (unsigned)((r + 31) / 32)), tb, 0, ctx->stream>>>(Wt.p, cols, cols, r,
                                                                                                     X.p, r);
        ctx->launches++;
        TCI_CUDA(ctx, cudaMemcpyAsync(U2_host, X.p, (size_t)(r * cols) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // X goes out of scope
    }
    TCI_CUDA(ctx, cudaGetLastError());
    tm.stop();
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

// ---- batched B * A^-1 for small pivot matrices ------------------------------------------------------------------
// fillsitetensors! solves T_b = Pi1_b P_b^-1 for every site (tensorci2.jl:391); at chi <= 256 each solve is a chain of ~30
// tiny launches (extract, gather, recursive TRSM x 2, scatter), and the whole stage is launch-bound (config 3: 2.0 ms
// for 19 sites).  Here ONE launch handles all sites: a CTA takes RD_RB rows of one site, keeps them in shared memory and
// runs both triangular solves on them against the factors as they sit in the factorised matrix (rows physically
// permuted, columns through colperm; L unit lower, U upper: the leftorthogonal form tci_fill_sitetensors uses):
//   W = B[:, colperm];  W <- W U^-1 (columns left to right);  W <- W L^-1 (right to left);  X[:, rowperm[c]] = W[:, c].
// Blocked by 32 columns: a thread per row solves the diagonal block, then all threads update the columns still to come.
#define RD_RB 32
#define RD_THREADS 256
struct RdivJob {
    const double *A; // factorised P, lda
    i64 lda;
    const i64 *colperm, *rowperm;
    const double *B;
    i64 ldb;
    double *X;
    i64 ldx;
    int rows, k, cta0; // first CTA of this job
};

__global__ void __launch_bounds__(RD_THREADS) k_rdiv_small(const RdivJob *__restrict__ jobs, int njobs)
{
    extern __shared__ __align__(16) double rd_sm[]; // W[c][RD_RB] (column c of the row block contiguous)
    __shared__ i64 cp_s[256];
    int lo = 0, hi = njobs - 1; // the job this CTA belongs to
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].cta0 <= (int)blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const RdivJob jb = jobs[lo];
    const int k = jb.k, tid = threadIdx.x;
    const int r0 = ((int)blockIdx.x - jb.cta0) * RD_RB;
    const int nr = min(RD_RB, jb.rows - r0);
    for (int c = tid; c < k; c += RD_THREADS) cp_s[c] = jb.colperm[c];
    __syncthreads();
    for (int e = tid; e < k * RD_RB; e += RD_THREADS) {
        const int r = e % RD_RB, c = e / RD_RB;
        rd_sm[e] = r < nr ? jb.B[r0 + r + jb.ldb * cp_s[c]] : 0.0;
    }
    __syncthreads();
    const int r = tid % RD_RB, cg = tid / RD_RB; // row, column group (RD_THREADS / RD_RB = 8 groups)
    constexpr int NG = RD_THREADS / RD_RB;
    __shared__ double dg[32][33]; // the 32 x 32 diagonal block of the factor, dg[j][c - cb]
    // one column of the trailing update: W[:, c] -= sum_{j in [cb, ce)} W[:, j] * f[j], f = column c of the factor.  The
    // 32 factor values are fetched first (independent loads: they come from L2, the factor does not fit L1), then used.
    auto update_col = [&](const double *fcol, int c, int cb, int ce) {
        double f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = cb + j < ce ? fcol[cb + j] : 0.0;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (cb + j < ce) a0 = fma(rd_sm[(cb + j) * RD_RB + r], f[j], a0);
            if (cb + j + 1 < ce) a1 = fma(rd_sm[(cb + j + 1) * RD_RB + r], f[j + 1], a1);
            if (cb + j + 2 < ce) a2 = fma(rd_sm[(cb + j + 2) * RD_RB + r], f[j + 2], a2);
            if (cb + j + 3 < ce) a3 = fma(rd_sm[(cb + j + 3) * RD_RB + r], f[j + 3], a3);
        }
        rd_sm[c * RD_RB + r] -= (a0 + a1) + (a2 + a3);
    };
    // ---- W <- W U^-1 ----
    for (int cb = 0; cb < k; cb += 32) {
        const int ce = min(cb + 32, k);
        for (int e = tid; e < 32 * 32; e += RD_THREADS) { // stage the diagonal block
            const int j = e % 32, c = e / 32;
            dg[j][c] = (cb + j < ce && cb + c < ce) ? jb.A[cb + j + jb.lda * cp_s[cb + c]] : 0.0;
        }
        __syncthreads();
        if (tid < RD_RB) { // diagonal block, a thread per row
            for (int c = cb; c < ce; ++c) {
                double acc = rd_sm[c * RD_RB + r];
                for (int j = cb; j < c; ++j) acc = fma(-rd_sm[j * RD_RB + r], dg[j - cb][c - cb], acc);
                rd_sm[c * RD_RB + r] = acc / dg[c - cb][c - cb];
            }
        }
        __syncthreads();
        for (int c = ce + cg; c < k; c += NG) update_col(jb.A + jb.lda * cp_s[c], c, cb, ce); // columns still to come
        __syncthreads();
    }
    // ---- W <- W L^-1 (unit lower; columns right to left) ----
    for (int ce = k; ce > 0; ce -= 32) {
        const int cb = max(ce - 32, 0);
        for (int e = tid; e < 32 * 32; e += RD_THREADS) {
            const int j = e % 32, c = e / 32;
            dg[j][c] = (cb + j < ce && cb + c < ce) ? jb.A[cb + j + jb.lda * cp_s[cb + c]] : 0.0;
        }
        __syncthreads();
        if (tid < RD_RB) {
            for (int c = ce - 1; c >= cb; --c) {
                double acc = rd_sm[c * RD_RB + r];
                for (int j = c + 1; j < ce; ++j) acc = fma(-rd_sm[j * RD_RB + r], dg[j - cb][c - cb], acc);
                rd_sm[c * RD_RB + r] = acc;
            }
        }
        __syncthreads();
        for (int c = cg; c < cb; c += NG) update_col(jb.A + jb.lda * cp_s[c], c, cb, ce);
        __syncthreads();
    }
    for (int e = tid; e < k * RD_RB; e += RD_THREADS) {
        const int rr = e % RD_RB, c = e / RD_RB;
        if (rr < nr) jb.X[r0 + rr + jb.ldx * jb.rowperm[c]] = rd_sm[e];
    }
}

// all solves of a fillsitetensors! call whose pivot matrix has k <= 256 in one launch; done[q] tells the caller which
// jobs were taken (the others go through lu_rdiv_enqueue)
int lu_rdiv_batched_small(tci_ctx *ctx, int n, tci_lu *const *lus, const double *const *B, const i64 *ldb, const i64 *rows,
                          double *const *X, const i64 *ldx, std::vector<char> &done)
{
    done.assign((size_t)n, 0);
    static const bool off = getenv("TCI_NO_BATCHED_RDIV") != nullptr;
    if (off) return TCI_OK;
    std::vector<RdivJob> jobs;
    int cta = 0, kmax = 0;
    for (int q = 0; q < n; ++q) {
        tci_lu *lu = lus[q];
        if (!lu || lu->is_complex || !lu->leftorthogonal || lu->r > 256 || lu->r < 1 || lu->m != lu->n || lu->r != lu->m || rows[q] < 1)
            continue;
        RdivJob jb;
        jb.A = lu->A->p;
        jb.lda = lu->A->ld;
        jb.colperm = lu->d_colperm;
        jb.rowperm = lu->d_rowperm;
        jb.B = B[q];
        jb.ldb = ldb[q];
        jb.X = X[q];
        jb.ldx = ldx[q];
        jb.rows = (int)rows[q];
        jb.k = (int)lu->r;
        jb.cta0 = cta;
        cta += (int)((rows[q] + RD_RB - 1) / RD_RB);
        kmax = std::max(kmax, jb.k);
        jobs.push_back(jb);
        done[q] = 1;
    }
    if (jobs.empty()) return TCI_OK;
    DevBuf<RdivJob> dj(ctx);
    TCI_CUDA(ctx, dj.upload(jobs.data(), jobs.size()));
    const size_t smem = (size_t)kmax * RD_RB * sizeof(double);
    TCI_CUDA(ctx, ctx_func_smem(ctx, (const void *)k_rdiv_small, 256 * RD_RB * (int)sizeof(double)));
    k_rdiv_small<<<(unsigned)cta, RD_THREADS, smem, ctx->stream>>>(dj.p, (int)jobs.size());
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

// enqueue only: X (rows x k, ldx) = B (rows x k, ldb) * A^-1 with the factors of `lu` (assumed of full rank k)
int lu_rdiv_enqueue(tci_lu *lu, const double *B, i64 ldb, i64 rows, double *X, i64 ldx)
{
    tci_ctx *ctx = lu->ctx;
    const i64 k = lu->r;
    int rc = 0;
    DevBuf<double> L(ctx), U(ctx), W(ctx);
    TCI_CUDA(ctx, L.alloc((size_t)(k * k)));
    TCI_CUDA(ctx, U.alloc((size_t)(k * k)));
    TCI_CUDA(ctx, W.alloc((size_t)(rows * k)));
    rc = lu_extract(lu, L.p, k, U.p, k);
    if (!rc) {
        k_gather_cols<<<(unsigned)((rows * k + 255) / 256), 256, 0, ctx->stream>>>(B, ldb, rows, k, lu->d_colperm, W.p,
                                                                               rows);
        ctx->launches++;
    }
    if (!rc) rc = trsm_right_upper(ctx, W.p, rows, rows, U.p, k, k, !lu->leftorthogonal); // Y U = W
    if (!rc) rc = trsm_right_lower(ctx, W.p, rows, rows, L.p, k, k, lu->leftorthogonal); // X' L = Y
    if (!rc) {
        k_scatter_cols<<<(unsigned)((rows * k + 255) / 256), 256, 0, ctx->stream>>>(W.p, rows, rows, k, lu->d_rowperm, X,
                                                                                ldx, 0);
        ctx->launches++;
    }
    if (!rc && cudaGetLastError() != cudaSuccess) rc = tci_fail(ctx, TCI_ERR_CUDA, "lu_rdiv launch failed");
    return rc;
}

extern "C" int tci_lu_rdiv(tci_lu *lu, tci_dmat *B, double *out_host, tci_dmat **out_dev)
{
    if (!lu || !B) return TCI_ERR_ARG;
    tci_ctx *ctx = lu->ctx;
    TCI_ENTER(ctx);
    if (out_dev) *out_dev = nullptr;
    const i64 k = lu->r, rows = B->m;
    if (lu->m != lu->n || k != lu->m)
        return tci_fail(ctx, TCI_ERR_ARG, "tci_lu_rdiv: the factorised matrix must be square and of full rank");
    if (B->n != k) return tci_fail(ctx, TCI_ERR_ARG, "tci_lu_rdiv: DimensionMismatch between B and the factorised matrix");
    dmat_wait_ready(ctx, B);
    tci_dmat *res = nullptr;
    int rc = dmat_alloc(ctx, rows, k, &res);
    if (rc) return rc;
    if (rows == 0 || k == 0) return finish(ctx, res, out_host, out_dev);
    {
        StageTimer tm(ctx, ST_LUCI);
        rc = lu_rdiv_enqueue(lu, B->p, B->ld, rows, res->p, res->ld);
    }
    if (rc) {
        tci_dmat_destroy(res);
        return rc;
    }
    return finish(ctx, res, out_host, out_dev);
}
