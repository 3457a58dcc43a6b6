// zgemm.cu -- ComplexF64 GEMM on the FP64 tensor path: C = alpha * op(A) * op(B) + beta * C, op = identity or plain
// transpose (never conjugated: the reference's `_contract` only permutes, contraction.jl:71-93), alpha / beta real.
//
// The ComplexF64 value type of SURVEY 8f-4 (the reference's contraction tests are complex, test_contraction.jl:39-46)
// needs the same building block the Float64 path has in dgemm.cu.  Matrices are column-major arrays of interleaved
// (re, im) pairs -- the memory of a Julia Matrix{ComplexF64}.  A complex product is four real ones,
//     C_re += A_re B_re - A_im B_im,    C_im += A_re B_im + A_im B_re,
// issued as four mma.sync.m8n8k4.f64 (SASS DMMA) per 8 x 8 x 4 complex tile on fragments loaded ONCE as 16-byte
// (re, im) pairs: a complex FMA is 8 real flops for 32 bytes of operands, so the kernel is even more firmly bound by
// the FP64 tensor path than the real one.  CTA tile 64 x 64 x 8, 4 warps as 2 x 2, warp tile 32 x 32 = 4 x 4 complex
// DMMA tiles (128 accumulator registers); operands staged [row][k] with the k-row padded to 12 pairs: the quarter-warp
// of an LDS.128 then touches 8 different 16-byte banks.  Strided-batched with per-batch element offsets, like
// dgemm_dev_batched_off, so that the MPO environment chains (mpo.cu) run unchanged on complex cores.
#include "tci_internal.h"

#define ZK 8
#define ZKP 12

__device__ __forceinline__ void zdmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(128)
    k_zgemm_mma(i64 M, i64 N, i64 K, double alpha, const double2 *__restrict__ A, i64 lda, i64 strideA,
                const double2 *__restrict__ B, i64 ldb, i64 strideB, double beta, double2 *__restrict__ C, i64 ldc,
                i64 strideC, const i64 *__restrict__ offA, const i64 *__restrict__ offB)
{
    constexpr int BM = 64, BN = 64, NT = 128, TI = 4, TJ = 4;
    constexpr int LA = BM * ZK / NT, LB = BN * ZK / NT; // pairs each thread stages per k-tile
    __shared__ __align__(16) double2 As[BM][ZKP];
    __shared__ __align__(16) double2 Bs[BN][ZKP];
    A += strideA * blockIdx.z + (offA ? offA[blockIdx.z] : 0);
    B += strideB * blockIdx.z + (offB ? offB[blockIdx.z] : 0);
    C += strideC * blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    const i64 m0 = (i64)blockIdx.x * BM, n0 = (i64)blockIdx.y * BN;

    double cre[TI][TJ][2], cim[TI][TJ][2];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0;

    double2 ra[LA], rb[LB];
    const double2 zero = make_double2(0.0, 0.0);
    auto gload = [&](i64 k0) {
#pragma unroll
        for (int q = 0; q < LA; ++q) {
            const int e = tid + q * NT;
            const int mm = TA ? e / ZK : e % BM, kk = TA ? e % ZK : e / BM;
            const i64 gm = m0 + mm, gk = k0 + kk;
            ra[q] = (gm < M && gk < K) ? (TA ? A[gk + lda * gm] : A[gm + lda * gk]) : zero;
        }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            const int e = tid + q * NT;
            const int nn = TB ? e % BN : e / ZK, kk = TB ? e / BN : e % ZK;
            const i64 gn = n0 + nn, gk = k0 + kk;
            rb[q] = (gn < N && gk < K) ? (TB ? B[gn + ldb * gk] : B[gk + ldb * gn]) : zero;
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int q = 0; q < LA; ++q) {
            const int e = tid + q * NT;
            As[TA ? e / ZK : e % BM][TA ? e % ZK : e / BM] = ra[q];
        }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            const int e = tid + q * NT;
            Bs[TB ? e % BN : e / ZK][TB ? e / BN : e % ZK] = rb[q];
        }
    };

    gload(0);
    for (i64 k0 = 0; k0 < K; k0 += ZK) {
        __syncthreads();
        sstore();
        __syncthreads();
        if (k0 + ZK < K) gload(k0 + ZK);
#pragma unroll
        for (int k4 = 0; k4 < ZK; k4 += 4) {
            double2 af[TI], bf[TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i) af[i] = As[wm + 8 * i + fr][k4 + fk];
#pragma unroll
            for (int j = 0; j < TJ; ++j) bf[j] = Bs[wn + 8 * j + fr][k4 + fk];
#pragma unroll
            for (int i = 0; i < TI; ++i) {
                const double nim = -af[i].y;
#pragma unroll
                for (int j = 0; j < TJ; ++j) {
                    zdmma(cre[i][j][0], cre[i][j][1], af[i].x, bf[j].x);
                    zdmma(cim[i][j][0], cim[i][j][1], af[i].x, bf[j].y);
                    zdmma(cre[i][j][0], cre[i][j][1], nim, bf[j].y);
                    zdmma(cim[i][j][0], cim[i][j][1], af[i].y, bf[j].x);
                }
            }
        }
    }
    // accumulator fragment: row = lane/4, columns 2*(lane%4) + {0,1}
#pragma unroll
    for (int j = 0; j < TJ; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const i64 gn = n0 + wn + 8 * j + 2 * fk + c;
            if (gn >= N) continue;
#pragma unroll
            for (int i = 0; i < TI; ++i) {
                const i64 gm = m0 + wm + 8 * i + fr;
                if (gm >= M) continue;
                double2 *cp = C + gm + ldc * gn;
                double2 v = make_double2(alpha * cre[i][j][c], alpha * cim[i][j][c]);
                if (beta != 0.0) {
                    const double2 old = *cp;
                    v.x = fma(beta, old.x, v.x);
                    v.y = fma(beta, old.y, v.y);
                }
                *cp = v;
            }
        }
}

__global__ void k_zscale(double2 *C, i64 M, i64 N, i64 ldc, i64 strideC, double beta)
{
    double2 *c = C + strideC * blockIdx.z;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < M * N; e += (i64)gridDim.x * blockDim.x) {
        const i64 i = e % M, j = e / M;
        double2 v = c[i + ldc * j];
        c[i + ldc * j] = beta == 0.0 ? make_double2(0.0, 0.0) : make_double2(beta * v.x, beta * v.y);
    }
}

int zgemm_dev_batched_off(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double2 *A, i64 lda,
                          i64 strideA, const double2 *B, i64 ldb, i64 strideB, double beta, double2 *C, i64 ldc,
                          i64 strideC, i64 batch, const i64 *offA, const i64 *offB)
{
    if (M <= 0 || N <= 0 || batch <= 0) return TCI_OK;
    if (batch > 65535) { // gridDim.z limit
        for (i64 b0 = 0; b0 < batch; b0 += 65535) {
            const i64 nb = std::min<i64>(65535, batch - b0);
            int rc = zgemm_dev_batched_off(ctx, tA, tB, M, N, K, alpha, A + strideA * b0, lda, strideA, B + strideB * b0,
                                           ldb, strideB, beta, C + strideC * b0, ldc, strideC, nb,
                                           offA ? offA + b0 : nullptr, offB ? offB + b0 : nullptr);
            if (rc) return rc;
        }
        return TCI_OK;
    }
    if (K <= 0) {
        dim3 grid((unsigned)std::min<i64>((M * N + 255) / 256, 1024), 1, (unsigned)batch);
        k_zscale<<<grid, 256, 0, ctx->stream>>>(C, M, N, ldc, strideC, beta);
    } else {
        dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)batch);
        if (!tA && !tB)
            k_zgemm_mma<false, false><<<grid, 128, 0, ctx->stream>>>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, offA, offB);
        else if (tA && !tB)
            k_zgemm_mma<true, false><<<grid, 128, 0, ctx->stream>>>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, offA, offB);
        else if (!tA && tB)
            k_zgemm_mma<false, true><<<grid, 128, 0, ctx->stream>>>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, offA, offB);
        else
            k_zgemm_mma<true, true><<<grid, 128, 0, ctx->stream>>>(M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC, offA, offB);
    }
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

int zgemm_dev(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double2 *A, i64 lda,
              const double2 *B, i64 ldb, double beta, double2 *C, i64 ldc)
{
    return zgemm_dev_batched_off(ctx, tA, tB, M, N, K, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1, nullptr, nullptr);
}

// C = op(A) * op(B) for host Matrix{ComplexF64} arrays (interleaved, column-major, tight)
extern "C" int tci_zgemm_host(tci_ctx *ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, const double *A,
                              const double *B, double *C)
{
    TCI_ENTER(ctx);
    if (M < 0 || N < 0 || K < 0) return tci_fail(ctx, TCI_ERR_ARG, "tci_zgemm_host: negative size");
    if (M * N == 0) return TCI_OK;
    const i64 lda = transA ? K : M, ldb = transB ? N : K;
    DevBuf<double> dA(ctx), dB(ctx), dC(ctx);
    {
        StageTimer tm(ctx, ST_H2D);
        TCI_CUDA(ctx, dA.upload(A, (size_t)(2 * M * K)));
        TCI_CUDA(ctx, dB.upload(B, (size_t)(2 * K * N)));
        TCI_CUDA(ctx, dC.alloc((size_t)(2 * M * N)));
    }
    {
        StageTimer tm(ctx, ST_GEMM);
        int rc = zgemm_dev(ctx, transA != 0, transB != 0, M, N, K, 1.0, (const double2 *)dA.p, lda ? lda : 1,
                           (const double2 *)dB.p, ldb ? ldb : 1, 0.0, (double2 *)dC.p, M);
        if (rc) return rc;
    }
    StageTimer tm(ctx, ST_D2H);
    TCI_CUDA(ctx, cudaMemcpyAsync(C, dC.p, 2 * M * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}
