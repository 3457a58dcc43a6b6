// dgemm.cu -- FP64 GEMM on the DFMA pipe: C = alpha * op(A) * op(B) + beta * C.
//
// Replaces the OpenBLAS dgemm calls behind matrixluci.jl:41,45 (colmatrix/rowmatrix),
// cachedtensortrain.jl:207,212, contraction.jl:92 (_contract) on the hot path.
// tcgen05 has no FP64 path; B200's FP64 peak is the DFMA pipe (64 FMA/clk/SM), so this
// is a shared-memory tiled, register blocked kernel: BMxBNx16 tiles, 256 threads, each
// thread an (BM/16)x(BN/16) micro tile made of 2-wide strips 32 apart so that every
// LDS.128 of a quarter warp is conflict free and C stores are 256 B contiguous.
// The next k-tile is prefetched into registers while the current one is consumed.
#include <cstdlib>

#include "tci_internal.h"

#define GK 16

template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(256)
    k_dgemm(i64 M, i64 N, i64 K, double alpha, const double *__restrict__ A, i64 lda, i64 strideA,
            const double *__restrict__ B, i64 ldb, i64 strideB, double beta, double *__restrict__ C, i64 ldc,
            i64 strideC, const i64 *__restrict__ offA, const i64 *__restrict__ offB)
{
    constexpr int TM = BM / 16, TN = BN / 16;       // micro tile
    constexpr int SM = TM / 2, SN = TN / 2;         // number of 2-wide strips
    constexpr int LA = BM * GK / 256, LB = BN * GK / 256; // elements each thread stages per tile
    __shared__ __align__(16) double As[GK][BM];
    __shared__ __align__(16) double Bs[GK][BN];

    A += strideA * blockIdx.z + (offA ? offA[blockIdx.z] : 0);
    B += strideB * blockIdx.z + (offB ? offB[blockIdx.z] : 0);
    C += strideC * blockIdx.z;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const i64 m0 = (i64)blockIdx.x * BM, n0 = (i64)blockIdx.y * BN;

    double acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;

    double ra[LA], rb[LB];
    auto gload = [&](i64 k0) {
#pragma unroll
        for (int q = 0; q < LA; ++q) {
            int e = tid + q * 256;
            int mm, kk;
            if (!TA) { // A is M x K: m contiguous
                mm = e % BM;
                kk = e / BM;
            } else { // A is K x M: k contiguous
                kk = e % GK;
                mm = e / GK;
            }
            i64 gm = m0 + mm, gk = k0 + kk;
            ra[q] = (gm < M && gk < K) ? (TA ? A[gk + lda * gm] : A[gm + lda * gk]) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            int e = tid + q * 256;
            int nn, kk;
            if (!TB) { // B is K x N: k contiguous
                kk = e % GK;
                nn = e / GK;
            } else { // B is N x K: n contiguous
                nn = e % BN;
                kk = e / BN;
            }
            i64 gn = n0 + nn, gk = k0 + kk;
            rb[q] = (gn < N && gk < K) ? (TB ? B[gn + ldb * gk] : B[gk + ldb * gn]) : 0.0;
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int q = 0; q < LA; ++q) {
            int e = tid + q * 256;
            int mm = TA ? e / GK : e % BM, kk = TA ? e % GK : e / BM;
            As[kk][mm] = ra[q];
        }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            int e = tid + q * 256;
            int nn = TB ? e % BN : e / GK, kk = TB ? e / BN : e % GK;
            Bs[kk][nn] = rb[q];
        }
    };

    gload(0);
    for (i64 k0 = 0; k0 < K; k0 += GK) {
        __syncthreads();
        sstore();
        __syncthreads();
        if (k0 + GK < K) gload(k0 + GK);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double av[TM], bv[TN];
#pragma unroll
            for (int s = 0; s < SM; ++s) {
                double2 t = *reinterpret_cast<const double2 *>(&As[kk][2 * tx + 32 * s]);
                av[2 * s] = t.x;
                av[2 * s + 1] = t.y;
            }
#pragma unroll
            for (int s = 0; s < SN; ++s) {
                double2 t = *reinterpret_cast<const double2 *>(&Bs[kk][2 * ty + 32 * s]);
                bv[2 * s] = t.x;
                bv[2 * s + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int sj = 0; sj < SN; ++sj)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            i64 gn = n0 + 2 * ty + 32 * sj + jj;
            if (gn >= N) continue;
#pragma unroll
            for (int si = 0; si < SM; ++si)
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) {
                    i64 gm = m0 + 2 * tx + 32 * si + ii;
                    if (gm >= M) continue;
                    double v = alpha * acc[2 * si + ii][2 * sj + jj];
                    double *c = C + gm + ldc * gn;
                    *c = (beta == 0.0) ? v : fma(beta, *c, v);
                }
        }
}

// ---- FP64 tensor-core variant: mma.sync.m8n8k4.f64 (SASS DMMA) ------------------------------
// tcgen05 has no FP64 kind; DMMA through mma.sync is the only tensor path for doubles on sm_100a.
// CTA tile 128x128x16, 8 warps as 2 (M) x 4 (N), warp tile 64x32 = 8 x 4 DMMA tiles (64 accumulator
// doubles per thread).  Operands are staged as As[m][k] / Bs[n][k] with the k-row padded to 20 doubles,
// which makes the per-lane 8-byte fragment loads (row = lane/4, k = lane%4) bank-conflict free.
#define MK 16
#define MKP 20

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int BM, int BN, int NWM, int NWN, bool TA, bool TB>
__global__ void __launch_bounds__(32 * NWM * NWN)
    k_dgemm_mma(i64 M, i64 N, i64 K, double alpha, const double *__restrict__ A, i64 lda, i64 strideA,
                const double *__restrict__ B, i64 ldb, i64 strideB, double beta, double *__restrict__ C, i64 ldc,
                i64 strideC, const i64 *__restrict__ offA, const i64 *__restrict__ offB)
{
    constexpr int NT = 32 * NWM * NWN;            // threads
    constexpr int TI = BM / NWM / 8, TJ = BN / NWN / 8; // DMMA tiles per warp
    constexpr int LA = BM * MK / NT, LB = BN * MK / NT;
    __shared__ __align__(16) double As[BM][MKP];
    __shared__ __align__(16) double Bs[BN][MKP];
    A += strideA * blockIdx.z + (offA ? offA[blockIdx.z] : 0);
    B += strideB * blockIdx.z + (offB ? offB[blockIdx.z] : 0);
    C += strideC * blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp % NWM) * (BM / NWM), wn = (warp / NWM) * (BN / NWN); // warp tile origin
    const int fr = lane >> 2, fk = lane & 3;                                  // fragment row / k index
    const i64 m0 = (i64)blockIdx.x * BM, n0 = (i64)blockIdx.y * BN;

    double acc[TI][TJ][2];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    double ra[LA], rb[LB];
    auto gload = [&](i64 k0) {
#pragma unroll
        for (int q = 0; q < LA; ++q) {
            const int e = tid + q * NT;
            const int mm = TA ? e / MK : e % BM, kk = TA ? e % MK : e / BM;
            const i64 gm = m0 + mm, gk = k0 + kk;
            ra[q] = (gm < M && gk < K) ? (TA ? A[gk + lda * gm] : A[gm + lda * gk]) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            const int e = tid + q * NT;
            const int nn = TB ? e % BN : e / MK, kk = TB ? e / BN : e % MK;
            const i64 gn = n0 + nn, gk = k0 + kk;
            rb[q] = (gn < N && gk < K) ? (TB ? B[gn + ldb * gk] : B[gk + ldb * gn]) : 0.0;
        }
    };
    auto sstore = [&]() {
#pragma unroll
        for (int q = 0; q < LA; ++q) {
            const int e = tid + q * NT;
            As[TA ? e / MK : e % BM][TA ? e % MK : e / BM] = ra[q];
        }
#pragma unroll
        for (int q = 0; q < LB; ++q) {
            const int e = tid + q * NT;
            Bs[TB ? e % BN : e / MK][TB ? e / BN : e % MK] = rb[q];
        }
    };

    gload(0);
    for (i64 k0 = 0; k0 < K; k0 += MK) {
        __syncthreads();
        sstore();
        __syncthreads();
        if (k0 + MK < K) gload(k0 + MK);
#pragma unroll
        for (int k4 = 0; k4 < MK; k4 += 4) {
            double af[TI], bf[TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i) af[i] = As[wm + 8 * i + fr][k4 + fk];
#pragma unroll
            for (int j = 0; j < TJ; ++j) bf[j] = Bs[wn + 8 * j + fr][k4 + fk];
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
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
                const double v = alpha * acc[i][j][c];
                double *cp = C + gm + ldc * gn;
                *cp = (beta == 0.0) ? v : fma(beta, *cp, v);
            }
        }
}

// ---- DMMA kernel with a 3-stage cp.async pipeline ---------------------------------------------
// Same tiling as k_dgemm_mma, but the operand tiles go global -> shared with 16-byte cp.async (no
// register staging, one __syncthreads per k-tile, two tiles in flight while the third is consumed).
// The shared layout follows the contiguous direction of each operand in global memory so that every
// 16-byte chunk is two adjacent doubles on both sides:
//   A not transposed (m contiguous): As[k][BM+4]      A transposed (k contiguous): As[m][MK+4]
//   B not transposed (k contiguous): Bs[n][MK+4]      B transposed (n contiguous): Bs[k][BN+4]
// (+4 padding: the 8-byte fragment loads of a half warp hit 16 different bank pairs in all four cases).
// Needs 16-byte aligned operands: even leading dimensions / strides / offsets, checked by the launcher.
#define ASTAGES 3

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int BM, int BN, int NWM, int NWN, bool TA, bool TB, int KT = MK, int NST = ASTAGES>
__global__ void __launch_bounds__(32 * NWM * NWN)
    k_dgemm_mma_async(i64 M, i64 N, i64 K, double alpha, const double *__restrict__ A, i64 lda, i64 strideA,
                      const double *__restrict__ B, i64 ldb, i64 strideB, double beta, double *__restrict__ C, i64 ldc,
                      i64 strideC, const i64 *__restrict__ offA, const i64 *__restrict__ offB)
{
    constexpr int NT = 32 * NWM * NWN;
    constexpr int TI = BM / NWM / 8, TJ = BN / NWN / 8;
    constexpr int A_ROW = TA ? KT + 4 : BM + 4, A_ROWS = TA ? BM : KT; // shared tile of A: A_ROWS x A_ROW
    constexpr int B_ROW = TB ? BN + 4 : KT + 4, B_ROWS = TB ? KT : BN;
    constexpr int A_SZ = A_ROW * A_ROWS, B_SZ = B_ROW * B_ROWS;
    extern __shared__ __align__(16) double dsm[];
    double *const Asm = dsm;                   // [NST][A_SZ]
    double *const Bsm = dsm + NST * A_SZ;  // [NST][B_SZ]

    A += strideA * blockIdx.z + (offA ? offA[blockIdx.z] : 0);
    B += strideB * blockIdx.z + (offB ? offB[blockIdx.z] : 0);
    C += strideC * blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp % NWM) * (BM / NWM), wn = (warp / NWM) * (BN / NWN);
    const int fr = lane >> 2, fk = lane & 3;
    const i64 m0 = (i64)blockIdx.x * BM, n0 = (i64)blockIdx.y * BN;

    auto load_stage = [&](int stage, i64 k0) {
        double *as = Asm + stage * A_SZ, *bs = Bsm + stage * B_SZ;
        // A: chunks of two doubles along the contiguous direction
#pragma unroll
        for (int c = tid; c < BM * KT / 2; c += NT) {
            if (!TA) { // pairs along m
                const int mm = (c % (BM / 2)) * 2, kk = c / (BM / 2);
                const i64 gm = m0 + mm, gk = k0 + kk;
                const int bytes = (gk < K && gm < M) ? (gm + 1 < M ? 16 : 8) : 0;
                cp_async16(as + kk * A_ROW + mm, A + (bytes ? gm + lda * gk : 0), bytes);
            } else { // pairs along k
                const int kk = (c % (KT / 2)) * 2, mm = c / (KT / 2);
                const i64 gm = m0 + mm, gk = k0 + kk;
                const int bytes = (gm < M && gk < K) ? (gk + 1 < K ? 16 : 8) : 0;
                cp_async16(as + mm * A_ROW + kk, A + (bytes ? gk + lda * gm : 0), bytes);
            }
        }
#pragma unroll
        for (int c = tid; c < BN * KT / 2; c += NT) {
            if (!TB) { // B is K x N: pairs along k
                const int kk = (c % (KT / 2)) * 2, nn = c / (KT / 2);
                const i64 gn = n0 + nn, gk = k0 + kk;
                const int bytes = (gn < N && gk < K) ? (gk + 1 < K ? 16 : 8) : 0;
                cp_async16(bs + nn * B_ROW + kk, B + (bytes ? gk + ldb * gn : 0), bytes);
            } else { // B is N x K: pairs along n
                const int nn = (c % (BN / 2)) * 2, kk = c / (BN / 2);
                const i64 gn = n0 + nn, gk = k0 + kk;
                const int bytes = (gk < K && gn < N) ? (gn + 1 < N ? 16 : 8) : 0;
                cp_async16(bs + kk * B_ROW + nn, B + (bytes ? gn + ldb * gk : 0), bytes);
            }
        }
    };

    double acc[TI][TJ][2];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const i64 nkt = (K + KT - 1) / KT;
    // Interior CTAs (the whole BM x BN tile inside the matrix, K a multiple of KT) copy through per-thread base
    // pointers that only advance by a constant per k-tile: a thread's chunks c = tid + q*NT differ by a fixed number
    // of columns (rows) of the operand, so the ~300 integer instructions of the bounds-checked path per k-tile shrink
    // to one 64-bit add per copy.  (The DMMA stream of a warp pauses for the copies; two warps share a scheduler.)
    const bool interior = m0 + BM <= M && n0 + BN <= N && K % KT == 0;
    constexpr int ACH = BM * KT / 2, BCH = BN * KT / 2;
    constexpr int AQ = ACH / NT, BQ = BCH / NT;
    static_assert(ACH % NT == 0 && BCH % NT == 0, "chunks must divide evenly over the threads");
    // chunk (tid, q) of A: !TA: mm = (tid % (BM/2))*2, kk = tid / (BM/2) + q*(NT/(BM/2)); TA: kk = (tid % (KT/2))*2, mm = tid/(KT/2) + q*(NT/(KT/2))
    constexpr int A_DIV = TA ? KT / 2 : BM / 2, B_DIV = TB ? BN / 2 : KT / 2;
    static_assert(NT % A_DIV == 0 && NT % B_DIV == 0, "thread count must be a multiple of the chunk row length");
    const int a_in = (tid % A_DIV) * 2, a_out = tid / A_DIV; // contiguous-direction element, other-direction index
    const int b_in = (tid % B_DIV) * 2, b_out = tid / B_DIV;
    const double *gA = TA ? A + (a_in + lda * (m0 + a_out)) : A + (m0 + a_in + lda * a_out);
    const double *gB = TB ? B + (n0 + b_in + ldb * b_out) : B + (b_in + ldb * (n0 + b_out));
    const i64 a_qstep = lda * (NT / A_DIV), b_qstep = ldb * (NT / B_DIV);     // between the chunks of one thread
    const i64 a_kstep = TA ? KT : (i64)KT * lda, b_kstep = TB ? (i64)KT * ldb : KT; // between k-tiles
    const int a_soff = a_out * A_ROW + a_in, b_soff = b_out * B_ROW + b_in;   // shared offsets of chunk q = 0
    auto load_fast = [&](int stage, i64 kt) {
        double *as = Asm + stage * A_SZ + a_soff, *bs = Bsm + stage * B_SZ + b_soff;
        const double *pa = gA + kt * a_kstep, *pb = gB + kt * b_kstep;
#pragma unroll
        for (int q = 0; q < AQ; ++q) cp_async16(as + q * (NT / A_DIV) * A_ROW, pa + q * a_qstep, 16);
#pragma unroll
        for (int q = 0; q < BQ; ++q) cp_async16(bs + q * (NT / B_DIV) * B_ROW, pb + q * b_qstep, 16);
    };
#pragma unroll
    for (int st = 0; st < NST - 1; ++st) {
        if (st < nkt) {
            if (interior)
                load_fast(st, st);
            else
                load_stage(st, (i64)st * KT);
        }
        cp_async_commit();
    }
    for (i64 kt = 0; kt < nkt; ++kt) {
        cp_async_wait<NST - 2>(); // tile kt has landed
        __syncthreads();              // ... for everybody, and everybody is done with tile kt-1
        if (kt + NST - 1 < nkt) {
            if (interior)
                load_fast((int)((kt + NST - 1) % NST), kt + NST - 1);
            else
                load_stage((int)((kt + NST - 1) % NST), (kt + NST - 1) * KT);
        }
        cp_async_commit();
        const double *as = Asm + (kt % NST) * A_SZ, *bs = Bsm + (kt % NST) * B_SZ;
#pragma unroll
        for (int k4 = 0; k4 < KT; k4 += 4) {
            double af[TI], bf[TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i)
                af[i] = TA ? as[(wm + 8 * i + fr) * A_ROW + k4 + fk] : as[(k4 + fk) * A_ROW + wm + 8 * i + fr];
#pragma unroll
            for (int j = 0; j < TJ; ++j)
                bf[j] = TB ? bs[(k4 + fk) * B_ROW + wn + 8 * j + fr] : bs[(wn + 8 * j + fr) * B_ROW + k4 + fk];
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
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
                const double v = alpha * acc[i][j][c];
                double *cp = C + gm + ldc * gn;
                *cp = (beta == 0.0) ? v : fma(beta, *cp, v);
            }
        }
}

// ---- tile-stream variant for short inner dimensions ---------------------------------------------------------
// The environment chains of a contraction Pi are batches of ~1000 products with K = 256 or 512: a 128 x 64 tile then has
// only 16 (32) k-tiles, and the pipeline fill at its start and the drain at its end are a visible fraction of it.  Here
// a CTA walks TPC consecutive tiles as ONE stream of k-tiles: the copies of the next tile's first k-tiles are issued
// while the last k-tiles of the current one are multiplied, so the DMMA stream only pauses for the accumulator
// stores.  Interior tiles only (M % 128 == 0, N % 64 == 0, K % 16 == 0; the launcher checks), 128 x 64 x 16 tiles,
// 4 warps, 3 stages, two CTAs per SM as in k_dgemm_mma_async<128, 64, 2, 2>.
template <bool TA, bool TB>
__global__ void __launch_bounds__(128)
    k_dgemm_mma_stream(i64 M, i64 N, i64 K, double alpha, const double *__restrict__ A, i64 lda, i64 strideA,
                       const double *__restrict__ B, i64 ldb, i64 strideB, double beta, double *__restrict__ C, i64 ldc,
                       i64 strideC, const i64 *__restrict__ offA, const i64 *__restrict__ offB, int mx, int ny, i64 total, int TPC)
{
    constexpr int BM = 128, BN = 64, NWM = 2, KT = 16, NST = 3, NT = 128;
    constexpr int TI = BM / NWM / 8, TJ = BN / 2 / 8;
    constexpr int A_ROW = TA ? KT + 4 : BM + 4, A_ROWS = TA ? BM : KT;
    constexpr int B_ROW = TB ? BN + 4 : KT + 4, B_ROWS = TB ? KT : BN;
    constexpr int A_SZ = A_ROW * A_ROWS, B_SZ = B_ROW * B_ROWS;
    constexpr int AQ = BM * KT / 2 / NT, BQ = BN * KT / 2 / NT;
    constexpr int A_DIV = TA ? KT / 2 : BM / 2, B_DIV = TB ? BN / 2 : KT / 2;
    extern __shared__ __align__(16) double dsm[];
    double *const Asm = dsm, *const Bsm = dsm + NST * A_SZ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp % NWM) * (BM / NWM), wn = (warp / NWM) * (BN / 2);
    const int fr = lane >> 2, fk = lane & 3;
    const i64 t0 = (i64)blockIdx.x * TPC;
    const int nt = (int)(total - t0 < TPC ? total - t0 : TPC);
    const int nkt = (int)(K / KT), F = nt * nkt;
    const int a_in = (tid % A_DIV) * 2, a_out = tid / A_DIV, b_in = (tid % B_DIV) * 2, b_out = tid / B_DIV;
    const i64 a_qstep = lda * (NT / A_DIV), b_qstep = ldb * (NT / B_DIV);
    const i64 a_kstep = TA ? KT : (i64)KT * lda, b_kstep = TB ? (i64)KT * ldb : KT;
    const int a_soff = a_out * A_ROW + a_in, b_soff = b_out * B_ROW + b_in;
    const int mxy = mx * ny;

    // producer cursor: tile lt, k-tile lk, this thread's operand pointers for tile lt
    int lt = 0, lk = 0;
    const double *gA = nullptr, *gB = nullptr;
    auto set_tile = [&](int t) {
        const i64 g = t0 + t;
        const i64 z = g / mxy;
        const int r = (int)(g - z * mxy), y = r / mx, x = r - y * mx;
        const double *Az = A + strideA * z + (offA ? offA[z] : 0), *Bz = B + strideB * z + (offB ? offB[z] : 0);
        const i64 m0 = (i64)x * BM, n0 = (i64)y * BN;
        gA = TA ? Az + (a_in + lda * (m0 + a_out)) : Az + (m0 + a_in + lda * a_out);
        gB = TB ? Bz + (n0 + b_in + ldb * b_out) : Bz + (b_in + ldb * (n0 + b_out));
    };
    auto produce = [&](int stage) { // copies of (lt, lk) into `stage`, then the cursor advances
        if (lt < nt) {
            double *as = Asm + stage * A_SZ + a_soff, *bs = Bsm + stage * B_SZ + b_soff;
            const double *pa = gA + lk * a_kstep, *pb = gB + lk * b_kstep;
#pragma unroll
            for (int q = 0; q < AQ; ++q) cp_async16(as + q * (NT / A_DIV) * A_ROW, pa + q * a_qstep, 16);
#pragma unroll
            for (int q = 0; q < BQ; ++q) cp_async16(bs + q * (NT / B_DIV) * B_ROW, pb + q * b_qstep, 16);
            if (++lk == nkt) {
                lk = 0;
                if (++lt < nt) set_tile(lt);
            }
        }
        cp_async_commit();
    };
    set_tile(0);
#pragma unroll
    for (int st = 0; st < NST - 1; ++st) produce(st);

    double acc[TI][TJ][2];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    int ct = 0, ck = 0;
    for (int f = 0; f < F; ++f) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        produce((f + NST - 1) % NST);
        const double *as = Asm + (f % NST) * A_SZ, *bs = Bsm + (f % NST) * B_SZ;
        // fragments of k-step k4 + 4 are fetched while the DMMAs of k-step k4 issue (explicit double buffering)
        double af[2][TI], bf[2][TJ];
        auto lfrag = [&](int buf, int k4) {
#pragma unroll
            for (int i = 0; i < TI; ++i)
                af[buf][i] = TA ? as[(wm + 8 * i + fr) * A_ROW + k4 + fk] : as[(k4 + fk) * A_ROW + wm + 8 * i + fr];
#pragma unroll
            for (int j = 0; j < TJ; ++j)
                bf[buf][j] = TB ? bs[(k4 + fk) * B_ROW + wn + 8 * j + fr] : bs[(wn + 8 * j + fr) * B_ROW + k4 + fk];
        };
        lfrag(0, 0);
#pragma unroll
        for (int k4 = 0; k4 < KT; k4 += 4) {
            const int cur = (k4 >> 2) & 1;
            if (k4 + 4 < KT) lfrag(cur ^ 1, k4 + 4);
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
        if (++ck == nkt) { // tile ct is complete: store it and start the next accumulator
            const i64 g = t0 + ct;
            const i64 z = g / mxy;
            const int r = (int)(g - z * mxy), y = r / mx, x = r - y * mx;
            double *Cz = C + strideC * z + ((i64)x * BM + wm + fr) + ldc * ((i64)y * BN + wn + 2 * fk);
#pragma unroll
            for (int j = 0; j < TJ; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int i = 0; i < TI; ++i) {
                        double *cp = Cz + 8 * i + ldc * (8 * j + c);
                        const double v = alpha * acc[i][j][c];
                        *cp = (beta == 0.0) ? v : fma(beta, *cp, v);
                        acc[i][j][c] = 0.0;
                    }
            ck = 0;
            ++ct;
        }
    }
}

static int launch_dgemm_stream(tci_ctx *ctx, int TPC, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
                               i64 sA, const double *B, i64 ldb, i64 sB, double beta, double *C, i64 ldc, i64 sC, i64 batch,
                               const i64 *offA, const i64 *offB)
{
    const int mx = (int)(M / 128), ny = (int)(N / 64);
    const i64 total = (i64)mx * ny * batch;
    const unsigned grid = (unsigned)((total + TPC - 1) / TPC);
    constexpr int KT = 16, NST = 3;
    const size_t smemTT = (size_t)NST * ((KT + 4) * 128 + (64 + 4) * KT) * 8, smemTF = (size_t)NST * ((KT + 4) * 128 + (KT + 4) * 64) * 8,
                 smemFT = (size_t)NST * ((128 + 4) * KT + (64 + 4) * KT) * 8, smemFF = (size_t)NST * ((128 + 4) * KT + (KT + 4) * 64) * 8;
#define TCI_STREAM_LAUNCH(TA_, TB_, SM_)                                                                                   \
    {                                                                                                                      \
        auto fn = k_dgemm_mma_stream<TA_, TB_>;                                                                       \
        TCI_CUDA(ctx, ctx_func_smem(ctx, (const void *)fn, (int)(SM_)));                                                   \
        fn<<<grid, 128, (SM_), ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB, mx, ny, \
                                              total, TPC);                                                                      \
    }
    if (tA && tB)
        TCI_STREAM_LAUNCH(true, true, smemTT)
    else if (tA && !tB)
        TCI_STREAM_LAUNCH(true, false, smemTF)
    else if (!tA && tB)
        TCI_STREAM_LAUNCH(false, true, smemFT)
    else
        TCI_STREAM_LAUNCH(false, false, smemFF)
#undef TCI_STREAM_LAUNCH
    ctx->launches++;
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

// (A variant fed by cp.async.bulk copies through an mbarrier ring -- one bulk copy per contiguous tile row, no CTA-wide
// barrier in the main loop -- was built and measured in round 2: 26.0 TFLOP/s at 4096^3 against 29.9 for the cp.async
// kernel below, also with both operands in 1 KB rows (24.2), so it was dropped.  What did help is halving the CTA:
// 128 x 64 tiles with 4 warps leave room for TWO CTAs per SM, which drift apart so that one issues DMMAs while the
// other sits at its barrier or issues its copies: 33.0 TFLOP/s at 4096^3, see dgemm_dev_batched_off.)
template <int BM, int BN, int NWM, int NWN, bool TA, bool TB, int KT = MK, int NST = ASTAGES>
static int launch_async_one(tci_ctx *ctx, dim3 grid, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda, i64 sA,
                            const double *B, i64 ldb, i64 sB, double beta, double *C, i64 ldc, i64 sC,
                            const i64 *offA, const i64 *offB)
{
    constexpr int A_SZ = (TA ? KT + 4 : BM + 4) * (TA ? BM : KT), B_SZ = (TB ? BN + 4 : KT + 4) * (TB ? KT : BN);
    constexpr size_t smem = (size_t)NST * (A_SZ + B_SZ) * sizeof(double);
    auto fn = k_dgemm_mma_async<BM, BN, NWM, NWN, TA, TB, KT, NST>;
    TCI_CUDA(ctx, ctx_func_smem(ctx, (const void *)fn, (int)smem));
    fn<<<grid, 32 * NWM * NWN, smem, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    ctx->launches++;
    return TCI_OK;
}

template <int BM, int BN, int NWM, int NWN, int KT = MK, int NST = ASTAGES>
static int launch_dgemm_mma_async(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A,
                                  i64 lda, i64 sA, const double *B, i64 ldb, i64 sB, double beta, double *C, i64 ldc,
                                  i64 sC, i64 batch, const i64 *offA, const i64 *offB)
{
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)batch);
    if (!tA && !tB)
        return launch_async_one<BM, BN, NWM, NWN, false, false, KT, NST>(ctx, grid, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    if (tA && !tB)
        return launch_async_one<BM, BN, NWM, NWN, true, false, KT, NST>(ctx, grid, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    if (!tA && tB)
        return launch_async_one<BM, BN, NWM, NWN, false, true, KT, NST>(ctx, grid, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    return launch_async_one<BM, BN, NWM, NWN, true, true, KT, NST>(ctx, grid, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
}

template <int BM, int BN, int NWM, int NWN>
static void launch_dgemm_mma(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A,
                             i64 lda, i64 sA, const double *B, i64 ldb, i64 sB, double beta, double *C, i64 ldc,
                             i64 sC, i64 batch, const i64 *offA, const i64 *offB)
{
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)batch);
    constexpr int NT = 32 * NWM * NWN;
    if (!tA && !tB)
        k_dgemm_mma<BM, BN, NWM, NWN, false, false><<<grid, NT, 0, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    else if (tA && !tB)
        k_dgemm_mma<BM, BN, NWM, NWN, true, false><<<grid, NT, 0, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    else if (!tA && tB)
        k_dgemm_mma<BM, BN, NWM, NWN, false, true><<<grid, NT, 0, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    else
        k_dgemm_mma<BM, BN, NWM, NWN, true, true><<<grid, NT, 0, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    ctx->launches++;
}

template <int BM, int BN>
static void launch_dgemm(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
                         i64 sA, const double *B, i64 ldb, i64 sB, double beta, double *C, i64 ldc, i64 sC, i64 batch,
                         const i64 *offA, const i64 *offB)
{
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)batch);
    if (!tA && !tB)
        k_dgemm<BM, BN, false, false><<<grid, 256, 0, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    else if (tA && !tB)
        k_dgemm<BM, BN, true, false><<<grid, 256, 0, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    else if (!tA && tB)
        k_dgemm<BM, BN, false, true><<<grid, 256, 0, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    else
        k_dgemm<BM, BN, true, true><<<grid, 256, 0, ctx->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, offA, offB);
    ctx->launches++;
}

__global__ void k_scale(double *C, i64 M, i64 N, i64 ldc, i64 strideC, double beta)
{
    double *c = C + strideC * blockIdx.z;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < M * N; e += (i64)gridDim.x * blockDim.x) {
        i64 i = e % M, j = e / M;
        c[i + ldc * j] = beta == 0.0 ? 0.0 : beta * c[i + ldc * j];
    }
}

// C = alpha * sum_s part[s] + beta * C, slices added in order
__global__ void k_splitk_reduce(const double *__restrict__ part, i64 M, i64 N, int S, double alpha, double beta,
                                double *__restrict__ C, i64 ldc)
{
    const i64 MN = M * N;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < MN; e += (i64)gridDim.x * blockDim.x) {
        double acc = part[e];
        for (int s = 1; s < S; ++s) acc += part[e + MN * s];
        double *c = C + e % M + ldc * (e / M);
        *c = beta == 0.0 ? alpha * acc : fma(beta, *c, alpha * acc);
    }
}
// number of k-slices for a single (unbatched) product, 0 = do not split: the output tiles must cover less than two
// waves of the 128 x 64 kernel (two CTAs per SM) and every slice keeps at least 1024 of the inner dimension
static int splitk_slices(tci_ctx *ctx, i64 M, i64 N, i64 K, i64 batch, const i64 *offA, const i64 *offB)
{
    static const int enabled = getenv("TCI_DGEMM_NO_SPLITK") ? 0 : 1;
    if (!enabled || batch != 1 || offA || offB || K < 4096 || M < 32 || N < 32) return 0;
    const i64 tiles = ((M + 127) / 128) * ((N + 63) / 64);
    const i64 want = 4 * (i64)ctx->sm_count; // two waves of two CTAs per SM
    if (tiles >= want / 2) return 0;
    const i64 S = std::min<i64>((want + tiles - 1) / tiles, K / 1024);
    return S >= 2 ? (int)S : 0;
}

int dgemm_dev_batched_off(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
                          i64 strideA, const double *B, i64 ldb, i64 strideB, double beta, double *C, i64 ldc,
                          i64 strideC, i64 batch, const i64 *offA, const i64 *offB, bool offsets_even)
{
    if (M <= 0 || N <= 0 || batch <= 0) return TCI_OK;
    if (batch > 65535) { // gridDim.z limit
        for (i64 b0 = 0; b0 < batch; b0 += 65535) {
            i64 nb = std::min<i64>(65535, batch - b0);
            int rc = dgemm_dev_batched_off(ctx, tA, tB, M, N, K, alpha, A + strideA * b0, lda, strideA,
                                           B + strideB * b0, ldb, strideB, beta, C + strideC * b0, ldc, strideC, nb,
                                           offA ? offA + b0 : nullptr, offB ? offB + b0 : nullptr, offsets_even);
            if (rc) return rc;
        }
        return TCI_OK;
    }
    if (K <= 0) {
        dim3 grid((unsigned)std::min<i64>((M * N + 255) / 256, 1024), 1, (unsigned)batch);
        k_scale<<<grid, 256, 0, ctx->stream>>>(C, M, N, ldc, strideC, beta);
        ctx->launches++;
    } else if (int S = splitk_slices(ctx, M, N, K, batch, offA, offB)) {
        // Split-K.  A single product with few output tiles and a long inner dimension -- the last step of a contraction
        // Pi, (nL x Da*Db)(Da*Db x nR) with Da*Db = 65536: 64 tiles of 128 x 128 at nL = nR = 1024, 8 for the row block
        // of one GPU out of eight -- leaves most SMs idle.  The inner dimension is cut into S slices that run as one
        // strided-batched launch into S partial products, summed in slice order by k_splitk_reduce (deterministic).
        const i64 Ks = round_up((K + S - 1) / S, 32);
        S = (int)((K + Ks - 1) / Ks);
        DevBuf<double> part(ctx);
        TCI_CUDA(ctx, part.alloc((size_t)(M * N) * (size_t)S));
        const i64 kA = tA ? Ks : lda * Ks, kB = tB ? ldb * Ks : Ks;
        int rc = TCI_OK;
        if (S > 1)
            rc = dgemm_dev_batched_off(ctx, tA, tB, M, N, Ks, 1.0, A, lda, kA, B, ldb, kB, 0.0, part.p, M, M * N, S - 1,
                                       nullptr, nullptr, false);
        if (!rc)
            rc = dgemm_dev_batched_off(ctx, tA, tB, M, N, K - Ks * (S - 1), 1.0, A + kA * (S - 1), lda, 0,
                                       B + kB * (S - 1), ldb, 0, 0.0, part.p + M * N * (S - 1), M, 0, 1, nullptr, nullptr,
                                       false);
        if (rc) return rc;
        k_splitk_reduce<<<(unsigned)std::min<i64>((M * N + 255) / 256, (i64)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
            part.p, M, N, S, alpha, beta, C, ldc);
        ctx->launches++;
    } else {
        i64 big_ctas = ((M + 127) / 128) * ((N + 127) / 128) * batch;
        static const int use_mma = getenv("TCI_DGEMM_NO_MMA") ? 0 : 1;
        const i64 mid_ctas = ((M + 63) / 64) * ((N + 63) / 64) * batch;
        static const int force_tile = getenv("TCI_DGEMM_TILE") ? atoi(getenv("TCI_DGEMM_TILE")) : 0;
        if (use_mma && force_tile == 64)
            launch_dgemm_mma<64, 64, 2, 2>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,
                                           strideC, batch, offA, offB);
        else if (use_mma && force_tile == 12864)
            launch_dgemm_mma<128, 64, 4, 2>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,
                                            strideC, batch, offA, offB);
        else if (use_mma && force_tile == 6432)
            launch_dgemm_mma<64, 32, 2, 2>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,
                                           strideC, batch, offA, offB);
        if (use_mma && force_tile) {
            TCI_CUDA(ctx, cudaGetLastError());
            return TCI_OK;
        }
        // cp.async pipeline variant: operands must be 16-byte aligned along their contiguous direction
        const bool aligned16 = (((size_t)A | (size_t)B) & 15) == 0 && !((lda | ldb | strideA | strideB) & 1) &&
                               ((!offA && !offB) || offsets_even);
        static const int use_async = getenv("TCI_DGEMM_NO_ASYNC") ? 0 : 1;
        static const int variant = getenv("TCI_DGEMM_VARIANT") ? atoi(getenv("TCI_DGEMM_VARIANT")) : 0;
        // 128 x 64 tiles, 4 warps per CTA, TWO CTAs per SM: the CTAs drift apart, so one issues DMMAs while the other
        // sits at its barrier or issues its copies (measured: 4096^3 29.9 -> 33.0 TFLOP/s, 8192^2 x 512 28.7 -> 31.5, the
        // config-5 MPO Pi 27.6 -> 29.4; with fewer than ~4 CTAs per SM the 8-warp 128 x 128 kernel stays ahead:
        // 2048^2 x 256 24.9 vs 22.7).  Also tried: 16 warps of 32 x 32 (28.2), 32-deep k-tiles (30.8 with 8 warps, 25.9
        // with 4: shared memory then allows one CTA per SM), 4 stages (32.4), bulk copies + mbarriers (26.0).
        const i64 half_ctas = ((M + 127) / 128) * ((N + 63) / 64) * batch;
        // 32-deep k-tiles with two stages (same shared memory, half as many barriers and copy phases per flop) win once
        // the k-loop is long: 4096^3 33.5 -> 34.1 TFLOP/s, 8192^2 x 512 31.9 -> 32.3; at K = 256 (the MPO environment
        // steps) the longer pipeline fill loses: config-5 Pi 145.3 vs 147.6 ms
        // short inner dimension, many interior tiles: the tile-stream kernel (see k_dgemm_mma_stream)
        static const int stream_tpc = getenv("TCI_DGEMM_STREAM") ? atoi(getenv("TCI_DGEMM_STREAM")) : 2;
        if (stream_tpc > 0 && use_mma && use_async && aligned16 && M % 128 == 0 && N % 64 == 0 && K % 16 == 0 && K <= 1024 &&
            !(ldc & 1) && half_ctas >= 4 * (i64)ctx->sm_count) {
            // two tiles per CTA only when that still leaves many waves (a sharded chain level has ~1000 tiles)
            const int tpc = half_ctas >= 16 * (i64)ctx->sm_count ? stream_tpc : 1;
            return launch_dgemm_stream(ctx, tpc, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,
                                       strideC, batch, offA, offB);
        }
        static const i64 k32_min = getenv("TCI_DGEMM_K32_MIN") ? atoll(getenv("TCI_DGEMM_K32_MIN")) : 1024;
        if ((variant == 2 || (variant == 0 && K >= k32_min)) && use_mma && use_async && aligned16 && M >= 96 && N >= 64 && half_ctas >= 4 * (i64)ctx->sm_count) {
            int rc = launch_dgemm_mma_async<128, 64, 2, 2, 32, 2>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB,
                                                                  beta, C, ldc, strideC, batch, offA, offB);
            if (rc) return rc;
            TCI_CUDA(ctx, cudaGetLastError());
            return TCI_OK;
        }
        if (variant != 9 && use_mma && use_async && aligned16 && M >= 96 && N >= 64 && half_ctas >= 4 * (i64)ctx->sm_count) {
            int rc = launch_dgemm_mma_async<128, 64, 2, 2>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB,
                                                           beta, C, ldc, strideC, batch, offA, offB);
            if (rc) return rc;
            TCI_CUDA(ctx, cudaGetLastError());
            return TCI_OK;
        }
        if (use_mma && use_async && aligned16 && !force_tile && big_ctas >= 2 * ctx->sm_count && M >= 96 && N >= 96) {
            int rc = launch_dgemm_mma_async<128, 128, 2, 4>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB,
                                                            beta, C, ldc, strideC, batch, offA, offB);
            if (rc) return rc;
        } else if (use_mma && use_async && aligned16 && !force_tile && mid_ctas >= 16 && M >= 32 && N >= 32) {
            int rc = launch_dgemm_mma_async<64, 64, 2, 2>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB,
                                                          beta, C, ldc, strideC, batch, offA, offB);
            if (rc) return rc;
        } else if (use_mma && big_ctas >= 2 * ctx->sm_count && M >= 96 && N >= 96)
            launch_dgemm_mma<128, 128, 2, 4>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C,
                                             ldc, strideC, batch, offA, offB);
        else if (use_mma && mid_ctas >= 16 && M >= 32 && N >= 32)
            launch_dgemm_mma<64, 64, 2, 2>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,
                                           strideC, batch, offA, offB);
        else if (big_ctas >= ctx->sm_count && M >= 96 && N >= 96)
            launch_dgemm<128, 128>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,
                                   strideC, batch, offA, offB);
        else
            launch_dgemm<64, 64>(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC,
                                 batch, offA, offB);
    }
    TCI_CUDA(ctx, cudaGetLastError());
    return TCI_OK;
}

int dgemm_dev_batched(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
                      i64 strideA, const double *B, i64 ldb, i64 strideB, double beta, double *C, i64 ldc, i64 strideC,
                      i64 batch)
{
    return dgemm_dev_batched_off(ctx, tA, tB, M, N, K, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc, strideC,
                                 batch, nullptr, nullptr, false);
}

int dgemm_dev(tci_ctx *ctx, bool tA, bool tB, i64 M, i64 N, i64 K, double alpha, const double *A, i64 lda,
              const double *B, i64 ldb, double beta, double *C, i64 ldc)
{
    return dgemm_dev_batched(ctx, tA, tB, M, N, K, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1);
}

extern "C" int tci_dgemm_host(tci_ctx *ctx, int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha,
                              const double *A, const double *B, double beta, double *C)
{
    TCI_ENTER(ctx);
    if (M < 0 || N < 0 || K < 0) return tci_fail(ctx, TCI_ERR_ARG, "tci_dgemm_host: negative size");
    if (M * N == 0) return TCI_OK;
    const i64 lda = transA ? K : M, ldb = transB ? N : K;
    DevBuf<double> dA(ctx), dB(ctx), dC(ctx);
    {
        StageTimer tm(ctx, ST_H2D);
        TCI_CUDA(ctx, dA.upload(A, (size_t)(M * K)));
        TCI_CUDA(ctx, dB.upload(B, (size_t)(K * N)));
        if (beta != 0.0)
            TCI_CUDA(ctx, dC.upload(C, (size_t)(M * N)));
        else
            TCI_CUDA(ctx, dC.alloc((size_t)(M * N)));
    }
    {
        StageTimer tm(ctx, ST_GEMM);
        int rc = dgemm_dev(ctx, transA != 0, transB != 0, M, N, K, alpha, dA.p, lda ? lda : 1, dB.p, ldb ? ldb : 1,
                           beta, dC.p, M);
        if (rc) return rc;
    }
    StageTimer tm(ctx, ST_D2H);
    TCI_CUDA(ctx, cudaMemcpyAsync(C, dC.p, M * N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TCI_OK;
}

// ---- FP64 roofline denominators: register-resident DFMA and DMMA loops -----------------------
// 8 independent accumulator chains per thread; the result is stored so that the loops cannot be removed.
__global__ void __launch_bounds__(256) k_peak_dfma(double *out, int iters, double x, double y)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, x, y);
        a1 = fma(a1, x, y);
        a2 = fma(a2, x, y);
        a3 = fma(a3, x, y);
        a4 = fma(a4, x, y);
        a5 = fma(a5, x, y);
        a6 = fma(a6, x, y);
        a7 = fma(a7, x, y);
    }
    out[blockIdx.x * (i64)blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
__global__ void __launch_bounds__(256) k_peak_dmma(double *out, int iters, double x, double y)
{
    double d[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) d[q] = threadIdx.x + q;
    const double a = x + threadIdx.x * 1e-9, b = y;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) dmma884(d[2 * q], d[2 * q + 1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += d[q];
    out[blockIdx.x * (i64)blockDim.x + threadIdx.x] = s;
}

// out[0] = DFMA pipe TFLOP/s (2 flop per FMA), out[1] = DMMA (mma.sync.m8n8k4.f64: 2*8*8*4 flop per warp instruction)
extern "C" int tci_fp64_peak(tci_ctx *ctx, double *out)
{
    TCI_ENTER(ctx);
    if (!out) return tci_fail(ctx, TCI_ERR_ARG, "tci_fp64_peak: out missing");
    const int blocks = ctx->sm_count * 8, iters = 20000;
    DevBuf<double> buf(ctx);
    TCI_CUDA(ctx, buf.alloc((size_t)blocks * 256));
    for (int which = 0; which < 2; ++which) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(ctx->ev4, ctx->stream);
            if (which == 0)
                k_peak_dfma<<<blocks, 256, 0, ctx->stream>>>(buf.p, iters, 0.999999, 1e-7);
            else
                k_peak_dmma<<<blocks, 256, 0, ctx->stream>>>(buf.p, iters, 0.999999, 1e-7);
            cudaEventRecord(ctx->ev5, ctx->stream);
            TCI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev4, ctx->ev5);
            if (rep > 0 && ms < best) best = ms;
            ctx->launches++;
        }
        const double flop = which == 0 ? (double)blocks * 256 * iters * 8 * 2.0
                                       : (double)blocks * 8 /*warps*/ * iters * 8 * (2.0 * 8 * 8 * 4);
        out[which] = flop / (best * 1e-3) / 1e12;
    }
    return TCI_OK;
}
