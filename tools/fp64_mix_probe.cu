// fp64_mix_probe.cu -- do the FP64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA) and the DFMA pipe run concurrently
// on sm_100a, i.e. can a GEMM that feeds both exceed either peak?  Variants: all warps DMMA, all warps DFMA, both
// instruction kinds interleaved in every warp (ratio R DFMA per DMMA), and half the warps of each kind.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_mix_probe tools/fp64_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// MODE 0: DMMA only, 1: DFMA only, 2: interleaved (8 DMMA + 8*R DFMA per iteration), 3: even warps DMMA / odd warps DFMA
template <int MODE, int R> __global__ void __launch_bounds__(512) k(double *out, int iters, double x, double y)
{
    double d[16], f[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        d[q] = threadIdx.x + q;
        f[q] = threadIdx.x * 0.5 + q;
    }
    const double a = x + threadIdx.x * 1e-9, b = y;
    const int warp = threadIdx.x >> 5;
    const bool do_mma = MODE == 0 || MODE == 2 || (MODE == 3 && (warp & 1) == 0);
    const bool do_fma = MODE == 1 || MODE == 2 || (MODE == 3 && (warp & 1) == 1);
    if (MODE == 3) {
        if (do_mma) {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int q = 0; q < 8; ++q) dmma884(d[2 * q], d[2 * q + 1], a, b);
            }
        } else {
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int rr = 0; rr < R; ++rr)
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] = fma(f[q], a, b);
            }
        }
    } else {
        for (int i = 0; i < iters; ++i) {
            if (do_mma) {
#pragma unroll
                for (int q = 0; q < 8; ++q) dmma884(d[2 * q], d[2 * q + 1], a, b);
            }
            if (do_fma) {
#pragma unroll
                for (int rr = 0; rr < R; ++rr)
#pragma unroll
                    for (int q = 0; q < 16; ++q) f[q] = fma(f[q], a, b);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) s += d[q] + f[q];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}

template <int MODE, int R> void run(int warps, int sms, double *buf, const char *label)
{
    const int iters = 8192;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE, R><<<sms, warps * 32>>>(buf, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    double mma_warps = MODE == 0 || MODE == 2 ? warps : (MODE == 3 ? warps / 2 : 0);
    double fma_warps = MODE == 1 || MODE == 2 ? warps : (MODE == 3 ? warps / 2 : 0);
    const double fl_mma = (double)sms * mma_warps * iters * 8 * 512.0;
    const double fl_fma = (double)sms * fma_warps * 32 * iters * R * 16 * 2.0;
    printf("%-34s warps/SM %2d R %d: %6.3f ms  DMMA %6.2f + DFMA %6.2f = %6.2f TFLOP/s\n", label, warps, R, best,
           fl_mma / (best * 1e-3) / 1e12, fl_fma / (best * 1e-3) / 1e12, (fl_mma + fl_fma) / (best * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *buf;
    cudaMalloc(&buf, (size_t)p.multiProcessorCount * 1024 * 8);
    for (int w : {8, 16}) {
        run<0, 1>(w, p.multiProcessorCount, buf, "DMMA only");
        run<1, 1>(w, p.multiProcessorCount, buf, "DFMA only");
        run<2, 1>(w, p.multiProcessorCount, buf, "interleaved 8 DMMA : 16 DFMA");
        run<2, 2>(w, p.multiProcessorCount, buf, "interleaved 8 DMMA : 32 DFMA");
        run<2, 4>(w, p.multiProcessorCount, buf, "interleaved 8 DMMA : 64 DFMA");
        run<3, 1>(w, p.multiProcessorCount, buf, "even warps DMMA / odd warps DFMA");
        run<3, 4>(w, p.multiProcessorCount, buf, "even DMMA / odd DFMA (4x work)");
    }
    return 0;
}
