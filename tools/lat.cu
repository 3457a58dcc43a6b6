// micro-latencies that bound the rrLU per-pivot chain on B200 (build: nvcc -arch=sm_100a -o lat lat.cu)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *t, double seed, int iters)
{
    __shared__ double sm[1024];
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double a = seed, b = seed * 0.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = __dmul_rn(a, b);
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i) a = __dadd_rn(a, b);
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, b, b);
    long long t3 = clock64();
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) idx = (int)sm[idx & 1023] & 1023;
    long long t4 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    long long t5 = clock64();
    double c = a;
    for (int i = 0; i < iters; ++i) c = (c > b) ? c * 0.999 : b; // DSETP + select chain
    long long t6 = clock64();
    float f = (float)seed;
    for (int i = 0; i < iters; ++i) f = fmaf(f, 0.999f, 0.5f);
    long long t7 = clock64();
    for (int i = 0; i < iters; ++i) a = __ddiv_rn(a, b);
    long long t8 = clock64();
    if (threadIdx.x == 0) {
        t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; t[6] = t7 - t6; t[7] = t8 - t7;
    }
    out[threadIdx.x] = a + idx + c + f;
}
int main()
{
    double *out; long long *t; cudaMalloc(&out, 8192); cudaMalloc(&t, 64);
    for (int threads : {32, 256, 1024}) {
        int iters = 1000;
        k<<<1, threads>>>(out, t, 1.0000001, iters);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, t, 64, cudaMemcpyDeviceToHost);
        printf("threads %4d: cycles/iter dmul %.1f dadd %.1f dfma %.1f lds+cvt %.1f syncthreads %.1f dsetp+sel %.1f ffma %.1f ddiv %.1f\n", threads,
               h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[4] / (double)iters, h[5] / (double)iters, h[6] / (double)iters, h[7] / (double)iters);
    }
    return 0;
}
