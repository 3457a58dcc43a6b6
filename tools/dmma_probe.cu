// dmma_probe.cu -- how many warps per SM sub-partition and how many independent accumulators does the FP64 tensor
// path (mma.sync.m8n8k4.f64, SASS DMMA) need to reach its peak, and what does an interleaved LDS.64 cost?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_probe tools/dmma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int ILP, int LDS> __global__ void k(double *out, int iters, double x, double y)
{
    __shared__ double sm[2048];
    for (int t = threadIdx.x; t < 2048; t += blockDim.x) sm[t] = x + t * 1e-9;
    __syncthreads();
    double d[2 * ILP];
#pragma unroll
    for (int q = 0; q < 2 * ILP; ++q) d[q] = threadIdx.x + q;
    double a = x + threadIdx.x * 1e-9, b = y;
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < ILP; ++q) {
            if (LDS && (q % LDS) == 0) a = sm[(lane + 33 * q + i) & 2047]; // one LDS.64 every LDS DMMAs
            dmma884(d[2 * q], d[2 * q + 1], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 2 * ILP; ++q) s += d[q];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}

template <int ILP, int LDS> void run(int warps_per_sm, int sms, double *buf)
{
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<ILP, LDS><<<sms, warps_per_sm * 32>>>(buf, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    const double flop = (double)sms * warps_per_sm * iters * ILP * 512.0;
    printf("warps/SM %2d (per SMSP %d)  ILP %2d  LDS every %d: %7.2f TFLOP/s\n", warps_per_sm, warps_per_sm / 4, ILP, LDS,
           flop / (best * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *buf;
    cudaMalloc(&buf, (size_t)p.multiProcessorCount * 1024 * 8);
    for (int w : {4, 8, 16, 32}) {
        run<1, 0>(w, p.multiProcessorCount, buf);
        run<2, 0>(w, p.multiProcessorCount, buf);
        run<4, 0>(w, p.multiProcessorCount, buf);
        run<8, 0>(w, p.multiProcessorCount, buf);
        run<32, 0>(w, p.multiProcessorCount, buf);
        run<32, 8>(w, p.multiProcessorCount, buf);
        run<32, 3>(w, p.multiProcessorCount, buf);
        run<32, 1>(w, p.multiProcessorCount, buf);
    }
    return 0;
}
