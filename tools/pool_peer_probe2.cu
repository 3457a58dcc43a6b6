// Probe 2: (i) how far does a peer-mapped explicit pool grow in 256 MB steps, (ii) cudaMalloc of 9 GB written by the peer.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__global__ void k_fill(double *p, size_t n, double v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_sum(const double *p, size_t n, double *out)
{
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += p[i];
    atomicAdd(out, s);
}
int main()
{
    for (int a = 0; a < 2; ++a) { cudaSetDevice(a); cudaFree(0); printf("enable peer %d: %s\n", a, cudaGetErrorString(cudaDeviceEnablePeerAccess(1 - a, 0))); }
    cudaSetDevice(0);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = 0;
    cudaMemPool_t pool;
    cudaMemPoolCreate(&pool, &props);
    cudaMemAccessDesc d{};
    d.location.type = cudaMemLocationTypeDevice;
    d.location.id = 1;
    d.flags = cudaMemAccessFlagsProtReadWrite;
    cudaMemPoolSetAccess(pool, &d, 1);
    std::vector<void *> ps;
    for (int step : {256, 64, 1024}) {
        size_t total = 0;
        for (int i = 0; i < 64; ++i) {
            void *p;
            cudaError_t e = cudaMallocFromPoolAsync(&p, (size_t)step << 20, pool, s);
            if (e != cudaSuccess) { cudaGetLastError(); printf("pool: step %d MB failed at total %zu MB: %s\n", step, total, cudaGetErrorString(e)); break; }
            total += step;
            ps.push_back(p);
        }
        printf("pool: step %d MB reached %zu MB\n", step, total);
        for (void *p : ps) cudaFreeAsync(p, s);
        ps.clear();
        cudaStreamSynchronize(s);
        cudaMemPoolTrimTo(pool, 0);
    }
    // one fresh pool, one 2 GB allocation first thing
    {
        cudaMemPool_t p2;
        cudaMemPoolCreate(&p2, &props);
        cudaMemPoolSetAccess(p2, &d, 1);
        for (size_t gb : {2, 4, 9}) {
            void *p;
            cudaError_t e = cudaMallocFromPoolAsync(&p, gb << 30, p2, s);
            printf("fresh pool: %zu GB: %s\n", gb, cudaGetErrorString(e));
            if (e == cudaSuccess) cudaFreeAsync(p, s); else cudaGetLastError();
            cudaStreamSynchronize(s);
        }
    }
    // cudaMalloc 9 GB on device 0, filled from device 1, summed on device 0
    double *big = nullptr, *out = nullptr;
    const size_t n = (size_t)9 << 27;
    printf("cudaMalloc 9 GB: %s\n", cudaGetErrorString(cudaMalloc(&big, n * 8)));
    cudaMalloc(&out, 8);
    cudaMemset(out, 0, 8);
    cudaSetDevice(1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_fill<<<148 * 8, 256>>>(big, n, 1.0);
    cudaEventRecord(e1);
    printf("peer fill: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("peer fill 9 GB: %.2f ms = %.1f GB/s\n", ms, n * 8 / ms * 1e-6);
    cudaSetDevice(0);
    k_sum<<<148 * 8, 256>>>(big, n, out);
    double h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("sum = %.0f (expect %zu): %s\n", h, n, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
