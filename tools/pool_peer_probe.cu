// Probe: does growing peer-mapped DEFAULT stream-ordered pools fail?  (a) one thread, (b) two threads at once,
// (c) explicit peer-mapped pools.  nvcc -arch=sm_100a -o pool_peer_probe pool_peer_probe.cu -lpthread
#include <cuda_runtime.h>
#include <cstdio>
#include <thread>
#include <vector>
static void seq(int dev, cudaMemPool_t pool, const char *tag)
{
    cudaSetDevice(dev);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    size_t sizes[] = {8, 1 << 20, 33 << 20, 472 << 20, 118 << 20, 472u << 20, 1888u << 20, 472u << 20};
    std::vector<void *> ps;
    size_t total = 0;
    for (size_t b : sizes) {
        void *p = nullptr;
        cudaError_t e = pool ? cudaMallocFromPoolAsync(&p, b, pool, s) : cudaMallocAsync(&p, b, s);
        total += b;
        printf("[%s dev %d] alloc %zu MB (total %zu MB): %s\n", tag, dev, b >> 20, total >> 20, cudaGetErrorString(e));
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaMemPool_t pl = pool;
            if (!pl) cudaDeviceGetDefaultMemPool(&pl, dev);
            cudaStreamSynchronize(s);
            cudaMemPoolTrimTo(pl, 0);
            e = pool ? cudaMallocFromPoolAsync(&p, b, pool, s) : cudaMallocAsync(&p, b, s);
            printf("[%s dev %d]   after sync + trim: %s\n", tag, dev, cudaGetErrorString(e));
            if (e != cudaSuccess) { cudaGetLastError(); break; }
        }
        ps.push_back(p);
        if (ps.size() % 3 == 0) { cudaFreeAsync(ps[ps.size() - 2], s); ps[ps.size() - 2] = nullptr; }
    }
    for (void *p : ps) if (p) cudaFreeAsync(p, s);
    cudaStreamSynchronize(s);
    fflush(stdout);
}
int main(int argc, char **argv)
{
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    for (int a = 0; a < 2; ++a) { cudaSetDevice(a); cudaFree(0); cudaDeviceEnablePeerAccess(1 - a, 0); }
    cudaMemPool_t pools[2] = {nullptr, nullptr};
    for (int b = 0; b < 2; ++b) {
        cudaMemPool_t pool;
        if (mode == 2) {
            cudaMemPoolProps props{};
            props.allocType = cudaMemAllocationTypePinned;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = b;
            printf("create: %s\n", cudaGetErrorString(cudaMemPoolCreate(&pool, &props)));
            pools[b] = pool;
        } else
            cudaDeviceGetDefaultMemPool(&pool, b);
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (mode != 3) {
            cudaMemAccessDesc d{};
            d.location.type = cudaMemLocationTypeDevice;
            d.location.id = 1 - b;
            d.flags = cudaMemAccessFlagsProtReadWrite;
            printf("setaccess pool %d: %s\n", b, cudaGetErrorString(cudaMemPoolSetAccess(pool, &d, 1)));
        }
    }
    printf("--- mode %d: one thread\n", mode);
    seq(0, pools[0], "seq");
    seq(1, pools[1], "seq");
    printf("--- mode %d: two threads\n", mode);
    std::thread t0(seq, 0, pools[0], "par"), t1(seq, 1, pools[1], "par");
    t0.join();
    t1.join();
    return 0;
}
