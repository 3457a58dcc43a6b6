// user_target.cu -- user-defined device targets (SURVEY 7 hard part 5): a Julia closure cannot run on the GPU, so an
// arbitrary target f is handed over as CUDA source and compiled at run time with NVRTC (libnvrtc.so.12, bound with
// dlopen like NCCL) for sm_100a.  The source defines
//
//     __device__ double tci_user_f(const long long *x, int n, const double *params);
//
// (x[0..n): the 1-based local indices).  The generated module adds the two kernels every target needs -- the batched
// Pi / T evaluation with the fused max|.| (batcheval.jl:32-61, util.jl:1-10) and the point evaluation
// (batcheval.jl:11-13) -- and is loaded per device (cudaLibraryLoadData).  Compiled with --fmad=false so that a target
// written as the reference's closure (rounded multiply, rounded add) gives the reference's bits.
#include <dlfcn.h>

#include <cstring>

#include "tci_internal.h"

typedef struct _nvrtcProgram *nvrtcProgram_t;
struct NvrtcApi {
    int (*CreateProgram)(nvrtcProgram_t *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram_t, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram_t, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram_t, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram_t, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram_t, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram_t *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

static NvrtcApi &nvrtc()
{
    static NvrtcApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = nullptr;
        for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            api.why = std::string("libnvrtc.so.12 not found: ") + (dlerror() ? dlerror() : "");
            return;
        }
#define NVRTC_SYM(field, name)                                  \
    *(void **)(&api.field) = dlsym(h, name);                    \
    if (!api.field) {                                           \
        api.why = std::string("NVRTC symbol missing: ") + name; \
        return;                                                 \
    }
        NVRTC_SYM(CreateProgram, "nvrtcCreateProgram");
        NVRTC_SYM(CompileProgram, "nvrtcCompileProgram");
        NVRTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
        NVRTC_SYM(GetCUBIN, "nvrtcGetCUBIN");
        NVRTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        NVRTC_SYM(GetProgramLog, "nvrtcGetProgramLog");
        NVRTC_SYM(DestroyProgram, "nvrtcDestroyProgram");
        NVRTC_SYM(GetErrorString, "nvrtcGetErrorString");
#undef NVRTC_SYM
        api.ok = true;
    });
    return api;
}

static const char *USER_PRELUDE = R"SRC(
typedef long long i64;
#define TCI_USER_MAXSITES 128
__device__ double tci_user_f(const long long *x, int n, const double *params);
)SRC";

static const char *USER_KERNELS = R"SRC(
// out[(i + nI*c) + ldo*j] = f(I_i ++ c ++ J_j): thread = 2 consecutive rows x a strip of columns; max |.| (NaN sorts
// above Inf in the integer order of the absolute bit patterns) is reduced per block and committed with one atomicMax
extern "C" __global__ void __launch_bounds__(256)
tci_user_pi(const i64 *__restrict__ I, int nl, i64 nI, const i64 *__restrict__ J, int nr, i64 nJ, int M,
            const i64 *__restrict__ ld, i64 C, const double *__restrict__ params, double *__restrict__ out, i64 ldo,
            unsigned long long *gmax)
{
    const i64 rows = nI * C, total = rows * nJ;
    const int n = nl + M + nr;
    unsigned long long mx = 0ull;
    long long x[TCI_USER_MAXSITES];
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 row = e % rows, j = e / rows;
        const i64 i = row % nI;
        i64 c = row / nI;
        for (int k = 0; k < nl; ++k) x[k] = I[(i64)nl * i + k];
        for (int k = 0; k < M; ++k) { // first centre index fastest (Iterators.product, batcheval.jl:49-60)
            const i64 d = ld[nl + k];
            x[nl + k] = c % d + 1;
            c /= d;
        }
        for (int k = 0; k < nr; ++k) x[nl + M + k] = J[(i64)nr * j + k];
        const double v = tci_user_f(x, n, params);
        out[row + ldo * j] = v;
        const unsigned long long b = (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
        mx = b > mx ? b : mx;
    }
    __shared__ unsigned long long sm[256];
    sm[threadIdx.x] = mx;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w && sm[threadIdx.x + w] > sm[threadIdx.x]) sm[threadIdx.x] = sm[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0 && gmax && sm[0]) atomicMax(gmax, sm[0]);
}

extern "C" __global__ void tci_user_points(const i64 *__restrict__ idx, int n, i64 count, const double *__restrict__ params,
                                           double *__restrict__ out)
{
    const i64 q = blockIdx.x * (i64)blockDim.x + threadIdx.x;
    if (q >= count) return;
    out[q] = tci_user_f(idx + (i64)n * q, n, params);
}
)SRC";

int user_target_load(tci_ctx *ctx, TargetDev &t) // loads t.cubin on the context's device and resolves the kernels
{
    cudaLibrary_t lib = nullptr;
    cudaError_t e = cudaLibraryLoadData(&lib, t.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) return tci_fail(ctx, TCI_ERR_CUDA, std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e));
    t.user_lib = lib;
    cudaKernel_t k1 = nullptr, k2 = nullptr;
    e = cudaLibraryGetKernel(&k1, lib, "tci_user_pi");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k2, lib, "tci_user_points");
    if (e != cudaSuccess) return tci_fail(ctx, TCI_ERR_CUDA, std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e));
    t.user_pi = k1;
    t.user_points = k2;
    return TCI_OK;
}

void user_target_unload(TargetDev &t)
{
    if (t.user_lib) cudaLibraryUnload(static_cast<cudaLibrary_t>(t.user_lib));
    t.user_lib = nullptr;
}

extern "C" int tci_target_source(tci_ctx *ctx, const char *source, const double *params, int64_t nparams,
                                 const int64_t *localdims, int64_t nsites, int64_t *target_id)
{
    TCI_ENTER(ctx);
    if (!target_id || !source || nsites < 1 || nsites > 128 || !localdims || (nparams > 0 && !params))
        return tci_fail(ctx, TCI_ERR_ARG, "tci_target_source: bad arguments (at most 128 sites)");
    if (!nvrtc().ok) return tci_fail(ctx, TCI_ERR_UNSUPPORTED, "user-defined targets need NVRTC: " + nvrtc().why);
    const std::string src = std::string(USER_PRELUDE) + source + "\n" + USER_KERNELS;
    nvrtcProgram_t prog = nullptr;
    int r = nvrtc().CreateProgram(&prog, src.c_str(), "tci_user_target.cu", 0, nullptr, nullptr);
    if (r) return tci_fail(ctx, TCI_ERR_CUDA, std::string("nvrtcCreateProgram: ") + nvrtc().GetErrorString(r));
    const char *opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-lineinfo"};
    r = nvrtc().CompileProgram(prog, 4, opts);
    if (r) {
        size_t n = 0;
        nvrtc().GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) nvrtc().GetProgramLog(prog, &log[0]);
        nvrtc().DestroyProgram(&prog);
        return tci_fail(ctx, TCI_ERR_ARG, "the target source does not compile:\n" + log);
    }
    std::unique_ptr<TargetDev> t(new TargetDev());
    size_t nb = 0;
    nvrtc().GetCUBINSize(prog, &nb);
    t->cubin.resize(nb);
    r = nvrtc().GetCUBIN(prog, t->cubin.data());
    nvrtc().DestroyProgram(&prog);
    if (r || nb == 0) return tci_fail(ctx, TCI_ERR_CUDA, "nvrtcGetCUBIN failed");
    t->kind = 3;
    t->nsites = nsites;
    t->localdims.assign(localdims, localdims + nsites);
    t->nparams_alloc = std::max<i64>(nparams, 1);
    std::vector<double> p(params, params + nparams);
    if (p.empty()) p.push_back(0.0);
    TCI_CUDA(ctx, cudaMalloc(&t->d_params, p.size() * sizeof(double)));
    TCI_CUDA(ctx, cudaMemcpy(t->d_params, p.data(), p.size() * sizeof(double), cudaMemcpyHostToDevice));
    TCI_CUDA(ctx, cudaMalloc(&t->d_localdims, nsites * sizeof(i64)));
    TCI_CUDA(ctx, cudaMemcpy(t->d_localdims, localdims, nsites * sizeof(i64), cudaMemcpyHostToDevice));
    int rc = user_target_load(ctx, *t);
    if (rc) {
        target_free(ctx, *t);
        return rc;
    }
    const i64 id = ctx->next_target++;
    ctx->targets[id] = std::move(t);
    *target_id = id;
    return target_replicate(ctx, id);
}

int pi_eval_user(tci_ctx *ctx, TargetDev &t, const i64 *dI, i64 nl, i64 nI, const i64 *dJ, i64 nr, i64 nJ, i64 M,
                 tci_dmat *out, unsigned long long *d_maxbits)
{
    i64 C = 1;
    for (i64 k = 0; k < M; ++k) C *= t.localdims[nl + k];
    const i64 total = nI * C * nJ;
    int inl = (int)nl, inr = (int)nr, iM = (int)M;
    const i64 *ld = t.d_localdims;
    const double *params = t.d_params;
    double *op = out->p;
    i64 ldo = out->ld;
    void *args[] = {(void *)&dI, &inl, &nI, (void *)&dJ, &inr, &nJ, &iM, (void *)&ld, &C, (void *)&params, &op, &ldo,
                    (void *)&d_maxbits};
    const unsigned blocks = (unsigned)std::min<i64>((total + 255) / 256, (i64)ctx->sm_count * 16);
    TCI_CUDA(ctx, cudaLaunchKernel((const void *)t.user_pi, dim3(blocks), dim3(256), args, 0, ctx->stream));
    ctx->launches++;
    return TCI_OK;
}

int target_eval_user(tci_ctx *ctx, TargetDev &t, const i64 *d_idx, i64 count, double *d_out)
{
    int n = (int)t.nsites;
    const double *params = t.d_params;
    void *args[] = {(void *)&d_idx, &n, &count, (void *)&params, &d_out};
    TCI_CUDA(ctx, cudaLaunchKernel((const void *)t.user_points, dim3((unsigned)((count + 127) / 128)), dim3(128), args, 0,
                                   ctx->stream));
    ctx->launches++;
    return TCI_OK;
}
