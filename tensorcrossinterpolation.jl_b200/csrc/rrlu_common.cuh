// rrlu_common.cuh -- pieces shared by the two rrLU kernels (rrlu.cu: in-place update per pivot;
// rrlu_lazy.cu: deferred updates, committed every NB pivots).
#pragma once
#include "tci_internal.h"

#define RR_MAX_THREADS 1024
#define RR_XS_CAP 24576 // doubles of shared memory for the pivot column
#define RR_U 4          // 16-byte loads in flight per lane in the trailing update

// One candidate record = ONE aligned 16-byte word, written with a single 16-byte store and polled
// with single 16-byte loads: the phase bit (top bit of rowphase) flips every second step, so a
// reader can tell a fresh record from the one left two steps earlier in the same parity buffer
// without a separate flag word and without a second round trip through L2.
struct __align__(16) RRCand {
    double val;        // value of the candidate (abs2 is recomputed by the reader)
    unsigned rowphase; // row | phase << 31
    int colpos;        // column position, -1: this CTA has no finite candidate
};
#define RRLU_LAZY_NB 4 // pivots per commit of the deferred-update kernel (rrlu_lazy.cu)
#define RR_MAXQ 5 // candidate records per lane of the polling warp (G <= 160)

struct RRArgs {
    double *A;
    i64 m, n, ld;
    int maxrank;
    double reltol, abstol;
    int leftorth;
    int *colpos;     // [n]
    i64 *rowperm;    // [m] 0-based
    i64 *colperm;    // [n] position -> physical column
    double *pivvals; // [maxrank]
    int *pivrows;    // [maxrank] row picked at every step
    RRCand *cand;    // [2][G], zero initialised
    double *xbuf;    // [2][G][ldx]
    i64 ldx;
    int *result;        // [0] npivot, [1] flags (1: no finite candidate left)
    double *result_err; // lu.error
    int xs_in_smem;
    int maxown;
    i64 lds; // leading dimension of the shared-memory resident columns (RES mode)
    long long *dbg; // optional per-phase cycle counters (TCI_RRLU_DEBUG)
    int dbg_cta;
    int nxslots; // lazy kernel: xbuf is [nxslots][G][ldx]
    // lazy kernel, speed-weighted column ownership (all null: CTA g owns columns g, g+G, ...)
    const int *colmap;    // [G][maxown] physical column of (logical CTA, own slot)
    const int *colcnt;    // [G] columns owned by the logical CTA
    int *smids;           // [G] zeroed; smid+1 of every CTA, to rank the CTAs by the SM they run on
    unsigned *startbar;   // zeroed
    long long *passstats; // [2][G] cycles spent in the streaming passes, column tiles streamed
};

__device__ __forceinline__ void ld_relaxed_16(const RRCand *p, double &val, unsigned &rowphase, int &colpos)
{
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    val = __longlong_as_double((long long)a);
    rowphase = (unsigned)(b & 0xffffffffull);
    colpos = (int)(b >> 32);
}
__device__ __forceinline__ void st_relaxed_16(RRCand *p, double val, unsigned rowphase, int colpos)
{
    unsigned long long a = (unsigned long long)__double_as_longlong(val);
    unsigned long long b = (unsigned long long)rowphase | ((unsigned long long)(unsigned)colpos << 32);
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

template <bool EXACT> __device__ __forceinline__ double schur(double a, double x, double y)
{
    if (EXACT) return __dsub_rn(a, __dmul_rn(x, y)); // matrixlu.jl:132
    return fma(-x, y, a);
}

// abs2 value -> ordered integer (0 = no candidate); squares are >= 0 so the bit pattern is monotonic
__device__ __forceinline__ unsigned long long vbits(double v)
{
    return v == -INFINITY ? 0ull : (unsigned long long)__double_as_longlong(v) + 1ull;
}
// Warp arg-max with the reference's tie-break: max value bits, then min key.  Four REDUX operations
// instead of a five-step shuffle tree; every lane returns the winner.
__device__ __forceinline__ void warp_argmax(unsigned long long &vb, unsigned long long &key)
{
    const unsigned hi = (unsigned)(vb >> 32), lo = (unsigned)vb;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    const bool top = hi == mhi && lo == mlo;
    const unsigned khi = (unsigned)(key >> 32), klo = (unsigned)key;
    const unsigned nhi = __reduce_min_sync(0xffffffffu, top ? khi : 0xffffffffu);
    const unsigned nlo = __reduce_min_sync(0xffffffffu, (top && khi == nhi) ? klo : 0xffffffffu);
    vb = ((unsigned long long)mhi << 32) | mlo;
    key = ((unsigned long long)nhi << 32) | nlo;
}


// rrlu_lazy.cu
size_t rrlu_lazy_smem(int maxown, int nb);
int rrlu_lazy_launch(tci_ctx *ctx, RRArgs &args, int G, size_t smem, bool exact);
