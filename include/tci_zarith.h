/*
 * tci_zarith.h -- ComplexF64 scalar arithmetic as Julia Base defines it, for the ComplexF64 value type of the
 * TensorCI2 hot path.  matrixlu.jl is generic in T; on a Matrix{ComplexF64} its operations are
 *
 *   abs2(z)  (pivot metric, matrixlu.jl:152)      = real(z)*real(z) + imag(z)*imag(z)            base/complex.jl
 *   abs(z)   (lu.error, stop rule, :153-155)      = hypot(real(z), imag(z))                       base/math.jl _hypot
 *   x / piv  (scaling, :120-124)                  = the robust complex division of base/complex.jl
 *                                                   (Baudin & Smith, arXiv:1210.4539, with the over/underflow scaling)
 *   a - x*y  (Schur update, :132)                 = Complex(re(x)re(y) - im(x)im(y), re(x)im(y) + im(x)re(y)), then -
 *
 * none of which Julia contracts into fused multiply-adds (only `muladd`/`fma` calls fuse, and _hypot makes them
 * explicitly).  Julia Base is not part of /root/reference (it ships with the Julia binary, compat julia = "1.9"), so
 * these are restatements of its PUBLISHED source, not of files that can be cited by line here; the hypot branch taken
 * is the one for hardware with a native fma (every x86-64 since Haswell), which is correctly rounded.
 *
 * Like tci_targets.h this header is compiled into BOTH the CUDA library (device code) and the CPU oracle (host code,
 * -ffp-contract=off), so that the two sides apply the same sequence of IEEE operations and agree to the last bit.
 */
#ifndef TCI_ZARITH_H
#define TCI_ZARITH_H

#include <math.h>

#if defined(__CUDACC__)
#define TCI_ZHD __host__ __device__ __forceinline__
#else
#define TCI_ZHD static inline
#endif

#if defined(__CUDA_ARCH__)
#define TCI_ZMUL(a, b) __dmul_rn((a), (b))
#define TCI_ZADD(a, b) __dadd_rn((a), (b))
#define TCI_ZSUB(a, b) __dsub_rn((a), (b))
#define TCI_ZDIV(a, b) __ddiv_rn((a), (b))
#define TCI_ZFMA(a, b, c) __fma_rn((a), (b), (c))
#define TCI_ZSQRT(a) __dsqrt_rn((a))
#else
#define TCI_ZMUL(a, b) ((a) * (b))
#define TCI_ZADD(a, b) ((a) + (b))
#define TCI_ZSUB(a, b) ((a) - (b))
#define TCI_ZDIV(a, b) ((a) / (b))
#define TCI_ZFMA(a, b, c) fma((a), (b), (c)) /* libm: correctly rounded with or without hardware FMA */
#define TCI_ZSQRT(a) sqrt((a))
#endif

typedef struct {
    double re, im;
} tci_z; /* memory layout of Julia's ComplexF64 */

TCI_ZHD tci_z tci_zmake(double re, double im)
{
    tci_z r;
    r.re = re;
    r.im = im;
    return r;
}

/* abs2(z::Complex) = real(z)*real(z) + imag(z)*imag(z) */
TCI_ZHD double tci_zabs2(tci_z z) { return TCI_ZADD(TCI_ZMUL(z.re, z.re), TCI_ZMUL(z.im, z.im)); }

/* *(z::Complex, w::Complex) */
TCI_ZHD tci_z tci_zmul(tci_z z, tci_z w)
{
    return tci_zmake(TCI_ZSUB(TCI_ZMUL(z.re, w.re), TCI_ZMUL(z.im, w.im)),
                     TCI_ZADD(TCI_ZMUL(z.re, w.im), TCI_ZMUL(z.im, w.re)));
}
TCI_ZHD tci_z tci_zsub(tci_z z, tci_z w) { return tci_zmake(TCI_ZSUB(z.re, w.re), TCI_ZSUB(z.im, w.im)); }
TCI_ZHD tci_z tci_zadd(tci_z z, tci_z w) { return tci_zmake(TCI_ZADD(z.re, w.re), TCI_ZADD(z.im, w.im)); }

/* hypot(x::Float64, y::Float64) of base/math.jl (_hypot, the branch for a native fma: correctly rounded) */
TCI_ZHD double tci_hypot(double x, double y)
{
    double ax = fabs(x), ay = fabs(y);
    if (isinf(ax) || isinf(ay)) return INFINITY;
    if (ay > ax) {
        double t = ax;
        ax = ay;
        ay = t;
    }
    /* widely varying operands (also ay == 0); NaN falls through and propagates */
    if (ay <= TCI_ZMUL(ax, 0x1.6a09e667f3bcdp-27 /* sqrt(eps/2) */)) return ax;
    double scale = 0x1.0p-563; /* eps(Float64) * sqrt(floatmin(Float64)) */
    if (ax > 0x1.6a09e667f3bccp+511 /* sqrt(floatmax/2) */) {
        ax = TCI_ZMUL(ax, scale);
        ay = TCI_ZMUL(ay, scale);
        scale = TCI_ZDIV(1.0, scale);
    } else if (ay < 0x1.0p-511 /* sqrt(floatmin) */) {
        ax = TCI_ZDIV(ax, scale);
        ay = TCI_ZDIV(ay, scale);
    } else {
        scale = 1.0;
    }
    double h = TCI_ZSQRT(TCI_ZFMA(ax, ax, TCI_ZMUL(ay, ay)));
    const double hsquared = TCI_ZMUL(h, h), axsquared = TCI_ZMUL(ax, ax);
    const double corr = TCI_ZSUB(TCI_ZADD(TCI_ZFMA(-ay, ay, TCI_ZSUB(hsquared, axsquared)), TCI_ZFMA(h, h, -hsquared)),
                                 TCI_ZFMA(ax, ax, -axsquared));
    h = TCI_ZSUB(h, TCI_ZDIV(corr, TCI_ZMUL(2.0, h)));
    return TCI_ZMUL(h, scale);
}

/* abs(z::Complex) = hypot(real(z), imag(z)) */
TCI_ZHD double tci_zabs(tci_z z) { return tci_hypot(z.re, z.im); }

/* robust_cdiv2 / robust_cdiv1 / cdiv of base/complex.jl */
TCI_ZHD double tci_cdiv2(double a, double b, double c, double d, double r, double t)
{
    if (r != 0.0) {
        const double br = TCI_ZMUL(b, r);
        return br != 0.0 ? TCI_ZMUL(TCI_ZADD(a, br), t) : TCI_ZADD(TCI_ZMUL(a, t), TCI_ZMUL(TCI_ZMUL(b, t), r));
    }
    return TCI_ZMUL(TCI_ZADD(a, TCI_ZMUL(d, TCI_ZDIV(b, c))), t);
}
TCI_ZHD void tci_cdiv1(double a, double b, double c, double d, double *p, double *q)
{
    const double r = TCI_ZDIV(d, c);
    const double t = TCI_ZDIV(1.0, TCI_ZADD(c, TCI_ZMUL(d, r)));
    *p = tci_cdiv2(a, b, c, d, r, t);
    *q = tci_cdiv2(b, -a, c, d, r, t);
}
TCI_ZHD void tci_cdiv(double a, double b, double c, double d, double *p, double *q)
{
    if (fabs(d) <= fabs(c)) {
        tci_cdiv1(a, b, c, d, p, q);
    } else {
        tci_cdiv1(b, a, d, c, p, q);
        *q = -*q;
    }
}

/* /(z::ComplexF64, w::ComplexF64) */
TCI_ZHD tci_z tci_zdiv(tci_z z, tci_z w)
{
    double a = z.re, b = z.im, c = w.re, d = w.im;
    const double absa = fabs(a), absb = fabs(b), absc = fabs(c), absd = fabs(d);
    const double ab = absa >= absb ? absa : absb, cd = absc >= absd ? absc : absd;
    const double halfov = 0x1.fffffffffffffp+1022; /* 0.5 * floatmax(Float64) */
    const double twounep = 0x1.0p-969;            /* floatmin(Float64) * 2 / eps(Float64) */
    double p, q;
    if (ab >= halfov || ab <= twounep || cd >= halfov || cd <= twounep) { /* scaling_cdiv */
        const double bs = 0x1.0p+105; /* 2 / (eps * eps) */
        double s = 1.0;
        if (ab >= halfov) {
            a = TCI_ZMUL(a, 0.5);
            b = TCI_ZMUL(b, 0.5);
            s = TCI_ZMUL(s, 2.0);
        } else if (ab <= twounep) {
            a = TCI_ZMUL(a, bs);
            b = TCI_ZMUL(b, bs);
            s = TCI_ZDIV(s, bs);
        }
        if (cd >= halfov) {
            c = TCI_ZMUL(c, 0.5);
            d = TCI_ZMUL(d, 0.5);
            s = TCI_ZMUL(s, 0.5);
        } else if (cd <= twounep) {
            c = TCI_ZMUL(c, bs);
            d = TCI_ZMUL(d, bs);
            s = TCI_ZMUL(s, bs);
        }
        tci_cdiv(a, b, c, d, &p, &q);
        p = TCI_ZMUL(p, s);
        q = TCI_ZMUL(q, s);
    } else {
        tci_cdiv(a, b, c, d, &p, &q);
    }
    return tci_zmake(p, q);
}

#endif /* TCI_ZARITH_H */
