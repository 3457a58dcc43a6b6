/*
 * tci_targets.h -- definitions of the built-in *target functions* f(v) that the
 * TensorCI2 hot path interpolates.  A target is INPUT DATA of the path (the
 * reference takes an arbitrary Julia closure `f`, src/tensorci2.jl:943-953); a
 * device kernel cannot call a Julia closure, so targets are registered by ID
 * (tci_target_builtin in tci_b200.h).  This header is the single definition of
 * every built-in target, compiled both into the CUDA library (device code) and
 * into the CPU oracle (host code), so that "the same f" is evaluated on both
 * sides down to the last bit:
 *
 *   - every floating point operation goes through TCI_MUL/ADD/SUB/DIV, which map
 *     to the round-to-nearest, never-contracted intrinsics on the device and to
 *     plain operators on the host (the host is compiled with -ffp-contract=off);
 *   - transcendental functions are implemented here (Cody-Waite reduction +
 *     the classic fdlibm minimax kernels) instead of calling libm / libdevice,
 *     whose results differ from each other by 1-2 ulp.
 *
 * Every target has the additive-state form
 *        s[j] = sum_k phi_j(k, sigma_k)  (j < nstate),   f(v) = F(s)
 * The *scalar* definition accumulates k = 0..n-1 in order.  Targets marked
 * "exact" have phi values whose partial sums are exactly representable in
 * binary64 (small integers / dyadic rationals), hence any association of the
 * sum gives the same bits; the batched device kernel exploits this to combine
 * a per-row state with a per-column state.
 *
 * Indices sigma_k are 1-based, as in the reference (MultiIndex = Vector{Int},
 * src/abstracttensortrain.jl:6-7).
 */
#ifndef TCI_TARGETS_H
#define TCI_TARGETS_H

#include <stdint.h>

#if defined(__CUDACC__)
#define TCI_HD __host__ __device__ __forceinline__
#else
#define TCI_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define TCI_MUL(a, b) __dmul_rn((a), (b))
#define TCI_ADD(a, b) __dadd_rn((a), (b))
#define TCI_SUB(a, b) __dsub_rn((a), (b))
#define TCI_DIV(a, b) __ddiv_rn((a), (b))
#define TCI_RCP(a) __drcp_rn((a))
#else
#define TCI_MUL(a, b) ((a) * (b))
#define TCI_ADD(a, b) ((a) + (b))
#define TCI_SUB(a, b) ((a) - (b))
#define TCI_DIV(a, b) ((a) / (b))
#define TCI_RCP(a) (1.0 / (a))
#endif

#define TCI_MAX_STATE 6

enum tci_target_kind {
    TCI_TARGET_LORENTZ = 1,    /* params: [coeff]           f = coeff/(1+sum v_k^2)      exact */
    TCI_TARGET_SUM = 2,        /* params: []                f = sum v_k                  exact */
    TCI_TARGET_QUANTICS2D = 3, /* params: [layout, R]       BASELINE config 3            exact */
    TCI_TARGET_SEPCOS = 4,     /* params: [nterms, w[n], a[nterms], omega[nterms*n]]  config 4 */
    TCI_TARGET_TABLE = 5,      /* params: [stride[n], table[prod d]]  dense lookup       exact */
    TCI_TARGET_QUANTICS1D = 6, /* params: [R, fid]          exp(-x) (+1e-3 sin(1000x))   exact */
    TCI_TARGET_GKCOSEXP = 7    /* params: [q, node[q], weight[q]]  Gauss-Kronrod weighted
                                  1000 cos(10 sum x^2) exp(-(sum x)^4/1000) (test_integration.jl:61-70)
                                  NOT exact: evaluated site by site in order on every path */
};

typedef struct {
    int32_t kind;
    int32_t nsites;
    int32_t nstate;
    int32_t reserved;
    int64_t nparams;
    const double *params;     /* host pointer in the oracle, device pointer in kernels */
    const int64_t *localdims; /* same */
} tci_analytic_t;

/* ---------------------------------------------------------------- math ---- */

TCI_HD double tci_rint(double x)
{
    /* round half to even via the 2^52 trick; valid for |x| < 2^51 */
    const double big = 6755399441055744.0; /* 1.5 * 2^52 */
    double t = TCI_ADD(x, big);
    return TCI_SUB(t, big);
}

TCI_HD double tci_ksin(double x)
{ /* |x| <= pi/4 ; fdlibm __kernel_sin coefficients */
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = TCI_MUL(x, x);
    double v = TCI_MUL(z, x);
    double r = TCI_ADD(S2, TCI_MUL(z, TCI_ADD(S3, TCI_MUL(z, TCI_ADD(S4, TCI_MUL(z, TCI_ADD(S5, TCI_MUL(z, S6))))))));
    return TCI_ADD(x, TCI_MUL(v, TCI_ADD(S1, TCI_MUL(z, r))));
}

TCI_HD double tci_kcos(double x)
{ /* |x| <= pi/4 ; fdlibm __kernel_cos coefficients */
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = TCI_MUL(x, x);
    double r = TCI_MUL(z, TCI_ADD(C1, TCI_MUL(z, TCI_ADD(C2, TCI_MUL(z, TCI_ADD(C3, TCI_MUL(z, TCI_ADD(C4, TCI_MUL(z, TCI_ADD(C5, TCI_MUL(z, C6)))))))))));
    double hz = TCI_MUL(0.5, z);
    double w = TCI_SUB(1.0, hz);
    /* 1 - z/2 + z*r with the rounding error of (1 - hz) folded back in */
    return TCI_ADD(w, TCI_ADD(TCI_SUB(TCI_SUB(1.0, w), hz), TCI_MUL(z, r)));
}

TCI_HD double tci_sincos_quadrant(double r, int q, int want_cos)
{
    /* value of sin (want_cos=0) or cos (want_cos=1) of (r + q*pi/2), |r|<=pi/4 */
    int n = (q + want_cos) & 3;
    double v = (n & 1) ? tci_kcos(r) : tci_ksin(r);
    return (n & 2) ? -v : v;
}

/* sin / cos of 2*pi*t with an exact argument reduction (t - rint(t) is exact). */
TCI_HD double tci_trig2pi(double t, int want_cos)
{
    const double twopi = 6.283185307179586476925287;
    double r = TCI_SUB(t, tci_rint(t));          /* [-1/2, 1/2], exact */
    double q = tci_rint(TCI_MUL(4.0, r));        /* -2..2 */
    double rr = TCI_SUB(r, TCI_MUL(0.25, q));    /* [-1/8, 1/8], exact */
    return tci_sincos_quadrant(TCI_MUL(twopi, rr), (int)q, want_cos);
}
TCI_HD double tci_cos2pi(double t) { return tci_trig2pi(t, 1); }
TCI_HD double tci_sin2pi(double t) { return tci_trig2pi(t, 0); }

/* sin / cos of x in radians, |x| < ~1e6 (three-term Cody-Waite reduction). */
TCI_HD double tci_trig(double x, int want_cos)
{
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
    const double pio2_2 = 6.07710050630396597660e-11;  /* second 33 bits */
    const double pio2_3 = 2.02226624871116645580e-21;  /* third 33 bits */
    const double pio2_3t = 8.47842766036889956997e-32; /* tail */
    double fn = tci_rint(TCI_MUL(x, invpio2));
    double r = TCI_SUB(x, TCI_MUL(fn, pio2_1));
    r = TCI_SUB(r, TCI_MUL(fn, pio2_2));
    r = TCI_SUB(r, TCI_MUL(fn, pio2_3));
    r = TCI_SUB(r, TCI_MUL(fn, pio2_3t));
    /* (int)fn via int64 keeps the two low bits also for negative fn */
    return tci_sincos_quadrant(r, (int)((int64_t)fn & 3), want_cos);
}
TCI_HD double tci_cos(double x) { return tci_trig(x, 1); }
TCI_HD double tci_sin(double x) { return tci_trig(x, 0); }

/* exp(x) for |x| < 700 (fdlibm e_exp.c scheme). */
TCI_HD double tci_exp(double x)
{
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
                 invln2 = 1.44269504088896338700e+00;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                 P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                 P5 = 4.13813679705723846039e-08;
    double fk = tci_rint(TCI_MUL(x, invln2));
    double hi = TCI_SUB(x, TCI_MUL(fk, ln2HI));
    double lo = TCI_MUL(fk, ln2LO);
    double r = TCI_SUB(hi, lo);
    double t = TCI_MUL(r, r);
    double c = TCI_SUB(r, TCI_MUL(t, TCI_ADD(P1, TCI_MUL(t, TCI_ADD(P2, TCI_MUL(t, TCI_ADD(P3, TCI_MUL(t, TCI_ADD(P4, TCI_MUL(t, P5))))))))));
    double y = TCI_SUB(1.0, TCI_SUB(TCI_SUB(lo, TCI_DIV(TCI_MUL(r, c), TCI_SUB(2.0, c))), hi));
    /* scale by 2^k in two exact steps (k may reach +-1022) */
    int64_t k = (int64_t)fk;
    int64_t k1 = k / 2, k2 = k - k1;
    union { uint64_t u; double d; } a, b;
    a.u = (uint64_t)(1023 + k1) << 52;
    b.u = (uint64_t)(1023 + k2) << 52;
    return TCI_MUL(TCI_MUL(y, a.d), b.d);
}

/* ------------------------------------------------------------- targets ---- */

TCI_HD int tci_target_nstate(int kind, const double *params)
{
    switch (kind) {
    case TCI_TARGET_LORENTZ: return 1;
    case TCI_TARGET_SUM: return 1;
    case TCI_TARGET_QUANTICS2D: return 2;
    case TCI_TARGET_SEPCOS: return 1 + (int)params[0];
    case TCI_TARGET_TABLE: return 1;
    case TCI_TARGET_QUANTICS1D: return 1;
    case TCI_TARGET_GKCOSEXP: return 3;
    default: return 0;
    }
}

/* 1 if partial sums of the state are exact in binary64 for admissible parameters, i.e. the
 * state of a multi-index may be formed as state(left part) + state(right part). */
TCI_HD int tci_target_exact(int kind) { return kind != TCI_TARGET_GKCOSEXP; }

TCI_HD void tci_target_init(const tci_analytic_t *t, double *s)
{
    for (int j = 0; j < TCI_MAX_STATE; ++j) s[j] = 0.0;
    if (t->kind == TCI_TARGET_GKCOSEXP) s[2] = 1.0;
}

TCI_HD double tci_pow2neg(int b)
{ /* 2^-b, 0 <= b <= 1000 */
    union { uint64_t u; double d; } a;
    a.u = (uint64_t)(1023 - b) << 52;
    return a.d;
}

/* s[j] += phi_j(site, sigma); site is 0-based, sigma 1-based. */
TCI_HD void tci_target_accum(const tci_analytic_t *t, int site, int64_t sigma, double *s)
{
    const double *p = t->params;
    switch (t->kind) {
    case TCI_TARGET_LORENTZ: {
        double v = (double)sigma;
        s[0] = TCI_ADD(s[0], TCI_MUL(v, v));
        break;
    }
    case TCI_TARGET_SUM:
        s[0] = TCI_ADD(s[0], (double)sigma);
        break;
    case TCI_TARGET_QUANTICS2D: {
        int layout = (int)p[0];
        if (layout == 0) { /* fused: sigma-1 = bx + 2*by, site b carries bit b+1 */
            int64_t q = sigma - 1;
            double w = tci_pow2neg(site + 1);
            s[0] = TCI_ADD(s[0], TCI_MUL((double)(q & 1), w));
            s[1] = TCI_ADD(s[1], TCI_MUL((double)((q >> 1) & 1), w));
        } else { /* interleaved: x1 y1 x2 y2 ... */
            double w = tci_pow2neg(site / 2 + 1);
            int j = site & 1;
            s[j] = TCI_ADD(s[j], TCI_MUL((double)(sigma - 1), w));
        }
        break;
    }
    case TCI_TARGET_SEPCOS: {
        int nt = (int)p[0];
        int n = t->nsites;
        const double *w = p + 1;
        const double *om = p + 1 + n + nt;
        double d = (double)t->localdims[site];
        double x = TCI_DIV((double)(2 * sigma - 1), TCI_MUL(2.0, d));
        s[0] = TCI_ADD(s[0], TCI_MUL(w[site], TCI_MUL(x, x)));
        for (int m = 0; m < nt; ++m)
            s[1 + m] = TCI_ADD(s[1 + m], TCI_MUL(om[m * n + site], x));
        break;
    }
    case TCI_TARGET_TABLE:
        s[0] = TCI_ADD(s[0], TCI_MUL((double)(sigma - 1), p[site]));
        break;
    case TCI_TARGET_QUANTICS1D:
        s[0] = TCI_ADD(s[0], TCI_MUL((double)(sigma - 1), tci_pow2neg(site + 1)));
        break;
    case TCI_TARGET_GKCOSEXP: {
        int q = (int)p[0];
        double x = p[1 + (sigma - 1)];
        double w = TCI_MUL(p[1 + q + (sigma - 1)], (double)q);
        s[0] = TCI_ADD(s[0], TCI_MUL(x, x));
        s[1] = TCI_ADD(s[1], x);
        s[2] = TCI_MUL(s[2], w);
        break;
    }
    default:
        break;
    }
}

TCI_HD double tci_target_finalize(const tci_analytic_t *t, const double *s)
{
    const double *p = t->params;
    switch (t->kind) {
    case TCI_TARGET_LORENTZ: {
        double coeff = (t->nparams > 0) ? p[0] : 1.0;
        double den = TCI_ADD(s[0], 1.0);
        /* 1/den correctly rounded is the same number whichever way it is computed */
        return coeff == 1.0 ? TCI_RCP(den) : TCI_DIV(coeff, den);
    }
    case TCI_TARGET_SUM:
        return s[0];
    case TCI_TARGET_QUANTICS2D: {
        /* cos(2pi(37x+23y)) exp(-((x-1/2)^2+(y-1/2)^2)/0.1) + 1/2 sin(2pi 101 x y) */
        double x = s[0], y = s[1];
        double ph = TCI_ADD(TCI_MUL(37.0, x), TCI_MUL(23.0, y));
        double dx = TCI_SUB(x, 0.5), dy = TCI_SUB(y, 0.5);
        double u = TCI_DIV(TCI_ADD(TCI_MUL(dx, dx), TCI_MUL(dy, dy)), 0.1);
        double g = TCI_MUL(tci_cos2pi(ph), tci_exp(-u));
        double h = tci_sin2pi(TCI_MUL(101.0, TCI_MUL(x, y)));
        return TCI_ADD(g, TCI_MUL(0.5, h));
    }
    case TCI_TARGET_SEPCOS: {
        int nt = (int)p[0];
        const double *a = p + 1 + t->nsites;
        double f = TCI_DIV(1.0, TCI_ADD(1.0, s[0]));
        for (int m = 0; m < nt; ++m)
            f = TCI_ADD(f, TCI_MUL(a[m], tci_cos2pi(s[1 + m])));
        return f;
    }
    case TCI_TARGET_TABLE:
        return p[t->nsites + (int64_t)s[0]];
    case TCI_TARGET_QUANTICS1D: {
        int fid = (int)p[1];
        double e = tci_exp(-s[0]);
        if (fid == 1)
            e = TCI_ADD(e, TCI_MUL(1e-3, tci_sin(TCI_MUL(1000.0, s[0]))));
        return e;
    }
    case TCI_TARGET_GKCOSEXP: {
        double s2 = TCI_MUL(s[1], s[1]);
        double e = tci_exp(-TCI_DIV(TCI_MUL(s2, s2), 1000.0));
        double c = tci_cos(TCI_MUL(10.0, s[0]));
        return TCI_MUL(s[2], TCI_MUL(TCI_MUL(1000.0, c), e));
    }
    default:
        return 0.0;
    }
}

/* scalar definition: accumulate sites in order, then finalize. */
TCI_HD double tci_target_eval(const tci_analytic_t *t, const int64_t *v)
{
    double s[TCI_MAX_STATE];
    tci_target_init(t, s);
    for (int k = 0; k < t->nsites; ++k) tci_target_accum(t, k, v[k], s);
    return tci_target_finalize(t, s);
}

/* counter-based generator shared by oracle, library and tests (no
 * Julia RNG; random choices are injected).  splitmix64 finalizer. */
TCI_HD uint64_t tci_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
TCI_HD double tci_uniform01(uint64_t seed, uint64_t index)
{ /* 53-bit uniform in [0,1) */
    uint64_t h = tci_splitmix64(tci_splitmix64(seed) ^ tci_splitmix64(index + 0x632BE59BD9B4E019ull));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

#endif /* TCI_TARGETS_H */
