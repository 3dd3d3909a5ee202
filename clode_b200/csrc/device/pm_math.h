/* pm_math.h — portable double-precision exp / log / pow / sin / cos.
 *
 * Purpose: the reference leaves its transcendental builtins to whatever OpenCL
 * device compiler it runs on (call sites: clode/cpp/steppers/adaptive_explicit_step.clh:51,66
 * `pow`, clode/cpp/clODE_random.cl:107 `log`, user RHS `exp`), so the last bit of every
 * result is vendor-defined.  To make "identical accepted-step counts / event counts /
 * RNG streams" a testable statement, the bit-exact build of the CUDA kernels and the
 * CPU oracle both route these five functions to THIS file.  It uses only IEEE-754
 * add/sub/mul/div and integer bit operations, so — as long as the compiler does not
 * contract a*b+c into an FMA (nvcc/NVRTC --fmad=false, gcc -ffp-contract=off) — every
 * operation rounds identically on the GPU and on the CPU.
 *
 * Accuracy (measured against glibc in tests/test_pm_math.py): exp, log, sin, cos
 * <= 2 ulp on the tested ranges; pow(x, y) = exp(y*log x) <= (2 + |y ln x|) ulp,
 * i.e. inside OpenCL C 1.2's 16-ulp bound for pow whenever |y ln x| <= 14
 * (the step-size controller uses |y ln x| < 5).  Special values follow C99 Annex F
 * for the cases an ODE right-hand side can reach (NaN, +-inf, +-0, x<0 with integer y).
 *
 * The file is plain C11 and CUDA C++ at the same time.
 */
#ifndef CLODE_PM_MATH_H
#define CLODE_PM_MATH_H

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define PM_FN static __device__ __forceinline__
#define PM_D2U(x) ((unsigned long long)__double_as_longlong(x))
#define PM_U2D(u) (__longlong_as_double((long long)(u)))
#else
#include <string.h>
#define PM_FN static inline
static inline unsigned long long pm_d2u_(double x) { unsigned long long u; memcpy(&u, &x, 8); return u; }
static inline double pm_u2d_(unsigned long long u) { double x; memcpy(&x, &u, 8); return x; }
#define PM_D2U(x) pm_d2u_(x)
#define PM_U2D(u) pm_u2d_(u)
#endif

#define PM_LN2_HI 6.93147180369123816490e-01 /* 0x3fe62e42fee00000: 32 significant bits */
#define PM_LN2_LO 1.90821492927058770002e-10 /* ln2 - PM_LN2_HI */
#define PM_INV_LN2 1.44269504088896338700e+00
#define PM_INF (PM_U2D(0x7ff0000000000000ULL))
#define PM_NAN (PM_U2D(0x7ff8000000000000ULL))

PM_FN int pm_isnan(double x) { return (PM_D2U(x) & 0x7fffffffffffffffULL) > 0x7ff0000000000000ULL; }

/* 2^k * x for any int k, via exponent-field arithmetic in up to three exact steps */
PM_FN double pm_scale2(double x, int k)
{
    while (k > 1000) { x *= PM_U2D((unsigned long long)(1023 + 1000) << 52); k -= 1000; }
    while (k < -1000) { x *= PM_U2D((unsigned long long)(1023 - 1000) << 52); k += 1000; }
    return x * PM_U2D((unsigned long long)(1023 + k) << 52);
}

PM_FN double pm_exp(double x)
{
    if (pm_isnan(x)) return x;
    if (x > 709.782712893384) return PM_INF;
    if (x < -745.2) return 0.0;
    /* k = nearest integer to x/ln2 (ties irrelevant) */
    double t = x * PM_INV_LN2;
    int k = (int)(t < 0.0 ? t - 0.5 : t + 0.5);
    double kd = (double)k;
    /* Cody-Waite: kd*LN2_HI is exact (|k| <= 1075 needs 11 bits, LN2_HI has 32) */
    double r = (x - kd * PM_LN2_HI) - kd * PM_LN2_LO;
    /* exp(r), |r| <= 0.3466: Taylor to degree 13 (truncation 4e-18), Horner */
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    /* 1 + r + r^2 * p, summed small-to-large */
    double e = 1.0 + (r + (r * r) * p);
    return pm_scale2(e, k);
}

PM_FN double pm_log(double x)
{
    unsigned long long u = PM_D2U(x);
    if (pm_isnan(x)) return x;
    if ((u << 1) == 0) return -PM_INF;        /* log(+-0) = -inf */
    if (u >> 63) return PM_NAN;               /* log(x<0) = NaN  */
    if (u == 0x7ff0000000000000ULL) return x; /* log(inf) = inf  */
    int e = 0;
    if ((u >> 52) == 0) { /* subnormal: normalise */
        x *= 18014398509481984.0; /* 2^54 */
        u = PM_D2U(x);
        e = -54;
    }
    e += (int)(u >> 52) - 1023;
    /* mantissa m in [1,2) */
    u = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m = PM_U2D(u);
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; } /* m in (sqrt2/2, sqrt2] */
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    /* R = 2*(z/3 + z^2/5 + ... + z^11/23), truncation < 1e-18 relative */
    double R = 2.0 / 23.0;
    R = R * z + 2.0 / 21.0;
    R = R * z + 2.0 / 19.0;
    R = R * z + 2.0 / 17.0;
    R = R * z + 2.0 / 15.0;
    R = R * z + 2.0 / 13.0;
    R = R * z + 2.0 / 11.0;
    R = R * z + 2.0 / 9.0;
    R = R * z + 2.0 / 7.0;
    R = R * z + 2.0 / 5.0;
    R = R * z + 2.0 / 3.0;
    R = R * z;
    double hfsq = 0.5 * f * f;
    double ed = (double)e;
    /* log(1+f) = f - (hfsq - s*(hfsq+R)); add e*ln2 split hi/lo */
    return ed * PM_LN2_HI + (f - (hfsq - (s * (hfsq + R) + ed * PM_LN2_LO)));
}

/* y is an integer?  returns 0 = no, 1 = odd integer, 2 = even integer */
PM_FN int pm_int_class(double y)
{
    unsigned long long u = PM_D2U(y) & 0x7fffffffffffffffULL;
    int e = (int)(u >> 52) - 1023;
    if (e < 0) return (u == 0) ? 2 : 0;
    if (e > 52) return 2;
    unsigned long long frac_mask = (e == 52) ? 0ULL : (0x000fffffffffffffULL >> e);
    if (u & frac_mask) return 0;
    unsigned long long unit = 1ULL << (52 - e);
    return (((u & 0x000fffffffffffffULL) | 0x0010000000000000ULL) & unit) ? 1 : 2;
}

/* x^(1/q), q = 3 or 5, for finite x > 0 — the step-size controller's only use of pow():
 * `pow(reltol / normErr, 1/(order+1))` (clode/cpp/steppers/adaptive_explicit_step.clh:51,66).  exp(y log x) costs
 * ~110 uncontracted operations per attempted step, a quarter of a Lorenz dopri5 attempt in the bit-exact tier; a
 * root needs 35: write x = m 2^(qE + r), w = m 2^r in [1, 2^q), start z ~ w^(-1/q) from the exponent-field
 * estimate C - bits(w)/q (< 6 % off), apply the cubically convergent inverse-root correction three times
 *     d = 1 - w z^q,   z <- z (1 + (d/q)(1 + (q+1)/(2q) d)),
 * and return 2^E w z^(q-1).  Multiplications and additions only (every compiler without contraction rounds them
 * identically); <= 5 ulp (q = 3) / <= 11 ulp (q = 5, measured over 2e7 arguments of every binade) of the true root (tests/test_pm_math.py), inside OpenCL C's 16-ulp bound for pow. */
PM_FN double pm_rootq(double x, int q)
{
    unsigned long long u = PM_D2U(x);
    int E = 0;
    if ((u >> 52) == 0) { /* subnormal: times 2^60 (60 = 3*20 = 5*12) */
        x *= 1152921504606846976.0;
        u = PM_D2U(x);
        E = -60 / q;
    }
    int e = (int)(u >> 52) - 1023;
    int Eq = (e >= 0 ? e : e - (q - 1)) / q; /* floor(e / q) */
    int r = e - q * Eq;                      /* 0 <= r < q  */
    E += Eq;
    double w = PM_U2D((u & 0x000fffffffffffffULL) | ((unsigned long long)(1023 + r) << 52));
    /* bits(w^(-1/q)) ~ (1 + 1/q) (1023 - 0.0450) 2^52 - bits(w)/q */
    unsigned long long c = q == 5 ? 0x4CB8BC2FFC470C00ULL : 0x553F09FC6DA44800ULL;
    double z = PM_U2D(c - PM_D2U(w) / (unsigned long long)q);
    const double iq = q == 5 ? 0.2 : 1.0 / 3.0;
    const double hq = q == 5 ? 0.6 : 2.0 / 3.0; /* (q+1)/(2q) */
    for (int it = 0; it < 3; ++it) {
        double z2 = z * z;
        double zq = q == 5 ? (z2 * z2) * z : z2 * z;
        double d = 1.0 - w * zq;
        z = z + z * ((d * iq) * (1.0 + hq * d));
    }
    double z2 = z * z;
    double y = q == 5 ? w * (z2 * z2) : w * z2;
    return y * PM_U2D((unsigned long long)(1023 + E) << 52);
}

PM_FN double pm_pow(double x, double y)
{
    unsigned long long ux = PM_D2U(x), uy = PM_D2U(y);
    if ((uy << 1) == 0) return 1.0;               /* pow(x, +-0) = 1, even for NaN */
    if (ux == 0x3ff0000000000000ULL) return 1.0;  /* pow(1, y) = 1, even for NaN  */
    if (pm_isnan(x) || pm_isnan(y)) return PM_NAN;
    int yneg = (int)(uy >> 63);
    int yclass = pm_int_class(y);
    double ax = PM_U2D(ux & 0x7fffffffffffffffULL);
    int xneg = (int)(ux >> 63);
    double sign = 1.0;
    if (xneg) {
        if ((ux << 1) != 0 && ax != PM_INF && yclass == 0) return PM_NAN; /* (-finite)^(non-integer) */
        if (yclass == 1) sign = -1.0;
    }
    if ((uy & 0x7fffffffffffffffULL) == 0x7ff0000000000000ULL) { /* y = +-inf */
        if (ax == 1.0) return 1.0;
        return ((ax > 1.0) != yneg) ? PM_INF : 0.0;
    }
    if ((ux << 1) == 0) /* x = +-0 */
        return yneg ? sign * PM_INF : sign * 0.0;
    if (ax == PM_INF)
        return yneg ? sign * 0.0 : sign * PM_INF;
    /* x is finite and positive here: a negative finite x with this (non-integer) y returned NaN above */
    if (uy == 0x3fc999999999999aULL) return pm_rootq(ax, 5); /* y = 1/5: dopri5's controller */
    if (uy == 0x3fd5555555555555ULL) return pm_rootq(ax, 3); /* y = 1/3: bs23's controller  */
    double l = pm_log(ax);
    double p = y * l;
    /* recover the rounding error of y*l with a Dekker product so that large |p|
       do not lose bits: p_lo = y*l - p (exact in double-double) */
    const double split = 134217729.0; /* 2^27 + 1 */
    double yh = y * split; yh = yh - (yh - y); double yl = y - yh;
    double lh = l * split; lh = lh - (lh - l); double ll = l - lh;
    double p_lo = ((yh * lh - p) + yh * ll + yl * lh) + yl * ll;
    if (p > 710.0) return sign * PM_INF;
    if (p < -746.0) return sign * 0.0;
    double r = pm_exp(p);
    r = r + r * p_lo;
    return sign * r;
}

/* sin/cos: Cody-Waite reduction by pi/2 in three pieces (exact products for
 * |n| < 2^20, i.e. |x| < ~1.6e6; beyond that accuracy degrades gracefully),
 * then Taylor kernels on [-pi/4, pi/4]. */
#define PM_PIO2_1 1.57079632673412561417e+00 /* first 33 bits of pi/2 */
#define PM_PIO2_2 6.07710050630396597660e-11 /* next 33 bits          */
#define PM_PIO2_3 2.02226624879595063154e-21 /* remainder             */
#define PM_2_OVER_PI 6.36619772367581382433e-01

PM_FN double pm_sin_kernel(double r)
{
    double z = r * r;
    double p = -1.0 / 1307674368000.0;      /* -1/15! */
    p = p * z + 1.0 / 6227020800.0;         /*  1/13! */
    p = p * z - 1.0 / 39916800.0;           /* -1/11! */
    p = p * z + 1.0 / 362880.0;             /*  1/9!  */
    p = p * z - 1.0 / 5040.0;               /* -1/7!  */
    p = p * z + 1.0 / 120.0;                /*  1/5!  */
    p = p * z - 1.0 / 6.0;                  /* -1/3!  */
    return r + r * (z * p);
}

PM_FN double pm_cos_kernel(double r)
{
    double z = r * r;
    double p = 1.0 / 20922789888000.0;      /*  1/16! */
    p = p * z - 1.0 / 87178291200.0;        /* -1/14! */
    p = p * z + 1.0 / 479001600.0;          /*  1/12! */
    p = p * z - 1.0 / 3628800.0;            /* -1/10! */
    p = p * z + 1.0 / 40320.0;              /*  1/8!  */
    p = p * z - 1.0 / 720.0;                /* -1/6!  */
    p = p * z + 1.0 / 24.0;                 /*  1/4!  */
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    /* 1 - z/2 + z^2*p, with the rounding error of (1 - hz) folded back in */
    return w + (((1.0 - w) - hz) + (z * z) * p);
}

PM_FN double pm_sincos_impl(double x, int want_cos)
{
    unsigned long long u = PM_D2U(x) & 0x7fffffffffffffffULL;
    if (u >= 0x7ff0000000000000ULL) return PM_NAN; /* inf or NaN */
    double t = x * PM_2_OVER_PI;
    double nd = (double)(long long)(t < 0.0 ? t - 0.5 : t + 0.5);
    if (u >= 0x4330000000000000ULL) nd = 0.0; /* |x| >= 2^52: give up on reduction */
    double r = ((x - nd * PM_PIO2_1) - nd * PM_PIO2_2) - nd * PM_PIO2_3;
    long long n = (long long)nd;
    int q = (int)(n & 3) + (want_cos ? 1 : 0);
    q &= 3;
    /* q: 0 -> sin r, 1 -> cos r, 2 -> -sin r, 3 -> -cos r */
    double v = (q & 1) ? pm_cos_kernel(r) : pm_sin_kernel(r);
    return (q & 2) ? -v : v;
}

PM_FN double pm_sin(double x) { return pm_sincos_impl(x, 0); }
PM_FN double pm_cos(double x) { return pm_sincos_impl(x, 1); }

#endif /* CLODE_PM_MATH_H */
