// rng.cuh — per-instance random stream, bit-compatible with the reference's
// xoroshiro128+ / polar Box-Muller (clode/cpp/clODE_random.cl:28-114).
//
// The integer recurrence is exact by construction.  The floating-point part is
// written with explicit round-to-nearest intrinsics so that FMA contraction (on in
// the production build) can never change the accept/reject decision of the polar
// method (`w >= 1`, clODE_random.cl:100-105) and desynchronise the stream from the
// CPU oracle; only log() (libdevice vs glibc) can differ in the last bit of the
// returned variate, never in the number of draws consumed.
#ifndef CLODE_RNG_CUH
#define CLODE_RNG_CUH

// production double builds of the stochastic stepper: the polar method's scale factor from fast_polar.cuh
#if defined(CLODE_FAST_POLAR) && defined(STOCHASTIC_EULER) && defined(CLODE_DOUBLE_PRECISION) && !defined(CLODE_BITEXACT) && \
    !defined(CLODE_REFERENCE_MATH) && !defined(__CUDACC_EMU__)
#include "fast_polar.cuh"
#define CLODE_HAVE_FAST_POLAR 1
#else
#define CLODE_HAVE_FAST_POLAR 0
#endif

struct RngStream {
    unsigned long long s0, s1;
    realtype spare;
    bool have_spare;
};

CLODE_DEV unsigned long long rng_next(RngStream &g)
{
    const unsigned long long a = g.s0;
    unsigned long long b = g.s1;
    const unsigned long long out = a + b;
    b ^= a;
    g.s0 = ((a << 55) | (a >> 9)) ^ b ^ (b << 14);
    g.s1 = (b << 36) | (b >> 28);
    return out;
}

#if defined(CLODE_SINGLE_PRECISION)
CLODE_DEV realtype rng_mul(realtype a, realtype b) { return __fmul_rn(a, b); }
CLODE_DEV realtype rng_add(realtype a, realtype b) { return __fadd_rn(a, b); }
CLODE_DEV realtype rng_from_u64(unsigned long long u) { return __ull2float_rn(u); }
#else
CLODE_DEV realtype rng_mul(realtype a, realtype b) { return __dmul_rn(a, b); }
CLODE_DEV realtype rng_add(realtype a, realtype b) { return __dadd_rn(a, b); }
CLODE_DEV realtype rng_from_u64(unsigned long long u) { return __ull2double_rn(u); }
#endif

// uniform on [0,1]: u64 -> real (rounded) times 2^-64  (clODE_random.cl:81-85)
CLODE_DEV realtype rng_uniform(RngStream &g)
{
    return rng_mul(rng_from_u64(rng_next(g)), RCONST(5.421010862427522e-20));
}

// 2 * uniform - 1 (clODE_random.cl:100-101).  The two scalings of the converted integer, by 2^-64 and by 2, are exact
// (powers of two; no underflow: the smallest non-zero value is 2^-64), so RN(RN(RN(u) 2^-64) 2 - 1) = RN(RN(u) 2^-63 - 1):
// ONE fused multiply-add gives the reference's value bit for bit in every tier, instead of two multiplications and an
// addition (2 of the ~150 FP64-pipe instructions of a C4 step, per candidate).
#if defined(CLODE_SINGLE_PRECISION)
CLODE_DEV realtype rng_symmetric(RngStream &g) { return __fmaf_rn(rng_from_u64(rng_next(g)), 1.0842021724855044e-19f, -ONE); }
#else
CLODE_DEV realtype rng_symmetric(RngStream &g) { return __fma_rn(rng_from_u64(rng_next(g)), 1.0842021724855044e-19, -ONE); }
#endif

// N(0,1) by the Marsaglia polar method; the second variate of each pair is kept
CLODE_DEV realtype rng_normal(RngStream &g)
{
    if (g.have_spare) {
        g.have_spare = false;
        return g.spare;
    }
    realtype a, b, q;
    do {
        a = rng_symmetric(g);
        b = rng_symmetric(g);
        q = rng_add(rng_mul(a, a), rng_mul(b, b));
    } while (q >= ONE);
#if CLODE_HAVE_FAST_POLAR
    q = clode_polar_scale(q);
#else
    q = sqrt(rng_mul(-RCONST(2.0), log(q)) / q);
#endif
    g.spare = rng_mul(b, q);
    g.have_spare = true;
    return rng_mul(a, q);
}

#endif // CLODE_RNG_CUH
