// cl_compat.cuh — what a clODE right-hand-side file may assume, restated for CUDA.
//
// clODE hands the user's getRHS source to an OpenCL C compiler after its own device
// code (clode/cpp/CLODE.cpp:144-152), so RHS files — hand written, or emitted by
// clode/function_converter.py:55-124 and clode/xpp_parser.py:73-190 — are free to use
// OpenCL C builtins, the `realtype`/`RCONST` precision macros (clode/cpp/realtype.cl:9-44)
// and clODE's helper functions (clode/cpp/clODE_utilities.cl).  This prelude gives the
// same names a CUDA meaning so those files compile unmodified under NVRTC for sm_100a.
//
// Build macros: CLODE_SINGLE_PRECISION | CLODE_DOUBLE_PRECISION, CLODE_BITEXACT
// (route exp/log/pow/sin/cos to pm_math.h; requires --fmad=false).
#ifndef CLODE_CL_COMPAT_CUH
#define CLODE_CL_COMPAT_CUH

#define CLODE_DEV static __device__ __forceinline__

// ---- precision ---------------------------------------------------------------
#if defined(CLODE_SINGLE_PRECISION)
typedef float realtype;
typedef float2 realtype2;
typedef float3 realtype3;
typedef float4 realtype4;
#define RCONST(x) (x##f)
#define BIG_REAL 3.402823466e+38f
#define SMALL_REAL 1.175494351e-38f
#define UNIT_ROUNDOFF 1.192092896e-07f
#define ZERO 0.0f
#define ONE 1.0f
#elif defined(CLODE_DOUBLE_PRECISION)
typedef double realtype;
typedef double2 realtype2;
typedef double3 realtype3;
typedef double4 realtype4;
#define RCONST(x) (x)
#define BIG_REAL 1.7976931348623157e+308
#define SMALL_REAL 2.2250738585072014e-308
#define UNIT_ROUNDOFF 2.2204460492503131e-16
#define ZERO 0.0
#define ONE 1.0
#else
#error "define CLODE_SINGLE_PRECISION or CLODE_DOUBLE_PRECISION"
#endif

// ---- OpenCL C scalar type names and constants -----------------------------------
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong; // 64-bit under LP64, as in OpenCL C
#ifndef M_PI
#define M_E 2.71828182845904523536
#define M_LOG2E 1.44269504088896340736
#define M_LOG10E 0.434294481903251827651
#define M_LN2 0.693147180559945309417
#define M_LN10 2.30258509299404568402
#define M_PI 3.14159265358979323846
#define M_PI_2 1.57079632679489661923
#define M_PI_4 0.785398163397448309616
#define M_1_PI 0.318309886183790671538
#define M_2_PI 0.636619772367581343076
#define M_2_SQRTPI 1.12837916709551257390
#define M_SQRT2 1.41421356237309504880
#define M_SQRT1_2 0.707106781186547524401
#endif
#define M_E_F 2.718281828f
#define M_PI_F 3.141592654f
#define M_PI_2_F 1.570796327f
#define M_PI_4_F 0.785398163f
#define M_1_PI_F 0.318309886f
#define M_2_PI_F 0.636619772f
#define M_SQRT2_F 1.414213562f
#define M_LN2_F 0.693147181f
#define M_LN10_F 2.302585093f
#define M_LOG2E_F 1.442695041f
#define M_LOG10E_F 0.434294482f
#define MAXFLOAT 3.402823466e+38f
#define FLT_MAX 3.402823466e+38f
#define FLT_MIN 1.175494351e-38f
#define FLT_EPSILON 1.192092896e-07f
#define DBL_MAX 1.7976931348623157e+308
#define DBL_MIN 2.2250738585072014e-308
#define DBL_EPSILON 2.2204460492503131e-16
#ifndef INFINITY
#define INFINITY (__int_as_float(0x7f800000))
#endif
#ifndef NAN
#define NAN (__int_as_float(0x7fffffff))
#endif

// ---- address-space qualifiers (no meaning for per-thread CUDA code) --------------
#define __global
#define __private
#define __local
#define __constant const
#define __kernel

// ---- math builtins OpenCL C has and CUDA lacks -----------------------------------
// (float and double overloads; everything else — exp, pow, sinpi, cbrt, erf, tgamma,
// ldexp, ilogb, nextafter, fdim, remainder, rint, ... — exists in CUDA with the same name)
CLODE_DEV double clamp(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
CLODE_DEV float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
CLODE_DEV double mad(double a, double b, double c) { return a * b + c; }
CLODE_DEV float mad(float a, float b, float c) { return a * b + c; }
CLODE_DEV double mix(double a, double b, double t) { return a + (b - a) * t; }
CLODE_DEV float mix(float a, float b, float t) { return a + (b - a) * t; }
CLODE_DEV double step(double edge, double x) { return x < edge ? 0.0 : 1.0; }
CLODE_DEV float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
CLODE_DEV double sign(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : (x == 0.0 ? x : 0.0)); }
CLODE_DEV float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : (x == 0.0f ? x : 0.0f)); }
CLODE_DEV double degrees(double r) { return r * (180.0 / M_PI); }
CLODE_DEV float degrees(float r) { return r * (180.0f / M_PI_F); }
CLODE_DEV double radians(double d) { return d * (M_PI / 180.0); }
CLODE_DEV float radians(float d) { return d * (M_PI_F / 180.0f); }
CLODE_DEV double smoothstep(double e0, double e1, double x)
{
    double t = clamp((x - e0) / (e1 - e0), 0.0, 1.0);
    return t * t * (3.0 - 2.0 * t);
}
CLODE_DEV float smoothstep(float e0, float e1, float x)
{
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
// pown: integer power by repeated multiplication, left to right (the CPU oracle uses
// the same definition, so both sides round identically); exponent is usually a literal
// and the loop unrolls away.
CLODE_DEV double pown(double x, int n)
{
    int m = n < 0 ? -n : n;
    double r = 1.0;
    for (int k = 0; k < m; ++k)
        r *= x;
    return n < 0 ? 1.0 / r : r;
}
CLODE_DEV float pown(float x, int n)
{
    int m = n < 0 ? -n : n;
    float r = 1.0f;
    for (int k = 0; k < m; ++k)
        r *= x;
    return n < 0 ? 1.0f / r : r;
}
CLODE_DEV double pown(double x, double n) { return pown(x, (int)n); }
CLODE_DEV float pown(float x, float n) { return pown(x, (int)n); }
CLODE_DEV double powr(double x, double y) { return x < 0.0 ? (double)NAN : pow(x, y); }
CLODE_DEV float powr(float x, float y) { return x < 0.0f ? NAN : powf(x, y); }
CLODE_DEV double rootn(double x, int n)
{
    if (n == 2) return sqrt(x);
    if (n == 3) return cbrt(x);
    if (x < 0.0 && (n & 1)) return -pow(-x, 1.0 / (double)n);
    return pow(x, 1.0 / (double)n);
}
CLODE_DEV float rootn(float x, int n)
{
    if (n == 2) return sqrtf(x);
    if (n == 3) return cbrtf(x);
    if (x < 0.0f && (n & 1)) return -powf(-x, 1.0f / (float)n);
    return powf(x, 1.0f / (float)n);
}
CLODE_DEV double rootn(double x, double n) { return rootn(x, (int)n); }
CLODE_DEV float rootn(float x, float n) { return rootn(x, (int)n); }
CLODE_DEV double acospi(double x) { return acos(x) * M_1_PI; }
CLODE_DEV float acospi(float x) { return acosf(x) * M_1_PI_F; }
CLODE_DEV double asinpi(double x) { return asin(x) * M_1_PI; }
CLODE_DEV float asinpi(float x) { return asinf(x) * M_1_PI_F; }
CLODE_DEV double atanpi(double x) { return atan(x) * M_1_PI; }
CLODE_DEV float atanpi(float x) { return atanf(x) * M_1_PI_F; }
CLODE_DEV double atan2pi(double y, double x) { return atan2(y, x) * M_1_PI; }
CLODE_DEV float atan2pi(float y, float x) { return atan2f(y, x) * M_1_PI_F; }
CLODE_DEV double tanpi(double x) { return sinpi(x) / cospi(x); }
CLODE_DEV float tanpi(float x) { return sinpif(x) / cospif(x); }
CLODE_DEV double fract(double x) { return fmin(x - floor(x), 0.99999999999999988898); }
CLODE_DEV float fract(float x) { return fminf(x - floorf(x), 0.99999994f); }
CLODE_DEV double maxmag(double a, double b) { return fabs(a) > fabs(b) ? a : (fabs(b) > fabs(a) ? b : fmax(a, b)); }
CLODE_DEV double minmag(double a, double b) { return fabs(a) < fabs(b) ? a : (fabs(b) < fabs(a) ? b : fmin(a, b)); }
CLODE_DEV float maxmag(float a, float b) { return fabsf(a) > fabsf(b) ? a : (fabsf(b) > fabsf(a) ? b : fmaxf(a, b)); }
CLODE_DEV float minmag(float a, float b) { return fabsf(a) < fabsf(b) ? a : (fabsf(b) < fabsf(a) ? b : fminf(a, b)); }
// select(a, b, c): b where c is true (scalar form)
CLODE_DEV double select(double a, double b, long c) { return c ? b : a; }
CLODE_DEV float select(float a, float b, int c) { return c ? b : a; }
// mixed-precision calls that OpenCL front ends accept (e.g. pow(x, -1.5f) emitted by
// xpp_parser with double realtype, test/xpp/van_der_pol_oscillator_reference.cl)
CLODE_DEV double pow(double x, float y) { return pow(x, (double)y); }
CLODE_DEV double pow(float x, double y) { return pow((double)x, y); }
CLODE_DEV double fmax(double a, float b) { return fmax(a, (double)b); }
CLODE_DEV double fmax(float a, double b) { return fmax((double)a, b); }
CLODE_DEV double fmin(double a, float b) { return fmin(a, (double)b); }
CLODE_DEV double fmin(float a, double b) { return fmin((double)a, b); }

// native_* / half_* : reduced-accuracy variants — OpenCL leaves their accuracy
// implementation-defined; map float versions to the SFU intrinsics, double to full precision.
CLODE_DEV float native_exp(float x) { return __expf(x); }
CLODE_DEV float native_exp2(float x) { return exp2f(x); }
CLODE_DEV float native_exp10(float x) { return __exp10f(x); }
CLODE_DEV float native_log(float x) { return __logf(x); }
CLODE_DEV float native_log2(float x) { return __log2f(x); }
CLODE_DEV float native_log10(float x) { return __log10f(x); }
CLODE_DEV float native_sin(float x) { return __sinf(x); }
CLODE_DEV float native_cos(float x) { return __cosf(x); }
CLODE_DEV float native_tan(float x) { return __tanf(x); }
CLODE_DEV float native_sqrt(float x) { return sqrtf(x); }
CLODE_DEV float native_rsqrt(float x) { return rsqrtf(x); }
CLODE_DEV float native_recip(float x) { return __frcp_rn(x); }
CLODE_DEV float native_divide(float a, float b) { return __fdividef(a, b); }
CLODE_DEV float native_powr(float x, float y) { return __powf(x, y); }
CLODE_DEV double native_exp(double x) { return exp(x); }
CLODE_DEV double native_exp2(double x) { return exp2(x); }
CLODE_DEV double native_exp10(double x) { return exp10(x); }
CLODE_DEV double native_log(double x) { return log(x); }
CLODE_DEV double native_log2(double x) { return log2(x); }
CLODE_DEV double native_log10(double x) { return log10(x); }
CLODE_DEV double native_sin(double x) { return sin(x); }
CLODE_DEV double native_cos(double x) { return cos(x); }
CLODE_DEV double native_tan(double x) { return tan(x); }
CLODE_DEV double native_sqrt(double x) { return sqrt(x); }
CLODE_DEV double native_rsqrt(double x) { return rsqrt(x); }
CLODE_DEV double native_recip(double x) { return 1.0 / x; }
CLODE_DEV double native_divide(double a, double b) { return a / b; }
CLODE_DEV double native_powr(double x, double y) { return pow(x, y); }
#define half_exp native_exp
#define half_exp2 native_exp2
#define half_exp10 native_exp10
#define half_log native_log
#define half_log2 native_log2
#define half_log10 native_log10
#define half_sin native_sin
#define half_cos native_cos
#define half_tan native_tan
#define half_sqrt native_sqrt
#define half_rsqrt native_rsqrt
#define half_recip native_recip
#define half_divide native_divide
#define half_powr native_powr

// ---- clODE helper functions visible to RHS code (clode/cpp/clODE_utilities.cl) ----
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#define heaviside(x) ((x) >= ZERO ? ONE : ZERO)

CLODE_DEV realtype norm_1(realtype v[], int N) // clODE_utilities.cl:22-28
{
    realtype s = ZERO;
    for (int k = 0; k < N; k++) s += fabs(v[k]);
    return s;
}
CLODE_DEV realtype norm_2(realtype v[], int N) // clODE_utilities.cl:31-37
{
    realtype s = ZERO;
    for (int k = 0; k < N; k++) s += v[k] * v[k];
    return sqrt(s);
}
CLODE_DEV realtype norm_inf(realtype v[], int N) // clODE_utilities.cl:40-46
{
    realtype s = ZERO;
    for (int k = 0; k < N; k++) s = fmax(fabs(v[k]), s);
    return s;
}
// first-occurrence extrema (clODE_utilities.cl:49-128)
CLODE_DEV void maxOfArray(realtype v[], int N, realtype *best, int *at)
{
    *best = -BIG_REAL; *at = 0;
    for (int k = 0; k < N; k++) if (v[k] > *best) { *best = v[k]; *at = k; }
}
CLODE_DEV void minOfArray(realtype v[], int N, realtype *best, int *at)
{
    *best = BIG_REAL; *at = 0;
    for (int k = 0; k < N; k++) if (v[k] < *best) { *best = v[k]; *at = k; }
}
CLODE_DEV realtype array_max(realtype v[], int N) { realtype b; int a; maxOfArray(v, N, &b, &a); return b; }
CLODE_DEV realtype array_min(realtype v[], int N) { realtype b; int a; minOfArray(v, N, &b, &a); return b; }
CLODE_DEV int array_argmax(realtype v[], int N) { realtype b; int a; maxOfArray(v, N, &b, &a); return a; }
CLODE_DEV int array_argmin(realtype v[], int N) { realtype b; int a; minOfArray(v, N, &b, &a); return a; }
// running statistics (clODE_utilities.cl:167-195)
CLODE_DEV realtype runningMeanTime(realtype mean, realtype v, realtype dt, realtype span)
{
    return mean + (v - mean) * dt / span;
}
CLODE_DEV void runningMean(realtype *mean, realtype v, unsigned int count)
{
    if (count == 1) *mean = v;
    else if (count > 1) *mean += (v - *mean) / (realtype)count;
}
CLODE_DEV void runningMeanVar(realtype *mean, realtype *var, realtype v, unsigned int count)
{
    if (count == 1) { *mean = v; *var = ZERO; }
    else if (count > 1) {
        realtype old = *mean;
        *mean = old + (v - old) / (realtype)count;
        *var = *var + (v - old) * (v - *mean);
    }
}
// interpolation (clODE_utilities.cl:201-241)
CLODE_DEV realtype linearInterp(realtype t0, realtype t1, realtype y0, realtype y1, realtype ti)
{
    return y0 + (ti - t0) * (y1 - y0) / (t1 - t0);
}
CLODE_DEV realtype linearInterpArray(realtype t[], realtype y[], realtype ti)
{
    return ti < t[1] ? linearInterp(t[0], t[1], y[0], y[1], ti) : linearInterp(t[1], t[2], y[1], y[2], ti);
}
CLODE_DEV realtype quadraticInterp(realtype t[], realtype y[], realtype ti)
{
    realtype b0 = y[0];
    realtype b1 = (y[1] - b0) / (t[1] - t[0]);
    realtype b2 = (y[2] - b0 - b1 * (t[2] - t[0])) / ((t[2] - t[0]) * (t[2] - t[1]));
    return b0 + b1 * (ti - t[0]) + b2 * (ti - t[0]) * (ti - t[1]);
}
CLODE_DEV void quadraticInterpVertex(realtype t[], realtype y[], realtype *tv, realtype *yv)
{
    realtype b0 = y[0];
    realtype b1 = (y[1] - b0) / (t[1] - t[0]);
    realtype b2 = (y[2] - b0 - b1 * (t[2] - t[0])) / ((t[2] - t[0]) * (t[2] - t[1]));
    *tv = -(b1 - b2 * (t[0] + t[1])) / (RCONST(2.0) * b2);
    *yv = b0 + b1 * (*tv - t[0]) + b2 * (*tv - t[0]) * (*tv - t[1]);
}

// ---- bit-exact tier: pin the transcendental builtins -------------------------------
#ifdef CLODE_BITEXACT
#ifndef CLODE_DOUBLE_PRECISION
#error "CLODE_BITEXACT requires double precision"
#endif
#include "pm_math.h"
#define exp(x) pm_exp(x)
#define log(x) pm_log(x)
#define pow(x, y) pm_pow((x), (y))
#define cos(x) pm_cos(x)
#define sin(x) pm_sin(x)
#endif

// ---- production double: exp() by table + short polynomial (fast_exp.cuh), unless the program asks for the library's
#if defined(CLODE_DOUBLE_PRECISION) && !defined(CLODE_BITEXACT) && !defined(CLODE_REFERENCE_MATH) && \
    !defined(CLODE_LIBRARY_EXP) && !defined(__CUDACC_EMU__)
#include "fast_exp.cuh"
#define exp(x) clode_fast_exp(x)
#define CLODE_HAVE_FAST_EXP 1
#else
#define CLODE_HAVE_FAST_EXP 0
#endif

#endif // CLODE_CL_COMPAT_CUH
