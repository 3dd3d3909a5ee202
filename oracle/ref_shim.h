/* TEST INFRASTRUCTURE — not part of the product.
 *
 * Host-C shim that lets gcc compile the reference's OpenCL C kernel sources
 * (clode/cpp/{transient,initializeObserver,features,trajectory}.cl and everything
 * they include) unmodified, as plain C11, so that the reference's own code can be
 * executed on the CPU as the parity oracle ("oracle/_ref", see build_ref.py).
 *
 * Only the handful of OpenCL-C tokens the reference uses are provided:
 *   address-space / kernel qualifiers  (transient.cl:9-17)
 *   get_global_id / get_global_size    (transient.cl:19-20)
 *   ulong, bool, clamp                 (clODE_random.cl:28, adaptive_explicit_step.clh:29)
 *   vector typedef names               (realtype.cl:12-15,32-35 — never instantiated)
 *   pown                               (user RHS files, e.g. examples/chay_keizer.cl:34)
 * Math builtins come from <tgmath.h> (glibc libm) or, with -DREF_PM_MATH, from
 * the portable-math header shared with the CUDA "bit-exact" build.
 */
#ifndef CLODE_ORACLE_REF_SHIM_H
#define CLODE_ORACLE_REF_SHIM_H

#include <float.h>
#include <stdbool.h>
#include <tgmath.h>
#undef I /* <complex.h> macro; RHS files are free to use the name */

typedef unsigned long ulong; /* 64-bit on LP64, like OpenCL's ulong */
typedef unsigned int uint;

#define __kernel
#define __global
#define __private
#define __local
#define __constant const
#define cl_khr_fp64 1

/* the vector typedefs in realtype.cl are never used; give the names a meaning */
typedef struct { float s[2]; } float2;
typedef struct { float s[4]; } float4;
typedef struct { float s[8]; } float8;
typedef struct { float s[16]; } float16;
typedef struct { double s[2]; } double2;
typedef struct { double s[4]; } double4;
typedef struct { double s[8]; } double8;
typedef struct { double s[16]; } double16;

/* one "work-item" at a time per host thread */
extern _Thread_local int ref_gid;
extern int ref_gsize;
#define get_global_id(d) (ref_gid)
#define get_global_size(d) (ref_gsize)

/* OpenCL clamp(x, lo, hi) = fmin(fmax(x, lo), hi) (OpenCL C 1.2 spec 6.12.4) */
#define clamp(x, lo, hi) fmin(fmax((x), (lo)), (hi))

#ifdef REF_PM_MATH
/* bit-exact tier: pin the "vendor math library" to the portable implementation */
#include "pm_math.h"
#ifdef CLODE_SINGLE_PRECISION
#error "portable math is double precision only"
#endif
#undef exp
#undef log
#undef pow
#undef cos
#undef sin
#define exp(x) pm_exp(x)
#define log(x) pm_log(x)
#define pow(x, y) pm_pow((x), (y))
#define cos(x) pm_cos(x)
#define sin(x) pm_sin(x)
#endif

/* pown(x, n): integer power by repeated multiplication (left to right), the
 * same definition the CUDA prelude uses, so both sides round identically. */
static inline double ref_pown_d(double x, int n)
{
    int m = n < 0 ? -n : n;
    double r = 1.0;
    for (int k = 0; k < m; ++k)
        r *= x;
    return n < 0 ? 1.0 / r : r;
}
static inline float ref_pown_f(float x, int n)
{
    int m = n < 0 ? -n : n;
    float r = 1.0f;
    for (int k = 0; k < m; ++k)
        r *= x;
    return n < 0 ? 1.0f / r : r;
}
#define pown(x, n) _Generic((x), float: ref_pown_f, default: ref_pown_d)((x), (n))

#endif
