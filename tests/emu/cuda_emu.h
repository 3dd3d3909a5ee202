// TEST INFRASTRUCTURE.  Minimal host stand-ins for the CUDA constructs the device headers
// (clode_b200/csrc/device/*.cuh) use, so the very same kernel source can be compiled with g++
// and executed on the CPU by tests/emu/emu_driver.cpp.  This is how device-code logic is
// unit-tested in the GPU-less build container; it is never part of the product path.
#pragma once
#include <math.h>
#include <stddef.h>
#include <string.h>

#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __constant__
#define __restrict__
#define __shared__ static thread_local   /* one host thread runs the threads of a block one after the other */

struct emu_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emu_dim3 blockIdx, blockDim, threadIdx, gridDim;

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct double3 { double x, y, z; };
struct double4 { double x, y, z, w; };

template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, T v) { *p = v; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __ull2double_rn(unsigned long long u) { return (double)u; }
static inline float __ull2float_rn(unsigned long long u) { return (float)u; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float emu__expf(float x) { return expf(x); }
static inline float emu__exp10f(float x) { return powf(10.0f, x); }
static inline float emu__logf(float x) { return logf(x); }
static inline float emu__log2f(float x) { return log2f(x); }
static inline float emu__log10f(float x) { return log10f(x); }
static inline float emu__sinf(float x) { return sinf(x); }
static inline float emu__cosf(float x) { return cosf(x); }
static inline float emu__tanf(float x) { return tanf(x); }
static inline float emu__powf(float x, float y) { return powf(x, y); }
static inline double sinpi(double x) { return sin(M_PI * x); }
static inline double cospi(double x) { return cos(M_PI * x); }
static inline float sinpif(float x) { return sinf((float)M_PI * x); }
static inline float cospif(float x) { return cosf((float)M_PI * x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#define __expf emu__expf
#define __exp10f emu__exp10f
#define __logf emu__logf
#define __log2f emu__log2f
#define __log10f emu__log10f
#define __sinf emu__sinf
#define __cosf emu__cosf
#define __tanf emu__tanf
#define __powf emu__powf
#define __CUDACC_EMU__ 1
static inline unsigned int atomicMax(unsigned int *p, unsigned int v) { unsigned int o = *p; if (v > o) *p = v; return o; }
