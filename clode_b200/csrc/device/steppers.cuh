// steppers.cuh — time steppers for one ODE instance held entirely in registers.
//
// Behaviour follows clode/cpp/steppers/*.clh exactly (same tableaux, same left-to-right
// evaluation order, same controller), but the control structure is different:
//   * fixed-step methods: one call = one step (fixed_explicit_step.clh:9-37);
//   * adaptive methods: one call = ONE ATTEMPT.  The reference retries rejected
//     steps in a per-work-item inner loop (adaptive_explicit_step.clh:20-62), which
//     under SIMT makes 31 lanes idle while one lane retries.  Here the kernel's time
//     loop is an attempt loop: every live lane performs one trial step per
//     iteration, accepted lanes commit and run the observer, rejected lanes only
//     shrink h.  Per-instance arithmetic and its order are unchanged, so accepted-step
//     counts are identical to the reference's.
//
// Stepper selection by macro, as in clode/cpp/steppers.cl:29-34:
//   EXPLICIT_EULER, EXPLICIT_HEUN, EXPLICIT_RK4, EXPLICIT_BS23, EXPLICIT_DOPRI5, STOCHASTIC_EULER
#ifndef CLODE_STEPPERS_CUH
#define CLODE_STEPPERS_CUH

#if defined(STOCHASTIC_EULER)
#define CLODE_STOCHASTIC 1
#else
#define CLODE_STOCHASTIC 0
#endif
#if defined(EXPLICIT_BS23) || defined(EXPLICIT_DOPRI5)
#define CLODE_ADAPTIVE 1
#else
#define CLODE_ADAPTIVE 0
#endif

#define NV N_VAR
#define NP_ (N_PAR > 0 ? N_PAR : 1)
#define NA_ (N_AUX > 0 ? N_AUX : 1)
#define NW_ (N_WIENER > 0 ? N_WIENER : 1)

// struct SolverParams — clode/cpp/clODE_struct_defs.cl:11-20 (rebuilt per thread from the constant argument block)
struct SolverParams {
    realtype dt, dtmax, abstol, reltol;
    unsigned int max_steps, max_store, nout;
};

// Arithmetic tier.  1: every engine expression is evaluated as the reference writes it (bit-exact tier,
// single precision, reference-math builds, the host emulation).  0: production double — same algorithm,
// cheaper instruction sequences where noted (DESIGN.md §3 "instruction diet").
#if defined(CLODE_BITEXACT) || defined(CLODE_SINGLE_PRECISION) || defined(CLODE_REFERENCE_MATH) || defined(__CUDACC_EMU__)
#define CLODE_EXACT_ARITH 1
#else
#define CLODE_EXACT_ARITH 0
#endif
// Production single precision (the reference's Python default): FP32 has native min/max/abs, so only the
// engine's own divisions, the controller root and the step floor get cheaper sequences (all within a few ulp).
#if defined(CLODE_SINGLE_PRECISION) && !defined(CLODE_REFERENCE_MATH) && !defined(__CUDACC_EMU__)
#define CLODE_FAST_SINGLE 1
#else
#define CLODE_FAST_SINGLE 0
#endif

// fmax/fmin where the SECOND operand is known not to be NaN (a running extreme that started finite, a
// solver parameter, ...).  Same value as fmax(v, m) / fmin(v, m) for every input — a NaN v yields m —
// except possibly the sign of a zero result.  In SASS the library fmax on doubles is DSETP + FSEL +
// SEL + LOP3 + moves (~7 issue slots, NaN quieting included); this form is DSETP + 2 SEL.  The step loop
// evaluates ~20 of them per attempt.
#if defined(CLODE_SINGLE_PRECISION)
// FP32 has native min/max instructions (FMNMX) with exactly these NaN semantics
CLODE_DEV realtype max_nn(const realtype v, const realtype m) { return fmaxf(v, m); }
CLODE_DEV realtype min_nn(const realtype v, const realtype m) { return fminf(v, m); }
#else
CLODE_DEV realtype max_nn(const realtype v, const realtype m) { return v > m ? v : m; }
CLODE_DEV realtype min_nn(const realtype v, const realtype m) { return v < m ? v : m; }
#endif
CLODE_DEV realtype clamp_nn(const realtype v, const realtype lo, const realtype hi) { return min_nn(max_nn(v, lo), hi); }

#if CLODE_EXACT_ARITH
CLODE_DEV realtype abs_nn(const realtype v) { return fabs(v); }
CLODE_DEV realtype opaque(const realtype c) { return c; }
#else
// |v| on the bit pattern: one LOP3 on the integer pipe.  fabs() of a double that is then selected on
// compiles to DADD -RZ,|v|, i.e. it occupies the FP64 pipe — the bottleneck of the step loop — for a move.
CLODE_DEV double abs_nn(const double v) { return __hiloint2double(__double2hiint(v) & 0x7fffffff, __double2loint(v)); }
// a constant the optimiser cannot see: `v > c ? v : c` with a literal c is rewritten into fmax(v, c), whose
// double-precision expansion (DSETP.MAX + NaN quieting) costs 5 more issue slots than compare + select
CLODE_DEV double opaque(const double c) { double r; asm("mov.f64 %0, %1;" : "=d"(r) : "d"(c)); return r; }
#endif

// a / b for the engine's own bookkeeping divisions (error normalisation, running means).
// Reference-arithmetic builds: the IEEE division, as written in the reference.  Production double:
// reciprocal by MUFU.RCP64H + two Newton steps, then one multiply — <= 2 ulp, no denormal / overflow
// slow path (the divisors here are max(|x|, abstol/reltol) and elapsed times: normal, finite numbers).
// libdevice's correctly-rounded division costs 8 FP64 + ~6 control instructions per call and a Lorenz
// dopri5 attempt makes four of them.  The user's RHS keeps the IEEE `/`.
#if CLODE_FAST_SINGLE
// MUFU.RCP + one Newton step + multiply: <= 1.5 ulp, branch-free (the IEEE float division is ~10 instructions
// with a range check and a slow-path call)
CLODE_DEV float div_nr(const float a, const float b)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    r = fmaf(r, fmaf(-b, r, 1.0f), r);
    return a * r;
}
#elif CLODE_EXACT_ARITH
CLODE_DEV realtype div_nr(const realtype a, const realtype b) { return a / b; }
#else
CLODE_DEV double div_nr(const double a, const double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(r, fma(-b, r, 1.0), r);
    r = fma(r, fma(-b, r, 1.0), r);
    return a * r;
}
#endif

// err / scale of the step-size controller's error norm.  The quotient feeds a comparison with reltol and a clamped
// fifth (third) root, so production double stops after ONE Newton step on the SFU seed (relative error < 1e-11,
// measured on the GPU against the IEEE quotient, tests/test_fast_exp.py): 3 FP64-pipe instructions per variable instead
// of 5 (C2 58.99 -> 57.38 ms).  Moving the norm's and the step clamps' compares to the integer pipe (bit-pattern order of
// non-negative doubles) was tried as well: 23 fewer FP64-pipe instructions but 27 more issue slots, and slower
// (59.47 ms) — with the divisions trimmed the loop is issue-bound, not FP64-pipe-bound
// (profiles/r01_integer_compares_and_norm_division_ab.log).
#if CLODE_EXACT_ARITH || CLODE_FAST_SINGLE
CLODE_DEV realtype div_norm(const realtype a, const realtype b) { return div_nr(a, b); }
#else
CLODE_DEV double div_norm(const double a, const double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(r, fma(-b, r, 1.0), r);
    return a * r;
}
#endif

// user right-hand side; the definition is appended after all engine code (clode/cpp/steppers.cl:50)
__device__ __forceinline__ void getRHS(const realtype t, const realtype x_[], const realtype p_[],
                                       realtype dx_[], realtype aux_[], const realtype w_[]);

// everything one instance owns while it is being integrated
struct Instance {
    realtype t, dt;
    realtype x[NV], k1[NV];
    realtype p[NP_], aux[NA_], w[NW_];
    RngStream rng;
};

// Noise of the next step, fixed_explicit_step.clh:29-33: w = randn / sqrt(dt).  dt is constant inside a kernel for the
// (fixed-step) stochastic method, so production double divides ONCE per kernel (the compiler hoists the square root
// but must keep the IEEE division in the loop) and multiplies by the reciprocal: the variate moves by at most one
// ulp, the integer RNG stream not at all.
CLODE_DEV void draw_noise(Instance &I)
{
#if CLODE_STOCHASTIC && !CLODE_EXACT_ARITH
    const realtype scale = ONE / sqrt(I.dt);
#endif
#pragma unroll
    for (int j = 0; j < N_WIENER; ++j)
#if CLODE_STOCHASTIC && !CLODE_EXACT_ARITH
        I.w[j] = rng_normal(I.rng) * scale;
#elif CLODE_STOCHASTIC
        I.w[j] = rng_normal(I.rng) / sqrt(I.dt);
#else
        I.w[j] = ZERO;
#endif
}

// ---------------------------------------------------------------------------------
#if !CLODE_ADAPTIVE

CLODE_DEV void step_fixed(Instance &I)
{
    const realtype h = I.dt;
#if defined(EXPLICIT_EULER) || defined(STOCHASTIC_EULER)
    // fixed_explicit_Euler.clh:5-18
#pragma unroll
    for (int j = 0; j < NV; ++j)
        I.x[j] += h * I.k1[j];
    I.t += h;
#elif defined(EXPLICIT_HEUN)
    // fixed_explicit_Trapezoidal.clh:5-23 (explicit fma in the predictor, as in the reference)
    const realtype t1 = I.t + h;
    realtype y[NV], k2[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = fma(h, I.k1[j], I.x[j]);
    getRHS(t1, y, I.p, k2, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        I.x[j] += h * RCONST(0.5) * (I.k1[j] + k2[j]);
    I.t = t1;
#elif defined(EXPLICIT_RK4)
    // fixed_explicit_RK4.clh:5-40
    const realtype hh = h * RCONST(0.5);
    const realtype tm = I.t + hh, t1 = I.t + h;
    realtype y[NV], k2[NV], k3[NV], k4[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + hh * I.k1[j];
    getRHS(tm, y, I.p, k2, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + hh * k2[j];
    getRHS(tm, y, I.p, k3, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * k3[j];
    getRHS(t1, y, I.p, k4, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        I.x[j] += h * (I.k1[j] + RCONST(2.0) * k2[j] + RCONST(2.0) * k3[j] + k4[j]) / RCONST(6.0);
    I.t = t1;
#else
#error "no stepper selected"
#endif
    // wrapper, fixed_explicit_step.clh:26-34: new noise, then the slope at the new point
    draw_noise(I);
    getRHS(I.t, I.x, I.p, I.k1, I.aux, I.w);
}

#else // CLODE_ADAPTIVE ---------------------------------------------------------------

// The error estimate is h * (sum of e_i k_i) per variable, and the controller only uses its weighted maximum norm
// max_j |err_j| / scale_j.  The effective step h is positive, so production double takes it out of the maximum — one
// multiplication per attempt instead of one per variable (the norm moves by an ulp; the bit-exact tier and single
// precision scale every component as the reference writes it, adaptive_bs23.clh:60, adaptive_dp45.clh:107).
#if CLODE_EXACT_ARITH || CLODE_FAST_SINGLE
#define CLODE_NORM_SCALED_ONCE 0
#define CLODE_ERR_SCALE(h) (h) *
#else
#define CLODE_NORM_SCALED_ONCE 1
#define CLODE_ERR_SCALE(h)
#endif

#if defined(EXPLICIT_BS23)
#define ERR_ORDER RCONST(2.0)
#define MAX_SHRINK RCONST(0.5)
// Tableau in the constant bank: each coefficient is then a c[bank][offset] operand of the DFMA/DMUL
// that uses it.  As literals the compiler re-materialises every 64-bit constant with two UMOVs per
// use inside the attempt loop (12 % of all issued instructions in the first profile).  Values are the
// reference's own quotient expressions, folded at compile time (adaptive_bs23.clh:9-25).
__constant__ realtype clode_tab[9] = {
    RCONST(0.5), RCONST(0.75),
    RCONST(2.0) / RCONST(9.0), RCONST(1.0) / RCONST(3.0), RCONST(4.0) / RCONST(9.0),
    RCONST(-5.0) / RCONST(72.0), RCONST(1.0) / RCONST(12.0), RCONST(1.0) / RCONST(9.0), RCONST(-1.0) / RCONST(8.0)};
// Bogacki-Shampine 3(2), adaptive_bs23.clh:27-63.  Returns the effective step.
CLODE_DEV realtype trial_step(Instance &I, const realtype h_in, realtype &t1, realtype xn[NV],
                              realtype kn[NV], realtype err[NV])
{
    const realtype *c = clode_tab;
    t1 = I.t + h_in;
    const realtype h = t1 - I.t;
    realtype y[NV], k2[NV], k3[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * c[0] * I.k1[j];
    getRHS(I.t + h * c[0], y, I.p, k2, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * c[1] * k2[j];
    getRHS(I.t + h * c[1], y, I.p, k3, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        xn[j] = I.x[j] + h * (c[2] * I.k1[j] + c[3] * k2[j] + c[4] * k3[j]);
    getRHS(t1, xn, I.p, kn, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        err[j] = CLODE_ERR_SCALE(h) (c[5] * I.k1[j] + c[6] * k2[j] + c[7] * k3[j] + c[8] * kn[j]);
    return h;
}
#else // EXPLICIT_DOPRI5
#define ERR_ORDER RCONST(4.0)
#define MAX_SHRINK RCONST(0.1)
// Tableau in the constant bank (see the note at the bs23 table); adaptive_dp45.clh:10-56.
enum { A2, A3, A4, A5, B21, B31, B32, B41, B42, B43, B51, B52, B53, B54, B61, B62, B63, B64, B65,
       C1, C3, C4, C5, C6, E1, E3, E4, E5, E6, E7, DP_TAB_SIZE };
__constant__ realtype clode_tab[DP_TAB_SIZE] = {
    RCONST(1.0) / RCONST(5.0), RCONST(3.0) / RCONST(10.0), RCONST(4.0) / RCONST(5.0), RCONST(8.0) / RCONST(9.0),
    RCONST(1.0) / RCONST(5.0),
    RCONST(3.0) / RCONST(40.0), RCONST(9.0) / RCONST(40.0),
    RCONST(44.0) / RCONST(45.0), RCONST(-56.0) / RCONST(15.0), RCONST(32.0) / RCONST(9.0),
    RCONST(19372.0) / RCONST(6561.0), RCONST(-25360.0) / RCONST(2187.0), RCONST(64448.0) / RCONST(6561.0),
    RCONST(-212.0) / RCONST(729.0),
    RCONST(9017.0) / RCONST(3168.0), RCONST(-355.0) / RCONST(33.0), RCONST(46732.0) / RCONST(5247.0),
    RCONST(49.0) / RCONST(176.0), RCONST(-5103.0) / RCONST(18656.0),
    RCONST(35.0) / RCONST(384.0), RCONST(500.0) / RCONST(1113.0), RCONST(125.0) / RCONST(192.0),
    RCONST(-2187.0) / RCONST(6784.0), RCONST(11.0) / RCONST(84.0),
    RCONST(71.0) / RCONST(57600.0), RCONST(-71.0) / RCONST(16695.0), RCONST(71.0) / RCONST(1920.0),
    RCONST(-17253.0) / RCONST(339200.0), RCONST(22.0) / RCONST(525.0), RCONST(-1.0) / RCONST(40.0)};
// Dormand-Prince 5(4), adaptive_dp45.clh:58-110.  Returns the effective step.
CLODE_DEV realtype trial_step(Instance &I, const realtype h_in, realtype &t1, realtype xn[NV],
                              realtype kn[NV], realtype err[NV])
{
    const realtype *c = clode_tab;
    t1 = I.t + h_in;
    const realtype h = t1 - I.t;
    realtype y[NV], k2[NV], k3[NV], k4[NV], k5[NV], k6[NV];
    const realtype *k1 = I.k1;
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (c[B21] * k1[j]);
    getRHS(I.t + c[A2] * h, y, I.p, k2, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (c[B31] * k1[j] + c[B32] * k2[j]);
    getRHS(I.t + c[A3] * h, y, I.p, k3, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (c[B41] * k1[j] + c[B42] * k2[j] + c[B43] * k3[j]);
    getRHS(I.t + c[A4] * h, y, I.p, k4, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (c[B51] * k1[j] + c[B52] * k2[j] + c[B53] * k3[j] + c[B54] * k4[j]);
    getRHS(I.t + c[A5] * h, y, I.p, k5, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (c[B61] * k1[j] + c[B62] * k2[j] + c[B63] * k3[j] + c[B64] * k4[j] + c[B65] * k5[j]);
    getRHS(I.t + h, y, I.p, k6, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        xn[j] = I.x[j] + h * (c[C1] * k1[j] + c[C3] * k3[j] + c[C4] * k4[j] + c[C5] * k5[j] + c[C6] * k6[j]);
    getRHS(t1, xn, I.p, kn, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        err[j] = CLODE_ERR_SCALE(h) (c[E1] * k1[j] + c[E3] * k3[j] + c[E4] * k4[j] + c[E5] * k5[j] + c[E6] * k6[j] + c[E7] * kn[j]);
    return h;
}
#endif

// 0.8 * (reltol/err)^(1/(order+1)) for the step-size controller (adaptive_explicit_step.clh:51,66).
// Bit-exact tier, reference-math builds and single precision evaluate it as written, with pow().
// Production double: libdevice's pow() costs ~100 FP64-pipe instructions on a ~40-deep dependent
// chain, plus an IEEE division for its argument — a third of a Lorenz dopri5 step — for a factor
// that is immediately clamped to [MAX_SHRINK, 5].  There the factor is computed division-free as
//     0.8 * reltol^(1/(p+1)) * err^(-1/(p+1)),
// the first two terms once per kernel, the last as an SFU seed (FP32 lg2/ex2) times one third-order
// series correction: 8-9 FP64 instructions on a short chain, <= 2 ulp, well inside OpenCL C's 16-ulp bound
// for pow.  err outside [1e-30, 1e30] saturates, which the clamps make indistinguishable.
struct Controller {
    realtype reltol, floor_;
    realtype scale; // 0.8 * reltol^(1/(p+1))   (production double only)
    int floor_hi_min; // step_floor's exponent-field form applies to hi(t) in [floor_hi_min, 0x7ff00000): empty unless t_end > 0
    int floor_hi_max; // production double: ... and 16 ulp(t) <= dtmax, i.e. hi(t) < floor_hi_max (adaptive_attempt)
};
#if CLODE_FAST_SINGLE
#define CLODE_EXACT_CONTROLLER 0
// single precision: the SFU pair lg2/ex2 alone is within ~8 ulp of pow (OpenCL C allows 16)
CLODE_DEV float controller_factor(const Controller &c, float nerr)
{
    const float x = fminf(fmaxf(nerr, 1e-30f), 1e30f);
    return c.scale * exp2f((-1.0f / (ERR_ORDER + 1.0f)) * __log2f(x));
}
#elif defined(CLODE_BITEXACT) || defined(CLODE_SINGLE_PRECISION) || defined(CLODE_REFERENCE_MATH)
#define CLODE_EXACT_CONTROLLER 1
CLODE_DEV realtype controller_factor(const Controller &c, realtype nerr)
{
    return RCONST(0.8) * pow(c.reltol / nerr, RCONST(1.0) / (ERR_ORDER + RCONST(1.0)));
}
#else
#define CLODE_EXACT_CONTROLLER 0
CLODE_DEV double controller_factor(const Controller &c, double nerr)
{
    // clamp to about [1e-30, 1e30] on the high word (nerr is >= 0 and never NaN: integer order == value order);
    // any value out there saturates the [MAX_SHRINK, 5] clamp that follows, so the low word may stay
    const int hi = min(max(__double2hiint(nerr), 0x39b4484b), 0x46293e59);
    const double x = __hiloint2double(hi, __double2loint(nerr));
    // seed z ~ x^(-1/q) from the SFU (relative error < 1e-5), then ONE third-order correction: with
    // d = x z^q - 1 the exact root is z (1 + d)^(-1/q), expanded to d^3 (|d| < 5e-5: truncation < 1e-18)
#if defined(EXPLICIT_BS23)
    double z = (double)exp2f(-0.33333334f * __log2f((float)x));
    const double d = fma(x, z * z * z, -1.0);
    z = fma(z, d * fma(d, fma(d, -14.0 / 81.0, 2.0 / 9.0), -1.0 / 3.0), z); // (1+d)^(-1/3) = 1 - d/3 + 2/9 d^2 - 14/81 d^3
#else
    double z = (double)exp2f(-0.2f * __log2f((float)x));
    const double z2 = z * z;
    const double d = fma(x, z2 * z2 * z, -1.0);
    z = fma(z, d * fma(d, fma(d, -11.0 / 125.0, 3.0 / 25.0), -0.2), z); // (1+d)^(-1/5) = 1 - d/5 + 3/25 d^2 - 11/125 d^3
#endif
    return c.scale * z;
}
#endif

CLODE_DEV Controller make_controller(const SolverParams &sp, const realtype t_end)
{
    Controller c;
    c.reltol = sp.reltol;
    c.floor_ = sp.abstol / sp.reltol;
    c.scale = RCONST(0.8) * pow(sp.reltol, RCONST(1.0) / (ERR_ORDER + RCONST(1.0)));
#if defined(CLODE_SINGLE_PRECISION)
    c.floor_hi_min = t_end > ZERO ? (20 << 23) : 0x7f800000;
    c.floor_hi_max = 0;
#else
    c.floor_hi_min = t_end > ZERO ? (49 << 20) : 0x7ff00000;
#if !CLODE_EXACT_ARITH
    {   // 16 ulp(t) = 2^(E_t - 1071) <= 2^(E_dtmax - 1023) <= dtmax  <=>  E_t <= E_dtmax + 48
        const int d_hi = __double2hiint((double)sp.dtmax);
        const int top = (d_hi & 0x7ff00000) + (49 << 20);
        c.floor_hi_max = (d_hi > 0 && d_hi < 0x7ff00000) ? (top < 0x7ff00000 ? top : 0x7ff00000) : 0;
    }
#else
    c.floor_hi_max = 0;
#endif
#endif
    return c;
}

// hmin = 16 * | |nextafter(t, 1.1 t_end)| - t |   (adaptive_explicit_step.clh:17; "16 eps(t)")
// nextafter written out on the bit pattern: same result as the library call for every non-NaN
// input, without its NaN / signalling paths (about a third of the instructions).
CLODE_DEV realtype step_floor(const realtype t, const realtype t_end, const int fast_hi_min)
{
#if defined(CLODE_SINGLE_PRECISION)
#if CLODE_FAST_SINGLE
    {   // 0 < t <= t_end, biased exponent E >= 20: 16 ulp(t) = 2^(E-146), from the exponent field (as for double below)
        const int bits = __float_as_int(t);
        if (bits >= fast_hi_min && bits < 0x7f800000)
            return __int_as_float((bits & 0x7f800000) - (19 << 23));
    }
#endif
    return RCONST(16.0) * fabs(fabs(nextafter(t, RCONST(1.1) * t_end)) - t);
#else
#if !CLODE_EXACT_ARITH
    // Common case: 0 < t <= t_end (every caller steps only while t <= t_end), t normal and >= 2^-974.
    // nextafter then moves one ulp up, so the result is 16 ulp(t) = 2^(E-1071) exactly (E = biased exponent
    // of t): built from the exponent field, no FP64-pipe instruction.  tests/test_pm_math.py pins the identity.
    {
        const int hi = __double2hiint(t);
        if (hi >= fast_hi_min && hi < 0x7ff00000)
            return __hiloint2double((hi & 0x7ff00000) - (48 << 20), 0);
    }
#endif
    const double target = 1.1 * t_end;
    long long bits = __double_as_longlong(t);
    double next;
    if (t == target || target != target)
        next = (target != target) ? target : t; // equal: unchanged; NaN target propagates
    else if (t == 0.0)
        next = __longlong_as_double(target > 0.0 ? 1LL : (long long)0x8000000000000001ULL); // smallest subnormal toward target
    else {
        bits += ((t < target) == (t > 0.0)) ? 1LL : -1LL; // away from zero when moving outward, else toward it
        next = __longlong_as_double(bits);
    }
    return 16.0 * fabs(fabs(next) - t);
#endif
}

// The trial step an instance enters its first attempt with (the stored dt).  Production double keeps h <= dtmax as a
// loop invariant so that the clamp at the top of an attempt is one compare; a NaN dt becomes 0, which that clamp raises
// to hmin exactly as clamp(NaN, hmin, dtmax) does in the reference's fmin/fmax arithmetic.
CLODE_DEV realtype attempt_entry_step(const realtype dt, const SolverParams &sp)
{
#if CLODE_EXACT_ARITH || CLODE_FAST_SINGLE
    (void)sp;
    return dt;
#else
    return dt == dt ? min_nn(dt, sp.dtmax) : ZERO;
#endif
}

// One attempt of the step-size controller, adaptive_explicit_step.clh:9-81.
// `h` is the trial step carried between attempts, `clean` is the reference's
// noFailedSteps.  Returns true when the reference's stepper() would have returned
// (step accepted, or abandoned at hmin with its -1 flag); false = try again.
CLODE_DEV bool adaptive_attempt(Instance &I, realtype &h, bool &clean, const SolverParams &sp, const Controller &ctl,
                                 const realtype t_end)
{
    const realtype floor_ = ctl.floor_;
    realtype t1, xn[NV], kn[NV], err[NV];
    // hmin = 16 ulp(t), then h = clamp(h, hmin, dtmax) (adaptive_explicit_step.clh:17,31)
#if CLODE_EXACT_ARITH || defined(CLODE_SINGLE_PRECISION)
    const realtype hmin = step_floor(I.t, t_end, ctl.floor_hi_min);
    h = clamp_nn(h, hmin, sp.dtmax);
#else
    // Production double.  Common case (0 < t <= t_end, t normal, and 16 ulp(t) <= dtmax — folded into floor_hi_max): hmin
    // is a power of two built from the exponent field of t (step_floor), its low word is zero, so `h < hmin` is a
    // comparison of HIGH WORDS on the integer pipe; and h <= dtmax holds on entry (attempt_entry_step, the shrink of a
    // rejected attempt, the clamp that ends an accepted one), so the clamp's upper bound cannot bind.  Same value as the
    // two FP64 compares of the generic form, which every other case still takes.
    realtype hmin;
    {
        const int t_hi = __double2hiint(I.t);
        if (t_hi >= ctl.floor_hi_min && t_hi < ctl.floor_hi_max) {
            const int f_hi = (t_hi & 0x7ff00000) - (48 << 20);
            hmin = __hiloint2double(f_hi, 0);
            if (__double2hiint(h) < f_hi) h = hmin;
        } else {
            hmin = step_floor(I.t, t_end, 0x7ff00000);
            h = clamp_nn(h, hmin, sp.dtmax);
        }
    }
#endif

    h = trial_step(I, h, t1, xn, kn, err);

    realtype nerr = opaque(ZERO);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        // fmax(fmax(|x|, |xn|), floor) and norm_inf's fmax(|e|, running), NaN operands ignored as in the reference
        err[j] = div_norm(err[j], max_nn(abs_nn(I.x[j]), max_nn(abs_nn(xn[j]), floor_)));
#if !CLODE_EXACT_ARITH && !defined(CLODE_SINGLE_PRECISION)
        if (j == 0) {
            // fmax(|e|, 0) only drops a NaN: a test of the high word (|e| has no sign; every NaN an arithmetic instruction
            // produces is quiet, high word > 0x7ff00000) on the integer pipe instead of an FP64 compare
            const double e0 = abs_nn(err[0]);
            nerr = __double2hiint(e0) <= 0x7ff00000 ? e0 : 0.0;
            continue;
        }
#endif
        nerr = max_nn(abs_nn(err[j]), nerr);
    }
#if CLODE_NORM_SCALED_ONCE
    nerr *= h; // h = t1 - t > 0 (hmin > 0): max_j |h s_j| / scale_j = h max_j |s_j| / scale_j
#endif
    const bool reject = nerr > sp.reltol;
    realtype factor = RCONST(0.5);
    if (clean)
        factor = controller_factor(ctl, nerr);
    if (reject) {
        if (h <= hmin) { // cannot shrink further: stepper() returns -1, state untouched
            I.dt = hmin;
            h = hmin;
            clean = true;
            return true;
        }
        h *= clean ? max_nn(factor, opaque(MAX_SHRINK)) : RCONST(0.5);
        clean = false;
        return false;
    }
    if (clean)
        h *= min_nn(factor, opaque(RCONST(5.0)));
    h = min_nn(t_end - t1, h); // fmin(h, t_end - t1): h is never NaN here
    h = clamp_nn(h, hmin, sp.dtmax);
    I.dt = h;
    I.t = t1;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        I.x[j] = xn[j];
        I.k1[j] = kn[j];
    }
    clean = true;
    return true;
}

#endif // CLODE_ADAPTIVE

#endif // CLODE_STEPPERS_CUH
