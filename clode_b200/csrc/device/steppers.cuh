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

// struct SolverParams — clode/cpp/clODE_struct_defs.cl:11-20 (passed by value as a kernel argument)
struct SolverParams {
    realtype dt, dtmax, abstol, reltol;
    unsigned int max_steps, max_store, nout;
};

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

CLODE_DEV void draw_noise(Instance &I)
{
#pragma unroll
    for (int j = 0; j < N_WIENER; ++j)
#if CLODE_STOCHASTIC
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

#if defined(EXPLICIT_BS23)
#define ERR_ORDER RCONST(2.0)
#define MAX_SHRINK RCONST(0.5)
// Bogacki-Shampine 3(2), adaptive_bs23.clh:27-63.  Returns the effective step.
CLODE_DEV realtype trial_step(Instance &I, const realtype h_in, realtype &t1, realtype xn[NV],
                              realtype kn[NV], realtype err[NV])
{
    t1 = I.t + h_in;
    const realtype h = t1 - I.t;
    realtype y[NV], k2[NV], k3[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * RCONST(0.5) * I.k1[j];
    getRHS(I.t + h * RCONST(0.5), y, I.p, k2, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * RCONST(0.75) * k2[j];
    getRHS(I.t + h * RCONST(0.75), y, I.p, k3, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        xn[j] = I.x[j] + h * (RCONST(2.0) / RCONST(9.0) * I.k1[j] + RCONST(1.0) / RCONST(3.0) * k2[j] +
                              RCONST(4.0) / RCONST(9.0) * k3[j]);
    getRHS(t1, xn, I.p, kn, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        err[j] = h * (RCONST(-5.0) / RCONST(72.0) * I.k1[j] + RCONST(1.0) / RCONST(12.0) * k2[j] +
                      RCONST(1.0) / RCONST(9.0) * k3[j] + RCONST(-1.0) / RCONST(8.0) * kn[j]);
    return h;
}
#else // EXPLICIT_DOPRI5
#define ERR_ORDER RCONST(4.0)
#define MAX_SHRINK RCONST(0.1)
// Dormand-Prince 5(4), adaptive_dp45.clh:58-110.  Returns the effective step.
CLODE_DEV realtype trial_step(Instance &I, const realtype h_in, realtype &t1, realtype xn[NV],
                              realtype kn[NV], realtype err[NV])
{
    t1 = I.t + h_in;
    const realtype h = t1 - I.t;
    realtype y[NV], k2[NV], k3[NV], k4[NV], k5[NV], k6[NV];
    const realtype *k1 = I.k1;
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (RCONST(1.0) / RCONST(5.0) * k1[j]);
    getRHS(I.t + RCONST(1.0) / RCONST(5.0) * h, y, I.p, k2, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (RCONST(3.0) / RCONST(40.0) * k1[j] + RCONST(9.0) / RCONST(40.0) * k2[j]);
    getRHS(I.t + RCONST(3.0) / RCONST(10.0) * h, y, I.p, k3, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (RCONST(44.0) / RCONST(45.0) * k1[j] + RCONST(-56.0) / RCONST(15.0) * k2[j] +
                             RCONST(32.0) / RCONST(9.0) * k3[j]);
    getRHS(I.t + RCONST(4.0) / RCONST(5.0) * h, y, I.p, k4, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (RCONST(19372.0) / RCONST(6561.0) * k1[j] + RCONST(-25360.0) / RCONST(2187.0) * k2[j] +
                             RCONST(64448.0) / RCONST(6561.0) * k3[j] + RCONST(-212.0) / RCONST(729.0) * k4[j]);
    getRHS(I.t + RCONST(8.0) / RCONST(9.0) * h, y, I.p, k5, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        y[j] = I.x[j] + h * (RCONST(9017.0) / RCONST(3168.0) * k1[j] + RCONST(-355.0) / RCONST(33.0) * k2[j] +
                             RCONST(46732.0) / RCONST(5247.0) * k3[j] + RCONST(49.0) / RCONST(176.0) * k4[j] +
                             RCONST(-5103.0) / RCONST(18656.0) * k5[j]);
    getRHS(I.t + h, y, I.p, k6, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        xn[j] = I.x[j] + h * (RCONST(35.0) / RCONST(384.0) * k1[j] + RCONST(500.0) / RCONST(1113.0) * k3[j] +
                              RCONST(125.0) / RCONST(192.0) * k4[j] + RCONST(-2187.0) / RCONST(6784.0) * k5[j] +
                              RCONST(11.0) / RCONST(84.0) * k6[j]);
    getRHS(t1, xn, I.p, kn, I.aux, I.w);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        err[j] = h * (RCONST(71.0) / RCONST(57600.0) * k1[j] + RCONST(-71.0) / RCONST(16695.0) * k3[j] +
                      RCONST(71.0) / RCONST(1920.0) * k4[j] + RCONST(-17253.0) / RCONST(339200.0) * k5[j] +
                      RCONST(22.0) / RCONST(525.0) * k6[j] + RCONST(-1.0) / RCONST(40.0) * kn[j]);
    return h;
}
#endif

// 0.8 * (reltol/err)^(1/(order+1)) for the step-size controller (adaptive_explicit_step.clh:51,66).
// Bit-exact tier, reference-math builds and single precision evaluate it as written, with pow().
// Production double: libdevice's pow() costs ~100 FP64-pipe instructions on a ~40-deep dependent
// chain, plus an IEEE division for its argument — a third of a Lorenz dopri5 step — for a factor
// that is immediately clamped to [MAX_SHRINK, 5].  There the factor is computed division-free as
//     0.8 * reltol^(1/(p+1)) * err^(-1/(p+1)),
// the first two terms once per kernel, the last by Newton's iteration on z^-(p+1) = err seeded from
// the SFU (FP32 lg2/ex2): ~14 FP64 instructions, <= 4 ulp, well inside OpenCL C's 16-ulp bound
// for pow.  err outside [1e-30, 1e30] saturates, which the clamps make indistinguishable.
struct Controller {
    realtype reltol, floor_;
    realtype scale; // 0.8 * reltol^(1/(p+1))   (production double only)
};
#if defined(CLODE_BITEXACT) || defined(CLODE_SINGLE_PRECISION) || defined(CLODE_REFERENCE_MATH)
#define CLODE_EXACT_CONTROLLER 1
CLODE_DEV realtype controller_factor(const Controller &c, realtype nerr)
{
    return RCONST(0.8) * pow(c.reltol / nerr, RCONST(1.0) / (ERR_ORDER + RCONST(1.0)));
}
#else
#define CLODE_EXACT_CONTROLLER 0
CLODE_DEV double controller_factor(const Controller &c, double nerr)
{
    const double x = fmin(fmax(nerr, 1e-30), 1e30);
#if defined(EXPLICIT_BS23)
    return c.scale * rcbrt(x);
#else
    double z = (double)exp2f(-0.2f * __log2f((float)x)); // ~ x^(-1/5), relative error ~1e-6
    const double fifth_x = -0.2 * x;
    // Newton on f(z) = z^-5 - x :  z <- z (1.2 - 0.2 x z^5), twice: 1e-6 -> 1e-11 -> rounding level
    double z2 = z * z;
    z = z * fma(fifth_x, z2 * z2 * z, 1.2);
    z2 = z * z;
    z = z * fma(fifth_x, z2 * z2 * z, 1.2);
    return c.scale * z;
#endif
}
#endif

CLODE_DEV Controller make_controller(const SolverParams &sp)
{
    Controller c;
    c.reltol = sp.reltol;
    c.floor_ = sp.abstol / sp.reltol;
    c.scale = RCONST(0.8) * pow(sp.reltol, RCONST(1.0) / (ERR_ORDER + RCONST(1.0)));
    return c;
}

// hmin = 16 * | |nextafter(t, 1.1 t_end)| - t |   (adaptive_explicit_step.clh:17; "16 eps(t)")
// nextafter written out on the bit pattern: same result as the library call for every non-NaN
// input, without its NaN / signalling paths (about a third of the instructions).
CLODE_DEV realtype step_floor(const realtype t, const realtype t_end)
{
#if defined(CLODE_SINGLE_PRECISION)
    return RCONST(16.0) * fabs(fabs(nextafter(t, RCONST(1.1) * t_end)) - t);
#else
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

// One attempt of the step-size controller, adaptive_explicit_step.clh:9-81.
// `h` is the trial step carried between attempts, `clean` is the reference's
// noFailedSteps.  Returns true when the reference's stepper() would have returned
// (step accepted, or abandoned at hmin with its -1 flag); false = try again.
CLODE_DEV bool adaptive_attempt(Instance &I, realtype &h, bool &clean, const SolverParams &sp, const Controller &ctl,
                                 const realtype t_end)
{
    const realtype floor_ = ctl.floor_;
    const realtype hmin = step_floor(I.t, t_end);
    realtype t1, xn[NV], kn[NV], err[NV];

    h = clamp(h, hmin, sp.dtmax);
    h = trial_step(I, h, t1, xn, kn, err);

    realtype nerr = ZERO;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        err[j] /= fmax(fmax(fabs(I.x[j]), fabs(xn[j])), floor_);
        nerr = fmax(fabs(err[j]), nerr);
    }
    const bool reject = nerr > sp.reltol;
    if (reject && h <= hmin) { // cannot shrink further: stepper() returns -1, state untouched
        I.dt = hmin;
        h = hmin;
        clean = true;
        return true;
    }
    realtype factor = RCONST(0.5);
    if (clean)
        factor = controller_factor(ctl, nerr);
    if (reject) {
        h *= clean ? fmax(MAX_SHRINK, factor) : RCONST(0.5);
        clean = false;
        return false;
    }
    if (clean)
        h *= fmin(RCONST(5.0), factor);
    h = fmin(h, t_end - t1);
    h = clamp(h, hmin, sp.dtmax);
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
