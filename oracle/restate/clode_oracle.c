/* TEST INFRASTRUCTURE — not part of the product; never linked or imported by clode_b200.
 *
 * clode_oracle.c — CPU restatement of the reference's ensemble-ODE hot path in plain
 * C11: the four kernels, six selectable steppers, six observers and the per-instance
 * RNG.  It is written independently of both the reference text and the CUDA kernels
 * (run-time dimensions, one generic observer record, scalar loops) and is pinned
 * bit-for-bit against oracle/_ref — the reference's own sources compiled as host C —
 * by tests/test_oracle_pinning.py.  Every function cites the reference lines it follows
 * (paths relative to the reference root).
 *
 * Compile (oracle/restate.py does this): one shared object per right-hand side,
 *   gcc -std=gnu11 -O2 -ffp-contract=off -fopenmp -shared -fPIC \
 *       -DOR_RHS_FILE='"model.cl"' [-DOR_SINGLE] [-DOR_PM_MATH] clode_oracle.c
 * Dimensions, stepper and observer are run-time arguments.
 */
#include <float.h>
#include <stdbool.h>
#include <stdint.h>
#include <string.h>
#include <tgmath.h>
#undef I /* <complex.h>, pulled in by <tgmath.h>, defines the imaginary unit as a macro */

/* ---- precision (clode/cpp/realtype.cl:9-44) -------------------------------- */
#ifdef OR_SINGLE
typedef float realtype;
#define RCONST(x) (x##f)
#define BIG_REAL FLT_MAX
#else
typedef double realtype;
#define RCONST(x) (x)
#define BIG_REAL DBL_MAX
#endif
#define ZERO RCONST(0.0)
#define ONE RCONST(1.0)

/* ---- what a user RHS may rely on (clode/cpp/clODE_utilities.cl:17-19) ------- */
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#define heaviside(x) ((x) >= ZERO ? ONE : ZERO)
#define clamp(x, lo, hi) fmin(fmax((x), (lo)), (hi))

#ifdef OR_PM_MATH
#include "pm_math.h"
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

static inline realtype or_pown(realtype x, int n)
{
    int m = n < 0 ? -n : n;
    realtype r = ONE;
    for (int k = 0; k < m; ++k)
        r *= x;
    return n < 0 ? ONE / r : r;
}
#define pown(x, n) or_pown((x), (n))

/* the right-hand side: void getRHS(t, x_[], p_[], dx_[], aux_[], w_[])  (clode/cpp/steppers.cl:50) */
#include OR_RHS_FILE

/* ---- limits of this restatement ------------------------------------------- */
#define OR_MAXV 16  /* state variables   */
#define OR_MAXA 64  /* aux variables     */
#define OR_MAXW 8   /* Wiener variables  */
#define OR_MAXS 16  /* stored events     */

enum { ST_EULER, ST_HEUN, ST_RK4, ST_BS23, ST_DOPRI5, ST_SEULER };
enum { OB_BASIC, OB_BASICALL, OB_LOCALMAX, OB_NHOOD1, OB_NHOOD2, OB_THRESH2 };

typedef struct {
    int n_var, n_par, n_aux, n_wiener;
    int stepper, observer, n_store;
} or_problem;

/* clode/cpp/clODE_struct_defs.cl:11-20 */
typedef struct {
    realtype dt, dtmax, abstol, reltol;
    unsigned max_steps, max_store, nout;
} or_solver;

/* clode/cpp/observers.cl:25-46 */
typedef struct {
    unsigned e_var, f_var, max_events, max_stamps;
    realtype min_amp, min_imi, radius, x_up, x_down, dx_up, dx_down, eps_dx;
} or_obspar;

/* ======================= RNG (clode/cpp/clODE_random.cl) ===================== */
typedef struct {
    uint64_t s[2];
    bool have_spare;
    realtype spare;
} or_rng;

/* xoroshiro128+ with the (55, 14, 36) constants — clODE_random.cl:28-44 */
static uint64_t or_next(uint64_t s[2])
{
    uint64_t a = s[0], b = s[1], out = a + b;
    b ^= a;
    s[0] = ((a << 55) | (a >> 9)) ^ b ^ (b << 14);
    s[1] = (b << 36) | (b >> 28);
    return out;
}

/* clODE_random.cl:81-85: u64 -> real conversion, then * 2^-64 */
static realtype or_uniform(uint64_t s[2])
{
    uint64_t u = or_next(s);
    return (realtype)u * RCONST(5.421010862427522e-20);
}

/* Marsaglia polar method with the spare variate cached — clODE_random.cl:88-114 */
static realtype or_normal(or_rng *g)
{
    if (g->have_spare) {
        g->have_spare = false;
        return g->spare;
    }
    realtype a, b, q;
    do {
        a = RCONST(2.0) * or_uniform(g->s) - ONE;
        b = RCONST(2.0) * or_uniform(g->s) - ONE;
        q = a * a + b * b;
    } while (q >= ONE);
    q = sqrt((-RCONST(2.0) * log(q)) / q);
    g->spare = b * q;
    g->have_spare = true;
    return a * q;
}

/* ============================ per-instance state ============================= */
typedef struct {
    const or_problem *pb;
    const or_solver *sp;
    const realtype *tspan;
    realtype t, dt;
    realtype x[OR_MAXV], k1[OR_MAXV], p[OR_MAXV * 4], aux[OR_MAXA], w[OR_MAXW];
    or_rng rng;
} or_inst;

static void or_rhs(or_inst *I, realtype t, const realtype *x, realtype *dx)
{
    getRHS(t, x, I->p, dx, I->aux, I->w);
}

/* new Wiener increments: w_j = N(0,1)/sqrt(dt) — transient.cl:43-48, fixed_explicit_step.clh:29-33 */
static void or_draw_noise(or_inst *I, bool stochastic)
{
    for (int j = 0; j < I->pb->n_wiener; ++j)
        I->w[j] = stochastic ? or_normal(&I->rng) / sqrt(I->dt) : ZERO;
}

/* kernel prologue shared by all four kernels — transient.cl:28-52 */
static void or_load(or_inst *I, const or_problem *pb, const or_solver *sp, const realtype *tspan, int i,
                    int n_pts, const realtype *x0, const realtype *pars, const uint64_t *rng,
                    const realtype *dt)
{
    I->pb = pb;
    I->sp = sp;
    I->tspan = tspan;
    I->t = tspan[0];
    I->dt = dt[i];
    for (int j = 0; j < pb->n_par; ++j)
        I->p[j] = pars[j * n_pts + i];
    for (int j = 0; j < pb->n_var; ++j)
        I->x[j] = x0[j * n_pts + i];
    I->rng.s[0] = rng[i];
    I->rng.s[1] = rng[n_pts + i];
    I->rng.have_spare = false;
    for (int j = 0; j < OR_MAXW; ++j)
        I->w[j] = ZERO;
    or_draw_noise(I, pb->stepper == ST_SEULER);
    or_rhs(I, I->t, I->x, I->k1);
}

/* kernel epilogue — transient.cl:64-76 */
static void or_store(const or_inst *I, int i, int n_pts, realtype *xf, uint64_t *rng, realtype *dt,
                     realtype *tf)
{
    for (int j = 0; j < I->pb->n_var; ++j)
        xf[j * n_pts + i] = I->x[j];
    rng[i] = I->rng.s[0];
    rng[n_pts + i] = I->rng.s[1];
    dt[i] = I->dt;
    tf[i] = I->t;
}

/* ================================ steppers =================================== */

/* fixed-step bodies: steppers/fixed_explicit_Euler.clh:5-18, fixed_explicit_Trapezoidal.clh:5-23,
 * fixed_explicit_RK4.clh:5-40 — then the wrapper steppers/fixed_explicit_step.clh:9-37 */
static int or_step_fixed(or_inst *I)
{
    const int n = I->pb->n_var;
    const realtype h = I->dt;
    realtype *x = I->x, *k1 = I->k1;
    realtype y[OR_MAXV], k2[OR_MAXV], k3[OR_MAXV], k4[OR_MAXV];

    switch (I->pb->stepper) {
    case ST_EULER:
    case ST_SEULER:
        for (int j = 0; j < n; ++j)
            x[j] += h * k1[j];
        I->t += h;
        break;
    case ST_HEUN: {
        realtype t1 = I->t + h;
        for (int j = 0; j < n; ++j)
            y[j] = fma(h, k1[j], x[j]);
        or_rhs(I, t1, y, k2);
        for (int j = 0; j < n; ++j)
            x[j] += h * RCONST(0.5) * (k1[j] + k2[j]);
        I->t = t1;
        break;
    }
    case ST_RK4: {
        realtype hh = h * RCONST(0.5);
        realtype tm = I->t + hh, t1 = I->t + h;
        for (int j = 0; j < n; ++j)
            y[j] = x[j] + hh * k1[j];
        or_rhs(I, tm, y, k2);
        for (int j = 0; j < n; ++j)
            y[j] = x[j] + hh * k2[j];
        or_rhs(I, tm, y, k3);
        for (int j = 0; j < n; ++j)
            y[j] = x[j] + h * k3[j];
        or_rhs(I, t1, y, k4);
        for (int j = 0; j < n; ++j)
            x[j] += h * (k1[j] + RCONST(2.0) * k2[j] + RCONST(2.0) * k3[j] + k4[j]) / RCONST(6.0);
        I->t = t1;
        break;
    }
    }
    if (I->pb->stepper == ST_SEULER)
        or_draw_noise(I, true);
    or_rhs(I, I->t, x, k1);
    return 0;
}

/* Bogacki-Shampine 3(2) trial step — steppers/adaptive_bs23.clh:27-63.
 * Advances (*t, x, k1) in place, fills err[], returns the effective step. */
static realtype or_try_bs23(or_inst *I, realtype *t, realtype *x, realtype *k1, realtype h_in, realtype *err)
{
    const int n = I->pb->n_var;
    realtype t1 = *t + h_in;
    realtype h = t1 - *t;
    realtype y[OR_MAXV], k2[OR_MAXV], k3[OR_MAXV], k4[OR_MAXV];
    for (int j = 0; j < n; ++j)
        y[j] = x[j] + h * RCONST(0.5) * k1[j];
    or_rhs(I, *t + h * RCONST(0.5), y, k2);
    for (int j = 0; j < n; ++j)
        y[j] = x[j] + h * RCONST(0.75) * k2[j];
    or_rhs(I, *t + h * RCONST(0.75), y, k3);
    for (int j = 0; j < n; ++j)
        x[j] = x[j] + h * (RCONST(2.0) / RCONST(9.0) * k1[j] + RCONST(1.0) / RCONST(3.0) * k2[j] +
                           RCONST(4.0) / RCONST(9.0) * k3[j]);
    or_rhs(I, t1, x, k4);
    for (int j = 0; j < n; ++j) {
        err[j] = h * (RCONST(-5.0) / RCONST(72.0) * k1[j] + RCONST(1.0) / RCONST(12.0) * k2[j] +
                      RCONST(1.0) / RCONST(9.0) * k3[j] + RCONST(-1.0) / RCONST(8.0) * k4[j]);
        k1[j] = k4[j];
    }
    *t = t1;
    return h;
}

/* Dormand-Prince 5(4) trial step — steppers/adaptive_dp45.clh:58-110.
 * Stage sums run left to right over the non-zero tableau entries. */
static realtype or_try_dopri5(or_inst *I, realtype *t, realtype *x, realtype *k1, realtype h_in, realtype *err)
{
    static const realtype c[7] = {ZERO, RCONST(1.0) / RCONST(5.0), RCONST(3.0) / RCONST(10.0),
                                  RCONST(4.0) / RCONST(5.0), RCONST(8.0) / RCONST(9.0), ONE, ONE};
    static const realtype a[7][6] = {
        {0},
        {RCONST(1.0) / RCONST(5.0)},
        {RCONST(3.0) / RCONST(40.0), RCONST(9.0) / RCONST(40.0)},
        {RCONST(44.0) / RCONST(45.0), RCONST(-56.0) / RCONST(15.0), RCONST(32.0) / RCONST(9.0)},
        {RCONST(19372.0) / RCONST(6561.0), RCONST(-25360.0) / RCONST(2187.0), RCONST(64448.0) / RCONST(6561.0),
         RCONST(-212.0) / RCONST(729.0)},
        {RCONST(9017.0) / RCONST(3168.0), RCONST(-355.0) / RCONST(33.0), RCONST(46732.0) / RCONST(5247.0),
         RCONST(49.0) / RCONST(176.0), RCONST(-5103.0) / RCONST(18656.0)},
        {RCONST(35.0) / RCONST(384.0), ZERO, RCONST(500.0) / RCONST(1113.0), RCONST(125.0) / RCONST(192.0),
         RCONST(-2187.0) / RCONST(6784.0), RCONST(11.0) / RCONST(84.0)},
    };
    static const realtype e[7] = {RCONST(71.0) / RCONST(57600.0), ZERO, RCONST(-71.0) / RCONST(16695.0),
                                  RCONST(71.0) / RCONST(1920.0), RCONST(-17253.0) / RCONST(339200.0),
                                  RCONST(22.0) / RCONST(525.0), RCONST(-1.0) / RCONST(40.0)};
    const int n = I->pb->n_var;
    realtype t1 = *t + h_in;
    realtype h = t1 - *t;
    realtype k[7][OR_MAXV], y[OR_MAXV];
    for (int j = 0; j < n; ++j)
        k[0][j] = k1[j];
    for (int s = 1; s <= 5; ++s) { /* stages k2..k6 at t + c_s h */
        for (int j = 0; j < n; ++j) {
            realtype acc = a[s][0] * k[0][j];
            for (int q = 1; q < s; ++q)
                acc += a[s][q] * k[q][j];
            y[j] = x[j] + h * acc;
        }
        /* the reference writes `*ti + A_s * newDt` for s<6 and `*ti + newDt` for k6 */
        or_rhs(I, s < 5 ? *t + c[s] * h : *t + h, y, k[s]);
    }
    for (int j = 0; j < n; ++j) { /* 5th-order solution: row 6 without the zero b2 entry */
        realtype acc = a[6][0] * k[0][j];
        for (int q = 2; q < 6; ++q)
            acc += a[6][q] * k[q][j];
        x[j] = x[j] + h * acc;
    }
    or_rhs(I, t1, x, k[6]);
    for (int j = 0; j < n; ++j) {
        realtype acc = e[0] * k[0][j];
        for (int q = 2; q < 7; ++q)
            acc += e[q] * k[q][j];
        err[j] = h * acc;
        k1[j] = k[6][j];
    }
    *t = t1;
    return h;
}

/* step-size controller — steppers/adaptive_explicit_step.clh:9-81 */
static int or_step_adaptive(or_inst *I)
{
    const int n = I->pb->n_var;
    const bool dp = I->pb->stepper == ST_DOPRI5;
    const realtype order = dp ? RCONST(4.0) : RCONST(2.0);        /* LOCAL_ERROR_ORDER        */
    const realtype max_shrink = dp ? RCONST(0.1) : RCONST(0.5);    /* ADAPTIVE_STEP_MAX_SHRINK */
    const realtype max_grow = RCONST(5.0);                         /* ADAPTIVE_STEP_MAX_GROW   */
    const realtype expo = RCONST(1.0) / (order + RCONST(1.0));     /* EXPON                    */
    const realtype rtol = I->sp->reltol;
    const realtype floor_ = I->sp->abstol / rtol;
    const realtype hmin = RCONST(16.0) * fabs(fabs(nextafter(I->t, RCONST(1.1) * I->tspan[1])) - I->t);

    realtype h = I->dt, t_new, nerr;
    realtype xn[OR_MAXV], kn[OR_MAXV], err[OR_MAXV];
    bool clean = true;
    for (;;) {
        t_new = I->t;
        memcpy(xn, I->x, sizeof(realtype) * n);
        memcpy(kn, I->k1, sizeof(realtype) * n);
        h = clamp(h, hmin, I->sp->dtmax);
        h = dp ? or_try_dopri5(I, &t_new, xn, kn, h, err) : or_try_bs23(I, &t_new, xn, kn, h, err);
        nerr = ZERO;
        for (int j = 0; j < n; ++j) {
            err[j] /= fmax(fmax(fabs(I->x[j]), fabs(xn[j])), floor_);
            nerr = fmax(fabs(err[j]), nerr); /* norm_inf, clODE_utilities.cl:40-46 */
        }
        if (!(nerr > rtol))
            break;
        if (h <= hmin) {
            I->dt = hmin;
            return -1;
        }
        if (clean) {
            clean = false;
            h *= fmax(max_shrink, RCONST(0.8) * pow(rtol / nerr, expo));
        } else {
            h *= RCONST(0.5);
        }
    }
    if (clean)
        h *= fmin(max_grow, RCONST(0.8) * pow(rtol / nerr, expo));
    h = fmin(h, I->tspan[1] - t_new);
    h = clamp(h, hmin, I->sp->dtmax);
    I->dt = h;
    I->t = t_new;
    memcpy(I->x, xn, sizeof(realtype) * n);
    memcpy(I->k1, kn, sizeof(realtype) * n);
    return 0;
}

static int or_step(or_inst *I)
{
    int s = I->pb->stepper;
    return (s == ST_BS23 || s == ST_DOPRI5) ? or_step_adaptive(I) : or_step_fixed(I);
}

/* ================================ observers ================================== */

/* one record serves all six observers; unused members stay zero */
typedef struct {
    realtype tb[3], xb[OR_MAXV][3], dxb[OR_MAXV][3];
    realtype xmax[OR_MAXV], xmin[OR_MAXV], xmean[OR_MAXV], dxmax[OR_MAXV], dxmin[OR_MAXV];
    realtype xrange[OR_MAXV], center[OR_MAXV];
    realtype amax[OR_MAXA], amin[OR_MAXA], amean[OR_MAXA];
    realtype list_a[OR_MAXS], list_b[OR_MAXS], list_c[OR_MAXS], list_d[OR_MAXS];
    realtype tri[7][3]; /* (max, min, mean) accumulators */
    realtype t_start, t_last, t_event, t_down, t_max, t_min, x_min, down_mean;
    realtype g_xmax, g_xmin, g_dxmax, g_dxmin, x_up, x_down, dx_up, dx_down, x_thresh;
    realtype norm_now, norm_prev;
    unsigned peaks, steps, events, up, found, inside;
} or_obs;

enum { TRI_A, TRI_B, TRI_C, TRI_D, TRI_E, TRI_F, TRI_DT };

static void tri_reset(realtype tri[3])
{
    tri[0] = -BIG_REAL;
    tri[1] = BIG_REAL;
    tri[2] = ZERO;
}

/* runningMean — clODE_utilities.cl:172-177 */
static void mean_count(realtype *m, realtype v, unsigned cnt)
{
    if (cnt == 1)
        *m = v;
    else if (cnt > 1)
        *m += (v - *m) / (realtype)cnt;
}

/* runningMeanTime — clODE_utilities.cl:167-169 */
static realtype mean_time(realtype m, realtype v, realtype dt, realtype total)
{
    return m + (v - m) * dt / total;
}

static void tri_push(realtype tri[3], realtype v, unsigned cnt)
{
    tri[0] = fmax(v, tri[0]);
    tri[1] = fmin(v, tri[1]);
    mean_count(&tri[2], v, cnt);
}

/* first-occurrence arg-extrema over three samples — clODE_utilities.cl:49-60, 74-87, 90-101, 115-128 */
static int argmax3(const realtype v[3])
{
    realtype best = -BIG_REAL;
    int ix = 0;
    for (int k = 0; k < 3; ++k)
        if (v[k] > best) {
            best = v[k];
            ix = k;
        }
    return ix;
}
static int argmin3(const realtype v[3])
{
    realtype best = BIG_REAL;
    int ix = 0;
    for (int k = 0; k < 3; ++k)
        if (v[k] < best) {
            best = v[k];
            ix = k;
        }
    return ix;
}
/* maxOfArray/minOfArray return the running extreme, which stays +-BIG_REAL if no
 * element beats it (e.g. all NaN) */
static realtype max3(const realtype v[3], int *ix)
{
    realtype best = -BIG_REAL;
    *ix = 0;
    for (int k = 0; k < 3; ++k)
        if (v[k] > best) {
            best = v[k];
            *ix = k;
        }
    return best;
}
static realtype min3(const realtype v[3], int *ix)
{
    realtype best = BIG_REAL;
    *ix = 0;
    for (int k = 0; k < 3; ++k)
        if (v[k] < best) {
            best = v[k];
            *ix = k;
        }
    return best;
}

/* 3-deep history shift used by localmax/nhood1/nhood2/thresh2, e.g. observer_local_maximum.clh:218-230 */
static void ob_shift(or_obs *o, const or_inst *I)
{
    o->tb[0] = o->tb[1];
    o->tb[1] = o->tb[2];
    o->tb[2] = I->t;
    for (int j = 0; j < I->pb->n_var; ++j) {
        o->xb[j][0] = o->xb[j][1];
        o->xb[j][1] = o->xb[j][2];
        o->xb[j][2] = I->x[j];
        o->dxb[j][0] = o->dxb[j][1];
        o->dxb[j][1] = o->dxb[j][2];
        o->dxb[j][2] = I->k1[j];
    }
}

/* extent + time-weighted mean of every variable, slope and aux, e.g. observer_basic_allVar.clh:92-103 */
static void ob_extents(or_obs *o, const or_inst *I, realtype dt, realtype elapsed, bool count_mean)
{
    for (int j = 0; j < I->pb->n_var; ++j) {
        o->xmax[j] = fmax(I->x[j], o->xmax[j]);
        o->xmin[j] = fmin(I->x[j], o->xmin[j]);
        if (count_mean)
            mean_count(&o->xmean[j], I->x[j], o->steps); /* nhood1 only: observer_neighborhood_1.clh:253 */
        else
            o->xmean[j] = mean_time(o->xmean[j], I->x[j], dt, elapsed);
        o->dxmax[j] = fmax(I->k1[j], o->dxmax[j]);
        o->dxmin[j] = fmin(I->k1[j], o->dxmin[j]);
    }
    for (int j = 0; j < I->pb->n_aux; ++j) {
        o->amax[j] = fmax(I->aux[j], o->amax[j]);
        o->amin[j] = fmin(I->aux[j], o->amin[j]);
        if (count_mean)
            mean_count(&o->amean[j], I->aux[j], o->steps);
        else
            o->amean[j] = mean_time(o->amean[j], I->aux[j], dt, elapsed);
    }
}

/* initializeObserverData of each observer:
 * observer_basic.clh:33-42, observer_basic_allVar.clh:54-70, observer_local_maximum.clh:99-141,
 * observer_neighborhood_1.clh:93-157, observer_neighborhood_2.clh:89-135, observer_threshold_2.clh:120-193 */
static void ob_init(or_obs *o, const or_inst *I)
{
    const or_problem *pb = I->pb;
    memset(o, 0, sizeof *o);
    o->tb[2] = I->t;
    for (int j = 0; j < pb->n_var; ++j) {
        o->xb[j][2] = I->x[j];
        o->dxb[j][2] = I->k1[j];
        o->xmax[j] = o->dxmax[j] = -BIG_REAL;
        o->xmin[j] = o->dxmin[j] = BIG_REAL;
        if (pb->observer == OB_NHOOD2)
            o->center[j] = I->x[j];
    }
    for (int j = 0; j < pb->n_aux; ++j) {
        o->amax[j] = -BIG_REAL;
        o->amin[j] = BIG_REAL;
    }
    for (int k = 0; k < 7; ++k)
        tri_reset(o->tri[k]);
    o->t_start = o->t_last = I->t;
    o->x_min = BIG_REAL;
    o->g_xmax = o->g_dxmax = -BIG_REAL;
    o->g_xmin = o->g_dxmin = BIG_REAL;
}

/* warmupObserverData (two-pass observers): observer_threshold_2.clh:196-203, observer_neighborhood_2.clh:139-147 */
static void ob_warmup(or_obs *o, const or_inst *I, const or_obspar *op)
{
    if (I->pb->observer == OB_THRESH2) {
        o->g_xmax = fmax(o->g_xmax, I->x[op->e_var]);
        o->g_xmin = fmin(o->g_xmin, I->x[op->e_var]);
        o->g_dxmax = fmax(o->g_dxmax, I->k1[op->e_var]);
        o->g_dxmin = fmin(o->g_dxmin, I->k1[op->e_var]);
    } else if (I->pb->observer == OB_NHOOD2) {
        for (int j = 0; j < I->pb->n_var; ++j) {
            o->xmax[j] = fmax(I->x[j], o->xmax[j]);
            o->xmin[j] = fmin(I->x[j], o->xmin[j]);
        }
    }
}

/* initializeEventDetector: observer_threshold_2.clh:206-224, observer_neighborhood_2.clh:150-154 */
static void ob_arm(or_obs *o, const or_inst *I, const or_obspar *op)
{
    if (I->pb->observer == OB_THRESH2) {
        realtype amp = o->g_xmax - o->g_xmin;
        o->x_up = o->g_xmin + op->x_up * amp;
        o->x_down = op->x_down > ZERO ? o->g_xmin + op->x_down * amp : o->x_up;
        o->dx_up = op->dx_up * o->g_dxmax;
        o->dx_down = op->dx_down > ZERO ? op->dx_down * o->g_dxmin : o->g_dxmin;
        o->up = I->x[op->e_var] > o->x_up ? 1 : 0;
    } else if (I->pb->observer == OB_NHOOD2) {
        for (int j = 0; j < I->pb->n_var; ++j)
            o->xrange[j] = o->xmax[j] - o->xmin[j];
        o->x_thresh = o->xmin[op->e_var] + op->x_down * o->xrange[op->e_var];
    }
}

/* updateObserverData: observer_basic.clh:62-73, observer_basic_allVar.clh:87-104,
 * observer_local_maximum.clh:213-283, observer_neighborhood_1.clh:229-307,
 * observer_neighborhood_2.clh:210-267, observer_threshold_2.clh:308-407 */
static void ob_update(or_obs *o, const or_inst *I, const or_obspar *op)
{
    const or_problem *pb = I->pb;
    const unsigned f = op->f_var, e = op->e_var;
    ++o->steps;
    if (pb->observer == OB_BASIC || pb->observer == OB_BASICALL) {
        realtype dt = I->t - o->t_last;
        o->t_last = I->t;
        realtype elapsed = I->t - o->t_start;
        if (pb->observer == OB_BASICALL) {
            ob_extents(o, I, dt, elapsed, false);
        } else {
            o->xmax[0] = fmax(I->x[f], o->xmax[0]);
            o->xmin[0] = fmin(I->x[f], o->xmin[0]);
            o->xmean[0] = mean_time(o->xmean[0], I->x[f], dt, elapsed);
            o->dxmax[0] = fmax(I->k1[f], o->dxmax[0]);
            o->dxmin[0] = fmin(I->k1[f], o->dxmin[0]);
        }
        return;
    }
    ob_shift(o, I);
    realtype dt = o->tb[2] - o->tb[1];
    realtype elapsed = I->t - o->t_start;
    if (pb->observer != OB_LOCALMAX)
        tri_push(o->tri[TRI_DT], dt, o->steps);
    ob_extents(o, I, dt, elapsed, pb->observer == OB_NHOOD1);
    if (o->steps < 2)
        return;

    const realtype d1 = o->dxb[f][1], d2 = o->dxb[f][2];
    int ix;
    switch (pb->observer) {
    case OB_LOCALMAX:
        if (d1 < 0.0 && d2 > 0.0) { /* local minimum of the feature variable */
            ix = argmin3(o->xb[f]);
            o->t_min = o->tb[ix];
            o->x_min = o->xb[f][ix];
            /* reference indexes [events-1] unguarded; guarded here as in oracle/_ref */
            if (o->events > 0 && o->events <= (unsigned)pb->n_store) {
                o->list_c[o->events - 1] = o->t_min;
                o->list_d[o->events - 1] = o->x_min;
            }
        }
        break;
    case OB_NHOOD1:
        if (!o->found) {
            if (o->dxb[e][1] <= 0.0 && o->dxb[e][2] > 0.0) { /* first local min of the event variable */
                (void)min3(o->xb[e], &ix);
                o->t_event = o->tb[ix];
                o->found = 1;
                for (int j = 0; j < pb->n_var; ++j)
                    o->center[j] = o->xb[j][ix];
            }
        } else if (d1 >= 0.0 && d2 < 0.0) {
            o->peaks++;
        }
        break;
    case OB_NHOOD2:
        if (o->found) {
            if (d1 >= 0.0 && d2 < 0.0)
                o->peaks++;
        } else if (o->xb[e][1] > o->x_thresh && o->xb[e][2] < o->x_thresh) {
            o->found = 1;
            o->inside = 1;
            for (int j = 0; j < pb->n_var; ++j)
                o->center[j] = I->x[j];
        }
        break;
    case OB_THRESH2:
        if (d1 > 0.0 && d2 < 0.0) {
            (void)max3(o->xb[f], &ix);
            o->peaks++;
            o->t_max = o->tb[ix];
        }
        if (d1 < 0.0 && d2 > 0.0) {
            o->x_min = min3(o->xb[f], &ix);
            o->t_min = o->tb[ix];
        }
        if (o->up) {
            if (I->x[e] <= o->x_down && I->k1[e] >= o->dx_down) {
                o->t_down = I->t;
                o->up = 0;
                if (o->events > 0 && o->events <= (unsigned)pb->n_store)
                    o->list_b[o->events - 1] = o->t_down;
                o->down_mean = I->x[f];
            }
        } else {
            realtype since = I->t - o->t_down;
            if (since > 0.0)
                o->down_mean = mean_time(o->down_mean, I->x[f], dt, since);
        }
        break;
    }
}

/* eventFunction: observer_local_maximum.clh:144-152, observer_neighborhood_1.clh:169-189,
 * observer_neighborhood_2.clh:157-175, observer_threshold_2.clh:227-240 */
static bool ob_event(or_obs *o, const or_inst *I, const or_obspar *op)
{
    const or_problem *pb = I->pb;
    const unsigned f = op->f_var, e = op->e_var;
    realtype d[OR_MAXV], acc;
    unsigned was;
    switch (pb->observer) {
    case OB_LOCALMAX:
        return o->steps >= 2 && o->dxb[f][1] > 0.0 && o->dxb[f][2] < 0.0;
    case OB_NHOOD1:
        if (o->steps < 2 || !o->found)
            return false;
        if (o->xmax[f] - o->xmin[f] < op->min_amp)
            return false;
        was = o->inside;
        o->norm_prev = o->norm_now;
        for (int j = 0; j < pb->n_var; ++j)
            d[j] = fabs(I->x[j] - o->center[j]) / (o->xmax[j] - o->xmin[j]);
        acc = ZERO;
        for (int j = 0; j < pb->n_var; ++j)
            acc += d[j] * d[j];
        o->norm_now = sqrt(acc);
        o->inside = o->norm_now <= op->radius;
        return o->inside & !was;
    case OB_NHOOD2:
        if (o->steps < 2 || !o->found)
            return false;
        for (int j = 0; j < pb->n_var; ++j)
            d[j] = (I->x[j] - o->center[j]) / o->xrange[j];
        acc = ZERO;
        for (int j = 0; j < pb->n_var; ++j)
            acc += d[j] * d[j];
        was = o->inside;
        o->inside = sqrt(acc) < op->radius;
        return was && !o->inside;
    case OB_THRESH2:
        if (o->steps < 2)
            return false;
        if (o->g_xmax - o->g_xmin < op->min_amp)
            return false;
        if (o->up)
            return false;
        return I->x[e] > o->x_up && I->k1[e] > o->dx_up;
    default:
        return false;
    }
}

/* computeEventFeatures (returns true on a terminal event): observer_local_maximum.clh:156-205,
 * observer_neighborhood_1.clh:193-221, observer_neighborhood_2.clh:178-207, observer_threshold_2.clh:243-305 */
static bool ob_on_event(or_obs *o, const or_inst *I, const or_obspar *op)
{
    const or_problem *pb = I->pb;
    const unsigned f = op->f_var;
    ++o->events;
    switch (pb->observer) {
    case OB_LOCALMAX: {
        int ix = argmax3(o->xb[f]);
        realtype t_pk = o->tb[ix], x_pk = o->xb[f][ix];
        if (o->events > 1) {
            tri_push(o->tri[TRI_A], t_pk - o->t_max, o->events - 1); /* inter-maximum interval */
            tri_push(o->tri[TRI_B], x_pk - o->x_min, o->events - 1); /* amplitude */
        }
        o->t_max = t_pk;
        if (o->events <= (unsigned)pb->n_store) {
            o->list_a[o->events - 1] = t_pk;
            o->list_b[o->events - 1] = x_pk;
        }
        return o->events == op->max_events;
    }
    case OB_NHOOD1:
    case OB_NHOOD2: {
        realtype now = I->t;
        if (o->events > 1) {
            tri_push(o->tri[TRI_A], (realtype)o->peaks, o->events - 1);
            tri_push(o->tri[TRI_B], now - o->t_event, o->events - 1);
        }
        o->t_event = now;
        o->peaks = 0;
        if (pb->observer == OB_NHOOD1)
            return o->events >= op->max_events;
        if (o->events <= (unsigned)pb->n_store)
            o->list_a[o->events - 1] = now;
        return o->events == op->max_events;
    }
    case OB_THRESH2: {
        realtype now = I->t;
        o->up = 1;
        if (o->events > 1) {
            unsigned np = o->events - 1;
            tri_push(o->tri[TRI_A], (realtype)o->peaks, np);
            realtype period = now - o->t_event;
            tri_push(o->tri[TRI_B], period, np);
            realtype up_for = o->t_down - o->t_event;
            tri_push(o->tri[TRI_C], up_for, np);
            tri_push(o->tri[TRI_D], now - o->t_down, np);
            tri_push(o->tri[TRI_E], up_for / period, np);
            tri_push(o->tri[TRI_F], o->down_mean - o->x_min, np);
        }
        if (o->events <= (unsigned)pb->n_store)
            o->list_a[o->events - 1] = now;
        o->t_event = now;
        o->peaks = 0;
        return o->events == op->max_events;
    }
    default:
        return false;
    }
}

/* finalizeFeatures: observer_basic.clh:76-84, observer_basic_allVar.clh:107-124,
 * observer_local_maximum.clh:288-316, observer_neighborhood_1.clh:310-348,
 * observer_neighborhood_2.clh:270-302, observer_threshold_2.clh:411-456 */
static void ob_emit(const or_obs *o, const or_inst *I, realtype *F, int i, int n_pts)
{
    const or_problem *pb = I->pb;
    int c = 0;
#define PUT(v) F[(c++) * n_pts + i] = (v)
    const bool multi = o->events > 1;
    switch (pb->observer) {
    case OB_BASIC:
        PUT(o->xmax[0]); PUT(o->xmin[0]); PUT(o->xmean[0]); PUT(o->dxmax[0]); PUT(o->dxmin[0]);
        PUT(o->steps);
        return;
    case OB_LOCALMAX:
        for (int k = 0; k < 2; ++k)
            for (int q = 0; q < 3; ++q)
                PUT(multi ? o->tri[k][q] : ZERO);
        break;
    case OB_NHOOD1:
    case OB_NHOOD2:
        for (int q = 0; q < 3; ++q) PUT(multi ? o->tri[TRI_B][q] : ZERO); /* period */
        for (int q = 0; q < 3; ++q) PUT(multi ? o->tri[TRI_A][q] : ZERO); /* peaks  */
        break;
    case OB_THRESH2:
        for (int q = 0; q < 3; ++q) PUT(multi ? o->tri[TRI_B][q] : ZERO); /* period */
        for (int q = 0; q < 3; ++q) PUT(multi ? o->tri[TRI_A][q] : ZERO); /* peaks  */
        for (int k = TRI_C; k <= TRI_F; ++k)
            for (int q = 0; q < 3; ++q)
                PUT(multi ? o->tri[k][q] : ZERO);
        break;
    default:
        break;
    }
    for (int j = 0; j < pb->n_var; ++j) {
        PUT(o->xmax[j]); PUT(o->xmin[j]); PUT(o->xmean[j]);
        if (pb->observer == OB_NHOOD2) {
            PUT(o->xrange[j]); PUT(o->center[j]);
        }
        PUT(o->dxmax[j]); PUT(o->dxmin[j]);
    }
    for (int j = 0; j < pb->n_aux; ++j) {
        PUT(o->amax[j]); PUT(o->amin[j]); PUT(o->amean[j]);
    }
    for (int j = 0; j < pb->n_store; ++j) {
        if (pb->observer == OB_LOCALMAX) {
            PUT(o->list_a[j]); PUT(o->list_b[j]); PUT(o->list_c[j]); PUT(o->list_d[j]);
        } else if (pb->observer == OB_THRESH2) {
            PUT(o->list_a[j]); PUT(o->list_b[j]);
        } else if (pb->observer == OB_NHOOD2) {
            PUT(o->list_a[j]);
        }
    }
    if (pb->observer == OB_NHOOD1)
        PUT(o->events - 1); /* unsigned wrap when no event, as in the reference (:343) */
    else if (pb->observer != OB_BASICALL)
        PUT(o->events);
    PUT(o->steps);
    if (pb->observer == OB_NHOOD1 || pb->observer == OB_NHOOD2 || pb->observer == OB_THRESH2)
        for (int q = 0; q < 3; ++q)
            PUT(o->tri[TRI_DT][q]);
#undef PUT
}

/* finalizeObserverData (time shift for continuation): observer_basic.clh:87-90,
 * observer_local_maximum.clh:319-328, observer_neighborhood_1.clh:351-360,
 * observer_neighborhood_2.clh:305-312, observer_threshold_2.clh:459-470 */
static void ob_rebase(or_obs *o, const or_inst *I)
{
    realtype T = I->t - I->tspan[0];
    o->t_start -= T;
    switch (I->pb->observer) {
    case OB_LOCALMAX:
        o->t_max = o->t_max - T;
        o->t_min = o->t_min - T;
        break;
    case OB_NHOOD1:
    case OB_NHOOD2:
        o->t_event -= T;
        break;
    case OB_THRESH2:
        o->t_event -= T;
        o->t_down -= T;
        o->t_max -= T;
        o->t_min -= T;
        break;
    default:
        return;
    }
    for (int k = 0; k < 3; ++k)
        o->tb[k] = o->tb[k] - T;
}

/* ================================= kernels =================================== */

#define OR_PARALLEL_FOR _Pragma("omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)")

long or_obs_size(void) { return (long)sizeof(or_obs); }
long or_real_size(void) { return (long)sizeof(realtype); }

/* clode/cpp/transient.cl:9-77 */
void or_transient(const or_problem *pb, int n_pts, int nthreads, const realtype *tspan, const realtype *x0,
                  const realtype *pars, const or_solver *sp, realtype *xf, uint64_t *rng, realtype *dt,
                  realtype *tf)
{
    OR_PARALLEL_FOR
    for (int i = 0; i < n_pts; ++i) {
        or_inst I;
        or_load(&I, pb, sp, tspan, i, n_pts, x0, pars, rng, dt);
        unsigned step = 0;
        while (I.t <= tspan[1] && step < sp->max_steps) {
            ++step;
            or_step(&I);
        }
        or_store(&I, i, n_pts, xf, rng, dt, tf);
    }
}

/* clode/cpp/initializeObserver.cl:9-83 */
void or_initialize_observer(const or_problem *pb, int n_pts, int nthreads, const realtype *tspan,
                            const realtype *x0, const realtype *pars, const or_solver *sp,
                            const uint64_t *rng, const realtype *dt, or_obs *odata, const or_obspar *op)
{
    const bool two_pass = pb->observer == OB_THRESH2 || pb->observer == OB_NHOOD2;
    OR_PARALLEL_FOR
    for (int i = 0; i < n_pts; ++i) {
        or_inst I;
        or_obs *o = &odata[i];
        or_load(&I, pb, sp, tspan, i, n_pts, x0, pars, rng, dt);
        ob_init(o, &I);
        if (two_pass) {
            unsigned step = 0;
            while (I.t < tspan[1] && step < sp->max_steps) { /* strict <, unlike the other kernels */
                ++step;
                or_step(&I);
                ob_warmup(o, &I, op);
            }
            I.t = tspan[0];
            for (int j = 0; j < pb->n_var; ++j)
                I.x[j] = x0[j * n_pts + i];
            or_rhs(&I, I.t, I.x, I.k1);
        }
        ob_arm(o, &I, op);
        /* dt and RNG state are deliberately not written back (initializeObserver.cl:79-82) */
    }
}

/* clode/cpp/features.cl:10-106 */
void or_features(const or_problem *pb, int n_pts, int nthreads, const realtype *tspan, const realtype *x0,
                 const realtype *pars, const or_solver *sp, realtype *xf, uint64_t *rng, realtype *dt,
                 realtype *tf, or_obs *odata, const or_obspar *op, realtype *F)
{
    OR_PARALLEL_FOR
    for (int i = 0; i < n_pts; ++i) {
        or_inst I;
        or_obs o = odata[i];
        or_load(&I, pb, sp, tspan, i, n_pts, x0, pars, rng, dt);
        unsigned step = 0;
        while (I.t <= tspan[1] && step < sp->max_steps) {
            ++step;
            or_step(&I);
            ob_update(&o, &I, op);
            if (ob_event(&o, &I, op) && ob_on_event(&o, &I, op))
                break;
        }
        ob_emit(&o, &I, F, i, n_pts);
        ob_rebase(&o, &I);
        odata[i] = o;
        or_store(&I, i, n_pts, xf, rng, dt, tf);
    }
}

/* clode/cpp/trajectory.cl:14-112. The output arrays must hold max_store+1 rows:
 * row index max_store can be written. */
void or_trajectory(const or_problem *pb, int n_pts, int nthreads, const realtype *tspan, const realtype *x0,
                   const realtype *pars, const or_solver *sp, realtype *xf, uint64_t *rng, realtype *dt,
                   realtype *tf, realtype *t, realtype *x, realtype *dx, realtype *aux, int *n_stored)
{
    const int nv = pb->n_var, na = pb->n_aux;
    OR_PARALLEL_FOR
    for (int i = 0; i < n_pts; ++i) {
        or_inst I;
        or_load(&I, pb, sp, tspan, i, n_pts, x0, pars, rng, dt);
        int row = 0;
        unsigned step = 0;
        for (;;) {
            t[(size_t)row * n_pts + i] = I.t;
            for (int j = 0; j < nv; ++j) {
                x[(size_t)row * n_pts * nv + (size_t)j * n_pts + i] = I.x[j];
                dx[(size_t)row * n_pts * nv + (size_t)j * n_pts + i] = I.k1[j];
            }
            for (int j = 0; j < na; ++j)
                aux[(size_t)row * n_pts * na + (size_t)j * n_pts + i] = I.aux[j];
            /* advance until the next stored step or the end of the run */
            bool more = false;
            while (I.t <= tspan[1] && step < sp->max_steps && (unsigned)row < sp->max_store) {
                ++step;
                or_step(&I);
                if (step % sp->nout == 0) {
                    ++row;
                    more = true;
                    break;
                }
            }
            if (!more)
                break;
        }
        n_stored[i] = row;
        or_store(&I, i, n_pts, xf, rng, dt, tf);
    }
}
