// observers.cuh — on-the-fly feature observers fused into the step loop.
//
// Semantics follow clode/cpp/observers/observer_*.clh function by function (init /
// warm-up / per-step update / event test / event features / feature read-out / time
// re-base for continuation); the data layout does not:
//   * every observer keeps ONLY the fields that can influence a feature (e.g. the
//     3-deep x/dx history is kept for the feature/event variable alone, where the
//     reference shifts 6*N_VAR values per step and reads 5 of them), so the state
//     lives in registers;
//   * the feature and event variable indices are compile-time constants
//     (F_VAR_IX / E_VAR_IX), so `x[fVarIx]` never forces a register array into
//     local memory;
//   * between kernels the state is spilled to a structure-of-arrays record
//     (field-major, one coalesced row per field) instead of the reference's
//     array-of-structs `ObserverData` (clode/cpp/features.cl:45,91).
//
// Observer selection by macro, as in the reference (`oi.define`):
//   USE_OBSERVER_BASIC, USE_OBSERVER_BASIC_ALLVAR, USE_OBSERVER_LOCAL_MAX,
//   USE_OBSERVER_NEIGHBORHOOD_1, USE_OBSERVER_NEIGHBORHOOD_2, USE_OBSERVER_THRESHOLD_2
#ifndef CLODE_OBSERVERS_CUH
#define CLODE_OBSERVERS_CUH

#ifndef N_STORE_EVENTS
#define N_STORE_EVENTS 0
#endif
#define NS_ (N_STORE_EVENTS > 0 ? N_STORE_EVENTS : 1)

// struct ObserverParams — clode/cpp/observers.cl:25-46 (rebuilt per thread from the constant argument block).
// eVarIx / fVarIx are carried for completeness; the kernels use E_VAR_IX / F_VAR_IX.
struct ObserverParams {
    unsigned int eVarIx, fVarIx, maxEventCount, maxEventTimestamps;
    realtype minXamp, minIMI, nHoodRadius, xUpThresh, xDownThresh, dxUpThresh, dxDownThresh, eps_dx;
};

// ---- small building blocks -----------------------------------------------------------

// engine-internal forms of runningMeanTime / runningMean (clODE_utilities.cl:167-177).
// Reference-arithmetic builds evaluate mean + (v - mean) * dt / span exactly as written, every step.
//
// Production build: the time-weighted means are carried as INTEGRALS while a kernel runs.  With
// span_n = span_{n-1} + dt_n (true for every observer: dt is the time since the previous update, span the
// time since t_start) the recurrence  m_n = m_{n-1} + (v_n - m_{n-1}) dt_n / span_n  is, multiplied by span_n,
//     S_n = S_{n-1} + v_n dt_n,     S_n = m_n span_n,
// so one FMA per mean per step replaces a subtraction, an FMA and the shared division (9 FP64 instructions per
// step for `basic`).  open_means() converts the stored means on entry, S_0 = m_0 * (t_prev - t_start) — this also
// reproduces what the recurrence does when a continued run restarts the clock — and close_means() divides
// once before the features are emitted / the state is stored, so the persistent layout holds means in both tiers.
#if CLODE_FAST_SINGLE
// single precision keeps the recurrence (an FP32 running sum over 1e4..1e6 steps would lose digits) but forms the
// weight dt / span once per step and updates every mean with one FMA
#define CLODE_INTEGRAL_MEANS 0
struct MeanWeight {
    realtype w;
};
CLODE_DEV MeanWeight mean_weight(realtype dt, realtype span) { MeanWeight w = {div_nr(dt, span)}; return w; }
CLODE_DEV realtype mean_time(realtype mean, realtype v, const MeanWeight &w) { return fmaf(v - mean, w.w, mean); }
CLODE_DEV realtype mean_step(realtype mean, realtype v, realtype dt, realtype span) { return fmaf(v - mean, div_nr(dt, span), mean); }
#elif CLODE_EXACT_ARITH
#define CLODE_INTEGRAL_MEANS 0
struct MeanWeight {
    realtype dt, span;
};
CLODE_DEV MeanWeight mean_weight(realtype dt, realtype span) { MeanWeight w = {dt, span}; return w; }
CLODE_DEV realtype mean_time(realtype mean, realtype v, const MeanWeight &w) { return mean + (v - mean) * w.dt / w.span; }
CLODE_DEV realtype mean_step(realtype mean, realtype v, realtype dt, realtype span) { return mean + (v - mean) * dt / span; }
#else
#define CLODE_INTEGRAL_MEANS 1
struct MeanWeight {
    realtype dt;
};
CLODE_DEV MeanWeight mean_weight(realtype dt, realtype) { MeanWeight w = {dt}; return w; }
CLODE_DEV realtype mean_time(realtype sum, realtype v, const MeanWeight &w) { return fma(v, w.dt, sum); }
// a running mean that is consumed while it runs (thresh2's down-state mean) keeps the recurrence
CLODE_DEV realtype mean_step(realtype mean, realtype v, realtype dt, realtype span) { return fma(v - mean, div_nr(dt, span), mean); }
#endif
CLODE_DEV void mean_count(realtype *mean, realtype v, unsigned int count)
{
    if (count == 1) *mean = v;
    else if (count > 1) *mean += div_nr(v - *mean, (realtype)count);
}
// A mean over EVERY accepted step (the step-size statistic of the event observers, nhood1's per-step means): with
// count = 1, 2, 3, ... the recurrence above is the plain average, so the production build carries the SUM while a
// kernel runs — one addition per step instead of an int-to-double conversion and a division on a dependent chain — and
// converts on entry / exit like the time-weighted means (open_steps / close_steps).
#if CLODE_INTEGRAL_MEANS
CLODE_DEV void mean_steps(realtype *acc, realtype v, unsigned int) { *acc += v; }
CLODE_DEV realtype open_steps(realtype mean, unsigned int count) { return mean * (realtype)count; }
CLODE_DEV realtype close_steps(realtype sum, unsigned int count) { return count > 0 ? sum / (realtype)count : sum; }
#else
CLODE_DEV void mean_steps(realtype *acc, realtype v, unsigned int count) { mean_count(acc, v, count); }
CLODE_DEV realtype open_steps(realtype mean, unsigned int) { return mean; }
CLODE_DEV realtype close_steps(realtype mean, unsigned int) { return mean; }
#endif

// (max, min, running mean) accumulator used for every per-event statistic
struct Tri {
    realtype hi, lo, mean;
    __device__ __forceinline__ void reset() { hi = -BIG_REAL; lo = BIG_REAL; mean = ZERO; }
    __device__ __forceinline__ void push(realtype v, unsigned int count)
    {
        hi = max_nn(v, hi);
        lo = min_nn(v, lo);
        mean_count(&mean, v, count);
    }
    // the same for a statistic pushed on every accepted step (see mean_steps)
    __device__ __forceinline__ void push_step(realtype v, unsigned int count)
    {
        hi = max_nn(v, hi);
        lo = min_nn(v, lo);
        mean_steps(&mean, v, count);
    }
    template <class V> __device__ __forceinline__ void visit(V &v) { v(hi); v(lo); v(mean); }
};

// first-occurrence arg-extrema over a 3-sample window (clODE_utilities.cl:49-128)
CLODE_DEV int argmax3(realtype a, realtype b, realtype c, realtype &best)
{
    best = -BIG_REAL; int at = 0;
    if (a > best) { best = a; at = 0; }
    if (b > best) { best = b; at = 1; }
    if (c > best) { best = c; at = 2; }
    return at;
}
CLODE_DEV int argmin3(realtype a, realtype b, realtype c, realtype &best)
{
    best = BIG_REAL; int at = 0;
    if (a < best) { best = a; at = 0; }
    if (b < best) { best = b; at = 1; }
    if (c < best) { best = c; at = 2; }
    return at;
}
CLODE_DEV realtype pick3(const realtype v[3], int at) { return at == 0 ? v[0] : (at == 1 ? v[1] : v[2]); }

// extents and means of all variables, slopes and aux variables
//
// CLODE_EXT_SMEM: the 5 nVar + 3 nAux values live in shared memory, field-major [field][thread] (consecutive threads,
// consecutive words: conflict-free), instead of registers.  They are touched once per ACCEPTED step, outside the RK
// stages, and the extremes are almost never written after the first oscillation (a load and a compare per step), while
// in registers they cost the fat observers 46 registers (nVar = 4) that the trial step then has to spill around:
// C3's features kernel executed 72 local loads + 51 local stores per attempt before this.  The accesses are volatile
// so that the compiler does not promote the words back into registers for the whole time loop.
#if defined(CLODE_EXT_SMEM) && !defined(CLODE_OBS_SMEM) && (!defined(__CUDACC_EMU__) || defined(CLODE_EMU_EXT_SMEM))
#define CLODE_EXT_IN_SMEM 1
// thresh2 additionally keeps its four Schmitt-trigger thresholds there: written once when the observer is armed,
// read (two of them) on every accepted step of the features pass
#if defined(USE_OBSERVER_THRESHOLD_2)
#define CLODE_EXT_EXTRA_ROWS 4
#else
#define CLODE_EXT_EXTRA_ROWS 0
#endif
__shared__ volatile realtype clode_ext_smem[5 * NV + 3 * NA_ + CLODE_EXT_EXTRA_ROWS][CLODE_BLOCK];
template <int ROW> struct ExtScalar { // one per-thread real in the shared array, usable like a realtype member
    __device__ __forceinline__ operator realtype() const { return clode_ext_smem[ROW][threadIdx.x]; }
    __device__ __forceinline__ ExtScalar &operator=(const realtype v) { clode_ext_smem[ROW][threadIdx.x] = v; return *this; }
    __device__ __forceinline__ ExtScalar &operator=(const ExtScalar &o) { return *this = (realtype)o; }
};
template <int BASE, int STRIDE> struct ExtField {
    __device__ __forceinline__ volatile realtype &operator[](int j) const { return clode_ext_smem[BASE + STRIDE * j][threadIdx.x]; }
};
CLODE_DEV void ext_max(volatile realtype &m, const realtype v) { if (v > m) m = v; } // == max_nn(v, m), stored only on change
CLODE_DEV void ext_min(volatile realtype &m, const realtype v) { if (v < m) m = v; }
#else
#define CLODE_EXT_IN_SMEM 0
CLODE_DEV void ext_max(realtype &m, const realtype v) { m = max_nn(v, m); }
CLODE_DEV void ext_min(realtype &m, const realtype v) { m = min_nn(v, m); }
#endif
struct Extents {
#if CLODE_EXT_IN_SMEM
    ExtField<0, 5> xmax; ExtField<1, 5> xmin; ExtField<2, 5> xmean; ExtField<3, 5> dxmax; ExtField<4, 5> dxmin;
    ExtField<5 * NV + 0, 3> amax; ExtField<5 * NV + 1, 3> amin; ExtField<5 * NV + 2, 3> amean;
#else
    realtype xmax[NV], xmin[NV], xmean[NV], dxmax[NV], dxmin[NV];
    realtype amax[NA_], amin[NA_], amean[NA_];
#endif
    __device__ __forceinline__ void reset()
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            xmax[j] = -BIG_REAL; xmin[j] = BIG_REAL; xmean[j] = ZERO;
            dxmax[j] = -BIG_REAL; dxmin[j] = BIG_REAL;
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) { amax[j] = -BIG_REAL; amin[j] = BIG_REAL; amean[j] = ZERO; }
    }
    // time-weighted means (all observers except nhood1)
    __device__ __forceinline__ void update_time(const Instance &I, const MeanWeight &w)
    {
#if CLODE_EXT_IN_SMEM && defined(CLODE_EXT_BATCH)
        // Default since the end of round 2 (CLODE_EXT_BATCH=0 turns it off; enabled after the last GPU measurement of the round,
        // so its effect is NOT in the measured numbers — its results are checked bit for bit on the emulated device code,
        // tests/test_device_emu.py::test_extents_in_shared_memory_placement_matches_oracle): the volatile accesses keep the
        // words out of registers, but they also pin every load right before its use — the ncu source page of C3's features
        // launch shows 28 exposed shared-memory round trips per accepted step, a fifth of the launch's stall samples
        // (profiles/r02b_c3_features_source_hotspots.txt).  Here the words of a variable are loaded first (independent
        // LDS, in flight together), then compared / accumulated and stored exactly as below: same values, same stores.
        // One variable at a time — five loads in flight, four predicates live (a batch over all variables makes ptxas
        // spill predicates into a general register, ~50 LOP3 per step).
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const realtype c0 = xmax[j], c1 = xmin[j], c2 = xmean[j], c3 = dxmax[j], c4 = dxmin[j];
            if (I.x[j] > c0) xmax[j] = I.x[j];
            if (I.x[j] < c1) xmin[j] = I.x[j];
            xmean[j] = mean_time(c2, I.x[j], w);
            if (I.k1[j] > c3) dxmax[j] = I.k1[j];
            if (I.k1[j] < c4) dxmin[j] = I.k1[j];
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) {
            const realtype c0 = amax[j], c1 = amin[j], c2 = amean[j];
            if (I.aux[j] > c0) amax[j] = I.aux[j];
            if (I.aux[j] < c1) amin[j] = I.aux[j];
            amean[j] = mean_time(c2, I.aux[j], w);
        }
#else
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            ext_max(xmax[j], I.x[j]);
            ext_min(xmin[j], I.x[j]);
            xmean[j] = mean_time(xmean[j], I.x[j], w);
            ext_max(dxmax[j], I.k1[j]);
            ext_min(dxmin[j], I.k1[j]);
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) {
            ext_max(amax[j], I.aux[j]);
            ext_min(amin[j], I.aux[j]);
            amean[j] = mean_time(amean[j], I.aux[j], w);
        }
#endif
    }
    // mean <-> integral conversion of the time-weighted means (production build, see mean_time)
    __device__ __forceinline__ void open_means(realtype span)
    {
#if CLODE_INTEGRAL_MEANS
#pragma unroll
        for (int j = 0; j < NV; ++j) xmean[j] = xmean[j] * span;
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) amean[j] = amean[j] * span;
#endif
    }
    __device__ __forceinline__ void close_means(realtype span)
    {
#if CLODE_INTEGRAL_MEANS
        if (span != ZERO) { // span == 0: nothing was ever accumulated (sums are 0, as the means were)
#pragma unroll
            for (int j = 0; j < NV; ++j) xmean[j] = xmean[j] / span;
#pragma unroll
            for (int j = 0; j < N_AUX; ++j) amean[j] = amean[j] / span;
        }
#endif
    }
    __device__ __forceinline__ void open_counts(unsigned int count)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) xmean[j] = open_steps(xmean[j], count);
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) amean[j] = open_steps(amean[j], count);
    }
    __device__ __forceinline__ void close_counts(unsigned int count)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) xmean[j] = close_steps(xmean[j], count);
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) amean[j] = close_steps(amean[j], count);
    }
    // per-step means (nhood1: observer_neighborhood_1.clh:250-261)
    __device__ __forceinline__ void update_count(const Instance &I, unsigned int count)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            ext_max(xmax[j], I.x[j]);
            ext_min(xmin[j], I.x[j]);
            { realtype m = xmean[j]; mean_steps(&m, I.x[j], count); xmean[j] = m; }
            ext_max(dxmax[j], I.k1[j]);
            ext_min(dxmin[j], I.k1[j]);
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) {
            ext_max(amax[j], I.aux[j]);
            ext_min(amin[j], I.aux[j]);
            { realtype m = amean[j]; mean_steps(&m, I.aux[j], count); amean[j] = m; }
        }
    }
    // persistence visitor: through a temporary, so that it works for both placements
    template <class V, class F> __device__ __forceinline__ static void visit_one(V &v, F &&field)
    {
        realtype tmp = field;
        v(tmp);
        field = tmp;
    }
    template <class V> __device__ __forceinline__ void visit(V &v)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            visit_one(v, xmax[j]); visit_one(v, xmin[j]); visit_one(v, xmean[j]); visit_one(v, dxmax[j]); visit_one(v, dxmin[j]);
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) { visit_one(v, amax[j]); visit_one(v, amin[j]); visit_one(v, amean[j]); }
    }
};

// feature writer: F is feature-major, F[k * n + i]
struct FeatureOut {
    realtype *F;
    size_t n;
    size_t i;
    int k;
    __device__ __forceinline__ void put(realtype v) { F[(size_t)(k++) * n + i] = v; }
};

// write list[idx] = v without dynamic register indexing
CLODE_DEV void list_set(realtype list[NS_], unsigned int idx, realtype v)
{
#pragma unroll
    for (int j = 0; j < N_STORE_EVENTS; ++j)
        if ((unsigned int)j == idx) list[j] = v;
}

// =======================================================================================
#if defined(USE_OBSERVER_BASIC)
// observer_basic.clh:21-90 — extent and mean of one variable, no events.
#define CLODE_TWO_PASS 0
struct Observer {
    realtype xmax, xmin, xmean, dxmax, dxmin, t_last, t_start;
    unsigned int steps;

    __device__ __forceinline__ void init(const Instance &I)
    {
        xmax = -BIG_REAL; xmin = BIG_REAL; xmean = ZERO; dxmax = -BIG_REAL; dxmin = BIG_REAL;
        t_last = I.t; t_start = I.t; steps = 0;
    }
    __device__ __forceinline__ void warmup(const Instance &, const ObserverParams &) {}
    __device__ __forceinline__ void arm(const Instance &, const ObserverParams &) {}
    __device__ __forceinline__ void update(const Instance &I, const ObserverParams &)
    {
        ++steps;
        const realtype dt = I.t - t_last;
        t_last = I.t;
        const realtype span = I.t - t_start;
        xmax = max_nn(I.x[F_VAR_IX], xmax);
        xmin = min_nn(I.x[F_VAR_IX], xmin);
        xmean = mean_time(xmean, I.x[F_VAR_IX], mean_weight(dt, span));
        dxmax = max_nn(I.k1[F_VAR_IX], dxmax);
        dxmin = min_nn(I.k1[F_VAR_IX], dxmin);
    }
    __device__ __forceinline__ void open_means()
    {
#if CLODE_INTEGRAL_MEANS
        xmean *= t_last - t_start;
#endif
    }
    __device__ __forceinline__ void close_means()
    {
#if CLODE_INTEGRAL_MEANS
        const realtype span = t_last - t_start;
        if (span != ZERO) xmean /= span;
#endif
    }
    __device__ __forceinline__ bool event(const Instance &, const ObserverParams &) { return false; }
    __device__ __forceinline__ bool on_event(const Instance &, const ObserverParams &) { return false; }
    __device__ __forceinline__ void emit(FeatureOut &o) const
    {
        o.put(xmax); o.put(xmin); o.put(xmean); o.put(dxmax); o.put(dxmin); o.put((realtype)steps);
    }
    __device__ __forceinline__ void rebase(realtype T) { t_start -= T; }
    template <class V> __device__ __forceinline__ void visit(V &v)
    {
        v(xmax); v(xmin); v(xmean); v(dxmax); v(dxmin); v(t_last); v(t_start); v(steps);
    }    template <class V> __device__ __forceinline__ void visit_warmup(V &) {} // one-pass observer: no warm-up loop
};

// =======================================================================================
#elif defined(USE_OBSERVER_BASIC_ALLVAR)
// observer_basic_allVar.clh:35-132 — extents and means of all variables and aux, no events.
#define CLODE_TWO_PASS 0
struct Observer {
    Extents ext;
    realtype t_last, t_start;
    unsigned int steps;

    __device__ __forceinline__ void init(const Instance &I) { ext.reset(); t_last = I.t; t_start = I.t; steps = 0; }
    __device__ __forceinline__ void warmup(const Instance &, const ObserverParams &) {}
    __device__ __forceinline__ void arm(const Instance &, const ObserverParams &) {}
    __device__ __forceinline__ void update(const Instance &I, const ObserverParams &)
    {
        ++steps;
        const realtype dt = I.t - t_last;
        t_last = I.t;
        ext.update_time(I, mean_weight(dt, I.t - t_start));
    }
    __device__ __forceinline__ void open_means() { ext.open_means(t_last - t_start); }
    __device__ __forceinline__ void close_means() { ext.close_means(t_last - t_start); }
    __device__ __forceinline__ bool event(const Instance &, const ObserverParams &) { return false; }
    __device__ __forceinline__ bool on_event(const Instance &, const ObserverParams &) { return false; }
    __device__ __forceinline__ void emit(FeatureOut &o) const
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            o.put(ext.xmax[j]); o.put(ext.xmin[j]); o.put(ext.xmean[j]); o.put(ext.dxmax[j]); o.put(ext.dxmin[j]);
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) { o.put(ext.amax[j]); o.put(ext.amin[j]); o.put(ext.amean[j]); }
        o.put((realtype)steps);
    }
    __device__ __forceinline__ void rebase(realtype T) { t_start -= T; }
    template <class V> __device__ __forceinline__ void visit(V &v) { ext.visit(v); v(t_last); v(t_start); v(steps); }    template <class V> __device__ __forceinline__ void visit_warmup(V &) {} // one-pass observer: no warm-up loop
};

// =======================================================================================
#elif defined(USE_OBSERVER_LOCAL_MAX)
// observer_local_maximum.clh:54-328 — event = local maximum of x[fVar] (slope + -> -).
#define CLODE_TWO_PASS 0
struct Observer {
    realtype tb[3], xf[3]; // time / feature-variable history (oldest first)
    realtype d1, d2;       // feature-variable slope at the previous and the current step
    Extents ext;
    realtype t_peak[NS_], x_peak[NS_], t_dip[NS_], x_dip[NS_];
    Tri imi, amp;
    realtype t_start, t_last_max, t_last_min, x_last_min;
    unsigned int events, steps;

    __device__ __forceinline__ void init(const Instance &I)
    {
        tb[0] = tb[1] = ZERO; tb[2] = I.t;
        xf[0] = xf[1] = ZERO; xf[2] = I.x[F_VAR_IX];
        d1 = ZERO; d2 = I.k1[F_VAR_IX];
        ext.reset();
#pragma unroll
        for (int j = 0; j < NS_; ++j) { t_peak[j] = x_peak[j] = t_dip[j] = x_dip[j] = ZERO; }
        imi.reset(); amp.reset();
        t_start = I.t; t_last_max = ZERO; t_last_min = ZERO; x_last_min = BIG_REAL;
        events = 0; steps = 0;
    }
    __device__ __forceinline__ void warmup(const Instance &, const ObserverParams &) {}
    __device__ __forceinline__ void arm(const Instance &, const ObserverParams &) {}
    __device__ __forceinline__ void update(const Instance &I, const ObserverParams &)
    {
        ++steps;
        tb[0] = tb[1]; tb[1] = tb[2]; tb[2] = I.t;
        xf[0] = xf[1]; xf[1] = xf[2]; xf[2] = I.x[F_VAR_IX];
        d1 = d2; d2 = I.k1[F_VAR_IX];
        ext.update_time(I, mean_weight(tb[2] - tb[1], I.t - t_start));
        if (steps < 2) return;
        if (d1 < 0.0 && d2 > 0.0) { // local minimum between maxima (:269-282)
            realtype lowest;
            const int at = argmin3(xf[0], xf[1], xf[2], lowest);
            t_last_min = pick3(tb, at);
            x_last_min = pick3(xf, at); // the value AT the arg-min (:276), not the running best
            // the reference indexes [eventcount-1] without checking eventcount > 0
            // (out-of-bounds write); guarded here and in the oracle
            if (events > 0 && events <= N_STORE_EVENTS) {
                list_set(t_dip, events - 1, t_last_min);
                list_set(x_dip, events - 1, x_last_min);
            }
        }
    }
    __device__ __forceinline__ void open_means() { ext.open_means(tb[2] - t_start); }
    __device__ __forceinline__ void close_means() { ext.close_means(tb[2] - t_start); }
    __device__ __forceinline__ bool event(const Instance &, const ObserverParams &)
    {
        return steps >= 2 && d1 > 0.0 && d2 < 0.0;
    }
    __device__ __forceinline__ bool on_event(const Instance &, const ObserverParams &op)
    {
        realtype x_pk;
        const int at = argmax3(xf[0], xf[1], xf[2], x_pk);
        const realtype t_pk = pick3(tb, at);
        x_pk = pick3(xf, at); // value AT the arg-max (differs from the running best only for NaN input)
        ++events;
        if (events > 1) {
            imi.push(t_pk - t_last_max, events - 1);
            amp.push(x_pk - x_last_min, events - 1);
        }
        t_last_max = t_pk;
        if (events <= N_STORE_EVENTS) {
            list_set(t_peak, events - 1, t_pk);
            list_set(x_peak, events - 1, x_pk);
        }
        return events == op.maxEventCount;
    }
    __device__ __forceinline__ void emit(FeatureOut &o) const
    {
        const bool multi = events > 1;
        o.put(multi ? imi.hi : ZERO); o.put(multi ? imi.lo : ZERO); o.put(multi ? imi.mean : ZERO);
        o.put(multi ? amp.hi : ZERO); o.put(multi ? amp.lo : ZERO); o.put(multi ? amp.mean : ZERO);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            o.put(ext.xmax[j]); o.put(ext.xmin[j]); o.put(ext.xmean[j]); o.put(ext.dxmax[j]); o.put(ext.dxmin[j]);
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) { o.put(ext.amax[j]); o.put(ext.amin[j]); o.put(ext.amean[j]); }
#pragma unroll
        for (int j = 0; j < N_STORE_EVENTS; ++j) { o.put(t_peak[j]); o.put(x_peak[j]); o.put(t_dip[j]); o.put(x_dip[j]); }
        o.put((realtype)events);
        o.put((realtype)steps);
    }
    __device__ __forceinline__ void rebase(realtype T)
    {
        t_start -= T;
        t_last_max = t_last_max - T;
        t_last_min = t_last_min - T;
#pragma unroll
        for (int k = 0; k < 3; ++k) tb[k] = tb[k] - T;
    }
    template <class V> __device__ __forceinline__ void visit(V &v)
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) { v(tb[k]); v(xf[k]); }
        v(d1); v(d2);
        ext.visit(v);
#pragma unroll
        for (int j = 0; j < N_STORE_EVENTS; ++j) { v(t_peak[j]); v(x_peak[j]); v(t_dip[j]); v(x_dip[j]); }
        imi.visit(v); amp.visit(v);
        v(t_start); v(t_last_max); v(t_last_min); v(x_last_min);
        v(events); v(steps);
    }    template <class V> __device__ __forceinline__ void visit_warmup(V &) {} // one-pass observer: no warm-up loop
};

// =======================================================================================
#elif defined(USE_OBSERVER_NEIGHBORHOOD_1)
// observer_neighborhood_1.clh:44-361 — centre = state at the first local minimum of
// x[eVar]; event = ENTRY into the L2 ball around it (range-normalised coordinates).
#define CLODE_TWO_PASS 0
struct Observer {
    realtype tb[3], xb[NV][3]; // the centre is captured from the history, so all variables are kept
    realtype de1, de2, df1, df2; // slopes of the event / feature variable (previous, current)
    realtype center[NV];
    Extents ext;
    Tri peaks_stat, period, step_dt;
    realtype t_start, t_last_event;
    unsigned int peaks, events, steps, found, inside;

    __device__ __forceinline__ void init(const Instance &I)
    {
        tb[0] = tb[1] = ZERO; tb[2] = I.t;
#pragma unroll
        for (int j = 0; j < NV; ++j) { xb[j][0] = xb[j][1] = ZERO; xb[j][2] = I.x[j]; center[j] = ZERO; }
        de1 = df1 = ZERO; de2 = I.k1[E_VAR_IX]; df2 = I.k1[F_VAR_IX];
        ext.reset();
        peaks_stat.reset(); period.reset(); step_dt.reset();
        t_start = I.t; t_last_event = ZERO; // left unset by the reference (:93-157); never read before being written
        peaks = events = steps = found = inside = 0;
    }
    __device__ __forceinline__ void warmup(const Instance &, const ObserverParams &) {}
    __device__ __forceinline__ void arm(const Instance &, const ObserverParams &) {}
    __device__ __forceinline__ void update(const Instance &I, const ObserverParams &)
    {
        ++steps;
        tb[0] = tb[1]; tb[1] = tb[2]; tb[2] = I.t;
#pragma unroll
        for (int j = 0; j < NV; ++j) { xb[j][0] = xb[j][1]; xb[j][1] = xb[j][2]; xb[j][2] = I.x[j]; }
        de1 = de2; de2 = I.k1[E_VAR_IX];
        df1 = df2; df2 = I.k1[F_VAR_IX];
        step_dt.push_step(tb[2] - tb[1], steps);
        ext.update_count(I, steps);
        if (steps > 1) {
            if (!found) {
                if (de1 <= 0.0 && de2 > 0.0) {
                    realtype lowest;
                    const int at = argmin3(xb[E_VAR_IX][0], xb[E_VAR_IX][1], xb[E_VAR_IX][2], lowest);
                    t_last_event = pick3(tb, at);
                    found = 1;
#pragma unroll
                    for (int j = 0; j < NV; ++j) center[j] = pick3(xb[j], at);
                }
            } else if (df1 >= 0.0 && df2 < 0.0) {
                peaks++;
            }
        }
    }
    // per-step means only
    __device__ __forceinline__ void open_means() { ext.open_counts(steps); step_dt.mean = open_steps(step_dt.mean, steps); }
    __device__ __forceinline__ void close_means() { ext.close_counts(steps); step_dt.mean = close_steps(step_dt.mean, steps); }
    __device__ __forceinline__ bool event(const Instance &I, const ObserverParams &op)
    {
        if (steps < 2 || !found) return false;
        if (ext.xmax[F_VAR_IX] - ext.xmin[F_VAR_IX] < op.minXamp) return false;
        const unsigned int was = inside;
        realtype d[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) d[j] = fabs(I.x[j] - center[j]) / (ext.xmax[j] - ext.xmin[j]);
        inside = norm_2(d, NV) <= op.nHoodRadius;
        return inside & !was;
    }
    __device__ __forceinline__ bool on_event(const Instance &I, const ObserverParams &op)
    {
        ++events;
        if (events > 1) {
            peaks_stat.push((realtype)peaks, events - 1);
            period.push(I.t - t_last_event, events - 1);
        }
        t_last_event = I.t;
        peaks = 0;
        return events >= op.maxEventCount;
    }
    __device__ __forceinline__ void emit(FeatureOut &o) const
    {
        const bool multi = events > 1;
        o.put(multi ? period.hi : ZERO); o.put(multi ? period.lo : ZERO); o.put(multi ? period.mean : ZERO);
        o.put(multi ? peaks_stat.hi : ZERO); o.put(multi ? peaks_stat.lo : ZERO); o.put(multi ? peaks_stat.mean : ZERO);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            o.put(ext.xmax[j]); o.put(ext.xmin[j]); o.put(ext.xmean[j]); o.put(ext.dxmax[j]); o.put(ext.dxmin[j]);
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) { o.put(ext.amax[j]); o.put(ext.amin[j]); o.put(ext.amean[j]); }
        o.put((realtype)(events - 1u)); // "period count": unsigned wrap when no event, as the reference (:343)
        o.put((realtype)steps);
        o.put(step_dt.hi); o.put(step_dt.lo); o.put(step_dt.mean);
    }
    __device__ __forceinline__ void rebase(realtype T)
    {
        t_start -= T;
        t_last_event -= T;
#pragma unroll
        for (int k = 0; k < 3; ++k) tb[k] -= T;
    }
    template <class V> __device__ __forceinline__ void visit(V &v)
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) v(tb[k]);
#pragma unroll
        for (int j = 0; j < NV; ++j) { v(xb[j][0]); v(xb[j][1]); v(xb[j][2]); v(center[j]); }
        v(de1); v(de2); v(df1); v(df2);
        ext.visit(v);
        peaks_stat.visit(v); period.visit(v); step_dt.visit(v);
        v(t_start); v(t_last_event);
        v(peaks); v(events); v(steps); v(found); v(inside);
    }    template <class V> __device__ __forceinline__ void visit_warmup(V &) {} // one-pass observer: no warm-up loop
};

// =======================================================================================
#elif defined(USE_OBSERVER_NEIGHBORHOOD_2)
// observer_neighborhood_2.clh:50-313 — two-pass: warm-up finds per-variable ranges; centre =
// state when x[eVar] first drops through min + xDownThresh*range; event = EXIT from the ball.
#define CLODE_TWO_PASS 1
struct Observer {
    realtype tb[3];
    realtype xe1, xe2; // event-variable history (previous, current)
    realtype df1, df2; // feature-variable slope (previous, current)
    realtype center[NV], range[NV];
    Extents ext;
    realtype t_exit[NS_];
    Tri peaks_stat, period, step_dt;
    realtype t_start, t_last_event, x_threshold;
    unsigned int peaks, found, inside, events, steps;

    __device__ __forceinline__ void init(const Instance &I)
    {
        tb[0] = tb[1] = ZERO; tb[2] = I.t;
        xe1 = ZERO; xe2 = I.x[E_VAR_IX];
        df1 = ZERO; df2 = I.k1[F_VAR_IX];
#pragma unroll
        for (int j = 0; j < NV; ++j) { center[j] = I.x[j]; range[j] = ZERO; }
        ext.reset();
#pragma unroll
        for (int j = 0; j < NS_; ++j) t_exit[j] = ZERO;
        peaks_stat.reset(); period.reset(); step_dt.reset();
        t_start = I.t; t_last_event = ZERO; x_threshold = ZERO;
        peaks = found = inside = events = steps = 0;
    }
    __device__ __forceinline__ void warmup(const Instance &I, const ObserverParams &)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            ext_max(ext.xmax[j], I.x[j]);
            ext_min(ext.xmin[j], I.x[j]);
        }
    }
    __device__ __forceinline__ void arm(const Instance &, const ObserverParams &op)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) range[j] = ext.xmax[j] - ext.xmin[j];
        x_threshold = ext.xmin[E_VAR_IX] + op.xDownThresh * range[E_VAR_IX];
    }
    __device__ __forceinline__ void update(const Instance &I, const ObserverParams &)
    {
        ++steps;
        tb[0] = tb[1]; tb[1] = tb[2]; tb[2] = I.t;
        xe1 = xe2; xe2 = I.x[E_VAR_IX];
        df1 = df2; df2 = I.k1[F_VAR_IX];
        const realtype dt = tb[2] - tb[1];
        step_dt.push_step(dt, steps);
        ext.update_time(I, mean_weight(dt, I.t - t_start));
        if (steps < 2) return;
        if (found) {
            if (df1 >= 0.0 && df2 < 0.0) peaks++;
            return;
        }
        if (xe1 > x_threshold && xe2 < x_threshold) {
            found = 1;
            inside = 1;
#pragma unroll
            for (int j = 0; j < NV; ++j) center[j] = I.x[j];
        }
    }
    __device__ __forceinline__ void open_means() { ext.open_means(tb[2] - t_start); step_dt.mean = open_steps(step_dt.mean, steps); }
    __device__ __forceinline__ void close_means() { ext.close_means(tb[2] - t_start); step_dt.mean = close_steps(step_dt.mean, steps); }
    __device__ __forceinline__ bool event(const Instance &I, const ObserverParams &op)
    {
        if (steps < 2 || !found) return false;
        realtype d[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) d[j] = (I.x[j] - center[j]) / range[j];
        const unsigned int was = inside;
        inside = norm_2(d, NV) < op.nHoodRadius;
        return was && !inside;
    }
    __device__ __forceinline__ bool on_event(const Instance &I, const ObserverParams &op)
    {
        ++events;
        if (events > 1) {
            peaks_stat.push((realtype)peaks, events - 1);
            period.push(I.t - t_last_event, events - 1);
        }
        t_last_event = I.t;
        peaks = 0;
        if (events <= N_STORE_EVENTS) list_set(t_exit, events - 1, I.t);
        return events == op.maxEventCount;
    }
    __device__ __forceinline__ void emit(FeatureOut &o) const
    {
        const bool multi = events > 1;
        o.put(multi ? period.hi : ZERO); o.put(multi ? period.lo : ZERO); o.put(multi ? period.mean : ZERO);
        o.put(multi ? peaks_stat.hi : ZERO); o.put(multi ? peaks_stat.lo : ZERO); o.put(multi ? peaks_stat.mean : ZERO);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            o.put(ext.xmax[j]); o.put(ext.xmin[j]); o.put(ext.xmean[j]); o.put(range[j]); o.put(center[j]);
            o.put(ext.dxmax[j]); o.put(ext.dxmin[j]);
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) { o.put(ext.amax[j]); o.put(ext.amin[j]); o.put(ext.amean[j]); }
#pragma unroll
        for (int j = 0; j < N_STORE_EVENTS; ++j) o.put(t_exit[j]);
        o.put((realtype)events);
        o.put((realtype)steps);
        o.put(step_dt.hi); o.put(step_dt.lo); o.put(step_dt.mean);
    }
    __device__ __forceinline__ void rebase(realtype T)
    {
        t_start -= T;
        t_last_event -= T;
#pragma unroll
        for (int k = 0; k < 3; ++k) tb[k] -= T;
    }
    template <class V> __device__ __forceinline__ void visit(V &v)
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) v(tb[k]);
        v(xe1); v(xe2); v(df1); v(df2);
#pragma unroll
        for (int j = 0; j < NV; ++j) { v(center[j]); v(range[j]); }
        ext.visit(v);
#pragma unroll
        for (int j = 0; j < N_STORE_EVENTS; ++j) v(t_exit[j]);
        peaks_stat.visit(v); period.visit(v); step_dt.visit(v);
        v(t_start); v(t_last_event); v(x_threshold);
        v(peaks); v(found); v(inside); v(events); v(steps);
    }
    // the fields the warm-up pass accumulates (what a parked warm-up has to carry; the rest is init()'s)
    template <class V> __device__ __forceinline__ void visit_warmup(V &v)
    {
#pragma unroll
        for (int j = 0; j < NV; ++j) { Extents::visit_one(v, ext.xmax[j]); Extents::visit_one(v, ext.xmin[j]); }
    }
};

// =======================================================================================
#elif defined(USE_OBSERVER_THRESHOLD_2)
// observer_threshold_2.clh:63-470 — two-pass Schmitt trigger on x[eVar]: warm-up finds the
// global extent of x and dx, thresholds are fractions of it; event = upward crossing.
#define CLODE_TWO_PASS 1
struct Observer {
    realtype tb[3], xf[3]; // time / feature-variable history
    realtype d1, d2;       // feature-variable slope (previous, current)
    Extents ext;
    realtype t_up[NS_], t_down[NS_];
    Tri peaks_stat, period, up_time, down_time, duty, dip, step_dt;
    realtype down_mean;
    realtype g_xmax, g_xmin, g_dxmax, g_dxmin;
#if CLODE_EXT_IN_SMEM
    ExtScalar<5 * NV + 3 * NA_ + 0> x_up; ExtScalar<5 * NV + 3 * NA_ + 1> x_down;
    ExtScalar<5 * NV + 3 * NA_ + 2> dx_up; ExtScalar<5 * NV + 3 * NA_ + 3> dx_down;
#else
    realtype x_up, x_down, dx_up, dx_down;
#endif
    realtype t_start, t_last_event, t_this_down, x_last_min;
    unsigned int peaks, steps, events, up;

    __device__ __forceinline__ void init(const Instance &I)
    {
        tb[0] = tb[1] = ZERO; tb[2] = I.t;
        xf[0] = xf[1] = ZERO; xf[2] = I.x[F_VAR_IX];
        d1 = ZERO; d2 = I.k1[F_VAR_IX];
        ext.reset();
#pragma unroll
        for (int j = 0; j < NS_; ++j) { t_up[j] = ZERO; t_down[j] = ZERO; }
        peaks_stat.reset(); period.reset(); up_time.reset(); down_time.reset(); duty.reset(); dip.reset();
        step_dt.reset();
        down_mean = ZERO;
        g_xmax = -BIG_REAL; g_xmin = BIG_REAL; g_dxmax = -BIG_REAL; g_dxmin = BIG_REAL;
        x_up = ZERO; x_down = ZERO; dx_up = ZERO; dx_down = ZERO;
        t_start = I.t; t_last_event = ZERO; t_this_down = ZERO; x_last_min = BIG_REAL;
        peaks = steps = events = up = 0;
    }
    __device__ __forceinline__ void warmup(const Instance &I, const ObserverParams &)
    {
        g_xmax = max_nn(I.x[E_VAR_IX], g_xmax);
        g_xmin = min_nn(I.x[E_VAR_IX], g_xmin);
        g_dxmax = max_nn(I.k1[E_VAR_IX], g_dxmax);
        g_dxmin = min_nn(I.k1[E_VAR_IX], g_dxmin);
    }
    __device__ __forceinline__ void arm(const Instance &I, const ObserverParams &op)
    {
        const realtype amp = g_xmax - g_xmin;
        const realtype xu = g_xmin + op.xUpThresh * amp;
        x_up = xu;
        x_down = op.xDownThresh > ZERO ? g_xmin + op.xDownThresh * amp : xu;
        dx_up = op.dxUpThresh * g_dxmax;
        dx_down = op.dxDownThresh > ZERO ? op.dxDownThresh * g_dxmin : g_dxmin;
        up = I.x[E_VAR_IX] > xu ? 1 : 0;
    }
    __device__ __forceinline__ void update(const Instance &I, const ObserverParams &)
    {
        ++steps;
        tb[0] = tb[1]; tb[1] = tb[2]; tb[2] = I.t;
        xf[0] = xf[1]; xf[1] = xf[2]; xf[2] = I.x[F_VAR_IX];
        d1 = d2; d2 = I.k1[F_VAR_IX];
        const realtype dt = tb[2] - tb[1];
        step_dt.push_step(dt, steps);
        ext.update_time(I, mean_weight(dt, I.t - t_start));
        if (steps > 1) {
            if (d1 > 0.0 && d2 < 0.0) peaks++; // local maximum of the feature variable
            if (d1 < 0.0 && d2 > 0.0) {        // local minimum: remember its value
                (void)argmin3(xf[0], xf[1], xf[2], x_last_min);
            }
            if (up) {
                if (I.x[E_VAR_IX] <= x_down && I.k1[E_VAR_IX] >= dx_down) { // downward crossing
                    t_this_down = I.t;
                    up = 0;
                    if (events > 0 && events <= N_STORE_EVENTS) list_set(t_down, events - 1, t_this_down);
#if CLODE_INTEGRAL_MEANS
                    down_mean = ZERO; // the integral restarts; the sample at the crossing has zero weight in the reference too
#else
                    down_mean = I.x[F_VAR_IX];
#endif
                }
            } else {
                const realtype since = I.t - t_this_down;
#if CLODE_INTEGRAL_MEANS
                if (since > 0.0) down_mean = fma(I.x[F_VAR_IX], dt, down_mean);
#else
                if (since > 0.0) down_mean = mean_step(down_mean, I.x[F_VAR_IX], dt, since);
#endif
            }
        }
    }
    // down_mean: the time-weighted mean of x[fVar] since the last downward crossing.  The reference's recurrence
    // m += (x - m) dt / (t - t_down) is, multiplied by (t - t_down), the integral S += x dt (t - dt is the previous
    // sample's time), with S = 0 at a crossing; the mean is only READ when the next event fires.  Production build:
    // one FMA per step instead of a division, the quotient once per event (down_mean_now) and on exit.
    __device__ __forceinline__ void open_means()
    {
        ext.open_means(tb[2] - t_start);
        step_dt.mean = open_steps(step_dt.mean, steps);
#if CLODE_INTEGRAL_MEANS
        const realtype since = tb[2] - t_this_down;
        if (since > ZERO) down_mean *= since;
#endif
    }
    __device__ __forceinline__ void close_means()
    {
        ext.close_means(tb[2] - t_start);
        step_dt.mean = close_steps(step_dt.mean, steps);
#if CLODE_INTEGRAL_MEANS
        const realtype since = tb[2] - t_this_down;
        if (since > ZERO) down_mean /= since;
#endif
    }
    __device__ __forceinline__ realtype down_mean_now(realtype now) const
    {
#if CLODE_INTEGRAL_MEANS
        const realtype since = now - t_this_down;
        return since > ZERO ? down_mean / since : down_mean;
#else
        return down_mean;
#endif
    }
    __device__ __forceinline__ bool event(const Instance &I, const ObserverParams &op)
    {
        if (steps < 2) return false;
        if (up) return false;
        if (!(I.x[E_VAR_IX] > x_up && I.k1[E_VAR_IX] > dx_up)) return false;
        // the amplitude test (observer_threshold_2.clh:331) last: it only decides at an upward crossing, where the
        // reference evaluates it first — same result, two FP64 instructions fewer on every other step
        return !(g_xmax - g_xmin < op.minXamp);
    }
    __device__ __forceinline__ bool on_event(const Instance &I, const ObserverParams &op)
    {
        ++events;
        up = 1;
        const realtype now = I.t;
        if (events > 1) {
            const unsigned int n = events - 1;
            peaks_stat.push((realtype)peaks, n);
            const realtype this_period = now - t_last_event;
            period.push(this_period, n);
            const realtype this_up = t_this_down - t_last_event;
            up_time.push(this_up, n);
            down_time.push(now - t_this_down, n);
            duty.push(this_up / this_period, n);
            dip.push(down_mean_now(now) - x_last_min, n);
        }
        if (events <= N_STORE_EVENTS) list_set(t_up, events - 1, now);
        t_last_event = now;
        peaks = 0;
        return events == op.maxEventCount;
    }
    __device__ __forceinline__ void emit(FeatureOut &o) const
    {
        const bool multi = events > 1;
#define PUT3(T) o.put(multi ? T.hi : ZERO); o.put(multi ? T.lo : ZERO); o.put(multi ? T.mean : ZERO)
        PUT3(period); PUT3(peaks_stat); PUT3(up_time); PUT3(down_time); PUT3(duty); PUT3(dip);
#undef PUT3
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            o.put(ext.xmax[j]); o.put(ext.xmin[j]); o.put(ext.xmean[j]); o.put(ext.dxmax[j]); o.put(ext.dxmin[j]);
        }
#pragma unroll
        for (int j = 0; j < N_AUX; ++j) { o.put(ext.amax[j]); o.put(ext.amin[j]); o.put(ext.amean[j]); }
#pragma unroll
        for (int j = 0; j < N_STORE_EVENTS; ++j) { o.put(t_up[j]); o.put(t_down[j]); }
        o.put((realtype)events);
        o.put((realtype)steps);
        o.put(step_dt.hi); o.put(step_dt.lo); o.put(step_dt.mean);
    }
    __device__ __forceinline__ void rebase(realtype T)
    {
        t_start -= T;
        t_last_event -= T;
        t_this_down -= T;
#pragma unroll
        for (int k = 0; k < 3; ++k) tb[k] -= T;
    }
    template <class V> __device__ __forceinline__ void visit(V &v)
    {
#pragma unroll
        for (int k = 0; k < 3; ++k) { v(tb[k]); v(xf[k]); }
        v(d1); v(d2);
        ext.visit(v);
#pragma unroll
        for (int j = 0; j < N_STORE_EVENTS; ++j) { v(t_up[j]); v(t_down[j]); }
        peaks_stat.visit(v); period.visit(v); up_time.visit(v); down_time.visit(v); duty.visit(v); dip.visit(v);
        step_dt.visit(v);
        v(down_mean);
        v(g_xmax); v(g_xmin); v(g_dxmax); v(g_dxmin);
        Extents::visit_one(v, x_up); Extents::visit_one(v, x_down); Extents::visit_one(v, dx_up); Extents::visit_one(v, dx_down);
        v(t_start); v(t_last_event); v(t_this_down); v(x_last_min);
        v(peaks); v(steps); v(events); v(up);
    }
    // the fields the warm-up pass accumulates (what a parked warm-up has to carry; the rest is init()'s)
    template <class V> __device__ __forceinline__ void visit_warmup(V &v) { v(g_xmax); v(g_xmin); v(g_dxmax); v(g_dxmin); }
};

#else
#error "no observer selected"
#endif

// ---- structure-of-arrays persistence of the observer state -----------------------------
struct ObsCount {
    int nreal, nuint;
    __device__ __forceinline__ void operator()(realtype &) { ++nreal; }
    __device__ __forceinline__ void operator()(unsigned int &) { ++nuint; }
};
struct ObsStore {
    realtype *r; unsigned int *u; size_t n, i; int kr, ku;
    __device__ __forceinline__ void operator()(realtype &x) { r[(size_t)(kr++) * n + i] = x; }
    __device__ __forceinline__ void operator()(unsigned int &x) { u[(size_t)(ku++) * n + i] = x; }
};
struct ObsLoad {
    const realtype *r; const unsigned int *u; size_t n, i; int kr, ku;
    __device__ __forceinline__ void operator()(realtype &x) { x = r[(size_t)(kr++) * n + i]; }
    __device__ __forceinline__ void operator()(unsigned int &x) { x = u[(size_t)(ku++) * n + i]; }
};

#endif // CLODE_OBSERVERS_CUH
