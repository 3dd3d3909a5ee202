// kernels.cuh — the four ensemble kernels (sm_100a), one ODE instance per thread.
//
//   clode_transient          <- clode/cpp/transient.cl:9-77
//   clode_initialize_observer<- clode/cpp/initializeObserver.cl:9-83
//   clode_features           <- clode/cpp/features.cl:10-106
//   clode_trajectory         <- clode/cpp/trajectory.cl:14-112
//
// Data layout (identical to the reference API): every ensemble array is
// variable-major with row pitch n = number of instances in the launch, so lane i of a
// warp touches element [row*n + i] and each warp-wide access is one contiguous,
// fully-coalesced 128/256-byte segment.  State, RK stages, parameters, aux and observer
// data of an instance live in registers for the whole time loop; HBM is touched only in
// the prologue / epilogue (and by trajectory stores).
//
// Time loop: an ATTEMPT loop (see steppers.cuh).  A lane leaves it when its instance is
// finished; with CLODE_WORK_QUEUE the lane then pulls the next unprocessed instance from
// a global counter instead of idling (persistent threads, warp-aggregated atomics).
#ifndef CLODE_KERNELS_CUH
#define CLODE_KERNELS_CUH

#ifndef CLODE_BLOCK
#define CLODE_BLOCK 128
#endif
#ifndef CLODE_MIN_BLOCKS
#define CLODE_MIN_BLOCKS 4
#endif
// Occupancy target (second __launch_bounds__ argument) per kernel: the kernels of one program carry very different
// amounts of state — initializeObserver's warm-up pass holds four extents where the features pass holds the whole
// thresh2 / nhood2 record — so the runtime picks each kernel's register budget separately (clode_sim_build).
#ifndef CLODE_MIN_BLOCKS_TRANSIENT
#define CLODE_MIN_BLOCKS_TRANSIENT CLODE_MIN_BLOCKS
#endif
#ifndef CLODE_MIN_BLOCKS_INIT
#define CLODE_MIN_BLOCKS_INIT CLODE_MIN_BLOCKS
#endif
#ifndef CLODE_MIN_BLOCKS_FEATURES
#define CLODE_MIN_BLOCKS_FEATURES CLODE_MIN_BLOCKS
#endif
#ifndef CLODE_MIN_BLOCKS_TRAJECTORY
#define CLODE_MIN_BLOCKS_TRAJECTORY CLODE_MIN_BLOCKS
#endif

// Every kernel starts by staging the tables of the production-double math (exp: cl_compat.cuh / fast_exp.cuh; the polar
// method's logarithm: rng.cuh / fast_polar.cuh) in shared memory; all threads of the block, before any of them returns.
CLODE_DEV void clode_kernel_prologue()
{
#if CLODE_HAVE_FAST_EXP
    clode_stage_exp_table();
#endif
#if CLODE_HAVE_FAST_POLAR
    clode_stage_polar_table();
#endif
}
#define CLODE_KERNEL_PROLOGUE() clode_kernel_prologue()

// Launch arguments: one struct in __constant__ memory (`clode_args`), written by the host
// with an in-stream copy before each launch; every field is a warp-uniform constant-bank load.
// NOTE: deliberately NOT a by-value kernel parameter.  With a by-value struct larger than
// ~128 bytes (with or without __grid_constant__) NVVM 12.9 keeps the argument in param space
// behind a pointer and then MISCOMPILES the time loop: the floating-point exit test
// `t <= t_end` is dropped (the loop runs max_steps iterations) and t is advanced once.
// Found on the first GPU run by the parity tests; reproduced with a 20-line kernel
// (profiles/r01_nvvm_byval_param_miscompile.md).  __constant__ and global-pointer arguments
// compile correctly.
// Scalars travel as double and are narrowed on the device when realtype is float — the
// same round-to-nearest conversion the reference does on the host (CLODE.cpp:292-300, 377-394).
struct KernelArgs {
    double t0, t1;
    double sp_dt, sp_dtmax, sp_abstol, sp_reltol;
    unsigned int sp_max_steps, sp_max_store, sp_nout;
    unsigned int op_max_event_count;
    double op_min_x_amp, op_min_imi, op_nhood_radius, op_x_up, op_x_down, op_dx_up, op_dx_down, op_eps_dx;
    unsigned long long n;          // instances in this launch == row pitch of every array
    const void *x0, *pars;         // [N_VAR][n], [N_PAR][n]
    void *xf;                      // [N_VAR][n]
    unsigned long long *rng;       // [2][n]
    void *dt, *tf;                 // [n]
    unsigned int *steps;           // [n] accepted steps of this call (may be null)
    void *od_real;                 // observer state, [n_od_real][n]
    unsigned int *od_uint;         // observer state, [n_od_uint][n]
    void *F;                       // [n_features][n]
    void *tr_t, *tr_x, *tr_dx, *tr_aux; // [rows][n], [rows][N_VAR][n], [rows][N_VAR][n], [rows][N_AUX][n]
    int *n_stored;                 // [n]
    unsigned long long *queue;     // work-queue head (CLODE_WORK_QUEUE)
    // chunked ("streamed") trajectory: one launch stores global rows [row_begin, row_end) into buffers that hold
    // only those rows; the per-instance state needed to resume travels in xf/tf/dt/rng plus the two arrays below
    void *rs_real;                 // [1 + N_WIENER][n]: cached normal variate, noise values of the next step
    unsigned int *rs_uint;         // [3][n]: accepted steps so far, rows stored so far, variate cached?
    unsigned int *chunk_flags;     // [2]: an instance is still live after this launch / highest row index stored
    unsigned int row_begin, row_end, resume; // monolithic launch: 0, 0xffffffff, 0
    // Block -> chunk of CLODE_BLOCK consecutive instances.  Blocks are dispatched in index order, so when the cost of
    // an instance grows along the ensemble (a sorted parameter sweep) the most expensive warps start last and the
    // device drains behind them; walking the ensemble backwards puts the cheap ones last (longest-processing-time
    // first).  0: forward, 1: reverse, 2: decide from cost_in — the accepted steps the previous launch on this
    // ensemble spent in the lower / upper half of the index range (every kernel accumulates them into cost_out).
    unsigned int block_order;
    const unsigned long long *cost_in; // [2] or null
    unsigned long long *cost_out;      // [2] or null
    // Cost-sorted, chunked execution of the adaptive time loops (run_ensemble, "Scheduling" below).  A launch covers
    // `n_slots` thread slots; slot s integrates instance perm[s] (identity when perm is null) for at most
    // `attempt_budget` attempts and then either finishes it (end) or PARKS it — the complete loop-carried state goes
    // to park_real / park_uint (+ the observer record) — so that a later launch (`sched_resume`) continues it
    // bit-identically on whatever lane the scheduler assigns.
    const unsigned int *perm;      // [n_slots] or null
    unsigned long long n_slots;    // 0 = n
    unsigned int attempt_budget;   // 0xffffffff: run to completion
    unsigned int sched_resume;     // 1: continue parked instances
    void *park_real;               // [2 N_VAR + NA_ + 3][n]: x, k1, aux, t, dt, trial step h
    unsigned int *park_uint;       // [2][n]: accepted steps so far; flags (bit 0 `clean`, bit 1 finished)
    // written by clode_sched_scan on the device, so that a whole schedule is enqueued without a host round trip:
    // [0] live instances (= slots of the next launch), [1] dearest non-empty bucket, [2] attempt budget of the next launch
    const unsigned int *sched_state; // or null: n_slots / attempt_budget above apply
};

extern "C" __constant__ KernelArgs clode_args;

CLODE_DEV SolverParams solver_params(const KernelArgs &a)
{
    SolverParams sp;
    sp.dt = (realtype)a.sp_dt; sp.dtmax = (realtype)a.sp_dtmax;
    sp.abstol = (realtype)a.sp_abstol; sp.reltol = (realtype)a.sp_reltol;
    sp.max_steps = a.sp_max_steps; sp.max_store = a.sp_max_store; sp.nout = a.sp_nout;
    return sp;
}
CLODE_DEV ObserverParams observer_params(const KernelArgs &a)
{
    ObserverParams op;
    op.eVarIx = E_VAR_IX; op.fVarIx = F_VAR_IX;
    op.maxEventCount = a.op_max_event_count; op.maxEventTimestamps = N_STORE_EVENTS;
    op.minXamp = (realtype)a.op_min_x_amp; op.minIMI = (realtype)a.op_min_imi;
    op.nHoodRadius = (realtype)a.op_nhood_radius;
    op.xUpThresh = (realtype)a.op_x_up; op.xDownThresh = (realtype)a.op_x_down;
    op.dxUpThresh = (realtype)a.op_dx_up; op.dxDownThresh = (realtype)a.op_dx_down;
    op.eps_dx = (realtype)a.op_eps_dx;
    return op;
}

// prologue shared by all kernels (transient.cl:28-52): coalesced loads, first noise draw, slope at t0
CLODE_DEV void load_instance(Instance &I, const KernelArgs &a, const size_t i)
{
    const size_t n = a.n;
    const realtype *x0 = (const realtype *)a.x0, *pars = (const realtype *)a.pars;
    I.t = (realtype)a.t0;
    I.dt = ((const realtype *)a.dt)[i];
#pragma unroll
    for (int j = 0; j < N_PAR; ++j)
        I.p[j] = __ldg(pars + (size_t)j * n + i);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        I.x[j] = x0[(size_t)j * n + i];
    I.rng.s0 = a.rng[i];
    I.rng.s1 = a.rng[n + i];
    I.rng.have_spare = false;
    I.rng.spare = ZERO;
#pragma unroll
    for (int j = 0; j < NA_; ++j)
        I.aux[j] = ZERO;
#pragma unroll
    for (int j = 0; j < NW_; ++j)
        I.w[j] = ZERO;
    draw_noise(I);
    getRHS(I.t, I.x, I.p, I.k1, I.aux, I.w);
}

// epilogue (transient.cl:64-76)
CLODE_DEV void store_instance(const Instance &I, const KernelArgs &a, const size_t i, const unsigned int steps)
{
    const size_t n = a.n;
    realtype *xf = (realtype *)a.xf;
#pragma unroll
    for (int j = 0; j < NV; ++j)
        xf[(size_t)j * n + i] = I.x[j];
    a.rng[i] = I.rng.s0;
    a.rng[n + i] = I.rng.s1;
    ((realtype *)a.dt)[i] = I.dt;
    ((realtype *)a.tf)[i] = I.t;
    if (a.steps) a.steps[i] = steps;
}

// ---- parked state of an instance between two launches of a chunked time loop ---------------------------------------
// Everything the attempt loop carries: x, the FSAL slope k1 (stored, not recomputed: a recomputed slope is a different
// inlined copy of getRHS and may contract differently), aux, t, dt, the trial step h, the accepted-step counter and the
// controller's `clean` flag.  Parameters are re-read from pars; the RNG state is not touched by adaptive steppers
// (only they are scheduled) and stays where it is.
#define PARK_X 0
#define PARK_K1 (NV)
#define PARK_AUX (2 * NV)
#define PARK_T (2 * NV + NA_)
#define PARK_DT (2 * NV + NA_ + 1)
#define PARK_H (2 * NV + NA_ + 2)
#define PARK_ROWS (2 * NV + NA_ + 3)
#define PARK_FLAG_CLEAN 1u
#define PARK_FLAG_FINISHED 2u

CLODE_DEV void park_instance(const Instance &I, const KernelArgs &a, const size_t i, const unsigned int step,
                             const realtype h, const bool clean)
{
    const size_t n = a.n;
    realtype *pr = (realtype *)a.park_real + i;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        pr[(size_t)(PARK_X + j) * n] = I.x[j];
        pr[(size_t)(PARK_K1 + j) * n] = I.k1[j];
    }
#pragma unroll
    for (int j = 0; j < NA_; ++j)
        pr[(size_t)(PARK_AUX + j) * n] = I.aux[j];
    pr[(size_t)PARK_T * n] = I.t;
    pr[(size_t)PARK_DT * n] = I.dt;
    pr[(size_t)PARK_H * n] = h;
    a.park_uint[i] = step;
    a.park_uint[n + i] = clean ? PARK_FLAG_CLEAN : 0u;
}

CLODE_DEV void unpark_instance(Instance &I, const KernelArgs &a, const size_t i, unsigned int &step, realtype &h, bool &clean)
{
    const size_t n = a.n;
    const realtype *pr = (const realtype *)a.park_real + i, *pars = (const realtype *)a.pars;
#pragma unroll
    for (int j = 0; j < N_PAR; ++j)
        I.p[j] = __ldg(pars + (size_t)j * n + i);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        I.x[j] = pr[(size_t)(PARK_X + j) * n];
        I.k1[j] = pr[(size_t)(PARK_K1 + j) * n];
    }
#pragma unroll
    for (int j = 0; j < NA_; ++j)
        I.aux[j] = pr[(size_t)(PARK_AUX + j) * n];
    I.t = pr[(size_t)PARK_T * n];
    I.dt = pr[(size_t)PARK_DT * n];
    h = pr[(size_t)PARK_H * n];
    I.rng.s0 = a.rng[i];
    I.rng.s1 = a.rng[n + i];
    I.rng.have_spare = false;
    I.rng.spare = ZERO;
#pragma unroll
    for (int j = 0; j < NW_; ++j)
        I.w[j] = ZERO;
    step = a.park_uint[i];
    clean = (a.park_uint[n + i] & PARK_FLAG_CLEAN) != 0u;
}

// an instance finished inside a scheduled launch: the scheduler drops it, and its accepted-step count is the cost
// estimate for a later pass over the same trajectory (warm-up -> features of the two-pass observers)
CLODE_DEV void mark_finished(const KernelArgs &a, const size_t i, const unsigned int step)
{
    if (a.park_uint) {
        a.park_uint[i] = step;
        a.park_uint[a.n + i] = PARK_FLAG_FINISHED;
    }
}

// advance by one ATTEMPT; true when an accepted (or abandoned, flag -1) step completed
#if !CLODE_ADAPTIVE
struct Controller {};
CLODE_DEV Controller make_controller(const SolverParams &, realtype) { return Controller(); }
CLODE_DEV realtype attempt_entry_step(const realtype dt, const SolverParams &) { return dt; }
#endif

CLODE_DEV bool advance(Instance &I, realtype &h, bool &clean, const SolverParams &sp, const Controller &ctl,
                       const realtype t_end)
{
#if CLODE_ADAPTIVE
    return adaptive_attempt(I, h, clean, sp, ctl, t_end);
#else
    step_fixed(I);
    return true;
#endif
}

// ------------------------------------------------------------------------------------------
// Ensemble drivers.  A "job" is one instance's life inside a kernel: begin(i) loads it,
// live() says whether another attempt is due, attempt() performs one, end(i) writes results.
//
//  * default: thread <-> instance, grid = ceil(n / block).  The attempt loop keeps the lanes of
//    a warp in step; a lane whose instance is finished idles until the warp's slowest lane is done.
//  * CLODE_WORK_QUEUE: persistent threads with PER-LANE REFILL.  The grid is sized to the
//    device (SMs x resident blocks); every warp keeps looping while any lane has work, and when
//    __ballot_sync shows at least CLODE_REFILL_LANES idle lanes (or the whole warp is idle) the idle
//    lanes take the next unprocessed instances from a global counter — one warp-aggregated atomicAdd
//    per refill — and load them while the busy lanes wait.  Results are indexed by instance id, so
//    they do not depend on which lane integrated what.  This pays when instances of one warp have very
//    different step counts (shuffled or randomly sampled parameter sets); for sorted grids the plain
//    mapping is already >96 % lane-efficient (profiles/).
#ifndef CLODE_REFILL_LANES
#define CLODE_REFILL_LANES 8
#endif

template <class Job> CLODE_DEV void run_ensemble(const KernelArgs &a, Job &job)
{
#ifndef CLODE_WORK_QUEUE
    bool reverse = a.block_order == 1u;
#ifndef __CUDACC_EMU__
    if (a.block_order == 2u) {
        const unsigned long long lower = a.cost_in[0], upper = a.cost_in[1];
        reverse = upper > lower + (lower >> 5); // 3 % hysteresis
    }
#endif
    const size_t chunk = reverse ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
    const size_t slot = chunk * (size_t)blockDim.x + threadIdx.x;
    const size_t n_slots = a.sched_state ? (size_t)a.sched_state[0] : (a.n_slots ? (size_t)a.n_slots : (size_t)a.n);
    if (slot >= n_slots) return;
    // Scheduling: the host sorts the unfinished instances by predicted remaining cost (clode_sched_* below) so that
    // the lanes of a warp, and the warps of a block, hold instances of similar cost and the dearest start first
    const size_t i = a.perm ? (size_t)a.perm[slot] : slot;
    if (a.sched_resume) job.resume(i);
    else job.begin(i);
    // the budget of a sorted round comes from the scan kernel unless the host asks for a run to completion
    unsigned int left = (a.sched_state && a.attempt_budget != 0xffffffffu) ? a.sched_state[2] : a.attempt_budget;
    while (job.live() && left != 0u) {
        job.attempt();
        --left;
    }
    if (job.live()) {
        job.park(i);
        return;
    }
    job.end(i);
    mark_finished(a, i, job.step);
#ifndef __CUDACC_EMU__
    if (a.cost_out) { // one atomic per warp and half
        const unsigned int mask = __activemask();
        const bool in_upper = 2 * i >= a.n;
        const unsigned int s = min(job.step, 1u << 26);
        const unsigned int lo = __reduce_add_sync(mask, in_upper ? 0u : s), hi = __reduce_add_sync(mask, in_upper ? s : 0u);
        if ((threadIdx.x & 31u) == (unsigned int)(__ffs(mask) - 1)) {
            if (lo) atomicAdd(a.cost_out, (unsigned long long)lo);
            if (hi) atomicAdd(a.cost_out + 1, (unsigned long long)hi);
        }
    }
#endif
#else
    const unsigned int FULL = 0xffffffffu;
    const unsigned int lane = threadIdx.x & 31u;
    size_t i = 0;
    bool have = false;    // this lane holds an unfinished instance
    bool drained = false; // the queue has been seen empty (warp-uniform)
    for (;;) {
        const unsigned int idle = __ballot_sync(FULL, !have);
        if (idle == FULL && drained) break;
        if (!drained && (idle == FULL || __popc(idle) >= CLODE_REFILL_LANES)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.queue, (unsigned long long)__popc(idle));
            base = __shfl_sync(FULL, base, 0);
            drained = base + __popc(idle) >= a.n;
            if (!have) {
                i = (size_t)(base + __popc(idle & ((1u << lane) - 1u)));
                if (i < a.n) {
                    job.begin(i);
                    have = true;
                }
            }
        }
        if (have) {
            if (job.live())
                job.attempt();
            if (!job.live()) {
                job.end(i);
                have = false;
            }
        }
    }
#endif
}

// ---- transient (clode/cpp/transient.cl:9-77) -------------------------------------------------
struct TransientJob {
    const KernelArgs &a;
    SolverParams sp;
    Controller ctl;
    realtype t_end;
    Instance I;
    unsigned int step;
    realtype h;
    bool clean;
    __device__ __forceinline__ TransientJob(const KernelArgs &a_) : a(a_), sp(solver_params(a_)), ctl(make_controller(sp, (realtype)a_.t1)), t_end((realtype)a_.t1) {}
    __device__ __forceinline__ void begin(size_t i) { load_instance(I, a, i); step = 0; h = attempt_entry_step(I.dt, sp); clean = true; }
    __device__ __forceinline__ bool live() const { return I.t <= t_end && step < sp.max_steps; }
    __device__ __forceinline__ void attempt() { if (advance(I, h, clean, sp, ctl, t_end)) ++step; }
    __device__ __forceinline__ void end(size_t i) { store_instance(I, a, i, step); }
    __device__ __forceinline__ void park(size_t i) { park_instance(I, a, i, step, h, clean); }
    __device__ __forceinline__ void resume(size_t i) { unpark_instance(I, a, i, step, h, clean); }
};

extern "C" __global__ void __launch_bounds__(CLODE_BLOCK, CLODE_MIN_BLOCKS_TRANSIENT)
clode_transient()
{
    CLODE_KERNEL_PROLOGUE();
    TransientJob job(clode_args);
    run_ensemble(clode_args, job);
}

// ---- Scheduling: cost-sorted instance order for the adaptive time loops ------------------------------------------
// Adaptive instances of one ensemble differ in cost by an order of magnitude (Lorenz r-sweep: 300 .. 6500 attempts;
// a bursting-model grid mixes silent, spiking and bursting cells).  A warp runs as long as its dearest lane and a launch
// as long as its last block, so WHICH instances share a warp and WHEN the dear ones start decide the lane and SM
// utilisation — on an unsorted parameter set half the lanes idle (C2 shuffled: 122 ms against 70 ms sorted).
// The runtime therefore runs every adaptive time loop as a short PILOT launch in the caller's order (a fixed
// attempt budget, nobody finishes, no divergence), then sorts the unfinished instances by predicted remaining cost,
//     remaining = accepted steps so far * (t_end - t) / (t - t0)        (average cost per unit of time so far),
// longest first, and continues them in launches of bounded attempt budget, re-sorting in between: lanes of a warp hold
// instances of equal predicted cost, the dearest blocks start first (longest-processing-time-first), a misprediction
// costs at most one budget, and the tail of a launch is short because its last blocks are its cheapest.  For the
// features pass of a two-pass observer the warm-up pass has just integrated the same trajectories, so its step counts
// are the exact costs (`by_steps`).  Results do not depend on the order: every instance's arithmetic is its own.
//
// The sort is a one-pass counting sort on a logarithmic key: bucket = the top 5 mantissa bits and the exponent of the
// cost as a float (2 % resolution, 1024 buckets, dearest first): histogram (shared-memory privatised), scan (one
// block), scatter (block-aggregated cursors).  Order inside a bucket is arbitrary.
#ifndef __CUDACC_EMU__
#define SCHED_BUCKETS 1024
#define SCHED_BLOCK 256

CLODE_DEV unsigned int sched_bucket_of(const float cost)
{
    // cost in [1, 2^31): biased exponent 127..157; NaN / Inf / huge -> dearest bucket, < 1 -> cheapest
    const int q = (int)(__float_as_uint(cost) >> 18) - (127 << 5);
    const int b = cost != cost ? SCHED_BUCKETS - 1 : (q < 0 ? 0 : (q > SCHED_BUCKETS - 1 ? SCHED_BUCKETS - 1 : q));
    return (unsigned int)(SCHED_BUCKETS - 1 - b); // bucket 0 = dearest
}

// pass 1: bucket of every live slot (0xffffffff: finished, dropped) + global histogram
extern "C" __global__ void __launch_bounds__(SCHED_BLOCK)
clode_sched_histogram(const unsigned int *perm_in, const unsigned int *slots_dev, const unsigned long long n,
                      const void *park_real, const unsigned int *park_uint, const double t0, const double t1,
                      const unsigned int by_steps, unsigned int *bucket, unsigned int *hist)
{
    const size_t n_slots = slots_dev ? (size_t)slots_dev[0] : (size_t)n; // slots of the order being re-sorted
    __shared__ unsigned int h[SCHED_BUCKETS];
    for (int k = threadIdx.x; k < SCHED_BUCKETS; k += SCHED_BLOCK) h[k] = 0u;
    __syncthreads();
    const size_t slot = blockIdx.x * (size_t)SCHED_BLOCK + threadIdx.x;
    if (slot < n_slots) {
        const size_t i = perm_in ? (size_t)perm_in[slot] : slot;
        unsigned int b = 0xffffffffu;
        const unsigned int steps = park_uint[i];
        if (by_steps) {
            b = sched_bucket_of((float)steps);
        } else if (!(park_uint[n + i] & PARK_FLAG_FINISHED)) {
            const realtype t = ((const realtype *)park_real)[(size_t)PARK_T * n + i];
            const float done = (float)((double)t - t0), left = (float)(t1 - (double)t);
            b = sched_bucket_of(done > 0.0f ? (float)steps * (left / done) : 3.0e38f);
        }
        bucket[slot] = b;
        if (b != 0xffffffffu) atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SCHED_BUCKETS; k += SCHED_BLOCK)
        if (h[k]) atomicAdd(&hist[k], h[k]);
}

// pass 2 (one block): exclusive scan of the histogram -> bucket cursors; result = {live instances, dearest non-empty
// bucket, attempt budget of the next launch} (KernelArgs::sched_state)
extern "C" __global__ void __launch_bounds__(SCHED_BUCKETS)
clode_sched_scan(const unsigned int *hist, unsigned int *cursor, unsigned int *result, const float budget_fraction,
                 const unsigned int budget_min)
{
    __shared__ unsigned int warp_sum[SCHED_BUCKETS / 32];
    __shared__ unsigned int first;
    const unsigned int k = threadIdx.x, lane = k & 31u, w = k >> 5;
    if (k == 0) first = SCHED_BUCKETS;
    const unsigned int v = hist[k];
    unsigned int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned int)d) incl += up;
    }
    if (lane == 31u) warp_sum[w] = incl;
    __syncthreads();
    if (w == 0) {
        unsigned int s = warp_sum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int up = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= (unsigned int)d) s += up;
        }
        warp_sum[lane] = s; // inclusive over warps
    }
    __syncthreads();
    const unsigned int before = (w ? warp_sum[w - 1] : 0u) + incl - v;
    cursor[k] = before;
    if (v) atomicMin(&first, k);
    __syncthreads();
    if (k == SCHED_BUCKETS - 1) {
        result[0] = before + v;
        result[1] = first;
        // attempt budget of the next launch: a fraction of the dearest live instance's predicted remaining cost
        // (upper edge of its bucket), so that mispredictions are re-sorted after at most that many attempts
        const unsigned int q = (unsigned int)(SCHED_BUCKETS - first) + (127u << 5); // bucket `first` spans [2^.., next edge)
        const float top = first < SCHED_BUCKETS ? __uint_as_float(q << 18) : 0.0f;
        const float want = fminf(top * budget_fraction, 1.0e9f);
        result[2] = max(budget_min, (unsigned int)want);
    }
}

// pass 3: instance indices into their bucket ranges
extern "C" __global__ void __launch_bounds__(SCHED_BLOCK)
clode_sched_scatter(const unsigned int *perm_in, const unsigned int *slots_dev, const unsigned long long n,
                    const unsigned int *bucket, unsigned int *cursor, unsigned int *perm_out)
{
    const size_t n_slots = slots_dev ? (size_t)slots_dev[0] : (size_t)n;
    __shared__ unsigned int h[SCHED_BUCKETS];
    for (int k = threadIdx.x; k < SCHED_BUCKETS; k += SCHED_BLOCK) h[k] = 0u;
    __syncthreads();
    const size_t slot = blockIdx.x * (size_t)SCHED_BLOCK + threadIdx.x;
    unsigned int b = 0xffffffffu, rank = 0u, inst = 0u;
    if (slot < n_slots) {
        b = bucket[slot];
        inst = perm_in ? perm_in[slot] : (unsigned int)slot;
        if (b != 0xffffffffu) rank = atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SCHED_BUCKETS; k += SCHED_BLOCK)
        if (h[k]) h[k] = atomicAdd(&cursor[k], h[k]); // count -> base of this block's range in bucket k
    __syncthreads();
    if (b != 0xffffffffu) perm_out[h[b] + rank] = inst;
}

// last step of the NVLink gather (clode_gather_rows): a shard's [rows][count] block, already on the root GPU, into the
// global array at instances first, first + stride, ... — variable-major [rows][n_total] (the API layout) or, for the
// Python front end's record arrays, instance-major [n_total][rows]
extern "C" __global__ void __launch_bounds__(256)
clode_interleave_rows(realtype *dst, const realtype *src, const unsigned long long rows, const unsigned long long count,
                      const unsigned long long n_total, const unsigned long long first, const unsigned long long stride,
                      const unsigned int instance_major)
{
    const size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= count) return;
    const size_t inst = first + j * stride;
    for (size_t r = blockIdx.y; r < rows; r += gridDim.y)
        dst[instance_major ? inst * rows + r : r * n_total + inst] = src[r * count + j];
}

// upload side of the same: records of `cols` reals per instance (the Python front end's (ensemble, nVar) matrices, moved to
// the device as they are) into the variable-major device layout dst[c * n + i]
extern "C" __global__ void __launch_bounds__(256)
clode_records_to_rows(realtype *dst, const double *src, const unsigned long long n, const unsigned int cols)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (unsigned int c = 0; c < cols; ++c)
        dst[(size_t)c * n + i] = (realtype)src[i * cols + c];
}

// The multi-GPU form of the same upload.  The host array was cut into G CONTIGUOUS chunks, chunk h moved as it is to GPU h
// (dense DMA, all GPUs at once); this kernel, running on the GPU that owns the interleaved shard g, g+G, ..., PULLS its
// records out of the chunks — peer loads over NVLink for the chunks that live on other GPUs — and transposes them into
// the variable-major buffer: the de-interleaving that would cost the host a strided pass over the whole array per shard
// happens on the GPUs.  src[h] = device pointer of chunk h (records of `cols` doubles), `chunk` = records per chunk.
struct PeerChunks { const double *src[16]; };
extern "C" __global__ void __launch_bounds__(256)
clode_records_pull(realtype *dst, const PeerChunks peers, const unsigned long long n_local, const unsigned int cols,
                   const unsigned long long first, const unsigned long long stride, const unsigned long long chunk)
{
    const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= n_local) return;
    const size_t i = first + k * stride; // global instance
    const size_t h = i / chunk;
    const double *rec = peers.src[h] + (i - h * chunk) * cols;
    for (unsigned int c = 0; c < cols; ++c)
        dst[(size_t)c * n_local + k] = (realtype)rec[c];
}
#endif // !__CUDACC_EMU__

#ifdef CLODE_WITH_FEATURES
// Where the observer state lives while an instance is being integrated.
//  default            : registers (best for the light observers: basic, basicall)
//  CLODE_OBS_SMEM     : a per-thread slot in dynamic shared memory.  The observer is touched once per ACCEPTED
//                       step, outside the RK stage evaluations, so shared-memory latency is irrelevant, while the
//                       fat observers (thresh2 / nhood1 / nhood2 / localmax: 60-90 reals for nVar = 4) otherwise
//                       push the kernel past 250 registers and into local-memory spills.  The slot size is an odd
//                       number of 8-byte words, which makes the thread-strided (array-of-structs) accesses of a
//                       warp bank-conflict free for the 64-bit fields.
#ifdef CLODE_OBS_SMEM
#define CLODE_OBS_SLOT_WORDS (((sizeof(Observer) + 7) / 8) | 1)
extern __shared__ double clode_dynamic_smem[];
CLODE_DEV Observer &observer_slot() { return *reinterpret_cast<Observer *>(clode_dynamic_smem + (size_t)threadIdx.x * CLODE_OBS_SLOT_WORDS); }
#define CLODE_OBSERVER_MEMBER Observer &ob
#define CLODE_OBSERVER_INIT , ob(observer_slot())
// The per-step observer work as an out-of-line call on the shared-memory slot.  Inlined, the compiler
// promotes the slot's fields back into registers for the whole time loop (no aliasing to stop it), which
// defeats the purpose; behind a call boundary the observer's registers exist only inside the call.
struct StepView {
    realtype t, x[NV], k1[NV], aux[NA_];
};
static __device__ __noinline__ bool observe_step(Observer *ob, const StepView v)
{
    const ObserverParams op = observer_params(clode_args);
    Instance J;
    J.t = v.t;
#pragma unroll
    for (int j = 0; j < NV; ++j) { J.x[j] = v.x[j]; J.k1[j] = v.k1[j]; }
#pragma unroll
    for (int j = 0; j < NA_; ++j) J.aux[j] = v.aux[j];
    ob->update(J, op);
    bool terminal = false;
    if (ob->event(J, op))
        terminal = ob->on_event(J, op);
    return terminal;
}
#else
#define CLODE_OBS_SLOT_WORDS 0
#define CLODE_OBSERVER_MEMBER Observer ob
#define CLODE_OBSERVER_INIT
#endif

// ---- initializeObserver (clode/cpp/initializeObserver.cl:9-83) -------------------------------
struct WarmupJob {
    const KernelArgs &a;
    SolverParams sp;
    ObserverParams op;
    Controller ctl;
    realtype t_end;
    Instance I;
    CLODE_OBSERVER_MEMBER;
    unsigned int step;
    realtype h;
    bool clean;
    __device__ __forceinline__ WarmupJob(const KernelArgs &a_)
        : a(a_), sp(solver_params(a_)), op(observer_params(a_)), ctl(make_controller(sp, (realtype)a_.t1)), t_end((realtype)a_.t1) CLODE_OBSERVER_INIT {}
    __device__ __forceinline__ void begin(size_t i) { load_instance(I, a, i); ob.init(I); step = 0; h = attempt_entry_step(I.dt, sp); clean = true; }
    // strict '<' (initializeObserver.cl:62); one-pass observers do no warm-up integration at all
    __device__ __forceinline__ bool live() const { return CLODE_TWO_PASS && I.t < t_end && step < sp.max_steps; }
    __device__ __forceinline__ void attempt()
    {
        if (advance(I, h, clean, sp, ctl, t_end)) {
            ++step;
            ob.warmup(I, op);
        }
    }
    __device__ __forceinline__ void end(size_t i)
    {
#if CLODE_TWO_PASS
        // rewind; dt and the RNG state are NOT written back (initializeObserver.cl:72-82)
        I.t = (realtype)a.t0;
        const realtype *x0 = (const realtype *)a.x0;
#pragma unroll
        for (int j = 0; j < NV; ++j)
            I.x[j] = x0[(size_t)j * a.n + i];
        getRHS(I.t, I.x, I.p, I.k1, I.aux, I.w);
#endif
        ob.arm(I, op);
        ObsStore st = {(realtype *)a.od_real, a.od_uint, (size_t)a.n, i, 0, 0};
        ob.visit(st);
    }
    // A parked warm-up carries only what the warm-up accumulates (visit_warmup: a few extents), in the first rows of
    // the observer record, which end() overwrites with the complete record anyway.  Everything else in the observer is
    // what init() derives from the initial state, so resume() re-derives it from x0 exactly as begin() does — loading
    // the whole record instead would keep 60-90 reals alive through the warm-up loop (+600 B of spills for thresh2).
    __device__ __forceinline__ void park(size_t i)
    {
        park_instance(I, a, i, step, h, clean);
        ObsStore st = {(realtype *)a.od_real, a.od_uint, (size_t)a.n, i, 0, 0};
        ob.visit_warmup(st);
    }
    __device__ __forceinline__ void resume(size_t i)
    {
        load_instance(I, a, i);
        ob.init(I);
        unpark_instance(I, a, i, step, h, clean);
        ObsLoad ld = {(const realtype *)a.od_real, a.od_uint, (size_t)a.n, i, 0, 0};
        ob.visit_warmup(ld);
    }
};

extern "C" __global__ void __launch_bounds__(CLODE_BLOCK, CLODE_MIN_BLOCKS_INIT)
clode_initialize_observer()
{
    CLODE_KERNEL_PROLOGUE();
    WarmupJob job(clode_args);
    run_ensemble(clode_args, job);
}

// ---- features (clode/cpp/features.cl:10-106) ---------------------------------------------------
struct FeaturesJob {
    const KernelArgs &a;
    SolverParams sp;
    ObserverParams op;
    Controller ctl;
    realtype t_end;
    Instance I;
    CLODE_OBSERVER_MEMBER;
    unsigned int step;
    realtype h;
    bool clean, alive;
    __device__ __forceinline__ FeaturesJob(const KernelArgs &a_)
        : a(a_), sp(solver_params(a_)), op(observer_params(a_)), ctl(make_controller(sp, (realtype)a_.t1)), t_end((realtype)a_.t1) CLODE_OBSERVER_INIT {}
    __device__ __forceinline__ void begin(size_t i)
    {
        load_instance(I, a, i);
        ObsLoad ld = {(const realtype *)a.od_real, a.od_uint, (size_t)a.n, i, 0, 0};
        ob.visit(ld);
        ob.open_means();
        step = 0; h = attempt_entry_step(I.dt, sp); clean = true;
        alive = I.t <= t_end && step < sp.max_steps;
    }
    __device__ __forceinline__ bool live() const { return alive; }
    __device__ __forceinline__ void attempt()
    {
        if (advance(I, h, clean, sp, ctl, t_end)) {
            ++step;
            // features.cl:71-81: update, then event test, then event features (a terminal event ends the run)
#ifdef CLODE_OBS_SMEM
            StepView v;
            v.t = I.t;
#pragma unroll
            for (int j = 0; j < NV; ++j) { v.x[j] = I.x[j]; v.k1[j] = I.k1[j]; }
#pragma unroll
            for (int j = 0; j < NA_; ++j) v.aux[j] = I.aux[j];
            const bool terminal = observe_step(&ob, v);
#else
            ob.update(I, op);
            bool terminal = false;
            if (ob.event(I, op))
                terminal = ob.on_event(I, op);
#endif
            alive = !terminal && I.t <= t_end && step < sp.max_steps;
        }
    }
    __device__ __forceinline__ void end(size_t i)
    {
        FeatureOut out = {(realtype *)a.F, (size_t)a.n, i, 0};
        ob.close_means();
        ob.emit(out);
        ob.rebase(I.t - (realtype)a.t0);
        ObsStore st = {(realtype *)a.od_real, a.od_uint, (size_t)a.n, i, 0, 0};
        ob.visit(st);
        store_instance(I, a, i, step);
    }
    // parked as it is in the loop: the time-weighted means stay in their in-kernel (integral) form, no re-basing
    __device__ __forceinline__ void park(size_t i)
    {
        park_instance(I, a, i, step, h, clean);
        ObsStore st = {(realtype *)a.od_real, a.od_uint, (size_t)a.n, i, 0, 0};
        ob.visit(st);
    }
    __device__ __forceinline__ void resume(size_t i)
    {
        unpark_instance(I, a, i, step, h, clean);
        ObsLoad ld = {(const realtype *)a.od_real, a.od_uint, (size_t)a.n, i, 0, 0};
        ob.visit(ld);
        alive = true; // it was parked because it was
    }
};

extern "C" __global__ void __launch_bounds__(CLODE_BLOCK, CLODE_MIN_BLOCKS_FEATURES)
clode_features()
{
    CLODE_KERNEL_PROLOGUE();
    FeaturesJob job(clode_args);
    run_ensemble(clode_args, job);
}

// number of observer-state rows, for the host allocator
extern "C" __global__ void clode_observer_layout(int *out)
{
    Observer ob;
    ObsCount c = {0, 0};
    ob.visit(c);
    out[0] = c.nreal;
    out[1] = c.nuint;
    out[2] = CLODE_TWO_PASS;
    out[3] = (int)(CLODE_OBS_SLOT_WORDS * 8); // bytes of dynamic shared memory per thread (0: observer in registers)
}
#endif // CLODE_WITH_FEATURES

#ifdef CLODE_WITH_TRAJECTORY
// ---- trajectory (clode/cpp/trajectory.cl:14-112) -----------------------------------------------
// Row r of the outputs holds stored point r of every instance:
//   t[r*n + i], x[(r*N_VAR + j)*n + i], dx[...], aux[(r*N_AUX + j)*n + i].
// A warp therefore writes one contiguous 32*sizeof(realtype) segment per variable per row.
// The host allocates max_store+1 rows: row index max_store can be written.
CLODE_DEV void store_point(const Instance &I, const KernelArgs &a, const size_t i, const size_t row)
{
    const size_t n = a.n;
    __stcs((realtype *)a.tr_t + row * n + i, I.t);
    realtype *x = (realtype *)a.tr_x + row * n * NV + i;
    realtype *dx = (realtype *)a.tr_dx + row * n * NV + i;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        __stcs(x + (size_t)j * n, I.x[j]);
        __stcs(dx + (size_t)j * n, I.k1[j]);
    }
#if N_AUX > 0
    realtype *aux = (realtype *)a.tr_aux + row * n * N_AUX + i;
#pragma unroll
    for (int j = 0; j < N_AUX; ++j)
        __stcs(aux + (size_t)j * n, I.aux[j]);
#endif
}

// state of a chunked trajectory between two launches: x, t, dt and the RNG words are where store_instance puts
// them; the polar method's cached variate, the noise already drawn for the next step and the two counters go to
// rs_real / rs_uint.  The slope and the aux variables are recomputed (same inputs, same getRHS).
CLODE_DEV void suspend_instance(const Instance &I, const KernelArgs &a, const size_t i, const unsigned int step,
                                const unsigned int row, const bool unfinished)
{
    const size_t n = a.n;
    realtype *rr = (realtype *)a.rs_real;
    rr[i] = I.rng.spare;
#pragma unroll
    for (int j = 0; j < N_WIENER; ++j)
        rr[(size_t)(1 + j) * n + i] = I.w[j];
    a.rs_uint[i] = step;
    a.rs_uint[n + i] = row;
    a.rs_uint[2 * n + i] = I.rng.have_spare ? 1u : 0u;
    if (unfinished) a.chunk_flags[0] = 1u; // benign race: every writer stores the same value
    if (row >= a.row_begin) atomicMax(a.chunk_flags + 1, row);
}

CLODE_DEV void resume_instance(Instance &I, const KernelArgs &a, const size_t i, unsigned int &step, unsigned int &row)
{
    const size_t n = a.n;
    const realtype *xf = (const realtype *)a.xf, *pars = (const realtype *)a.pars, *rr = (const realtype *)a.rs_real;
    I.t = ((const realtype *)a.tf)[i];
    I.dt = ((const realtype *)a.dt)[i];
#pragma unroll
    for (int j = 0; j < N_PAR; ++j)
        I.p[j] = __ldg(pars + (size_t)j * n + i);
#pragma unroll
    for (int j = 0; j < NV; ++j)
        I.x[j] = xf[(size_t)j * n + i];
    I.rng.s0 = a.rng[i];
    I.rng.s1 = a.rng[n + i];
    I.rng.have_spare = a.rs_uint[2 * n + i] != 0u;
    I.rng.spare = rr[i];
#pragma unroll
    for (int j = 0; j < NA_; ++j)
        I.aux[j] = ZERO;
#pragma unroll
    for (int j = 0; j < NW_; ++j)
        I.w[j] = ZERO;
#pragma unroll
    for (int j = 0; j < N_WIENER; ++j)
        I.w[j] = rr[(size_t)(1 + j) * n + i];
    step = a.rs_uint[i];
    row = a.rs_uint[n + i];
    getRHS(I.t, I.x, I.p, I.k1, I.aux, I.w);
}

struct TrajectoryJob {
    const KernelArgs &a;
    SolverParams sp;
    Controller ctl;
    realtype t_end;
    Instance I;
    size_t inst;
    unsigned int step, row;
    realtype h;
    bool clean;
    __device__ __forceinline__ TrajectoryJob(const KernelArgs &a_) : a(a_), sp(solver_params(a_)), ctl(make_controller(sp, (realtype)a_.t1)), t_end((realtype)a_.t1) {}
    __device__ __forceinline__ void begin(size_t i)
    {
        inst = i; clean = true;
        if (!a.resume) {
            load_instance(I, a, i);
            step = 0; row = 0;
            store_point(I, a, i, 0);
        } else {
            resume_instance(I, a, i, step, row);
        }
        h = attempt_entry_step(I.dt, sp);
    }
    // trajectory.cl:76; `row + 1 < row_end`: the next point still belongs to this launch's rows
    __device__ __forceinline__ bool unfinished() const { return I.t <= t_end && step < sp.max_steps && row < sp.max_store; }
    __device__ __forceinline__ bool live() const { return unfinished() && row + 1 < a.row_end; }
    __device__ __forceinline__ void attempt()
    {
        if (advance(I, h, clean, sp, ctl, t_end)) {
            ++step;
            if (step % sp.nout == 0) {
                ++row;
                store_point(I, a, inst, row - a.row_begin);
            }
        }
    }
    __device__ __forceinline__ void end(size_t i)
    {
        a.n_stored[i] = (int)row;
        store_instance(I, a, i, step);
        if (a.rs_uint) suspend_instance(I, a, i, step, row, unfinished());
    }
    // trajectories are chunked by stored rows (suspend_instance / resume_instance), not by the attempt scheduler:
    // a permuted instance order would turn the coalesced row stores into scattered ones
    __device__ __forceinline__ void park(size_t i) { end(i); }
    __device__ __forceinline__ void resume(size_t i) { begin(i); }
};

#if defined(CLODE_TRAJ_STAGED) && !CLODE_ADAPTIVE && !defined(CLODE_WORK_QUEUE)
// ---- shared-memory staged stores (fixed-step methods) ------------------------------------------
// With a fixed-step method every instance of a block reaches stored row r in the same loop iteration,
// so the block's values for one (row, variable) line are CLODE_BLOCK consecutive reals in global memory.
// The threads write their values into a double-buffered shared-memory tile [lines][CLODE_BLOCK]; one
// thread then hands each line to the TMA as a bulk shared->global copy (cp.async.bulk, `UBLKCP` in SASS),
// which drains asynchronously while the block integrates the next step: the store traffic leaves the
// LSU / issue path of the compute warps entirely.  Lanes that finished earlier leave stale values in
// their columns; those land in rows beyond their own nStored, which the API never returns.
#define TRAJ_LINES (1 + 2 * N_VAR + N_AUX)

CLODE_DEV void tile_put(realtype (*tile)[CLODE_BLOCK], const Instance &I)
{
    const unsigned int c = threadIdx.x;
    tile[0][c] = I.t;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        tile[1 + j][c] = I.x[j];
        tile[1 + NV + j][c] = I.k1[j];
    }
#pragma unroll
    for (int j = 0; j < N_AUX; ++j)
        tile[1 + 2 * NV + j][c] = I.aux[j];
}

// one thread: bulk-copy every line of the tile to row `row` of the outputs, `cols` valid columns
CLODE_DEV void tile_flush(realtype (*tile)[CLODE_BLOCK], const KernelArgs &a, const size_t base, const size_t row,
                          const unsigned int cols)
{
    const size_t n = a.n;
    const unsigned int bytes = cols * (unsigned int)sizeof(realtype);
#pragma unroll
    for (int line = 0; line < TRAJ_LINES; ++line) {
        realtype *g;
        if (line == 0) g = (realtype *)a.tr_t + row * n + base;
        else if (line < 1 + NV) g = (realtype *)a.tr_x + (row * NV + (line - 1)) * n + base;
        else if (line < 1 + 2 * NV) g = (realtype *)a.tr_dx + (row * NV + (line - 1 - NV)) * n + base;
        else g = (realtype *)a.tr_aux + (row * N_AUX + (line - 1 - 2 * NV)) * n + base;
        const unsigned int src = (unsigned int)__cvta_generic_to_shared(&tile[line][0]);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

extern "C" __global__ void __launch_bounds__(CLODE_BLOCK, CLODE_MIN_BLOCKS_TRAJECTORY)
clode_trajectory()
{
    CLODE_KERNEL_PROLOGUE();
    const KernelArgs &a = clode_args;
    __shared__ __align__(128) realtype tile[2][TRAJ_LINES][CLODE_BLOCK];
    const size_t base = blockIdx.x * (size_t)blockDim.x;
    const size_t i = base + threadIdx.x;
    const bool valid = i < a.n;
    const unsigned int cols = (unsigned int)(a.n - base < (size_t)CLODE_BLOCK ? a.n - base : (size_t)CLODE_BLOCK);
    // bulk copies need 16-byte aligned addresses and sizes: even row pitch and an even column count
    const bool bulk_ok = (a.n * sizeof(realtype)) % 16 == 0 && (cols * sizeof(realtype)) % 16 == 0;
    TrajectoryJob job(a);
    if (!bulk_ok) { // rare shapes: plain per-thread stores
        if (valid) {
            job.begin(i);
            while (job.live()) job.attempt();
            job.end(i);
        }
        return;
    }
    if (valid) {
        load_instance(job.I, a, i);
        job.inst = i; job.step = 0; job.row = 0; job.h = job.I.dt; job.clean = true;
        tile_put(tile[0], job.I);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) tile_flush(tile[0], a, base, 0, cols);
    int buf = 1;
    for (unsigned int k = 1;; ++k) {
        const bool live = valid && job.live();
        if (!__syncthreads_or(live)) break;
        if (live) {
            step_fixed(job.I);
            ++job.step;
        }
        if (k % job.sp.nout == 0) { // block-uniform: every live lane stores row k / nout now
            // the buffer about to be overwritten was handed to the TMA two stores ago: wait until it has been read
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
            if (live) {
                ++job.row;
                tile_put(tile[buf], job.I);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) tile_flush(tile[buf], a, base, (size_t)(k / job.sp.nout), cols);
            buf ^= 1;
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (valid) job.end(i);
}
#else
extern "C" __global__ void __launch_bounds__(CLODE_BLOCK, CLODE_MIN_BLOCKS_TRAJECTORY)
clode_trajectory()
{
    CLODE_KERNEL_PROLOGUE();
    TrajectoryJob job(clode_args);
    run_ensemble(clode_args, job);
}
#endif
#endif // CLODE_WITH_TRAJECTORY

#endif // CLODE_KERNELS_CUH
