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

// advance by one ATTEMPT; true when an accepted (or abandoned, flag -1) step completed
#if !CLODE_ADAPTIVE
struct Controller {};
CLODE_DEV Controller make_controller(const SolverParams &, realtype) { return Controller(); }
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
    const size_t i = chunk * (size_t)blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    job.begin(i);
    while (job.live())
        job.attempt();
    job.end(i);
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
    __device__ __forceinline__ void begin(size_t i) { load_instance(I, a, i); step = 0; h = I.dt; clean = true; }
    __device__ __forceinline__ bool live() const { return I.t <= t_end && step < sp.max_steps; }
    __device__ __forceinline__ void attempt() { if (advance(I, h, clean, sp, ctl, t_end)) ++step; }
    __device__ __forceinline__ void end(size_t i) { store_instance(I, a, i, step); }
};

extern "C" __global__ void __launch_bounds__(CLODE_BLOCK, CLODE_MIN_BLOCKS_TRANSIENT)
clode_transient()
{
    CLODE_KERNEL_PROLOGUE();
    TransientJob job(clode_args);
    run_ensemble(clode_args, job);
}

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
    __device__ __forceinline__ void begin(size_t i) { load_instance(I, a, i); ob.init(I); step = 0; h = I.dt; clean = true; }
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
        step = 0; h = I.dt; clean = true;
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
        h = I.dt;
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
