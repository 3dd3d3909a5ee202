/* clode_rt.h — C ABI of the B200 ensemble-ODE runtime (libclode_rt.so).
 *
 * This is the drop-in boundary below clODE's C++ classes: it replaces what the reference
 * obtains from `OpenCLResource` + `cl::Buffer` + `cl::Kernel` (clode/cpp/OpenCLResource.hpp:73-144,
 * clode/cpp/CLODE.hpp:131-141, clode/cpp/CLODEfeatures.hpp:41-43, clode/cpp/CLODEtrajectory.hpp:29-30)
 * with a CUDA driver-API + NVRTC implementation for sm_100a.  Signatures are plain C:
 * opaque handles, POD structs, host pointers and element counts.  No torch, no C++ types.
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero `clode_status` on failure;
 *     `clode_last_error()` then returns a human-readable message (thread-local),
 *     including the NVRTC build log for build failures (cf. OpenCLResource.cpp:280-295);
 *   - ensemble arrays are flat, variable-major: x0[j*nPts + i], pars[j*nPts + i],
 *     F[k*nPts + i], t[s*nPts + i], x[(s*nVar + j)*nPts + i]  (SURVEY.md §8b);
 *   - host arrays are always double (the reference API's `std::vector<cl_double>`); with
 *     single_precision the runtime narrows/widens on transfer like the reference does
 *     (clode/cpp/CLODE.cpp:292-300, 534-549);
 *   - all calls are synchronous with respect to the caller, like the reference's
 *     enqueue + finish pairs (clode/cpp/CLODE.cpp:484-485).
 */
#ifndef CLODE_RT_H
#define CLODE_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CLODE_API __attribute__((visibility("default")))
#else
#define CLODE_API
#endif

typedef enum clode_status {
    CLODE_OK = 0,
    CLODE_ERR_INVALID = 1,   /* bad argument / wrong size / unknown name        */
    CLODE_ERR_NO_DRIVER = 2, /* libcuda / libnvrtc could not be loaded, no GPU  */
    CLODE_ERR_CUDA = 3,      /* a CUDA driver call failed                       */
    CLODE_ERR_BUILD = 4,     /* NVRTC compilation failed (see clode_last_error) */
    CLODE_ERR_STATE = 5,     /* call sequence error (e.g. run before build)     */
    CLODE_ERR_MEMORY = 6     /* allocation larger than the device allows        */
} clode_status;

CLODE_API const char *clode_last_error(void);
CLODE_API const char *clode_version(void);

/* ---- devices: replaces queryOpenCL / OpenCLResource device selection --------------------
 * (clode/cpp/OpenCLResource.hpp:48-71 `deviceInfo`, :131-144) */
typedef struct clode_device_info {
    char name[256];
    int cc_major, cc_minor;
    int multiprocessors;        /* -> deviceInfo.computeUnits     */
    int clock_mhz;              /* -> deviceInfo.maxClock         */
    int max_threads_per_block;  /* -> deviceInfo.maxWorkGroupSize */
    uint64_t total_memory;      /* -> deviceInfo.deviceMemSize    */
    uint64_t max_alloc;         /* -> deviceInfo.maxMemAllocSize  */
    int driver_version;
} clode_device_info;

CLODE_API int clode_device_count(int *count);
CLODE_API int clode_device_get_info(int device, clode_device_info *info);
/* Roofline denominator for the transient/features kernels: measured FP64 FMA throughput of the
 * device (register-resident DFMA chains, best of `repeats` launches, CUDA events), in TFLOP/s
 * counting an FMA as 2 flops.  MEASURED_PEAKS.json has no FP64 entry. */
CLODE_API int clode_measure_fp64_peak(int device, int repeats, double *tflops, double *ms_best);

/* ---- program: replaces CLODE::setCLbuildOpts / buildProgram / cl::Kernel creation -------
 * (clode/cpp/CLODE.cpp:109-172, clode/cpp/CLODEfeatures.cpp:34-61, clode/cpp/CLODEtrajectory.cpp:25-43).
 * One program = user RHS x stepper x observer x precision x dimensions, JIT-specialised. */
enum { CLODE_KERNEL_TRANSIENT = 1, CLODE_KERNEL_FEATURES = 2, CLODE_KERNEL_TRAJECTORY = 4 };

typedef struct clode_program_desc {
    const char *rhs_source;   /* text of the getRHS translation unit (ProblemInfo.clRHSfilename contents) */
    const char *stepper;      /* "euler" "heun" "rk4" "bs23" "dopri5" "seuler" (steppers.cl:29-34)        */
    const char *observer;     /* "basic" "basicall" "localmax" "nhood1" "nhood2" "thresh2"; NULL = basic  */
    int single_precision;     /* CLODE_SINGLE_PRECISION vs CLODE_DOUBLE_PRECISION                         */
    int n_var, n_par, n_aux, n_wiener; /* N_VAR, N_PAR, N_AUX, N_WIENER                                   */
    int f_var_ix, e_var_ix;   /* ObserverParams.fVarIx / eVarIx, baked as constants                      */
    int n_store_events;       /* N_STORE_EVENTS = ObserverParams.maxEventTimestamps                      */
    int kernels;              /* bit mask of CLODE_KERNEL_*; transient is always built                   */
    int bit_exact;            /* 1: portable transcendental math + no FMA contraction (parity tier)     */
    int work_queue;           /* 1: persistent threads pulling instances from a global queue            */
    int block_size;           /* threads per block; 0 = default                                          */
    int min_blocks_per_sm;    /* __launch_bounds__ second argument; 0 = chosen by spill check            */
    int staged_trajectory;    /* 1: trajectory rows staged in shared memory and written by TMA bulk copies
                                 (fixed-step methods); 0: per-thread coalesced stores                     */
    int observer_in_shared;   /* 1: observer state in a per-thread shared-memory slot instead of registers
                                 (for the fat observers); 0: registers                                    */
    int ieee_constant_division; /* 0 (default, ignored when bit_exact): `x / literal` in the program's PTX is
                                 evaluated as q = x*y, q + (x - c*q)*y with y = RN(1/c) from the host — the
                                 correctly rounded quotient for |x| in [2^-511, 2^512), one ulp outside —
                                 instead of ptxas' generic sequence, which Newton-refines the literal's
                                 reciprocal at run time; and (double precision) `1 / x`, `a / x` with a variable x
                                 are ptxas' own fast-path sequence with selects instead of the branch to its
                                 slow path (bit-identical to IEEE for normal operands and results, zeros,
                                 infinities and NaNs; subnormal or >= 2^1022 divisors are flushed; environment
                                 CLODE_BRANCHLESS=0 keeps ptxas' expansion); 1: leave every division to ptxas */
    int library_exp;          /* 0 (default): in production double builds exp() is the engine's table + polynomial
                                 (device/fast_exp.cuh: branch-free, 2048-entry table + cubic, <= 1.06 ulp, 9 FP64
                                 instructions; with CLODE_BRANCHLESS=0 the 128-entry hi/lo table + degree 5,
                                 <= 0.52 ulp, 11 instructions and a branch); 1: CUDA's libdevice exp (1 ulp,
                                 15 FP64 instructions + coefficient moves).  Ignored by bit_exact and
                                 single-precision builds                                                    */
} clode_program_desc;

/* compile only (no GPU needed): returns malloc'd cubin + log; caller frees with clode_free */
CLODE_API int clode_compile(const clode_program_desc *desc, void **cubin, size_t *cubin_size, char **log);
/* the full generated CUDA source for a description (cf. CLODE::getProgramString, CLODE.hpp:187) */
CLODE_API int clode_program_source(const clode_program_desc *desc, char **source);
CLODE_API void clode_free(void *p);

/* ---- solver / observer parameters (always double on the host side) -----------------------
 * mirrors of SolverParams<cl_double> (clode/cpp/clODE_struct_defs.cl:11-20) and
 * ObserverParams<cl_double> (clode/cpp/observers.cl:25-46) */
typedef struct clode_solver_params {
    double dt, dtmax, abstol, reltol;
    unsigned int max_steps, max_store, nout;
} clode_solver_params;

typedef struct clode_observer_params {
    unsigned int e_var_ix, f_var_ix, max_event_count, max_event_timestamps;
    double min_x_amp, min_imi, nhood_radius, x_up_thresh, x_down_thresh, dx_up_thresh, dx_down_thresh, eps_dx;
} clode_observer_params;

/* ---- simulation object: device buffers + kernels of one CLODE* instance on one GPU -------- */
typedef struct clode_sim clode_sim;

CLODE_API int clode_sim_create(int device, clode_sim **out);
CLODE_API int clode_sim_destroy(clode_sim *sim);

/* CLODE::buildCL / CLODEfeatures::buildCL / CLODEtrajectory::buildCL */
CLODE_API int clode_sim_build(clode_sim *sim, const clode_program_desc *desc);
CLODE_API int clode_sim_build_log(clode_sim *sim, const char **log);

/* CLODE::setNpts (CLODE.cpp:174-243): (re)allocate per-instance buffers; dt is filled with fill_dt.
 * RNG state is left untouched — seed it with clode_sim_seed_rng / clode_sim_set_rng_state. */
CLODE_API int clode_sim_set_npts(clode_sim *sim, size_t n_pts, double fill_dt);
CLODE_API int clode_sim_get_npts(clode_sim *sim, size_t *n_pts);

CLODE_API int clode_sim_set_x0(clode_sim *sim, const double *x0, size_t count);      /* CLODE::setX0   */
CLODE_API int clode_sim_set_pars(clode_sim *sim, const double *pars, size_t count);  /* CLODE::setPars */
CLODE_API int clode_sim_set_dt(clode_sim *sim, const double *dt, size_t count);      /* per-instance dt (continuation) */
CLODE_API int clode_sim_set_tspan(clode_sim *sim, double t0, double t1);             /* CLODE::setTspan */
CLODE_API int clode_sim_set_solver_params(clode_sim *sim, const clode_solver_params *sp);     /* CLODE::setSolverParams */
CLODE_API int clode_sim_set_observer_params(clode_sim *sim, const clode_observer_params *op); /* CLODEfeatures::setObserverParams */

/* CLODE::seedRNG(cl_int) (CLODE.cpp:447-465): state word k of the GLOBAL ensemble is seed+k, instance i
 * owns words i and n_global+i.  A shard holding instances [offset, offset+nPts) of a larger ensemble
 * passes its offset and the global size so that sharded and unsharded runs draw identical streams. */
CLODE_API int clode_sim_seed_rng(clode_sim *sim, int64_t seed, uint64_t offset, uint64_t n_global);
CLODE_API int clode_sim_set_rng_state(clode_sim *sim, const uint64_t *state, size_t count); /* [2][nPts] */
CLODE_API int clode_sim_get_rng_state(clode_sim *sim, uint64_t *state, size_t count);

/* simulation routines */
CLODE_API int clode_sim_transient(clode_sim *sim);            /* CLODE::transient (CLODE.cpp:468-493)          */
CLODE_API int clode_sim_initialize_observer(clode_sim *sim);  /* CLODEfeatures::initializeObserver (:180-209)   */
CLODE_API int clode_sim_features(clode_sim *sim, int initialize); /* CLODEfeatures::features; initialize: 1 force, 0 continue, -1 only if needed */
CLODE_API int clode_sim_observer_initialized(clode_sim *sim, int *flag);
CLODE_API int clode_sim_trajectory(clode_sim *sim);           /* CLODEtrajectory::trajectory (:97-130)          */
CLODE_API int clode_sim_shift_x0(clode_sim *sim);             /* CLODE::shiftX0 (CLODE.cpp:502-515), device to device */

/* Streamed trajectory: the same integration and the same results as clode_sim_trajectory followed by
 * clode_sim_get(CLODE_BUF_T / _X / _DX / _AUX) and clode_sim_get_n_stored, but cut into launches of `chunk_rows`
 * stored points.  The device holds two chunks instead of all max_store rows, and the copy of chunk k to the host
 * overlaps the integration of chunk k+1.  Replaces the single nPts*max_store allocation of
 * CLODEtrajectory::resizeTrajectoryVariables (CLODEtrajectory.cpp:45-95; "TODO" at :47 and clode/trajectory.py:166)
 * and the four copies of CLODEtrajectory::getT/getX/getDx/getAux (:132-205).
 * Host arrays are full size and may be NULL to skip an output: t[max_store][n], x[max_store][nVar][n],
 * dx[max_store][nVar][n], aux[max_store][nAux][n], n_stored[n]; rows at or beyond an instance's n_stored read 0.
 * Arrays from clode_host_alloc are copied at the full PCIe rate; any other host memory works, slower. */
CLODE_API int clode_sim_trajectory_stream(clode_sim *sim, size_t chunk_rows, double *t, double *x, double *dx,
                                          double *aux, int *n_stored);
CLODE_API void *clode_host_alloc(int device, size_t bytes); /* page-locked host memory; NULL on failure */
CLODE_API void clode_host_free(void *p);

/* Non-blocking variants, for callers that drive several GPUs from one thread (one clode_sim per
 * device, the reference's unused `OpenCLResource(platformID, std::vector<deviceIDs>)` hook,
 * OpenCLResource.hpp:104): enqueue on every shard, then wait on every shard.  `kernel` is one of
 * CLODE_KERNEL_*; `initialize` as in clode_sim_features. */
CLODE_API int clode_sim_enqueue(clode_sim *sim, int kernel, int initialize);
CLODE_API int clode_sim_wait(clode_sim *sim);

/* results: which = one of CLODE_BUF_*; out has `count` doubles (checked) */
enum {
    CLODE_BUF_X0 = 0, CLODE_BUF_PARS = 1, CLODE_BUF_XF = 2, CLODE_BUF_DT = 3, CLODE_BUF_TF = 4,
    CLODE_BUF_F = 5, CLODE_BUF_T = 6, CLODE_BUF_X = 7, CLODE_BUF_DX = 8, CLODE_BUF_AUX = 9,
    CLODE_BUF_RNG = 10, CLODE_BUF_STEPS = 11, CLODE_BUF_NSTORED = 12
};
CLODE_API int clode_sim_get(clode_sim *sim, int which, double *out, size_t count);
CLODE_API int clode_sim_get_n_stored(clode_sim *sim, int *out, size_t count);        /* CLODEtrajectory::getNstored */
CLODE_API int clode_sim_get_steps(clode_sim *sim, uint32_t *out, size_t count);      /* accepted steps of the last call */
CLODE_API int clode_sim_n_features(clode_sim *sim, int *n_features);

/* Strided host access, for callers that keep ONE host array for an ensemble spread over several GPUs (the reference's
 * unused multi-device hook, OpenCLResource.hpp:104): the simulation object holds the instances first, first+stride, ...
 * of a global ensemble whose host arrays are [rows][host_pitch].  set_rows gathers that column set from `host` into the
 * device buffer `which` ([rows][nPts] on the device), get_rows scatters the device buffer back.  Both go through a
 * page-locked staging ring inside the runtime: the gather / scatter (and the float<->double conversion of single-
 * precision programs) happens in the same pass that fills the ring, while the previous chunk is on the wire.
 * first = 0, stride = 1, host_pitch = nPts is the plain whole-array transfer (CLODE::setX0, CLODE::getXf, ...).
 * Element (row r, local column k) lives at host[r*host_pitch + first + k*stride]; nothing else is assumed, so an
 * instance-major matrix a[i*nVar + j] (the Python front end's (ensemble, nVar) arrays) is passed as host_pitch = 1,
 * stride = nVar and transposed into the variable-major device layout by the same staging pass. */
CLODE_API int clode_sim_set_rows(clode_sim *sim, int which, const double *host, size_t rows, size_t host_pitch,
                                 size_t first, size_t stride);
CLODE_API int clode_sim_get_rows(clode_sim *sim, int which, double *host, size_t rows, size_t host_pitch,
                                 size_t first, size_t stride);
/* Instance-major upload: host holds one RECORD of `cols` consecutive doubles per instance, record k of this simulation
 * object at host + (first + k*stride) * record_pitch (record_pitch >= cols, in doubles).  The records travel to the
 * device as they are (one contiguous staging pass, no per-element gather on the CPU) and a kernel transposes them into
 * the variable-major device buffer `which` (x0 or pars; cols = nVar or nPar). */
CLODE_API int clode_sim_set_records(clode_sim *sim, int which, const double *host, size_t cols, size_t record_pitch,
                                    size_t first, size_t stride);

/* The same for an ensemble spread over several GPUs, without a strided pass over the host array per shard: the host array
 * of n_total records is cut into n_shards CONTIGUOUS chunks of ceil(n_total / n_shards) records; clode_sim_stage_records
 * moves chunk h to shards[h]'s GPU as it is (call it for every shard, from one host thread per GPU), then
 * clode_scatter_records lets every GPU pull the records of ITS interleaved shard out of all chunks with peer loads over
 * NVLink and transpose them into its variable-major buffer `which` (x0 or pars). */
CLODE_API int clode_sim_stage_records(clode_sim *sim, const double *chunk, size_t n_records, size_t cols);
CLODE_API int clode_scatter_records(clode_sim *const *shards, int n_shards, int which, size_t cols, size_t n_total);

/* The path's one exchange step (north_star: "only a final NVLink gather of features and final states"): the buffers
 * `which` of `n_shards` simulation objects — shard g holding instances g, g+n_shards, ... of a global ensemble of
 * n_total — are copied device-to-device over NVLink to `shards[0]`'s GPU (cuMemcpyPeerAsync, ordered behind each shard's
 * pending kernels by events, no host synchronisation in between), interleaved there into the global [rows][n_total]
 * layout by a small kernel, and brought to `host` with ONE device-to-host copy (full PCIe rate when `host` comes from
 * clode_host_alloc).  host may be NULL: the gathered array then stays on shards[0]'s GPU (clode_gathered_device_ptr). */
CLODE_API int clode_gather_rows(clode_sim *const *shards, int n_shards, int which, size_t rows, size_t n_total, double *host);
/* the same with the result transposed on the GPU to instance-major host[i*rows + r] — the layout of the Python front
 * end's record arrays (ObserverOutput.F, clode/features.py:51): one coalesced-read kernel instead of a host-side
 * unstructured_to_structured pass.  n_shards may be 1. */
CLODE_API int clode_gather_rows_instance_major(clode_sim *const *shards, int n_shards, int which, size_t rows, size_t n_total, double *host);
CLODE_API int clode_gathered_device_ptr(clode_sim *root, uint64_t *device_ptr, size_t *bytes);

/* raw device access for zero-copy consumers (e.g. wrap as a torch tensor for an NCCL gather):
 * pointer, size in bytes, element size (4 or 8) */
CLODE_API int clode_sim_device_buffer(clode_sim *sim, int which, uint64_t *device_ptr, size_t *bytes, int *elem_size);

/* measurement: device time of the kernel(s) of the last simulation call (CUDA events on the
 * launch stream), and static kernel facts */
CLODE_API int clode_sim_last_kernel_ms(clode_sim *sim, float *ms);
CLODE_API int clode_sim_launch_count(clode_sim *sim, uint64_t *launches);
typedef struct clode_kernel_info {
    int registers, local_bytes, shared_bytes, const_bytes, max_threads;
    int block_size, blocks_per_sm, grid_size;
} clode_kernel_info;
CLODE_API int clode_sim_kernel_info(clode_sim *sim, int kernel /* CLODE_KERNEL_* */, clode_kernel_info *info);

#ifdef __cplusplus
}
#endif
#endif /* CLODE_RT_H */
