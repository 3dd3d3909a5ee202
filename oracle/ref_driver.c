/* TEST INFRASTRUCTURE — not part of the product.
 *
 * oracle/_ref driver: compiles the REFERENCE's own kernel sources as host C and
 * runs them one work-item at a time.  The four kernels are included verbatim from
 * the reference tree (passed with -I by build_ref.py; nothing is copied into this
 * repository), followed by the right-hand-side file, exactly as the reference
 * concatenates them before handing the text to the OpenCL compiler
 * (clode/cpp/CLODE.cpp:144-152, clode/cpp/CLODEfeatures.cpp:15-16,
 * clode/cpp/CLODEtrajectory.cpp:12).
 *
 * Build-time macros (the reference's own -D options, clode/cpp/CLODE.cpp:118-138,
 * clode/cpp/CLODEfeatures.cpp:37-38): CLODE_{SINGLE,DOUBLE}_PRECISION, the stepper
 * define, N_PAR, N_VAR, N_AUX, N_WIENER, USE_OBSERVER_*, N_STORE_EVENTS, and
 * REF_RHS_FILE = path of the getRHS source.
 *
 * Each exported function is the host-side launch of one kernel over nPts
 * work-items (`enqueueNDRangeKernel(..., NDRange(nPts))`, clode/cpp/CLODE.cpp:484,
 * clode/cpp/CLODEfeatures.cpp:200,248, clode/cpp/CLODEtrajectory.cpp:121),
 * optionally spread over host threads with OpenMP (used only for CPU-baseline timing).
 */
#include "ref_shim.h"

_Thread_local int ref_gid = 0;
int ref_gsize = 0;

#include "transient.cl"
#include "initializeObserver.cl"
#include "features.cl"
#include "trajectory.cl"

#include REF_RHS_FILE

#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_FOR_EACH_WORKITEM(nPts, nthreads, CALL)                          \
    do {                                                                     \
        ref_gsize = (nPts);                                                  \
        if ((nthreads) <= 1) {                                               \
            for (int i_ = 0; i_ < (nPts); ++i_) { ref_gid = i_; CALL; }      \
        } else {                                                             \
            _Pragma("omp parallel for schedule(dynamic, 16) num_threads(nthreads)") \
            for (int i_ = 0; i_ < (nPts); ++i_) { ref_gid = i_; CALL; }      \
        }                                                                    \
    } while (0)

/* layout facts the Python wrapper needs */
void ref_info(long out[8])
{
    out[0] = (long)sizeof(realtype);
    out[1] = (long)sizeof(ObserverData);
    out[2] = (long)sizeof(struct SolverParams);
    out[3] = (long)sizeof(struct ObserverParams);
    out[4] = N_VAR;
    out[5] = N_PAR;
    out[6] = N_AUX;
    out[7] = N_WIENER;
}

void ref_transient(int nPts, int nthreads, const realtype *tspan, realtype *x0, const realtype *pars,
                   const struct SolverParams *sp, realtype *xf, ulong *rng, realtype *dt, realtype *tf)
{
    REF_FOR_EACH_WORKITEM(nPts, nthreads, transient(tspan, x0, pars, sp, xf, rng, dt, tf));
}

void ref_initialize_observer(int nPts, int nthreads, const realtype *tspan, realtype *x0, const realtype *pars,
                             const struct SolverParams *sp, ulong *rng, realtype *dt, void *odata,
                             const struct ObserverParams *op)
{
    REF_FOR_EACH_WORKITEM(nPts, nthreads,
                          initializeObserver(tspan, x0, pars, sp, rng, dt, (ObserverData *)odata, op));
}

void ref_features(int nPts, int nthreads, const realtype *tspan, realtype *x0, const realtype *pars,
                  const struct SolverParams *sp, realtype *xf, ulong *rng, realtype *dt, realtype *tf,
                  void *odata, const struct ObserverParams *op, realtype *F)
{
    REF_FOR_EACH_WORKITEM(nPts, nthreads,
                          features(tspan, x0, pars, sp, xf, rng, dt, tf, (ObserverData *)odata, op, F));
}

void ref_trajectory(int nPts, int nthreads, const realtype *tspan, realtype *x0, const realtype *pars,
                    const struct SolverParams *sp, realtype *xf, ulong *rng, realtype *dt, realtype *tf,
                    realtype *t, realtype *x, realtype *dx, realtype *aux, int *nStored)
{
    REF_FOR_EACH_WORKITEM(nPts, nthreads,
                          trajectory(tspan, x0, pars, sp, xf, rng, dt, tf, t, x, dx, aux, nStored));
}
