// Host-side parameter structs and scalar type names of the clODE C++ API.
//   SolverParams<T>   — clode/cpp/clODE_struct_defs.cl:9-20
//   ObserverParams<T> — clode/cpp/observers.cl:22-46
// The cl_* aliases keep caller code written against the reference headers
// (`std::vector<cl_double>`, `cl_int nPts`, ...) compiling unchanged.
#pragma once

#include <cstdint>

typedef double cl_double;
typedef float cl_float;
typedef int cl_int;
typedef unsigned int cl_uint;
typedef std::uint64_t cl_ulong;
typedef cl_ulong cl_device_type;
typedef cl_uint cl_bool;

template <typename realtype> struct SolverParams {
    realtype dt;
    realtype dtmax;
    realtype abstol;
    realtype reltol;
    unsigned int max_steps;
    unsigned int max_store;
    unsigned int nout;
};

template <typename realtype> struct ObserverParams {
    unsigned int eVarIx;             // variable for event detection
    unsigned int fVarIx;             // variable for features
    unsigned int maxEventCount;      // time-loop limiter
    unsigned int maxEventTimestamps; // number of event timestamps to store
    realtype minXamp;
    realtype minIMI;
    realtype nHoodRadius;
    realtype xUpThresh;
    realtype xDownThresh;
    realtype dxUpThresh;
    realtype dxDownThresh;
    realtype eps_dx;
};
