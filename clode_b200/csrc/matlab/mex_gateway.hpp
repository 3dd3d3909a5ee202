// mex_gateway.hpp — the MATLAB entry points of clODE on the B200 runtime.
//
// The reference ships three mex files, one per class, that its MATLAB classes (matlab/clODE.m, clODEfeatures.m,
// clODEtrajectory.m through cppclass.m) drive with (command string, instance handle, arguments...):
//   matlab/clODEmex.cpp:59-90, clODEfeaturesmex.cpp, clODEtrajectorymex.cpp; argument structs in clODEmexHelpers.hpp.
// This header is ONE dispatcher templated on the class; clODEmex.cpp / clODEfeaturesmex.cpp / clODEtrajectorymex.cpp
// next to it instantiate it.  Command names, argument positions, struct field names and the shapes of the returned
// arrays are the reference's, so the reference's .m classes work unchanged on top.  Differences, all forced by the
// C++ API the reference's own mex sources have drifted away from:
//   * 'initialize' = setProblemData + setTspan + setSolverParams (+ setObserverParams): CLODE::initialize is gone;
//   * 'setnpts' is accepted and ignored with a warning: nPts follows the problem data (CLODE::setNpts is protected);
//   * features 'new' takes the observer-parameter struct as an optional 8th argument.
// Errors of the runtime surface as MATLAB errors (mexErrMsgIdAndTxt "clODE:runtime") carrying the runtime's message.
#pragma once

#include "mex.h"

#include "CLODE.hpp"
#include "CLODEfeatures.hpp"
#include "CLODEtrajectory.hpp"

#include <algorithm>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#ifndef CLODE_ROOT
#define CLODE_ROOT "" // the engine sources are embedded in the runtime; the reference reads its .cl files from here
#endif

namespace clode_mex {

inline std::string toString(const mxArray *a)
{
    if (!a || !mxIsChar(a)) mexErrMsgIdAndTxt("clODE:args", "expected a character array");
    char *c = mxArrayToString(a);
    std::string s = c ? c : "";
    mxFree(c);
    return s;
}

inline std::vector<double> toVector(const mxArray *a)
{
    if (!a || !mxIsDouble(a)) mexErrMsgIdAndTxt("clODE:args", "expected a real double array");
    const double *p = mxGetPr(a);
    return std::vector<double>(p, p + mxGetNumberOfElements(a));
}

inline double fieldScalar(const mxArray *s, const char *name, bool required = true, double fallback = 0.0)
{
    const mxArray *f = mxIsStruct(s) ? mxGetField(s, 0, name) : nullptr;
    if (!f) {
        if (required) mexErrMsgIdAndTxt("clODE:args", "struct field '%s' is missing", name);
        return fallback;
    }
    return mxGetScalar(f);
}

inline std::vector<std::string> fieldNames(const mxArray *s, const char *name)
{
    std::vector<std::string> out;
    const mxArray *f = mxGetField(s, 0, name);
    if (!f) return out;
    if (mxIsChar(f)) { // a single name
        out.push_back(toString(f));
        return out;
    }
    if (!mxIsCell(f)) mexErrMsgIdAndTxt("clODE:args", "struct field '%s' must be a cell array of names", name);
    for (size_t i = 0; i < mxGetNumberOfElements(f); ++i) out.push_back(toString(mxGetCell(f, i)));
    return out;
}

// clODEmexHelpers.hpp:44-75 — fields clRHSfilename, nVar, nPar, nAux, nWiener, varNames, parNames, auxNames
inline ProblemInfo toProblemInfo(const mxArray *s)
{
    if (!s || !mxIsStruct(s)) mexErrMsgIdAndTxt("clODE:args", "expected the problem-info struct");
    ProblemInfo p;
    p.clRHSfilename = toString(mxGetField(s, 0, "clRHSfilename"));
    p.nVar = (cl_int)fieldScalar(s, "nVar");
    p.nPar = (cl_int)fieldScalar(s, "nPar");
    p.nAux = (cl_int)fieldScalar(s, "nAux");
    p.nWiener = (cl_int)fieldScalar(s, "nWiener");
    p.varNames = fieldNames(s, "varNames");
    p.parNames = fieldNames(s, "parNames");
    p.auxNames = fieldNames(s, "auxNames");
    return p;
}

// clODEmexHelpers.hpp:32-42
inline SolverParams<cl_double> toSolverParams(const mxArray *s)
{
    if (!s || !mxIsStruct(s)) mexErrMsgIdAndTxt("clODE:args", "expected the solver-parameter struct");
    SolverParams<cl_double> sp;
    sp.dt = fieldScalar(s, "dt");
    sp.dtmax = fieldScalar(s, "dtmax");
    sp.abstol = fieldScalar(s, "abstol");
    sp.reltol = fieldScalar(s, "reltol");
    sp.max_steps = (unsigned int)fieldScalar(s, "max_steps");
    sp.max_store = (unsigned int)fieldScalar(s, "max_store");
    sp.nout = (unsigned int)fieldScalar(s, "nout");
    return sp;
}

// clODEfeaturesmex.cpp getMatlabOPstruct; maxEventTimestamps is not in the reference's struct (defaults to 0)
inline ObserverParams<cl_double> toObserverParams(const mxArray *s)
{
    if (!s || !mxIsStruct(s)) mexErrMsgIdAndTxt("clODE:args", "expected the observer-parameter struct");
    ObserverParams<cl_double> op{};
    op.eVarIx = (unsigned int)fieldScalar(s, "eVarIx");
    op.fVarIx = (unsigned int)fieldScalar(s, "fVarIx");
    op.maxEventCount = (unsigned int)fieldScalar(s, "maxEventCount");
    op.maxEventTimestamps = (unsigned int)fieldScalar(s, "maxEventTimestamps", false, 0.0);
    op.minXamp = fieldScalar(s, "minXamp");
    op.minIMI = fieldScalar(s, "minIMI");
    op.nHoodRadius = fieldScalar(s, "nHoodRadius");
    op.xUpThresh = fieldScalar(s, "xUpThresh");
    op.xDownThresh = fieldScalar(s, "xDownThresh");
    op.dxUpThresh = fieldScalar(s, "dxUpThresh");
    op.dxDownThresh = fieldScalar(s, "dxDownThresh");
    op.eps_dx = fieldScalar(s, "eps_dx");
    return op;
}

inline ObserverParams<cl_double> defaultObserverParams()
{
    // CLODEpython.cpp:296-308 defaults
    ObserverParams<cl_double> op{};
    op.maxEventCount = 100;
    op.nHoodRadius = 0.05;
    op.xUpThresh = 0.2;
    op.xDownThresh = 0.2;
    return op;
}

inline mxArray *fromVector(const std::vector<double> &v, bool row = false)
{
    mxArray *a = row ? mxCreateDoubleMatrix(1, v.size(), mxREAL) : mxCreateDoubleMatrix(v.size(), 1, mxREAL);
    std::copy(v.begin(), v.end(), mxGetPr(a));
    return a;
}

inline mxArray *fromNames(const std::vector<std::string> &names)
{
    mxArray *c = mxCreateCellMatrix(names.size(), 1);
    for (size_t i = 0; i < names.size(); ++i) mxSetCell(c, i, mxCreateString(names[i].c_str()));
    return c;
}

template <class T> struct Traits;
template <> struct Traits<CLODE> {
    static std::shared_ptr<CLODE> make(int nrhs, const mxArray *prhs[], const ProblemInfo &p, const std::string &stepper,
                                       bool single, unsigned int platform, unsigned int device)
    {
        return std::make_shared<CLODE>(p, stepper, single, platform, device, CLODE_ROOT);
    }
};
template <> struct Traits<CLODEfeatures> {
    static std::shared_ptr<CLODEfeatures> make(int nrhs, const mxArray *prhs[], const ProblemInfo &p,
                                               const std::string &stepper, bool single, unsigned int platform,
                                               unsigned int device)
    {
        if (nrhs < 7) mexErrMsgIdAndTxt("clODE:args", "clODEfeatures constructor: the observer name is missing");
        const std::string observer = toString(prhs[6]);
        const ObserverParams<cl_double> op = nrhs > 7 ? toObserverParams(prhs[7]) : defaultObserverParams();
        return std::make_shared<CLODEfeatures>(p, stepper, observer, op, single, platform, device, CLODE_ROOT);
    }
};
template <> struct Traits<CLODEtrajectory> {
    static std::shared_ptr<CLODEtrajectory> make(int nrhs, const mxArray *prhs[], const ProblemInfo &p,
                                                 const std::string &stepper, bool single, unsigned int platform,
                                                 unsigned int device)
    {
        return std::make_shared<CLODEtrajectory>(p, stepper, single, platform, device, CLODE_ROOT);
    }
};

// commands only the derived classes understand; return true when handled
inline bool extra(CLODE &, const std::string &, int, mxArray *[], int, const mxArray *[]) { return false; }

inline bool extra(CLODEfeatures &f, const std::string &cmd, int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (cmd == "setobserverpars") f.setObserverParams(toObserverParams(prhs[2]));
    else if (cmd == "setobserver") f.setObserver(toString(prhs[2]));
    else if (cmd == "initobserver") f.initializeObserver();
    else if (cmd == "features") {
        if (nrhs > 2) f.features(mxGetScalar(prhs[2]) != 0.0);
        else f.features();
    } else if (cmd == "getnfeatures") plhs[0] = mxCreateDoubleScalar(f.getNFeatures());
    else if (cmd == "getf") plhs[0] = fromVector(f.getF());
    else if (cmd == "getfeaturenames") plhs[0] = fromNames(f.getFeatureNames());
    else if (cmd == "getobservernames") plhs[0] = fromNames(f.getAvailableObservers());
    else return false;
    return true;
}

inline bool extra(CLODEtrajectory &t, const std::string &cmd, int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (cmd == "trajectory") t.trajectory();
    else if (cmd == "gett") plhs[0] = fromVector(t.getT());
    else if (cmd == "getx") plhs[0] = fromVector(t.getX());
    else if (cmd == "getdx") plhs[0] = fromVector(t.getDx());
    else if (cmd == "getaux") plhs[0] = fromVector(t.getAux());
    else if (cmd == "getnstored") {
        const std::vector<cl_int> n = t.getNstored();
        plhs[0] = fromVector(std::vector<double>(n.begin(), n.end()));
    } else return false;
    return true;
}

template <class T> void initializeAll(T &obj, int nrhs, const mxArray *prhs[])
{
    if (nrhs < 6) mexErrMsgIdAndTxt("clODE:args", "initialize: expected tspan, x0, pars, solver-parameter struct");
    // solver parameters first: the per-instance dt buffer is filled with sp.dt when the problem data fixes nPts and is
    // not refilled by a later setSolverParams (reference semantics, CLODE.cpp:201-227, 382)
    obj.setSolverParams(toSolverParams(prhs[5]));
    obj.setTspan(toVector(prhs[2]));
    obj.setProblemData(toVector(prhs[3]), toVector(prhs[4]));
    if constexpr (std::is_same<T, CLODEfeatures>::value)
        if (nrhs > 6) obj.setObserverParams(toObserverParams(prhs[6]));
}

// the body of mexFunction for class T
template <class T> void dispatch(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    static std::map<unsigned int, std::shared_ptr<T>> instances; // lives as long as the mex file is loaded (mexLock)
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("clODE:args", "first argument must be a command string");
    const std::string cmd = toString(prhs[0]);
    try {
        if (cmd == "new") {
            if (nrhs < 6)
                mexErrMsgIdAndTxt("clODE:args", "new: expected problem struct, stepper, single-precision flag, platformID, deviceID");
            const ProblemInfo p = toProblemInfo(prhs[1]);
            const std::string stepper = toString(prhs[2]);
            const bool single = mxGetScalar(prhs[3]) != 0.0;
            const unsigned int platform = (unsigned int)mxGetScalar(prhs[4]), device = (unsigned int)mxGetScalar(prhs[5]);
            const unsigned int handle = instances.empty() ? 1u : instances.rbegin()->first + 1u;
            instances[handle] = Traits<T>::make(nrhs, prhs, p, stepper, single, platform, device);
            mexLock();
            plhs[0] = mxCreateDoubleScalar((double)handle);
            return;
        }
        if (nrhs < 2 || mxGetNumberOfElements(prhs[1]) != 1) mexErrMsgIdAndTxt("clODE:args", "second argument must be an instance handle");
        const unsigned int handle = (unsigned int)mxGetScalar(prhs[1]);
        auto it = instances.find(handle);
        if (it == instances.end()) mexErrMsgIdAndTxt("clODE:handle", "no instance with handle %u", handle);
        T &obj = *it->second;
        auto need = [&](int n) { if (nrhs < n) mexErrMsgIdAndTxt("clODE:args", "%s: too few arguments", cmd.c_str()); };

        if (cmd == "delete") {
            instances.erase(it);
            mexUnlock();
            plhs[0] = mxCreateLogicalScalar(instances.empty());
        } else if (cmd == "setProblemInfo") { need(3); obj.setProblemInfo(toProblemInfo(prhs[2])); }
        else if (cmd == "setstepper") { need(3); obj.setStepper(toString(prhs[2])); }
        else if (cmd == "setprecision") { need(3); obj.setPrecision(mxGetScalar(prhs[2]) != 0.0); }
        else if (cmd == "setopencl") { need(4); obj.setOpenCL((unsigned int)mxGetScalar(prhs[2]), (unsigned int)mxGetScalar(prhs[3])); }
        else if (cmd == "buildcl") obj.buildCL();
        else if (cmd == "initialize") initializeAll(obj, nrhs, prhs);
        else if (cmd == "setnpts") mexWarnMsgTxt("clODE: setnpts is ignored, the number of points follows setproblemdata");
        else if (cmd == "setproblemdata") { need(4); obj.setProblemData(toVector(prhs[2]), toVector(prhs[3])); }
        else if (cmd == "settspan") { need(3); obj.setTspan(toVector(prhs[2])); }
        else if (cmd == "setx0") { need(3); obj.setX0(toVector(prhs[2])); }
        else if (cmd == "setpars") { need(3); obj.setPars(toVector(prhs[2])); }
        else if (cmd == "setsolverpars") { need(3); obj.setSolverParams(toSolverParams(prhs[2])); }
        else if (cmd == "seedrng") {
            if (nrhs > 2) obj.seedRNG((cl_int)mxGetScalar(prhs[2]));
            else obj.seedRNG();
        } else if (cmd == "transient") obj.transient();
        else if (cmd == "shifttspan") obj.shiftTspan();
        else if (cmd == "shiftx0") obj.shiftX0();
        else if (cmd == "gettspan") plhs[0] = fromVector(obj.getTspan());
        else if (cmd == "getx0") plhs[0] = fromVector(obj.getX0());
        else if (cmd == "getxf") plhs[0] = fromVector(obj.getXf(), true); // a row, as in the reference (clODEmex.cpp:324)
        else if (cmd == "getsteppernames") plhs[0] = fromNames(obj.getAvailableSteppers());
        else if (cmd == "getprogramstring") {
            plhs[0] = mxCreateCellMatrix(1, 1);
            mxSetCell(plhs[0], 0, mxCreateString(obj.getProgramString().c_str()));
        } else if (cmd == "printstatus") obj.printStatus();
        else if (!extra(obj, cmd, nlhs, plhs, nrhs, prhs))
            mexErrMsgIdAndTxt("clODE:command", "unrecognized command: %s", cmd.c_str());
    } catch (const std::exception &e) {
        mexErrMsgIdAndTxt("clODE:runtime", "%s: %s", cmd.c_str(), e.what());
    }
}

} // namespace clode_mex
