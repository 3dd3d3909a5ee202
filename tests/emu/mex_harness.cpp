// mex_harness.cpp — drives the MATLAB entry points (clode_b200/csrc/matlab) through the in-process mex stub.
// Built three times by tests/test_mex_gateway.py, once per mex file (-DWHICH=0 base, 1 features, 2 trajectory);
// prints "key value..." lines the test parses.   usage: mex_harness <rhs file> cpu|gpu
#include "mex.h"
#include "mex_stub.hpp"

#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

static mxArray *strcell(const std::vector<std::string> &v)
{
    mxArray *c = mxCreateCellMatrix(v.size(), 1);
    for (size_t i = 0; i < v.size(); ++i) mxSetCell(c, i, mxCreateString(v[i].c_str()));
    return c;
}
static mxArray *scalar_struct(const std::vector<std::pair<std::string, mxArray *>> &fields)
{
    std::vector<const char *> names;
    for (auto &f : fields) names.push_back(f.first.c_str());
    mxArray *s = mxCreateStructMatrix(1, 1, (int)names.size(), names.data());
    for (auto &f : fields) mxSetField(s, 0, f.first.c_str(), f.second);
    return s;
}
static mxArray *num(double v) { return mxCreateDoubleScalar(v); }
static mxArray *vec(const std::vector<double> &v)
{
    mxArray *a = mxCreateDoubleMatrix(v.size(), 1, mxREAL);
    for (size_t i = 0; i < v.size(); ++i) mxGetPr(a)[i] = v[i];
    return a;
}
static mxArray *call(const std::vector<mxArray *> &args, int nlhs = 1)
{
    mxArray *out[2] = {nullptr, nullptr};
    mexFunction(nlhs, out, (int)args.size(), const_cast<const mxArray **>(args.data()));
    return out[0];
}
static void print_vec(const char *key, const mxArray *a, size_t limit = 8)
{
    std::printf("%s %zu x %zu :", key, mxGetM(a), mxGetN(a));
    for (size_t i = 0; i < mxGetNumberOfElements(a) && i < limit; ++i) std::printf(" %.17g", mxGetPr(a)[i]);
    std::printf("\n");
}

int main(int argc, char **argv)
{
    const std::string rhs = argv[1], mode = argc > 2 ? argv[2] : "cpu";
    // Lorenz: 3 variables, 3 parameters, 1 aux (clode_b200/models/lorenz63.cl)
    mxArray *prob = scalar_struct({{"clRHSfilename", mxCreateString(rhs.c_str())}, {"nVar", num(3)}, {"nPar", num(3)},
                                   {"nAux", num(1)}, {"nWiener", num(0)}, {"varNames", strcell({"x", "y", "z"})},
                                   {"parNames", strcell({"r", "s", "b"})}, {"auxNames", strcell({"dx"})}});
    mxArray *sp = scalar_struct({{"dt", num(0.01)}, {"dtmax", num(1.0)}, {"abstol", num(1e-6)}, {"reltol", num(1e-6)},
                                 {"max_steps", num(100000)}, {"max_store", num(50)}, {"nout", num(1)}});
    mxArray *op = scalar_struct({{"eVarIx", num(0)}, {"fVarIx", num(0)}, {"maxEventCount", num(100)}, {"minXamp", num(0)},
                                 {"minIMI", num(0)}, {"nHoodRadius", num(0.05)}, {"xUpThresh", num(0.3)},
                                 {"xDownThresh", num(0.2)}, {"dxUpThresh", num(0)}, {"dxDownThresh", num(0)}, {"eps_dx", num(0)}});
    (void)op;
    // argument errors never reach the runtime
    try { call({mxCreateString("transient")}); std::printf("nohandle no-error\n"); }
    catch (const MexError &e) { std::printf("nohandle %s\n", e.id.c_str()); }
    try { call({mxCreateString("transient"), num(42)}); std::printf("badhandle no-error\n"); }
    catch (const MexError &e) { std::printf("badhandle %s\n", e.id.c_str()); }
    try { call({num(1)}); std::printf("nocommand no-error\n"); }
    catch (const MexError &e) { std::printf("nocommand %s\n", e.id.c_str()); }

    mxArray *handle = nullptr;
    try {
        std::vector<mxArray *> ctor = {mxCreateString("new"), prob, mxCreateString("dopri5"), num(0), num(0), num(0)};
#if WHICH == 1
        ctor.push_back(mxCreateString("basic"));
        ctor.push_back(op);
#endif
        handle = call(ctor);
    } catch (const MexError &e) {
        // without a GPU the constructor fails loudly inside the runtime and the gateway turns that into a MATLAB error
        std::printf("new-error %s | %s\n", e.id.c_str(), e.what());
        return mode == "cpu" ? 0 : 1;
    }
    std::printf("handle %g locks %d\n", mxGetScalar(handle), mex_stub_lock_count);
    const int n = 64;
    std::vector<double> x0(3 * n, 1.0), pars(3 * n);
    for (int i = 0; i < n; ++i) { pars[i] = 5.0 + 20.0 * i / (n - 1); pars[n + i] = 10.0; pars[2 * n + i] = 8.0 / 3.0; }
    std::vector<mxArray *> init = {mxCreateString("initialize"), handle, vec({0.0, 5.0}), vec(x0), vec(pars), sp};
#if WHICH == 1
    init.push_back(op);
#endif
    call(init, 0);
    call({mxCreateString("seedrng"), handle, num(1)}, 0);
    mxArray *names = call({mxCreateString("getsteppernames"), handle});
    std::printf("steppers %zu\n", mxGetNumberOfElements(names));
    call({mxCreateString("transient"), handle}, 0);
    print_vec("xf", call({mxCreateString("getxf"), handle}));
    print_vec("tspan", call({mxCreateString("gettspan"), handle}));
    call({mxCreateString("setnpts"), handle, num(n)}, 0);
    std::printf("warning %s\n", mex_stub_last_warning.c_str());
#if WHICH == 1
    call({mxCreateString("features"), handle}, 0);
    std::printf("nfeatures %g\n", mxGetScalar(call({mxCreateString("getnfeatures"), handle})));
    print_vec("F", call({mxCreateString("getf"), handle}));
    std::printf("featurenames %zu observers %zu\n", mxGetNumberOfElements(call({mxCreateString("getfeaturenames"), handle})),
                mxGetNumberOfElements(call({mxCreateString("getobservernames"), handle})));
#elif WHICH == 2
    call({mxCreateString("trajectory"), handle}, 0);
    print_vec("nstored", call({mxCreateString("getnstored"), handle}));
    print_vec("t", call({mxCreateString("gett"), handle}));
    print_vec("x", call({mxCreateString("getx"), handle}));
#endif
    try { call({mxCreateString("nosuchcommand"), handle}); std::printf("unknown no-error\n"); }
    catch (const MexError &e) { std::printf("unknown %s\n", e.id.c_str()); }
    mxArray *empty = call({mxCreateString("delete"), handle});
    std::printf("deleted %g locks %d\n", mxGetScalar(empty), mex_stub_lock_count);
    return 0;
}
