// CLODE — host class of the clODE C++ API (clode/cpp/CLODE.hpp:22-189), re-implemented on the
// B200 runtime (include/clode_rt.h).  Public names, signatures and semantics are the reference's;
// what changed is underneath: no cl::Buffer / cl::Kernel members, but one runtime simulation
// object per selected GPU, each holding a contiguous shard of the ensemble.
#pragma once

#include "OpenCLResource.hpp"
#include "clODE_struct_defs.hpp"
#include "clode_rt.h"

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

struct ProblemInfo {
    std::string clRHSfilename;
    cl_int nVar = 0, nPar = 0, nAux = 0, nWiener = 0;
    std::vector<std::string> varNames, parNames, auxNames;

    ProblemInfo(std::string clRHSfilename, cl_int nVar, cl_int nPar, cl_int nAux, cl_int nWiener,
                std::vector<std::string> varNames, std::vector<std::string> parNames, std::vector<std::string> auxNames)
        : clRHSfilename(clRHSfilename), nVar(nVar), nPar(nPar), nAux(nAux), nWiener(nWiener), varNames(varNames),
          parNames(parNames), auxNames(auxNames)
    {
    }
    ProblemInfo(std::string clRHSfilename, std::vector<std::string> varNames, std::vector<std::string> parNames,
                std::vector<std::string> auxNames = std::vector<std::string>(), cl_int nWiener = 0)
        : clRHSfilename(clRHSfilename), nVar((cl_int)varNames.size()), nPar((cl_int)parNames.size()),
          nAux((cl_int)auxNames.size()), nWiener(nWiener), varNames(varNames), parNames(parNames), auxNames(auxNames)
    {
    }
    ProblemInfo() {}

    void setVarNames(std::vector<std::string> v) { varNames = v; nVar = (cl_int)v.size(); }
    void setParNames(std::vector<std::string> v) { parNames = v; nPar = (cl_int)v.size(); }
    void setAuxNames(std::vector<std::string> v) { auxNames = v; nAux = (cl_int)v.size(); }
    std::vector<std::string> getVarNames() { return varNames; }
    std::vector<std::string> getParNames() { return parNames; }
    std::vector<std::string> getAuxNames() { return auxNames; }
};

class CLODE
{
protected:
    ProblemInfo prob;
    std::string clRHSfilename;
    cl_int nVar = 0, nPar = 0, nAux = 0, nWiener = 0;
    cl_int nPts = 0;

    std::string stepper;
    std::vector<std::string> availableSteppers;
    std::map<std::string, std::string> stepperDefineMap;

    bool clSinglePrecision = false;
    size_t realSize = 8;

    OpenCLResource opencl;
    std::string clodeRoot;

    cl_int nRNGstate = 2;

    SolverParams<cl_double> sp{0.1, 0.5, 1e-6, 1e-3, 1000000u, 1000000u, 1u};
    std::vector<cl_double> tspan{0.0, 1.0}, x0, pars, xf, dt, tf;
    size_t x0elements = 0, parselements = 0, RNGelements = 0;
    std::vector<cl_ulong> RNGstate;

    std::string clprogramstring, buildOptions, ODEsystemsource;

    // ---- runtime objects: one per GPU.  Shard g of G owns instances g, g+G, g+2G, ... (cost-balanced:
    // sweeps are usually sorted grids whose cost varies smoothly, so contiguous ranges would load the GPUs
    // unevenly); on its GPU a shard is stored contiguously, so device accesses stay coalesced.
    struct Shard {
        clode_sim *sim = nullptr;
        int device = 0;
        size_t first = 0, stride = 1, count = 0; // global instance of local column k: first + k * stride
    };
    template <typename T> static void takeShard(const std::vector<T> &full, size_t nTotal, int rows, const Shard &s, std::vector<T> &part)
    {
        part.resize((size_t)rows * s.count);
        for (int r = 0; r < rows; ++r)
            for (size_t k = 0; k < s.count; ++k) part[(size_t)r * s.count + k] = full[(size_t)r * nTotal + s.first + k * s.stride];
    }
    template <typename T> static void putShard(std::vector<T> &full, size_t nTotal, int rows, const Shard &s, const std::vector<T> &part)
    {
        for (int r = 0; r < rows; ++r)
            for (size_t k = 0; k < s.count; ++k) full[(size_t)r * nTotal + s.first + k * s.stride] = part[(size_t)r * s.count + k];
    }
    struct Runtime; // owns the clode_sim handles; shared so that copies of a CLODE stay valid
    std::shared_ptr<Runtime> runtime;
    std::vector<Shard> &shards();
    bool programBuilt = false;

    // what the derived class needs compiled into the program
    virtual int kernelMask() const { return CLODE_KERNEL_TRANSIENT; }
    virtual void fillProgramDesc(clode_program_desc &) const {}
    virtual void onNptsChanged() {}

    void check(int status, const char *where) const; // log + throw on a runtime error
    void makeShards();
    void buildProgram();
    void pushSolverParams();
    void setNpts(cl_int newNpts);
    // host [rows][nPts] <-> the shards: every shard moves its own columns (strided, through the runtime's page-locked
    // staging ring), all shards concurrently; results of several GPUs come back through the NVLink gather
    void uploadRows(const cl_double *full, int rows, int which, const char *where);
    void uploadMatrix(const cl_double *a, int rows, size_t rowStride, size_t instStride, int which, const char *where);
    void downloadRows(cl_double *full, int rows, int which, const char *where);
    void downloadRows(std::vector<cl_double> &full, int rows, int which, const char *where);
    void forEachShard(const std::function<int(Shard &)> &fn, const char *where);
    mutable bool parsOnDeviceOnly = false; // setPars(pointer) skipped the host mirror; getPars() fetches it on demand
    void runOnShards(int kernel, int initialize, const char *where);
    std::string getStepperDefine();

public:
    CLODE(ProblemInfo prob, std::string stepper, bool clSinglePrecision, OpenCLResource opencl, const std::string clodeRoot);
    CLODE(ProblemInfo prob, std::string stepper, bool clSinglePrecision, unsigned int platformID, unsigned int deviceID,
          const std::string clodeRoot);
    virtual ~CLODE();

    void setProblemInfo(ProblemInfo prob);
    void setStepper(std::string newStepper);
    void setPrecision(bool clSinglePrecision);
    void setOpenCL(OpenCLResource opencl);
    void setOpenCL(unsigned int platformID, unsigned int deviceID);

    virtual void buildCL();

    void setProblemData(std::vector<cl_double> newX0, std::vector<cl_double> newPars);
    // additions: the same calls on caller-owned memory (numpy buffers): no std::vector copy on the way in, results
    // written straight into caller-owned (ideally page-locked, clode_host_alloc) memory on the way out
    void setProblemData(const cl_double *newX0, size_t nX0, const cl_double *newPars, size_t nParsValues);
    void setX0(const cl_double *newX0, size_t count);
    void setPars(const cl_double *newPars, size_t count);
    void fetch(int which, int rows, cl_double *out, const char *where = "CLODE::fetch"); // which = CLODE_BUF_*
    void fetchInstanceMajor(int which, int rows, cl_double *out, const char *where = "CLODE::fetch"); // out[i*rows + r]
    // (ensemble x nVar) / (ensemble x nPar) matrices with arbitrary element strides (numpy arrays as the Python front end
    // holds them): transposed into the variable-major device layout inside the staging copy, no host-side flatten
    void setProblemDataMatrix(const cl_double *x0, size_t nX0rows, ptrdiff_t x0InstStride, ptrdiff_t x0VarStride,
                              const cl_double *pars, size_t nParsRows, ptrdiff_t parsInstStride, ptrdiff_t parsParStride);
    void setX0Matrix(const cl_double *x0, size_t rows, ptrdiff_t instStride, ptrdiff_t varStride);
    void setParsMatrix(const cl_double *pars, size_t rows, ptrdiff_t instStride, ptrdiff_t parStride);
    void setTspan(std::vector<cl_double> newTspan);
    void setX0(std::vector<cl_double> newX0);
    void setPars(std::vector<cl_double> newPars);
    void setSolverParams(SolverParams<cl_double> newSp);

    void seedRNG();
    void seedRNG(cl_int mySeed);

    void transient();

    void shiftTspan();
    void shiftX0();

    const ProblemInfo getProblemInfo() const { return prob; }
    const std::vector<cl_double> getTspan() const { return tspan; }
    const SolverParams<cl_double> getSolverParams() const { return sp; }
    const std::vector<cl_double> getPars() const;
    const std::vector<cl_double> getX0();
    const std::vector<cl_double> getXf();
    const std::vector<cl_double> getDt();
    const std::vector<cl_double> getTf();
    const std::vector<std::string> getAvailableSteppers() const { return availableSteppers; }

    const std::string getProgramString() const { return buildOptions + clprogramstring + ODEsystemsource; }
    void printStatus();

    int getNpts() const { return nPts; }
    int getNvar() const { return nVar; }
    int getNpar() const { return nPar; }
    // additions (not in the reference): measurement hooks used by bench/tests
    double getLastKernelMilliseconds() const;
    std::vector<unsigned int> getStepCounts();
    std::vector<cl_ulong> getRNGstate();
};
