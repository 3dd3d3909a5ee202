#include "CLODE.hpp"

#include "clode_log.hpp"

#include <algorithm>
#include <cmath>
#include <random>
#include <stdexcept>
#include <thread>

namespace lg = clode_log;

struct CLODE::Runtime {
    std::vector<Shard> shards;
    ~Runtime()
    {
        for (auto &s : shards)
            if (s.sim) clode_sim_destroy(s.sim);
    }
};

std::vector<CLODE::Shard> &CLODE::shards() { return runtime->shards; }

// name -> build define, clode/cpp/steppers.cl:26-43
static void getStepperDefineMap(std::map<std::string, std::string> &m, std::vector<std::string> &names)
{
    m = {{"euler", "EXPLICIT_EULER"}, {"heun", "EXPLICIT_HEUN"}, {"rk4", "EXPLICIT_RK4"},
         {"bs23", "EXPLICIT_BS23"},   {"dopri5", "EXPLICIT_DOPRI5"}, {"seuler", "STOCHASTIC_EULER"}};
    names.clear();
    for (auto const &e : m) names.push_back(e.first);
}

CLODE::CLODE(ProblemInfo prob, std::string stepper, bool clSinglePrecision, OpenCLResource opencl, const std::string clodeRoot)
    : opencl(opencl), clodeRoot(clodeRoot), runtime(std::make_shared<Runtime>())
{
    getStepperDefineMap(stepperDefineMap, availableSteppers);
    this->stepper = "rk4";
    setProblemInfo(prob);
    setStepper(stepper);
    setPrecision(clSinglePrecision);
    lg::debug_("constructor clODE");
}

CLODE::CLODE(ProblemInfo prob, std::string stepper, bool clSinglePrecision, unsigned int platformID, unsigned int deviceID,
             const std::string clodeRoot)
    : CLODE(prob, stepper, clSinglePrecision, OpenCLResource(platformID, deviceID), clodeRoot)
{
}

CLODE::~CLODE() {}

void CLODE::check(int status, const char *where) const
{
    if (status == CLODE_OK) return;
    std::string msg = clode_last_error();
    lg::error_("{}:{}({})", where, msg, CLErrorString(status));
    if (status == CLODE_ERR_MEMORY) throw std::invalid_argument(msg);
    throw std::runtime_error(std::string(where) + ": " + msg);
}

void CLODE::setProblemInfo(ProblemInfo newProb)
{
    prob = newProb;
    clRHSfilename = newProb.clRHSfilename;
    ODEsystemsource = read_file(clRHSfilename);
    nVar = newProb.nVar;
    nPar = newProb.nPar;
    nAux = newProb.nAux;
    nWiener = newProb.nWiener;
    programBuilt = false;
    lg::debug_("set new problem");
}

void CLODE::setStepper(std::string newStepper)
{
    if (stepperDefineMap.find(newStepper) != stepperDefineMap.end()) {
        stepper = newStepper;
        programBuilt = false;
    } else {
        lg::warn_("Unknown stepper: {}. Stepper method unchanged", newStepper);
    }
    lg::debug_("set stepper");
}

void CLODE::setPrecision(bool newPrecision)
{
    clSinglePrecision = newPrecision;
    realSize = newPrecision ? sizeof(cl_float) : sizeof(cl_double);
    programBuilt = false;
    lg::debug_("set precision");
}

void CLODE::setOpenCL(OpenCLResource newOpencl)
{
    opencl = newOpencl;
    runtime = std::make_shared<Runtime>(); // device set changed: all device state is dropped
    programBuilt = false;
    nPts = 0;
    lg::debug_("set OpenCL");
}

void CLODE::setOpenCL(unsigned int platformID, unsigned int deviceID) { setOpenCL(OpenCLResource(platformID, deviceID)); }

std::string CLODE::getStepperDefine() { return stepperDefineMap.at(stepper); }

void CLODE::makeShards()
{
    if (!shards().empty()) return;
    for (int dev : opencl.getDeviceOrdinals()) {
        Shard s;
        s.device = dev;
        check(clode_sim_create(dev, &s.sim), "CLODE::buildCL(): create runtime");
        shards().push_back(s);
    }
}

// CLODE::setCLbuildOpts + buildProgram (clode/cpp/CLODE.cpp:109-152): one JIT-specialised program per GPU
void CLODE::buildProgram()
{
    makeShards();
    clode_program_desc d{};
    d.rhs_source = ODEsystemsource.c_str();
    d.stepper = stepper.c_str();
    d.observer = nullptr;
    d.single_precision = clSinglePrecision ? 1 : 0;
    d.n_var = nVar; d.n_par = nPar; d.n_aux = nAux; d.n_wiener = nWiener;
    d.kernels = kernelMask();
    fillProgramDesc(d);
    for (auto &s : shards()) {
        int rc = clode_sim_build(s.sim, &d);
        if (rc == CLODE_ERR_BUILD) lg::error_("Program build failed. Build log:\n{}", clode_last_error());
        check(rc, "CLODE::buildProgram");
    }
    char *src = nullptr;
    if (clode_program_source(&d, &src) == CLODE_OK && src) {
        std::string all(src);
        clode_free(src);
        size_t nl = all.find('\n');
        buildOptions = all.substr(0, nl + 1);
        clprogramstring = all.substr(nl + 1, all.size() - nl - 1 - ODEsystemsource.size() - 1);
    }
    programBuilt = true;
    // device buffers were dropped if precision / dimensions changed; re-create them lazily
    size_t have = 0;
    for (auto &s : shards()) {
        size_t n = 0;
        clode_sim_get_npts(s.sim, &n);
        have += n;
    }
    if (have != (size_t)nPts) nPts = 0;
    pushSolverParams();
    for (auto &s : shards()) check(clode_sim_set_tspan(s.sim, tspan[0], tspan[1]), "CLODE::setTspan");
    lg::trace_("{}", clprogramstring + ODEsystemsource);
    lg::debug_("CLODE buildProgram finished");
}

void CLODE::buildCL()
{
    lg::info_("Running CLODE buildCL");
    buildProgram();
    lg::debug_("Created kernel");
}

// CLODE::setNpts (clode/cpp/CLODE.cpp:174-243)
void CLODE::setNpts(cl_int newNpts)
{
    size_t largestAlloc = (size_t)std::max(nVar, std::max(nPar, nAux)) * (size_t)newNpts * realSize;
    if (largestAlloc > opencl.getMaxMemAllocSize()) throw std::invalid_argument("nPts*nVar, nPts*nPar, or nPts*nAux is too large");
    if (newNpts == nPts) return;
    if (!programBuilt) buildCL();
    nPts = newNpts;
    x0elements = (size_t)nVar * nPts;
    parselements = (size_t)nPar * nPts;
    RNGelements = (size_t)nRNGstate * nPts;
    x0.resize(x0elements);
    pars.resize(parselements);
    RNGstate.resize(RNGelements);
    dt.assign(nPts, sp.dt);
    tf.resize(nPts);
    xf.resize(x0elements);

    // interleaved shards, one per GPU
    const size_t g = shards().size();
    for (size_t k = 0; k < g; ++k) {
        Shard &s = shards()[k];
        s.first = k;
        s.stride = g;
        s.count = (size_t)nPts > k ? ((size_t)nPts - k + g - 1) / g : 0;
        check(clode_sim_set_npts(s.sim, s.count, sp.dt), "CLODE::setNpts");
    }
    onNptsChanged();
    seedRNG(); // must follow the allocation of the RNG state (CLODE.cpp:239)
    lg::debug_("set nPts={}", nPts);
}

// run fn on every non-empty shard, concurrently when there are several (one host thread per GPU: the staging copies
// are CPU work, the DMAs go over separate PCIe links).  The runtime's error text is thread-local, so it travels back
// with the status.
void CLODE::forEachShard(const std::function<int(Shard &)> &fn, const char *where)
{
    std::vector<Shard *> busy;
    for (auto &s : shards())
        if (s.count) busy.push_back(&s);
    if (busy.size() <= 1) {
        for (Shard *s : busy) check(fn(*s), where);
        return;
    }
    std::vector<int> rc(busy.size(), CLODE_OK);
    std::vector<std::string> msg(busy.size());
    std::vector<std::thread> workers;
    for (size_t k = 0; k < busy.size(); ++k)
        workers.emplace_back([&, k] {
            rc[k] = fn(*busy[k]);
            if (rc[k]) msg[k] = clode_last_error();
        });
    for (auto &w : workers) w.join();
    for (size_t k = 0; k < busy.size(); ++k)
        if (rc[k]) {
            lg::error_("{}:{}({})", where, msg[k], CLErrorString(rc[k]));
            if (rc[k] == CLODE_ERR_MEMORY) throw std::invalid_argument(msg[k]);
            throw std::runtime_error(std::string(where) + ": " + msg[k]);
        }
}

// host [rows][nPts] -> per-shard [rows][count]
void CLODE::uploadRows(const cl_double *full, int rows, int which, const char *where)
{
    const size_t pitch = (size_t)nPts;
    forEachShard([&](Shard &s) { return clode_sim_set_rows(s.sim, which, full, (size_t)rows, pitch, s.first, s.stride); }, where);
}

// host element (row r, instance i) at a[r*rowStride + i*instStride] -> per-shard [rows][count]
void CLODE::uploadMatrix(const cl_double *a, int rows, size_t rowStride, size_t instStride, int which, const char *where)
{
    if (rowStride == 1 && instStride == (size_t)rows && shards().size() > 1 && shards().size() <= 16 &&
        (size_t)nPts >= shards().size()) {
        // dense records on several GPUs: contiguous chunk h to GPU h (dense DMA, one host thread per GPU), then every GPU
        // pulls its interleaved shard out of all chunks with peer loads over NVLink (clode_scatter_records) — instead of
        // one strided pass over the whole host array per shard
        const size_t G = shards().size(), chunk = ((size_t)nPts + G - 1) / G;
        std::vector<clode_sim *> sims;
        for (auto &s : shards()) sims.push_back(s.sim);
        std::vector<int> rc(G, CLODE_OK);
        std::vector<std::string> msg(G);
        std::vector<std::thread> workers;
        for (size_t h = 0; h < G; ++h)
            workers.emplace_back([&, h] {
                const size_t lo = std::min((size_t)nPts, h * chunk), hi = std::min((size_t)nPts, lo + chunk);
                rc[h] = clode_sim_stage_records(sims[h], a + lo * rows, hi - lo, (size_t)rows);
                if (rc[h]) msg[h] = clode_last_error();
            });
        for (auto &w : workers) w.join();
        for (size_t h = 0; h < G; ++h)
            if (rc[h]) throw std::runtime_error(std::string(where) + ": " + msg[h]);
        int status = clode_scatter_records(sims.data(), (int)G, which, (size_t)rows, (size_t)nPts);
        if (status == CLODE_OK) return;
        lg::warn_("{}: peer scatter unavailable ({}), falling back to per-shard strided uploads", where, clode_last_error());
    }
    if (rowStride == 1 && instStride >= (size_t)rows) { // records of `rows` consecutive values: moved as they are, transposed on the GPU
        forEachShard([&](Shard &s) { return clode_sim_set_records(s.sim, which, a, (size_t)rows, instStride, s.first, s.stride); }, where);
        return;
    }
    forEachShard([&](Shard &s) {
        return clode_sim_set_rows(s.sim, which, a, (size_t)rows, rowStride, s.first * instStride, s.stride * instStride);
    }, where);
}

void CLODE::setX0Matrix(const cl_double *a, size_t rows, ptrdiff_t instStride, ptrdiff_t varStride)
{
    if (rows != (size_t)nPts || instStride <= 0 || varStride <= 0) {
        lg::info_("...Initial conditions were not updated!");
        return;
    }
    if (nVar > 0 && nPts > 0) uploadMatrix(a, nVar, (size_t)varStride, (size_t)instStride, CLODE_BUF_X0, "CLODE::setX0");
    lg::debug_("set X0");
}

void CLODE::setParsMatrix(const cl_double *a, size_t rows, ptrdiff_t instStride, ptrdiff_t parStride)
{
    if (rows != (size_t)nPts || instStride <= 0 || parStride <= 0) {
        lg::info_("Invalid parameter vector: Expected {}*{} elements, recieved {}", nPts, nPar, rows * nPar);
        lg::info_("...Parameters were not updated!");
        return;
    }
    if (nPar > 0 && nPts > 0) uploadMatrix(a, nPar, (size_t)parStride, (size_t)instStride, CLODE_BUF_PARS, "CLODE::setPars");
    parsOnDeviceOnly = true;
    lg::debug_("set P");
}

void CLODE::setProblemDataMatrix(const cl_double *x0m, size_t nX0rows, ptrdiff_t x0InstStride, ptrdiff_t x0VarStride,
                                 const cl_double *parsm, size_t nParsRows, ptrdiff_t parsInstStride, ptrdiff_t parsParStride)
{
    if (nPar > 0 && nX0rows != nParsRows) {
        lg::info_("Initial contition and parameter vector dimensions don't match");
        lg::info_("...Expected {} sets of each, recieved {} for x0 and {} for pars", nPts, nX0rows, nParsRows);
        lg::info_("...Problem data was not updated!");
        return;
    }
    setNpts((cl_int)nX0rows);
    setX0Matrix(x0m, nX0rows, x0InstStride, x0VarStride);
    setParsMatrix(parsm, nParsRows, parsInstStride, parsParStride);
    lg::debug_("set problem data");
}

// per-shard [rows][count] -> host [rows][nPts]
void CLODE::downloadRows(cl_double *full, int rows, int which, const char *where)
{
    if (shards().size() == 1) {
        check(clode_sim_get_rows(shards()[0].sim, which, full, (size_t)rows, (size_t)nPts, 0, 1), where);
        return;
    }
    // several GPUs: one NVLink gather to the first shard's GPU, one copy to the host (clode_gather_rows)
    std::vector<clode_sim *> sims;
    for (auto &s : shards()) sims.push_back(s.sim);
    check(clode_gather_rows(sims.data(), (int)sims.size(), which, (size_t)rows, (size_t)nPts, full), where);
}

void CLODE::downloadRows(std::vector<cl_double> &full, int rows, int which, const char *where)
{
    full.resize((size_t)rows * nPts);
    downloadRows(full.data(), rows, which, where);
}

void CLODE::fetch(int which, int rows, cl_double *out, const char *where)
{
    if (nPts) downloadRows(out, rows, which, where);
}

// gathered (NVLink, when sharded) and transposed on the GPU, then one copy to the host
void CLODE::fetchInstanceMajor(int which, int rows, cl_double *out, const char *where)
{
    if (!nPts) return;
    std::vector<clode_sim *> sims;
    for (auto &s : shards()) sims.push_back(s.sim);
    check(clode_gather_rows_instance_major(sims.data(), (int)sims.size(), which, (size_t)rows, (size_t)nPts, out), where);
}

void CLODE::setProblemData(std::vector<cl_double> newX0, std::vector<cl_double> newPars)
{
    if (nVar == 0 || newX0.size() % nVar != 0) {
        lg::info_("Invalid initial condition vector: not a multiple of nVar={}", nVar);
        lg::info_("...Initial conditions were not updated!");
        return;
    }
    if (nPar > 0 && newPars.size() % nPar != 0) {
        lg::info_("Invalid parameter vector: not a multiple of nPar={}", nPar);
        lg::info_("...Parameters were not updated!");
        return;
    }
    cl_int nPtsX0 = (cl_int)(newX0.size() / nVar);
    cl_int nPtsPars = nPar > 0 ? (cl_int)(newPars.size() / nPar) : nPtsX0;
    if (nPtsX0 != nPtsPars) {
        lg::info_("Initial contition and parameter vector dimensions don't match");
        lg::info_("...Expected {} sets of each, recieved {} for x0 and {} for pars", nPts, nPtsX0, nPtsPars);
        lg::info_("...Problem data was not updated!");
        return;
    }
    setNpts(nPtsX0);
    setX0(newX0);
    setPars(newPars);
    lg::debug_("set problem data");
}

void CLODE::setX0(std::vector<cl_double> newX0)
{
    if (newX0.size() == (size_t)nPts * nVar) x0 = newX0;
    setX0(newX0.data(), newX0.size());
}

void CLODE::setX0(const cl_double *newX0, size_t count)
{
    if (count == (size_t)nPts * nVar) {
        if (nVar > 0 && nPts > 0) uploadRows(newX0, nVar, CLODE_BUF_X0, "CLODE::setX0");
        lg::debug_("set X0");
    } else {
        lg::info_("...Initial conditions were not updated!");
    }
}

void CLODE::setPars(std::vector<cl_double> newPars)
{
    const bool ok = newPars.size() == (size_t)nPts * nPar;
    setPars(newPars.data(), newPars.size());
    if (ok) {
        pars = newPars;
        parsOnDeviceOnly = false;
    }
}

void CLODE::setPars(const cl_double *newPars, size_t count)
{
    if (count == (size_t)nPts * nPar) {
        if (nPar > 0 && nPts > 0) uploadRows(newPars, nPar, CLODE_BUF_PARS, "CLODE::setPars");
        parsOnDeviceOnly = true; // the host mirror is refreshed on demand (getPars)
        lg::debug_("set P");
    } else {
        lg::info_("Invalid parameter vector: Expected {}*{} elements, recieved {}", nPts, nPar, count);
        lg::info_("...Parameters were not updated!");
    }
}

const std::vector<cl_double> CLODE::getPars() const
{
    if (parsOnDeviceOnly && nPts && nPar) {
        CLODE *self = const_cast<CLODE *>(this);
        self->downloadRows(self->pars, nPar, CLODE_BUF_PARS, "CLODE::getPars");
        parsOnDeviceOnly = false;
    }
    return pars;
}

void CLODE::setProblemData(const cl_double *newX0, size_t nX0, const cl_double *newPars, size_t nParsValues)
{
    if (nVar == 0 || nX0 % nVar != 0) {
        lg::info_("Invalid initial condition vector: not a multiple of nVar={}", nVar);
        lg::info_("...Initial conditions were not updated!");
        return;
    }
    if (nPar > 0 && nParsValues % nPar != 0) {
        lg::info_("Invalid parameter vector: not a multiple of nPar={}", nPar);
        lg::info_("...Parameters were not updated!");
        return;
    }
    cl_int nPtsX0 = (cl_int)(nX0 / nVar);
    cl_int nPtsPars = nPar > 0 ? (cl_int)(nParsValues / nPar) : nPtsX0;
    if (nPtsX0 != nPtsPars) {
        lg::info_("Initial contition and parameter vector dimensions don't match");
        lg::info_("...Expected {} sets of each, recieved {} for x0 and {} for pars", nPts, nPtsX0, nPtsPars);
        lg::info_("...Problem data was not updated!");
        return;
    }
    setNpts(nPtsX0);
    setX0(newX0, nX0);
    setPars(newPars, nParsValues);
    lg::debug_("set problem data");
}

void CLODE::setTspan(std::vector<cl_double> newTspan)
{
    if (newTspan.size() != 2) throw std::invalid_argument("tspan must have two elements");
    tspan = newTspan;
    for (auto &s : shards()) check(clode_sim_set_tspan(s.sim, tspan[0], tspan[1]), "CLODE::setTspan");
    lg::debug_("set tspan");
}

void CLODE::pushSolverParams()
{
    clode_solver_params c{sp.dt, sp.dtmax, sp.abstol, sp.reltol, sp.max_steps, sp.max_store, sp.nout};
    for (auto &s : shards()) check(clode_sim_set_solver_params(s.sim, &c), "CLODE::setSolverParams");
}

// as in the reference (CLODE.cpp:377-400) the device-side per-instance dt is NOT refreshed here:
// it is (re)filled with sp.dt only when nPts changes
void CLODE::setSolverParams(SolverParams<cl_double> newSp)
{
    sp = newSp;
    // the reference also fills its host copy of dt here (CLODE.cpp:382) and never uploads it; that copy is only ever
    // the landing buffer of getDt(), so the fill (64 MB at 8 Mi instances) is skipped
    pushSolverParams();
    lg::debug_("set SolverParams");
}

// push the host copy of the RNG state to the shards
// CLODE::seedRNG() (CLODE.cpp:420-444): nRNGstate x nPts random 64-bit words
void CLODE::seedRNG()
{
    std::random_device rd;
    std::mt19937_64 gen(rd());
    std::uniform_int_distribution<cl_ulong> dis;
    for (size_t i = 0; i < RNGstate.size(); ++i) RNGstate[i] = dis(gen);
    std::vector<cl_ulong> part;
    for (auto &s : shards()) {
        if (s.count == 0) continue;
        takeShard(RNGstate, (size_t)nPts, 2, s, part);
        check(clode_sim_set_rng_state(s.sim, part.data(), part.size()), "CLODE::seedRNG");
    }
    lg::debug_("set random RNG seed");
}

// CLODE::seedRNG(cl_int) (CLODE.cpp:447-465): word k of the GLOBAL state array is seed + k, whatever the sharding
void CLODE::seedRNG(cl_int mySeed)
{
    for (size_t i = 0; i < RNGstate.size(); ++i) RNGstate[i] = (cl_ulong)(mySeed + (cl_int)i);
    std::vector<cl_ulong> part;
    for (auto &s : shards()) {
        if (s.count == 0) continue;
        takeShard(RNGstate, (size_t)nPts, 2, s, part);
        check(clode_sim_set_rng_state(s.sim, part.data(), part.size()), "CLODE::seedRNG(int mySeed)");
    }
    lg::debug_("set fixed RNG seed");
}

void CLODE::runOnShards(int kernel, int initialize, const char *where)
{
    if (!programBuilt) buildCL();
    if (nPts == 0) throw std::runtime_error(std::string(where) + ": no problem data (call setProblemData first)");
    for (auto &s : shards())
        if (s.count) check(clode_sim_enqueue(s.sim, kernel, initialize), where);
    for (auto &s : shards())
        if (s.count) check(clode_sim_wait(s.sim), where);
}

void CLODE::transient()
{
    runOnShards(CLODE_KERNEL_TRANSIENT, 0, "CLODE::transient");
    lg::info_("run transient");
}

void CLODE::shiftTspan()
{
    setTspan(std::vector<cl_double>({tspan[1], tspan[1] + (tspan[1] - tspan[0])}));
    lg::debug_("shift tspan");
}

void CLODE::shiftX0()
{
    for (auto &s : shards())
        if (s.count) check(clode_sim_shift_x0(s.sim), "CLODE::shiftX0");
    lg::debug_("shift X0");
}

const std::vector<cl_double> CLODE::getX0()
{
    if (nPts) downloadRows(x0, nVar, CLODE_BUF_X0, "CLODE::getX0");
    return x0;
}
const std::vector<cl_double> CLODE::getXf()
{
    if (nPts) downloadRows(xf, nVar, CLODE_BUF_XF, "CLODE::getXf");
    return xf;
}
const std::vector<cl_double> CLODE::getDt()
{
    if (nPts) downloadRows(dt, 1, CLODE_BUF_DT, "CLODE::getDt");
    return dt;
}
const std::vector<cl_double> CLODE::getTf()
{
    if (nPts) downloadRows(tf, 1, CLODE_BUF_TF, "CLODE::getTf");
    return tf;
}

void CLODE::printStatus()
{
    lg::info_("------------------");
    lg::info_("   {}", clRHSfilename);
    lg::info_("   nVar={}", nVar);
    lg::info_("   nPar={}", nPar);
    lg::info_("   nAux={}", nAux);
    lg::info_("   nWiener={}", nWiener);
    lg::info_("Using {} precision.", (clSinglePrecision ? "single" : "double"));
    lg::info_("Using stepper: {} ", stepper);
    lg::info_("Using {} GPU(s), nPts={}", opencl.getDeviceOrdinals().size(), nPts);
}

double CLODE::getLastKernelMilliseconds() const
{
    double worst = 0.0;
    for (auto &s : runtime->shards) {
        float ms = 0.f;
        if (s.sim && s.count && clode_sim_last_kernel_ms(s.sim, &ms) == CLODE_OK) worst = std::max(worst, (double)ms);
    }
    return worst;
}

std::vector<unsigned int> CLODE::getStepCounts()
{
    std::vector<unsigned int> out(nPts), part;
    for (auto &s : shards()) {
        if (s.count == 0) continue;
        part.resize(s.count);
        check(clode_sim_get_steps(s.sim, part.data(), s.count), "CLODE::getStepCounts");
        putShard(out, (size_t)nPts, 1, s, part);
    }
    return out;
}

std::vector<cl_ulong> CLODE::getRNGstate()
{
    std::vector<cl_ulong> part;
    for (auto &s : shards()) {
        if (s.count == 0) continue;
        part.resize(2 * s.count);
        check(clode_sim_get_rng_state(s.sim, part.data(), part.size()), "CLODE::getRNGstate");
        putShard(RNGstate, (size_t)nPts, 2, s, part);
    }
    return RNGstate;
}
