#include "CLODEfeatures.hpp"

#include "clode_log.hpp"

#include <stdexcept>

namespace lg = clode_log;

// ---- feature-name manifests: the `getObserverInfo_*` functions of clode/cpp/observers/*.clh -------------
static void pushVarBlock(std::vector<std::string> &f, const ProblemInfo &pi, bool nhood2)
{
    for (int j = 0; j < pi.nVar; ++j) {
        const std::string &v = pi.varNames[j];
        f.push_back("max " + v);
        f.push_back("min " + v);
        f.push_back("mean " + v);
        if (nhood2) {
            f.push_back("range " + v);
            f.push_back("nhood center " + v);
        }
        f.push_back("max d" + v + "/dt");
        f.push_back("min d" + v + "/dt");
    }
    for (int j = 0; j < pi.nAux; ++j) {
        const std::string &a = pi.auxNames[j];
        f.push_back("max " + a);
        f.push_back("min " + a);
        f.push_back("mean " + a);
    }
}

static void pushStats(std::vector<std::string> &f, std::initializer_list<const char *> quantities)
{
    for (const char *q : quantities)
        for (const char *s : {"max ", "min ", "mean "}) f.push_back(std::string(s) + q);
}

void getObserverDefineMap(const ProblemInfo pi, const unsigned int fVarIx, const unsigned int eVarIx,
                          const unsigned int nStoredEvents, std::map<std::string, ObserverInfo> &observerDefineMap,
                          std::vector<std::string> &availableObserverNames)
{
    (void)eVarIx;
    std::map<std::string, ObserverInfo> m;
    {   // observer_basic.clh:6-15
        ObserverInfo oi;
        oi.define = "USE_OBSERVER_BASIC";
        std::string v = fVarIx < pi.varNames.size() ? pi.varNames[fVarIx] : "";
        oi.featureNames = {"max " + v, "min " + v, "mean " + v, "max d" + v + "/dt", "min d" + v + "/dt", "step count"};
        m["basic"] = oi;
    }
    {   // observer_basic_allVar.clh:7-29
        ObserverInfo oi;
        oi.define = "USE_OBSERVER_BASIC_ALLVAR";
        pushVarBlock(oi.featureNames, pi, false);
        oi.featureNames.push_back("step count");
        m["basicall"] = oi;
    }
    {   // observer_local_maximum.clh:9-48
        ObserverInfo oi;
        oi.define = "USE_OBSERVER_LOCAL_MAX";
        pushStats(oi.featureNames, {"IMI", "amplitude"});
        pushVarBlock(oi.featureNames, pi, false);
        for (unsigned int j = 0; j < nStoredEvents; ++j) {
            const std::string k = std::to_string(j);
            oi.featureNames.push_back("localmax event time " + k);
            oi.featureNames.push_back("localmax event evar " + k);
            oi.featureNames.push_back("localmin event time " + k);
            oi.featureNames.push_back("localmin event evar " + k);
        }
        oi.featureNames.push_back("event count");
        oi.featureNames.push_back("step count");
        m["localmax"] = oi;
    }
    {   // observer_neighborhood_1.clh:6-38
        ObserverInfo oi;
        oi.define = "USE_OBSERVER_NEIGHBORHOOD_1";
        pushStats(oi.featureNames, {"period", "peaks"});
        pushVarBlock(oi.featureNames, pi, false);
        for (const char *s : {"period count", "step count", "max dt", "min dt", "mean dt"}) oi.featureNames.push_back(s);
        m["nhood1"] = oi;
    }
    {   // observer_neighborhood_2.clh:6-43
        ObserverInfo oi;
        oi.define = "USE_OBSERVER_NEIGHBORHOOD_2";
        pushStats(oi.featureNames, {"period", "peaks"});
        pushVarBlock(oi.featureNames, pi, true);
        for (unsigned int j = 0; j < nStoredEvents; ++j) oi.featureNames.push_back("nhood event time " + std::to_string(j));
        for (const char *s : {"event count", "step count", "max dt", "min dt", "mean dt"}) oi.featureNames.push_back(s);
        m["nhood2"] = oi;
    }
    {   // observer_threshold_2.clh:7-56
        ObserverInfo oi;
        oi.define = "USE_OBSERVER_THRESHOLD_2";
        pushStats(oi.featureNames, {"period", "peaks", "upDuration", "downDuration", "duty", "activeDip"});
        pushVarBlock(oi.featureNames, pi, false);
        for (unsigned int j = 0; j < nStoredEvents; ++j) {
            oi.featureNames.push_back("up event time " + std::to_string(j));
            oi.featureNames.push_back("down event time " + std::to_string(j));
        }
        for (const char *s : {"event count", "step count", "max dt", "min dt", "mean dt"}) oi.featureNames.push_back(s);
        m["thresh2"] = oi;
    }
    std::vector<std::string> names;
    for (auto const &e : m) names.push_back(e.first);
    observerDefineMap = m;
    availableObserverNames = names;
}

// ---------------------------------------------------------------------------------------------------
CLODEfeatures::CLODEfeatures(ProblemInfo prob, std::string stepper, std::string observer, ObserverParams<cl_double> op,
                             bool clSinglePrecision, OpenCLResource opencl, const std::string clodeRoot)
    : CLODE(prob, stepper, clSinglePrecision, opencl, clodeRoot), observer(observer)
{
    this->op = op;
    updateObserverDefineMap();
    if (observerDefineMap.find(observer) == observerDefineMap.end()) {
        lg::warn_("unknown observer: {}. Using basic", observer);
        this->observer = "basic";
        updateObserverDefineMap();
    }
    lg::debug_("constructor clODEfeatures");
}

CLODEfeatures::CLODEfeatures(ProblemInfo prob, std::string stepper, std::string observer, ObserverParams<cl_double> op,
                             bool clSinglePrecision, unsigned int platformID, unsigned int deviceID,
                             const std::string clodeRoot)
    : CLODEfeatures(prob, stepper, observer, op, clSinglePrecision, OpenCLResource(platformID, deviceID), clodeRoot)
{
}

CLODEfeatures::~CLODEfeatures() {}

void CLODEfeatures::fillProgramDesc(clode_program_desc &d) const
{
    d.observer = observer.c_str();
    d.f_var_ix = (int)op.fVarIx;
    d.e_var_ix = (int)op.eVarIx;
    d.n_store_events = (int)op.maxEventTimestamps;
}

void CLODEfeatures::buildCL()
{
    lg::info_("Running CLODEFeatures buildCL");
    observerBuildOpts = " -D" + observerDefineMap.at(observer).define;
    observerBuildOpts += " -DN_STORE_EVENTS=" + std::to_string((long long)op.maxEventTimestamps);
    buildProgram();
    builtFVarIx = op.fVarIx;
    builtEVarIx = op.eVarIx;
    builtStoreEvents = op.maxEventTimestamps;
    builtObserver = observer;
    observerInitialized = false;
    pushObserverParams();
    lg::debug_("created features kernels");
    lg::debug_("Using observer: {}", observer);
}

// fVarIx / eVarIx / N_STORE_EVENTS / the observer are compile-time constants of the kernels
void CLODEfeatures::rebuildIfNeeded()
{
    if (!programBuilt || builtFVarIx != op.fVarIx || builtEVarIx != op.eVarIx || builtStoreEvents != op.maxEventTimestamps ||
        builtObserver != observer) {
        const cl_int keep = nPts;
        buildCL();
        if (keep > 0 && nPts == 0) { // device data was dropped: restore it from the host copies
            std::vector<cl_double> x0c = x0, pc = pars;
            setProblemData(x0c, pc);
        }
    }
}

void CLODEfeatures::setObserver(std::string newObserver)
{
    if (observerDefineMap.find(newObserver) != observerDefineMap.end()) {
        observer = newObserver;
        updateObserverDefineMap();
    } else {
        lg::warn_("unknown observer: {}. Observer method unchanged", newObserver);
    }
    lg::debug_("set observer");
}

void CLODEfeatures::pushObserverParams()
{
    clode_observer_params c{op.eVarIx, op.fVarIx, op.maxEventCount, op.maxEventTimestamps, op.minXamp, op.minIMI,
                            op.nHoodRadius, op.xUpThresh, op.xDownThresh, op.dxUpThresh, op.dxDownThresh, op.eps_dx};
    for (auto &s : shards()) check(clode_sim_set_observer_params(s.sim, &c), "CLODEfeatures::setObserverParams");
}

void CLODEfeatures::setObserverParams(ObserverParams<cl_double> newOp)
{
    op = newOp;
    pushObserverParams();
    updateObserverDefineMap();
    lg::debug_("set observer params");
}

void CLODEfeatures::updateObserverDefineMap()
{
    getObserverDefineMap(prob, op.fVarIx, op.eVarIx, op.maxEventTimestamps, observerDefineMap, availableObserverNames);
    auto it = observerDefineMap.find(observer);
    if (it == observerDefineMap.end()) return;
    observerName = observer;
    observerBuildOpts = " -D" + it->second.define;
    observerBuildOpts += " -DN_STORE_EVENTS=" + std::to_string((long long)op.maxEventTimestamps);
    nFeatures = (int)it->second.featureNames.size();
    featureNames = it->second.featureNames;
}

void CLODEfeatures::initializeObserver()
{
    rebuildIfNeeded();
    if (nPts == 0) throw std::runtime_error("CLODEfeatures::initializeObserver: no problem data");
    for (auto &s : shards())
        if (s.count) check(clode_sim_initialize_observer(s.sim), "CLODEfeatures::initializeObserver");
    observerInitialized = true;
    lg::debug_("run initializeObserver");
}

void CLODEfeatures::features(bool reinitialize_observer)
{
    observerInitialized = !reinitialize_observer;
    features();
}

void CLODEfeatures::features()
{
    rebuildIfNeeded();
    runOnShards(CLODE_KERNEL_FEATURES, observerInitialized ? 0 : 1, "CLODEfeatures::features");
    observerInitialized = true;
    lg::debug_("run features");
}

const std::vector<cl_double> CLODEfeatures::getF()
{
    if (nPts) downloadRows(F, nFeatures, CLODE_BUF_F, "CLODEfeatures::getF");
    return F;
}
