// CLODEfeatures — clode/cpp/CLODEfeatures.hpp:23-76 on the B200 runtime.
#pragma once

#include "CLODE.hpp"

#include <map>
#include <string>
#include <vector>

// per-observer facts the host needs (clode/cpp/observers.cl:51-58 `struct ObserverInfo`)
struct ObserverInfo {
    std::string define;
    std::vector<std::string> featureNames;
};

// clode/cpp/observers.cl:122-143
void getObserverDefineMap(const ProblemInfo pi, const unsigned int fVarIx, const unsigned int eVarIx,
                          const unsigned int nStoredEvents, std::map<std::string, ObserverInfo> &observerDefineMap,
                          std::vector<std::string> &availableObserverNames);

class CLODEfeatures : public CLODE
{
protected:
    std::string observer;
    std::map<std::string, ObserverInfo> observerDefineMap;
    std::vector<std::string> featureNames;
    std::vector<std::string> availableObserverNames;

    int nFeatures = 0;
    std::vector<cl_double> F;
    ObserverParams<cl_double> op{};
    bool observerInitialized = false;

    std::string observerBuildOpts;
    std::string observerName;

    // values baked into the compiled program; a change triggers a rebuild before the next launch
    unsigned int builtFVarIx = 0, builtEVarIx = 0, builtStoreEvents = 0;
    std::string builtObserver;

    int kernelMask() const override { return CLODE_KERNEL_TRANSIENT | CLODE_KERNEL_FEATURES; }
    void fillProgramDesc(clode_program_desc &d) const override;
    void onNptsChanged() override { observerInitialized = false; }
    void updateObserverDefineMap();
    void pushObserverParams();
    void rebuildIfNeeded();

public:
    CLODEfeatures(ProblemInfo prob, std::string stepper, std::string observer, ObserverParams<cl_double> op,
                  bool clSinglePrecision, OpenCLResource opencl, const std::string clodeRoot);
    CLODEfeatures(ProblemInfo prob, std::string stepper, std::string observer, ObserverParams<cl_double> op,
                  bool clSinglePrecision, unsigned int platformID, unsigned int deviceID, const std::string clodeRoot);
    virtual ~CLODEfeatures();

    void buildCL() override;

    void setObserverParams(ObserverParams<cl_double> newOp);
    void setObserver(std::string newObserver);

    void initializeObserver();
    void features();
    void features(bool reinitialize_observer);
    bool isObserverInitialized() { return observerInitialized; }

    const ObserverParams<cl_double> getObserverParams() const { return op; }
    const std::string getObserverName() const { return observerName; }
    const std::vector<cl_double> getF();
    const int getNFeatures() const { return nFeatures; }
    const std::vector<std::string> getFeatureNames() const { return featureNames; }
    const std::vector<std::string> getAvailableObservers() const { return availableObserverNames; }
};
