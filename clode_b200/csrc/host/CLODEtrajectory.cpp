#include "CLODEtrajectory.hpp"

#include "clode_log.hpp"

#include <algorithm>
#include <stdexcept>

namespace lg = clode_log;

CLODEtrajectory::CLODEtrajectory(ProblemInfo prob, std::string stepper, bool clSinglePrecision, OpenCLResource opencl,
                                 const std::string clodeRoot)
    : CLODE(prob, stepper, clSinglePrecision, opencl, clodeRoot)
{
    lg::debug_("constructor clODEtrajectory");
}

CLODEtrajectory::CLODEtrajectory(ProblemInfo prob, std::string stepper, bool clSinglePrecision, unsigned int platformID,
                                 unsigned int deviceID, const std::string clodeRoot)
    : CLODE(prob, stepper, clSinglePrecision, platformID, deviceID, clodeRoot)
{
    lg::debug_("constructor clODEtrajectory");
}

CLODEtrajectory::~CLODEtrajectory() {}

void CLODEtrajectory::buildCL()
{
    lg::info_("Running CLODEtrajectory buildCL");
    buildProgram();
    lg::debug_("Created trajectory kernels");
}

void CLODEtrajectory::trajectory()
{
    // sizes as in CLODEtrajectory::resizeTrajectoryVariables (CLODEtrajectory.cpp:45-95); the device
    // buffers themselves (with the extra row) are owned by the runtime
    const size_t largest = (size_t)std::max(1, std::max(nVar, nAux)) * (size_t)nPts * sp.max_store * realSize;
    if (largest > opencl.getMaxMemAllocSize()) {
        lg::error_("Storage requested exceeds device maximum variable size. Try reducing storage or nPts.");
        throw std::invalid_argument("nPts*nStoreMax*nVar*realSize or nPts*nStoreMax*nAux*realSize is too big");
    }
    nStoreMax = (cl_int)sp.max_store;
    telements = (size_t)nStoreMax * nPts;
    xelements = (size_t)nVar * telements;
    auxelements = nAux > 0 ? (size_t)nAux * telements : 1;
    runOnShards(CLODE_KERNEL_TRAJECTORY, 0, "CLODEtrajectory::trajectory()");
    lg::debug_("run trajectory");
}

void CLODEtrajectory::downloadStored(std::vector<cl_double> &full, int width, int which, const char *where)
{
    const size_t rows = (size_t)nStoreMax;
    full.assign(std::max<size_t>(rows * width * nPts, 1), 0.0);
    if (width == 0 || nPts == 0 || rows == 0) return;
    if (shards().size() == 1) {
        check(clode_sim_get(shards()[0].sim, which, full.data(), rows * width * nPts), where);
        return;
    }
    std::vector<double> part;
    for (auto &s : shards()) {
        if (s.count == 0) continue;
        part.resize(rows * width * s.count);
        check(clode_sim_get(s.sim, which, part.data(), part.size()), where);
        putShard(full, (size_t)nPts, (int)(rows * width), s, part);
    }
}

std::vector<cl_double> CLODEtrajectory::getT() { downloadStored(t, 1, CLODE_BUF_T, "CLODEtrajectory::getT"); return t; }
std::vector<cl_double> CLODEtrajectory::getX() { downloadStored(x, nVar, CLODE_BUF_X, "CLODEtrajectory::getX"); return x; }
std::vector<cl_double> CLODEtrajectory::getDx() { downloadStored(dx, nVar, CLODE_BUF_DX, "CLODEtrajectory::getDx"); return dx; }
std::vector<cl_double> CLODEtrajectory::getAux() { downloadStored(aux, nAux, CLODE_BUF_AUX, "CLODEtrajectory::getAux"); return aux; }

std::vector<cl_int> CLODEtrajectory::getNstored()
{
    nStored.resize(nPts);
    std::vector<cl_int> part;
    for (auto &s : shards()) {
        if (s.count == 0) continue;
        part.resize(s.count);
        check(clode_sim_get_n_stored(s.sim, part.data(), s.count), "CLODEtrajectory::getNstored");
        putShard(nStored, (size_t)nPts, 1, s, part);
    }
    return nStored;
}
