#include "CLODEtrajectory.hpp"

#include "clode_log.hpp"

#include <algorithm>
#include <stdexcept>
#include <thread>

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
    streamed = false;
    if (streamChunk > 0) {
        trajectoryStreamed();
        return;
    }
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

void CLODEtrajectory::trajectoryStreamed()
{
    nStoreMax = (cl_int)sp.max_store;
    telements = (size_t)nStoreMax * nPts;
    xelements = (size_t)nVar * telements;
    auxelements = nAux > 0 ? (size_t)nAux * telements : 1;
    const size_t rows = (size_t)nStoreMax;
    t.assign(std::max<size_t>(telements, 1), 0.0);
    x.assign(std::max<size_t>(xelements, 1), 0.0);
    dx.assign(std::max<size_t>(xelements, 1), 0.0);
    aux.assign(std::max<size_t>(auxelements, 1), 0.0);
    nStored.assign(nPts, 0);
    if (nPts == 0 || rows == 0) { streamed = true; return; }
    const char *where = "CLODEtrajectory::trajectory() [streamed]";
    if (shards().size() == 1) {
        check(clode_sim_trajectory_stream(shards()[0].sim, streamChunk, t.data(), x.data(), dx.data(),
                                          nAux > 0 ? aux.data() : nullptr, nStored.data()), where);
        streamed = true;
        return;
    }
    // several GPUs: every shard streams into its own compact arrays concurrently, then the columns are interleaved
    struct Part { std::vector<double> t, x, dx, aux; std::vector<cl_int> n; int rc = 0; std::string err; };
    std::vector<Part> parts(shards().size());
    std::vector<std::thread> workers;
    for (size_t g = 0; g < shards().size(); ++g) {
        auto &s = shards()[g];
        if (s.count == 0) continue;
        Part &p = parts[g];
        p.t.assign(rows * s.count, 0.0);
        p.x.assign(rows * nVar * s.count, 0.0);
        p.dx.assign(rows * nVar * s.count, 0.0);
        p.aux.assign(std::max<size_t>(rows * nAux * s.count, 1), 0.0);
        p.n.assign(s.count, 0);
        workers.emplace_back([&p, &s, this] {
            p.rc = clode_sim_trajectory_stream(s.sim, streamChunk, p.t.data(), p.x.data(), p.dx.data(),
                                               nAux > 0 ? p.aux.data() : nullptr, p.n.data());
            if (p.rc) p.err = clode_last_error();
        });
    }
    for (auto &w : workers) w.join();
    for (size_t g = 0; g < shards().size(); ++g) {
        auto &s = shards()[g];
        if (s.count == 0) continue;
        Part &p = parts[g];
        if (p.rc) throw std::runtime_error(std::string(where) + ": " + p.err);
        putShard(t, (size_t)nPts, (int)rows, s, p.t);
        putShard(x, (size_t)nPts, (int)(rows * nVar), s, p.x);
        putShard(dx, (size_t)nPts, (int)(rows * nVar), s, p.dx);
        if (nAux > 0) putShard(aux, (size_t)nPts, (int)(rows * nAux), s, p.aux);
        putShard(nStored, (size_t)nPts, 1, s, p.n);
    }
    streamed = true;
}

void CLODEtrajectory::downloadStored(std::vector<cl_double> &full, int width, int which, const char *where)
{
    if (streamed) return; // trajectory() already delivered the host arrays
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
    if (streamed) return nStored;
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
