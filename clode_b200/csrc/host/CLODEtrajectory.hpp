// CLODEtrajectory — clode/cpp/CLODEtrajectory.hpp:21-51 on the B200 runtime.
#pragma once

#include "CLODE.hpp"

#include <string>
#include <vector>

class CLODEtrajectory : public CLODE
{
protected:
    cl_int nStoreMax = 0;
    std::vector<cl_int> nStored;
    std::vector<cl_double> t, x, dx, aux;
    size_t telements = 0, xelements = 0, auxelements = 0;
    unsigned int streamChunk = 0; // > 0: trajectory() runs in launches of this many stored points (streamed to the host)
    bool streamed = false;        // t, x, dx, aux already hold the last run's result

    int kernelMask() const override { return CLODE_KERNEL_TRANSIENT | CLODE_KERNEL_TRAJECTORY; }
    // per-shard [rows][width][count] -> host [max_store][width][nPts]
    void downloadStored(std::vector<cl_double> &full, int width, int which, const char *where);
    void trajectoryStreamed();

public:
    CLODEtrajectory(ProblemInfo prob, std::string stepper, bool clSinglePrecision, OpenCLResource opencl,
                    const std::string clodeRoot);
    CLODEtrajectory(ProblemInfo prob, std::string stepper, bool clSinglePrecision, unsigned int platformID,
                    unsigned int deviceID, const std::string clodeRoot);
    virtual ~CLODEtrajectory();

    void buildCL() override;
    void trajectory();
    // Not in the reference (its TODO at CLODEtrajectory.cpp:47): integrate in chunks of `rows` stored points and
    // copy each chunk to the host while the next one integrates; the device then holds two chunks instead of
    // nPts*max_store points.  0 restores the single launch.  Results of getT/getX/getDx/getAux are unchanged.
    void setStreamChunk(unsigned int rows) { streamChunk = rows; }
    unsigned int getStreamChunk() const { return streamChunk; }

    std::vector<cl_double> getT();
    std::vector<cl_double> getX();
    std::vector<cl_double> getDx();
    std::vector<cl_double> getAux();
    std::vector<cl_int> getNstored();
};
