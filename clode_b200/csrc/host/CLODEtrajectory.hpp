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

    int kernelMask() const override { return CLODE_KERNEL_TRANSIENT | CLODE_KERNEL_TRAJECTORY; }
    // per-shard [rows][width][count] -> host [max_store][width][nPts]
    void downloadStored(std::vector<cl_double> &full, int width, int which, const char *where);

public:
    CLODEtrajectory(ProblemInfo prob, std::string stepper, bool clSinglePrecision, OpenCLResource opencl,
                    const std::string clodeRoot);
    CLODEtrajectory(ProblemInfo prob, std::string stepper, bool clSinglePrecision, unsigned int platformID,
                    unsigned int deviceID, const std::string clodeRoot);
    virtual ~CLODEtrajectory();

    void buildCL() override;
    void trajectory();

    std::vector<cl_double> getT();
    std::vector<cl_double> getX();
    std::vector<cl_double> getDx();
    std::vector<cl_double> getAux();
    std::vector<cl_int> getNstored();
};
