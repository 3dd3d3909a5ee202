// queryOpenCL — MATLAB entry point listing the compute devices (replaces matlab/queryOpenCL.cpp of the reference):
// a struct array with the reference's field names, one element per (platform, device); here platform 0 = CUDA.
#include "mex.h"

#include "OpenCLResource.hpp"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nlhs > 1) mexErrMsgIdAndTxt("clODE:args", "more than one output argument is not supported");
    const std::vector<platformInfo> platforms = queryOpenCL();
    size_t total = 0;
    for (const platformInfo &p : platforms) total += p.nDevices;
    const char *fields[] = {"platformID", "deviceID", "name", "type", "vendor", "computeUnits", "maxClock", "memSize",
                            "doubleSupport", "available"};
    mwSize dims[2] = {(mwSize)total, 1};
    plhs[0] = mxCreateStructArray(2, dims, 10, fields);
    size_t ix = 0;
    for (size_t i = 0; i < platforms.size(); ++i)
        for (size_t j = 0; j < platforms[i].nDevices; ++j, ++ix) {
            const deviceInfo &d = platforms[i].device_info[j];
            mxSetField(plhs[0], ix, "platformID", mxCreateDoubleScalar((double)i));
            mxSetField(plhs[0], ix, "deviceID", mxCreateDoubleScalar((double)j));
            mxSetField(plhs[0], ix, "name", mxCreateString(d.name.c_str()));
            mxSetField(plhs[0], ix, "type", mxCreateString(d.devTypeStr.c_str()));
            mxSetField(plhs[0], ix, "vendor", mxCreateString(d.vendor.c_str()));
            mxSetField(plhs[0], ix, "computeUnits", mxCreateDoubleScalar(d.computeUnits));
            mxSetField(plhs[0], ix, "maxClock", mxCreateDoubleScalar(d.maxClock));
            mxSetField(plhs[0], ix, "memSize", mxCreateDoubleScalar((double)(d.deviceMemSize / 1024 / 1024)));
            mxSetField(plhs[0], ix, "doubleSupport", mxCreateDoubleScalar(d.doubleSupport));
            mxSetField(plhs[0], ix, "available", mxCreateDoubleScalar(d.deviceAvailable));
        }
}
