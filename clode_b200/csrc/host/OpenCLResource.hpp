// OpenCLResource — device selection object of the clODE C++ API, re-targeted at CUDA.
//
// Same name, constructors, queries and free functions as clode/cpp/OpenCLResource.hpp:73-144 so the
// binding layers (pybind module, mex files, samples) keep working; what it holds is no longer an
// OpenCL context + queues but a list of CUDA device ordinals.  "Platform 0" is the CUDA driver;
// `OpenCLResource(platformID, std::vector<deviceIDs>)` — which the reference accepts but never uses
// beyond device 0 (OpenCLResource.hpp:104,110) — selects the GPUs an ensemble is sharded over.
// Program building and queues live in libclode_rt (include/clode_rt.h).
#pragma once

#include "clODE_struct_defs.hpp"

#include <string>
#include <vector>

typedef cl_device_type cl_deviceType;

enum cl_vendor { VENDOR_ANY = 0, VENDOR_NVIDIA, VENDOR_AMD, VENDOR_INTEL };

// numeric values of the OpenCL CL_DEVICE_TYPE_* bit masks
enum e_cl_device_type {
    DEVICE_TYPE_ALL = 0xFFFFFFFF,
    DEVICE_TYPE_CPU = 1 << 1,
    DEVICE_TYPE_GPU = 1 << 2,
    DEVICE_TYPE_ACCELERATOR = 1 << 3,
    DEVICE_TYPE_DEFAULT = 1 << 0,
    DEVICE_TYPE_CUSTOM = 1 << 4
};

typedef struct deviceInfo {
    std::string name;
    std::string vendor;
    std::string version;
    cl_device_type devType;
    std::string devTypeStr;
    cl_uint computeUnits;
    cl_uint maxClock;
    size_t maxWorkGroupSize;
    cl_ulong deviceMemSize;
    cl_ulong maxMemAllocSize;
    std::string extensions;
    bool doubleSupport;
    cl_bool deviceAvailable;
} deviceInfo;

typedef struct platformInfo {
    std::string name;
    std::string vendor;
    std::string version;
    std::vector<deviceInfo> device_info;
    unsigned int nDevices;
} platformInfo;

class OpenCLResource
{
    platformInfo platform_info;
    std::vector<int> deviceOrdinals; // CUDA device indices of this resource

    void getPlatformAndDevices(cl_deviceType type = DEVICE_TYPE_ALL, cl_vendor vendor = VENDOR_ANY);
    void getPlatformAndDevices(unsigned int platformID, std::vector<unsigned int> deviceID);

public:
    OpenCLResource();
    OpenCLResource(cl_deviceType type);
    OpenCLResource(cl_vendor vendor);
    OpenCLResource(cl_deviceType type, cl_vendor vendor);
    OpenCLResource(e_cl_device_type type, cl_vendor vendor);
    OpenCLResource(int argc, char **argv); // "--device gpu/cpu/accel" and/or "--vendor amd/intel/nvidia"
    OpenCLResource(unsigned int platformID, unsigned int deviceID);
    OpenCLResource(unsigned int platformID, std::vector<unsigned int> deviceID);

    cl_int error = 0;

    bool getDoubleSupport(cl_uint deviceID = 0) { return platform_info.device_info.at(deviceID).doubleSupport; }
    cl_ulong getMaxMemAllocSize(cl_uint deviceID = 0) { return platform_info.device_info.at(deviceID).maxMemAllocSize; }
    std::string getDeviceCLVersion(cl_uint deviceID = 0) { return platform_info.device_info.at(deviceID).version; }
    cl_device_type getDeviceType(cl_uint deviceID = 0) { return platform_info.device_info.at(deviceID).devType; }

    // CUDA device ordinals selected by this resource (new; used by CLODE to create its runtime objects)
    const std::vector<int> &getDeviceOrdinals() const { return deviceOrdinals; }

    void print();
};

std::vector<platformInfo> queryOpenCL();
void printOpenCL();
void printOpenCL(std::vector<platformInfo>);
void printPlatformInfo(platformInfo pi);
void printDeviceInfo(deviceInfo di);

// status code of the runtime (include/clode_rt.h `clode_status`) as text; replaces the OpenCL error table
std::string CLErrorString(cl_int cl_error);

std::string read_file(std::string filename);
