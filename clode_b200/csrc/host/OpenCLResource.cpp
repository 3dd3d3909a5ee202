#include "OpenCLResource.hpp"

#include "clode_log.hpp"
#include "clode_rt.h"

#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace lg = clode_log;

static deviceInfo cudaDeviceInfo(int ordinal)
{
    clode_device_info ci;
    if (clode_device_get_info(ordinal, &ci) != CLODE_OK) throw std::runtime_error(clode_last_error());
    deviceInfo di;
    di.name = ci.name;
    di.vendor = "NVIDIA Corporation";
    di.version = "CUDA sm_" + std::to_string(ci.cc_major) + std::to_string(ci.cc_minor) + " (driver " +
                 std::to_string(ci.driver_version / 1000) + "." + std::to_string((ci.driver_version % 1000) / 10) + ")";
    di.devType = DEVICE_TYPE_GPU;
    di.devTypeStr = "GPU";
    di.computeUnits = ci.multiprocessors;
    di.maxClock = ci.clock_mhz;
    di.maxWorkGroupSize = ci.max_threads_per_block;
    di.deviceMemSize = ci.total_memory;
    di.maxMemAllocSize = ci.max_alloc;
    di.extensions = "cl_khr_fp64 (native FP64), NVRTC JIT";
    di.doubleSupport = true;
    di.deviceAvailable = 1;
    return di;
}

static platformInfo cudaPlatform(const std::vector<int> &ordinals)
{
    platformInfo pi;
    pi.name = "NVIDIA CUDA (clode_b200 runtime; takes the place of the OpenCL platform)";
    pi.vendor = "NVIDIA Corporation";
    pi.version = clode_version();
    for (int o : ordinals) pi.device_info.push_back(cudaDeviceInfo(o));
    pi.nDevices = (unsigned int)pi.device_info.size();
    return pi;
}

std::vector<platformInfo> queryOpenCL()
{
    int count = 0;
    if (clode_device_count(&count) != CLODE_OK) throw std::runtime_error(clode_last_error());
    std::vector<int> all;
    for (int i = 0; i < count; ++i) all.push_back(i);
    return {cudaPlatform(all)};
}

// selection by type/vendor (OpenCLResource.cpp:128-196): only GPUs of vendor NVIDIA exist here
void OpenCLResource::getPlatformAndDevices(cl_deviceType type, cl_vendor vendor)
{
    if (vendor != VENDOR_ANY && vendor != VENDOR_NVIDIA)
        throw std::runtime_error("No OpenCL platforms were found for the requested vendor (this runtime drives NVIDIA GPUs only)");
    const bool wants_gpu = type == DEVICE_TYPE_ALL || (type & DEVICE_TYPE_GPU) || (type & DEVICE_TYPE_DEFAULT);
    if (!wants_gpu) throw std::runtime_error("No devices of the requested type were found (this runtime drives GPUs only)");
    int count = 0;
    if (clode_device_count(&count) != CLODE_OK) throw std::runtime_error(clode_last_error());
    if (count == 0) throw std::runtime_error("No CUDA devices were found");
    deviceOrdinals = {0}; // the reference also settles on the first matching device
    platform_info = cudaPlatform(deviceOrdinals);
}

void OpenCLResource::getPlatformAndDevices(unsigned int platformID, std::vector<unsigned int> deviceIDs)
{
    if (platformID != 0) throw std::out_of_range("Specified platformID exceeds number of available platforms");
    int count = 0;
    if (clode_device_count(&count) != CLODE_OK) throw std::runtime_error(clode_last_error());
    deviceOrdinals.clear();
    for (unsigned int id : deviceIDs) {
        if ((int)id >= count) throw std::out_of_range("Specified deviceID exceeds the number devices on the selected platform");
        deviceOrdinals.push_back((int)id);
    }
    if (deviceOrdinals.empty())
        for (int i = 0; i < count; ++i) deviceOrdinals.push_back(i); // "default uses all available devices"
    platform_info = cudaPlatform(deviceOrdinals);
}

OpenCLResource::OpenCLResource() { getPlatformAndDevices(DEVICE_TYPE_DEFAULT, VENDOR_ANY); }
OpenCLResource::OpenCLResource(cl_deviceType type) { getPlatformAndDevices(type, VENDOR_ANY); }
OpenCLResource::OpenCLResource(cl_vendor vendor) { getPlatformAndDevices(DEVICE_TYPE_DEFAULT, vendor); }
OpenCLResource::OpenCLResource(cl_deviceType type, cl_vendor vendor) { getPlatformAndDevices(type, vendor); }
OpenCLResource::OpenCLResource(e_cl_device_type type, cl_vendor vendor) { getPlatformAndDevices((cl_deviceType)type, vendor); }
OpenCLResource::OpenCLResource(unsigned int platformID, unsigned int deviceID)
{
    getPlatformAndDevices(platformID, std::vector<unsigned int>{deviceID});
}
OpenCLResource::OpenCLResource(unsigned int platformID, std::vector<unsigned int> deviceIDs)
{
    getPlatformAndDevices(platformID, deviceIDs);
}

// command-line selection, OpenCLResource.cpp:54-111
OpenCLResource::OpenCLResource(int argc, char **argv)
{
    cl_deviceType type = DEVICE_TYPE_DEFAULT;
    cl_vendor vendor = VENDOR_ANY;
    for (int i = 1; i + 1 < argc; ++i) {
        if (!std::strcmp(argv[i], "--device")) {
            const char *v = argv[++i];
            if (!std::strcmp(v, "cpu")) type = DEVICE_TYPE_CPU;
            else if (!std::strcmp(v, "gpu")) type = DEVICE_TYPE_GPU;
            else if (!std::strcmp(v, "accel")) type = DEVICE_TYPE_ACCELERATOR;
        } else if (!std::strcmp(argv[i], "--vendor")) {
            const char *v = argv[++i];
            if (!std::strcmp(v, "amd")) vendor = VENDOR_AMD;
            else if (!std::strcmp(v, "intel")) vendor = VENDOR_INTEL;
            else if (!std::strcmp(v, "nvidia")) vendor = VENDOR_NVIDIA;
        }
    }
    getPlatformAndDevices(type, vendor);
}

void OpenCLResource::print() { printPlatformInfo(platform_info); }

void printOpenCL() { printOpenCL(queryOpenCL()); }

void printOpenCL(std::vector<platformInfo> pinfo)
{
    lg::info_("Querying OpenCL-equivalent platforms (CUDA)...");
    lg::info_("Number of platforms found: {}", pinfo.size());
    for (auto &p : pinfo) printPlatformInfo(p);
}

void printPlatformInfo(platformInfo pi)
{
    lg::info_("Platform (OpenCL API slot): {}", pi.name);
    lg::info_("Vendor: {}", pi.vendor);
    lg::info_("Version: {}", pi.version);
    for (unsigned int j = 0; j < pi.nDevices; ++j) {
        lg::info_("Device {}:", j);
        printDeviceInfo(pi.device_info[j]);
    }
}

void printDeviceInfo(deviceInfo di)
{
    lg::info_("  Name: {}", di.name);
    lg::info_("  Type: {}", di.devTypeStr);
    lg::info_("  Version: {}", di.version);
    lg::info_("  Compute units (SMs): {}", di.computeUnits);
    lg::info_("  Clock frequency: {} MHz", di.maxClock);
    lg::info_("  Maximum memory allocation size: {} MB", di.maxMemAllocSize / (1024 * 1024));
    lg::info_("  Global memory: {} MB", di.deviceMemSize / (1024 * 1024));
    lg::info_("  Max work group size: {}", di.maxWorkGroupSize);
    lg::info_("  Supports double precision: {}", di.doubleSupport ? "yes" : "no");
}

std::string CLErrorString(cl_int e)
{
    switch (e) {
    case CLODE_OK: return "CLODE_OK";
    case CLODE_ERR_INVALID: return "CLODE_ERR_INVALID";
    case CLODE_ERR_NO_DRIVER: return "CLODE_ERR_NO_DRIVER";
    case CLODE_ERR_CUDA: return "CLODE_ERR_CUDA";
    case CLODE_ERR_BUILD: return "CLODE_ERR_BUILD";
    case CLODE_ERR_STATE: return "CLODE_ERR_STATE";
    case CLODE_ERR_MEMORY: return "CLODE_ERR_MEMORY";
    }
    return "unknown error";
}

std::string read_file(std::string filename)
{
    std::ifstream f(filename, std::ios::in | std::ios::binary);
    if (!f) throw std::runtime_error("Could not open file: " + filename);
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
