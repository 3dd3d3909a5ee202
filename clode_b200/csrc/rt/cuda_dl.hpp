// cuda_dl.hpp — CUDA driver API and NVRTC loaded at run time with dlopen.
//
// libclode_rt.so has no link-time dependency on libcuda/libnvrtc, so it loads (and its
// exported symbols can be checked) on a machine without a GPU driver; any call that needs
// the driver fails loudly with CLODE_ERR_NO_DRIVER instead of falling back to the CPU.
#pragma once

#include <cuda.h>
#include <nvrtc.h>

#include <dlfcn.h>

#include <mutex>
#include <string>

namespace clode {

#define CLODE_STR2(x) #x
#define CLODE_STR(x) CLODE_STR2(x)

// name, as written in source; the cuda.h macros map e.g. cuMemAlloc -> cuMemAlloc_v2 and the
// stringification below happens after that expansion, so the versioned symbol is looked up.
#define CLODE_DRIVER_FUNCS(X)                                                                     \
    X(cuInit) X(cuDriverGetVersion) X(cuGetErrorString) X(cuGetErrorName)                         \
    X(cuDeviceGetCount) X(cuDeviceGet) X(cuDeviceGetName) X(cuDeviceGetAttribute)                 \
    X(cuDeviceTotalMem) X(cuDevicePrimaryCtxRetain) X(cuDevicePrimaryCtxRelease)                  \
    X(cuCtxPushCurrent) X(cuCtxPopCurrent)                                                        \
    X(cuMemAlloc) X(cuMemFree) X(cuMemcpyHtoD) X(cuMemcpyDtoH) X(cuMemcpyDtoDAsync)               \
    X(cuMemcpyHtoDAsync) X(cuMemcpyDtoHAsync) X(cuMemsetD8Async) X(cuMemsetD32Async)              \
    X(cuMemHostAlloc) X(cuMemFreeHost) X(cuPointerGetAttribute) X(cuMemcpyPeerAsync)              \
    X(cuCtxEnablePeerAccess) X(cuDeviceCanAccessPeer) X(cuStreamWaitEvent)                        \
    X(cuStreamCreate) X(cuStreamDestroy) X(cuStreamSynchronize)                                   \
    X(cuEventCreate) X(cuEventDestroy) X(cuEventRecord) X(cuEventSynchronize) X(cuEventElapsedTime) \
    X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction) X(cuModuleGetGlobal)                                  \
    X(cuFuncGetAttribute) X(cuFuncSetAttribute) X(cuLaunchKernel)                                 \
    X(cuOccupancyMaxActiveBlocksPerMultiprocessor)

#define CLODE_NVRTC_FUNCS(X)                                                                      \
    X(nvrtcVersion) X(nvrtcGetErrorString) X(nvrtcCreateProgram) X(nvrtcDestroyProgram)           \
    X(nvrtcCompileProgram) X(nvrtcGetCUBINSize) X(nvrtcGetCUBIN) X(nvrtcGetProgramLogSize)        \
    X(nvrtcGetProgramLog) X(nvrtcGetPTXSize) X(nvrtcGetPTX)

struct DriverApi {
#define X(name) decltype(&::name) name = nullptr;
    CLODE_DRIVER_FUNCS(X)
#undef X
    CUresult (*cuFuncLoad_opt)(CUfunction) = nullptr; // CUDA >= 12.4: load a lazily-loaded kernel now (optional)
    void *handle = nullptr;
    std::string error;

    bool load()
    {
        const char *names[] = {"libcuda.so.1", "libcuda.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("cannot load the CUDA driver (libcuda.so.1): ") + dlerror();
            return false;
        }
#define X(fn)                                                                  \
    fn = reinterpret_cast<decltype(fn)>(dlsym(handle, CLODE_STR(fn)));         \
    if (!fn) {                                                                 \
        error = std::string("CUDA driver lacks symbol ") + CLODE_STR(fn);      \
        return false;                                                          \
    }
        CLODE_DRIVER_FUNCS(X)
#undef X
        cuFuncLoad_opt = reinterpret_cast<decltype(cuFuncLoad_opt)>(dlsym(handle, "cuFuncLoad"));
        CUresult r = cuInit(0);
        if (r != CUDA_SUCCESS) {
            const char *s = nullptr;
            cuGetErrorString(r, &s);
            error = std::string("cuInit failed: ") + (s ? s : "unknown error");
            return false;
        }
        return true;
    }
};

struct NvrtcApi {
#define X(name) decltype(&::name) name = nullptr;
    CLODE_NVRTC_FUNCS(X)
#undef X
    void *handle = nullptr;
    std::string error;

    bool load()
    {
        const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("cannot load NVRTC (libnvrtc.so.12): ") + dlerror();
            return false;
        }
#define X(fn)                                                                  \
    fn = reinterpret_cast<decltype(fn)>(dlsym(handle, CLODE_STR(fn)));         \
    if (!fn) {                                                                 \
        error = std::string("NVRTC lacks symbol ") + CLODE_STR(fn);            \
        return false;                                                          \
    }
        CLODE_NVRTC_FUNCS(X)
#undef X
        return true;
    }
};

// nvJitLink: PTX -> sm_100a cubin without a GPU (the same ptxas NVRTC embeds).  Used when the runtime
// post-processes the PTX of a program (constant-divisor rewrite, clode_rt.cpp).  The library exports its entry
// points under versioned names (__nvJitLinkCreate_12_0 ... _12_9); the _12_0 names exist in every 12.x release.
struct JitLinkApi {
    typedef struct nvJitLink *Handle;
    int (*Create)(Handle *, unsigned int, const char **) = nullptr;
    int (*Destroy)(Handle *) = nullptr;
    int (*AddData)(Handle, int /* nvJitLinkInputType */, const void *, size_t, const char *) = nullptr;
    int (*Complete)(Handle) = nullptr;
    int (*GetLinkedCubinSize)(Handle, size_t *) = nullptr;
    int (*GetLinkedCubin)(Handle, void *) = nullptr;
    int (*GetErrorLogSize)(Handle, size_t *) = nullptr;
    int (*GetErrorLog)(Handle, char *) = nullptr;
    enum { INPUT_PTX = 2 }; // NVJITLINK_INPUT_PTX (nvJitLink.h: NONE, CUBIN, PTX, ...)
    void *handle = nullptr;
    std::string error;

    bool load()
    {
        const char *names[] = {"libnvJitLink.so.12", "/usr/local/cuda/lib64/libnvJitLink.so.12", "libnvJitLink.so",
                               "/usr/local/cuda/lib64/libnvJitLink.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (handle) break;
        }
        if (!handle) {
            error = std::string("cannot load nvJitLink (libnvJitLink.so.12): ") + dlerror();
            return false;
        }
#define X(member, sym)                                                         \
    member = reinterpret_cast<decltype(member)>(dlsym(handle, sym));           \
    if (!member) {                                                             \
        error = std::string("nvJitLink lacks symbol ") + sym;                  \
        return false;                                                          \
    }
        X(Create, "__nvJitLinkCreate_12_0") X(Destroy, "__nvJitLinkDestroy_12_0") X(AddData, "__nvJitLinkAddData_12_0")
        X(Complete, "__nvJitLinkComplete_12_0") X(GetLinkedCubinSize, "__nvJitLinkGetLinkedCubinSize_12_0")
        X(GetLinkedCubin, "__nvJitLinkGetLinkedCubin_12_0") X(GetErrorLogSize, "__nvJitLinkGetErrorLogSize_12_0")
        X(GetErrorLog, "__nvJitLinkGetErrorLog_12_0")
#undef X
        return true;
    }
};

// lazily-initialised singletons; nullptr + message on failure
inline DriverApi *driver(std::string *why = nullptr)
{
    static DriverApi api;
    static bool ok = false;
    static std::once_flag once;
    std::call_once(once, [] { ok = api.load(); });
    if (!ok && why) *why = api.error;
    return ok ? &api : nullptr;
}

inline NvrtcApi *nvrtc(std::string *why = nullptr)
{
    static NvrtcApi api;
    static bool ok = false;
    static std::once_flag once;
    std::call_once(once, [] { ok = api.load(); });
    if (!ok && why) *why = api.error;
    return ok ? &api : nullptr;
}

inline JitLinkApi *jitlink(std::string *why = nullptr)
{
    static JitLinkApi api;
    static bool ok = false;
    static std::once_flag once;
    std::call_once(once, [] { ok = api.load(); });
    if (!ok && why) *why = api.error;
    return ok ? &api : nullptr;
}

} // namespace clode
