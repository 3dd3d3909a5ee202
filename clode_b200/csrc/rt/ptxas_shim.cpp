// ptxas_shim.cpp -> clode_b200/libclode_ptxas.so: PTX -> sm_100a cubin, no GPU needed.
//
// ptxas as a library (CUDA's libnvptxcompiler_static.a, linked statically into this one small shared object so
// that libclode_rt.so itself keeps no link-time CUDA dependency; the runtime dlopen()s it from its own directory).
// It compiles the PTX as a WHOLE PROGRAM, exactly like NVRTC's embedded ptxas does: __constant__ data (the RK
// tableaux) stay direct c[bank][offset] operands.  nvJitLink, the other PTX entry point of the toolkit, compiles
// relocatable code and re-loads every such constant with LDCU inside the time loop (+30 instructions per Lorenz
// dopri5 attempt), which is why it is only the fallback.
#include <nvPTXCompiler.h>

#include <cstdlib>
#include <cstring>

extern "C" {

// returns 0 and malloc'd *cubin (caller frees with clode_ptxas_free) or non-zero and a malloc'd *log (may be null)
__attribute__((visibility("default"))) int clode_ptxas(const char *ptx, size_t ptx_size, int lineinfo, void **cubin,
                                                        size_t *cubin_size, char **log)
{
    *cubin = nullptr;
    *cubin_size = 0;
    *log = nullptr;
    nvPTXCompilerHandle h = nullptr;
    if (nvPTXCompilerCreate(&h, ptx_size, ptx) != NVPTXCOMPILE_SUCCESS) return 1;
    const char *opts[] = {"--gpu-name=sm_100a", "-lineinfo"};
    const nvPTXCompileResult r = nvPTXCompilerCompile(h, lineinfo ? 2 : 1, opts);
    if (r != NVPTXCOMPILE_SUCCESS) {
        size_t n = 0;
        if (nvPTXCompilerGetErrorLogSize(h, &n) == NVPTXCOMPILE_SUCCESS && n > 0) {
            *log = (char *)std::calloc(n + 1, 1);
            if (*log) nvPTXCompilerGetErrorLog(h, *log);
        }
        nvPTXCompilerDestroy(&h);
        return 2;
    }
    size_t n = 0;
    if (nvPTXCompilerGetCompiledProgramSize(h, &n) != NVPTXCOMPILE_SUCCESS || n == 0) {
        nvPTXCompilerDestroy(&h);
        return 3;
    }
    *cubin = std::malloc(n);
    if (!*cubin || nvPTXCompilerGetCompiledProgram(h, *cubin) != NVPTXCOMPILE_SUCCESS) {
        std::free(*cubin);
        *cubin = nullptr;
        nvPTXCompilerDestroy(&h);
        return 4;
    }
    *cubin_size = n;
    nvPTXCompilerDestroy(&h);
    return 0;
}

__attribute__((visibility("default"))) void clode_ptxas_free(void *p) { std::free(p); }

__attribute__((visibility("default"))) int clode_ptxas_version(unsigned int *major, unsigned int *minor)
{
    return nvPTXCompilerGetVersion(major, minor) == NVPTXCOMPILE_SUCCESS ? 0 : 1;
}
}
