// TEST INFRASTRUCTURE.  Runs the product's kernel source on the CPU: the device headers are
// compiled unmodified with g++ behind cuda_emu.h and each "kernel" is called once per
// (block, thread).  Build flags mirror the runtime's NVRTC options (csrc/rt/clode_rt.cpp
// compile_options).  Used by tests/test_device_emu.py to compare the device code with the
// oracle without a GPU.
#include "cuda_emu.h"

#include "cl_compat.cuh"
#include "rng.cuh"
#include "steppers.cuh"
#include "observers.cuh"
#include "kernels.cuh"
#define global
#define local
#define constant const
#include EMU_RHS_FILE

extern "C" {
KernelArgs clode_args; // stands in for the device's __constant__ argument block
}

static void run(void (*kernel)(), const KernelArgs *a)
{
    clode_args = *a;
    const unsigned block = CLODE_BLOCK;
    const size_t slots = a->n_slots ? (size_t)a->n_slots : (size_t)a->n;
    const unsigned grid = (unsigned)((slots + block - 1) / block);
    blockDim.x = block;
    gridDim.x = grid;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) {
            blockIdx.x = b;
            threadIdx.x = t;
            kernel();
        }
}

extern "C" {
int emu_args_size() { return (int)sizeof(KernelArgs); }
void emu_transient(const KernelArgs *a) { run(clode_transient, a); }
#ifdef CLODE_WITH_FEATURES
void emu_initialize_observer(const KernelArgs *a) { run(clode_initialize_observer, a); }
void emu_features(const KernelArgs *a) { run(clode_features, a); }
void emu_observer_layout(int *out) { clode_observer_layout(out); }
#endif
#ifdef CLODE_WITH_TRAJECTORY
void emu_trajectory(const KernelArgs *a) { run(clode_trajectory, a); }
#endif
}
