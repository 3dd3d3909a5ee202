// Micro-benchmarks of the FP64 pipe on one GPU: issue rate per instruction kind (all SMs, many warps, 8 independent
// accumulators per thread) and dependent-issue latency (one warp).  Used to interpret ncu's
// sm__inst_executed_pipe_fp64 percentage for the step loop.  Build: see run_probe.sh.
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__constant__ double cc[4] = {1.0000001, 1e-9, 0.9999999, 0.5};

enum { DFMA_R, DFMA_C, DMUL_R, DADD_R, DSETP_SEL, DFMA_FSEL, DFMA_DSETP, NMODES };
static const char *names[] = {"DFMA reg operands", "DFMA const operand", "DMUL", "DADD", "DSETP + SEL(int)",
                              "DFMA + 2 FSEL (1:2)", "DFMA:DSETP 4:1 (+2 FSEL), DFMAs counted"};

template <int MODE> __global__ void __launch_bounds__(128) rate(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + i * b + threadIdx.x * 1e-12;
    int cnt = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == DFMA_R) x[i] = fma(x[i], a, b);
                if (MODE == DFMA_C) x[i] = fma(x[i], cc[0], cc[1]);
                if (MODE == DMUL_R) x[i] = x[i] * a;
                if (MODE == DADD_R) x[i] = x[i] + b;
                if (MODE == DSETP_SEL)
                    asm volatile("{.reg .pred p; setp.gt.f64 p, %1, %2; selp.s32 %0, %3, %0, p;}" : "+r"(cnt) : "d"(x[i]), "d"(a), "r"(u + it));
                if (MODE == DFMA_FSEL) {
                    x[i] = fma(x[i], a, b);
                    asm volatile("{.reg .pred p; setp.ne.s32 p, %1, 12345; selp.f64 %0, %0, %2, p;}" : "+d"(x[i]) : "r"(it), "d"(b));
                }
                if (MODE == DFMA_DSETP) {
                    x[i] = fma(x[i], a, b);
                    if ((i & 3) == 3)
                        asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %1, %0, p;}" : "+d"(x[i]) : "d"(a));
                }
            }
        }
    }
    double s = cnt;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent chain: latency in cycles per instruction
template <int MODE> __global__ void latency(double *out, long long *cycles, int iters, double a, double b)
{
    double x = a;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 64; ++u) {
            if (MODE == 0) x = fma(x, a, b);
            if (MODE == 1) x = x * a;
            if (MODE == 2) x = x + b;
            if (MODE == 3) asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %1, %0, p;}" : "+d"(x) : "d"(b));
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE> int run_rate(double *d, int sms, int per_it, double clock_ghz)
{
    const int blocks = sms * 5, iters = 4000;
    rate<MODE><<<blocks, 128>>>(d, 100, 1.0000001, 1e-9);
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    CHECK(cudaEventRecord(e0));
    rate<MODE><<<blocks, 128>>>(d, iters, 1.0000001, 1e-9);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double warp_instr = (double)blocks * 4 * iters * 64 * per_it; // FP64-pipe warp instructions
    const double per_clk_sm = warp_instr / (ms * 1e-3 * clock_ghz * 1e9) / sms;
    printf("%-55s %8.3f ms  %6.3f FP64 warp-instr/clk/SM (nominal 2.0)\n", names[MODE], ms, per_clk_sm);
    return 0;
}

int main()
{
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    int khz;
    CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    double *d;
    long long *c;
    CHECK(cudaMalloc(&d, sizeof(double) * p.multiProcessorCount * 5 * 128));
    CHECK(cudaMalloc(&c, 8));
    const int sms = p.multiProcessorCount;
    if (run_rate<DFMA_R>(d, sms, 1, ghz)) return 1;
    if (run_rate<DFMA_C>(d, sms, 1, ghz)) return 1;
    if (run_rate<DMUL_R>(d, sms, 1, ghz)) return 1;
    if (run_rate<DADD_R>(d, sms, 1, ghz)) return 1;
    if (run_rate<DSETP_SEL>(d, sms, 1, ghz)) return 1;
    if (run_rate<DFMA_FSEL>(d, sms, 1, ghz)) return 1;
    if (run_rate<DFMA_DSETP>(d, sms, 1, ghz)) return 1; // counts the DFMAs only: 8 DFMA + 2 DSETP per 8
    const char *ln[] = {"DFMA", "DMUL", "DADD", "DSETP+2 FSEL"};
    for (int m = 0; m < 4; ++m) {
        long long h;
        if (m == 0) latency<0><<<1, 32>>>(d, c, 1000, 1.0000001, 1e-9);
        if (m == 1) latency<1><<<1, 32>>>(d, c, 1000, 1.0000001, 1e-9);
        if (m == 2) latency<2><<<1, 32>>>(d, c, 1000, 1.0000001, 1e-9);
        if (m == 3) latency<3><<<1, 32>>>(d, c, 1000, 1.0000001, 1e-9);
        CHECK(cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost));
        printf("dependent %-14s %6.2f cycles per link\n", ln[m], (double)h / (1000.0 * 64));
    }
    return 0;
}
