// Host check of clode_b200/csrc/device/fast_exp.cuh: the device function compiled as host C++ (same text, same
// IEEE fma) against the 80-bit expl of the host.  Driven by tests/test_fast_exp.py.
// Build: g++ -O2 -march=x86-64-v3 -ffp-contract=off [-DCLODE_EXP_2K]   (the branch-free 2048-entry variant: <= 1.1 ulp)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#define __device__
#define __forceinline__ inline
#define __noinline__
#define CLODE_EXP_HOST_CHECK
#define __constant__ static const
struct double2 { double x, y; };
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int __double2hiint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)b; }
static inline double __hiloint2double(int hi, int lo)
{
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
using std::fma;

#include "fast_exp.cuh"

#ifdef CLODE_EXP_2K
static const double BOUND = 1.1;
#else
static const double BOUND = 0.53;
#endif

static double ulps(double y, double x)
{
    const long double ref = expl((long double)x);
    if (std::isinf((double)ref) || ref == 0.0L) return y == (double)ref ? 0.0 : 1e9;
    int e;
    frexpl(ref, &e);
    if (e < -1021) e = -1021; // subnormal results: spacing 2^-1074
    return (double)(fabsl((long double)y - ref) / ldexpl(1.0L, e - 53));
}

int main(int argc, char **argv)
{
    const long count = argc > 1 ? std::atol(argv[1]) : 1000000;
    std::mt19937_64 g(7);
    const double ranges[][2] = {{-1, 1}, {-10, 10}, {-100, 100}, {-707.9, 707.9}, {-0.01, 0.01}, {-40, 5}, {-750, 720}};
    double worst = 0;
    for (auto &rg : ranges) {
        std::uniform_real_distribution<double> d(rg[0], rg[1]);
        double w = 0;
        for (long i = 0; i < count; ++i) {
            const double x = d(g);
            const double err = ulps(clode_fast_exp(x), x);
            if (err > w) w = err;
        }
        std::printf("range [%g, %g]: max error %.4f ulp\n", rg[0], rg[1], w);
        // results that are subnormal are rounded twice (once to 53 bits, once to the subnormal grid): up to 1 ulp there
        const bool edge = rg[0] < -708.0;
        if (edge ? w > BOUND + 0.5 : w > BOUND) worst = 9.0;
        if (!edge && w > worst) worst = w;
    }
    // special values take the library path
    const double specials[] = {0.0, -0.0, 708.0, -708.0, 709.78, 710.0, -745.0, -746.0, INFINITY, -INFINITY, 0x1p-1074, 1e-300,
                               1023.9, -1023.9, 1024.0, -1024.0, 1e7, -1e7, 1.2e7, -1.2e7, 1e15, -1e15, 1e300, -1e300,
                               0x1.fffffffffffffp1023, -0x1.fffffffffffffp1023, 709.782712893384, -745.1332191019412};
    int bad = 0;
    for (double x : specials)
        if (ulps(clode_fast_exp(x), x) > BOUND + 0.5) { std::printf("special %a wrong: %a\n", x, clode_fast_exp(x)); ++bad; }
#ifdef CLODE_EXP_2K
    // every entry of the table built at kernel entry is the correctly rounded 2^(j/2048)
    for (unsigned int j = 0; j < 2048; ++j) {
        const long double ref = exp2l((long double)j / 2048.0L);
        if (clode_exp2k_entry(j) != (double)ref) { std::printf("table entry %u: %a, want %a\n", j, clode_exp2k_entry(j), (double)ref); ++bad; }
    }
#endif
    if (!std::isnan(clode_fast_exp(NAN))) { std::printf("NaN not propagated\n"); ++bad; }
    if (clode_fast_exp(0.0) != 1.0) { std::printf("exp(0) != 1\n"); ++bad; }
    std::printf("worst=%.4f bad=%d\n", worst, bad);
    return (worst < BOUND && bad == 0) ? 0 : 1;
}
