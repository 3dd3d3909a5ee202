// Host check of clode_b200/csrc/device/fast_polar.cuh: the device function compiled as host C++ (same text, same IEEE
// fma, a MODEL of the SFU reciprocal-square-root seed that is less accurate than the hardware's) against 80-bit
// arithmetic.  Driven by tests/test_fast_exp.py.   Build: g++ -O2 -march=x86-64-v3 -ffp-contract=off
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#define __device__
#define __forceinline__ inline
#define __constant__ static const
#define CLODE_POLAR_HOST_CHECK
static inline int __double2hiint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)b; }
static inline double __hiloint2double(int hi, int lo)
{
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
using std::fma;
static std::mt19937_64 g_noise(3);
// MUFU.RSQ64H: works on the high word, returns a high word; modelled with a relative error of up to 2^-19
static double clode_rsqrt_seed(double u)
{
    uint64_t b;
    std::memcpy(&b, &u, 8);
    b &= 0xffffffff00000000ull;
    double uh;
    std::memcpy(&uh, &b, 8);
    double y = 1.0 / std::sqrt(uh);
    y *= 1.0 + ((double)(int64_t)(g_noise() % 2001) - 1000.0) / 1000.0 * 0x1p-19;
    std::memcpy(&b, &y, 8);
    b &= 0xffffffff00000000ull;
    std::memcpy(&y, &b, 8);
    return y;
}

#include "fast_polar.cuh"

static double ulps(double got, double q)
{
    const long double ref = sqrtl(-2.0L * logl((long double)q) / (long double)q);
    int e;
    frexpl(ref, &e);
    return (double)(fabsl((long double)got - ref) / ldexpl(1.0L, e - 53));
}

int main(int argc, char **argv)
{
    const long count = argc > 1 ? std::atol(argv[1]) : 1000000;
    std::mt19937_64 g(17);
    double worst = 0;
    // q = a^2 + b^2 of the polar method is uniform-ish on (0, 1); plus the two ends, where a naive log loses digits
    const double ranges[][2] = {{0x1p-53, 1.0}, {0.99, 1.0}, {1.0 - 0x1p-20, 1.0}, {0x1p-60, 0x1p-20}, {0.7, 0.8}, {0.49, 0.51}, {0x1p-110, 0x1p-100}};
    for (auto &rg : ranges) {
        std::uniform_real_distribution<double> d(rg[0], rg[1]);
        double w = 0;
        for (long i = 0; i < count; ++i) {
            double q = d(g);
            if (!(q > 0.0 && q < 1.0)) continue;
            const double err = ulps(clode_polar_scale(q), q);
            if (err > w) w = err;
        }
        std::printf("q in [%g, %g): max error %.3f ulp\n", rg[0], rg[1], w);
        if (w > worst) worst = w;
    }
    // the largest q below 1 and its neighbours, exact powers of two, interval edges of the table
    int bad = 0;
    for (int k = 1; k < 2000; ++k) {
        const double near1 = 1.0 - k * 0x1p-53, edge = (1.0 + (k % 256) / 256.0) * std::ldexp(1.0, -1 - k % 60);
        for (double q : {near1, edge, std::nextafter(edge, 0.0), std::ldexp(1.0, -k % 1000 - 1)})
            if (q > 0.0 && q < 1.0 && ulps(clode_polar_scale(q), q) > 3.0) { std::printf("q=%a: %.3f ulp\n", q, ulps(clode_polar_scale(q), q)); ++bad; }
    }
    std::printf("worst=%.3f bad=%d\n", worst, bad);
    return (worst <= 3.0 && bad == 0) ? 0 : 1;
}
