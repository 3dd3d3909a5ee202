// fast_exp.cuh — exp() for production double builds (not the bit-exact tier, not single precision).
//
// libdevice's exp is a degree-11 polynomial after a ln2 range reduction: 15 FP64-pipe instructions on one
// dependent chain plus ~24 moves that materialise its coefficients — and a gating-variable model calls it three to
// eight times per right-hand side (samples/lactotroph.cl: 3, examples/chay_keizer.cl: 8), i.e. up to 32 times per RK4
// step.  This one reduces by ln2/128 instead, looks 2^(j/128) up in a 2 KiB table and needs a degree-5 polynomial:
// 11 FP64-pipe instructions, chain depth 9.  The table is staged in SHARED memory at kernel entry: read through
// L1 (__ldg) it shares the cache with the local-memory spills of the fat-observer kernels, which evict it — C3
// (lactotroph, thresh2, 600 B of spills per thread) was 2 % slower with the L1 table than with libdevice's exp.
//     x = (128 m + j) ln2/128 + r,  |r| <= ln2/256:   exp(x) = 2^m * T[j] * (1 + p(r)),  p(r) = r + r^2/2 + ... + r^5/120
// The table entry is carried as hi + lo, so the only full rounding is the last addition: error <= 0.52 ulp measured
// over 2e7 arguments against 80-bit expl (tests/test_fast_exp.py) — CUDA documents 1 ulp for its own exp, OpenCL C
// allows 3.  Overflow, subnormal results (rounded twice, <= 1 ulp of the subnormal grid), Inf and NaN take an inline
// branch, as in libdevice; there is no call.
#ifndef CLODE_FAST_EXP_CUH
#define CLODE_FAST_EXP_CUH

// T[j] = 2^(j/128) as an unevaluated sum hi + lo (mpmath, 300 bits): one 16-byte load
__device__ const double2 clode_exp2_table[128] = {
    {0x1.0000000000000p+0, 0x0.0p+0}, {0x1.0163da9fb3335p+0, 0x1.b61299ab8cdb7p-54},
    {0x1.02c9a3e778061p+0, -0x1.19083535b085dp-56}, {0x1.04315e86e7f85p+0, -0x1.0a31c1977c96ep-54},
    {0x1.059b0d3158574p+0, 0x1.d73e2a475b465p-55}, {0x1.0706b29ddf6dep+0, -0x1.c91dfe2b13c27p-55},
    {0x1.0874518759bc8p+0, 0x1.186be4bb284ffp-57}, {0x1.09e3ecac6f383p+0, 0x1.1487818316136p-54},
    {0x1.0b5586cf9890fp+0, 0x1.8a62e4adc610bp-54}, {0x1.0cc922b7247f7p+0, 0x1.01edc16e24f71p-54},
    {0x1.0e3ec32d3d1a2p+0, 0x1.03a1727c57b53p-59}, {0x1.0fb66affed31bp+0, -0x1.b9bedc44ebd7bp-57},
    {0x1.11301d0125b51p+0, -0x1.6c51039449b3ap-54}, {0x1.12abdc06c31ccp+0, -0x1.1b514b36ca5c7p-58},
    {0x1.1429aaea92de0p+0, -0x1.32fbf9af1369ep-54}, {0x1.15a98c8a58e51p+0, 0x1.2406ab9eeab0ap-55},
    {0x1.172b83c7d517bp+0, -0x1.19041b9d78a76p-55}, {0x1.18af9388c8deap+0, -0x1.11023d1970f6cp-54},
    {0x1.1a35beb6fcb75p+0, 0x1.e5b4c7b4968e4p-55}, {0x1.1bbe084045cd4p+0, -0x1.95386352ef607p-54},
    {0x1.1d4873168b9aap+0, 0x1.e016e00a2643cp-54}, {0x1.1ed5022fcd91dp+0, -0x1.1df98027bb78cp-54},
    {0x1.2063b88628cd6p+0, 0x1.dc775814a8495p-55}, {0x1.21f49917ddc96p+0, 0x1.2a97e9494a5eep-55},
    {0x1.2387a6e756238p+0, 0x1.9b07eb6c70573p-54}, {0x1.251ce4fb2a63fp+0, 0x1.ac155bef4f4a4p-55},
    {0x1.26b4565e27cddp+0, 0x1.2bd339940e9d9p-55}, {0x1.284dfe1f56381p+0, -0x1.a4c3a8c3f0d7ep-54},
    {0x1.29e9df51fdee1p+0, 0x1.612e8afad1255p-55}, {0x1.2b87fd0dad990p+0, -0x1.10adcd6381aa4p-59},
    {0x1.2d285a6e4030bp+0, 0x1.0024754db41d5p-54}, {0x1.2ecafa93e2f56p+0, 0x1.1ca0f45d52383p-56},
    {0x1.306fe0a31b715p+0, 0x1.6f46ad23182e4p-55}, {0x1.32170fc4cd831p+0, 0x1.a9ce78e18047cp-55},
    {0x1.33c08b26416ffp+0, 0x1.32721843659a6p-54}, {0x1.356c55f929ff1p+0, -0x1.b5cee5c4e4628p-55},
    {0x1.371a7373aa9cbp+0, -0x1.63aeabf42eae2p-54}, {0x1.38cae6d05d866p+0, -0x1.e958d3c9904bdp-54},
    {0x1.3a7db34e59ff7p+0, -0x1.5e436d661f5e3p-56}, {0x1.3c32dc313a8e5p+0, -0x1.efff8375d29c3p-54},
    {0x1.3dea64c123422p+0, 0x1.ada0911f09ebcp-55}, {0x1.3fa4504ac801cp+0, -0x1.7d023f956f9f3p-54},
    {0x1.4160a21f72e2ap+0, -0x1.ef3691c309278p-58}, {0x1.431f5d950a897p+0, -0x1.1c7dde35f7999p-55},
    {0x1.44e086061892dp+0, 0x1.89b7a04ef80d0p-59}, {0x1.46a41ed1d0057p+0, 0x1.c944bd1648a76p-54},
    {0x1.486a2b5c13cd0p+0, 0x1.3c1a3b69062f0p-56}, {0x1.4a32af0d7d3dep+0, 0x1.9cb62f3d1be56p-54},
    {0x1.4bfdad5362a27p+0, 0x1.d4397afec42e2p-56}, {0x1.4dcb299fddd0dp+0, 0x1.8ecdbbc6a7833p-54},
    {0x1.4f9b2769d2ca7p+0, -0x1.4b309d25957e3p-54}, {0x1.516daa2cf6642p+0, -0x1.f768569bd93efp-55},
    {0x1.5342b569d4f82p+0, -0x1.07abe1db13cadp-55}, {0x1.551a4ca5d920fp+0, -0x1.d689cefede59bp-55},
    {0x1.56f4736b527dap+0, 0x1.9bb2c011d93adp-54}, {0x1.58d12d497c7fdp+0, 0x1.295e15b9a1de8p-55},
    {0x1.5ab07dd485429p+0, 0x1.6324c054647adp-54}, {0x1.5c9268a5946b7p+0, 0x1.c4b1b816986a2p-60},
    {0x1.5e76f15ad2148p+0, 0x1.ba6f93080e65ep-54}, {0x1.605e1b976dc09p+0, -0x1.3e2429b56de47p-54},
    {0x1.6247eb03a5585p+0, -0x1.383c17e40b497p-54}, {0x1.6434634ccc320p+0, -0x1.c483c759d8933p-55},
    {0x1.6623882552225p+0, -0x1.bb60987591c34p-54}, {0x1.68155d44ca973p+0, 0x1.038ae44f73e65p-57},
    {0x1.6a09e667f3bcdp+0, -0x1.bdd3413b26456p-54}, {0x1.6c012750bdabfp+0, -0x1.2895667ff0b0dp-56},
    {0x1.6dfb23c651a2fp+0, -0x1.bbe3a683c88abp-57}, {0x1.6ff7df9519484p+0, -0x1.83c0f25860ef6p-55},
    {0x1.71f75e8ec5f74p+0, -0x1.16e4786887a99p-55}, {0x1.73f9a48a58174p+0, -0x1.0a8d96c65d53cp-54},
    {0x1.75feb564267c9p+0, -0x1.0245957316dd3p-54}, {0x1.780694fde5d3fp+0, 0x1.866b80a02162dp-54},
    {0x1.7a11473eb0187p+0, -0x1.41577ee04992fp-55}, {0x1.7c1ed0130c132p+0, 0x1.f124cd1164dd6p-54},
    {0x1.7e2f336cf4e62p+0, 0x1.05d02ba15797ep-56}, {0x1.80427543e1a12p+0, -0x1.27c86626d972bp-54},
    {0x1.82589994cce13p+0, -0x1.d4c1dd41532d8p-54}, {0x1.8471a4623c7adp+0, -0x1.8d684a341cdfbp-55},
    {0x1.868d99b4492edp+0, -0x1.fc6f89bd4f6bap-54}, {0x1.88ac7d98a6699p+0, 0x1.994c2f37cb53ap-54},
    {0x1.8ace5422aa0dbp+0, 0x1.6e9f156864b27p-54}, {0x1.8cf3216b5448cp+0, -0x1.0d55e32e9e3aap-56},
    {0x1.8f1ae99157736p+0, 0x1.5cc13a2e3976cp-55}, {0x1.9145b0b91ffc6p+0, -0x1.dd6792e582524p-54},
    {0x1.93737b0cdc5e5p+0, -0x1.75fc781b57ebcp-57}, {0x1.95a44cbc8520fp+0, -0x1.64b7c96a5f039p-56},
    {0x1.97d829fde4e50p+0, -0x1.d185b7c1b85d1p-54}, {0x1.9a0f170ca07bap+0, -0x1.173bd91cee632p-54},
    {0x1.9c49182a3f090p+0, 0x1.c7c46b071f2bep-56}, {0x1.9e86319e32323p+0, 0x1.824ca78e64c6ep-56},
    {0x1.a0c667b5de565p+0, -0x1.359495d1cd533p-54}, {0x1.a309bec4a2d33p+0, 0x1.6305c7ddc36abp-54},
    {0x1.a5503b23e255dp+0, -0x1.d2f6edb8d41e1p-54}, {0x1.a799e1330b358p+0, 0x1.bcb7ecac563c7p-54},
    {0x1.a9e6b5579fdbfp+0, 0x1.0fac90ef7fd31p-54}, {0x1.ac36bbfd3f37ap+0, -0x1.f9234cae76cd0p-55},
    {0x1.ae89f995ad3adp+0, 0x1.7a1cd345dcc81p-54}, {0x1.b0e07298db666p+0, -0x1.bdef54c80e425p-54},
    {0x1.b33a2b84f15fbp+0, -0x1.2805e3084d708p-57}, {0x1.b59728de5593ap+0, -0x1.c71dfbbba6de3p-54},
    {0x1.b7f76f2fb5e47p+0, -0x1.5584f7e54ac3bp-56}, {0x1.ba5b030a1064ap+0, -0x1.efcd30e54292ep-54},
    {0x1.bcc1e904bc1d2p+0, 0x1.23dd07a2d9e84p-55}, {0x1.bf2c25bd71e09p+0, -0x1.efdca3f6b9c73p-54},
    {0x1.c199bdd85529cp+0, 0x1.11065895048ddp-55}, {0x1.c40ab5fffd07ap+0, 0x1.b4537e083c60ap-54},
    {0x1.c67f12e57d14bp+0, 0x1.2884dff483cadp-54}, {0x1.c8f6d9406e7b5p+0, 0x1.1acbc48805c44p-56},
    {0x1.cb720dcef9069p+0, 0x1.503cbd1e949dbp-56}, {0x1.cdf0b555dc3fap+0, -0x1.dd83b53829d72p-55},
    {0x1.d072d4a07897cp+0, -0x1.cbc3743797a9cp-54}, {0x1.d2f87080d89f2p+0, -0x1.d487b719d8578p-54},
    {0x1.d5818dcfba487p+0, 0x1.2ed02d75b3707p-55}, {0x1.d80e316c98398p+0, -0x1.11ec18beddfe8p-54},
    {0x1.da9e603db3285p+0, 0x1.c2300696db532p-54}, {0x1.dd321f301b460p+0, 0x1.2da5778f018c3p-54},
    {0x1.dfc97337b9b5fp+0, -0x1.1a5cd4f184b5cp-54}, {0x1.e264614f5a129p+0, -0x1.7b627817a1496p-54},
    {0x1.e502ee78b3ff6p+0, 0x1.39e8980a9cc8fp-55}, {0x1.e7a51fbc74c83p+0, 0x1.2d522ca0c8de2p-54},
    {0x1.ea4afa2a490dap+0, -0x1.e9c23179c2893p-54}, {0x1.ecf482d8e67f1p+0, -0x1.c93f3b411ad8cp-54},
    {0x1.efa1bee615a27p+0, 0x1.dc7f486a4b6b0p-54}, {0x1.f252b376bba97p+0, 0x1.3a1a5bf0d8e43p-54},
    {0x1.f50765b6e4540p+0, 0x1.9d3e12dd8a18bp-54}, {0x1.f7bfdad9cbe14p+0, -0x1.dbb12d006350ap-54},
    {0x1.fa7c1819e90d8p+0, 0x1.74853f3a5931ep-55}, {0x1.fd3c22b8f71f1p+0, 0x1.2eb74966579e7p-57},
};

#if !defined(CLODE_EXP_2K)
#ifndef CLODE_EXP_HOST_CHECK
__shared__ double2 clode_exp2_smem[128];
// every thread of the block, before any thread leaves the kernel
static __device__ __forceinline__ void clode_stage_exp_table()
{
    for (unsigned int j = threadIdx.x; j < 128u; j += blockDim.x)
        clode_exp2_smem[j] = clode_exp2_table[j];
    __syncthreads();
}
#define CLODE_EXP2_ENTRY(j) clode_exp2_smem[j]
#else
#define CLODE_EXP2_ENTRY(j) clode_exp2_table[j]
#endif

// Scalars in the constant bank: each is then a c[bank][offset] operand of the DFMA that uses it; as literals every
// call site re-materialises them with two UMOVs apiece inside the time loop (cf. the RK tableaux, steppers.cuh).
__constant__ double clode_exp_c[8] = {
    0x1.71547652b82fep+7,   // 128 / ln2
    0x1.8p52,               // 1.5 * 2^52: adding it leaves round(x * 128/ln2) in the low word
    -0x1.62e42fefa39efp-8,  // -(ln2 / 128), high part
    -0x1.abc9e3b39803fp-63, // -(ln2 / 128), low part
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};

static __device__ __forceinline__ double clode_fast_exp(const double x)
{
    const double *c = clode_exp_c;
    const double kf = fma(x, c[0], c[1]);
    const int k = __double2loint(kf);
    const double kd = kf - c[1];
    double r = fma(kd, c[2], x);
    r = fma(kd, c[3], r);
    const double r2 = r * r;
    double p = fma(r, fma(r, fma(r, c[4], c[5]), c[6]), c[7]);
    p = fma(r2, p, r);
    const double2 t = CLODE_EXP2_ENTRY(k & 127);
    const double y = t.x + fma(t.x, p, t.y); // in [1, 2.01)
    const int m = k >> 7;
    const int hi_abs = __double2hiint(x) & 0x7fffffff;
    if (hi_abs < 0x40862000) // |x| < 708: y * 2^m is a normal number, add m to the exponent field
        return __hiloint2double(__double2hiint(y) + (m << 20), __double2loint(y));
    // rare: overflow, subnormal results, Inf, NaN (the same cases libdevice branches on)
    if (hi_abs >= 0x40900000) // |x| >= 1024, Inf or NaN
        return x != x ? x + x : (x > 0.0 ? __hiloint2double(0x7ff00000, 0) : 0.0);
    // 708 <= |x| < 1024: two exact power-of-two factors; the second multiplication rounds once (to a subnormal,
    // zero or infinity where that is the answer)
    const int m1 = m >> 1;
    return y * __hiloint2double((1023 + m1) << 20, 0) * __hiloint2double((1023 + m - m1) << 20, 0);
}

#else // CLODE_EXP_2K ------------------------------------------------------------------------------------------------
// The BRANCH-FREE variant (round 2).  In the version above every call ends in a branch around the rare cases, and a
// branch is a scheduling barrier: the three sigmoids of a gating-variable right-hand side are evaluated one after the
// other, each a ~20-deep dependent FP64 chain, which is what the low-occupancy kernels (C3 / C4 / C5: 16 warps per SM)
// stall on (`wait` 40 %).  This one has no branch, so that — together with the branch-free reciprocal / division of
// the PTX pass (rt/ptx_pass.hpp) — a right-hand side is ONE basic block and the chains interleave:
//   * reduction by ln2/2048, 2^(j/2048) from a 16 KiB shared-memory table (one 8-byte word per entry), cubic
//     polynomial r + r^2 (a + r/6) with a = 1/2 + R^2/24 (the Chebyshev economisation of the r^4/24 term on
//     |r| <= R = ln2/4096): 8 FP64-pipe instructions instead of 11, error <= 1.1 ulp (table entry 0.5, final FMA 0.5,
//     truncation 0.04; CUDA documents 1 ulp for its own exp, OpenCL C allows 3);
//   * the scaling by 2^m as two factors ALWAYS, m = m1 + m2 with m1 = m >> 1: 2^m1 goes into the exponent field of the
//     table entry (a normal number for every m: NaN-safe, unlike an exponent add on the result), 2^m2 is the one
//     extra multiplication — it rounds only where the result is subnormal, zero or infinite, i.e. exactly the IEEE
//     answer for every finite argument without a range test;
//   * |x| >= 1024 and +-Inf (the low word of x * 2048/ln2 + 1.5 2^52 no longer holds an integer) are turned into
//     p = 0, m = -+2044 by selects, which the same two factors scale to 0 / +Inf; a NaN flows through r, p and the
//     final FMA untouched.
// The table is built at kernel entry from the 128-entry hi/lo table above times 2^(i/2048), i < 16, in
// double-double arithmetic (the rounded product differs from the correctly rounded entry in no case of the 2048:
// tests/emu/fast_exp_check.cpp compares every entry with 80-bit arithmetic).
__device__ const double2 clode_exp2_fine[16] = {
    {0x1.0000000000000p+0, 0x0.0p+0}, {0x1.00162f3904052p+0, -0x1.7b5d0d58ea8f4p-58},
    {0x1.002c605e2e8cfp+0, -0x1.d7c96f201bb2fp-55}, {0x1.0042936faa3d8p+0, -0x1.0484245243777p-55},
    {0x1.0058c86da1c0ap+0, -0x1.5e00e62d6b30dp-56}, {0x1.006eff583fc3dp+0, -0x1.4acf197a00142p-54},
    {0x1.0085382faef83p+0, 0x1.da93f90835f75p-56}, {0x1.009b72f41a12bp+0, 0x1.86364f8fbe8f8p-54},
    {0x1.00b1afa5abcbfp+0, -0x1.4f6b2a7609f71p-55}, {0x1.00c7ee448ee02p+0, 0x1.4362ca5bc26f1p-56},
    {0x1.00de2ed0ee0f5p+0, -0x1.406ac4e81a645p-57}, {0x1.00f4714af41d3p+0, -0x1.91b2060859321p-54},
    {0x1.010ab5b2cbd11p+0, 0x1.c1d0660524e08p-54}, {0x1.0120fc089ff63p+0, 0x1.843aa8b9cbbc6p-55},
    {0x1.0137444c9b5b5p+0, -0x1.2b6aeb6176892p-56}, {0x1.014d8e7ee8d2fp+0, 0x1.2edc08e5da99ap-56},
};

// 2^(j/2048) = 2^((j >> 4)/128) * 2^((j & 15)/2048), hi/lo product rounded once
static __device__ __forceinline__ double clode_exp2k_entry(const unsigned int j)
{
    const double2 a = clode_exp2_table[j >> 4], b = clode_exp2_fine[j & 15u];
    const double p = a.x * b.x;
    const double e = fma(a.x, b.x, -p);
    return p + fma(a.x, b.y, fma(a.y, b.x, e));
}

#ifndef CLODE_EXP_HOST_CHECK
__shared__ double clode_exp2k_smem[2048];
static __device__ __forceinline__ void clode_stage_exp_table()
{
    for (unsigned int j = threadIdx.x; j < 2048u; j += blockDim.x)
        clode_exp2k_smem[j] = clode_exp2k_entry(j);
    __syncthreads();
}
#define CLODE_EXP2K_ENTRY(j) clode_exp2k_smem[j]
#else
#define CLODE_EXP2K_ENTRY(j) clode_exp2k_entry(j)
#endif

__constant__ double clode_exp_c[6] = {
    0x1.71547652b82fep+11,  // 2048 / ln2
    0x1.8p52,               // 1.5 * 2^52: adding it leaves round(x * 2048/ln2) in the low word
    -0x1.62e42fefa39efp-12, // -(ln2 / 2048), high part
    -0x1.abc9e3b39803fp-67, // -(ln2 / 2048), low part
    0x1.5555555555555p-3,   // 1/6
    0x1.0000000a3fea0p-1};  // 1/2 + (ln2/4096)^2 / 24

static __device__ __forceinline__ double clode_fast_exp(const double x)
{
    const double *c = clode_exp_c;
    const double kf = fma(x, c[0], c[1]);
    const int k = __double2loint(kf);
    const double kd = kf - c[1];
    double r = fma(kd, c[2], x);
    r = fma(kd, c[3], r);
    double p = fma(r * r, fma(r, c[4], c[5]), r);
    const int hi = __double2hiint(x);
    // 1024 <= |x| <= Inf, not NaN (a NaN has hi_abs > 0x7ff00000, or == with a non-zero low word: such a NaN cannot
    // come out of an arithmetic instruction — the hardware quiets every NaN it produces)
    const bool huge = (unsigned int)((hi & 0x7fffffff) - 0x40900000) <= (unsigned int)(0x7ff00000 - 0x40900000);
    int m = k >> 11;
    if (huge) { // selects
        p = 0.0;
        m = hi < 0 ? -2044 : 2044;
    }
    const int m1 = m >> 1;
    const double t = CLODE_EXP2K_ENTRY(k & 2047);
    const double ts = __hiloint2double(__double2hiint(t) + (m1 << 20), __double2loint(t)); // 2^(j/2048) 2^m1: normal
    return fma(ts, p, ts) * __hiloint2double((1023 + m - m1) << 20, 0);
}
#endif // CLODE_EXP_2K
#endif // CLODE_FAST_EXP_CUH
