// rvlp_math.cuh — device math for the batched RV log-probability path (sm_100a).
//
// Everything here is __host__ __device__ so that the Kepler solver's convergence can be
// verified exhaustively on the CPU (tests/host/solver_check.cpp, a `not gpu` test) as well
// as on the GPU.  On the device the fp32 reciprocals / rsqrt use the MUFU approximations.
//
// What is computed is the reference's arithmetic (paths relative to
// /root/reference/src/ravest/): model.py:23-243 (Kepler solve -> true anomaly -> RV),
// param.py:198-234, 17-105 (conversions, validity), prior.py:49-508 (log-priors).
// HOW it is computed is new: see DESIGN.md "Kepler solver".
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/ravest_b200.h"

#if defined(__CUDACC__)
#define RV_HD __host__ __device__ __forceinline__
#define RV_SLOW __host__ __device__ __noinline__
#else
#define RV_HD inline
#define RV_SLOW inline
#endif

// Micro-optimisation switches of the Kepler stage (bit mask; tools/kernel_sweep.py times them one by one):
//   1  two-term Cody-Waite reduction (the third term is 1.3e-21 |M|, five orders below the rounding of M itself)
//   2  reciprocals with fewer fp64 operations (one Newton step where 2^-40 suffices, e + e^2 correction elsewhere)
//   4  fp32 <-> fp64 conversions of non-negative values by integer bit operations (ALU pipe) instead of F2F (XU
//      pipe, 8 cycles per warp);  8  the signed one as well (five ALU instructions: costs more issue slots than it frees)
//  16  one rejection test per group of W anomalies (integer max of the last steps / of |M|) instead of one per anomaly
//  32  samples whose planets all take the lite plan (0 < e <= 0.65) run a loop without the per-planet plan dispatch
#ifndef RVLP_OPT
#define RVLP_OPT 51
#endif

namespace rvlp {

// 2*pi split into 33 + 33 + 53 bits: k * TWO_PI_1 and k * TWO_PI_2 are exact for |k| < 2^20.
constexpr double TWO_PI_1 = 6.2831853069365025;
constexpr double TWO_PI_2 = 2.4308402025215864e-10;
constexpr double TWO_PI_3 = 8.089064995183803e-21;
constexpr double INV_2PI = 0.15915494309189535;
constexpr double PIO2_H = 1.5707963267341256;      // 33-bit head of pi/2
constexpr double PIO2_L = 6.077100506506192e-11;
constexpr double PI_D = 3.141592653589793;
constexpr double RINT_MAGIC = 6755399441055744.0;  // 1.5 * 2^52
constexpr double BIG_M = 6.0e6;                    // |M| beyond this takes the exact-reduction slow path

RV_HD double ffma(double a, double b, double c) { return ::fma(a, b, c); }
RV_HD float ffmaf(float a, float b, float c) { return ::fmaf(a, b, c); }

// ---------------------------------------------------------------- bit helpers
RV_HD int hi32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  int64_t b;
  memcpy(&b, &x, 8);
  return (int)(b >> 32);
#endif
}
RV_HD double xor_hi(double x, int mask) {   // flip bits of the high word (sign manipulation)
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(x) ^ mask, __double2loint(x));
#else
  int64_t b;
  memcpy(&b, &x, 8);
  b ^= ((int64_t)(uint32_t)mask) << 32;
  memcpy(&x, &b, 8);
  return x;
#endif
}
RV_HD double from_hi(int hi) {              // double with the given high word and a zero low word
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, 0);
#else
  int64_t b = ((int64_t)(uint32_t)hi) << 32;
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}
#if defined(__CUDA_ARCH__)
#define RV_ANY(pred) __any_sync(__activemask(), (pred))
#else
#define RV_ANY(pred) (pred)
#endif

// Polynomial coefficients live in constant memory on the device so that DFMA reads them as
// c[bank][offset] operands (as immediates they cost two UMOVs per use: ~10% of all issue slots).
#define RV_SINCOS_COEFS                                                                         \
  { 1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,        \
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,       \
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,       \
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,        \
    /* 12: -pi/2 head, tail */ -1.5707963267341256, -6.077100506506192e-11,                     \
    /* 14: 1/2pi, rint magic, -2pi split */ 0.15915494309189535, 6755399441055744.0,            \
    -6.2831853069365025, -2.4308402025215864e-10, -8.089064995183803e-21,                       \
    /* 19: 1/6, 1/24, -1/6 */ 1.0 / 6.0, 1.0 / 24.0, -1.0 / 6.0 }
#if defined(__CUDACC__)
__constant__ double kCoefDev[22] = RV_SINCOS_COEFS;
#endif
static const double kCoefHost[22] = RV_SINCOS_COEFS;
#if defined(__CUDA_ARCH__)
#define RVK(i) kCoefDev[i]
#else
#define RVK(i) kCoefHost[i]
#endif

// sin / cos at the grid points j / 512, j = 0 .. kTabN-1 (covers [0, pi]): the fp64 stage looks up
// the grid point nearest to E0 and rotates by the (exactly representable) remainder |E0 - j/512| <= 2^-10,
// which needs only a cubic / quartic series - 10 fp64 operations instead of the 19 of the polynomial
// kernels, and no quadrant logic.  Filled once per device from host libm (rvlp_capi.cu) / on first use (host).
constexpr int kTabN = 1610;
#ifndef RVLP_SINCOS_TABLE
#define RVLP_SINCOS_TABLE 1
#endif
#if defined(__CUDACC__)
__device__ double2 kSinCosTabDev[2048];   // kTabN entries used; padded so that an 11-bit index can never leave it
#endif
struct SinCosPair { double s, c; };
inline const SinCosPair* sincos_table_host() {
  static SinCosPair tab[kTabN];
  static bool init = false;
  if (!init) {
    for (int j = 0; j < kTabN; ++j) { tab[j].s = sin(j / 512.0); tab[j].c = cos(j / 512.0); }
    init = true;
  }
  return tab;
}

// ---------------------------------------------------------------- log of a positive normal number
// One log per sample and lane closes the chi^2 / log-det reduction (the variances' mantissa product); libm's log is
// ~90 instructions, 13 % of all instructions at 120 epochs x 2 planets.  x = 2^e2 f, f in [1, 2): 64-entry table of
// (1 / c_j rounded, -log of that rounded value) for c_j = 1 + (j + 1/2) / 64, r = f / c_j - 1 by ONE fma (|r| < 2^-7),
// degree-7 log1p.  Absolute error < 5e-16 + 2.5e-16 |result| (tests/host/log_check.cpp); the result is a term of a sum of
// O(N) magnitude that is held to 1e-7.
#define RV_LOG_TABLE {                                                                       \
    {0x1.fc07f01fc07f0p-1, 0x1.fe02a6b106799p-8}, {0x1.f44659e4a4271p-1, 0x1.7b91b07d5b126p-6}, \
    {0x1.ecc07b301ecc0p-1, 0x1.39e87b9febd68p-5}, {0x1.e573ac901e574p-1, 0x1.b42dd711971b9p-5}, \
    {0x1.de5d6e3f8868ap-1, 0x1.16536eea37ae3p-4}, {0x1.d77b654b82c34p-1, 0x1.51b073f06183cp-4}, \
    {0x1.d0cb58f6ec074p-1, 0x1.8c345d6319b23p-4}, {0x1.ca4b3055ee191p-1, 0x1.c5e548f5bc743p-4}, \
    {0x1.c3f8f01c3f8f0p-1, 0x1.fec9131dbeabcp-4}, {0x1.bdd2b899406f7p-1, 0x1.1b72ad52f67a2p-3}, \
    {0x1.b7d6c3dda338bp-1, 0x1.371fc201e8f75p-3}, {0x1.b2036406c80d9p-1, 0x1.526e5e3a1b438p-3}, \
    {0x1.ac5701ac5701bp-1, 0x1.6d60fe719d21bp-3}, {0x1.a6d01a6d01a6dp-1, 0x1.87fa06520c911p-3}, \
    {0x1.a16d3f97a4b02p-1, 0x1.a23bc1fe2b561p-3}, {0x1.9c2d14ee4a102p-1, 0x1.bc286742d8cd4p-3}, \
    {0x1.970e4f80cb872p-1, 0x1.d5c216b4fbb94p-3}, {0x1.920fb49d0e229p-1, 0x1.ef0adcbdc5935p-3}, \
    {0x1.8d3018d3018d3p-1, 0x1.0402594b4d041p-2}, {0x1.886e5f0abb04ap-1, 0x1.1058bf9ae4ad4p-2}, \
    {0x1.83c977ab2beddp-1, 0x1.1c898c16999fbp-2}, {0x1.7f405fd017f40p-1, 0x1.2895a13de86a4p-2}, \
    {0x1.7ad2208e0ecc3p-1, 0x1.347dd9a987d56p-2}, {0x1.767dce434a9b1p-1, 0x1.404308686a7e4p-2}, \
    {0x1.724287f46debcp-1, 0x1.4be5f957778a1p-2}, {0x1.6e1f76b4337c7p-1, 0x1.5767717455a6cp-2}, \
    {0x1.6a13cd1537290p-1, 0x1.62c82f2b9c796p-2}, {0x1.661ec6a5122f9p-1, 0x1.6e08eaa2ba1e4p-2}, \
    {0x1.623fa77016240p-1, 0x1.792a55fdd47a1p-2}, {0x1.5e75bb8d015e7p-1, 0x1.842d1da1e8b18p-2}, \
    {0x1.5ac056b015ac0p-1, 0x1.8f11e873662c8p-2}, {0x1.571ed3c506b3ap-1, 0x1.99d958117e08ap-2}, \
    {0x1.5390948f40febp-1, 0x1.a484090e5bb09p-2}, {0x1.5015015015015p-1, 0x1.af1293247786bp-2}, \
    {0x1.4cab88725af6ep-1, 0x1.b9858969310fdp-2}, {0x1.49539e3b2d067p-1, 0x1.c3dd7a7cdad4dp-2}, \
    {0x1.460cbc7f5cf9ap-1, 0x1.ce1af0b85f3ecp-2}, {0x1.42d6625d51f87p-1, 0x1.d83e7258a2f3ep-2}, \
    {0x1.3fb013fb013fbp-1, 0x1.e24881a7c6c26p-2}, {0x1.3c995a47babe7p-1, 0x1.ec399d2468cc1p-2}, \
    {0x1.3991c2c187f63p-1, 0x1.f6123fa7028adp-2}, {0x1.3698df3de0748p-1, 0x1.ffd2e0857f497p-2}, \
    {0x1.33ae45b57bcb2p-1, 0x1.04bdf9da926d2p-1}, {0x1.30d190130d190p-1, 0x1.0986f4f573521p-1}, \
    {0x1.2e025c04b8097p-1, 0x1.0e44985d1cc8cp-1}, {0x1.2b404ad012b40p-1, 0x1.12f719593efbdp-1}, \
    {0x1.288b01288b013p-1, 0x1.179eabbd899a0p-1}, {0x1.25e22708092f1p-1, 0x1.1c3b81f713c25p-1}, \
    {0x1.23456789abcdfp-1, 0x1.20cdcd192ab6ep-1}, {0x1.20b470c67c0d9p-1, 0x1.2555bce98f7cap-1}, \
    {0x1.1e2ef3b3fb874p-1, 0x1.29d37fec2b08bp-1}, {0x1.1bb4a4046ed29p-1, 0x1.2e47436e40268p-1}, \
    {0x1.19453808ca29cp-1, 0x1.32b1339121d71p-1}, {0x1.16e0689427379p-1, 0x1.37117b54747b6p-1}, \
    {0x1.1485f0e0acd3bp-1, 0x1.3b68449fffc23p-1}, {0x1.12358e75d3033p-1, 0x1.3fb5b84d16f43p-1}, \
    {0x1.0fef010fef011p-1, 0x1.43f9fe2f9ce67p-1}, {0x1.0db20a88f4696p-1, 0x1.48353d1ea88dfp-1}, \
    {0x1.0b7e6ec259dc8p-1, 0x1.4c679afccee39p-1}, {0x1.0953f39010954p-1, 0x1.50913cc01686bp-1}, \
    {0x1.073260a47f7c6p-1, 0x1.54b2467999498p-1}, {0x1.05197f7d73404p-1, 0x1.58cadb5cd7989p-1}, \
    {0x1.03091b51f5e1ap-1, 0x1.5cdb1dc6c1765p-1}, {0x1.0101010101010p-1, 0x1.60e32f44788d9p-1}, \
  }
struct LogTabEntry { double inv, lg; };
#if defined(__CUDACC__)
__device__ const LogTabEntry kLogTabDev[64] = RV_LOG_TABLE;
__constant__ double kLogCoefDev[7] = {1.0 / 7.0, -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5, 0.6931471805599453};
#endif
static const LogTabEntry kLogTabHost[64] = RV_LOG_TABLE;
static const double kLogCoefHost[7] = {1.0 / 7.0, -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5, 0.6931471805599453};
RV_HD double log_pos_normal(double x) {
  const int hi = hi32(x);
  const int e2 = (hi >> 20) - 1023;
  const int j = (hi >> 14) & 63;
#if defined(__CUDA_ARCH__)
  const double f = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
  const double2 t = __ldg(reinterpret_cast<const double2*>(kLogTabDev) + j);
  const double inv = t.x, lg = t.y;
  const double* K = kLogCoefDev;
#else
  int64_t b;
  memcpy(&b, &x, 8);
  b = (b & 0x000fffffffffffffll) | 0x3ff0000000000000ll;
  double f;
  memcpy(&f, &b, 8);
  const double inv = kLogTabHost[j].inv, lg = kLogTabHost[j].lg;
  const double* K = kLogCoefHost;
#endif
  const double r = ffma(f, inv, -1.0);
  double w = ffma(r, K[0], K[1]);
  w = ffma(r, w, K[2]);
  w = ffma(r, w, K[3]);
  w = ffma(r, w, K[4]);
  w = ffma(r, w, K[5]);
  const double p = ffma(r * r, w, r);                       // log1p(r)
  return ffma((double)e2, K[6], lg + p);
}

// ---------------------------------------------------------------- reciprocals
RV_HD double rcp64(double x) {
#if defined(__CUDA_ARCH__)
  // MUFU.RCP64H seed (~2^-20) + two Newton steps; x is a normal positive number at the call sites
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = ffma(-x, y, 1.0);
  y = ffma(y, e, y);
  e = ffma(-x, y, 1.0);
  y = ffma(y, e, y);
  return y;
#else
  return 1.0 / x;
#endif
}

// 1/x to ~2^-40 (one Newton step): enough where the result only scales a quantity that is itself ~1e-6 (the Newton
// step of the lite plan) and the final reciprocal is re-derived against the true denominator afterwards.
#if !defined(__CUDA_ARCH__)
// host stand-in for MUFU.RCP64H (seed good to ~2^-20): worst-case relative error with a pseudo-random sign under
// RVLP_EMULATE_MUFU, exact otherwise; the Newton steps below are the device's.
inline double rcp64_seed_host(double x) {
  double y = 1.0 / x;
#if defined(RVLP_EMULATE_MUFU)
  uint64_t b;
  memcpy(&b, &x, 8);
  b *= 0x9E3779B97F4A7C15ull;
  y *= (b >> 63) ? (1.0 + 9.6e-7) : (1.0 - 9.6e-7);
#endif
  return y;
}
#endif
RV_HD double rcp64_40(double x) {
  double y;
#if defined(__CUDA_ARCH__)
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  y = rcp64_seed_host(x);
#endif
  const double e = ffma(-x, y, 1.0);
  return ffma(y, e, y);
}

// 1/x to round-off in three fp64 operations: seed (2^-20), e = 1 - x y, y (1 + e + e^2); e^3 ~ 1e-18.
RV_HD double rcp64_3(double x) {
  double y;
#if defined(__CUDA_ARCH__)
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  y = rcp64_seed_host(x);
#endif
  const double e = ffma(-x, y, 1.0);
  return ffma(y, ffma(e, e, e), y);
}

// fp32 <-> fp64 by integer bit operations.  f2d_pos: x >= 0 (x = 0 gives 2^-127, denormals are off by < 2^-126: both
// are absolute errors of < 1.2e-38 on an angle).  f2d_signed: any finite x.  d2f_trunc: 0 <= x < 2^127, truncating
// (result <= x, within one fp32 ulp); anything below 2^-126 becomes 0.
RV_HD double f2d_pos(float x) {
#if defined(__CUDA_ARCH__) && (RVLP_OPT & 4)
  const int b = __float_as_int(x);
  return __hiloint2double((int)((unsigned)b >> 3) + 0x38000000, b << 29);
#else
  return (double)x;
#endif
}
RV_HD double f2d_signed(float x) {
#if defined(__CUDA_ARCH__) && (RVLP_OPT & 8)
  const int b = __float_as_int(x);
  const int mag = b & 0x7fffffff;
  return __hiloint2double(((mag >> 3) + 0x38000000) | (b & (int)0x80000000), b << 29);
#else
  return (double)x;
#endif
}
RV_HD float d2f_trunc(double x) {
#if defined(__CUDA_ARCH__) && (RVLP_OPT & 4)
  const int hi = __double2hiint(x);
  const int f = __funnelshift_l(__double2loint(x), hi - 0x38000000, 3);
  return __int_as_float(hi < 0x38100000 ? 0 : f);
#elif (RVLP_OPT & 4)
  // host build: same truncation, so that tests/host/solver_check.cpp sweeps what the device computes
  float f = (float)x;
  if ((double)f > x) {
    uint32_t b;
    memcpy(&b, &f, 4);
    b -= 1;
    memcpy(&f, &b, 4);
  }
  return x < 1.1754943508222875e-38 ? 0.0f : f;
#else
  return (float)x;
#endif
}

RV_HD float rcp32(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.0f / x;
#endif
}

RV_HD float rsqrt32(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.0f / sqrtf(x);
#endif
}

// ---------------------------------------------------------------- fp32 sin/cos on [0, pi]
// The fp32 stage only has to deliver a starter good to ~1e-5: on the device the MUFU
// approximations (abs error ~4e-7 on [0, pi]) are enough and cost two issue slots each.
RV_HD void sincos_0pi_f32(float E, float& s, float& c) {
#if defined(__CUDA_ARCH__)
  s = __sinf(E);
  c = __cosf(E);
#else
  s = sinf(E);
  c = cosf(E);
#if defined(RVLP_EMULATE_MUFU)
  // host-only stress knob for tests/host/solver_check.cpp: worst-case MUFU.SIN/COS error (2^-20.9 abs)
  // with a pseudo-random sign, plus the fp32 range-reduction multiply's rounding
  uint32_t b;
  memcpy(&b, &E, 4);
  b = b * 2654435761u;
  s += (b & 0x10000u) ? 5.2e-7f : -5.2e-7f;
  c += (b & 0x20000u) ? 5.2e-7f : -5.2e-7f;
#endif
#endif
}

// ---------------------------------------------------------------- fp64 sin/cos kernels
// |r| <= pi/4 (+ slack): the classic degree-13 / degree-14 minimax kernels (fdlibm constants).
RV_HD void sincos_kernel(double r, double& s, double& c) {
  const double z = r * r;
  double ps = ffma(RVK(0), z, RVK(1));
  ps = ffma(ps, z, RVK(2));
  ps = ffma(ps, z, RVK(3));
  ps = ffma(ps, z, RVK(4));
  ps = ffma(ps, z, RVK(5));
  double pc = ffma(RVK(6), z, RVK(7));
  pc = ffma(pc, z, RVK(8));
  pc = ffma(pc, z, RVK(9));
  pc = ffma(pc, z, RVK(10));
  pc = ffma(pc, z, RVK(11));
  s = ffma(r * z, ps, r);
  c = ffma(z, ffma(z, pc, -0.5), 1.0);
}

// sin/cos of x in [0, pi] (a hair outside is fine); xf is x rounded to fp32 and only picks
// the quadrant q in {0, 1, 2}, so any xf within ~1e-3 of x works.  r = x - q pi/2 by two FMAs
// with q built from integer selects; the quadrant swap is two selects + two sign xors.
RV_HD void sincos_0pi(double x, float xf, double& s, double& c) {
  const bool q1 = xf > 0.78539816f;
  const bool q2 = xf > 2.35619449f;
  const double qd = from_hi(q2 ? 0x40000000 : (q1 ? 0x3ff00000 : 0));   // 2.0 / 1.0 / 0.0
  double r = ffma(qd, RVK(12), x);
  r = ffma(qd, RVK(13), r);
  double sr, cr;
  sincos_kernel(r, sr, cr);
  const bool swap = q1 && !q2;
  const double s0 = swap ? cr : sr;
  const double c0 = swap ? sr : cr;
  s = xor_hi(s0, q2 ? (int)0x80000000 : 0);
  c = xor_hi(c0, q1 ? (int)0x80000000 : 0);
}

// sin / cos of a float-exact E0 in [0, pi] by table + rotation (see kSinCosTabDev).
RV_HD void sincos_0pi_table(float E0f, double& s, double& c) {
  const float t = ffmaf(E0f, 512.0f, 12582912.0f);       // 1.5 * 2^23 + rint(512 E0)
  const float jf = t - 12582912.0f;
  const float ebf = ffmaf(jf, -0.001953125f, E0f);        // E0 - j / 512, exact
  const double eb = f2d_signed(ebf);
#if defined(__CUDA_ARCH__)
  // t = 1.5 * 2^23 + j: j sits in the low mantissa bits.  Masking (instead of subtracting the bias) gives an unsigned
  // 11-bit index: one LOP3 + one IMAD.WIDE.U32 for the address instead of five 64-bit integer instructions.
  const unsigned j = (unsigned)__float_as_int(t) & 0x7ffu;
  const double2 sc = __ldg(&kSinCosTabDev[j]);
  const double sa = sc.x, ca = sc.y;
#else
  const int j = (int)jf;
  const SinCosPair sc = sincos_table_host()[j];
  const double sa = sc.s, ca = sc.c;
#endif
  // |eb| <= 2^-10:  sin eb = eb - eb^3/6 (+ 7e-18),  cos eb - 1 = -eb^2/2 + eb^4/24 (- 1e-21)
  const double z = eb * eb;
  const double sb = ffma(eb * z, RVK(21), eb);
  const double cb1 = z * ffma(z, RVK(20), -0.5);
  s = ffma(sa, cb1, ffma(ca, sb, sa));
  c = ffma(ca, cb1, ffma(-sa, sb, ca));
}

// Reduce a mean anomaly to m in [0, pi] and a sign bit:  M = 2 pi k + sign * m.
// Three-term Cody-Waite, exact to ~1e-20 |k| for |M| < BIG_M; larger |M| is flagged by the
// caller (anomaly_is_big) and redone through libm's exact reduction (reduce_big).
RV_SLOW double reduce_big(double M) {
  return atan2(sin(M), cos(M));   // NaN / inf propagate as NaN
}
RV_HD bool anomaly_is_big(double M) {       // |M| >= BIG_M, inf or NaN, by an integer compare
  return (hi32(M) & 0x7fffffff) >= 0x4156e360;   // high word of 6.0e6
}
RV_HD void split_sign(double r, double& m, int& sign) {
  sign = hi32(r) & (int)0x80000000;
  m = fabs(r);               // an operand modifier of the consuming fp64 instruction, not an instruction
}
RV_HD void reduce_anomaly(double M, double& m, int& sign) {
  const double k = ffma(M, RVK(14), RVK(15)) - RVK(15);
  double r = ffma(k, RVK(16), M);
  r = ffma(k, RVK(17), r);
#if !(RVLP_OPT & 1)
  r = ffma(k, RVK(18), r);
#endif
  split_sign(r, m, sign);
}

// ---------------------------------------------------------------- Kepler solver
// Solves E - e sin E = M for W independent anomalies that share one eccentricity
// (one warp works on one (sample, planet), so e and the iteration counts are warp-uniform
// and there is no divergence).  Returns cos E, sin E (model.py:23-70's outputs).
//
//  1. reduce M to m in [0, pi] (symmetry E(-M) = -E(M));
//  2. fp32 starter on the otherwise idle FP32/MUFU pipes:
//     E0 = m + e sin m / sqrt(1 - 2 e cos m + e^2), then n32 fp32 Halley steps;
//  3. fp64: ONE full sin/cos of E0, then n64 fourth-order (Householder-3) steps, each of
//     which updates (sin E, cos E) by an angle-addition with a short series in the step
//     instead of calling sin/cos again;
//  4. `dlast` = last step: the caller checks |dlast| against the planet's tolerance (an
//     integer compare that also catches NaN) and sends the (rare) lanes that did not
//     contract enough, or whose |M| is huge, through kepler_robust.
struct SolverPlan {
  int n32;       // fp32 Halley iterations
  int n64;       // fp64 steps; 0 selects the "lite" plan (one Halley step, see kepler_fast)
  double tol;    // acceptance bound on the last fp64 step
};

RV_HD SolverPlan plan_for(double e) {
  SolverPlan p;
  // Thresholds from tests/host/solver_check.cpp sweeps (worst last step per plan):
  //   (1,1): 4.5e-5 at e = 0.80 (2.1e-4 at 0.85, inaccurate from 0.9)
  //   (2,1): 1.2e-5 at e = 0.97 (8.4e-5 at 0.98, inaccurate from 0.99)
  //   (2,2) holds to 0.99, (3,3) to 0.999.
  //   lite (one fp32 + one fp64 *Halley* step): the fp32 stage leaves <= 5.4e-7 at e = 0.6 and
  //   2.4e-6 at e = 0.7, so a third-order step is already at round-off for e <= 0.65.
  if (e <= 0.65) { p.n32 = 1; p.n64 = 0; }
  else if (e <= 0.80) { p.n32 = 1; p.n64 = 1; }
  else if (e <= 0.97) { p.n32 = 2; p.n64 = 1; }
  else if (e <= 0.99) { p.n32 = 2; p.n64 = 2; }
  else { p.n32 = 3; p.n64 = 3; }
  // a last step of size d leaves an error ~ K4 d^4; 2.5e-4 measured safe for every plan
  p.tol = (e <= 0.65) ? 4.0e-6 : ((e <= 0.97) ? 2.5e-4 : 1.0e-5);
  if (!(e <= 0.999)) p.tol = -1.0;   // always take the robust path (also e = NaN)
  return p;
}

// true when the last step is too large (or NaN): compares the high words as integers
RV_HD bool step_rejected(double dlast, double tol) {
  return (hi32(dlast) & 0x7fffffff) >= hi32(tol) || hi32(tol) < 0;
}

// The solver in two stages so that callers can software-pipeline them (stage A of the next
// planet is independent of stage B of the current one, and uses different pipes).
//
// Stage A (fp32 / MUFU pipes + 5 fp64 ops): reduce M, starter, N32 fp32 Halley steps.
template <int W>
struct StarterOut {
  double m[W];     // reduced anomaly in [0, pi]
  float Ef[W];     // fp32 estimate of E, clamped to [0, pi]
  int sign[W];     // sign bit of the reduced anomaly (E(-M) = -E(M))
};

template <int W, int N32 = -1>
RV_HD void kepler_stage_a(const double (&M)[W], double e, int n32, StarterOut<W>& o) {
  const float ef = (float)e;
  const float one_p_e2 = ffmaf(ef, ef, 1.0f);
  const float m2e = -2.0f * ef;
  float mf[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    reduce_anomaly(M[i], o.m[i], o.sign[i]);
    mf[i] = d2f_trunc(o.m[i]);
    float s, c;
    sincos_0pi_f32(mf[i], s, c);
    // q = |1 - e exp(i m)|^2 >= (1 - e)^2 > 0; a rounding-negative q gives NaN, which the clamp at the
    // end of the fp32 stage turns into E0 = 0 and the step test then rejects
    const float q = ffmaf(m2e, c, one_p_e2);
    o.Ef[i] = ffmaf(ef * s, rsqrt32(q), mf[i]);
  }
#pragma unroll
  for (int it = 0; it < (N32 >= 0 ? N32 : n32); ++it) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      float s, c;
      sincos_0pi_f32(o.Ef[i], s, c);
      const float es = ef * s;
      const float f = (o.Ef[i] - mf[i]) - es;
      const float fp = ffmaf(-ef, c, 1.0f);
      const float den = ffmaf(fp, fp, -0.5f * f * es);
      const float d = f * fp * rcp32(den);
      o.Ef[i] = o.Ef[i] - d;
    }
  }
#pragma unroll
  for (int i = 0; i < W; ++i)   // the fp64 kernels need E0 in [0, pi]; NaN -> 0
    o.Ef[i] = fminf(fmaxf(o.Ef[i], 0.0f), 3.14159274f);
}

// Stage B (fp64 pipe): one sincos of E0, then the fp64 step(s).
// N64 >= 0 fixes the step count at compile time (straight-line code for the plans that cover
// e <= 0.97); -1 reads it from n64.
// N64 == 0 is the "lite" plan for e <= 0.65: the fp32 stage already lands within ~1e-6, so ONE
// third-order (Halley) step written as a series in f/f' suffices, the rotation needs only second
// order, and 1/(1 - e cos E) follows from 1/f'(E0) by a two-term Newton update (no second MUFU seed).
// rinv = 1 / (1 - e cos E), the factor model.py:119-121 divides by.
template <int W, int N64 = -1>
RV_HD void kepler_stage_b(const StarterOut<W>& o, double e, int n64, double (&cosE)[W], double (&sinE)[W],
                          double (&dlast)[W], double (&rinv)[W]) {
  const double (&m)[W] = o.m;
  const int (&sign)[W] = o.sign;
  double E[W], s[W], c[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    E[i] = f2d_pos(o.Ef[i]);
#if RVLP_SINCOS_TABLE
    sincos_0pi_table(o.Ef[i], s[i], c[i]);
#else
    sincos_0pi(E[i], o.Ef[i], s[i], c[i]);
#endif
    dlast[i] = 0.0;
  }
  if (N64 == 0) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const double es = e * s[i];
      const double f = (E[i] - m[i]) - es;     // NaN m (NaN / inf anomaly) poisons d -> rejected
      const double a = ffma(-e, c[i], 1.0);
#if RVLP_OPT & 2
      const double ra = rcp64_40(a);           // 2^-40 on a step of ~1e-6; rinv below is re-derived against e cos E
#else
      const double ra = rcp64(a);
#endif
      const double x = f * ra;                 // Newton step, |x| <~ 1e-6
      const double y = es * ra;
#if RVLP_OPT & 2
      const double u = -0.5 * x;
      const double d = x * ffma(u, y, -1.0);           // Halley: -x (1 + x y / 2), next term ~1e-19
      const double h = u * x;                          // -d^2 / 2 to 2e-18: d = -x (1 + O(1e-6)), |h| ~ 1e-12
#else
      const double d = -x * ffma(0.5 * x, y, 1.0);     // Halley: -x / (1 - x y / 2), next term ~1e-19
      const double h = -0.5 * (d * d);
#endif
      const double sn = ffma(s[i], h, ffma(c[i], d, s[i]));
      const double cn = ffma(c[i], h, ffma(-s[i], d, c[i]));
      const double eps = ffma(ffma(e, cn, -1.0), ra, 1.0);   // 1 - (1 - e cos E) / a
      rinv[i] = ffma(ra, ffma(eps, eps, eps), ra);
      cosE[i] = cn;
      sinE[i] = xor_hi(sn, sign[i]);
      dlast[i] = d;
    }
  } else {
#pragma unroll
  for (int it = 0; it < (N64 >= 0 ? N64 : n64); ++it) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const double es = e * s[i];
      const double ec = e * c[i];
      const double f = (E[i] - m[i]) - es;     // NaN m (NaN / inf anomaly) poisons d -> rejected
      const double a = 1.0 - ec;
      const double a2 = a * a;
      const double t = ffma(-0.5 * f, es, a2);
      const double u = ffma(-f, es, a2);
      const double v = (f * f) * (ec * RVK(19));
      const double den = ffma(a, u, v);
#if RVLP_OPT & 2
      const double d = -(f * t) * rcp64_3(den);
#else
      const double d = -(f * t) * rcp64(den);
#endif
      // rotate (s, c) by d:  sin d = d - d^3/6,  cos d - 1 = -d^2/2 + d^4/24
      const double d2 = d * d;
      double sd, cd1;
      if (N64 >= 0) {   // single-step plans: |d| <= 2.5e-4, series error < 1e-20
        sd = ffma(d * d2, RVK(21), d);
        cd1 = d2 * ffma(d2, RVK(20), -0.5);
      } else {          // multi-step plans (e > 0.97): early steps reach 3e-3 and 1/(1-e) amplifies any drift
        sd = ffma(d * d2, ffma(d2, 1.0 / 120.0, RVK(21)), d);
        cd1 = d2 * ffma(d2, ffma(d2, -1.0 / 720.0, RVK(20)), -0.5);
      }
      const double sn = ffma(s[i], cd1, ffma(c[i], sd, s[i]));
      const double cn = ffma(c[i], cd1, ffma(-s[i], sd, c[i]));
      s[i] = sn;
      c[i] = cn;
      E[i] += d;
      // Multi-step plans: the rotation's series is only good to ~|d|^7/5040, and a later step cannot
      // repair a drifted (s, c); an early step above 3e-3 therefore poisons the verdict (-> robust path).
      if (it > 0 && (hi32(dlast[i]) & 0x7fffffff) >= 0x3f689374) dlast[i] = 1.0;
      else dlast[i] = (it > 0 && dlast[i] == 1.0) ? 1.0 : d;
    }
  }
#pragma unroll
  for (int i = 0; i < W; ++i) {
    cosE[i] = c[i];
    sinE[i] = xor_hi(s[i], sign[i]);
#if RVLP_OPT & 2
    rinv[i] = rcp64_3(ffma(-e, c[i], 1.0));
#else
    rinv[i] = rcp64(ffma(-e, c[i], 1.0));
#endif
  }
  }
}

template <int W, int N32 = -1, int N64 = -1>
RV_HD void kepler_fast(const double (&M)[W], double e, int n32, int n64, double (&cosE)[W],
                       double (&sinE)[W], double (&dlast)[W], double (&rinv)[W]) {
  StarterOut<W> o;
  kepler_stage_a<W, N32>(M, e, n32, o);
  kepler_stage_b<W, N64>(o, e, n64, cosE, sinE, dlast, rinv);
}

// Robust scalar fallback: f is increasing and convex on [0, pi], so Newton started from the
// right of the root (f >= 0) descends monotonically onto it; bisection bounds guard the
// round-off end game.  Convergence-tested, capped.
struct CosSin { double c, s; };

RV_SLOW CosSin kepler_robust(double M, double e) {
  CosSin out;
  double m;
  int sign;
  if (anomaly_is_big(M)) split_sign(reduce_big(M), m, sign);
  else reduce_anomaly(M, m, sign);
  if (!(m == m)) { out.c = m; out.s = m; return out; }
  double lo = m, hi = m + e;
  if (hi > PI_D) hi = PI_D;
  if (hi < lo) hi = lo;
  double E = hi;
  for (int it = 0; it < 200; ++it) {
    double s, c;
    sincos_0pi(E, (float)E, s, c);
    const double f = (E - m) - e * s;
    if (f > 0) hi = E; else lo = E;
    const double fp = 1.0 - e * c;
    double En = E - f / fp;
    if (!(En > lo && En < hi)) En = 0.5 * (lo + hi);
    const double step = fabs(En - E);
    E = En;
    if (step <= 4.0e-16 * (E > 1.0 ? E : 1.0) || hi - lo <= 0.0) break;
  }
  double s, c;
  sincos_0pi(E, (float)E, s, c);
  out.c = c;
  out.s = xor_hi(s, sign);
  return out;
}

// ---------------------------------------------------------------- per-planet constants
// What Planet.__init__ + _njit_kepler_rv's prologue derive once per sample
// (model.py:203-206, 302; param.py:299-362, 88-105).
struct PlanetConst {
  double n;      // 2 pi / P
  double tp;     // time of periastron
  double e;
  double A;      // K cos w
  double B;      // K sqrt(1-e^2) sin w
  double C;      // K e cos w
  double w;      // only used by the circular branch
  double K;
};

struct DefaultPars { double P, K, e, w, tp; bool conv_error; bool invalid; };

// param.py:299-362 + validity param.py:17-105.  `in` = the five values in pars order.
RV_HD DefaultPars to_default(int par, const double* in) {
  DefaultPars d;
  d.conv_error = false;
  d.P = in[0];
  d.K = in[1];
  if (par == RVLP_PAR_PKEWTP || par == RVLP_PAR_PKEWTC) {
    d.e = in[2];
    d.w = in[3];
  } else {
    // param.py:232 — secosw**2 + sesinw**2 with numpy's three roundings (no FMA contraction): the
    // validity tests compare e with 0 and 1 exactly
#if defined(__CUDA_ARCH__)
    d.e = __dadd_rn(__dmul_rn(in[2], in[2]), __dmul_rn(in[3], in[3]));
#else
    {
      volatile double uu = in[2] * in[2], vv = in[3] * in[3];
      d.e = uu + vv;
    }
#endif
    d.w = atan2(in[3], in[2]);             // param.py:233
  }
  if (par == RVLP_PAR_PKEWTC || par == RVLP_PAR_PKSECTC) {
    if (d.e < 0 || d.e >= 1.0) {           // param.py:209 raises before any arithmetic
      d.conv_error = true;
      d.tp = NAN;
    } else {                               // param.py:206-215
      const double th = (PI_D / 2) - d.w;
      const double E = 2 * atan(sqrt((1 - d.e) / (1 + d.e)) * tan(th / 2));
      const double Mc = E - (d.e * sin(E));
      d.tp = in[4] - (d.P / (2 * PI_D)) * Mc;
    }
  } else {
    d.tp = in[4];
  }
  // NaN passes every comparison exactly as in the reference (param.py:25-86)
  d.invalid = d.conv_error || (d.P <= 0) || (d.K <= 0) || (d.e < 0) || (d.e >= 1.0) ||
              !(-PI_D <= d.w && d.w < PI_D);
  return d;
}

RV_HD PlanetConst planet_const(const DefaultPars& d) {
  PlanetConst c;
  c.n = (2 * PI_D) / d.P;                  // model.py:302
  c.tp = d.tp;
  c.e = d.e;
  const double s1 = sqrt(1.0 - d.e * d.e);  // model.py:203
  const double cw = cos(d.w), sw = sin(d.w);
  c.A = d.K * cw;
  c.B = d.K * s1 * sw;
  c.C = d.K * (d.e * cw);
  c.w = d.w;
  c.K = d.K;
  return c;
}

// One planet's RV at W epochs (model.py:216-243 + 327). Mean anomaly keeps the reference's
// two roundings: M = n * (t - tp).
RV_HD double mean_anomaly(double n, double t, double tp) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(n, __dsub_rn(t, tp));
#else
  volatile double dt = t - tp;
  return n * dt;
#endif
}

RV_SLOW double cos_big(double x) { return cos(x); }

RV_HD double cos_full(double x) {
  // cos of an unreduced angle via the same reduction + kernels (circular branch, model.py:242)
  double m;
  int sign;
  reduce_anomaly(x, m, sign);
  double s, c;
  sincos_0pi(m, (float)m, s, c);
  return c;
}

// Adds one planet's RV at W epochs into rv[] (model.py:216-243 + 327; fit.py:3630's `+=`).
// The constant K e cos w term is NOT added here: the caller folds the planets' C terms into one
// per-sample constant.
template <int W>
RV_HD void planet_rv_add(const PlanetConst& pc, const SolverPlan& plan, const double (&t)[W],
                         double (&rv)[W]) {
  double M[W];
#pragma unroll
  for (int i = 0; i < W; ++i) M[i] = mean_anomaly(pc.n, t[i], pc.tp);
  if (pc.e == 0) {                          // model.py:239-242, e * cos(w) == 0 exactly
    bool big = false;
    double x[W], cx[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
#if defined(__CUDA_ARCH__)
      x[i] = __dadd_rn(M[i], pc.w);
#else
      volatile double x0 = M[i] + pc.w;
      x[i] = x0;
#endif
      big |= anomaly_is_big(x[i]);
      cx[i] = cos_full(x[i]);
    }
    if (RV_ANY(big)) {
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (anomaly_is_big(x[i])) cx[i] = cos_big(x[i]);
    }
#pragma unroll
    for (int i = 0; i < W; ++i) rv[i] = ffma(pc.K, cx[i], rv[i]);
    return;
  }
  double cE[W], sE[W], dl[W], ri[W];
  if (plan.n64 == 0) kepler_fast<W, 1, 0>(M, pc.e, 1, 0, cE, sE, dl, ri);
  else if (plan.n64 == 1 && plan.n32 == 1) kepler_fast<W, 1, 1>(M, pc.e, 1, 1, cE, sE, dl, ri);
  else if (plan.n64 == 1 && plan.n32 == 2) kepler_fast<W, 2, 1>(M, pc.e, 2, 1, cE, sE, dl, ri);
  else kepler_fast<W>(M, pc.e, plan.n32, plan.n64, cE, sE, dl, ri);
#if RVLP_OPT & 16
  // one test per group: the largest |last step| and the largest |M| of the W anomalies, compared as high words
  // (NaN / inf have the largest high words of all)
  int dmax = 0, mmax = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const int dh = hi32(dl[i]) & 0x7fffffff, mh = hi32(M[i]) & 0x7fffffff;
    dmax = dh > dmax ? dh : dmax;
    mmax = mh > mmax ? mh : mmax;
  }
  const bool bad = dmax >= hi32(plan.tol) || hi32(plan.tol) < 0 || mmax >= 0x4156e360;
#else
  bool bad = false;
#pragma unroll
  for (int i = 0; i < W; ++i) bad |= step_rejected(dl[i], plan.tol) || anomaly_is_big(M[i]);
#endif
  if (RV_ANY(bad)) {                        // warp-uniform branch; rare
#pragma unroll
    for (int i = 0; i < W; ++i) {
      if (step_rejected(dl[i], plan.tol) || anomaly_is_big(M[i])) {
        const CosSin cs = kepler_robust(M[i], pc.e);
        cE[i] = cs.c;
        sE[i] = cs.s;
        ri[i] = 1.0 / (1.0 - pc.e * cs.c);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < W; ++i) {
    // model.py:119-121, 170 with K, cos w, sin w folded into A, B, C (C == e A)
    const double u = ffma(-sE[i], pc.B, ffma(cE[i], pc.A, -pc.C));
    rv[i] = ffma(ri[i], u, rv[i]);
  }
}

// planet_rv_add for a planet known to take the lite plan (0 < e <= 0.65): same arithmetic, same bits, without the
// circular-orbit test and the plan dispatch.
template <int W>
RV_HD void planet_rv_add_lite(double n, double tp, double e, double A, double B, double C, const double (&t)[W],
                              double (&rv)[W]) {
  double M[W], cE[W], sE[W], dl[W], ri[W];
#pragma unroll
  for (int i = 0; i < W; ++i) M[i] = mean_anomaly(n, t[i], tp);
  kepler_fast<W, 1, 0>(M, e, 1, 0, cE, sE, dl, ri);
  int dmax = 0, mmax = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const int dh = hi32(dl[i]) & 0x7fffffff, mh = hi32(M[i]) & 0x7fffffff;
    dmax = dh > dmax ? dh : dmax;
    mmax = mh > mmax ? mh : mmax;
  }
  if (RV_ANY(dmax >= 0x3ed0c6f7 || mmax >= 0x4156e360)) {   // high words of the lite tolerance 4.0e-6 and of BIG_M
#pragma unroll
    for (int i = 0; i < W; ++i) {
      if (step_rejected(dl[i], 4.0e-6) || anomaly_is_big(M[i])) {
        const CosSin cs = kepler_robust(M[i], e);
        cE[i] = cs.c;
        sE[i] = cs.s;
        ri[i] = 1.0 / (1.0 - e * cs.c);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const double u = ffma(-sE[i], B, ffma(cE[i], A, -C));
    rv[i] = ffma(ri[i], u, rv[i]);
  }
}

// ---------------------------------------------------------------- priors (prior.py)
RV_HD double prior_logpdf(const rvlp_prior& pr, double x) {
  const double* p = pr.p;
  switch (pr.kind) {
    case RVLP_PRIOR_UNIFORM:
      return (x < p[0] || x > p[1]) ? -INFINITY : pr.c[0];
    case RVLP_PRIOR_ECC_UNIFORM:
      return (x < 0 || x >= p[0]) ? -INFINITY : pr.c[0];
    case RVLP_PRIOR_NORMAL: {
      const double z = (x - p[0]) / p[1];
      return -0.5 * (z * z) - pr.c[0];
    }
    case RVLP_PRIOR_TRUNC_NORMAL: {
      if (x < p[2] || x > p[3]) return -INFINITY;
      const double z = (x - p[0]) / p[1];
      return -(z * z) / 2.0 + pr.c[0];
    }
    case RVLP_PRIOR_HALF_NORMAL: {
      if (x < 0.0) return -INFINITY;
      const double y = x / p[0];
      return -(y * y) / 2.0 + pr.c[0];
    }
    case RVLP_PRIOR_RAYLEIGH: {
      if (x < 0.0 || x == INFINITY) return -INFINITY;   // scipy: -inf outside the open support
      const double r = x / p[0];
      return log(r) - 0.5 * r * r + pr.c[0];
    }
    case RVLP_PRIOR_VANEYLEN19: {
      if (x < 0.0 || x == INFINITY) return -INFINITY;
      const double y = x / p[0], r = x / p[1], f = p[2];
      double lh = -(y * y) / 2.0 + pr.c[0];
      double lr = log(r) - 0.5 * r * r + pr.c[1];
      // scipy.special.logsumexp(a, b=[1-f, f]): zero-weight terms are dropped before the max
      if (1.0 - f == 0.0) lh = -INFINITY;
      if (f == 0.0) lr = -INFINITY;
      const double mx = lh > lr ? lh : lr;
      if (mx == -INFINITY) return -INFINITY;
      double s = 0.0;
      if (lh != -INFINITY) s += (1.0 - f) * exp(lh - mx);
      if (lr != -INFINITY) s += f * exp(lr - mx);
      return log(s) + mx;
    }
    case RVLP_PRIOR_BETA: {
      if (x < 0.0 || x > 1.0) return -INFINITY;
      const double a1 = p[0] - 1.0, b1 = p[1] - 1.0;
      const double t1 = (a1 == 0.0 && x == x) ? 0.0 : a1 * log(x);
      const double t2 = (b1 == 0.0 && x == x) ? 0.0 : b1 * log1p(-x);
      return t1 + t2 - pr.c[0];
    }
  }
  return NAN;
}

}  // namespace rvlp
