// rvlp_gpcov.cuh — the quasi-periodic covariance function of the GP kernels (K3, K7), branch-free.
//
//   k(tau) = A^2 exp(-Gamma sin^2(pi |tau| / P_gp) - tau^2 / (2 lambda_e^2)),   Gamma = 1 / (2 lambda_p^2)
//   (/root/reference/src/ravest/gp.py:145-156: ExpSquared(scale = lambda_e) * ExpSineSquared(scale = P_gp, gamma = Gamma),
//    scaled by A^2; tinygp evaluates exp(-Gamma sin^2(pi r)) with r = |tau| / P_gp.)
//
// libm's sinpi + exp cost ~90 instructions per element with data-dependent branches, so the 36 elements of
// a register tile ran one after the other (1000 cycles each, profiles/r01i_gp_phase_timing.md).  Here:
//   * sin(pi r): r reduced to [-1/2, 1/2] by a magic-number rint (exact), x = pi |r| in [0, pi/2], then the
//     1/512-radian sin/cos grid of the Kepler stage (kSinCosTabDev) + a cubic / quartic rotation;
//   * exp(y): y = (n / 64) ln 2 + r with |r| <= ln 2 / 128, a 64-entry 2^(j/64) table, a degree-5 polynomial
//     and an exact power-of-two scale (clamped: exp(y) below 2^-1022 is returned as ~2^-1022 * 2^(j/64) A^2,
//     i.e. 1e-308 instead of a denormal or 0 - far below one ulp of any diagonal entry).
// ~33 fp64 + ~12 integer / load instructions, no branches: the compiler interleaves the elements of a tile.
// Accuracy (tests/host/gpcov_check.cpp): relative error <= 3e-16 against long-double over the hyperparameter
// ranges of config 5 and far beyond.  Works on the host as well (the host check and the table are shared).
#pragma once
#include "rvlp_math.cuh"

namespace rvlp {

#define RV_EXP2_TABLE                                                                               \
  { 0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,         \
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,         \
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,         \
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,         \
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,         \
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,         \
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,         \
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,         \
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,         \
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,         \
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,         \
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,         \
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,         \
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,         \
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,         \
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0 }
#if defined(__CUDACC__)
__device__ const double kExp2TabDev[64] = RV_EXP2_TABLE;
// 0: 64/ln2   1: -ln2/64 head (32 bits)   2: -ln2/64 tail   3..6: 1/120, 1/24, 1/6, 1/2   7: pi   8: 512   9: -1/512
__constant__ double kGpCoefDev[10] = {0x1.71547652b82fep+6, -0x1.62e42fee00000p-7, -0x1.a39ef35793c76p-39,
                                      1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 3.141592653589793, 512.0, -0.001953125};
#endif
static const double kExp2TabHost[64] = RV_EXP2_TABLE;
static const double kGpCoefHost[10] = {0x1.71547652b82fep+6, -0x1.62e42fee00000p-7, -0x1.a39ef35793c76p-39,
                                       1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 3.141592653589793, 512.0, -0.001953125};
#if defined(__CUDA_ARCH__)
#define RVG(i) kGpCoefDev[i]
#else
#define RVG(i) kGpCoefHost[i]
#endif

RV_HD int lo32(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  int64_t b;
  memcpy(&b, &x, 8);
  return (int)(uint32_t)b;
#endif
}

// Per-sample constants of the covariance function (gp.py:145-156).
struct GpHyper {
  double inv_P, inv_le, gamma, A2;
};
RV_HD GpHyper gp_hyper(double A, double lambda_e, double lambda_p, double P_gp) {
  GpHyper h;
  h.gamma = 1.0 / (2.0 * (lambda_p * lambda_p));   // gp.py:152
  h.A2 = A * A;
  h.inv_le = 1.0 / lambda_e;
  h.inv_P = 1.0 / P_gp;
  return h;
}

// sin(pi r) for r = u - rint(u), returned as its square's ingredients: s = sin(pi |r|) in [0, 1].
RV_HD double gp_sinpi_frac(double u) {
  const double t = u + RINT_MAGIC;                 // rint(u) for |u| < 2^51
  const double kf = t - RINT_MAGIC;
  const double r = fabs(u - kf);                   // exact, in [0, 1/2]
  const double x = RVG(7) * r;                     // pi r in [0, pi/2]; one rounding (1.7e-16 absolute)
  const double jt = ffma(x, RVG(8), RINT_MAGIC);   // rint(512 x)
  const double jf = jt - RINT_MAGIC;
  const double eb = ffma(jf, RVG(9), x);           // x - j / 512, exact, |eb| <= 2^-10
#if defined(__CUDA_ARCH__)
  const unsigned j = (unsigned)lo32(jt) & 0x7ffu;  // rint(512 x) <= 805 in the low bits of the magic sum; u = inf / NaN:
                                                   // any 11-bit index stays inside the (padded) table, the result is NaN
#else
  int j = lo32(jt);
  j = j < 0 ? 0 : (j > kTabN - 1 ? kTabN - 1 : j);
#endif
#if defined(__CUDA_ARCH__)
  const double2 sc = __ldg(&kSinCosTabDev[j]);
  const double sa = sc.x, ca = sc.y;
#else
  const SinCosPair sc = sincos_table_host()[j];
  const double sa = sc.s, ca = sc.c;
#endif
  const double z = eb * eb;
  const double sb = ffma(eb * z, RVK(21), eb);     // sin eb = eb - eb^3/6 (+ 7e-18)
  const double cb1 = z * ffma(z, RVK(20), -0.5);   // cos eb - 1 = -eb^2/2 + eb^4/24
  return ffma(sa, cb1, ffma(ca, sb, sa));
}

// scale * exp(y), y <= 0 in normal use (any y gives a finite-or-NaN answer without branching).
RV_HD double gp_exp_scaled(double y, double scale) {
  const double t = ffma(y, RVG(0), RINT_MAGIC);    // n = rint(64 y / ln 2)
  const double nf = t - RINT_MAGIC;
  double r = ffma(nf, RVG(1), y);
  r = ffma(nf, RVG(2), r);                         // |r| <= ln2 / 128
  int n = lo32(t);
  // y = -inf / NaN / beyond +-709: keep the scale factor a normal number (see the header comment)
  const bool in_range = fabs(y) < 708.0;
  n = in_range ? n : (y < 0.0 ? -(1022 << 6) : 0);
  r = (in_range || y != y) ? r : 0.0;
#if defined(__CUDA_ARCH__)
  const double tj = __ldg(&kExp2TabDev[n & 63]);
#else
  const double tj = kExp2TabHost[n & 63];
#endif
  const double two_e = from_hi(((n >> 6) + 1023) << 20);
  double w = ffma(r, RVG(3), RVG(4));
  w = ffma(r, w, RVG(5));
  w = ffma(r, w, RVG(6));
  const double p = ffma(r * r, w, r);              // exp(r) - 1
  const double ts = tj * two_e;                    // exact
  return ffma(ts, p, ts) * scale;
}

// The same for y <= 0 (or NaN), the only values the covariance function produces: one clamp instead of the range logic
// (exp(y) below 2^-1021 is returned as ~3e-308 * scale; NaN propagates).
RV_HD double gp_exp_scaled_neg(double y, double scale) {
  y = y < -708.0 ? -708.0 : y;
  const double t = ffma(y, RVG(0), RINT_MAGIC);    // n = rint(64 y / ln 2)
  const double nf = t - RINT_MAGIC;
  double r = ffma(nf, RVG(1), y);
  r = ffma(nf, RVG(2), r);                         // |r| <= ln2 / 128
  const int n = lo32(t);
#if defined(__CUDA_ARCH__)
  const double tj = __ldg(&kExp2TabDev[n & 63]);
#else
  const double tj = kExp2TabHost[n & 63];
#endif
  const double two_e = from_hi(((n >> 6) + 1023) << 20);
  double w = ffma(r, RVG(3), RVG(4));
  w = ffma(r, w, RVG(5));
  w = ffma(r, w, RVG(6));
  const double p = ffma(r * r, w, r);              // exp(r) - 1
  const double ts = tj * two_e;                    // exact
  return ffma(ts, p, ts) * scale;
}

// k(tau) without the white-noise diagonal.
RV_HD double gp_cov(double tau, const GpHyper& h) {
  const double s = gp_sinpi_frac(fabs(tau) * h.inv_P);
  const double q = tau * h.inv_le;
  const double y = ffma(-h.gamma, s * s, (-0.5 * q) * q);
  return gp_exp_scaled(y, h.A2);
}

}  // namespace rvlp
