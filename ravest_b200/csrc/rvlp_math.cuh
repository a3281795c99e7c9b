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

#include "../../include/ravest_b200.h"

#if defined(__CUDACC__)
#define RV_HD __host__ __device__ __forceinline__
#define RV_SLOW __host__ __device__ __noinline__
#else
#define RV_HD inline
#define RV_SLOW inline
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

// ---------------------------------------------------------------- reciprocals
RV_HD double rcp64(double x) {
#if defined(__CUDA_ARCH__)
  // MUFU.RCP64H seed (~2^-20) + two Newton steps; x is never 0/inf/denormal at the call sites
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
// x = E - pi/2 in [-pi/2, pi/2]:  sin E = cos x,  cos E = -sin x.  Degree 9 / 8 near-minimax
// fits (abs error ~1e-7): the fp32 stage only has to deliver a starter good to ~1e-5.
RV_HD void sincos_0pi_f32(float E, float& s, float& c) {
  const float x = E - 1.57079632679f;
  const float z = x * x;
  float ps = ffmaf(2.630042900e-06f, z, -1.982125978e-04f);
  ps = ffmaf(ps, z, 8.333231322e-03f);
  ps = ffmaf(ps, z, -1.666666567e-01f);
  float pc = ffmaf(2.342478365e-05f, z, -1.386700082e-03f);
  pc = ffmaf(pc, z, 4.166554660e-02f);
  pc = ffmaf(pc, z, -4.999999106e-01f);
  const float sx = ffmaf(x * z, ps, x);
  const float cx = ffmaf(z, pc, 1.0f);
  s = cx;
  c = -sx;
}

// ---------------------------------------------------------------- fp64 sin/cos kernels
// |r| <= pi/4 (+ slack): the classic degree-13 / degree-14 minimax kernels (fdlibm constants).
RV_HD void sincos_kernel(double r, double& s, double& c) {
  const double z = r * r;
  double ps = ffma(1.58969099521155010221e-10, z, -2.50507602534068634195e-08);
  ps = ffma(ps, z, 2.75573137070700676789e-06);
  ps = ffma(ps, z, -1.98412698298579493134e-04);
  ps = ffma(ps, z, 8.33333333332248946124e-03);
  ps = ffma(ps, z, -1.66666666666666324348e-01);
  double pc = ffma(-1.13596475577881948265e-11, z, 2.08757232129817482790e-09);
  pc = ffma(pc, z, -2.75573143513906633035e-07);
  pc = ffma(pc, z, 2.48015872894767294178e-05);
  pc = ffma(pc, z, -1.38888888888741095749e-03);
  pc = ffma(pc, z, 4.16666666666666019037e-02);
  s = ffma(r * z, ps, r);
  c = ffma(z, ffma(z, pc, -0.5), 1.0);
}

// sin/cos of x in [0, pi] (a hair outside is fine); xf is x rounded to fp32 and only picks
// the quadrant, so any xf within ~1e-3 of x works.
RV_HD void sincos_0pi(double x, float xf, double& s, double& c) {
  const bool q1 = xf > 0.78539816f;
  const bool q2 = xf > 2.35619449f;
  const double kh = q2 ? 2.0 * PIO2_H : (q1 ? PIO2_H : 0.0);
  const double kl = q2 ? 2.0 * PIO2_L : (q1 ? PIO2_L : 0.0);
  const double r = (x - kh) - kl;
  double sr, cr;
  sincos_kernel(r, sr, cr);
  s = q2 ? -sr : (q1 ? cr : sr);
  c = q2 ? -cr : (q1 ? -sr : cr);
}

// Reduce a mean anomaly to m in [0, pi] and a sign:  M = 2 pi k + sg * m.
// Three-term Cody-Waite; exact to ~1e-20 |k| for |M| < BIG_M.  Larger |M| (far outside any
// real data set; the reference itself carries ulp(M) ~ 1e-9 rad of noise there) goes through
// libm's exact reduction instead.
RV_SLOW double reduce_big(double M) {
  return atan2(sin(M), cos(M));   // NaN / inf propagate as NaN
}

RV_HD void reduce_anomaly(double M, double& m, bool& neg) {
  double r;
  if (fabs(M) < BIG_M) {
    const double k = ffma(M, INV_2PI, RINT_MAGIC) - RINT_MAGIC;
    r = ffma(-k, TWO_PI_1, M);
    r = ffma(-k, TWO_PI_2, r);
    r = ffma(-k, TWO_PI_3, r);
  } else {
    r = reduce_big(M);
  }
  neg = r < 0.0;
  m = fabs(r);
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
//  4. `dlast` = |last step|: the caller checks it against the planet's tolerance and falls
//     back to kepler_robust for the (rare) lanes that did not contract enough.
struct SolverPlan {
  int n32;       // fp32 Halley iterations
  int n64;       // fp64 Householder steps
  double tol;    // acceptance bound on the last fp64 step
};

RV_HD SolverPlan plan_for(double e) {
  SolverPlan p;
  // Thresholds from tests/host/solver_check.cpp sweeps (worst last step per plan):
  //   (1,1): 4.5e-5 at e = 0.80 (2.1e-4 at 0.85, inaccurate from 0.9)
  //   (2,1): 1.2e-5 at e = 0.97 (8.4e-5 at 0.98, inaccurate from 0.99)
  //   (2,2) holds to 0.99, (3,3) to 0.999.
  if (e <= 0.80) { p.n32 = 1; p.n64 = 1; }
  else if (e <= 0.97) { p.n32 = 2; p.n64 = 1; }
  else if (e <= 0.99) { p.n32 = 2; p.n64 = 2; }
  else { p.n32 = 3; p.n64 = 3; }
  // a last step of size d leaves an error ~ K4 d^4; 2.5e-4 measured safe for every plan
  p.tol = (e <= 0.97) ? 2.5e-4 : 1.0e-5;
  if (!(e <= 0.999)) p.tol = -1.0;   // always take the robust path (also e = NaN)
  return p;
}

template <int W>
RV_HD void kepler_fast(const double (&M)[W], double e, int n32, int n64, double (&cosE)[W],
                       double (&sinE)[W], double (&dlast)[W]) {
  const float ef = (float)e;
  const float one_p_e2 = ffmaf(ef, ef, 1.0f);
  const float m2e = -2.0f * ef;
  double m[W];
  bool neg[W];
  float Ef[W], mf[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    reduce_anomaly(M[i], m[i], neg[i]);
    mf[i] = (float)m[i];
    float s, c;
    sincos_0pi_f32(mf[i], s, c);
    const float q = fmaxf(ffmaf(m2e, c, one_p_e2), 1e-12f);
    Ef[i] = ffmaf(ef * s, rsqrt32(q), mf[i]);
    Ef[i] = fminf(Ef[i], 3.14159274f);
  }
  for (int it = 0; it < n32; ++it) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      float s, c;
      sincos_0pi_f32(Ef[i], s, c);
      const float es = ef * s;
      const float f = (Ef[i] - mf[i]) - es;
      const float fp = ffmaf(-ef, c, 1.0f);
      const float den = ffmaf(fp, fp, -0.5f * f * es);
      const float d = f * fp * rcp32(den);
      Ef[i] = fminf(fmaxf(Ef[i] - d, 0.0f), 3.14159274f);
    }
  }
  double E[W], s[W], c[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    // NaN anomalies: keep them NaN (the reference returns NaN), comparisons above drop them
    E[i] = (m[i] == m[i]) ? (double)Ef[i] : m[i];
    sincos_0pi(E[i], Ef[i], s[i], c[i]);
    dlast[i] = 0.0;
  }
  for (int it = 0; it < n64; ++it) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const double es = e * s[i];
      const double ec = e * c[i];
      const double f = (E[i] - m[i]) - es;
      const double a = 1.0 - ec;
      const double a2 = a * a;
      const double t = ffma(-0.5 * f, es, a2);
      const double u = ffma(-f, es, a2);
      const double v = (f * f) * (ec * (1.0 / 6.0));
      const double den = ffma(a, u, v);
      const double d = -(f * t) * rcp64(den);
      // rotate (s, c) by d:  sin d = d - d^3/6,  cos d - 1 = -d^2/2 + d^4/24
      const double d2 = d * d;
      const double sd = ffma(d * d2, -1.0 / 6.0, d);
      const double cd1 = d2 * ffma(d2, 1.0 / 24.0, -0.5);
      const double sn = ffma(s[i], cd1, ffma(c[i], sd, s[i]));
      const double cn = ffma(c[i], cd1, ffma(-s[i], sd, c[i]));
      s[i] = sn;
      c[i] = cn;
      E[i] += d;
      dlast[i] = fabs(d);
    }
  }
#pragma unroll
  for (int i = 0; i < W; ++i) {
    cosE[i] = c[i];
    sinE[i] = neg[i] ? -s[i] : s[i];
  }
}

// Robust scalar fallback: f is increasing and convex on [0, pi], so Newton started from the
// right of the root (f >= 0) descends monotonically onto it; bisection bounds guard the
// round-off end game.  Full-precision libm sin/cos, convergence-tested, capped.
struct CosSin { double c, s; };

RV_SLOW CosSin kepler_robust(double M, double e) {
  CosSin out;
  double m;
  bool neg;
  reduce_anomaly(M, m, neg);
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
  out.s = neg ? -s : s;
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

RV_HD double cos_full(double x) {
  // cos of an unreduced angle via the same reduction + kernels (circular branch, model.py:242)
  double m;
  bool neg;
  reduce_anomaly(x, m, neg);
  double s, c;
  sincos_0pi(m, (float)m, s, c);
  return c;
}

template <int W>
RV_HD void planet_rv(const PlanetConst& pc, const SolverPlan& plan, const double (&t)[W],
                     double (&rv)[W]) {
  double M[W];
#pragma unroll
  for (int i = 0; i < W; ++i) M[i] = mean_anomaly(pc.n, t[i], pc.tp);
  if (pc.e == 0) {                          // model.py:239-242
#pragma unroll
    for (int i = 0; i < W; ++i) {
#if defined(__CUDA_ARCH__)
      const double x = __dadd_rn(M[i], pc.w);
#else
      volatile double x0 = M[i] + pc.w;
      const double x = x0;
#endif
      rv[i] = pc.K * cos_full(x);            // e * cos(w) == 0 exactly
    }
    return;
  }
  double cE[W], sE[W], dl[W];
  kepler_fast<W>(M, pc.e, plan.n32, plan.n64, cE, sE, dl);
#pragma unroll
  for (int i = 0; i < W; ++i) {
    if (!(dl[i] <= plan.tol) && (M[i] == M[i])) {
      const CosSin cs = kepler_robust(M[i], pc.e);
      cE[i] = cs.c;
      sE[i] = cs.s;
    }
    // model.py:119-121, 170 with K, cos w, sin w folded into A, B, C
    const double r = rcp64(ffma(-pc.e, cE[i], 1.0));
    const double u = ffma(-sE[i], pc.B, ffma(cE[i], pc.A, -pc.e * pc.A));
    rv[i] = ffma(r, u, pc.C);
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
