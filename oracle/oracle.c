/*
 * oracle.c — plain-C CPU restatement of ravest's batched RV log-probability path.
 *
 * TEST INFRASTRUCTURE ONLY.  Linked/loaded only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs; never by the product package.
 *
 * Parity: PINNED for the white-noise path — tests/test_oracle.py checks this file against
 * the golden vectors produced by the unmodified reference (tests/golden/*.json) and against
 * oracle_py.py (a dict-level restatement that is bit-identical to the reference on those
 * vectors).  GP path: "parity unpinned" (tinygp 0.3.0 / jax 0.8.3 absent; formula restated
 * from src/ravest/gp.py:145-156 and SURVEY.md Appendix A.5).
 *
 * It follows the REFERENCE's algorithm (Halley from E0 = M, tol 1.48e-8, libm sin/cos per
 * iteration), not the CUDA kernel's, so it doubles as the CPU baseline ("port") that
 * bench.py times on the host cores.  Citations: /root/reference/src/ravest/<file>:<line>.
 *
 * It consumes the same flat descriptor (include/ravest_b200.h) as the CUDA library so both
 * can be driven from one host-side compilation of a problem.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/ravest_b200.h"

#define ORACLE_PI 3.141592653589793

/* model.py:23-70 */
static void solve_kepler(double Mi, double e, double* cosE, double* sinE) {
  const double tol = 1.48e-08;
  double Ei = Mi, s = 0.0, c = 0.0;
  for (int it = 0; it < 50; ++it) {
    s = sin(Ei);
    c = cos(Ei);
    double f = Ei - e * s - Mi;
    double fp = 1.0 - e * c;
    double fpp = e * s;
    double En = Ei - f / (fp - (f * fpp) / (2.0 * fp));
    if (fabs(En - Ei) < tol) {
      s = sin(En);
      c = cos(En);
      break;
    }
    Ei = En;
  }
  *cosE = c;
  *sinE = s;
}

/* model.py:173-243 (_compute_rv incl. the e == 0 shortcut), accumulating into rv[] */
static void kepler_rv_add(const double* t, int64_t n, double nmot, double tp, double e, double K,
                          double w, double* rv, int accumulate) {
  if (e == 0) {
    double ecw = e * cos(w);
    for (int64_t i = 0; i < n; ++i) {
      double M = nmot * (t[i] - tp);                       /* model.py:327 */
      double v = K * (cos(M + w) + ecw);                   /* model.py:242 */
      rv[i] = accumulate ? rv[i] + v : v;
    }
    return;
  }
  double s1me2 = sqrt(1.0 - e * e), cw = cos(w), sw = sin(w), ecw = e * cw; /* model.py:203-206 */
  for (int64_t i = 0; i < n; ++i) {
    double M = nmot * (t[i] - tp);
    double cE, sE;
    solve_kepler(M, e, &cE, &sE);
    double d = 1.0 - e * cE;                               /* model.py:119-121 */
    double cf = (cE - e) / d;
    double sf = s1me2 * sE / d;
    double v = K * (cf * cw - sf * sw + ecw);              /* model.py:170 */
    rv[i] = accumulate ? rv[i] + v : v;
  }
}

int oracle_kepler_rv(const double* M, int64_t n, double e, double K, double w, double* rv) {
  /* the raw kernel on mean anomalies: nmot = 1, tp = 0 reproduces M exactly */
  kepler_rv_add(M, n, 1.0, 0.0, e, K, w, rv, 0);
  return 0;
}

/* param.py:299-362 + 88-105. Returns 0 when the planet is valid, 1 = ValueError. */
static int to_default(int par, const double* in, double* P, double* K, double* e, double* w,
                      double* tp, int* conv_error) {
  *conv_error = 0;
  *P = in[0];
  *K = in[1];
  if (par == RVLP_PAR_PKEWTP || par == RVLP_PAR_PKEWTC) {
    *e = in[2];
    *w = in[3];
  } else {
    *e = in[2] * in[2] + in[3] * in[3];                    /* param.py:232-233 */
    *w = atan2(in[3], in[2]);
  }
  if (par == RVLP_PAR_PKEWTC || par == RVLP_PAR_PKSECTC) { /* param.py:198-215 */
    double th = (ORACLE_PI / 2) - *w;
    if (*e < 0 || *e >= 1.0) {
      *conv_error = 1;
      *tp = NAN;
      return 1;
    }
    double E = 2 * atan(sqrt((1 - *e) / (1 + *e)) * tan(th / 2));
    double Mc = E - (*e * sin(E));
    *tp = in[4] - (*P / (2 * ORACLE_PI)) * Mc;
  } else {
    *tp = in[4];
  }
  if (*P <= 0) return 1;                                   /* param.py:26 */
  if (*K <= 0) return 1;                                   /* param.py:43 */
  if (*e < 0 || *e >= 1.0) return 1;                       /* param.py:60-63 */
  if (!(-ORACLE_PI <= *w && *w < ORACLE_PI)) return 1;     /* param.py:81 */
  return 0;
}

int oracle_convert_to_default(int par, const double* in, int64_t n, double* out, int32_t* valid) {
  for (int64_t i = 0; i < n; ++i) {
    int ce;
    int bad = to_default(par, in + 5 * i, out + 5 * i, out + 5 * i + 1, out + 5 * i + 2,
                         out + 5 * i + 3, out + 5 * i + 4, &ce);
    if (valid) valid[i] = !bad;
  }
  return 0;
}

/* prior.py — see include/ravest_b200.h for the p[] / c[] layout */
double oracle_prior(const rvlp_prior* pr, double x) {
  const double* p = pr->p;
  switch (pr->kind) {
    case RVLP_PRIOR_UNIFORM:                               /* prior.py:62-65 */
      return (x < p[0] || x > p[1]) ? -INFINITY : -log(p[1] - p[0]);
    case RVLP_PRIOR_ECC_UNIFORM:                           /* prior.py:119-122 */
      return (x < 0 || x >= p[0]) ? -INFINITY : -log(p[0]);
    case RVLP_PRIOR_NORMAL: {                              /* prior.py:156,171 */
      double z = (x - p[0]) / p[1];
      return -0.5 * (z * z) - pr->c[0];
    }
    case RVLP_PRIOR_TRUNC_NORMAL: {                        /* prior.py:243-246, scipy truncnorm.logpdf */
      if (x < p[2] || x > p[3]) return -INFINITY;
      double z = (x - p[0]) / p[1];
      return -(z * z) / 2.0 + pr->c[0];
    }
    case RVLP_PRIOR_HALF_NORMAL: {                         /* prior.py:300-303, scipy halfnorm.logpdf */
      if (x < 0.0) return -INFINITY;
      double y = x / p[0];
      return -(y * y) / 2.0 + pr->c[0];
    }
    case RVLP_PRIOR_RAYLEIGH: {                            /* prior.py:356-359, scipy rayleigh.logpdf */
      if (x < 0.0 || x == INFINITY) return -INFINITY;      /* scipy: -inf outside the open support */
      double r = x / p[0];
      return log(r) - 0.5 * r * r + pr->c[0];
    }
    case RVLP_PRIOR_VANEYLEN19: {                          /* prior.py:431-440 */
      if (x < 0.0 || x == INFINITY) return -INFINITY;
      double y = x / p[0], r = x / p[1], f = p[2];
      double lh = -(y * y) / 2.0 + pr->c[0];
      double lr = log(r) - 0.5 * r * r + pr->c[1];
      /* scipy.special.logsumexp(a, b=[1-f, f]): zero-weight terms are dropped before the max */
      if (1.0 - f == 0.0) lh = -INFINITY;
      if (f == 0.0) lr = -INFINITY;
      double m = lh > lr ? lh : lr;
      if (m == -INFINITY) return -INFINITY;
      double s = 0.0;
      if (lh != -INFINITY) s += (1.0 - f) * exp(lh - m);
      if (lr != -INFINITY) s += f * exp(lr - m);
      return log(s) + m;
    }
    case RVLP_PRIOR_BETA: {                                /* prior.py:503-508 */
      if (x < 0.0 || x > 1.0) return -INFINITY;
      double a1 = p[0] - 1.0, b1 = p[1] - 1.0;
      double t1 = (a1 == 0.0 && !isnan(x)) ? 0.0 : a1 * log(x);        /* xlogy   */
      double t2 = (b1 == 0.0 && !isnan(x)) ? 0.0 : b1 * log1p(-x);     /* xlog1py */
      return t1 + t2 - pr->c[0];
    }
  }
  return NAN;
}

static int n_model_of(const rvlp_desc* d) { return 5 * d->n_planets + 2 + 2 * d->n_inst; }

static inline double param_value(const rvlp_desc* d, const double* row, int i) {
  return d->src_col[i] >= 0 ? row[d->src_col[i]] : d->src_const[i];
}

/* One sample of LogPosterior.log_probability (fit.py:3448-3495).  scratch: [n_epochs] doubles.
 * gp != 0 evaluates GPLogPosterior.log_probability (fit.py:7836-7901) with scratch2 [N*N + N]. */
static double logprob_one(const rvlp_desc* d, const double* t, const double* vel,
                          const double* velerr_sq, const int32_t* inst, int64_t N,
                          const double* row, double* rv, double* ll_out, double* lp_out,
                          int gp, double* scratch2) {
  const int npl = d->n_planets, nin = d->n_inst, nm = n_model_of(d);
  const int i_gd = 5 * npl, i_g = i_gd + 2, i_jit = i_g + nin;
  if (ll_out) *ll_out = NAN;
  if (lp_out) *lp_out = NAN;
  /* `parts` mode (ll_out != NULL) evaluates the likelihood even where the posterior fast-fails */
  const int parts = ll_out != NULL;
  int reject = 0;

  for (int j = 0; j < nin; ++j)                            /* fit.py:3465-3468 */
    if (param_value(d, row, i_jit + j) < 0) reject = 1;
  if (reject && !parts) return -INFINITY;

  double hyp[4] = {0, 0, 0, 0};
  if (gp) {                                                /* gp.py:98-108 */
    for (int k = 0; k < 4; ++k) {
      hyp[k] = param_value(d, row, nm + k);
      if (!isfinite(hyp[k]) || hyp[k] <= 0) return -INFINITY;
    }
  }

  /* planets: conversion + validity */
  double P[64], K[64], e[64], w[64], tp[64];
  int invalid = 0, conv_err = 0;
  for (int k = 0; k < npl; ++k) {
    double in[5];
    for (int q = 0; q < 5; ++q) in[q] = param_value(d, row, 5 * k + q);
    int ce;
    int bad = to_default(d->parameterisation, in, &P[k], &K[k], &e[k], &w[k], &tp[k], &ce);
    invalid |= bad;
    conv_err |= ce;
  }

  /* priors (fit.py:3475-3482, 3672-3691; hyper-priors fit.py:7884-7886) */
  double lp = 0.0, lhp = 0.0;
  int needs_conv = 0;
  for (int j = 0; j < d->n_priors; ++j)
    if (d->priors[j].target != RVLP_TARGET_COLUMN) needs_conv = 1;
  if (needs_conv && conv_err) {                            /* ValueError in conversion */
    if (!parts) return -INFINITY;
    reject = 1;
  }
  for (int j = 0; j < d->n_priors; ++j) {
    const rvlp_prior* pr = &d->priors[j];
    double x;
    switch (pr->target) {
      case RVLP_TARGET_COLUMN: x = row[pr->index]; break;
      case RVLP_TARGET_P: x = P[pr->index]; break;
      case RVLP_TARGET_K: x = K[pr->index]; break;
      case RVLP_TARGET_E: x = e[pr->index]; break;
      case RVLP_TARGET_W: x = w[pr->index]; break;
      default: x = tp[pr->index]; break;
    }
    if (pr->is_hyper) lhp += oracle_prior(pr, x);
    else lp += oracle_prior(pr, x);
  }
  if (lp_out) *lp_out = lp;
  if (!isfinite(lp) || (gp && !isfinite(lhp))) {
    if (!parts) return -INFINITY;
    reject = 1;
  }

  /* likelihood (fit.py:3600-3660) */
  double ll;
  if (invalid) {
    ll = -INFINITY;                                        /* fit.py:3625-3627 */
  } else {
    for (int64_t i = 0; i < N; ++i) rv[i] = 0.0;
    for (int k = 0; k < npl; ++k) {
      double nmot = 2 * ORACLE_PI / P[k];                  /* model.py:302 */
      kepler_rv_add(t, N, nmot, tp[k], e[k], K[k], w[k], rv, 1);
    }
    double gd = param_value(d, row, i_gd), gdd = param_value(d, row, i_gd + 1);
    for (int64_t i = 0; i < N; ++i) {                      /* model.py:483-509 */
      double tr = 0.0;
      if (gd != 0) tr += gd * (t[i] - d->t0);
      if (gdd != 0) tr += gdd * ((t[i] - d->t0) * (t[i] - d->t0));
      rv[i] += tr;
      rv[i] += param_value(d, row, i_g + inst[i]);         /* fit.py:3642-3644 */
    }
    if (!gp) {
      const double log2pi = log(2 * ORACLE_PI);
      double sum = 0.0;
      for (int64_t i = 0; i < N; ++i) {                    /* fit.py:3652-3658 */
        double jit = param_value(d, row, i_jit + inst[i]);
        double var = velerr_sq[i] + jit * jit;
        double res = rv[i] - vel[i];
        sum += res * res / var + (log2pi + log(var));
      }
      ll = -0.5 * sum;
    } else {
      /* fit.py:8062-8105 + gp.py:145-156; dense Cholesky */
      int finite = 1;
      for (int64_t i = 0; i < N; ++i) finite &= isfinite(rv[i]) != 0;
      if (!finite) {
        ll = -INFINITY;                                    /* fit.py:8082-8083 */
      } else {
        double A = hyp[0], le = hyp[1], lpp = hyp[2], Pg = hyp[3];
        double gamma = 1.0 / (2.0 * (lpp * lpp));
        double* C = scratch2;
        double* r = scratch2 + N * N;
        for (int64_t i = 0; i < N; ++i) {
          for (int64_t j = 0; j <= i; ++j) {
            double tau = t[i] - t[j];
            double sn = sin(ORACLE_PI * fabs(tau) / Pg);
            double q = tau / le;
            C[i * N + j] = (A * A) * exp(-gamma * (sn * sn)) * exp(-0.5 * (q * q));
          }
          double jit = param_value(d, row, i_jit + inst[i]);
          C[i * N + i] += velerr_sq[i] + jit * jit;
          r[i] = vel[i] - rv[i];
        }
        double logdet = 0.0, quad = 0.0;
        int ok = 1;
        for (int64_t j = 0; j < N && ok; ++j) {            /* Cholesky-Banachiewicz, row by row */
          for (int64_t k2 = 0; k2 <= j; ++k2) {
            double s = C[j * N + k2];
            for (int64_t m = 0; m < k2; ++m) s -= C[j * N + m] * C[k2 * N + m];
            if (k2 == j) {
              if (!(s > 0)) { ok = 0; break; }
              C[j * N + j] = sqrt(s);
            } else {
              C[j * N + k2] = s / C[k2 * N + k2];
            }
          }
        }
        if (!ok) {
          ll = NAN;                                        /* jax cholesky yields NaN */
        } else {
          for (int64_t i = 0; i < N; ++i) {
            double s = r[i];
            for (int64_t m = 0; m < i; ++m) s -= C[i * N + m] * r[m];
            r[i] = s / C[i * N + i];
            quad += r[i] * r[i];
            logdet += log(C[i * N + i]);
          }
          ll = -0.5 * quad - logdet - 0.5 * (double)N * log(2 * ORACLE_PI);
        }
      }
    }
  }
  if (ll_out) *ll_out = ll;
  if (reject) return -INFINITY;
  double logprob = gp ? (ll + lp + lhp) : (ll + lp);       /* fit.py:3492-3495 / 7898-7900 */
  logprob += d->jacobian;
  logprob += d->renorm;
  return logprob;
}

static int run_batch(const rvlp_desc* d, const double* t, const double* vel, const double* velerr,
                     const int32_t* inst, int64_t N, const double* theta, int64_t S, double* out,
                     double* ll_out, double* lp_out, int gp, int nthreads) {
  if (d->n_planets > 64) return -1;
  double* vsq = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  for (int64_t i = 0; i < N; ++i) vsq[i] = velerr[i] * velerr[i];  /* fit.py:3598 */
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* rv = (double*)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    double* s2 = gp ? (double*)malloc(sizeof(double) * (size_t)(N * N + N + 1)) : NULL;
#pragma omp for schedule(dynamic, 16)
    for (int64_t s = 0; s < S; ++s) {
      double ll, lp;
      double v = logprob_one(d, t, vel, vsq, inst, N, theta + s * d->ndim, rv, &ll, &lp, gp, s2);
      if (out) out[s] = v;
      if (ll_out) ll_out[s] = ll;
      if (lp_out) lp_out[s] = lp;
    }
    free(rv);
    free(s2);
  }
  free(vsq);
  return 0;
}

int oracle_logprob_batch(const rvlp_desc* d, const double* t, const double* vel,
                         const double* velerr, const int32_t* inst, int64_t N,
                         const double* theta, int64_t S, double* out, int nthreads) {
  return run_batch(d, t, vel, velerr, inst, N, theta, S, out, NULL, NULL, 0, nthreads);
}

int oracle_logprob_parts_batch(const rvlp_desc* d, const double* t, const double* vel,
                               const double* velerr, const int32_t* inst, int64_t N,
                               const double* theta, int64_t S, double* ll, double* lp, int nthreads) {
  return run_batch(d, t, vel, velerr, inst, N, theta, S, NULL, ll, lp, 0, nthreads);
}

int oracle_gp_logprob_batch(const rvlp_desc* d, const double* t, const double* vel,
                            const double* velerr, const int32_t* inst, int64_t N,
                            const double* theta, int64_t S, double* out, int nthreads) {
  if (d->n_hyper != 4) return -1;
  return run_batch(d, t, vel, velerr, inst, N, theta, S, out, NULL, NULL, 1, nthreads);
}

/* fit.py:2690-2824; component >= 0 planet, -1 trend, -2 total. Invalid planet -> NaN row. */
int oracle_rv_batch(const rvlp_desc* d, const double* theta, int64_t S, const double* times,
                    int64_t T, int32_t component, double* out, int nthreads) {
  const int npl = d->n_planets;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t s = 0; s < S; ++s) {
    const double* row = theta + s * d->ndim;
    double* rv = out + s * T;
    int bad = 0;
    for (int64_t i = 0; i < T; ++i) rv[i] = 0.0;
    if (component == RVLP_RV_TREND || component == RVLP_RV_TOTAL) {
      double gd = param_value(d, row, 5 * npl), gdd = param_value(d, row, 5 * npl + 1);
      for (int64_t i = 0; i < T; ++i) {
        double tr = 0.0;
        if (gd != 0) tr += gd * (times[i] - d->t0);
        if (gdd != 0) tr += gdd * ((times[i] - d->t0) * (times[i] - d->t0));
        rv[i] += tr;
      }
    }
    for (int k = 0; k < npl; ++k) {
      if (!(component == k || component == RVLP_RV_TOTAL)) continue;
      double in[5], P, K, e, w, tp;
      int ce;
      for (int q = 0; q < 5; ++q) in[q] = param_value(d, row, 5 * k + q);
      if (to_default(d->parameterisation, in, &P, &K, &e, &w, &tp, &ce)) { bad = 1; continue; }
      kepler_rv_add(times, T, 2 * ORACLE_PI / P, tp, e, K, w, rv, 1);
    }
    if (bad) for (int64_t i = 0; i < T; ++i) rv[i] = NAN;
  }
  return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
