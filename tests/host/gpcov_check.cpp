// Host-side accuracy check of the branch-free GP covariance function in ravest_b200/csrc/rvlp_gpcov.cuh.
// Build: g++ -O2 -std=c++17 -mfma -ffp-contract=off -o gpcov_check gpcov_check.cpp -lm
// (contraction off: the algorithm is tested as written; a contracted |tau| * (1/P) - rint(.) is MORE accurate
// than the rounded quotient this check feeds to the long-double side)
// Compares k(tau) with a long-double evaluation of the same formula (gp.py:145-156) over hyperparameter
// ranges well beyond config 5, and checks the edge inputs.  The rounding of u = |tau| / P_gp (shared with
// the reference, which forms the same quotient in double) is taken out of the comparison: the long-double
// value is computed from the SAME double u and q.  Exit code 1 if a bound is broken.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../../ravest_b200/csrc/rvlp_gpcov.cuh"

using namespace rvlp;

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 2000000;
  std::mt19937_64 rng(777);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  double max_rel = 0, max_sin = 0, max_exp = 0;
  int bad = 0;
  for (long i = 0; i < n; ++i) {
    const double A = 0.01 + 30.0 * U(rng), le = std::pow(10.0, -1.0 + 4.0 * U(rng));
    const double lp = std::pow(10.0, -1.5 + 2.5 * U(rng)), Pg = std::pow(10.0, -0.5 + 3.0 * U(rng));
    const double tau = (U(rng) < 0.05 ? 0.0 : (U(rng) - 0.5) * 2.0 * std::pow(10.0, -3.0 + 6.5 * U(rng)));
    const GpHyper h = gp_hyper(A, le, lp, Pg);
    const double got = gp_cov(tau, h);
    const double u = std::fabs(tau) * h.inv_P, q = tau * h.inv_le;
    const long double ul = (long double)u - rintl((long double)u);
    const long double s = sinl(M_PIl * fabsl(ul));
    const long double y = -(long double)h.gamma * s * s - 0.5L * (long double)q * (long double)q;
    const long double want = (long double)h.A2 * expl(y);
    // sin and exp separately
    const double sg = gp_sinpi_frac(u);
    max_sin = std::fmax(max_sin, (double)fabsl(sg - s));
    const double yd = (double)y;
    if (yd > -700) {
      const double eg = gp_exp_scaled(yd, 1.0);
      max_exp = std::fmax(max_exp, (double)(fabsl(eg - expl((long double)yd)) / expl((long double)yd)));
    }
    if (want > 1e-290L) {
      // the error of y is absolute (gamma * 2 s * d_sin + rounding of y itself): allow |y| ulps on top
      const double rel = (double)(fabsl(got - want) / want);
      const double allow = 4e-16 + 2.3e-16 * std::fabs(yd) + (double)h.gamma * 3e-16;
      if (rel > allow) {
        if (bad < 10) printf("BAD rel %.3e allow %.3e: A %.4g le %.4g lp %.4g P %.4g tau %.17g\n", rel, allow, A, le, lp, Pg, tau);
        ++bad;
      }
      if (std::fabs(yd) < 30 && h.gamma < 50) max_rel = std::fmax(max_rel, rel);
    } else if (!(got >= 0.0 && got < 1e-280)) {
      printf("BAD tiny: got %.3e want %.3Le\n", got, want);
      ++bad;
    }
  }
  printf("n = %ld: max |d sin| = %.3e, max rel exp = %.3e, max rel k (|y| < 30, gamma < 50) = %.3e, bad = %d\n", n, max_sin, max_exp,
         max_rel, bad);
  if (max_sin > 2.5e-16 || max_exp > 3.5e-16) { printf("FAIL: component bound\n"); bad++; }
  // edge inputs: no crash, sensible values
  const GpHyper h = gp_hyper(3.0, 20.0, 0.5, 15.0);
  const double k0 = gp_cov(0.0, h);
  if (k0 != 9.0) { printf("FAIL: k(0) = %.17g\n", k0); bad++; }
  if (gp_cov(1e6, h) > 1e-300 || gp_cov(1e6, h) < 0) { printf("FAIL: far tail %.3e\n", gp_cov(1e6, h)); bad++; }
  if (!std::isnan(gp_cov(NAN, h))) { printf("FAIL: NaN tau\n"); bad++; }
  const double kinf = gp_cov(INFINITY, h);
  if (!(std::isnan(kinf) || kinf < 1e-300)) { printf("FAIL: inf tau %.3e\n", kinf); bad++; }
  const GpHyper hz = gp_hyper(3.0, 1e-320, 0.5, 1e-320);   // reciprocals overflow
  const double kz = gp_cov(1.0, hz), kzz = gp_cov(0.0, hz);
  if (!(std::isnan(kz) || kz < 1e-300)) { printf("FAIL: overflowed hyper %.3e\n", kz); bad++; }
  (void)kzz;
  // symmetric in tau
  for (int i = 0; i < 1000; ++i) {
    const double t = (U(rng) - 0.5) * 500;
    if (gp_cov(t, h) != gp_cov(-t, h)) { printf("FAIL: asymmetric at %.17g\n", t); bad++; break; }
  }
  printf(bad ? "FAILED\n" : "OK\n");
  return bad ? 1 : 0;
}
