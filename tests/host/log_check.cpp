// Host-side accuracy check of log_pos_normal (ravest_b200/csrc/rvlp_math.cuh) against long-double logl.
// Build: g++ -O2 -std=c++17 -mfma -ffp-contract=off -o log_check log_check.cpp -lm
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../../ravest_b200/csrc/rvlp_math.cuh"
using namespace rvlp;
int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 4000000;
  std::mt19937_64 rng(99);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  double worst = 0;
  int bad = 0;
  for (long i = 0; i < n; ++i) {
    double x;
    const double u = U(rng);
    if (u < 0.5) x = std::ldexp(1.0 + U(rng), (int)(U(rng) * 64));             // mantissa products: [1, 2^64)
    else if (u < 0.8) x = std::ldexp(1.0 + U(rng), (int)(U(rng) * 2040) - 1020);  // any normal exponent
    else x = 1.0 + std::ldexp(U(rng), -(int)(U(rng) * 50));                     // close to 1
    const long double want = logl((long double)x);
    const double got = log_pos_normal(x);
    const double err = (double)fabsl(got - want);
    const double allow = 5e-16 + 2.5e-16 * std::fabs((double)want);   // ~1 ulp of the result (ln 2 is one double)
    if (err > allow) { if (bad < 5) printf("BAD x = %.17g got %.17g want %.17Lg err %.3e\n", x, got, want, err); ++bad; }
    worst = std::fmax(worst, err / allow);
  }
  for (int j = 0; j < 64; ++j) {                                                // table bin edges
    const double lo = 1.0 + j / 64.0, hi = std::nextafter(1.0 + (j + 1) / 64.0, 0.0);
    for (double x : {lo, hi}) {
      const double err = (double)fabsl(log_pos_normal(x) - logl((long double)x));
      if (err > 5e-16) { printf("BAD edge x = %.17g err %.3e\n", x, err); ++bad; }
    }
  }
  if (log_pos_normal(1.0) != 0.0 && std::fabs(log_pos_normal(1.0)) > 2e-17) { printf("BAD log(1) = %.3e\n", log_pos_normal(1.0)); ++bad; }
  printf("n = %ld: worst error / allowance = %.3f, bad = %d\n%s\n", n, worst, bad, bad ? "FAILED" : "OK");
  return bad ? 1 : 0;
}
