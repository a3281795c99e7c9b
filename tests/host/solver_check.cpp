// Host-side exhaustive check of the Kepler solver in ravest_b200/csrc/rvlp_math.cuh.
// Build: g++ -O2 -std=c++17 -mfma -o solver_check solver_check.cpp -lm
// Prints one line per eccentricity: max backward error |E - e sin E - m| (long double),
// max forward error in (cos E, sin E), fallback fraction.  Exit code 1 if a bound is broken.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include "../../ravest_b200/csrc/rvlp_math.cuh"

using namespace rvlp;

static long double true_E(long double m, long double e) {
  long double lo = m, hi = m + e;
  if (hi > M_PIl) hi = M_PIl;
  long double E = hi;
  for (int it = 0; it < 300; ++it) {
    long double f = E - e * sinl(E) - m;
    if (f > 0) hi = E; else lo = E;
    long double En = E - f / (1 - e * cosl(E));
    if (!(En > lo && En < hi)) En = 0.5L * (lo + hi);
    if (fabsl(En - E) < 1e-19L) { E = En; break; }
    E = En;
  }
  return E;
}

int main(int argc, char** argv) {
  int nM = argc > 1 ? atoi(argv[1]) : 200000;
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  const double es[] = {1e-300, 1e-9, 0.01, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.62, 0.65, 0.66, 0.7, 0.72, 0.75, 0.8, 0.85, 0.9,
                       0.93, 0.95, 0.97, 0.975, 0.98, 0.99, 0.995, 0.999, 0.9995, 0.99999};
  int bad = 0;
  for (double e : es) {
    SolverPlan plan = plan_for(e);
    double max_fwd = 0, max_bwd = 0, max_rv = 0;
    long nfb = 0, ntot = 0;
    for (int j = 0; j < nM; j += 2) {
      double M[2], cE[2], sE[2], dl[2], ri[2];
      for (int q = 0; q < 2; ++q) {
        double u = U(rng);
        int mode = (j / 2 + q) % 4;
        if (mode == 0) M[q] = (2 * u - 1) * M_PI;                       // uniform in [-pi, pi]
        else if (mode == 1) M[q] = (2 * u - 1) * 6000.0;                // many revolutions
        else if (mode == 2) M[q] = std::pow(10.0, -12 * u) * ((j & 2) ? 1 : -1);   // near periastron
        else M[q] = 2 * M_PI * std::floor(u * 100) + std::pow(10.0, -8 * U(rng)) * (1 - e) ;  // cusp region
      }
      if (plan.n64 == 0) kepler_fast<2, 1, 0>(M, e, 1, 0, cE, sE, dl, ri); else kepler_fast<2>(M, e, plan.n32, plan.n64, cE, sE, dl, ri);
      for (int q = 0; q < 2; ++q) {
        ++ntot;
        if (step_rejected(dl[q], plan.tol) || anomaly_is_big(M[q])) { ++nfb; CosSin cs = kepler_robust(M[q], e); cE[q] = cs.c; sE[q] = cs.s; ri[q] = 1.0 / (1.0 - e * cs.c); }
        // exact reduction in long double for the truth
        long double Ml = (long double)M[q];
        long double k = rintl(Ml / (2 * M_PIl));
        long double ml = fabsl(Ml - k * 2 * M_PIl);
        long double Et = true_E(ml, (long double)e);
        long double ct = cosl(Et), st = sinl(Et) * ((Ml - k * 2 * M_PIl) < 0 ? -1 : 1);
        double fwd = std::max(fabs((double)(cE[q] - ct)), fabs((double)(sE[q] - st)));
        // rv-relevant: (cosE - e)/(1 - e cosE) and sinE/(1 - e cosE) scaled errors
        long double dt = 1 - e * ct, dc = 1.0L / (long double)ri[q];
        double rverr = std::max(fabs((double)((cE[q] - e) / dc - (ct - e) / dt)),
                                fabs((double)(sE[q] / dc - st / dt))) * sqrt(1 - e * e);
        // conditioning of the reference's own answer: ulp(M) / (1 - e cos E)
        double cond = (double)(fabs(M[q]) * 2.2e-16 / dt / dt + 4e-16 / dt);
        if (fwd > max_fwd) max_fwd = fwd;
        if (rverr / cond > max_rv) max_rv = rverr / cond;
        long double Ec = atan2l((long double)fabs(sE[q]), (long double)cE[q]);
        double bwd = fabs((double)(Ec - e * sinl(Ec) - ml));
        if (bwd > max_bwd) max_bwd = bwd;
      }
    }
    printf("e=%-8g n32=%d n64=%d tol=%.1e  fallback=%.4f%%  max|d(cos,sin)|=%.2e  backward=%.2e  rv_err/cond=%.2f\n",
           e, plan.n32, plan.n64, plan.tol, 100.0 * nfb / ntot, max_fwd, max_bwd, max_rv);
    if (max_rv > 8.0) bad = 1;
  }
  return bad;
}
