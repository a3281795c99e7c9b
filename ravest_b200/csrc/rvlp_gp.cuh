// rvlp_gp.cuh — K3: batched quasi-periodic GP log-posterior (config 5).
//
// Replaces GPLogPosterior.log_probability / GPLogLikelihood.__call__
// (/root/reference/src/ravest/fit.py:7836-7901, 8062-8105) and the kernel of gp.py:126-156.
// The dense algebra lives in tinygp 0.3.0 (absent, "parity unpinned"): restated as
//   C_ij = A^2 exp(-Gamma sin^2(pi |t_i - t_j| / P_gp)) exp(-(t_i - t_j)^2 / (2 lambda_e^2))
//          + delta_ij (sigma_i^2 + jit_i^2),            Gamma = 1 / (2 lambda_p^2)
//   C = L L^T,  alpha = L^-1 (v - mean),  ll = -1/2 alpha.alpha - sum log L_ii - N/2 log 2 pi.
//
// This header holds what the GP kernels share.  The kernels themselves:
//   rvlp_gp_pipe.cuh   N <= 219 epochs: one CTA per sample, software-pipelined register-tile Cholesky, factor in shared
//                      memory (K3, K7) - small batches, and 81..149 epochs
//   rvlp_gp_batch.cuh  any N <= 16384: level-synchronous batched Cholesky, one kernel per block column over all samples,
//                      16 x 16 blocks in a global workspace, DMMA (fp64 tensor core) Gram sums
// (The lock-step shared-memory / column / blocked predecessors of round 1 and a one-CTA-per-sample global-workspace
// kernel tried in round 2 were removed; each path is the other's independent second implementation in the tests,
// `RVLP_GP_KERNEL=pipe|batch`.)
#pragma once
#include "rvlp_kernels.cuh"
#include "rvlp_gpcov.cuh"

namespace rvlp {

// 1 / sqrt(x) for the pivots of the diagonal-tile factorisation, which sits on K3's critical path (one thread
// works, the CTA waits): MUFU.RSQ64H seed (~2^-20) + two Newton-Raphson steps (-> 2^-40 -> full precision) instead
// of libm's rsqrt.  Negative / zero / NaN pivots (matrix not positive definite) still give NaN / inf, as jax.
__device__ __forceinline__ double pivot_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}

// D(8x8) += A(8x4) B(4x8), fp64 tensor-core MMA.  Fragments (PTX ISA, m8n8k4 .f64): a = A[lane / 4][lane % 4],
// b = B[lane % 4][lane / 4], {c0, c1} = C[lane / 4][2 (lane % 4) + {0, 1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Register-tile size of the pipelined kernels for a given epoch count: 22 tile rows of TT epochs (incl. the residual
// row) per CTA; 0 = too many epochs for them (rvlp_gp_big.cuh takes over).
__host__ __device__ inline int gp_tile_for(int n_epochs) {
  const int t[5] = {2, 4, 6, 8, 10};
  for (int i = 0; i < 5; ++i)
    if (n_epochs + 1 <= 22 * t[i]) return t[i];
  return 0;
}

}  // namespace rvlp
