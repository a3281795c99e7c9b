// rvlp_gp_big.cuh — K3b: the quasi-periodic GP log-posterior (and the solve half of the conditioning path) for MANY
// epochs: N >= 220, where the packed factor no longer fits the shared memory / register tiles of rvlp_gp_pipe.cuh.
//
// Same arithmetic contract as K3 (GPLogPosterior.log_probability, /root/reference/src/ravest/fit.py:7836-7901,
// 8062-8105; kernel gp.py:145-156; conditioning fit.py:7494-7554, chi^2 fit.py:5386-5429):
//   C = K(t, t) + diag(sigma^2 + jit^2) = L L^T,  alpha = L^-1 r,  ll = -1/2 alpha.alpha - sum ln L_ii - N/2 ln 2 pi.
//
// One CTA (8 warps) per sample, two CTAs per SM, persistent over the samples.  The factor lives in a per-CTA global
// workspace ((N_pad + 17) x N_pad doubles: 2.2 MB at N = 512, L2-resident across the 296 CTAs up to N ~ 250 and streamed
// from HBM beyond), row-major, in 16 x 16 blocks.  LEFT-looking blocked Cholesky, per block column J:
//   1. every warp takes row blocks I = J + warp, J + warp + 8, ...: S = sum_{K<J} L_IK L_JK^T on the TENSOR CORES -
//      `mma.sync.aligned.m8n8k4.f64` (SASS DMMA.884), operands loaded straight from the workspace in fragment layout
//      (one 32-byte row segment per quad), four 8 x 8 accumulator tiles per warp - then C_IJ - S with C generated on the
//      fly by the branch-free covariance function (rvlp_gpcov.cuh); the residual r rides along as row block N_pad/16,
//      so that its forward substitution alpha = L^-1 r falls out of the same sweeps;
//   2. warp 0, which owns the diagonal block, factors it in shared memory (16 rank-1 steps);
//   3. one thread per row below solves X L_JJ^T = P against the shared-memory block (the block's TRSM).
// Two __syncthreads per block column.  Rows / columns N..N_pad-1 are identity padding (ln 1 = 0, alpha = 0).
// With PRED the kernel continues with the blocked back substitution beta = L^-T alpha (thread per column, coalesced)
// and writes chi^2 = alpha.alpha and beta for gp_mean_kernel.
//
// Deterministic: a sample's bits depend on its own row only (fixed task -> warp map, fixed DMMA accumulation order).
#pragma once
#include "rvlp_gp.cuh"

namespace rvlp {

constexpr int kBigNB = 16;
constexpr int kBigMaxWarps = 8;
constexpr int kBigDiagLd = 17;      // padded leading dimension of the shared-memory diagonal block

__host__ __device__ inline int gp_big_npad(int N) { return (N + kBigNB - 1) / kBigNB * kBigNB; }
// doubles of workspace per CTA: N_pad rows of L, 16 rows for the residual block, 1 row holding the raw residual
__host__ __device__ inline size_t gp_big_ws_doubles(int N) {
  const size_t np = (size_t)gp_big_npad(N);
  return (np + kBigNB + 1) * np;
}

struct GpBigSmem { int off_diag, off_inv, off_red, off_z, total; };
__host__ __device__ inline GpBigSmem gp_big_smem(const DevProblem& P, const SmemLayout& L, bool pred) {
  GpBigSmem G;
  int o = (L.total + 15) & ~15;
  G.off_diag = o; o += kBigNB * kBigDiagLd * 8;
  G.off_inv = o; o += kBigNB * 8;
  G.off_red = o; o += (kBigMaxWarps + 2) * 8;
  G.off_z = o; o += pred ? gp_big_npad(P.n_epochs) * 8 : 0;
  G.total = o;
  return G;
}

// NW warps per CTA (= per sample), MB CTAs per SM.
template <bool PRED, int NW, int MB>
__global__ void __launch_bounds__(32 * NW, MB)
gp_big_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, double* __restrict__ out,
              double* __restrict__ beta_out, double* ws_all) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);              // P.epochs_global == 1: the epoch arrays stay in global memory
  const GpBigSmem G = gp_big_smem(P, L, PRED);
  stage_problem(P, L, smem);
  const Tables T = tables_of<true>(P, L, smem);
  constexpr int kBigThreads = 32 * NW, kBigWarps = NW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  double* sr_w = reinterpret_cast<double*>(smem + L.off_scratch);
  double* diag = reinterpret_cast<double*>(smem + G.off_diag);
  double* invd = reinterpret_cast<double*>(smem + G.off_inv);
  double* red = reinterpret_cast<double*>(smem + G.off_red);
  const int N = P.n_epochs;
  const int np = gp_big_npad(N), ld = np;
  const int NBk = np / kBigNB;                      // block columns; row block NBk is the residual block
  double* W = ws_all + (size_t)blockIdx.x * gp_big_ws_doubles(N);
  double* resid = W + (size_t)(np + kBigNB) * ld;   // the raw residual, [np]
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  // rows np+1 .. np+15 of the residual block are never used: keep them zero so that the tensor-core tiles that carry
  // them along compute zeros (the workspace arrives uninitialised)
  for (size_t i = tid; i < (size_t)(kBigNB - 1) * ld; i += kBigThreads) W[(size_t)(np + 1) * ld + i] = 0.0;

  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    __syncthreads();
    if (warp == 0)
      sample_prologue(P, T, theta, s, s + 1, sr_w, rec, lane, !PRED, 1, reinterpret_cast<double*>(smem + L.off_pv));
    __syncthreads();
    const double* sr = sr_w;
    const int flags = __double2loint(sr[1]);
    const double lp = sr[0], lhp = sr[4];
    if (PRED ? (flags & (F_PLANET | F_HYPER)) : (flags & (F_JIT | F_HYPER | F_PRIOR))) {
      if (PRED) {                                            // the reference raises: NaN rows
        for (int j = tid; j < N; j += kBigThreads) beta_out[s * N + j] = qnan;
        if (tid == 0 && out) out[s] = qnan;
      } else if (tid == 0) {
        out[s] = -INFINITY;                                  // fit.py:7857-7886
      }
      continue;
    }
    int nonfinite = (!PRED && (flags & F_PLANET)) ? 1 : 0;  // fit.py:8022-8024
    const double* row = theta + s * P.ndim;
    const GpHyper hyp = gp_hyper(model_param(T, row, P.n_model + 0), model_param(T, row, P.n_model + 1),
                                 model_param(T, row, P.n_model + 2), model_param(T, row, P.n_model + 3));
    if (!nonfinite) {
      for (int i = tid; i < np; i += kBigThreads) {          // residual, fit.py:7994-8043, 8059 / 7543-7550
        double r = 0.0;
        if (i < N) {
          double tt[1] = {T.t[i]}, rv[1];
          model_rv<1>(P, sr, tt, rv, -1, true);
          if (PRED) {
            r = (T.v[i] - sr[kHdr + T.inst[i]]) - rv[0];
          } else {
            const double mean = rv[0] + sr[kHdr + T.inst[i]];
            if (!(fabs(mean) <= 1.79769313486231570e308)) nonfinite = 1;
            r = T.v[i] - mean;
          }
        }
        resid[i] = r;
      }
    }
    if (__syncthreads_or(nonfinite)) {                       // fit.py:8082-8083
      if (tid == 0) {
        double r = -INFINITY + lp + lhp;
        r += P.jacobian;
        r += P.renorm;
        out[s] = r;
      }
      continue;
    }

    // C[i][j] of the augmented, padded matrix (lower triangle; the diagonal block's upper half is never used)
    auto entry = [&](int i, int j) -> double {
      if (i >= np) return i == np ? resid[j] : 0.0;          // residual block: row 0 = r, rows 1..15 unused
      if (i >= N || j >= N) return i == j ? 1.0 : 0.0;       // identity padding
      double c = gp_cov(T.t[i] - T.t[j], hyp);               // gp.py:145-156
      if (i == j) c += T.e2[i] + sr[kHdr + P.n_inst + T.inst[i]];   // fit.py:8094-8096
      return c;
    };

    double logdet = 0.0;                                     // lanes 0..15 of warp 0: sum of ln L_kk of "their" row
    for (int J = 0; J < NBk; ++J) {
      const int kend = J * kBigNB;
      // ---- 1. panel update on the tensor cores
      for (int I = J + warp; I <= NBk; I += kBigWarps) {
        double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
        const double* a0p = W + (size_t)(I * kBigNB + g) * ld + q;
        const double* a1p = a0p + (size_t)8 * ld;
        const double* b0p = W + (size_t)(J * kBigNB + g) * ld + q;
        const double* b1p = b0p + (size_t)8 * ld;
#pragma unroll 4
        for (int k0 = 0; k0 < kend; k0 += 4) {
          const double a0 = a0p[k0], a1 = a1p[k0], b0 = b0p[k0], b1 = b1p[k0];
          dmma884(acc[0][0], acc[0][1], a0, b0);
          dmma884(acc[1][0], acc[1][1], a0, b1);
          dmma884(acc[2][0], acc[2][1], a1, b0);
          dmma884(acc[3][0], acc[3][1], a1, b1);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int r = I * kBigNB + (t >> 1) * 8 + g, c = J * kBigNB + (t & 1) * 8 + 2 * q;
          double e01[2];
#pragma unroll 1
          for (int u = 0; u < 2; ++u) e01[u] = entry(r, c + u);     // one covariance chain at a time: registers
          const double v0 = e01[0] - acc[t][0], v1 = e01[1] - acc[t][1];
          if (I == J) {
            double* d = diag + ((t >> 1) * 8 + g) * kBigDiagLd + (t & 1) * 8 + 2 * q;
            d[0] = v0;
            d[1] = v1;
          } else if (r <= np) {
            *reinterpret_cast<double2*>(W + (size_t)r * ld + c) = make_double2(v0, v1);
          }
        }
        // ---- 2. the diagonal block: factor in shared memory (warp 0 owns I == J)
        if (I == J) {
          __syncwarp();
          for (int k = 0; k < kBigNB; ++k) {
            const double dkk = diag[k * kBigDiagLd + k];
            const double inv = pivot_rsqrt(dkk);             // NaN when not positive definite (as jax)
            double lrk = 0.0;
            if (lane < kBigNB && lane >= k) {
              lrk = diag[lane * kBigDiagLd + k] * inv;       // lane == k: sqrt(dkk)
              diag[lane * kBigDiagLd + k] = lrk;
              if (lane == k) invd[k] = inv;
            }
            __syncwarp();
            if (lane < kBigNB && lane > k)
              for (int c = k + 1; c <= lane; ++c)
                diag[lane * kBigDiagLd + c] = fma(-lrk, diag[c * kBigDiagLd + k], diag[lane * kBigDiagLd + c]);
            __syncwarp();
          }
          if (lane < kBigNB) {
            logdet += log(diag[lane * kBigDiagLd + lane]);
            if (PRED)                                        // the back substitution reads L_JJ from the workspace
              for (int c = 0; c <= lane; ++c) W[(size_t)(J * kBigNB + lane) * ld + J * kBigNB + c] = diag[lane * kBigDiagLd + c];
          }
        }
      }
      __syncthreads();
      // ---- 3. rows below the diagonal block: X L_JJ^T = P, one thread per row (incl. the residual row np)
      for (int r = (J + 1) * kBigNB + tid; r <= np; r += kBigThreads) {
        double* pr = W + (size_t)r * ld + J * kBigNB;
        double x[kBigNB];
#pragma unroll
        for (int c = 0; c < kBigNB; c += 2) {
          const double2 v = *reinterpret_cast<const double2*>(pr + c);
          x[c] = v.x;
          x[c + 1] = v.y;
        }
#pragma unroll
        for (int c = 0; c < kBigNB; ++c) {
          double sacc = x[c];
#pragma unroll
          for (int k = 0; k < c; ++k) sacc = fma(-x[k], diag[c * kBigDiagLd + k], sacc);
          x[c] = sacc * invd[c];
        }
#pragma unroll
        for (int c = 0; c < kBigNB; c += 2) *reinterpret_cast<double2*>(pr + c) = make_double2(x[c], x[c + 1]);
      }
      __syncthreads();
    }

    // alpha = row np of the workspace; chi^2 = alpha . alpha in a fixed order
    const double* alpha = W + (size_t)np * ld;
    double qs = 0.0;
    for (int i = tid; i < np; i += kBigThreads) qs = fma(alpha[i], alpha[i], qs);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qs += __shfl_xor_sync(0xffffffffu, qs, o);
    if (warp == 0) {
      double ld_sum = lane < kBigNB ? logdet : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ld_sum += __shfl_xor_sync(0xffffffffu, ld_sum, o);
      if (lane == 0) red[kBigWarps] = ld_sum;
    }
    if (lane == 0) red[warp] = qs;
    __syncthreads();
    double quad = 0.0;
    for (int w = 0; w < kBigWarps; ++w) quad += red[w];
    if (!PRED) {
      if (tid == 0) {
        const double ll = -0.5 * quad - red[kBigWarps] - 0.5 * (double)N * kLog2Pi;
        double r = ll + lp + lhp;                            // fit.py:7898-7900
        r += P.jacobian;
        r += P.renorm;
        out[s] = r;
      }
      continue;
    }
    // ---- PRED: beta = L^-T alpha by blocked back substitution (fit.py:7536-7554 via gp.condition)
    if (tid == 0 && out) out[s] = quad;                      // fit.py:5428-5429
    double* z = reinterpret_cast<double*>(smem + G.off_z);
    for (int i = tid; i < np; i += kBigThreads) z[i] = alpha[i];
    __syncthreads();
    for (int J = NBk - 1; J >= 0; --J) {
      if (warp == 0) {
        // L_JJ^T b = z_J: columns right to left, lane c owns z_c of the block
        const double* Ljj = W + (size_t)(J * kBigNB) * ld + J * kBigNB;
        double zc = lane < kBigNB ? z[J * kBigNB + lane] : 0.0;
        for (int c = kBigNB - 1; c >= 0; --c) {
          const double bc = __shfl_sync(0xffffffffu, zc, c) / Ljj[(size_t)c * ld + c];
          if (lane == c) zc = bc;
          else if (lane < c) zc = fma(-Ljj[(size_t)c * ld + lane], bc, zc);
        }
        if (lane < kBigNB) {
          z[J * kBigNB + lane] = zc;
          if (J * kBigNB + lane < N) beta_out[s * N + J * kBigNB + lane] = zc;
        }
      }
      __syncthreads();
      // z_K -= L_JK^T beta_J for every column left of the block: one thread per column, rows of the block in order
      for (int c = tid; c < J * kBigNB; c += kBigThreads) {
        double zc = z[c];
#pragma unroll
        for (int r = 0; r < kBigNB; ++r) zc = fma(-W[(size_t)(J * kBigNB + r) * ld + c], z[J * kBigNB + r], zc);
        z[c] = zc;
      }
      __syncthreads();
    }
  }
}

}  // namespace rvlp
