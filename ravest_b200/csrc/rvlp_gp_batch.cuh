// rvlp_gp_batch.cuh — K3 / K7-solve for LARGE BATCHES: level-synchronous batched Cholesky on the tensor cores.
//
// The quasi-periodic GP log-posterior (GPLogPosterior.log_probability, /root/reference/src/ravest/fit.py:7836-7901,
// 8062-8105; kernel gp.py:145-156) and the solve half of the conditioning path (fit.py:7494-7554, 5386-5429):
//   C = K(t, t) + diag(sigma^2 + jit^2) = L L^T,  alpha = L^-1 r,  ll = -1/2 alpha.alpha - sum ln L_ii - N/2 ln 2 pi.
//
// One sample's factor is 58 KB at N = 120 and 263 KB at N = 256: an SM holds two or three of them at most, so a
// one-CTA-per-sample kernel (rvlp_gp_pipe.cuh, rvlp_gp_big.cuh) is bound by the barriers and latencies of ITS sample's
// dependency chain.  With thousands of samples in the batch the parallelism is better taken ACROSS samples:
//   * every sample's factor lives in HBM / L2 as 16 x 16 blocks (2 KB contiguous each, lower block triangle, the
//     residual r riding along as row N so that alpha = L^-1 r falls out of the same sweeps);
//   * ONE KERNEL PER BLOCK COLUMN J: one warp per (sample, row block I > J) computes the tile
//         X_IJ = (C_IJ - sum_{K<J} L_IK L_JK^T) L_JJ^-T
//     with the sum on the tensor cores (`mma.sync.aligned.m8n8k4.f64`, SASS DMMA.884, operands loaded from the blocks in
//     fragment layout), C generated on the fly by the branch-free covariance function, and the 16 x 16 triangular solve
//     by one lane per row from shared memory.  The warp that owns I = J + 1 then factors the NEXT diagonal block (its
//     Gram sum rides along in the same operand stream: L_IK L_IK^T costs four more DMMA per k-step and no loads), so
//     there is no separate diagonal kernel and no barrier inside a sample at all - the kernel boundary is the only
//     synchronisation, and thousands of independent tiles fill the machine between two boundaries;
//   * chi^2 = alpha.alpha and sum ln L_kk are accumulated per sample in block-column order by one lane each (fixed
//     order: a sample's bits depend on its own row only, not on the batch, the grid or the chunking).
// HBM traffic is ~N^3 / 12 bytes per sample (0.15 MB at N = 120, 1.4 MB at N = 256) against N^3 / 3 FLOP: at 7 TB/s
// the factorisation is tensor-core / fp64-issue bound from N ~ 100 on.  Launches per call: N / 16 + 3.
#pragma once
#include "rvlp_gp.cuh"

namespace rvlp {

constexpr int kGbNB = 16;           // block size
constexpr int kGbWarps = 4;         // warps per CTA of the step kernel = row blocks of one sample per task group
constexpr int kGbThreads = 32 * kGbWarps;
constexpr int kGbLd = 17;           // padded leading dimension of shared-memory tiles
constexpr int kGbLdT = 16;          // leading dimension of the per-warp L_JJ copy (read as broadcasts)
#ifndef RVLP_GPB_ABLATE           // timing experiments only (wrong results): 1 no covariance chains, 2 no triangular solve,
#define RVLP_GPB_ABLATE 0         // 4 no Gram loop, 8 no diagonal-block factorisation
#endif
#ifndef RVLP_GPB_TUNE             // 1 fix-up-free covariance path for interior tiles, 4 clamp-only exp (both +1 % at
#define RVLP_GPB_TUNE 5           // N = 120, profiles/r02s_gpb_tune.log; a transposed L_JJ for the solve lost 3 %)
#endif
#ifndef RVLP_GPB_MB
#define RVLP_GPB_MB 4               // CTAs per SM the step kernel is compiled for (register cap = 65536 / (128 MB))
#endif

// Block geometry for N epochs: the residual is row N of the augmented matrix, rows N+1 .. 16 nbr - 1 are identity padding.
struct GpbDims {
  int N, IR, rr, nbr, nblk, np;
};
__host__ __device__ inline GpbDims gpb_dims(int N) {
  GpbDims d;
  d.N = N;
  d.IR = N / kGbNB;                 // block row of the residual row
  d.rr = N % kGbNB;                 // its row inside that block
  d.nbr = d.IR + 1;                 // block rows = block columns
  d.nblk = d.nbr * (d.nbr + 1) / 2;
  d.np = d.nbr * kGbNB;
  return d;
}
// bytes of workspace per sample (all arrays of GpbWork)
__host__ __device__ inline size_t gpb_bytes_per_sample(int N, int n_inst) {
  const GpbDims d = gpb_dims(N);
  return ((size_t)d.nblk * 256 + 4 * (size_t)d.np + 4 + (size_t)n_inst + 4) * 8 + 8;
}

struct GpbWork {
  double* L;        // [S][nblk][16][16]   block (I, K) at I (I + 1) / 2 + K, row-major inside
  double* invd;     // [S][np]             1 / L_kk
  double* resid;    // [S][np]             r = v - mean (pads 0)
  double* cph;      // [S][np]             cos / sin of the epochs' phases 2 pi (t_i - t_0) / P_gp (factored periodic term)
  double* sph;
  double* hyp;      // [S][4]              inv_P, inv_le, gamma, A^2
  double* jit2;     // [S][n_inst]
  double* lp;       // [S]  log-prior, [S] log-hyperprior, [S] chi^2, [S] sum ln L_kk
  double* lhp;
  double* chi2;
  double* logdet;
  int* status;      // [S]  0 factorise, 1 rejected (-inf; PRED: NaN row), 2 non-finite mean model (-inf + lp + lhp)
};

__device__ __forceinline__ double* gpb_block(const GpbWork& w, const GpbDims& d, int64_t s, int I, int K) {
  return w.L + ((size_t)s * d.nblk + (size_t)(I * (I + 1) / 2 + K)) * 256;
}

// Per-sample view of what the covariance entries need.
struct GpbSample {
  GpHyper hyp;
  const double* t;
  const double* e2;
  const int* inst;
  const double* jit2;
  const double* resid;
  const double* cph;
  const double* sph;
  int N;
};

// Entries (i, j) and (i, j + 1) of the augmented, padded matrix, for two rows at once: four independent covariance
// chains (branch-free) followed by the fix-ups - white-noise diagonal (fit.py:8094-8096), residual row N, identity
// padding below it.  Only the lower triangle is ever used.
// INTERIOR (compile time): the tile lies strictly below the diagonal and above row N - no fix-ups, no clamps.
template <bool INTERIOR>
__device__ __forceinline__ void gpb_entries4(const GpbSample& sm, int i0, int i1, int j, double (&v)[4]) {
  const int N = sm.N;
  const int ii[2] = {i0, i1};
  double tv[2], tj[2], ci[2], si[2], cj[2], sj[2];
  int ic[2], jc[2];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    ic[a] = (INTERIOR || ii[a] < N) ? ii[a] : N - 1;
    jc[a] = (INTERIOR || j + a < N) ? j + a : N - 1;
    tv[a] = sm.t[ic[a]];
    tj[a] = sm.t[jc[a]];
    ci[a] = sm.cph[ic[a]]; si[a] = sm.sph[ic[a]];
    cj[a] = sm.cph[jc[a]]; sj[a] = sm.sph[jc[a]];
  }
  // gp.py:145-156 with the periodic factor split as in the pipelined kernels (rvlp_gp_pipe.cuh):
  //   sin^2(pi (t_i - t_j) / P) = (1 - cos(b_i - b_j)) / 2,  cos(b_i - b_j) = cos b_i cos b_j + sin b_i sin b_j,
  // 19 instead of 34 fp64 instructions per element and no table sine.
  const double g2 = 0.5 * sm.hyp.gamma;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const double cd = fma(ci[a], cj[b], si[a] * sj[b]);
      const double q = (tv[a] - tj[b]) * sm.hyp.inv_le;
      v[a * 2 + b] = (RVLP_GPB_ABLATE & 1) ? cd + q : ((RVLP_GPB_TUNE & 4) ? gp_exp_scaled_neg(fma(g2, cd, fma(-0.5 * q, q, -g2)), sm.hyp.A2) : gp_exp_scaled(fma(g2, cd, fma(-0.5 * q, q, -g2)), sm.hyp.A2));
    }
  if (INTERIOR) return;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = ii[a], jj = j + b;
      double x = v[a * 2 + b];
      if (i == jj && i < N) x += sm.e2[ic[a]] + sm.jit2[sm.inst[ic[a]]];                  // fit.py:8094-8096
      if (i == N) x = jj < N ? sm.resid[jc[b]] : 1.0e300;                                 // the residual row; (N, N) is unused
      if (i > N || jj > N) x = i == jj ? 1.0 : 0.0;                                       // identity padding
      if (i < N && jj == N) x = 0.0;
      v[a * 2 + b] = x;
    }
}

// Cholesky of a 16 x 16 block held one ROW PER LANE in registers (lanes 0..15; a[c], c <= lane, is the lower triangle).
// Column k: every lane fetches the pivot from lane k, forms 1 / sqrt(pivot) redundantly (no broadcast afterwards),
// scales its entry, and the rank-1 update pulls l_ck from lane c by shuffle.  No shared memory, no __syncwarp:
// ~2 k cycles instead of the ~12 k of a shared-memory version whose rows cannot be kept in registers.
// Returns 1 / L_kk of the lane's own row in `inv_own`.  Not positive definite -> NaN (as jax).
__device__ __forceinline__ void gpb_factor_rows(double (&a)[kGbNB], double& inv_own, int lane) {
  inv_own = 0.0;
#pragma unroll
  for (int k = 0; k < kGbNB; ++k) {
    const double piv = __shfl_sync(0xffffffffu, a[k], k);
    const double inv = pivot_rsqrt(piv);
    const double lrk = a[k] * inv;                       // lane k: sqrt(piv); lanes < k: unused
    a[k] = lrk;
    if (lane == k) inv_own = inv;
#pragma unroll
    for (int c = k + 1; c < kGbNB; ++c) {
      const double lck = __shfl_sync(0xffffffffu, lrk, c);
      a[c] = fma(-lrk, lck, a[c]);                       // only c <= lane is meaningful; the rest is never read
    }
  }
}

// After the diagonal block Jd of sample s has been factored (row per lane): publish it, and account for ln L_kk / the
// residual row's entries in it.
__device__ __forceinline__ void gpb_publish_rows(const GpbWork& w, const GpbDims& d, int64_t s, int Jd,
                                                 const double (&a)[kGbNB], double inv_own, int lane) {
  double lg = 0.0;
  if (lane < kGbNB) {
    double* dst = gpb_block(w, d, s, Jd, Jd) + lane * kGbNB;
#pragma unroll
    for (int c = 0; c < kGbNB; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(a[c], a[c + 1]);
    w.invd[(size_t)s * d.np + Jd * kGbNB + lane] = inv_own;
    double own = 1.0;
#pragma unroll
    for (int c = 0; c < kGbNB; ++c) own = c == lane ? a[c] : own;
    if (Jd * kGbNB + lane < d.N) lg = log(own);
    if (Jd == d.IR && lane == d.rr) {                    // alpha's last entries sit in this block's row rr
      double q = 0.0;
#pragma unroll
      for (int c = 0; c < kGbNB; ++c) q = c < d.rr ? fma(a[c], a[c], q) : q;
      w.chi2[s] += q;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
  if (lane == 0) w.logdet[s] += lg;
}

// ------------------------------------------------------------------ prologue: records, residual, hyperparameters
// One warp per sample: priors / conversions / reject flags (the K4 prologue), the mean model at the N epochs (Kepler
// stage, two epochs per lane in flight) and the per-sample constants of the covariance function.
template <bool PRED>
__global__ void __launch_bounds__(kThreads)
gpb_prologue_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, GpbWork w) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);              // P.epochs_global == 1 in the copy this kernel receives
  stage_problem(P, L, smem);
  const Tables T = tables_of<true>(P, L, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  const int cap = P.batch_cap > kG ? P.batch_cap : kG;
  double* sr = reinterpret_cast<double*>(smem + L.off_scratch) + warp * cap * rec;
  double* pv = reinterpret_cast<double*>(smem + L.off_pv) + warp * cap * P.n_priors;
  const GpbDims d = gpb_dims(P.n_epochs);
  const int N = d.N;
  const int64_t gw = (int64_t)blockIdx.x * kWarps + warp, nw = (int64_t)gridDim.x * kWarps;
  for (int64_t s = gw; s < S; s += nw) {
    sample_prologue(P, T, theta, s, s + 1, sr, rec, lane, !PRED, 1, pv);
    const int flags = __double2loint(sr[1]);
    int st = 0;
    if (PRED ? (flags & (F_PLANET | F_HYPER)) : (flags & (F_JIT | F_HYPER | F_PRIOR))) {
      st = 1;                                                // fit.py:7857-7886 (PRED: the reference raises)
    } else {
      int nonfinite = (!PRED && (flags & F_PLANET)) ? 1 : 0; // fit.py:8022-8024
      double* resid = w.resid + (size_t)s * d.np;
      if (!nonfinite) {
        constexpr int kRW = 2;
#pragma unroll 1
        for (int base = 0; base < d.np; base += 32 * kRW) {  // residual, fit.py:7994-8043, 8059 / 7543-7550
          double tt[kRW], rv[kRW];
          int idx[kRW];
#pragma unroll
          for (int j = 0; j < kRW; ++j) {
            idx[j] = base + j * 32 + lane;
            tt[j] = T.t[idx[j] < N ? idx[j] : N - 1];
          }
          model_rv<kRW>(P, sr, tt, rv, -1, true);
#pragma unroll
          for (int j = 0; j < kRW; ++j) {
            if (idx[j] < N) {
              if (PRED) {
                resid[idx[j]] = (T.v[idx[j]] - sr[kHdr + T.inst[idx[j]]]) - rv[j];
              } else {
                const double mean = rv[j] + sr[kHdr + T.inst[idx[j]]];
                if (!(fabs(mean) <= 1.79769313486231570e308)) nonfinite = 1;
                resid[idx[j]] = T.v[idx[j]] - mean;
              }
            } else if (idx[j] < d.np) {
              resid[idx[j]] = 0.0;
            }
          }
        }
      }
      if (__any_sync(0xffffffffu, nonfinite)) st = 2;        // fit.py:8082-8083
    }
    if (st == 0) {                                            // phases of the epochs for the factored periodic term
      const double inv_P = 1.0 / model_param(T, theta + s * P.ndim, P.n_model + 3);
      const double t_ref = T.t[0];
      double* cp = w.cph + (size_t)s * d.np;
      double* sp = w.sph + (size_t)s * d.np;
      for (int j = lane; j < N; j += 32) {
        double sj, cj;
        sincospi(2.0 * ((T.t[j] - t_ref) * inv_P), &sj, &cj);
        cp[j] = cj;
        sp[j] = sj;
      }
    }
    if (lane == 0) {
      const double* row = theta + s * P.ndim;
      const GpHyper h = gp_hyper(model_param(T, row, P.n_model + 0), model_param(T, row, P.n_model + 1),
                                 model_param(T, row, P.n_model + 2), model_param(T, row, P.n_model + 3));
      double* hp = w.hyp + (size_t)s * 4;
      hp[0] = h.inv_P; hp[1] = h.inv_le; hp[2] = h.gamma; hp[3] = h.A2;
      for (int j = 0; j < P.n_inst; ++j) w.jit2[(size_t)s * P.n_inst + j] = sr[kHdr + P.n_inst + j];
      w.lp[s] = sr[0];
      w.lhp[s] = sr[4];
      w.chi2[s] = 0.0;
      w.logdet[s] = 0.0;
      w.status[s] = st;
    }
    __syncwarp();
  }
}

__device__ __forceinline__ GpbSample gpb_sample(const DevProblem& P, const GpbWork& w, const GpbDims& d, int64_t s) {
  GpbSample sm;
  const double* hp = w.hyp + (size_t)s * 4;
  sm.hyp.inv_P = hp[0]; sm.hyp.inv_le = hp[1]; sm.hyp.gamma = hp[2]; sm.hyp.A2 = hp[3];
  sm.t = P.epochs;
  sm.e2 = P.epochs + 2 * (size_t)P.n_pad;
  sm.inst = reinterpret_cast<const int*>(P.epochs + 3 * (size_t)P.n_pad);
  sm.jit2 = w.jit2 + (size_t)s * P.n_inst;
  sm.resid = w.resid + (size_t)s * d.np;
  sm.cph = w.cph + (size_t)s * d.np;
  sm.sph = w.sph + (size_t)s * d.np;
  sm.N = d.N;
  return sm;
}

// C tile (I, J) minus the accumulated Gram sums, from the DMMA accumulator layout into a shared-memory tile.
template <bool INTERIOR>
__device__ __forceinline__ void gpb_tile_to_smem_t(const GpbSample& sm, int I, int J, const double (&acc)[4][2], double* tile,
                                                   int g, int q) {
#pragma unroll
  for (int cb = 0; cb < 2; ++cb) {                          // column half: tiles (0, cb) and (1, cb) share the columns
    double v[4];
    gpb_entries4<INTERIOR>(sm, I * kGbNB + g, I * kGbNB + 8 + g, J * kGbNB + cb * 8 + 2 * q, v);
    tile[g * kGbLd + cb * 8 + 2 * q] = v[0] - acc[cb][0];
    tile[g * kGbLd + cb * 8 + 2 * q + 1] = v[1] - acc[cb][1];
    tile[(8 + g) * kGbLd + cb * 8 + 2 * q] = v[2] - acc[2 + cb][0];
    tile[(8 + g) * kGbLd + cb * 8 + 2 * q + 1] = v[3] - acc[2 + cb][1];
  }
}
__device__ __forceinline__ void gpb_tile_to_smem(const GpbSample& sm, int I, int J, const double (&acc)[4][2], double* tile,
                                                 int g, int q) {
  if ((RVLP_GPB_TUNE & 1) && I > J && I * kGbNB + kGbNB - 1 < sm.N) gpb_tile_to_smem_t<true>(sm, I, J, acc, tile, g, q);   // warp-uniform
  else gpb_tile_to_smem_t<false>(sm, I, J, acc, tile, g, q);
}

// ------------------------------------------------------------------ diagonal block 0
__global__ void __launch_bounds__(kGbThreads)
gpb_diag0_kernel(DevProblem P, int64_t S, GpbWork w) {
  __shared__ double tiles[kGbWarps][kGbNB * kGbLd];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  const GpbDims d = gpb_dims(P.n_epochs);
  double* tile = tiles[warp];
  for (int64_t s = (int64_t)blockIdx.x * kGbWarps + warp; s < S; s += (int64_t)gridDim.x * kGbWarps) {
    if (w.status[s] != 0) continue;
    const GpbSample sm = gpb_sample(P, w, d, s);
    const double acc[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
    gpb_tile_to_smem(sm, 0, 0, acc, tile, g, q);
    __syncwarp();
    double a[kGbNB], inv_own;
    const int rl = lane & (kGbNB - 1);
#pragma unroll
    for (int c = 0; c < kGbNB; ++c) a[c] = tile[rl * kGbLd + c];
    gpb_factor_rows(a, inv_own, lane);
    gpb_publish_rows(w, d, s, 0, a, inv_own, lane);
    __syncwarp();
  }
}

// ------------------------------------------------------------------ block column J
// One WARP per (sample, group of TPW consecutive row blocks I > J); warps never synchronise with each other.
// TPW = 2 (two 16 x 16 tiles per warp): the B operand (block row J) is loaded once for both, all 32 lanes work in the
// triangular solve (one row each) and the per-task fixed costs are halved - faster from ~300 epochs on (N = 1024:
// 16.7 -> 15.2 ms per 600 samples); TPW = 1 keeps the registers and twice the tasks - faster below (N = 120: 1.34 vs
// 1.40 ms per 1e4).  Consecutive warps take consecutive groups of the same sample, so the shared operands hit in L1.
template <int TPW>
__global__ void __launch_bounds__(kGbThreads, RVLP_GPB_MB)
gpb_step_kernel(DevProblem P, int64_t S, GpbWork w, int J) {
  __shared__ __align__(16) double ljj_s[kGbWarps][kGbNB * kGbLdT];
  __shared__ double invd_s[kGbWarps][kGbNB];
  __shared__ double tiles[kGbWarps][TPW][kGbNB * kGbLd];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const GpbDims d = gpb_dims(P.n_epochs);
  const int nI = d.nbr - 1 - J;                                  // row blocks below the diagonal block
  const int nP = (nI + TPW - 1) / TPW;                           // groups of them
  const int64_t total = S * (int64_t)nP;
  double* ljj = ljj_s[warp];
  double* invd_j = invd_s[warp];
  const int rl = lane & (kGbNB - 1), tsel = TPW == 2 ? lane >> 4 : 0;   // solve: row rl of tile tsel
  for (int64_t task = (int64_t)blockIdx.x * kGbWarps + warp; task < total; task += (int64_t)gridDim.x * kGbWarps) {
    const int64_t s = task / nP;
    const int I0 = J + 1 + TPW * (int)(task - s * nP);
    if (w.status[s] != 0) continue;
    const bool two = TPW == 2 && I0 + 1 < d.nbr;                 // the last group of an odd count has one tile
    const int I1 = two ? I0 + 1 : I0;                            // (the second tile then repeats the first and stores nothing)
    const bool next_diag = I0 == J + 1;
    // L_JJ and 1 / diag: the loads are issued here and parked in registers; they go to shared memory after the Gram
    // loop, whose operand loads they overlap with
    double2 ljj_r[4];
    {
      const double2* src = reinterpret_cast<const double2*>(gpb_block(w, d, s, J, J));
#pragma unroll
      for (int e = 0; e < 4; ++e) ljj_r[e] = src[e * 32 + lane];   // 128 double2 = the 2 KB block, coalesced
    }
    const double invd_r = w.invd[(size_t)s * d.np + J * kGbNB + rl];
    const GpbSample sm = gpb_sample(P, w, d, s);
    // ---- sum_{K<J} L_IK L_JK^T per tile (and L_I0K L_I0K^T for the next diagonal block) on the tensor cores
    double acc[TPW][4][2] = {};
    double gram[4][2] = {};
    const double* arow[TPW];
    arow[0] = gpb_block(w, d, s, I0, 0) + g * kGbNB + q;          // blocks (I, 0..J-1) are contiguous
    if (TPW == 2) arow[TPW - 1] = gpb_block(w, d, s, I1, 0) + g * kGbNB + q;
    const double* brow = gpb_block(w, d, s, J, 0) + g * kGbNB + q;
    const int Jend = (RVLP_GPB_ABLATE & 4) ? 0 : J;
    if (next_diag) {
#pragma unroll 1
      for (int K = 0; K < Jend; ++K) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int o = K * 256 + 4 * kk;
          const double b0 = brow[o], b1 = brow[o + 128];
          double a0[TPW], a1[TPW];
#pragma unroll
          for (int t = 0; t < TPW; ++t) { a0[t] = arow[t][o]; a1[t] = arow[t][o + 128]; }
#pragma unroll
          for (int t = 0; t < TPW; ++t) {
            dmma884(acc[t][0][0], acc[t][0][1], a0[t], b0);
            dmma884(acc[t][1][0], acc[t][1][1], a0[t], b1);
            dmma884(acc[t][2][0], acc[t][2][1], a1[t], b0);
            dmma884(acc[t][3][0], acc[t][3][1], a1[t], b1);
          }
          dmma884(gram[0][0], gram[0][1], a0[0], a0[0]);
          dmma884(gram[1][0], gram[1][1], a0[0], a1[0]);
          dmma884(gram[2][0], gram[2][1], a1[0], a0[0]);
          dmma884(gram[3][0], gram[3][1], a1[0], a1[0]);
        }
      }
    } else {
#pragma unroll 1
      for (int K = 0; K < Jend; ++K) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int o = K * 256 + 4 * kk;
          const double b0 = brow[o], b1 = brow[o + 128];
          double a0[TPW], a1[TPW];
#pragma unroll
          for (int t = 0; t < TPW; ++t) { a0[t] = arow[t][o]; a1[t] = arow[t][o + 128]; }
#pragma unroll
          for (int t = 0; t < TPW; ++t) {
            dmma884(acc[t][0][0], acc[t][0][1], a0[t], b0);
            dmma884(acc[t][1][0], acc[t][1][1], a0[t], b1);
            dmma884(acc[t][2][0], acc[t][2][1], a1[t], b0);
            dmma884(acc[t][3][0], acc[t][3][1], a1[t], b1);
          }
        }
      }
    }
    __syncwarp();                                                // the previous task is done with this warp's buffers
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = e * 32 + lane;
      *reinterpret_cast<double2*>(ljj + (idx >> 3) * kGbLdT + 2 * (idx & 7)) = ljj_r[e];
    }
    if (lane < kGbNB) invd_j[lane] = invd_r;
    gpb_tile_to_smem(sm, I0, J, acc[0], tiles[warp][0], g, q);
    if (TPW == 2) gpb_tile_to_smem(sm, I1, J, acc[TPW - 1], tiles[warp][TPW - 1], g, q);
    __syncwarp();
    // ---- X L_JJ^T = P: one lane per row (TPW = 1: lanes 16..31 shadow rows 0..15 and store nothing).  Column-oriented:
    // once x_k is final, the updates of x_{k+1..15} are independent of each other (a 32-deep dependency chain, not 136).
    double x[kGbNB];
    {
      const double* trow = tiles[warp][tsel] + rl * kGbLd;
#pragma unroll
      for (int c = 0; c < kGbNB; ++c) x[c] = trow[c];
    }
#pragma unroll
    for (int k = 0; k < ((RVLP_GPB_ABLATE & 2) ? 1 : kGbNB); ++k) {
      x[k] *= invd_j[k];
#pragma unroll
      for (int c = k + 1; c < kGbNB; ++c) x[c] = fma(-x[k], ljj[c * kGbLdT + k], x[c]);
    }
    if (TPW == 2 ? (tsel == 0 || two) : lane < kGbNB) {
      const int It = tsel ? I1 : I0;
      double* dst = gpb_block(w, d, s, It, J) + rl * kGbNB;
#pragma unroll
      for (int c = 0; c < kGbNB; c += 2) *reinterpret_cast<double2*>(dst + c) = make_double2(x[c], x[c + 1]);
      if (It == d.IR && rl == d.rr) {                             // alpha's entries of block column J (all < N)
        double qs = 0.0;
#pragma unroll
        for (int c = 0; c < kGbNB; ++c) qs = fma(x[c], x[c], qs);
        w.chi2[s] += qs;
      }
    }
    // ---- the next diagonal block: Gram sum over blocks 0..J (the last one from the tile just solved), factor, publish
    if (next_diag && !(RVLP_GPB_ABLATE & 8)) {
      double* tile = tiles[warp][0];
      __syncwarp();
      if (TPW == 2 ? tsel == 0 : lane < kGbNB) {
#pragma unroll
        for (int c = 0; c < kGbNB; ++c) tile[rl * kGbLd + c] = x[c];
      }
      __syncwarp();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const double a0 = tile[g * kGbLd + 4 * kk + q], a1 = tile[(8 + g) * kGbLd + 4 * kk + q];
        dmma884(gram[0][0], gram[0][1], a0, a0);
        dmma884(gram[1][0], gram[1][1], a0, a1);
        dmma884(gram[2][0], gram[2][1], a1, a0);
        dmma884(gram[3][0], gram[3][1], a1, a1);
      }
      __syncwarp();
      gpb_tile_to_smem(sm, I0, I0, gram, tile, g, q);
      __syncwarp();
      double inv_own;
#pragma unroll
      for (int c = 0; c < kGbNB; ++c) x[c] = tile[rl * kGbLd + c];
      gpb_factor_rows(x, inv_own, lane);
      gpb_publish_rows(w, d, s, I0, x, inv_own, lane);
    }
  }
}

// ------------------------------------------------------------------ finish
template <bool PRED>
__global__ void gpb_finish_kernel(DevProblem P, int64_t S, GpbWork w, double* __restrict__ out, double* __restrict__ beta_out) {
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  const int N = P.n_epochs;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
    const int st = w.status[s];
    if (PRED) {
      if (st != 0) {                                             // the reference raises: NaN rows
        for (int j = 0; j < N; ++j) beta_out[s * N + j] = qnan;
        if (out) out[s] = qnan;
      } else if (out) {
        out[s] = w.chi2[s];                                      // fit.py:5428-5429
      }
      continue;
    }
    double r;
    if (st == 1) {
      r = -INFINITY;                                             // fit.py:7857-7886
    } else {
      const double ll = st == 2 ? -INFINITY                      // fit.py:8082-8083
                                : -0.5 * w.chi2[s] - w.logdet[s] - 0.5 * (double)N * kLog2Pi;
      r = ll + w.lp[s] + w.lhp[s];                               // fit.py:7898-7900
      r += P.jacobian;
      r += P.renorm;
    }
    out[s] = r;
  }
}

// ------------------------------------------------------------------ PRED: beta = L^-T alpha
// One CTA per sample, blocked back substitution over the leading N x N factor (rows >= N are the residual / padding).
__global__ void __launch_bounds__(kGbThreads)
gpb_backsub_kernel(DevProblem P, int64_t S, GpbWork w, double* __restrict__ beta_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* z = reinterpret_cast<double*>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const GpbDims d = gpb_dims(P.n_epochs);
  const int N = d.N;
  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    if (w.status[s] != 0) continue;
    __syncthreads();
    for (int c = tid; c < N; c += kGbThreads) z[c] = gpb_block(w, d, s, d.IR, c / kGbNB)[d.rr * kGbNB + (c % kGbNB)];
    __syncthreads();
    for (int J = d.IR; J >= 0; --J) {
      const int nrow = J == d.IR ? d.rr : kGbNB;                  // matrix rows in this block row
      if (nrow == 0) continue;
      if (warp == 0) {
        const double* Ljj = gpb_block(w, d, s, J, J);
        double zc = lane < nrow ? z[J * kGbNB + lane] : 0.0;
        for (int c = nrow - 1; c >= 0; --c) {
          const double bc = __shfl_sync(0xffffffffu, zc, c) / Ljj[c * kGbNB + c];
          if (lane == c) zc = bc;
          else if (lane < c) zc = fma(-Ljj[c * kGbNB + lane], bc, zc);
        }
        if (lane < nrow) {
          z[J * kGbNB + lane] = zc;
          beta_out[s * N + J * kGbNB + lane] = zc;
        }
      }
      __syncthreads();
      for (int c = tid; c < J * kGbNB; c += kGbThreads) {         // z_K -= L_JK^T beta_J, one thread per column
        const double* blk = gpb_block(w, d, s, J, c / kGbNB) + (c % kGbNB);
        double zc = z[c];
        for (int r = 0; r < nrow; ++r) zc = fma(-blk[r * kGbNB], z[J * kGbNB + r], zc);
        z[c] = zc;
      }
      __syncthreads();
    }
  }
}

}  // namespace rvlp
