// rvlp_gp.cuh — K3: batched quasi-periodic GP log-posterior (config 5).
//
// Replaces GPLogPosterior.log_probability / GPLogLikelihood.__call__
// (/root/reference/src/ravest/fit.py:7836-7901, 8062-8105) and the kernel of gp.py:126-156.
// The dense algebra lives in tinygp 0.3.0 (absent, "parity unpinned"): restated as
//   C_ij = A^2 exp(-Gamma sin^2(pi |t_i - t_j| / P_gp)) exp(-(t_i - t_j)^2 / (2 lambda_e^2))
//          + delta_ij (sigma_i^2 + jit_i^2),            Gamma = 1 / (2 lambda_p^2)
//   C = L L^T,  alpha = L^-1 (v - mean),  ll = -1/2 alpha.alpha - sum log L_ii - N/2 log 2 pi.
//
// One CTA per sample.  The packed lower triangle of C (plus the residual as an extra row, so
// the factorisation's column sweeps produce alpha for free) sits in shared memory:
// (N+1)(N+2)/2 doubles = 59 KB at N = 120, three CTAs per SM.  Right-looking Cholesky with
// the trailing update spread over the CTA.  fp64 throughout.
#pragma once
#include "rvlp_kernels.cuh"
#include "rvlp_gpcov.cuh"

namespace rvlp {

__host__ __device__ inline int gp_tri_doubles(int N) { return (N + 1) * (N + 2) / 2; }

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
__device__ __forceinline__ int tri(int i, int j) { return i * (i + 1) / 2 + j; }

struct GpSmem {
  int off_tri, off_red, total;
};
__host__ __device__ inline GpSmem gp_smem(const DevProblem& P, const SmemLayout& L) {
  GpSmem G;
  int o = (L.total + 15) & ~15;
  G.off_tri = o; o += gp_tri_doubles(P.n_epochs) * 8;
  G.off_red = o; o += 64 * 8;
  G.total = o;
  return G;
}

__global__ void __launch_bounds__(kThreads)
gp_logprob_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);
  const GpSmem G = gp_smem(P, L);
  stage_problem(P, L, smem);
  const Tables T = tables_of(P, L, smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  double* scratch = reinterpret_cast<double*>(smem + L.off_scratch);   // warp 0's slot, sample g = 0
  double* Cm = reinterpret_cast<double*>(smem + G.off_tri);
  double* red = reinterpret_cast<double*>(smem + G.off_red);
  const int N = P.n_epochs;

  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    if (warp == 0) sample_prologue(P, T, theta, s, s + 1, scratch, rec, lane, true, 1, reinterpret_cast<double*>(smem + L.off_pv));
    __syncthreads();
    const double* sr = scratch;
    const int flags = __double2loint(sr[1]);
    const double lp = sr[0], lhp = sr[4];
    if (flags & (F_JIT | F_HYPER | F_PRIOR)) {               // fit.py:7857-7886
      if (tid == 0) out[s] = -INFINITY;
      __syncthreads();
      continue;
    }
    if (flags & F_PLANET) {                                  // fit.py:8022-8024, 8082-8083
      if (tid == 0) {
        double r = -INFINITY + lp + lhp;
        r += P.jacobian;
        r += P.renorm;
        out[s] = r;
      }
      __syncthreads();
      continue;
    }
    const double* row = theta + s * P.ndim;
    const double A = model_param(T, row, P.n_model + 0), le = model_param(T, row, P.n_model + 1);
    const double lpp = model_param(T, row, P.n_model + 2), Pg = model_param(T, row, P.n_model + 3);
    const GpHyper hyp = gp_hyper(A, le, lpp, Pg);            // gp.py:145-156

    // residual row (index N): v - (planets + trend + gamma_inst)   fit.py:7994-8043, 8059
    int nonfinite = 0;
    for (int i = tid; i < N; i += kThreads) {
      double tt[1] = {T.t[i]}, rv[1];
      model_rv<1>(P, sr, tt, rv, -1, true);
      const double mean = rv[0] + sr[kHdr + T.inst[i]];
      if (!(fabs(mean) <= 1.79769313486231570e308)) nonfinite = 1;
      Cm[tri(N, i)] = T.v[i] - mean;
    }
    // covariance, packed lower triangle                        gp.py:145-156, fit.py:8094-8096
    const int npairs = N * (N + 1) / 2;
    for (int p = tid; p < npairs; p += kThreads) {
      int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
      while (tri(i + 1, 0) <= p) ++i;
      while (tri(i, 0) > p) --i;
      const int j = p - tri(i, 0);
      double c = gp_cov(T.t[i] - T.t[j], hyp);
      if (i == j) c += T.e2[i] + sr[kHdr + P.n_inst + T.inst[i]];
      Cm[p] = c;
    }
    if (__syncthreads_or(nonfinite)) {                       // fit.py:8082-8083
      if (tid == 0) {
        double r = -INFINITY + lp + lhp;
        r += P.jacobian;
        r += P.renorm;
        out[s] = r;
      }
      __syncthreads();
      continue;
    }
    // right-looking Cholesky on rows 0..N (row N = residual -> alpha)
    double logdet_part = 0.0;
    for (int j = 0; j < N; ++j) {
      const double djj = sqrt(Cm[tri(j, j)]);                // NaN when not positive definite (as jax)
      const double inv = 1.0 / djj;
      if (tid == 0) logdet_part += log(djj);
      __syncthreads();                                       // everyone has read C_jj
      for (int i = j + 1 + tid; i <= N; i += kThreads) Cm[tri(i, j)] *= inv;
      if (tid == 0) Cm[tri(j, j)] = djj;
      __syncthreads();
      // trailing update: C_ik -= L_ij L_kj for j < k <= i <= N (diagonal of row N not needed)
      const int m = N - j;                                   // rows j+1..N
      const int cnt = m * (m + 1) / 2;
      for (int p = tid; p < cnt; p += kThreads) {
        int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= p) ++a;
        while (a * (a + 1) / 2 > p) --a;
        const int b = p - a * (a + 1) / 2;
        const int i = j + 1 + a, k = j + 1 + b;
        if (i == N && k == N) continue;
        Cm[tri(i, k)] = fma(-Cm[tri(i, j)], Cm[tri(k, j)], Cm[tri(i, k)]);
      }
      __syncthreads();
    }
    // quad = alpha . alpha
    double q = 0.0;
    for (int i = tid; i < N; i += kThreads) q = fma(Cm[tri(N, i)], Cm[tri(N, i)], q);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (tid == 0) {
      double quad = 0.0;
      for (int w = 0; w < kWarps; ++w) quad += red[w];
      const double ll = -0.5 * quad - logdet_part - 0.5 * (double)N * kLog2Pi;
      double r = ll + lp + lhp;                              // fit.py:7898-7900
      r += P.jacobian;
      r += P.renorm;
      out[s] = r;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ K3, register-tiled version
// The whole lower triangle lives in REGISTERS: thread (I, J), I >= J, owns the T x T tile of rows
// I*T.. and columns J*T..; row N is the residual (so the column sweeps forward-substitute it and
// alpha = L^-1 r falls out).  Per column j: the owners of column j publish it to a double-buffered
// shared-memory vector, ONE __syncthreads, then every thread scales by 1/sqrt(pivot) and applies the
// rank-1 update to its tile with T + T shared loads per T*T FMAs.  N <= 22*T - 1 (T = 8: N <= 175).
struct GpTiledSmem { int off_resid, off_col, total; };
__host__ __device__ inline GpTiledSmem gp_tiled_smem(const DevProblem& P, const SmemLayout& L) {
  GpTiledSmem G;
  int o = (L.total + 15) & ~15;
  const int rows = ((P.n_epochs + 1 + 16) + 1) & ~1;  // + padding rows of the last tile row; even: 16-byte rows
  G.off_resid = o; o += rows * 8;
  G.off_col = o; o += 2 * rows * 8;
  G.total = o;
  return G;
}
__host__ __device__ inline int gp_tile_for(int n_epochs) {
  const int t[5] = {2, 4, 6, 8, 10};       // 10: pipelined kernels only (rvlp_gp_pipe.cuh), N <= 219
  for (int i = 0; i < 5; ++i)
    if (n_epochs + 1 <= 22 * t[i]) return t[i];
  return 0;
}

template <int TT>
__global__ void __launch_bounds__(kThreads, (TT >= 8 ? 1 : 2))
gp_logprob_tiled_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);
  const GpTiledSmem G = gp_tiled_smem(P, L);
  stage_problem(P, L, smem);
  const Tables T = tables_of(P, L, smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  double* scratch = reinterpret_cast<double*>(smem + L.off_scratch);
  double* resid = reinterpret_cast<double*>(smem + G.off_resid);
  double* colbuf = reinterpret_cast<double*>(smem + G.off_col);
  const int N = P.n_epochs;
  const int rows = ((N + 1 + 16) + 1) & ~1;
  const int nt = (N + 1 + TT - 1) / TT;
  // Tile coordinates of this thread: lower triangle enumerated COLUMN-major, so the lanes of a warp
  // share (almost) one tile column J and therefore retire from the sweep together.
  int J = 0, rem = tid;
  while (J < nt && rem >= nt - J) { rem -= nt - J; ++J; }
  const int I = J + rem;
  const bool has_tile = J < nt;
  const int r0 = I * TT, c0 = J * TT;

  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    if (warp == 0) sample_prologue(P, T, theta, s, s + 1, scratch, rec, lane, true, 1, reinterpret_cast<double*>(smem + L.off_pv));
    __syncthreads();
    const double* sr = scratch;
    const int flags = __double2loint(sr[1]);
    const double lp = sr[0], lhp = sr[4];
    if (flags & (F_JIT | F_HYPER | F_PRIOR)) {               // fit.py:7857-7886
      if (tid == 0) out[s] = -INFINITY;
      __syncthreads();
      continue;
    }
    int nonfinite = (flags & F_PLANET) ? 1 : 0;              // fit.py:8022-8024
    if (!nonfinite) {
      for (int i = tid; i < N; i += kThreads) {              // residual v - mean, fit.py:7994-8043, 8059
        double tt[1] = {T.t[i]}, rv[1];
        model_rv<1>(P, sr, tt, rv, -1, true);
        const double mean = rv[0] + sr[kHdr + T.inst[i]];
        if (!(fabs(mean) <= 1.79769313486231570e308)) nonfinite = 1;
        resid[i] = T.v[i] - mean;
      }
    }
    if (__syncthreads_or(nonfinite)) {                       // fit.py:8082-8083
      if (tid == 0) {
        double r = -INFINITY + lp + lhp;
        r += P.jacobian;
        r += P.renorm;
        out[s] = r;
      }
      __syncthreads();
      continue;
    }
    const double* row = theta + s * P.ndim;
    const double Aamp = model_param(T, row, P.n_model + 0), le = model_param(T, row, P.n_model + 1);
    const double lpp = model_param(T, row, P.n_model + 2), Pg = model_param(T, row, P.n_model + 3);
    const GpHyper hyp = gp_hyper(Aamp, le, lpp, Pg);         // gp.py:145-156

    // build this thread's tile                               gp.py:145-156, fit.py:8094-8096
    double a[TT][TT];
#pragma unroll
    for (int r = 0; r < TT; ++r) {
      const int i = r0 + r;
#pragma unroll
      for (int c = 0; c < TT; ++c) {
        const int k = c0 + c;
        double v = 0.0;
#ifdef RVLP_GP_SKIP_BUILD
        if (has_tile && i < N && k <= i) { v = (i == k) ? 10.0 + i : 0.001; } else
#endif
        {
          // branch-free (rvlp_gpcov.cuh): every element is evaluated, out-of-triangle ones are discarded, so
          // the compiler interleaves the TT*TT independent chains instead of running them one after the other
          const int ic = i < N ? i : N - 1, kc = k < N ? k : N - 1;
          const double kv = gp_cov(T.t[ic] - T.t[kc], hyp);
          const bool in_tri = has_tile && i < N && k <= i;
          v = in_tri ? kv : 0.0;
          if (in_tri && i == k) v += T.e2[i] + sr[kHdr + P.n_inst + T.inst[i]];
          if (has_tile && i == N && k < N) v = resid[k];
        }
        a[r][c] = v;
      }
    }
    // Column sweep.  a_ik -= (a_ij / p)(a_kj) with p the pivot: no square roots anywhere
    // (alpha_j^2 = a_Nj^2 / p, ln L_jj = ln(p) / 2).  Padding rows hold zeros; entries in
    // columns >= N are never read, so only the "row / column already finished" masks remain and
    // those are needed only while j runs through the tile's own rows / columns.
    double quad = 0.0, prodm = 1.0;
    int exsum = 0, par = 0;
    // entry (i, k) only receives updates from columns j < k: the tile is final once j reaches its last column
    const int jlast = has_tile ? c0 + TT - 1 : -1;           // last column this tile publishes
    const int warp_last = __reduce_max_sync(0xffffffffu, jlast);
    // Two-level column loop so that the column-within-tile index cj is a compile-time constant: the
    // publishing threads then read a[r][cj] with static register indices (a run-time cj compiles to an
    // indexed branch per column, which measured as ~40% of the sweep).
    for (int Jt = 0; Jt < nt; ++Jt) {
#pragma unroll
    for (int cj = 0; cj < TT; ++cj) {
      const int j = Jt * TT + cj;
      if (j >= N) break;
      double* col = colbuf + par * rows;
      par ^= 1;
      if (has_tile && J == Jt) {                              // owners publish column j
#pragma unroll
        for (int r = 0; r < TT; ++r) col[r0 + r] = a[r][cj];
      }
      __syncthreads();
#ifndef RVLP_GP_ABL_NOSCALAR
      if (tid == 0) {                                         // alpha_j^2 and ln L_jj
        const double piv = col[j], aN = col[N];
        quad = fma(aN * aN, 1.0 / piv, quad);
        const int h = __double2hiint(piv);
        if ((unsigned)(h - 0x00100000) < 0x7fe00000u) {
          prodm *= __hiloint2double((h & 0x000fffff) | 0x3ff00000, __double2loint(piv));
          exsum += (h >> 20) - 1023;
        } else {
          prodm *= piv;                                        // 0 / negative / NaN: let log() say so
        }
      }
#endif
#ifdef RVLP_GP_ABL_NOUPDATE
      continue;
#endif
      if (j >= warp_last) continue;                           // every tile of this warp is final
      if (j < jlast) {
        const double ip = 1.0 / col[j];                        // inf / NaN when not positive definite
        double Li[TT], Lk[TT];
        const double2* ci = reinterpret_cast<const double2*>(col + r0);   // r0, c0 multiples of an even TT
        const double2* ck = reinterpret_cast<const double2*>(col + c0);
#pragma unroll
        for (int r = 0; r < TT; r += 2) {
          const double2 v = ci[r / 2];
          Li[r] = v.x;
          Li[r + 1] = v.y;
        }
#pragma unroll
        for (int c = 0; c < TT; c += 2) {
          const double2 v = ck[c / 2];
          Lk[c] = v.x * ip;
          Lk[c + 1] = v.y * ip;
        }
        if (j >= c0) {                                         // sweeping through this tile's own columns / rows
#pragma unroll
          for (int c = 0; c < TT; ++c)
            if (c0 + c <= j) Lk[c] = 0.0;
#pragma unroll
          for (int r = 0; r < TT; ++r)
            if (r0 + r <= j) Li[r] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < TT; ++r)
#pragma unroll
          for (int c = 0; c < TT; ++c) a[r][c] = fma(-Li[r], Lk[c], a[r][c]);
      }
    }
    }
    if (tid == 0) {
      // sum_j ln L_jj = 1/2 ln prod piv_j
      const double logdet = 0.5 * fma((double)exsum, 0.6931471805599453, log(prodm));
      const double ll = -0.5 * quad - logdet - 0.5 * (double)N * kLog2Pi;
      double r = ll + lp + lhp;                                // fit.py:7898-7900
      r += P.jacobian;
      r += P.renorm;
      out[s] = r;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ K3, blocked (panel) version
// Same register tiling, but the sweep advances one PANEL of TT columns per step instead of one column:
//   1. the diagonal tile's owner factorises its TT x TT tile in registers and publishes L_d and 1/diag;
//   2. the tiles below it solve X L_d^T = A in registers (TRSM) and publish X, stored k-major;
//   3. every trailing tile subtracts P_I P_J^T: TT rank-1 updates with no barrier, mask or division between.
// Two barriers per panel (N/TT panels) instead of one per column.  The residual row N rides along as before:
// after step 2 its entries are alpha_j; ln L_jj comes from the pivots.  Deterministic: fixed tile ownership,
// fixed summation order, fixed-order final reduction.
struct GpBlockedSmem { int off_resid, off_d, off_p, off_red, total; };
__host__ __device__ inline GpBlockedSmem gp_blocked_smem(const DevProblem& P, const SmemLayout& L, int TT) {
  GpBlockedSmem G;
  int o = (L.total + 15) & ~15;
  const int nt = (P.n_epochs + 1 + TT - 1) / TT;
  G.off_resid = o; o += ((P.n_epochs + 2) & ~1) * 8;
  G.off_d = o; o += (TT * TT + TT) * 8;
  o = (o + 15) & ~15;
  G.off_p = o; o += nt * TT * TT * 8;
  G.off_red = o; o += 2 * kThreads * 8;
  G.total = o;
  return G;
}

#ifdef RVLP_GP_TIMING
// Phase timing of the blocked kernel (experiments only; tools/gp_phase_time.py): block 0, thread of the last tile.
__device__ unsigned long long g_gp_timing[32];
#ifndef RVLP_GP_TIMING_SEL
#define RVLP_GP_TIMING_SEL -1   /* -1: all phases (spills: indicative only); k: phase k alone, 4 registers */
#endif
#if RVLP_GP_TIMING_SEL < 0
#define GPT_DECL unsigned long long gpt[16] = {0}; long long gpt_t = 0; const bool gpt_on = blockIdx.x == 0 && tid == nt * (nt + 1) / 2 - 1;
#define GPT_START() do { if (gpt_on) gpt_t = clock64(); } while (0)
#define GPT_LAP(k) do { if (gpt_on) { const long long n_ = clock64(); gpt[k] += n_ - gpt_t; gpt_t = n_; } } while (0)
#define GPT_COUNT() do { if (gpt_on) gpt[8] += 1; } while (0)
#define GPT_FLUSH() do { if (gpt_on) for (int k = 0; k < 16; ++k) atomicAdd(&g_gp_timing[k], gpt[k]); } while (0)
#else
#define GPT_DECL unsigned long long gpt_acc = 0, gpt_n = 0; long long gpt_t = 0; const bool gpt_on = blockIdx.x == 0 && tid == nt * (nt + 1) / 2 - 1;
#define GPT_START() do { gpt_t = clock64(); } while (0)
#define GPT_LAP(k) do { const long long n_ = clock64(); if ((k) == RVLP_GP_TIMING_SEL) gpt_acc += n_ - gpt_t; gpt_t = n_; } while (0)
#define GPT_COUNT() do { gpt_n += 1; } while (0)
#define GPT_FLUSH() do { if (gpt_on) { atomicAdd(&g_gp_timing[RVLP_GP_TIMING_SEL], gpt_acc); atomicAdd(&g_gp_timing[8], gpt_n); } } while (0)
#endif
#else
#define GPT_DECL
#define GPT_START() do {} while (0)
#define GPT_LAP(k) do {} while (0)
#define GPT_COUNT() do {} while (0)
#define GPT_FLUSH() do {} while (0)
#endif

template <int TT>
__global__ void __launch_bounds__(kThreads, (TT >= 8 ? 1 : 2))
gp_logprob_blocked_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);
  const GpBlockedSmem G = gp_blocked_smem(P, L, TT);
  stage_problem(P, L, smem);
  const Tables T = tables_of(P, L, smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  double* scratch = reinterpret_cast<double*>(smem + L.off_scratch);
  double* resid = reinterpret_cast<double*>(smem + G.off_resid);
  double* dbuf = reinterpret_cast<double*>(smem + G.off_d);
  double* pbuf = reinterpret_cast<double*>(smem + G.off_p);
  double* red = reinterpret_cast<double*>(smem + G.off_red);
  const int N = P.n_epochs;
  const int nt = (N + 1 + TT - 1) / TT;          // tile rows (incl. the residual row N)
  const int ntc = (N + TT - 1) / TT;             // panels (columns 0..N-1)
  int J = 0, rem = tid;
  while (J < nt && rem >= nt - J) { rem -= nt - J; ++J; }
  const int I = J + rem;
  const bool has_tile = J < nt;
  const int r0 = I * TT, c0 = J * TT;
  const int IN = N / TT, rN = N - IN * TT;       // where the residual row lives
  GPT_DECL

  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    GPT_START();
    if (warp == 0) sample_prologue(P, T, theta, s, s + 1, scratch, rec, lane, true, 1, reinterpret_cast<double*>(smem + L.off_pv));
    __syncthreads();
    GPT_LAP(0);
    const double* sr = scratch;
    const int flags = __double2loint(sr[1]);
    const double lp = sr[0], lhp = sr[4];
    if (flags & (F_JIT | F_HYPER | F_PRIOR)) {               // fit.py:7857-7886
      if (tid == 0) out[s] = -INFINITY;
      __syncthreads();
      continue;
    }
    int nonfinite = (flags & F_PLANET) ? 1 : 0;              // fit.py:8022-8024
    if (!nonfinite) {
      for (int i = tid; i < N; i += kThreads) {              // residual v - mean, fit.py:7994-8043, 8059
        double tt[1] = {T.t[i]}, rv[1];
        model_rv<1>(P, sr, tt, rv, -1, true);
        const double mean = rv[0] + sr[kHdr + T.inst[i]];
        if (!(fabs(mean) <= 1.79769313486231570e308)) nonfinite = 1;
        resid[i] = T.v[i] - mean;
      }
    }
    if (__syncthreads_or(nonfinite)) {                       // fit.py:8082-8083
      if (tid == 0) {
        double r = -INFINITY + lp + lhp;
        r += P.jacobian;
        r += P.renorm;
        out[s] = r;
      }
      __syncthreads();
      continue;
    }
    GPT_LAP(1);
    const double* row = theta + s * P.ndim;
    const double Aamp = model_param(T, row, P.n_model + 0), le = model_param(T, row, P.n_model + 1);
    const double lpp = model_param(T, row, P.n_model + 2), Pg = model_param(T, row, P.n_model + 3);
    const GpHyper hyp = gp_hyper(Aamp, le, lpp, Pg);         // gp.py:145-156

    double a[TT][TT];                                         // gp.py:145-156, fit.py:8094-8096
#pragma unroll
    for (int r = 0; r < TT; ++r) {
      const int i = r0 + r;
#pragma unroll
      for (int c = 0; c < TT; ++c) {
        const int k = c0 + c;
        double v = 0.0;
        {
          // branch-free (rvlp_gpcov.cuh): every element is evaluated, out-of-triangle ones are discarded, so
          // the compiler interleaves the TT*TT independent chains instead of running them one after the other
          const int ic = i < N ? i : N - 1, kc = k < N ? k : N - 1;
          const double kv = gp_cov(T.t[ic] - T.t[kc], hyp);
          const bool in_tri = has_tile && i < N && k <= i;
          v = in_tri ? kv : 0.0;
          if (in_tri && i == k) v += T.e2[i] + sr[kHdr + P.n_inst + T.inst[i]];
          if (has_tile && i == N && k < N) v = resid[k];
        }
        a[r][c] = v;
      }
    }
    double quad = 0.0, prodm = 1.0;      // partial alpha.alpha and pivot product of THIS thread
    int exsum = 0;
    GPT_LAP(2);
    for (int Jt = 0; Jt < ntc; ++Jt) {
      // ---- 1. diagonal tile: unblocked Cholesky in registers
      if (has_tile && I == Jt && J == Jt) {
#if defined(RVLP_GP_TIMING) && RVLP_GP_TIMING_SEL < 0
        const long long d0_ = clock64();
#endif
        double invd[TT];
#pragma unroll
        for (int c = 0; c < TT; ++c) {
          const bool valid = c0 + c < N;                       // columns >= N: identity (no-op) column
          const double d = a[c][c];
          const double inv = valid ? pivot_rsqrt(d) : 1.0;     // NaN when not positive definite (as jax)
          invd[c] = inv;
          a[c][c] = valid ? d * inv : 1.0;
          if (valid) {
            const int h = __double2hiint(d);
            if ((unsigned)(h - 0x00100000) < 0x7fe00000u) {
              prodm *= __hiloint2double((h & 0x000fffff) | 0x3ff00000, __double2loint(d));
              exsum += (h >> 20) - 1023;
            } else {
              prodm *= d;                                      // 0 / negative / NaN: let log() say so
            }
          }
#pragma unroll
          for (int r = c + 1; r < TT; ++r) a[r][c] = valid ? a[r][c] * inv : 0.0;
#pragma unroll
          for (int r = c + 1; r < TT; ++r)
#pragma unroll
            for (int k = c + 1; k <= r; ++k) a[r][k] = fma(-a[r][c], a[k][c], a[r][k]);
        }
        if (IN == Jt) {                                        // the residual row sits in this tile
#pragma unroll
          for (int r = 0; r < TT; ++r)
            if (r == rN) {
#pragma unroll
              for (int c = 0; c < TT; ++c)
                if (c < r) quad = fma(a[r][c], a[r][c], quad);
            }
        }
#pragma unroll
        for (int r = 0; r < TT; ++r)
#pragma unroll
          for (int c = 0; c < TT; ++c) dbuf[r * TT + c] = (c <= r) ? a[r][c] : 0.0;
#pragma unroll
        for (int c = 0; c < TT; ++c) dbuf[TT * TT + c] = invd[c];
#if defined(RVLP_GP_TIMING) && RVLP_GP_TIMING_SEL < 0
        if (blockIdx.x == 0) atomicAdd(&g_gp_timing[9], (unsigned long long)(clock64() - d0_));
#endif
      }
      __syncthreads();
      GPT_LAP(Jt == 0 ? 3 : 4);
      // ---- 2. panel tiles: X L_d^T = A, publish X k-major
      if (has_tile && J == Jt && I > Jt) {
#if defined(RVLP_GP_TIMING) && RVLP_GP_TIMING_SEL < 0
        const long long d0_ = clock64();
#endif
#pragma unroll
        for (int c = 0; c < TT; ++c) {
          const double inv = dbuf[TT * TT + c];
#pragma unroll
          for (int r = 0; r < TT; ++r) {
            double x = a[r][c];
#pragma unroll
            for (int k = 0; k < c; ++k) x = fma(-a[r][k], dbuf[c * TT + k], x);
            a[r][c] = x * inv;
          }
        }
        if (I == IN) {                                         // alpha_j for this panel's columns
#pragma unroll
          for (int r = 0; r < TT; ++r)
            if (r == rN) {
#pragma unroll
              for (int c = 0; c < TT; ++c) quad = fma(a[r][c], a[r][c], quad);
            }
        }
        double* pb = pbuf + I * TT * TT;
#pragma unroll
        for (int k = 0; k < TT; ++k)
#pragma unroll
          for (int r = 0; r < TT; ++r) pb[k * TT + r] = a[r][k];
#if defined(RVLP_GP_TIMING) && RVLP_GP_TIMING_SEL < 0
        if (blockIdx.x == 0 && I == Jt + 1) atomicAdd(&g_gp_timing[10], (unsigned long long)(clock64() - d0_));
#endif
      }
      __syncthreads();
      GPT_LAP(5);
      // ---- 3. trailing tiles: a -= P_I P_J^T
      if (has_tile && J > Jt) {
        const double2* pi = reinterpret_cast<const double2*>(pbuf + I * TT * TT);
        const double2* pj = reinterpret_cast<const double2*>(pbuf + J * TT * TT);
#pragma unroll
        for (int k = 0; k < TT; ++k) {
          double Li[TT], Lk[TT];
#pragma unroll
          for (int r = 0; r < TT; r += 2) {
            const double2 u = pi[(k * TT + r) / 2], w = pj[(k * TT + r) / 2];
            Li[r] = u.x; Li[r + 1] = u.y;
            Lk[r] = w.x; Lk[r + 1] = w.y;
          }
#pragma unroll
          for (int r = 0; r < TT; ++r)
#pragma unroll
            for (int c = 0; c < TT; ++c) a[r][c] = fma(-Li[r], Lk[c], a[r][c]);
        }
      }
      GPT_LAP(6);
    }
    // fixed-order reduction of the per-thread partials (only a few threads hold non-trivial ones)
    red[tid] = quad;
    red[kThreads + tid] = 0.5 * fma((double)exsum, 0.6931471805599453, log(prodm));   // sum ln L_jj of this thread
    __syncthreads();
    if (tid == 0) {
      double q = 0.0, logdet = 0.0;
      for (int t = 0; t < kThreads; ++t) { q += red[t]; logdet += red[kThreads + t]; }
      const double ll = -0.5 * q - logdet - 0.5 * (double)N * kLog2Pi;
      double r = ll + lp + lhp;                                // fit.py:7898-7900
      r += P.jacobian;
      r += P.renorm;
      out[s] = r;
    }
    __syncthreads();
    GPT_LAP(7);
    GPT_COUNT();
  }
  GPT_FLUSH();
}

// ------------------------------------------------------------------ K7: GP conditioning (row f-4)
// Replaces the per-sample loop of GPFitter's posterior predictions (fit.py:6383-6414, 7494-7554:
// `gp.condition(y = vel - gamma - planets - trend, X_test = times).gp.mean`) and GPFitter._compute_gp_chi2
// (fit.py:5386-5429, alpha.alpha with alpha = L^-1 r).  tinygp's conditional mean with its default zero mean
// function is  mu*(t*) = k(t*, t)^T C^-1 r  (restated; "parity unpinned" like K3).
// One CTA per sample: the smem factorisation of gp_logprob_kernel (row N = residual -> alpha), then a
// column-oriented back substitution beta = L^-T alpha that walks L's rows (contiguous in the packed
// triangle), then mu*_i = sum_j k(t*_i - t_j) beta_j with one thread per test time.
struct GpPredictSmem { int off_tri, off_beta, off_red, total; };
__host__ __device__ inline GpPredictSmem gp_predict_smem(const DevProblem& P, const SmemLayout& L) {
  GpPredictSmem G;
  int o = (L.total + 15) & ~15;
  G.off_tri = o; o += gp_tri_doubles(P.n_epochs) * 8;
  G.off_beta = o; o += ((P.n_epochs + 1) & ~1) * 8;
  G.off_red = o; o += 64 * 8;
  G.total = o;
  return G;
}

__global__ void __launch_bounds__(kThreads)
gp_predict_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, const double* __restrict__ times,
                  int64_t T_n, double* __restrict__ mean_out, double* __restrict__ chi2_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);
  const GpPredictSmem G = gp_predict_smem(P, L);
  stage_problem(P, L, smem);
  const Tables T = tables_of(P, L, smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  double* scratch = reinterpret_cast<double*>(smem + L.off_scratch);
  double* Cm = reinterpret_cast<double*>(smem + G.off_tri);
  double* beta = reinterpret_cast<double*>(smem + G.off_beta);
  double* red = reinterpret_cast<double*>(smem + G.off_red);
  const int N = P.n_epochs;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);

  for (int64_t s = blockIdx.x; s < S; s += gridDim.x) {
    if (warp == 0) sample_prologue(P, T, theta, s, s + 1, scratch, rec, lane, false, 1, nullptr);
    __syncthreads();
    const double* sr = scratch;
    const int flags = __double2loint(sr[1]);
    // The reference raises for such a sample (Planet.__init__ ValueError / build_kernel ValueError); rows are NaN.
    int bad = (flags & (F_PLANET | F_HYPER)) ? 1 : 0;
    const double* row = theta + s * P.ndim;
    const double A = model_param(T, row, P.n_model + 0), le = model_param(T, row, P.n_model + 1);
    const double lpp = model_param(T, row, P.n_model + 2), Pg = model_param(T, row, P.n_model + 3);
    const GpHyper hyp = gp_hyper(A, le, lpp, Pg);            // gp.py:145-156
    if (!bad) {
      for (int i = tid; i < N; i += kThreads) {              // fit.py:6375-6380, 7536-7550
        double tt[1] = {T.t[i]}, rv[1];
        model_rv<1>(P, sr, tt, rv, -1, true);
        Cm[tri(N, i)] = (T.v[i] - sr[kHdr + T.inst[i]]) - rv[0];
      }
      const int npairs = N * (N + 1) / 2;
      for (int p = tid; p < npairs; p += kThreads) {         // gp.py:145-156; diag fit.py:6399, 7531
        int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
        while (tri(i + 1, 0) <= p) ++i;
        while (tri(i, 0) > p) --i;
        const int j = p - tri(i, 0);
        double c = gp_cov(T.t[i] - T.t[j], hyp);
        if (i == j) c += T.e2[i] + sr[kHdr + P.n_inst + T.inst[i]];
        Cm[p] = c;
      }
    }
    if (__syncthreads_or(bad)) {
      for (int64_t i = tid; i < T_n; i += kThreads) mean_out[s * T_n + i] = qnan;
      if (tid == 0 && chi2_out) chi2_out[s] = qnan;
      __syncthreads();
      continue;
    }
    for (int j = 0; j < N; ++j) {                            // right-looking Cholesky, rows 0..N
      const double djj = sqrt(Cm[tri(j, j)]);
      const double inv = 1.0 / djj;
      __syncthreads();
      for (int i = j + 1 + tid; i <= N; i += kThreads) Cm[tri(i, j)] *= inv;
      if (tid == 0) Cm[tri(j, j)] = djj;
      __syncthreads();
      const int m = N - j;
      const int cnt = m * (m + 1) / 2;
      for (int p = tid; p < cnt; p += kThreads) {
        int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= p) ++a;
        while (a * (a + 1) / 2 > p) --a;
        const int b = p - a * (a + 1) / 2;
        const int i = j + 1 + a, k = j + 1 + b;
        if (i == N && k == N) continue;
        Cm[tri(i, k)] = fma(-Cm[tri(i, j)], Cm[tri(k, j)], Cm[tri(i, k)]);
      }
      __syncthreads();
    }
    // chi^2 = alpha . alpha (fit.py:5428-5429), fixed-order reduction
    double q = 0.0;
    for (int i = tid; i < N; i += kThreads) {
      const double a = Cm[tri(N, i)];
      beta[i] = a;
      q = fma(a, a, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (tid == 0 && chi2_out) {
      double quad = 0.0;
      for (int w = 0; w < kWarps; ++w) quad += red[w];
      chi2_out[s] = quad;
    }
    // beta = L^-T alpha, last unknown first: beta_j = y_j / L_jj, then y_i -= L_ji beta_j for i < j
    for (int j = N - 1; j >= 0; --j) {
      const double bj = beta[j] / Cm[tri(j, j)];
      __syncthreads();                                       // everyone has read beta[j]
      if (tid == 0) beta[j] = bj;
      for (int i = tid; i < j; i += kThreads) beta[i] = fma(-Cm[tri(j, i)], bj, beta[i]);
      __syncthreads();
    }
    // conditional mean at the requested times
    for (int64_t i = tid; i < T_n; i += kThreads) {
      const double ts = times[i];
      double acc = 0.0;
#pragma unroll 4
      for (int j = 0; j < N; ++j) acc = fma(gp_cov(ts - T.t[j], hyp), beta[j], acc);
      mean_out[s * T_n + i] = acc;
    }
    __syncthreads();
  }
}

}  // namespace rvlp
