// rvlp_gp_smem.cuh — K3 for up to ~136 epochs: one 4-warp CTA per sample in flight, the factor in SHARED MEMORY as
// 8 x 8 tiles in tensor-core fragment order, every Gram sum on the fp64 tensor cores.
//
// GPLogPosterior.log_probability (/root/reference/src/ravest/fit.py:7836-7901, 8062-8105; kernel gp.py:145-156):
//   C = K(t, t) + diag(sigma^2 + jit^2) = L L^T,  alpha = L^-1 r,  ll = -1/2 alpha.alpha - sum ln L_ii - N/2 ln 2 pi.
//
// Why another kernel.  rvlp_gp_pipe.cuh keeps the triangle in REGISTERS (6 x 6 tiles, DFMA updates): two samples per
// SM, a barrier-bound dependency chain, 25 % of the fp64 peak at N = 120.  rvlp_gp_batch.cuh keeps it in HBM and pays
// ~N^3 / 12 bytes of traffic per sample.  Here the factor (54 KB at N = 120: the strictly lower 8 x 8 tiles) stays in
// shared memory, THREE samples per SM, and the O(N^3) part is `mma.sync.m8n8k4.f64` (SASS DMMA.884): one warp
// instruction per 256 FMAs keeps an SM sub-partition's fp64 pipe busy for 16 cycles, so the few warps that fit are
// enough to fill it.  Left-looking by block column j (NT = ceil(N / 8) of them), tile rows dealt round-robin to the
// four warps starting at the owner of row j:
//   A. every warp: G_ij = sum_{k<j} L_ik L_jk^T for its tiles (i, j), i >= j - both operands are tiles of L in
//      "A-fragment order" (lane l holds L[l / 4][l % 4] and L[l / 4][4 + l % 4]; L^T as the B operand has the SAME
//      lane map), one conflict-free LDS.128 per tile and k; then C_ij - G_ij with the covariance generated on the fly
//      in accumulator layout (phase-factored periodic term, 19 fp64 instructions per element; C is never stored);
//   B. the owner of row j does the diagonal tile FIRST: 8 x 8 Cholesky in accumulator layout (quad shuffles; the
//      reciprocal of the pivot is off the shuffle chain), the residual's entries alpha_j = L_jj^-1 (r_j - sum_k L_jk
//      alpha_k) ride along as a ninth row; it publishes L_jj, `bar.arrive`s and goes on with its other tiles;
//   C. the others `bar.sync` on that barrier once their Gram sums are done, solve X L_jj^T = C - G in accumulator
//      layout and store the tile in fragment order (two STS.64 per lane, no shuffles);
//   one __syncthreads per block column.
// Per sample: 1120 DMMA at N = 120 instead of 23.7 k DFMA warp instructions, 15 + 15 barriers instead of ~60.
// The mean model / priors / reject flags come from gpb_prologue_kernel (one warp per sample, rvlp_gp_batch.cuh).
// Deterministic: fixed summation order per tile, integer ticket only decides WHICH CTA takes a sample - out[s] depends
// on (theta[s], epochs) only.
#pragma once
#include "rvlp_gp_batch.cuh"

namespace rvlp {

#ifndef RVLP_GPS_WORKERS
#define RVLP_GPS_WORKERS 3
#endif
constexpr int kGsWorkers = RVLP_GPS_WORKERS;        // worker warps: worker w owns the tile rows i with i % kGsWorkers == w
constexpr int kGsThreads = 32 * (kGsWorkers + 1);   // + the diagonal warp (with 3 workers: an SM sub-partition of its own)
#ifndef RVLP_GPS_MB
#define RVLP_GPS_MB 3
#endif

__host__ __device__ inline int gps_tile_rows(int N) { return (N + 7) / 8; }
__host__ __device__ inline int gps_smem_bytes(int N) {
  const int nt = gps_tile_rows(N), np = nt * 8;
  return nt * (nt - 1) / 2 * 512 + (64 + 8) * 8 + 2 * (64 + 32) * 8 + 6 * np * 8 + 64;
}

__device__ __forceinline__ void gps_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void gps_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Per-sample shared-memory views and constants of the block-column steps.
struct GpsView {
  double2* Ls;
  double *ljj, *invd, *dtile, *t, *cph, *sph, *dn, *r, *al;
  double inv_le, g2, A2;
  int N, NT;
};

// Covariance entries (rows 8 i + g, columns 8 j + 2 q + {0, 1}) in accumulator layout, without the white-noise term.
// gp.py:145-156 with sin^2(pi tau / P) = (1 - cos(b_i - b_j)) / 2 (rvlp_gp_batch.cuh: gpb_entries4).
__device__ __forceinline__ void gps_cov2(const GpsView& V, int i, int j, int g, int q, double& v0, double& v1) {
  const int col = j * 8 + 2 * q, row = i * 8 + g;
  const double tc0 = V.t[col], tc1 = V.t[col + 1];
  const double cc0 = V.cph[col], cc1 = V.cph[col + 1], sc0 = V.sph[col], sc1 = V.sph[col + 1];
  const double tr = V.t[row], cr = V.cph[row], sr = V.sph[row];
  const double cd0 = fma(cr, cc0, sr * sc0), cd1 = fma(cr, cc1, sr * sc1);
  const double q0 = (tr - tc0) * V.inv_le, q1 = (tr - tc1) * V.inv_le;
  v0 = gp_exp_scaled_neg(fma(V.g2, cd0, fma(-0.5 * q0, q0, -V.g2)), V.A2);
  v1 = gp_exp_scaled_neg(fma(V.g2, cd1, fma(-0.5 * q1, q1, -V.g2)), V.A2);
  if (i == V.NT - 1) {                                     // identity padding beyond the N epochs (warp-uniform test;
    if (row >= V.N || col >= V.N) v0 = row == col ? 1.0 : 0.0;        // j <= i, so the last tile ROW covers the last column too)
    if (row >= V.N || col + 1 >= V.N) v1 = row == col + 1 ? 1.0 : 0.0;
  }
}

// ---- workers.  A tile's own 512-byte slot carries it through both phases, so the phases are runtime loops over small
// GROUPS of G tiles (rows i0, i0 + step, ...) with no register state between them - one copy of the code instead of
// one per tile count (the per-count unrolled version stalled on instruction fetch: ncu no_instruction 1.3 per issue).
//
// ahead, column jn, one column EARLY (while the diagonal warp factors tile (jn - 1, jn - 1)): the part of the Gram sum
// that is already final, sum_{k < jn - 1} L_ik L_jn,k^T, subtracted from the covariance entries; the partial tile is
// parked in its slot in accumulator layout (lane l: doubles 2 l, 2 l + 1).
template <int G>
__device__ __forceinline__ void gps_ahead_group(const GpsView& V, int jn, int i0, int step, int lane) {
  const int g = lane >> 2, q = lane & 3;
  double acc[G][2];
  double2* ai[G];
#pragma unroll
  for (int t = 0; t < G; ++t) {
    acc[t][0] = acc[t][1] = 0.0;
    const int i = i0 + step * t;
    ai[t] = V.Ls + (size_t)(i * (i - 1) / 2) * 32 + lane;
  }
  const double2* bj = V.Ls + (size_t)(jn * (jn - 1) / 2) * 32 + lane;
#pragma unroll 2
  for (int k = 0; k < jn - 1; ++k) {
    const double2 b = bj[k * 32];
    double2 a[G];
#pragma unroll
    for (int t = 0; t < G; ++t) a[t] = ai[t][k * 32];
#pragma unroll
    for (int t = 0; t < G; ++t) dmma884(acc[t][0], acc[t][1], a[t].x, b.x);
#pragma unroll
    for (int t = 0; t < G; ++t) dmma884(acc[t][0], acc[t][1], a[t].y, b.y);
  }
  double v[G][2];
#pragma unroll
  for (int t = 0; t < G; ++t) gps_cov2(V, i0 + step * t, jn, g, q, v[t][0], v[t][1]);
#pragma unroll
  for (int t = 0; t < G; ++t) ai[t][jn * 32] = make_double2(v[t][0] - acc[t][0], v[t][1] - acc[t][1]);
}

// The worker that owns tile ROW jn also prepares the diagonal tile for the diagonal warp: C_jn,jn - sum_{k < jn - 1}
// L_jn,k L_jn,k^T (accumulator layout) and the matching part of the residual row's sum_k L_jn,k alpha_k go to the
// double-buffered hand-over block V.dtile.
__device__ __forceinline__ void gps_ahead_diag(const GpsView& V, int jn, int lane) {
  const int g = lane >> 2, q = lane & 3;
  double dg[2] = {0.0, 0.0}, eg[2] = {0.0, 0.0}, part = 0.0;
  const double2* bj = V.Ls + (size_t)(jn * (jn - 1) / 2) * 32 + lane;
  const double* al = V.al + q;
#pragma unroll 2
  for (int k = 0; k < jn - 1; ++k) {
    const double2 b = bj[k * 32];
    dmma884(dg[0], dg[1], b.x, b.x);
    dmma884(eg[0], eg[1], b.y, b.y);
    part = fma(b.x, al[k * 8], fma(b.y, al[k * 8 + 4], part));
  }
  double d0, d1;
  gps_cov2(V, jn, jn, g, q, d0, d1);
  const int row = jn * 8 + g, col = jn * 8 + 2 * q;
  if (row < V.N) {                                         // white noise on the diagonal (fit.py:8094-8096)
    const double dn = V.dn[row];
    d0 += row == col ? dn : 0.0;
    d1 += row == col + 1 ? dn : 0.0;
  }
  double* h = V.dtile + (jn & 1) * 96;
  *reinterpret_cast<double2*>(h + 2 * lane) = make_double2(d0 - (dg[0] + eg[0]), d1 - (dg[1] + eg[1]));
  h[64 + lane] = part;
}

// solve, column j (L_jj is published): the last term of the Gram sum (k = j - 1, stored by the previous column's
// solve), X L_jj^T = C - G in accumulator layout (the published L_jj is STRICTLY lower, zeros elsewhere: no per-lane
// conditions), then the tile goes back to its slot in fragment order: element (row g, column cc) at double
// 2 (4 g + cc % 4) + cc / 4, two STS.64 per lane.
struct GpsLjj {
  double lr0[8], lr1[8], ivd[8], i0s, i1s;                 // rows 2 q and 2 q + 1 of L_jj (strictly lower part), 1 / diag
};
__device__ __forceinline__ void gps_load_ljj(const GpsView& V, int lane, GpsLjj& L) {
  const int q = lane & 3;
#pragma unroll
  for (int m = 0; m < 8; m += 2) {
    const double2 x0 = *reinterpret_cast<const double2*>(V.ljj + (2 * q) * 8 + m);
    const double2 x1 = *reinterpret_cast<const double2*>(V.ljj + (2 * q + 1) * 8 + m);
    const double2 iv = *reinterpret_cast<const double2*>(V.invd + m);
    L.lr0[m] = x0.x; L.lr0[m + 1] = x0.y;
    L.lr1[m] = x1.x; L.lr1[m + 1] = x1.y;
    L.ivd[m] = iv.x; L.ivd[m + 1] = iv.y;
  }
  L.i0s = V.invd[2 * q];
  L.i1s = V.invd[2 * q + 1];
}
template <int G>
__device__ __forceinline__ void gps_solve_group(const GpsView& V, const GpsLjj& L, int j, int i0, int step, int lane) {
  const int g = lane >> 2, q = lane & 3, quad = lane & ~3;
  double2* slot[G];
  double c[G][2];
#pragma unroll
  for (int t = 0; t < G; ++t) {
    const int i = i0 + step * t;
    slot[t] = V.Ls + (size_t)(i * (i - 1) / 2 + j) * 32;
    const double2 p = slot[t][lane];
    c[t][0] = p.x;
    c[t][1] = p.y;
  }
  if (j > 0) {
    const double2 b = V.Ls[(size_t)(j * (j - 1) / 2 + j - 1) * 32 + lane];
    double e[G][2];
#pragma unroll
    for (int t = 0; t < G; ++t) {
      const double2 a = slot[t][lane - 32];                // tile (i, j - 1) sits right before (i, j)
      e[t][0] = e[t][1] = 0.0;
      dmma884(e[t][0], e[t][1], a.x, b.x);
      dmma884(e[t][0], e[t][1], a.y, b.y);
    }
#pragma unroll
    for (int t = 0; t < G; ++t) {
      c[t][0] -= e[t][0];
      c[t][1] -= e[t][1];
    }
  }
#pragma unroll
  for (int m = 0; m < 7; ++m) {
#pragma unroll
    for (int t = 0; t < G; ++t) {
      const double xm = __shfl_sync(0xffffffffu, (m & 1) ? c[t][1] : c[t][0], quad | (m >> 1)) * L.ivd[m];   // X[g][m]
      if (m < 6) c[t][0] = fma(-xm, L.lr0[m], c[t][0]);    // (column 2 q <= 6)
      c[t][1] = fma(-xm, L.lr1[m], c[t][1]);
    }
  }
  const int e0 = 2 * (4 * g + ((2 * q) & 3)) + ((2 * q) >> 2), e1 = 2 * (4 * g + ((2 * q + 1) & 3)) + ((2 * q + 1) >> 2);
#pragma unroll
  for (int t = 0; t < G; ++t) {
    double* tile = reinterpret_cast<double*>(slot[t]);
    tile[e0] = c[t][0] * L.i0s;
    tile[e1] = c[t][1] * L.i1s;
  }
}

// ---- diagonal warp: 8 x 8 Cholesky of tile (j, j) in accumulator layout, alpha_j as a ninth row; publishes the
// strictly lower part of L_jj (row-major, zeros elsewhere), 1 / diag, alpha_j, and adds this block's chi^2 / sum ln L_kk.
// The pivot chain is shuffle -> reciprocal (seed + 3 fp64 operations) -> two multiplies -> shuffle; the reciprocal
// square roots that scale L are off it.
__device__ __forceinline__ void gps_factor_diag(const GpsView& V, int j, int lane, double c0, double c1, double part, double* red) {
  const int g = lane >> 2, q = lane & 3, quad = lane & ~3;
  part += __shfl_xor_sync(0xffffffffu, part, 1);
  part += __shfl_xor_sync(0xffffffffu, part, 2);
  double z = V.r[j * 8 + g] - part;                        // lanes (g, 0) carry the residual row
  double inv[8];
  double p8 = 1.0, chi = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double mine = (k & 1) ? c1 : c0;                 // this lane's entry of column pair k / 2
    const double low = g > k ? mine : 0.0;                 // rows <= k take no part in the update of the trailing block
    const double piv = __shfl_sync(0xffffffffu, mine, k * 4 + (k >> 1));
    const double cgk = __shfl_sync(0xffffffffu, low, quad | (k >> 1));             // C[g][k], g > k
    const double ck0 = __shfl_sync(0xffffffffu, low, (2 * q) * 4 + (k >> 1));      // C[2q][k], 2q > k
    const double ck1 = __shfl_sync(0xffffffffu, low, (2 * q + 1) * 4 + (k >> 1));  // C[2q+1][k], 2q+1 > k
    const double zk = __shfl_sync(0xffffffffu, z, k * 4);
    const double sg = cgk * rcp64_3(piv);
    c0 = fma(-sg, ck0, c0);
    c1 = fma(-sg, ck1, c1);
    inv[k] = pivot_rsqrt(piv);
    p8 = piv > 0.0 ? p8 * piv : __longlong_as_double(0x7ff8000000000000ll);        // not positive definite -> NaN (as jax)
    const double ak = zk * inv[k];                         // alpha[8 j + k]
    chi = fma(ak, ak, chi);
    z = fma(-ak * inv[k], cgk, z);                         // L[g][k] = C[g][k] / sqrt(pivot)
    if (lane == 0) V.al[j * 8 + k] = ak;
  }
  double i0s = inv[0], i1s = inv[1];
#pragma unroll
  for (int m = 1; m < 4; ++m) {
    i0s = q == m ? inv[2 * m] : i0s;
    i1s = q == m ? inv[2 * m + 1] : i1s;
  }
  *reinterpret_cast<double2*>(V.ljj + g * 8 + 2 * q) = make_double2(2 * q < g ? c0 * i0s : 0.0, 2 * q + 1 < g ? c1 * i1s : 0.0);
  if (lane < 8) {
    double iv = inv[0];
#pragma unroll
    for (int m = 1; m < 8; ++m) iv = lane == m ? inv[m] : iv;
    V.invd[lane] = iv;
  }
  if (lane == 0) {
    red[0] += chi;
    red[1] += 0.5 * log(p8);
  }
}

// One kernel for every epoch count that fits (the tile count per worker is a runtime loop).
__global__ void __launch_bounds__(kGsThreads, RVLP_GPS_MB)
gps_factor_kernel(DevProblem P, int64_t S, GpbWork w, unsigned long long* __restrict__ ticket, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int N = P.n_epochs, NT = gps_tile_rows(N), NP = NT * 8;
  GpsView V;
  V.N = N; V.NT = NT;
  V.Ls = reinterpret_cast<double2*>(smem);                                 // [NT (NT - 1) / 2][32] tile (i, k) at i (i - 1) / 2 + k
  V.ljj = reinterpret_cast<double*>(smem + (size_t)NT * (NT - 1) / 2 * 512);   // [8][8] row-major
  V.invd = V.ljj + 64;                                                     // [8]
  V.dtile = V.invd + 8;                                                    // [2][64 + 32] diagonal-tile hand-over
  V.t = V.dtile + 2 * 96;                                                  // [NP] each
  V.cph = V.t + NP;
  V.sph = V.cph + NP;
  V.dn = V.sph + NP;                                                       // sigma^2 + jit^2 (fit.py:8094-8096)
  V.r = V.dn + NP;
  V.al = V.r + NP;
  double* red_s = V.al + NP;                                               // chi^2, sum ln L_kk
  long long* next_s = reinterpret_cast<long long*>(red_s + 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np_w = gpb_dims(N).np;                                         // row stride of the prologue's arrays
  const double* ep_t = P.epochs;
  const double* ep_e2 = P.epochs + 2 * (size_t)P.n_pad;
  const int* ep_inst = reinterpret_cast<const int*>(P.epochs + 3 * (size_t)P.n_pad);
  for (int i = tid; i < NP; i += kGsThreads) V.t[i] = ep_t[i < N ? i : N - 1];

  for (;;) {
    __syncthreads();                                       // the previous sample is done with shared memory
    if (tid == 0) *next_s = (long long)atomicAdd(ticket, 1ull);
    __syncthreads();
    const int64_t s = *next_s;
    if (s >= S) break;
    const int st = w.status[s];
    if (st != 0) {                                         // CTA-uniform
      if (tid == 0) {
        double r = -INFINITY;                              // fit.py:7857-7886
        if (st == 2) {                                     // non-finite mean model, fit.py:8082-8083
          r = -INFINITY + w.lp[s] + w.lhp[s];
          r += P.jacobian;
          r += P.renorm;
        }
        out[s] = r;
      }
      continue;
    }
    for (int i = tid; i < NP; i += kGsThreads) {
      const bool in = i < N;
      V.cph[i] = in ? w.cph[(size_t)s * np_w + i] : 1.0;
      V.sph[i] = in ? w.sph[(size_t)s * np_w + i] : 0.0;
      V.r[i] = in ? w.resid[(size_t)s * np_w + i] : 0.0;
      V.dn[i] = in ? ep_e2[i] + w.jit2[(size_t)s * P.n_inst + ep_inst[i]] : 0.0;
    }
    if (tid == 0) { red_s[0] = 0.0; red_s[1] = 0.0; }
    V.inv_le = w.hyp[(size_t)s * 4 + 1];
    V.g2 = 0.5 * w.hyp[(size_t)s * 4 + 2];
    V.A2 = w.hyp[(size_t)s * 4 + 3];
    __syncthreads();

#ifdef RVLP_GPS_TRACE
    const bool trace = blockIdx.x == 0 && s < 600;          // the caller over-allocates `out`: stamps go behind the S results
    const long long tr0 = clock64();
#define GPS_STAMP(slot) if (trace && lane == 0) { out[S + ((j * 4 + warp) * 6 + (slot))] = (double)(clock64() - tr0); }
#else
#define GPS_STAMP(slot)
#endif
    if (warp < kGsWorkers) {
      // ------------------------------------------------ worker: tiles (i, j), i > j, i % 4 == warp
      auto first_row = [&](int j) { return j + 1 + (warp + (kGsWorkers - 1) * (j + 1)) % kGsWorkers; };   // first i > j, i % W == warp
      auto ahead = [&](int jn) {
        int i = first_row(jn);
#pragma unroll 1
        for (; i + kGsWorkers < NT; i += 2 * kGsWorkers) gps_ahead_group<2>(V, jn, i, kGsWorkers, lane);
        if (i < NT) gps_ahead_group<1>(V, jn, i, kGsWorkers, lane);
        if (jn % kGsWorkers == warp) gps_ahead_diag(V, jn, lane);
      };
      ahead(0);
      gps_bar_sync(2, kGsThreads);                         // tile (0, 0) is handed over
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        GPS_STAMP(0)
        if (j + 1 < NT) ahead(j + 1);                      // next column's early part, while the diagonal warp factors
        GPS_STAMP(1)
        GPS_STAMP(2)
        gps_bar_sync(1, kGsThreads);                       // L_jj is published
        GPS_STAMP(3)
        int i = first_row(j);
        if (i < NT) {
          GpsLjj L;
          gps_load_ljj(V, lane, L);
#pragma unroll 1
          for (; i + kGsWorkers < NT; i += 2 * kGsWorkers) gps_solve_group<2>(V, L, j, i, kGsWorkers, lane);
          if (i < NT) gps_solve_group<1>(V, L, j, i, kGsWorkers, lane);
        }
        GPS_STAMP(4)
        gps_bar_sync(2, kGsThreads);                       // column j of L is in shared memory
        GPS_STAMP(5)
      }
    } else {
      // ------------------------------------------------ diagonal warp
      gps_bar_sync(2, kGsThreads);                         // tile (0, 0) is handed over
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        GPS_STAMP(0)
        const double* h = V.dtile + (j & 1) * 96;
        const double2 c01 = *reinterpret_cast<const double2*>(h + 2 * lane);
        double part = h[64 + lane];
        double e0 = 0.0, e1 = 0.0;
        if (j > 0) {                                       // last term, k = j - 1
          const double2 b = V.Ls[(size_t)(j * (j - 1) / 2 + j - 1) * 32 + lane];
          dmma884(e0, e1, b.x, b.x);
          dmma884(e0, e1, b.y, b.y);
          const int q = lane & 3;
          part = fma(b.x, V.al[(j - 1) * 8 + q], fma(b.y, V.al[(j - 1) * 8 + 4 + q], part));
        }
        GPS_STAMP(1)
        gps_factor_diag(V, j, lane, c01.x - e0, c01.y - e1, part, red_s);
        __threadfence_block();
        gps_bar_arrive(1, kGsThreads);
        GPS_STAMP(2)
        GPS_STAMP(3)
        GPS_STAMP(4)
        gps_bar_sync(2, kGsThreads);
        GPS_STAMP(5)
      }
    }
    __syncthreads();
    if (tid == 0) {
      const double ll = -0.5 * red_s[0] - red_s[1] - 0.5 * (double)N * kLog2Pi;
      double r = ll + w.lp[s] + w.lhp[s];                   // fit.py:7898-7900
      r += P.jacobian;
      r += P.renorm;
      out[s] = r;
    }
  }
}

}  // namespace rvlp
