// rvlp_gp_smem.cuh — K3 for 40..208 epochs (config 5): one 4-warp CTA per sample in flight, the factor in SHARED MEMORY
// as 8 x 8 tiles in tensor-core fragment order, the whole O(N^3) part on the fp64 tensor cores.
//
// GPLogPosterior.log_probability (/root/reference/src/ravest/fit.py:7836-7901, 8062-8105; kernel gp.py:145-156):
//   C = K(t, t) + diag(sigma^2 + jit^2) = L L^T,  alpha = L^-1 r,  ll = -1/2 alpha.alpha - sum ln L_ii - N/2 ln 2 pi.
//
// Why another kernel.  rvlp_gp_pipe.cuh keeps the triangle in REGISTERS (6 x 6 tiles, DFMA updates): two samples per
// SM, a barrier-bound dependency chain, 25 % of the fp64 peak at N = 120.  rvlp_gp_batch.cuh keeps it in HBM and pays
// ~N^3 / 12 bytes of traffic per sample.  Here the factor (the strictly lower 8 x 8 tiles; 38 KB of shared slots at
// N = 120, see gps_diag_slots) stays in shared memory, FOUR samples per SM, and the O(N^3) part is `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4): one warp
// instruction per 256 FMAs keeps an SM sub-partition's fp64 pipe busy for 16 cycles (latency 26: tools/dmma_lat.cu), so
// the few warps that fit are enough to feed it.  Left-looking by block column j (NT = ceil(N / 8) of them):
//   * a tile of L in "A-fragment order" (lane l holds L[l / 4][l % 4] and L[l / 4][4 + l % 4]) serves as the A operand
//     AND - L^T as the B operand has the same lane map - as the B operand: one conflict-free LDS.128 per tile and k;
//   * warps 0..2 are WORKERS (tile rows dealt mod 3).  A tile's own 512-byte slot carries it through three phases that
//     run ahead of the factorisation front: cov (column j + 2: C_ij generated in accumulator layout, phase-factored
//     periodic term, 19 fp64 instructions per element), gram (column j + 1: slot -= sum_{k < j} L_ik L_j+1,k^T - all
//     but the last term), solve (column j: last term, X = (C - G) L_jj^-T as two more DMMA against the published
//     INVERSE of the diagonal tile, slot = X in fragment order: two STS.64 per lane, no shuffles);
//   * warp 3 is the DIAGONAL WARP, alone on its SM sub-partition: 8 x 8 Cholesky of tile (j, j) in accumulator layout
//     (quad shuffles; pivot reciprocal = MUFU seed + 3 fp64 operations on the chain), its inverse built along by forward
//     substitution on the identity; it publishes W = L_jj^-1 and `bar.arrive`s; then, off the critical path,
//     alpha_j = W (r_j - sum_k L_jk alpha_k), chi^2, sum ln L_kk and the covariance entries of tile (j + 1, j + 1) (its
//     Gram sum comes from the worker that owns row j + 1, through a double-buffered hand-over block);
//   * two named barriers per block column: "W is published" (arrive / sync), "column j is stored".
// Per sample at N = 120: 1328 DMMA instead of 23.7 k DFMA warp instructions; 0.71 ms per 1e4 samples (1.44 pipelined).
// The mean model / priors / reject flags come from gpb_prologue_kernel (one warp per sample, rvlp_gp_batch.cuh).
// Deterministic: fixed summation order per tile; the ticket only decides WHICH CTA takes a sample - out[s] depends on
// (theta[s], epochs) only.  Phase stamps for tuning: -DRVLP_GPS_TRACE + tools/gp_smem_trace.py.
#pragma once
#include "rvlp_gp_batch.cuh"

namespace rvlp {

#ifndef RVLP_GPS_WORKERS
#define RVLP_GPS_WORKERS 3
#endif
constexpr int kGsWorkers = RVLP_GPS_WORKERS;        // worker warps: worker w owns the tile rows i with i % kGsWorkers == w
constexpr int kGsThreads = 32 * (kGsWorkers + 1);   // + the diagonal warp (with 3 workers: an SM sub-partition of its own)
#ifndef RVLP_GPS_MB
#define RVLP_GPS_MB 4
#endif

__host__ __device__ inline int gps_tile_rows(int N) { return (N + 7) / 8; }
// Tile slots.  Tile (i, k), i > k, is written first by `cov` during column k - 2 and read last during column i (as the
// pivot row's B operand), so along the sub-diagonal d = i - k the tiles k, k + d + 3, k + 2 (d + 3), ... never live at the
// same time and share a slot: sub-diagonal d needs min(d + 3, NT - d) slots instead of NT - d.  75 slots instead of 105
// at N = 120 (38 KB): FOUR samples per SM instead of three.  slot(i, k) = base[d] + k mod min(d + 3, NT - d), through a
// small table in shared memory (gps_fill_slot_table).
__host__ __device__ inline int gps_diag_slots(int nt, int d) { return d + 3 < nt - d ? d + 3 : nt - d; }
__host__ __device__ inline int gps_slot_count(int nt) {
  int n = 0;
  for (int d = 1; d < nt; ++d) n += gps_diag_slots(nt, d);
  return n;
}
__host__ __device__ inline int gps_tab_bytes(int nt) { return (nt * nt * 2 + 15) & ~15; }
// pred (K7, the conditioning path): beta = L^-T alpha needs the WHOLE factor at the end - every tile keeps its own slot
// and the inverse of every diagonal tile is kept as well.
__host__ __device__ inline int gps_tile_slots(int nt, bool pred) { return pred ? nt * (nt - 1) / 2 : gps_slot_count(nt); }
__host__ __device__ inline int gps_smem_bytes(int N, bool pred = false) {
  const int nt = gps_tile_rows(N), np = nt * 8;
  return gps_tile_slots(nt, pred) * 512 + gps_tab_bytes(nt) + ((pred ? nt : 1) * 64 + 8) * 8 + 2 * (64 + 32) * 8 + 6 * np * 8 + 64;
}

__device__ __forceinline__ void gps_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void gps_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Per-sample shared-memory views and constants of the block-column steps.
struct GpsView {
  double2* Ls;
  const unsigned short* tab;                               // [NT][NT] slot of tile (i, k)
  double *ljj, *invd, *dtile, *t, *cph, *sph, *dn, *r, *al;
  double inv_le, g2, A2;
  int N, NT;
};

__device__ __forceinline__ double2* gps_tile(const GpsView& V, int i, int k) {
  return V.Ls + (size_t)V.tab[i * V.NT + k] * 32;
}

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

// ---- workers.  A tile's own 512-byte slot carries it through its three phases, so every phase is a runtime loop over
// small GROUPS of G tiles (rows i0, i0 + step, ...) with no register state in between - one copy of the code instead of
// one per tile count (a per-count unrolled version stalled on instruction fetch: ncu no_instruction 1.3 per issue):
//   cov   (two columns early)  slot = C_ij, generated in accumulator layout (lane l: doubles 2 l, 2 l + 1);
//   gram  (one column early)   slot -= sum_{k < j - 1} L_ik L_jk^T - everything but the last term, while the diagonal
//                              warp still factors tile (j - 1, j - 1); four accumulators cut the DMMA dependency chain;
//   solve (L_jj^-1 published)  last term k = j - 1, X = (C - G) L_jj^-T as two more DMMA against the published inverse,
//                              slot = X in fragment order.
template <int G>
__device__ __forceinline__ void gps_cov_group(const GpsView& V, int jc, int i0, int step, int lane) {
  const int g = lane >> 2, q = lane & 3;
  double v[G][2];
#pragma unroll
  for (int t = 0; t < G; ++t) gps_cov2(V, i0 + step * t, jc, g, q, v[t][0], v[t][1]);
#pragma unroll
  for (int t = 0; t < G; ++t) {
    const int i = i0 + step * t;
    gps_tile(V, i, jc)[lane] = make_double2(v[t][0], v[t][1]);
  }
}

template <int G>
__device__ __forceinline__ void gps_gram_group(const GpsView& V, int jn, int i0, int step, int lane) {
  double acc[G][2][2];                                     // two accumulators per tile: DMMA latency 26, issue 16 cycles
  const unsigned short* ti[G];
#pragma unroll
  for (int t = 0; t < G; ++t) {
#pragma unroll
    for (int h = 0; h < 2; ++h) acc[t][h][0] = acc[t][h][1] = 0.0;
    ti[t] = V.tab + (i0 + step * t) * V.NT;
  }
  const unsigned short* tj = V.tab + jn * V.NT;
  const double2* L = V.Ls + lane;
  const int nk = jn - 1;
  int k = 0;
#pragma unroll 1
  for (; k + 1 < nk; k += 2) {
    const double2 b0 = L[tj[k] * 32], b1 = L[tj[k + 1] * 32];
    double2 a0[G], a1[G];
#pragma unroll
    for (int t = 0; t < G; ++t) { a0[t] = L[ti[t][k] * 32]; a1[t] = L[ti[t][k + 1] * 32]; }
#pragma unroll
    for (int t = 0; t < G; ++t) dmma884(acc[t][0][0], acc[t][0][1], a0[t].x, b0.x);
#pragma unroll
    for (int t = 0; t < G; ++t) dmma884(acc[t][1][0], acc[t][1][1], a0[t].y, b0.y);
#pragma unroll
    for (int t = 0; t < G; ++t) dmma884(acc[t][0][0], acc[t][0][1], a1[t].x, b1.x);
#pragma unroll
    for (int t = 0; t < G; ++t) dmma884(acc[t][1][0], acc[t][1][1], a1[t].y, b1.y);
  }
  if (k < nk) {
    const double2 b0 = L[tj[k] * 32];
#pragma unroll
    for (int t = 0; t < G; ++t) {
      const double2 a0 = L[ti[t][k] * 32];
      dmma884(acc[t][0][0], acc[t][0][1], a0.x, b0.x);
      dmma884(acc[t][1][0], acc[t][1][1], a0.y, b0.y);
    }
  }
#pragma unroll
  for (int t = 0; t < G; ++t) {
    double2* slot = V.Ls + ti[t][jn] * 32 + lane;
    double2 p = *slot;
    p.x -= acc[t][0][0] + acc[t][1][0];
    p.y -= acc[t][0][1] + acc[t][1][1];
    *slot = p;
  }
}

// The worker that owns tile ROW jn also prepares the diagonal tile's Gram sum for the diagonal warp (its operands are
// in this worker's hands anyway): sum_{k < jn - 1} L_jn,k L_jn,k^T (accumulator layout) and the matching part of the
// residual row's sum_k L_jn,k alpha_k go to the double-buffered hand-over block V.dtile.  (The covariance entries of
// the diagonal tile are the diagonal warp's own job, in its idle time: gps_diag_cov.)
__device__ __forceinline__ void gps_ahead_diag(const GpsView& V, int jn, int lane) {
  const int q = lane & 3;
  double dg[2] = {0.0, 0.0}, eg[2] = {0.0, 0.0}, part = 0.0;
  const unsigned short* tj = V.tab + jn * V.NT;
  const double* al = V.al + q;
#pragma unroll 2
  for (int k = 0; k < jn - 1; ++k) {
    const double2 b = V.Ls[tj[k] * 32 + lane];
    dmma884(dg[0], dg[1], b.x, b.x);
    dmma884(eg[0], eg[1], b.y, b.y);
    part = fma(b.x, al[k * 8], fma(b.y, al[k * 8 + 4], part));
  }
  double* h = V.dtile + (jn & 1) * 96;
  *reinterpret_cast<double2*>(h + 2 * lane) = make_double2(dg[0] + eg[0], dg[1] + eg[1]);
  h[64 + lane] = part;
}
__device__ __forceinline__ void gps_diag_cov(const GpsView& V, int jn, int lane, double& d0, double& d1) {
  const int g = lane >> 2, q = lane & 3;
  gps_cov2(V, jn, jn, g, q, d0, d1);
  const int row = jn * 8 + g, col = jn * 8 + 2 * q;
  if (row < V.N) {                                         // white noise on the diagonal (fit.py:8094-8096)
    const double dn = V.dn[row];
    d0 += row == col ? dn : 0.0;
    d1 += row == col + 1 ? dn : 0.0;
  }
}

template <int G>
__device__ __forceinline__ void gps_solve_group(const GpsView& V, double2 w, int j, int i0, int step, int lane) {
  const int g = lane >> 2, q = lane & 3, quad = lane & ~3;
  double2* slot[G];
  double c[G][2];
#pragma unroll
  for (int t = 0; t < G; ++t) {
    const int i = i0 + step * t;
    slot[t] = gps_tile(V, i, j);
    const double2 p = slot[t][lane];
    c[t][0] = p.x;
    c[t][1] = p.y;
  }
  if (j > 0) {
    const double2 b = gps_tile(V, j, j - 1)[lane];
    double e[G][2], f[G][2];
#pragma unroll
    for (int t = 0; t < G; ++t) {
      const double2 a = gps_tile(V, i0 + step * t, j - 1)[lane];
      e[t][0] = e[t][1] = f[t][0] = f[t][1] = 0.0;
      dmma884(e[t][0], e[t][1], a.x, b.x);
      dmma884(f[t][0], f[t][1], a.y, b.y);
    }
#pragma unroll
    for (int t = 0; t < G; ++t) {
      c[t][0] -= e[t][0] + f[t][0];
      c[t][1] -= e[t][1] + f[t][1];
    }
  }
  // accumulator layout -> A-fragment layout inside the quad: P[g][q] and P[g][4 + q]
  const int s0 = quad | (q >> 1), s1 = s0 | 2;
  const int e0 = 2 * (4 * g + ((2 * q) & 3)) + ((2 * q) >> 2), e1 = 2 * (4 * g + ((2 * q + 1) & 3)) + ((2 * q + 1) >> 2);
#pragma unroll
  for (int t = 0; t < G; ++t) {
    const double u0 = __shfl_sync(0xffffffffu, c[t][0], s0), u1 = __shfl_sync(0xffffffffu, c[t][1], s0);
    const double v0 = __shfl_sync(0xffffffffu, c[t][0], s1), v1 = __shfl_sync(0xffffffffu, c[t][1], s1);
    const double ax = (q & 1) ? u1 : u0, ay = (q & 1) ? v1 : v0;
    double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
    dmma884(x0, x1, ax, w.x);                              // X = P W^T, W = L_jj^-1 in fragment order
    dmma884(y0, y1, ay, w.y);
    double* tile = reinterpret_cast<double*>(slot[t]);
    tile[e0] = x0 + y0;
    tile[e1] = x1 + y1;
  }
}

// ---- diagonal warp: 8 x 8 Cholesky of tile (j, j) in accumulator layout with the INVERSE of the factor built along
// (forward substitution on the identity, one step behind the pivots); publishes W = L_jj^-1 in fragment order.  The
// pivot chain is shuffle -> reciprocal (seed + 3 fp64 operations) -> two multiplies -> shuffle; the reciprocal square
// roots that scale L and W are off it.  Returns W (accumulator layout) and the product of the pivots.
__device__ __forceinline__ void gps_factor_diag(double* wbuf, int lane, double c0, double c1, double& w0, double& w1, double& p8) {
  const int g = lane >> 2, q = lane & 3, quad = lane & ~3;
  double z0 = g == 2 * q ? 1.0 : 0.0, z1 = g == 2 * q + 1 ? 1.0 : 0.0;
  p8 = 1.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double mine = (k & 1) ? c1 : c0;                 // this lane's entry of column pair k / 2
    const double low = g > k ? mine : 0.0;                 // rows <= k take no part in the update of the trailing block
    const double piv = __shfl_sync(0xffffffffu, mine, k * 4 + (k >> 1));
    const double cgk = __shfl_sync(0xffffffffu, low, quad | (k >> 1));             // C[g][k], g > k
    const double ck0 = __shfl_sync(0xffffffffu, low, (2 * q) * 4 + (k >> 1));      // C[2q][k], 2q > k
    const double ck1 = __shfl_sync(0xffffffffu, low, (2 * q + 1) * 4 + (k >> 1));  // C[2q+1][k], 2q+1 > k
    const double zk0 = __shfl_sync(0xffffffffu, z0, k * 4 + q), zk1 = __shfl_sync(0xffffffffu, z1, k * 4 + q);
    const double sg = cgk * rcp64_3(piv);
    c0 = fma(-sg, ck0, c0);
    c1 = fma(-sg, ck1, c1);
    const double inv = pivot_rsqrt(piv);
    p8 = piv > 0.0 ? p8 * piv : __longlong_as_double(0x7ff8000000000000ll);        // not positive definite -> NaN (as jax)
    const double wk0 = zk0 * inv, wk1 = zk1 * inv;         // row k of W
    const double lgk = cgk * inv;                          // L[g][k], 0 for g <= k
    z0 = g == k ? wk0 : fma(-lgk, wk0, z0);
    z1 = g == k ? wk1 : fma(-lgk, wk1, z1);
  }
  w0 = z0;
  w1 = z1;
  const int e0 = 2 * (4 * g + ((2 * q) & 3)) + ((2 * q) >> 2), e1 = 2 * (4 * g + ((2 * q + 1) & 3)) + ((2 * q + 1) >> 2);
  wbuf[e0] = z0;
  wbuf[e1] = z1;
}

// One kernel for every epoch count that fits (the tile count per worker is a runtime loop).
// PRED (K7, /root/reference/src/ravest/fit.py:7494-7554, 5386-5429): residual without priors (the prologue's <true>
// flavour), every tile and every diagonal inverse kept, then beta = L^-T alpha by a right-looking blocked back
// substitution on the fragment-order tiles; out = chi^2 = alpha.alpha (may be null), beta_out [S, N].
template <bool PRED>
__global__ void __launch_bounds__(kGsThreads, PRED ? 3 : RVLP_GPS_MB)
gps_factor_kernel(DevProblem P, int64_t S, GpbWork w, unsigned long long* __restrict__ ticket, double* __restrict__ out,
                  double* __restrict__ beta_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int N = P.n_epochs, NT = gps_tile_rows(N), NP = NT * 8;
  GpsView V;
  V.N = N; V.NT = NT;
  const int n_slots = gps_tile_slots(NT, PRED);
  V.Ls = reinterpret_cast<double2*>(smem);                                 // [n_slots][32] tile (i, k) at slot tab[i][k]
  unsigned short* tab_w = reinterpret_cast<unsigned short*>(smem + (size_t)n_slots * 512);
  V.tab = tab_w;
  V.ljj = reinterpret_cast<double*>(smem + (size_t)n_slots * 512 + gps_tab_bytes(NT));   // W = L_jj^-1, fragment order (PRED: [NT] of them)
  V.invd = V.ljj + (PRED ? NT : 1) * 64;                                   // [8] (spare)
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
  for (int idx = tid; idx < NT * NT; idx += kGsThreads) {                  // the slot table (see gps_diag_slots)
    const int i = idx / NT, k = idx - i * NT;
    int slot = 0;
    if (i > k) {
      if (PRED) {
        slot = i * (i - 1) / 2 + k;
      } else {
        const int d = i - k;
        for (int e = 1; e < d; ++e) slot += gps_diag_slots(NT, e);
        slot += k % gps_diag_slots(NT, d);
      }
    }
    tab_w[idx] = (unsigned short)slot;
  }

  for (;;) {
    __syncthreads();                                       // the previous sample is done with shared memory
    if (tid == 0) *next_s = (long long)atomicAdd(ticket, 1ull);
    __syncthreads();
    const int64_t s = *next_s;
    if (s >= S) break;
    const int st = w.status[s];
    if (st != 0) {                                         // CTA-uniform
      if (PRED) {                                          // the reference raises: NaN rows
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        for (int i = tid; i < N; i += kGsThreads) beta_out[(size_t)s * N + i] = qnan;
        if (tid == 0 && out) out[s] = qnan;
      } else if (tid == 0) {
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
      auto cov_col = [&](int jc) {
        int i = first_row(jc);
#pragma unroll 1
        for (; i + kGsWorkers < NT; i += 2 * kGsWorkers) gps_cov_group<2>(V, jc, i, kGsWorkers, lane);
        if (i < NT) gps_cov_group<1>(V, jc, i, kGsWorkers, lane);
      };
      cov_col(0);
      if (warp == 0) gps_ahead_diag(V, 0, lane);
      if (NT > 1) cov_col(1);
      gps_bar_sync(2, kGsThreads);                         // tile (0, 0) is handed over
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        GPS_STAMP(0)
        if (j + 1 < NT) {                                  // next column's early part, while the diagonal warp factors
          if ((j + 1) % kGsWorkers == warp) gps_ahead_diag(V, j + 1, lane);
          int i = first_row(j + 1);
#pragma unroll 1
          for (; i + kGsWorkers < NT; i += 2 * kGsWorkers) gps_gram_group<2>(V, j + 1, i, kGsWorkers, lane);
          if (i < NT) gps_gram_group<1>(V, j + 1, i, kGsWorkers, lane);
        }
        GPS_STAMP(1)
        if (j + 2 < NT) cov_col(j + 2);
        GPS_STAMP(2)
        gps_bar_sync(1, kGsThreads);                       // W = L_jj^-1 is published
        GPS_STAMP(3)
        int i = first_row(j);
        if (i < NT) {
          const double2 wf = *reinterpret_cast<const double2*>(V.ljj + (PRED ? j * 64 : 0) + 2 * lane);
#pragma unroll 1
          for (; i + kGsWorkers < NT; i += 2 * kGsWorkers) gps_solve_group<2>(V, wf, j, i, kGsWorkers, lane);
          if (i < NT) gps_solve_group<1>(V, wf, j, i, kGsWorkers, lane);
        }
        GPS_STAMP(4)
        gps_bar_sync(2, kGsThreads);                       // column j of L is in shared memory
        GPS_STAMP(5)
      }
    } else {
      // ------------------------------------------------ diagonal warp
      const int g = lane >> 2, q = lane & 3;
      double chi = 0.0, logdet = 0.0;
      double dc0, dc1;
      gps_diag_cov(V, 0, lane, dc0, dc1);
      gps_bar_sync(2, kGsThreads);                         // tile (0, 0) is handed over
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        GPS_STAMP(0)
        const double* h = V.dtile + (j & 1) * 96;
        double2 c01 = *reinterpret_cast<const double2*>(h + 2 * lane);
        c01.x = dc0 - c01.x;
        c01.y = dc1 - c01.y;
        double part = h[64 + lane];
        double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
        if (j > 0) {                                       // last term, k = j - 1
          const double2 b = gps_tile(V, j, j - 1)[lane];
          dmma884(e0, e1, b.x, b.x);
          dmma884(f0, f1, b.y, b.y);
          part = fma(b.x, V.al[(j - 1) * 8 + q], fma(b.y, V.al[(j - 1) * 8 + 4 + q], part));
        }
        GPS_STAMP(1)
        double w0, w1, p8;
        gps_factor_diag(V.ljj + (PRED ? j * 64 : 0), lane, c01.x - (e0 + f0), c01.y - (e1 + f1), w0, w1, p8);
        __threadfence_block();
        gps_bar_arrive(1, kGsThreads);
        GPS_STAMP(2)
        // off the critical path: alpha_j = W (r_j - sum_k L_jk alpha_k), chi^2, sum ln L_kk
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        const double zg = V.r[j * 8 + g] - part;           // entry g of the right-hand side, in every lane of quad g
        const double za = __shfl_sync(0xffffffffu, zg, (2 * q) * 4), zb = __shfl_sync(0xffffffffu, zg, (2 * q + 1) * 4);
        double ag = fma(w0, za, w1 * zb);
        ag += __shfl_xor_sync(0xffffffffu, ag, 1);
        ag += __shfl_xor_sync(0xffffffffu, ag, 2);          // alpha[8 j + g]
        if (q == 0) V.al[j * 8 + g] = ag;
        double a2 = q == 0 ? ag * ag : 0.0;
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        chi += a2;
        logdet += 0.5 * log(p8);
        if (j + 1 < NT) gps_diag_cov(V, j + 1, lane, dc0, dc1);
        GPS_STAMP(3)
        GPS_STAMP(4)
        gps_bar_sync(2, kGsThreads);
        GPS_STAMP(5)
      }
      if (lane == 0) { red_s[0] = chi; red_s[1] = logdet; }
    }
    __syncthreads();
    if (PRED) {
      if (tid == 0 && out) out[s] = red_s[0];              // chi^2 = r^T C^-1 r (fit.py:5428-5429)
      // beta = L^-T alpha, right-looking: beta_i = W_i^T z_i by one warp, then every warp folds beta_i into the z_j of
      // its tiles (i, j), j < i (fragment order: lane (g, q) holds L[g][q], L[g][4 + q]; the sums over g are three
      // xor-shuffles).  z lives in V.al and becomes beta in place.
      const int g = lane >> 2, q = lane & 3;
#pragma unroll 1
      for (int i = NT - 1; i >= 0; --i) {
        if (warp == kGsWorkers) {
          const double2 wv = *reinterpret_cast<const double2*>(V.ljj + i * 64 + 2 * lane);
          const double zg = V.al[i * 8 + g];
          double p0 = wv.x * zg, p1 = wv.y * zg;
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
          }
          __syncwarp();                                    // every lane has read z_i before it is overwritten
          if (g == 0) {
            V.al[i * 8 + q] = p0;
            V.al[i * 8 + 4 + q] = p1;
            if (i * 8 + q < N) beta_out[(size_t)s * N + i * 8 + q] = p0;
            if (i * 8 + 4 + q < N) beta_out[(size_t)s * N + i * 8 + 4 + q] = p1;
          }
        }
        __syncthreads();
        const double bg = V.al[i * 8 + g];
        for (int j = warp; j < i; j += kGsWorkers + 1) {
          const double2 a = gps_tile(V, i, j)[lane];
          double p0 = a.x * bg, p1 = a.y * bg;
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            p0 += __shfl_xor_sync(0xffffffffu, p0, o);
            p1 += __shfl_xor_sync(0xffffffffu, p1, o);
          }
          if (g == 0) {
            V.al[j * 8 + q] -= p0;
            V.al[j * 8 + 4 + q] -= p1;
          }
        }
        __syncthreads();
      }
    } else if (tid == 0) {
      const double ll = -0.5 * red_s[0] - red_s[1] - 0.5 * (double)N * kLog2Pi;
      double r = ll + w.lp[s] + w.lhp[s];                   // fit.py:7898-7900
      r += P.jacobian;
      r += P.renorm;
      out[s] = r;
    }
  }
}

}  // namespace rvlp
