// rvlp_bands.cuh — K6: column-wise percentiles of an [S, T] matrix (row f-2).
//
// Replaces `np.percentile(rv_matrix, [15.85, 50, 84.15], axis=0)` at
// /root/reference/src/ravest/fit.py:2239-2240, 2493-2495 (and the GP twins fit.py:6440-6450): the step that
// follows the per-sample RV matrices of K2, so the S x T matrix never leaves the device and 3 x T comes back.
//
// numpy's default ("linear", Hyndman & Fan 7) needs, per column and percentile, the order statistics
// floor(v) and floor(v)+1 of the virtual index v = q (S-1) and blends them (numpy _lerp).  Exact selection,
// no sort: an MSB-first radix select on the order-preserving 64-bit image of the doubles, 8 bits per level,
// all columns and all targets at once.  One launch per level; launch L
//   1. scans level L-1's histograms and extends every target's key prefix to L digits;
//   2. streams its slab of rows ONCE (coalesced: a CTA owns kColBlock adjacent columns) and counts digit L of
//      every element that still matches a live prefix into shared-memory histograms (run-length aggregated
//      per thread: neighbouring samples of a column mostly share their leading digits), OR - as soon as a
//      column's live elements fit its candidate buffer (typically at L = 3) - copies them out instead;
//   3. merges the histograms into the level's global histogram with integer atomics.
// Columns whose candidates were collected stop streaming; the finishing kernel runs their remaining levels on
// the (<= kCandCap) candidates in shared memory.  Either way the selected keys ARE the order statistics
// (bit-exact); the finishing kernel blends and stores.  Integer counting only: the result does not depend on
// grid size, row split or arrival order.  HBM-bound: passes x S x T x 8 bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ravest_b200.h"

namespace rvlp {

constexpr int kColBlock = 8;              // adjacent columns per CTA: 64-byte row segments
constexpr int kBandThreads = 256;
constexpr int kBandWarps = kBandThreads / 32;
constexpr int kRowsPerIter = kBandThreads / kColBlock;
constexpr int kMaxTargets = 2 * RVLP_MAX_PERCENTILES;
constexpr int kLevels = 8;
constexpr int kCandCap = 1024;            // candidate doubles per column
constexpr int kModeFastDone = 1000;       // W.mode value: column already finished by the two-pass path (rvlp_bands_fast.cuh)
#ifndef RVLP_BAND_UNROLL
#define RVLP_BAND_UNROLL 8
#endif
#ifndef RVLP_BAND_UNROLL0      // level 0 is pure streaming + a run-length counter: more loads in flight
#define RVLP_BAND_UNROLL0 8
#endif

struct BandTargets {                      // host-resolved numpy index arithmetic (see rvlp_capi.cu)
  int n_q;
  uint32_t k[kMaxTargets];                // k[2q] = previous index, k[2q+1] = next index (0-based ranks)
  double gamma[RVLP_MAX_PERCENTILES];
};

// Workspace (per column c in [0, T)), R = 2 n_q targets:
//   hist   [3][T][R][256] uint32   level L fills L%3, launch L+1 reads it, launch L clears (L+1)%3
//   prefix [9][T][R]      uint64   prefix[L] = the L leading digits of each target's key
//   rank   [9][T][R]      uint32   remaining 0-based rank among the elements matching prefix[L]
//   mode   [T]            int32    -1 streaming; L >= 1: live elements were collected by launch L
//   ncand  [T]            uint32   candidates stored
//   nan    [T]            uint32   column holds a NaN -> numpy returns NaN for it
//   cand   [T][kCandCap]  double
struct BandWorkspace {
  uint32_t* hist;
  uint64_t* prefix;
  uint32_t* rank;
  int32_t* mode;
  uint32_t* ncand;
  uint32_t* nan;
  double* cand;
};
__host__ __device__ inline size_t band_align(size_t b) { return (b + 255) & ~(size_t)255; }
__host__ __device__ inline size_t band_ws_bytes(int64_t T, int R) {
  size_t b = 0;
  b += band_align((size_t)3 * T * R * 256 * 4);
  b += band_align((size_t)(kLevels + 1) * T * R * 8);
  b += band_align((size_t)(kLevels + 1) * T * R * 4);
  b += band_align((size_t)3 * T * 4);
  b += band_align((size_t)T * kCandCap * 8);
  return b;
}
__host__ __device__ inline BandWorkspace band_ws_carve(void* base, int64_t T, int R) {
  BandWorkspace W;
  unsigned char* p = reinterpret_cast<unsigned char*>(base);
  W.hist = reinterpret_cast<uint32_t*>(p); p += band_align((size_t)3 * T * R * 256 * 4);
  W.prefix = reinterpret_cast<uint64_t*>(p); p += band_align((size_t)(kLevels + 1) * T * R * 8);
  W.rank = reinterpret_cast<uint32_t*>(p); p += band_align((size_t)(kLevels + 1) * T * R * 4);
  W.mode = reinterpret_cast<int32_t*>(p);
  W.ncand = reinterpret_cast<uint32_t*>(p) + T;           // ncand | nan are cleared with one memset
  W.nan = reinterpret_cast<uint32_t*>(p) + 2 * T;
  p += band_align((size_t)3 * T * 4);
  W.cand = reinterpret_cast<double*>(p);
  return W;
}

// order-preserving map double -> uint64 (total order: -NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN)
__device__ __forceinline__ uint64_t key_of(double x) {
  const uint64_t b = (uint64_t)__double_as_longlong(x);
  return b ^ ((uint64_t)((int64_t)b >> 63) | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(uint64_t k) {
  const uint64_t b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
  return __longlong_as_double((long long)b);
}
// the L leading digits of a key (L in 0..8)
__device__ __forceinline__ uint64_t key_head(uint64_t k, int L) { return L == 0 ? 0ull : k >> (64 - 8 * L); }

// numpy/lib/_function_base_impl.py `_lerp`: a + (b - a) t, replaced by b - (b - a)(1 - t) where t >= 0.5;
// every operation rounded separately (no FMA), as numpy's ufunc loops do.
__device__ __forceinline__ double numpy_lerp(double a, double b, double t) {
  const double diff = __dsub_rn(b, a);
  if (t >= 0.5) return __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
  return __dadd_rn(a, __dmul_rn(diff, t));
}

// Given a 256-bin histogram spread over a warp (lane holds bins 8 lane .. 8 lane + 7), find the bin holding
// 0-based rank rk: returns the digit, the rank inside the bin and the bin's count (all lanes get the result).
__device__ __forceinline__ void warp_pick_bin(const uint32_t (&cnt)[8], uint32_t rk, int lane, int& digit,
                                              uint32_t& newrank, uint32_t& bincount) {
  uint32_t mine = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) mine += cnt[q];
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const uint32_t excl = incl - mine;
  const bool here = rk >= excl && rk < incl;               // exactly one lane: the counts sum to > rk
  int d = 0;
  uint32_t nr = 0, bc = 0;
  if (here) {
    uint32_t run = excl;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (rk >= run && rk < run + cnt[q]) { d = lane * 8 + q; nr = rk - run; bc = cnt[q]; }
      run += cnt[q];
    }
  }
  const unsigned who = __ballot_sync(0xffffffffu, here);
  const int src = who ? __ffs(who) - 1 : 0;
  digit = __shfl_sync(0xffffffffu, d, src);
  newrank = __shfl_sync(0xffffffffu, nr, src);
  bincount = __shfl_sync(0xffffffffu, bc, src);
}

// smem of the level kernel: hist[kColBlock][R][256] | prefix[Cb][R] | uniq prefix[Cb][R] | rank[Cb][R] |
// bin count[Cb][R] | uniq slot[Cb][R] | n_uniq[Cb] | mode[Cb]
// Level 0 has a single (empty) prefix per column, so it needs one histogram per column, not R.
__host__ __device__ inline int band_hist_slots(int level, int R) { return level == 0 ? 1 : R; }
// words between the histograms of adjacent columns: +1 so that the same digit in the 8 columns of a CTA
// falls into 8 different banks (ncu: 41 M bank conflicts per pass without it)
__host__ __device__ inline int band_col_stride(int level, int R) { return band_hist_slots(level, R) * 256 + 1; }
__host__ __device__ inline int band_smem_bytes(int level, int R) {
  return ((kColBlock * band_col_stride(level, R) * 4 + 15) & ~15) + kColBlock * R * (8 + 8 + 4 + 4 + 4) + kColBlock * 8 + 16;
}
__host__ __device__ constexpr int band_mode_of(int level) { return level == 0 ? 0 : (level <= 3 ? 1 : 2); }

// MODE 0: level 0 (every element counts, one histogram per column, digits cluster -> run-length aggregation);
// MODE 1: levels 1-3 (prefix and digit live in the key's high word: 32-bit classification);
// MODE 2: levels 4-8 (64-bit; level 8 only completes the prefixes).  Rarely streams: columns are normally
//         collected by level 3.
#ifndef RVLP_BAND_MINB0
#define RVLP_BAND_MINB0 4
#endif
#ifndef RVLP_BAND_MINB1
#define RVLP_BAND_MINB1 3      // levels 1-3: 8 loads in flight per thread (80 registers) beat a 4th resident CTA
#endif
// (the body takes the row-slab index / count as arguments: band_fallback_kernel runs all levels in one CTA per column block)
template <int MODE>
__device__ __forceinline__ void band_level_body(const double* __restrict__ A, int64_t S, int64_t T, int level, const BandTargets& tg,
                                                const BandWorkspace& W, const int slab, const int n_slabs) {
  extern __shared__ __align__(16) unsigned char bsm[];
  const int R = 2 * tg.n_q;
  const int HS = band_hist_slots(level, R);
  const int CS = band_col_stride(level, R);
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(bsm);
  uint64_t* prefix_s = reinterpret_cast<uint64_t*>(bsm + ((kColBlock * CS * 4 + 15) & ~15));
  uint64_t* uprefix_s = prefix_s + kColBlock * R;
  uint32_t* rank_s = reinterpret_cast<uint32_t*>(uprefix_s + kColBlock * R);
  uint32_t* binc_s = rank_s + kColBlock * R;               // elements matching the target's new prefix
  int* uslot_s = reinterpret_cast<int*>(binc_s + kColBlock * R);
  int* nuniq_s = uslot_s + kColBlock * R;
  int* mode_s = nuniq_s + kColBlock;                       // -1 stream + count, -2 nothing to do, L collect now

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * kColBlock;
  const int ncol = (int)min((int64_t)kColBlock, T - c0);
  const bool writer = slab == 0;

  // columns collected by an EARLIER launch are finished as far as this kernel is concerned (a value equal to
  // `level` can only be this launch's own writer CTA racing ahead: recompute, the decision is the same)
  if (tid < kColBlock) {
    int m = -2;
    if (tid < ncol) {
      const int g = W.mode[c0 + tid];
      m = (g >= 0 && (g < level || g == kModeFastDone)) ? -2 : -1;   // collected earlier, or finished by rvlp_bands_fast.cuh
    }
    mode_s[tid] = m;
  }
  __syncthreads();
  {
    bool any = false;
    for (int c = 0; c < ncol; ++c) any |= mode_s[c] == -1;
    if (!any) return;                                      // CTA-uniform
  }

  // ---- 1. prefix[level] from prefix[level-1] + hist[level-1]
  if (level == 0) {
    for (int i = tid; i < kColBlock * R; i += kBandThreads) {
      prefix_s[i] = 0;
      rank_s[i] = tg.k[i % R];
      binc_s[i] = (uint32_t)S;
    }
  } else {
    const uint32_t* hprev = W.hist + (size_t)((level - 1) % 3) * T * R * 256;
    const uint64_t* pprev = W.prefix + (size_t)(level - 1) * T * R;
    const uint32_t* rprev = W.rank + (size_t)(level - 1) * T * R;
    for (int i = warp; i < ncol * R; i += kBandWarps) {
      const int c = i / R, r = i - c * R;
      if (mode_s[c] != -1) continue;                       // warp-uniform
      const size_t g = (size_t)(c0 + c) * R;
      const uint64_t pre = pprev[g + r];
      int slot = r;
      for (int q = 0; q < r; ++q)                          // first target with the same prefix owns the histogram
        if (pprev[g + q] == pre) { slot = q; break; }
      const uint32_t* h = hprev + (g + slot) * 256 + lane * 8;
      const uint4 v0 = *reinterpret_cast<const uint4*>(h), v1 = *reinterpret_cast<const uint4*>(h + 4);
      const uint32_t cnt[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      int digit;
      uint32_t newrank, bincount;
      warp_pick_bin(cnt, rprev[g + r], lane, digit, newrank, bincount);
      if (lane == 0) {
        prefix_s[c * R + r] = (pre << 8) | (uint64_t)digit;
        rank_s[c * R + r] = newrank;
        binc_s[c * R + r] = bincount;
      }
    }
  }
  __syncthreads();
  // unique prefixes per column (targets that agree so far share one histogram); collect-now decision
  if (tid < ncol && mode_s[tid] == -1) {
    int n = 0;
    uint32_t live = 0;
    for (int r = 0; r < R; ++r) {
      bool first = true;
      for (int q = 0; q < r; ++q)
        if (prefix_s[tid * R + q] == prefix_s[tid * R + r]) { first = false; break; }
      if (first) {
        uprefix_s[tid * R + n] = prefix_s[tid * R + r];
        uslot_s[tid * R + n] = r;
        live += binc_s[tid * R + r];
        ++n;
      }
    }
    nuniq_s[tid] = n;
    if (level >= 1 && live <= (uint32_t)kCandCap) mode_s[tid] = level;
  }
  for (int i = tid; i < kColBlock * CS; i += kBandThreads) hist_s[i] = 0;
  __syncthreads();
  if (writer) {
    uint64_t* pout = W.prefix + (size_t)level * T * R;
    uint32_t* rout = W.rank + (size_t)level * T * R;
    for (int i = tid; i < ncol * R; i += kBandThreads) {
      if (mode_s[i / R] == -2) continue;
      pout[(size_t)c0 * R + i] = prefix_s[i];
      rout[(size_t)c0 * R + i] = rank_s[i];
    }
    if (tid < ncol && mode_s[tid] >= 0) W.mode[c0 + tid] = mode_s[tid];
    // clear the buffer the NEXT level accumulates into (last read by launch level-1)
    uint32_t* hclr = W.hist + (size_t)((level + 1) % 3) * T * R * 256 + (size_t)c0 * R * 256;
    for (int i = tid; i < ncol * R * 256; i += kBandThreads) hclr[i] = 0;
  }
  if (level == kLevels) return;                            // launch 8 only completes the prefixes

  // ---- 2. one pass over this CTA's slab of rows (software-pipelined: the next U loads are in flight while
  //         the current U elements are classified)
  const int c = tid % kColBlock, rl = tid / kColBlock;
  const int64_t rows_per = (S + n_slabs - 1) / n_slabs;
  const int64_t r_begin = (int64_t)slab * rows_per, r_end = min(S, r_begin + rows_per);
  bool saw_nan = false;
  const int my_mode = c < ncol ? mode_s[c] : -2;
  if (my_mode != -2) {
    const int nu = nuniq_s[c];
    const double* col = A + c0 + c;
    uint32_t* hc = hist_s + c * CS;
    const uint64_t* up = uprefix_s + c * R;
    const int* us = uslot_s + c * R;
    constexpr int U = MODE == 0 ? RVLP_BAND_UNROLL0 : RVLP_BAND_UNROLL;
    const int64_t step = (int64_t)kRowsPerIter * U;
    int run_bin = -1;                                      // MODE 0: run-length aggregation of the digit
    uint32_t run_len = 0;
    auto collect = [&](double x) {                         // <= kCandCap live elements in this column
      const uint32_t at = atomicAdd(W.ncand + c0 + c, 1u);
      W.cand[(size_t)(c0 + c) * kCandCap + at] = x;
    };
    // MODE 1: the first kRegPrefix unique prefixes sit in registers behind a 32-bit Bloom mask over the low
    // 5 bits of the candidate's head (most elements match nothing and leave after two instructions)
    constexpr int kRegPrefix = 6;
    uint32_t pr[kRegPrefix];
    int sl[kRegPrefix];
    uint32_t bloom = 0;
    if (MODE == 1) {
#pragma unroll
      for (int q = 0; q < kRegPrefix; ++q) {
        pr[q] = q < nu ? (uint32_t)up[q] : 0xffffffffu;    // heads have <= 24 bits: all-ones never matches
        sl[q] = q < nu ? us[q] * 256 : 0;
      }
      for (int q = 0; q < nu; ++q) bloom |= 1u << ((uint32_t)up[q] & 31u);
    }
    const int hs = 32 - 8 * level, ds = 24 - 8 * level, shift = 56 - 8 * level;
    auto classify = [&](double x) {
      if (MODE == 0) {                                     // digit 0 of every element; only the high word matters
        saw_nan |= x != x;
        const uint32_t hw = (uint32_t)__double2hiint(x);
        const uint32_t k = hw ^ ((uint32_t)((int32_t)hw >> 31) | 0x80000000u);
        const int bin = (int)(k >> 24);
        if (bin == run_bin) {
          ++run_len;
        } else {
          if (run_len) atomicAdd(hc + run_bin, run_len);
          run_bin = bin;
          run_len = 1;
        }
      } else if (MODE == 1) {
        const uint32_t hw = (uint32_t)__double2hiint(x);
        const uint32_t k = hw ^ ((uint32_t)((int32_t)hw >> 31) | 0x80000000u);
        const uint32_t hi = k >> hs;
        if ((bloom >> (hi & 31u)) & 1u) {
          int slot = -1;
#pragma unroll
          for (int q = 0; q < kRegPrefix; ++q)
            if (hi == pr[q]) slot = sl[q];
          for (int q = kRegPrefix; q < nu; ++q)
            if (hi == (uint32_t)up[q]) slot = us[q] * 256;
          if (slot >= 0) {
            if (my_mode >= 0) collect(x);
            else atomicAdd(hc + slot + (int)((k >> ds) & 255u), 1u);
          }
        }
      } else {
        const uint64_t k = key_of(x);
        const uint64_t hi = key_head(k, level);
        int slot = -1;
        for (int q = 0; q < nu; ++q)
          if (hi == up[q]) { slot = us[q] * 256; break; }
        if (slot >= 0) {
          if (my_mode >= 0) collect(x);
          else atomicAdd(hc + slot + (int)((k >> shift) & 255u), 1u);
        }
      }
    };
    // full iterations: no bounds checks; the next U loads are issued before the current U are classified
    const int64_t first = r_begin + rl;
    const int64_t n_full = first + (U - 1) * kRowsPerIter < r_end ? (r_end - first - (U - 1) * kRowsPerIter + step - 1) / step : 0;
    const double* ptr = col + first * T;
    const int64_t dstep = step * T, drow = (int64_t)kRowsPerIter * T;
    double cur[U], nxt[U];
    if (n_full > 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) cur[u] = __ldcs(ptr + u * drow);
      for (int64_t it = 1; it < n_full; ++it) {
        ptr += dstep;
#pragma unroll
        for (int u = 0; u < U; ++u) nxt[u] = __ldcs(ptr + u * drow);
#pragma unroll
        for (int u = 0; u < U; ++u) classify(cur[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) classify(cur[u]);
    }
    for (int64_t r = first + n_full * step; r < r_end; r += kRowsPerIter) classify(__ldcs(col + r * T));   // tail
    if (MODE == 0 && run_len) atomicAdd(hc + run_bin, run_len);
  }
  if (saw_nan) atomicOr(W.nan + c0 + c, 1u);
  __syncthreads();
  // ---- 3. merge into the level's global histogram
  uint32_t* hcur = W.hist + (size_t)(level % 3) * T * R * 256 + (size_t)c0 * R * 256;
  for (int i = tid; i < ncol * HS * 256; i += kBandThreads) {
    const int cc = i / (HS * 256), w = i - cc * (HS * 256);
    const uint32_t v = hist_s[cc * CS + w];
    if (v) atomicAdd(hcur + cc * (R * 256) + w, v);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kBandThreads, (MODE == 0 ? RVLP_BAND_MINB0 : RVLP_BAND_MINB1))
band_level_kernel(const double* __restrict__ A, int64_t S, int64_t T, int level, BandTargets tg, BandWorkspace W) {
  band_level_body<MODE>(A, S, T, level, tg, W, (int)blockIdx.y, (int)gridDim.y);
}

// Finishing kernel: one warp per column.  Streaming columns have complete keys in prefix[8]; collected columns
// run their remaining levels on the candidates (cached in shared memory).  Then numpy's blend.
constexpr int kFinishSmem = kBandWarps * (kCandCap * 8 + 256 * 4);
__device__ __forceinline__ void band_finish_body(int64_t T, const BandTargets& tg, const BandWorkspace& W, double* __restrict__ out,
                                                 int64_t col_block) {
  extern __shared__ __align__(16) unsigned char bsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* cand_s = reinterpret_cast<double*>(bsm) + warp * kCandCap;
  uint32_t* hist_w = reinterpret_cast<uint32_t*>(bsm + kBandWarps * kCandCap * 8) + warp * 256;
  const int R = 2 * tg.n_q;
  const int64_t c = col_block * kBandWarps + warp;
  if (c >= T) return;
  const int mode = W.mode[c];
  if (mode == kModeFastDone) return;                       // warp-uniform: the column's result is already in `out`
  const int L0 = mode >= 0 ? mode : kLevels;
  uint32_t n = 0;
  if (mode >= 0) {
    n = W.ncand[c];
    for (uint32_t i = lane; i < n; i += 32) cand_s[i] = W.cand[(size_t)c * kCandCap + i];
    __syncwarp();
  }
  const uint64_t* pre = W.prefix + (size_t)L0 * T * R + (size_t)c * R;
  const uint32_t* rk0 = W.rank + (size_t)L0 * T * R + (size_t)c * R;
  for (int q = 0; q < tg.n_q; ++q) {
    double v[2];
    for (int h = 0; h < 2; ++h) {
      const int r = 2 * q + h;
      uint64_t p = pre[r];
      uint32_t rk = rk0[r];
      for (int L = L0; L < kLevels; ++L) {                 // collected columns only
        for (int i = lane; i < 256; i += 32) hist_w[i] = 0;
        __syncwarp();
        const int shift = 56 - 8 * L;
        for (uint32_t i = lane; i < n; i += 32) {
          const uint64_t k = key_of(cand_s[i]);
          if (key_head(k, L) == p) atomicAdd(&hist_w[(k >> shift) & 255u], 1u);
        }
        __syncwarp();
        uint32_t cnt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) cnt[j] = hist_w[lane * 8 + j];
        int digit;
        uint32_t newrank, bincount;
        warp_pick_bin(cnt, rk, lane, digit, newrank, bincount);
        p = (p << 8) | (uint64_t)digit;
        rk = newrank;
        __syncwarp();
      }
      v[h] = value_of(p);
    }
    if (lane == 0) {
      double res = numpy_lerp(v[0], v[1], tg.gamma[q]);
      if (W.nan[c]) res = __longlong_as_double(0x7ff8000000000000ll);
      out[(size_t)q * T + c] = res;
    }
  }
}

__global__ void __launch_bounds__(kBandThreads)
band_finish_kernel(int64_t T, BandTargets tg, BandWorkspace W, double* __restrict__ out) {
  band_finish_body(T, tg, W, out, (int64_t)blockIdx.x);
}

// After the two-pass path (rvlp_bands_fast.cuh): the columns it could not finish (target bins that do not fit the
// candidate buffer: heavy ties, adversarial data) take the radix path here - ONE launch, one CTA per column block over
// ALL rows, so that the levels need no grid-wide dependency and run back to back inside the kernel; a CTA whose
// columns are all finished returns at once (the usual case: ~3 us instead of ten launches).  Same bodies, same bits.
static_assert(kColBlock == kBandWarps, "band_fallback_kernel finishes its column block with one warp per column");
__global__ void __launch_bounds__(kBandThreads, 2)
band_fallback_kernel(const double* __restrict__ A, int64_t S, int64_t T, BandTargets tg, BandWorkspace W, double* __restrict__ out) {
  __shared__ int todo_s;
  const int64_t c0 = (int64_t)blockIdx.x * kColBlock;
  if (threadIdx.x == 0) {
    int todo = 0;
    for (int c = 0; c < kColBlock && c0 + c < T; ++c) todo |= W.mode[c0 + c] != kModeFastDone;
    todo_s = todo;
  }
  __syncthreads();
  if (!todo_s) return;
  for (int level = 0; level <= kLevels; ++level) {
    if (band_mode_of(level) == 0) band_level_body<0>(A, S, T, level, tg, W, 0, 1);
    else if (band_mode_of(level) == 1) band_level_body<1>(A, S, T, level, tg, W, 0, 1);
    else band_level_body<2>(A, S, T, level, tg, W, 0, 1);
    __threadfence();
    __syncthreads();
  }
  band_finish_body(T, tg, W, out, (int64_t)blockIdx.x);
}

}  // namespace rvlp
