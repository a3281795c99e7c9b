// rvlp_bands.cuh — K6: column-wise percentiles of an [S, T] matrix (row f-2).
//
// Replaces `np.percentile(rv_matrix, [15.85, 50, 84.15], axis=0)` at
// /root/reference/src/ravest/fit.py:2239-2240, 2493-2495 (and the GP twins fit.py:6440-6450): the step that
// follows the per-sample RV matrices of K2, so the S x T matrix never leaves the device and 3 x T comes back.
//
// numpy's default ("linear", Hyndman & Fan 7) needs, per column and percentile, the order statistics
// floor(v) and floor(v)+1 of the virtual index v = q (S-1) and blends them (numpy _lerp).  Exact selection,
// no sort: an MSB-first radix select on the order-preserving 64-bit image of the doubles, 8 bits per level,
// all columns and all targets at once.  One launch per level; each launch
//   1. (level > 0) scans the previous level's histograms and extends every target's key prefix by one digit,
//   2. streams its slab of rows ONCE (coalesced: a CTA owns kColBlock adjacent columns), counting the next
//      digit of every element that still matches a live prefix into shared-memory histograms,
//   3. merges them into the global histogram of the level with integer atomics.
// After 8 levels the prefixes ARE the order statistics (bit-exact); the last launch blends and stores.
// Integer counting only: the result does not depend on grid size, row split or arrival order.
// HBM-bound: 8 passes x S x T x 8 bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ravest_b200.h"

namespace rvlp {

constexpr int kColBlock = 8;              // adjacent columns per CTA: 64-byte row segments
constexpr int kBandThreads = 256;
constexpr int kRowsPerIter = kBandThreads / kColBlock;
constexpr int kMaxTargets = 2 * RVLP_MAX_PERCENTILES;
constexpr int kLevels = 8;

struct BandTargets {                      // host-resolved numpy index arithmetic (see rvlp_capi.cu)
  int n_q;
  uint32_t k[kMaxTargets];                // k[2q] = previous index, k[2q+1] = next index (0-based ranks)
  double gamma[RVLP_MAX_PERCENTILES];
};

// Workspace layout (all per column c in [0, T)), R = 2 n_q targets:
//   hist   [3][T][R][256] uint32   ping-pong-pong histograms (level L fills L%3, reads (L-1)%3, clears (L+1)%3)
//   prefix [2][T][R]      uint64   key prefix found so far (level L writes L%2, reads (L-1)%2)
//   rank   [2][T][R]      uint32   remaining 0-based rank among the elements matching the prefix
//   nan    [T]            uint32   column holds a NaN -> numpy returns NaN for it
struct BandWorkspace {
  uint32_t* hist;
  uint64_t* prefix;
  uint32_t* rank;
  uint32_t* nan;
};
__host__ __device__ inline size_t band_ws_bytes(int64_t T, int R) {
  size_t b = 0;
  b += (size_t)3 * T * R * 256 * 4;
  b += (size_t)2 * T * R * 8;
  b += (size_t)2 * T * R * 4;
  b += (size_t)T * 4;
  return (b + 255) & ~(size_t)255;
}
__host__ __device__ inline BandWorkspace band_ws_carve(void* base, int64_t T, int R) {
  BandWorkspace W;
  unsigned char* p = reinterpret_cast<unsigned char*>(base);
  W.hist = reinterpret_cast<uint32_t*>(p); p += (size_t)3 * T * R * 256 * 4;
  W.prefix = reinterpret_cast<uint64_t*>(p); p += (size_t)2 * T * R * 8;
  W.rank = reinterpret_cast<uint32_t*>(p); p += (size_t)2 * T * R * 4;
  W.nan = reinterpret_cast<uint32_t*>(p);
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

// numpy/lib/_function_base_impl.py `_lerp`: a + (b - a) t, replaced by b - (b - a)(1 - t) where t >= 0.5;
// every operation rounded separately (no FMA), as numpy's ufunc loops do.
__device__ __forceinline__ double numpy_lerp(double a, double b, double t) {
  const double diff = __dsub_rn(b, a);
  if (t >= 0.5) return __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
  return __dadd_rn(a, __dmul_rn(diff, t));
}

// smem: hist[kColBlock][R][256] | prefix[kColBlock][R] | rank | slot | n_uniq | uniq prefix / slot lists
__host__ __device__ inline int band_smem_bytes(int R) {
  return kColBlock * R * 256 * 4 + kColBlock * R * (8 + 4 + 4 + 8 + 4) + kColBlock * 4 + 16;
}

__global__ void __launch_bounds__(kBandThreads)
band_level_kernel(const double* __restrict__ A, int64_t S, int64_t T, int level, BandTargets tg, BandWorkspace W,
                  double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char bsm[];
  const int R = 2 * tg.n_q;
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(bsm);
  uint64_t* prefix_s = reinterpret_cast<uint64_t*>(hist_s + kColBlock * R * 256);
  uint64_t* uprefix_s = prefix_s + kColBlock * R;
  uint32_t* rank_s = reinterpret_cast<uint32_t*>(uprefix_s + kColBlock * R);
  int* slot_s = reinterpret_cast<int*>(rank_s + kColBlock * R);      // target -> target whose histogram it shares
  int* uslot_s = slot_s + kColBlock * R;                             // unique list: histogram slot
  int* nuniq_s = uslot_s + kColBlock * R;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * kColBlock;
  const int ncol = (int)min((int64_t)kColBlock, T - c0);
  const bool writer = blockIdx.y == 0;

  // ---- 1. extend the prefixes by the digit the previous level's histogram selects
  if (level == 0) {
    for (int i = tid; i < kColBlock * R; i += kBandThreads) {
      prefix_s[i] = 0;
      rank_s[i] = tg.k[i % R];
      slot_s[i] = 0;
    }
  } else {
    const uint32_t* hprev = W.hist + (size_t)((level - 1) % 3) * T * R * 256;
    const uint64_t* pprev = W.prefix + (size_t)((level - 1) & 1) * T * R;
    const uint32_t* rprev = W.rank + (size_t)((level - 1) & 1) * T * R;
    for (int i = warp; i < ncol * R; i += kBandThreads / 32) {
      const int c = i / R, r = i - c * R;
      const size_t g = (size_t)(c0 + c) * R;
      uint64_t pre;
      uint32_t rk;
      int slot = r;
      if (level == 1) {
        pre = 0; rk = tg.k[r]; slot = 0;
      } else {
        pre = pprev[g + r]; rk = rprev[g + r];
        for (int q = 0; q < r; ++q)                      // first target with the same prefix owns the histogram
          if (pprev[g + q] == pre) { slot = q; break; }
      }
      const uint32_t* h = hprev + (g + slot) * 256 + lane * 8;
      const uint4 v0 = *reinterpret_cast<const uint4*>(h), v1 = *reinterpret_cast<const uint4*>(h + 4);
      const uint32_t cnt[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      uint32_t mine = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) mine += cnt[q];
      uint32_t incl = mine;                              // inclusive warp scan of the 32 lane sums
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const uint32_t excl = incl - mine;
      const bool here = rk >= excl && rk < incl;         // exactly one lane (counts sum to >= rk + 1)
      int digit = 0;
      uint32_t newrank = 0;
      if (here) {
        uint32_t run = excl;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (rk >= run && rk < run + cnt[q]) { digit = lane * 8 + q; newrank = rk - run; }
          run += cnt[q];
        }
      }
      const unsigned who = __ballot_sync(0xffffffffu, here);
      const int src = who ? __ffs(who) - 1 : 0;
      digit = __shfl_sync(0xffffffffu, digit, src);
      newrank = __shfl_sync(0xffffffffu, newrank, src);
      if (lane == 0) {
        prefix_s[c * R + r] = (pre << 8) | (uint64_t)digit;
        rank_s[c * R + r] = newrank;
      }
    }
  }
  __syncthreads();
  if (level >= 1 && writer) {                            // state for the next launch
    uint64_t* pnext = W.prefix + (size_t)(level & 1) * T * R;
    uint32_t* rnext = W.rank + (size_t)(level & 1) * T * R;
    for (int i = tid; i < ncol * R; i += kBandThreads) {
      pnext[(size_t)c0 * R + i] = prefix_s[i];
      rnext[(size_t)c0 * R + i] = rank_s[i];
    }
  }
  if (level == kLevels) {                                // ---- done: the prefixes are the order statistics
    if (writer) {
      for (int i = tid; i < ncol * tg.n_q; i += kBandThreads) {
        const int c = i / tg.n_q, q = i - c * tg.n_q;
        const double a = value_of(prefix_s[c * R + 2 * q]), b = value_of(prefix_s[c * R + 2 * q + 1]);
        double v = numpy_lerp(a, b, tg.gamma[q]);
        if (W.nan[c0 + c]) v = __longlong_as_double(0x7ff8000000000000ll);
        out[(size_t)q * T + c0 + c] = v;
      }
    }
    return;
  }
  // unique prefixes per column (targets that agree so far share one histogram)
  if (tid < ncol) {
    int n = 0;
    for (int r = 0; r < R; ++r) {
      int slot = r;
      for (int q = 0; q < r; ++q)
        if (prefix_s[tid * R + q] == prefix_s[tid * R + r]) { slot = q; break; }
      slot_s[tid * R + r] = slot;
      if (slot == r) { uprefix_s[tid * R + n] = prefix_s[tid * R + r]; uslot_s[tid * R + n] = r; ++n; }
    }
    nuniq_s[tid] = n;
  }
  for (int i = tid; i < kColBlock * R * 256; i += kBandThreads) hist_s[i] = 0;
  if (writer) {                                          // clear the buffer the NEXT level accumulates into
    uint32_t* hclr = W.hist + (size_t)((level + 1) % 3) * T * R * 256 + (size_t)c0 * R * 256;
    for (int i = tid; i < ncol * R * 256; i += kBandThreads) hclr[i] = 0;
  }
  __syncthreads();

  // ---- 2. one pass over this CTA's slab of rows
  const int c = tid % kColBlock, rl = tid / kColBlock;
  const int64_t rows_per = (S + gridDim.y - 1) / gridDim.y;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per, r_end = min(S, r_begin + rows_per);
  const int shift = 56 - 8 * level;
  bool saw_nan = false;
  if (c < ncol) {
    const int nu = nuniq_s[c];
    const double* col = A + c0 + c;
    uint32_t* hc = hist_s + c * R * 256;
    const uint64_t* up = uprefix_s + c * R;
    const int* us = uslot_s + c * R;
    constexpr int U = 4;
    for (int64_t r = r_begin + rl; r < r_end; r += (int64_t)kRowsPerIter * U) {
      double x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + (int64_t)u * kRowsPerIter;
        x[u] = rr < r_end ? __ldcs(col + rr * T) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + (int64_t)u * kRowsPerIter >= r_end) break;
        const uint64_t k = key_of(x[u]);
        const int digit = (int)((k >> shift) & 255u);
        if (level == 0) {
          saw_nan |= x[u] != x[u];
          atomicAdd(hc + digit, 1u);
        } else {
          const uint64_t hi = k >> (shift + 8);
          for (int q = 0; q < nu; ++q)
            if (hi == up[q]) atomicAdd(hc + us[q] * 256 + digit, 1u);
        }
      }
    }
  }
  if (saw_nan) atomicOr(W.nan + c0 + c, 1u);
  __syncthreads();
  // ---- 3. merge into the level's global histogram
  uint32_t* hcur = W.hist + (size_t)(level % 3) * T * R * 256 + (size_t)c0 * R * 256;
  for (int i = tid; i < ncol * R * 256; i += kBandThreads) {
    const uint32_t v = hist_s[i];
    if (v) atomicAdd(hcur + i, v);
  }
}


}  // namespace rvlp
