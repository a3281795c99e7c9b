// rvlp_bands_fast.cuh — K6, two streaming passes instead of four (row f-2: `np.percentile(matrix, q, axis=0)`,
// /root/reference/src/ravest/fit.py:2239-2240, 2493-2495, 6440-6450).
//
// The radix select of rvlp_bands.cuh bins by the leading BITS of the doubles.  A column of a posterior RV matrix spans
// one or two binary exponents, so its first two 8-bit digits (sign, exponent, four mantissa bits) separate almost
// nothing and it takes four passes over the S x T matrix until the elements around a target rank fit a candidate
// buffer.  Binning in VALUE space, around the targets, does it in two:
//   0. sample: 1024 evenly spaced rows per column, sorted in shared memory; [a, b] = the sample's order statistics a
//      safe distance (5 sigma of the sampling error + 4) outside the extreme target ranks (sample min / max at the ends);
//   1. count: one pass; bin(x) = 0 for x < a, 1023 for x >= b, 1 + floor((x - a) * 1022 / (b - a)) in between - a
//      MONOTONE map, so "elements in lower bins" is an exact rank offset whatever the sample looked like;
//   1b. plan: the bin and in-bin rank of every target rank follow from the histogram (one warp per column);
//   2. collect: a second pass copies the elements of the target bins (~90 per bin at S = 1e5) into the column's
//      candidate buffer;
//   3. finish: exact selection of the remaining rank among the candidates of the target's bin, numpy's `_lerp`.
// The result is the exact order statistic - bit-identical to numpy - for ANY data; only the speed depends on the
// sample.  A column whose target bins do not fit the buffer (heavy ties, adversarial data) is left to the radix path
// (band_fallback_kernel, rvlp_bands.cuh: one launch afterwards, skips every column finished here).  Constant samples (a == b, e.g. an all-zero trend
// column) use the three bins <a, ==a, >a and need no candidates when the targets fall on the constant.
// HBM: 2 x S x T x 8 bytes + 1 % for the sample; integer counting only, so the bits do not depend on the grid.
#pragma once
#include <type_traits>

#include "rvlp_bands.cuh"

namespace rvlp {

constexpr int kFastBins = 1024;
#ifndef RVLP_BANDF_SAMPLE
#define RVLP_BANDF_SAMPLE 256
#endif
constexpr int kFastSample = RVLP_BANDF_SAMPLE;   // sample rows per column: 256 -> the bracket is the sample's ~4 % / ~96 % points for
                                                 // [15.85, 50, 84.15]: 30 % wider bins than with 1024 rows, a quarter of the gather
constexpr int kFastCap = 4096;            // candidate doubles per column
constexpr int kFastQuad = 128;            // a key list up to this size is ranked by counting, longer ones by radix steps first
constexpr int kFastKeysSmem = 1024;       // finish: columns with more candidates keep their key lists in global scratch
constexpr int kFastStage = 448;           // collect pass: candidates staged per (CTA, column) in shared memory
constexpr int kFastMinRows = 8192;        // below this the radix path's candidate buffer is reached quickly anyway

struct BandFastWs {
  double* a;            // [T]
  double* scale;        // [T]  1022 / (b - a); 0 marks a constant sample
  uint32_t* hist;       // [T][kFastBins]
  int32_t* tbin;        // [T][R]  bin of each target rank (-1: column left to the radix path)
  uint32_t* trank;      // [T][R]  rank inside that bin
  uint32_t* ncand;      // [T]
  double* cand;         // [T][kFastCap]
  uint64_t* keys;       // [T][kFastCap]  finish-kernel scratch for columns with > kFastKeysSmem candidates
};
__host__ __device__ inline size_t band_fast_ws_bytes(int64_t T, int R) {
  size_t b = 0;
  b += band_align((size_t)T * 8) * 2;
  b += band_align((size_t)T * kFastBins * 4);
  b += band_align((size_t)T * R * 4) * 2;
  b += band_align((size_t)T * 4);
  b += band_align((size_t)T * kFastCap * 8) * 2;
  return b;
}
__host__ __device__ inline BandFastWs band_fast_carve(void* base, int64_t T, int R) {
  BandFastWs F;
  unsigned char* p = reinterpret_cast<unsigned char*>(base);
  F.a = reinterpret_cast<double*>(p); p += band_align((size_t)T * 8);
  F.scale = reinterpret_cast<double*>(p); p += band_align((size_t)T * 8);
  F.hist = reinterpret_cast<uint32_t*>(p); p += band_align((size_t)T * kFastBins * 4);
  F.tbin = reinterpret_cast<int32_t*>(p); p += band_align((size_t)T * R * 4);
  F.trank = reinterpret_cast<uint32_t*>(p); p += band_align((size_t)T * R * 4);
  F.ncand = reinterpret_cast<uint32_t*>(p); p += band_align((size_t)T * 4);
  F.cand = reinterpret_cast<double*>(p); p += band_align((size_t)T * kFastCap * 8);
  F.keys = reinterpret_cast<uint64_t*>(p);
  return F;
}

// the monotone value -> bin map (see the header); scale == 0: constant sample.
// (x - a) * scale -> floor -> integer clamps: five instructions, no double compares.  cvt.rmi.s32.f64 saturates
// (-inf -> INT_MIN -> bin 0, +inf -> INT_MAX -> bin 1023) and sends NaN to 0 -> bin 1 (the column is flagged NaN anyway).
__device__ __forceinline__ int fast_bin_lin(double x, double a, double scale) {
  const int i = __double2int_rd((x - a) * scale);
  return max(min(i, kFastBins - 2) + 1, 0);
}
__device__ __forceinline__ int fast_bin_const(double x, double a) {   // in key order, so that -0 / +0 / NaN sort as in the radix path
  const uint64_t kx = key_of(x), ka = key_of(a);
  return kx < ka ? 0 : (kx == ka ? 1 : kFastBins - 1);
}
__device__ __forceinline__ int fast_bin(double x, double a, double scale) {
  return scale == 0.0 ? fast_bin_const(x, a) : fast_bin_lin(x, a, scale);
}

// Exact selection by the whole CTA: the key of 0-based rank `rk` among list[0, nb) (shared or global memory).  Lists
// up to kFastQuad keys are ranked by counting (the key with `less <= rank < less + equal`); longer ones are first
// narrowed by radix steps on the eight leading bits in which the live keys still differ.  Integer compares only: the
// result does not depend on the order of the list.
struct BandSelState {
  unsigned long long mn, mx, res;
  uint32_t qn, digit, newrank, bincount;
};
// `rk2` (optional, CTA-uniform): a second rank answered by the same counting pass when the list is short - the two
// ranks of a percentile pair sit in one bin almost always; *res2 is written only then (returns true through `got2`).
__device__ __forceinline__ uint64_t band_cta_select(const uint64_t* list, uint32_t nb, uint32_t rk, uint64_t* q_s,
                                                    uint32_t* hist_s, BandSelState* st, uint32_t rk2 = 0xffffffffu,
                                                    uint64_t* res2 = nullptr, bool* got2 = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31;
  uint64_t lo = 0, hi = ~0ull;
  uint32_t live = nb;
  if (got2) *got2 = false;
  for (;;) {
    __syncthreads();                                       // the previous round's / call's reads of *st are done
    if (live <= (uint32_t)kFastQuad) {
      const bool pair = res2 != nullptr && live == nb;     // no radix step has narrowed the list: both ranks are in it
      if (tid == 0) st->qn = 0;
      __syncthreads();
      for (uint32_t i = tid; i < nb; i += kBandThreads) {
        const uint64_t k = list[i];
        if (k >= lo && k <= hi) q_s[atomicAdd(&st->qn, 1u)] = k;
      }
      __syncthreads();
      for (uint32_t i = tid; i < live; i += kBandThreads) {
        const uint64_t ki = q_s[i];
        uint32_t less = 0, eq = 0;
        for (uint32_t j = 0; j < live; ++j) {
          const uint64_t kj = q_s[j];
          less += kj < ki ? 1u : 0u;
          eq += kj == ki ? 1u : 0u;
        }
        if (less <= rk && rk < less + eq) st->res = ki;    // ties write the same key
        if (pair && less <= rk2 && rk2 < less + eq) st->mx = ki;
      }
      __syncthreads();
      if (pair) { *res2 = st->mx; *got2 = true; }
      return st->res;
    }
    if (tid == 0) { st->mn = ~0ull; st->mx = 0ull; }
    if (tid < 256) hist_s[tid] = 0;
    __syncthreads();
    unsigned long long mn = ~0ull, mx = 0ull;
    for (uint32_t i = tid; i < nb; i += kBandThreads) {
      const uint64_t k = list[i];
      if (k >= lo && k <= hi) { mn = k < mn ? k : mn; mx = k > mx ? k : mx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = a < mn ? a : mn;
      mx = b > mx ? b : mx;
    }
    if (lane == 0) { atomicMin(&st->mn, mn); atomicMax(&st->mx, mx); }
    __syncthreads();
    mn = st->mn;
    mx = st->mx;
    if (mn == mx) return mn;                               // CTA-uniform
    const int top = 63 - __clzll((long long)(mn ^ mx));    // the live keys agree above bit `top`
    const int shift = top >= 7 ? top - 7 : 0;
    for (uint32_t i = tid; i < nb; i += kBandThreads) {
      const uint64_t k = list[i];
      if (k >= lo && k <= hi) atomicAdd(&hist_s[(uint32_t)(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      uint32_t cnt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cnt[j] = hist_s[lane * 8 + j];
      int digit;
      uint32_t newrank, bincount;
      warp_pick_bin(cnt, rk, lane, digit, newrank, bincount);
      if (lane == 0) { st->digit = (uint32_t)digit; st->newrank = newrank; st->bincount = bincount; }
    }
    __syncthreads();
    const uint64_t base = shift + 8 >= 64 ? 0ull : (mn >> (shift + 8)) << (shift + 8);
    lo = base | ((uint64_t)st->digit << shift);
    hi = lo | (((uint64_t)1 << shift) - 1ull);
    rk = st->newrank;
    live = st->bincount;                                   // < previous live: bit `top` splits the list
  }
}


// Ascending bitonic sort of n = 2^m keys in shared memory by the whole CTA (one __syncthreads per stage).  A few hundred
// keys: ~40 stages of n / 2 compare-exchanges - a quarter of the instructions of ranking every key by counting.
__device__ __forceinline__ void band_cta_sort(uint64_t* k, int n) {
  const int tid = threadIdx.x;
  __syncthreads();
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int p = tid; p < n / 2; p += kBandThreads) {
        const int lo = 2 * p - (p & (stride - 1));         // partner pairs (lo, lo + stride)
        const bool up = (lo & size) == 0;
        const uint64_t x = k[lo], y = k[lo + stride];
        if ((x > y) == up) { k[lo] = y; k[lo + stride] = x; }
      }
      __syncthreads();
    }
  }
}

// ---- 0. sample + bracket: one CTA per column gathers the sample, sorts it (256 keys: 36 stages) and reads its two
// bracket order statistics.  The four columns of a 32-byte sector are fetched by four CTAs and meet in L2.
__global__ void __launch_bounds__(kBandThreads)
band_fast_sample_kernel(const double* __restrict__ A, int64_t S, int64_t T, BandTargets tg, BandFastWs F) {
  static_assert((kFastSample & (kFastSample - 1)) == 0, "bitonic sort: power of two");
  __shared__ uint64_t k[kFastSample];
  const int tid = threadIdx.x;
  const int64_t c = blockIdx.x;
  for (int i = tid; i < kFastSample; i += kBandThreads) {
    const int64_t row = ((int64_t)i * S) / kFastSample;    // i < 2^10, S < 2^31
    k[i] = key_of(A[row * T + c]);
  }
  band_cta_sort(k, kFastSample);
  if (tid == 0) {
    const int R = 2 * tg.n_q;
    uint32_t kmin = 0xffffffffu, kmax = 0;
    for (int r = 0; r < R; ++r) { kmin = min(kmin, tg.k[r]); kmax = max(kmax, tg.k[r]); }
    const double m = (double)kFastSample;
    const double plo = (double)kmin / (double)S, phi = (double)(kmax + 1) / (double)S;
    int ra = (int)floor(plo * m - 5.0 * sqrt(m * plo * (1.0 - plo)) - 4.0);
    int rb = (int)ceil(phi * m + 5.0 * sqrt(m * phi * (1.0 - phi)) + 4.0);
    ra = ra < 0 ? 0 : ra;
    rb = rb > kFastSample - 1 ? kFastSample - 1 : rb;
    const double a = value_of(k[ra]), b = value_of(k[rb]);
    double scale = 0.0;
    if (b > a) {
      scale = (double)(kFastBins - 2) / (b - a);
      if (!(scale < 1.79769313486231570e308) || !(scale > 0.0)) scale = 0.0;   // b - a denormal / non-finite ends
    }
    F.a[c] = a;                                            // NaN / inf order statistics: a compares false or b - a is not
    F.scale[c] = (a == a && fabs(a) <= 1.79769313486231570e308) ? scale : 0.0;   // finite -> constant mode
  }
}

// ---- 1. (COLLECT = false) count pass / 2. (COLLECT = true) collect pass.  Same streaming skeleton as band_level_kernel:
// a CTA owns kColBlock adjacent columns and a slab of rows.
#ifndef RVLP_BANDF_MINB
#define RVLP_BANDF_MINB 3
#endif
template <bool COLLECT>
__global__ void __launch_bounds__(kBandThreads, RVLP_BANDF_MINB)
band_fast_pass_kernel(const double* __restrict__ A, int64_t S, int64_t T, BandTargets tg, BandWorkspace W, BandFastWs F) {
  extern __shared__ __align__(16) unsigned char bsm[];
  constexpr int CS = kFastBins + 1;                        // +1: the same bin of adjacent columns in different banks
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(bsm);     // COUNT: [kColBlock][CS]
  int* tb_s = reinterpret_cast<int*>(bsm);                 // COLLECT: [kColBlock][kMaxTargets] target bins (-1 none)
  uint32_t* mask_s = reinterpret_cast<uint32_t*>(bsm) + kColBlock * kMaxTargets;   // COLLECT: [kColBlock][32] bin bitmap
  uint32_t* scnt_s = mask_s + kColBlock * (kFastBins / 32);                        // COLLECT: [kColBlock] staged | [kColBlock] base
  double* stage_s = reinterpret_cast<double*>(scnt_s + 2 * kColBlock);             // COLLECT: [kColBlock][kFastStage]
  static_assert((kColBlock * kMaxTargets + kColBlock * (kFastBins / 32) + 2 * kColBlock) * 4 + kColBlock * kFastStage * 8 <=
                kColBlock * (kFastBins + 1) * 4, "collect staging must fit the count pass's histogram block");
  const int R = 2 * tg.n_q;
  const int tid = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * kColBlock;
  const int ncol = (int)min((int64_t)kColBlock, T - c0);

  if (!COLLECT) {
    for (int i = tid; i < kColBlock * CS; i += kBandThreads) hist_s[i] = 0;
  } else {
    // the target bins were resolved once per column by band_fast_plan_kernel
    for (int i = tid; i < kColBlock * kMaxTargets; i += kBandThreads) tb_s[i] = -1;
    for (int i = tid; i < kColBlock * (kFastBins / 32); i += kBandThreads) mask_s[i] = 0;
    if (tid < 2 * kColBlock) scnt_s[tid] = 0;
    __syncthreads();
    if (tid < ncol && W.mode[c0 + tid] == kModeFastDone) {
      for (int r = 0; r < R; ++r) {
        const int b = F.tbin[(size_t)(c0 + tid) * R + r];
        tb_s[tid * kMaxTargets + r] = b;
        if (b >= 0) mask_s[tid * (kFastBins / 32) + (b >> 5)] |= 1u << (b & 31);
      }
    }
  }
  __syncthreads();

  const int c = tid % kColBlock, rl = tid / kColBlock;
  const int64_t rows_per = (S + gridDim.y - 1) / gridDim.y;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per, r_end = min(S, r_begin + rows_per);
  bool saw_nan = false;
  if (c < ncol) {
    const double a = F.a[c0 + c], scale = F.scale[c0 + c];
    const double* col = A + c0 + c;
    uint32_t* hc = hist_s + c * CS;
    int nt = 0, tlo = kFastBins, thi = -1;                 // COLLECT: [tlo, thi] = range of the target bins
    const uint32_t* mk = mask_s + c * (kFastBins / 32);
    if (COLLECT) {
      for (int r = 0; r < R; ++r) {
        const int b = tb_s[c * kMaxTargets + r];
        if (b >= 0) { ++nt; tlo = min(tlo, b); thi = max(thi, b); }
      }
    }
    if (!COLLECT || nt > 0) {
      // one element; CONST (a constant sample: three bins in key order) is a per-column property, hoisted out of the loop
      auto visit = [&](double x, auto is_const) {
        const int bin = decltype(is_const)::value ? fast_bin_const(x, a) : fast_bin_lin(x, a, scale);
        if (!COLLECT) {
          saw_nan |= x != x;
          atomicAdd(hc + bin, 1u);                         // (neighbouring samples rarely share a value-space bin: no run lengths)
        } else {
          if (bin >= tlo && bin <= thi && ((mk[bin >> 5] >> (bin & 31)) & 1u)) {
            // staged in shared memory (a global atomic's round trip per candidate stalled the whole warp); the few
            // that overflow the stage go straight to the column's buffer
            const uint32_t at = atomicAdd(scnt_s + c, 1u);
            if (at < (uint32_t)kFastStage) stage_s[c * kFastStage + at] = x;
            else F.cand[(size_t)(c0 + c) * kFastCap + atomicAdd(F.ncand + c0 + c, 1u)] = x;
          }
        }
      };
      // software pipeline without register moves: two register sets, each loaded while the other is classified
      auto stream = [&](auto is_const) {
        constexpr int U = 8;
        const int64_t drow = (int64_t)kRowsPerIter * T;
        const double* ptr = col + (r_begin + rl) * T;
        int64_t left = r_end > r_begin + rl ? (r_end - (r_begin + rl) + kRowsPerIter - 1) / kRowsPerIter : 0;   // this thread's rows
        double ra[U], rb[U];
        if (left >= U) {
#pragma unroll
          for (int u = 0; u < U; ++u) ra[u] = __ldcs(ptr + u * drow);
          ptr += U * drow;
          left -= U;
          while (left >= 2 * U) {
#pragma unroll
            for (int u = 0; u < U; ++u) rb[u] = __ldcs(ptr + u * drow);
#pragma unroll
            for (int u = 0; u < U; ++u) visit(ra[u], is_const);
#pragma unroll
            for (int u = 0; u < U; ++u) ra[u] = __ldcs(ptr + (U + u) * drow);
#pragma unroll
            for (int u = 0; u < U; ++u) visit(rb[u], is_const);
            ptr += 2 * U * drow;
            left -= 2 * U;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) visit(ra[u], is_const);
        }
        for (; left > 0; --left, ptr += drow) visit(__ldcs(ptr), is_const);   // tail
      };
      if (scale == 0.0) stream(std::true_type{});
      else stream(std::false_type{});
    }
  }
  if (COLLECT) {
    __syncthreads();
    if (tid < ncol) {
      const uint32_t n = min(scnt_s[tid], (uint32_t)kFastStage);
      scnt_s[tid] = n;
      scnt_s[kColBlock + tid] = n ? atomicAdd(F.ncand + c0 + tid, n) : 0u;
    }
    __syncthreads();
    for (int cc = 0; cc < ncol; ++cc) {
      const uint32_t n = scnt_s[cc], base = scnt_s[kColBlock + cc];
      double* dst = F.cand + (size_t)(c0 + cc) * kFastCap + base;
      for (uint32_t i = tid; i < n; i += kBandThreads) dst[i] = stage_s[cc * kFastStage + i];
    }
  } else {
    if (saw_nan) atomicOr(W.nan + c0 + c, 1u);
    __syncthreads();
    uint32_t* hg = F.hist + (size_t)c0 * kFastBins;
    for (int i = tid; i < ncol * kFastBins; i += kBandThreads) {
      const int cc = i / kFastBins, b = i - cc * kFastBins;
      const uint32_t v = hist_s[cc * CS + b];
      if (v) atomicAdd(hg + (size_t)cc * kFastBins + b, v);
    }
  }
}

// ---- 1b. plan: one warp per column turns the complete histogram into the bin and in-bin rank of every target rank,
// and decides whether the column's target bins fit the candidate buffer (else it is left to the radix path)
__global__ void __launch_bounds__(kBandThreads)
band_fast_plan_kernel(int64_t T, BandTargets tg, BandWorkspace W, BandFastWs F) {
  __shared__ uint32_t hs[kBandWarps][kFastBins + kFastBins / 32];   // bin b at b + b / 32: lane-strided reads hit 32 banks
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * kBandWarps + warp;
  if (c >= T) return;                                      // warp-uniform
  const int R = 2 * tg.n_q;
  const uint32_t* h = F.hist + (size_t)c * kFastBins;
  for (int j = 0; j < kFastBins / 32; ++j) hs[warp][j * 33 + lane] = h[j * 32 + lane];
  __syncwarp();
  uint32_t cnt[kFastBins / 32];
  uint32_t mine = 0;
#pragma unroll
  for (int j = 0; j < kFastBins / 32; ++j) { cnt[j] = hs[warp][lane * 33 + j]; mine += cnt[j]; }
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  const uint32_t excl = incl - mine;
  uint32_t live = 0;
  int seen[kMaxTargets];
  int nseen = 0;
  bool ok = true;
  const bool constant = F.scale[c] == 0.0;
  for (int r = 0; r < R; ++r) {
    const uint32_t rk = tg.k[r];
    int bin = -1;
    uint32_t inb = 0, bc = 0;
    if (rk >= excl && rk < incl) {
      uint32_t run = excl;
#pragma unroll
      for (int j = 0; j < kFastBins / 32; ++j) {
        if (rk >= run && rk < run + cnt[j]) { bin = lane * (kFastBins / 32) + j; inb = rk - run; bc = cnt[j]; }
        run += cnt[j];
      }
    }
    const unsigned who = __ballot_sync(0xffffffffu, bin >= 0);
    const int src = who ? __ffs(who) - 1 : 0;
    bin = __shfl_sync(0xffffffffu, bin, src);
    inb = __shfl_sync(0xffffffffu, inb, src);
    bc = __shfl_sync(0xffffffffu, bc, src);
    if (!who) ok = false;                                  // cannot happen (the counts sum to S > rk)
    const bool free_bin = constant && bin == 1;            // every element of this bin equals a: no candidates needed
    bool dup = false;
    for (int q = 0; q < nseen; ++q) dup |= seen[q] == bin;
    if (!dup) {
      seen[nseen++] = bin;
      if (!free_bin) live += bc;
    }
    if (lane == 0) {
      F.tbin[(size_t)c * R + r] = free_bin ? -2 : bin;
      F.trank[(size_t)c * R + r] = inb;
    }
  }
  if (live > (uint32_t)kFastCap) ok = false;
  if (lane == 0 && ok) W.mode[c] = kModeFastDone;          // otherwise mode stays -1: the radix path streams the column
}

// ---- 3. finish: one CTA per column.  The candidates are dealt to one key list per distinct target bin (shared
// memory); every target is then an exact selection inside its list: lists up to kFastQuad keys by counting (the key
// with `less <= rank < less + equal`), longer ones (S ~ 1e6, ties) first narrowed by radix steps on the bits in which
// the list still differs.  Integer compares only: the result does not depend on the order the candidates arrived in.
constexpr int kFastFinishSmem = kFastKeysSmem * 8 + kFastQuad * 8 + 256 * 4;
__global__ void __launch_bounds__(kBandThreads)
band_fast_finish_kernel(int64_t T, BandTargets tg, BandWorkspace W, BandFastWs F, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char bsm[];
  uint64_t* key_s = reinterpret_cast<uint64_t*>(bsm);      // [kFastKeysSmem], one segment per distinct target bin
  uint64_t* q_s = key_s + kFastKeysSmem;                   // [kFastQuad]
  uint32_t* hist_s = reinterpret_cast<uint32_t*>(q_s + kFastQuad);   // [256]
  __shared__ double v_s[kMaxTargets];
  __shared__ int dbin_s[kMaxTargets], tslot_s[kMaxTargets], nd_s;
  __shared__ uint32_t dcnt_s[kMaxTargets], doff_s[kMaxTargets], dpos_s[kMaxTargets];
  __shared__ BandSelState st;
  const int tid = threadIdx.x;
  const int R = 2 * tg.n_q;
  const int64_t c = blockIdx.x;
  if (W.mode[c] != kModeFastDone) return;
  if (tid == 0) {
    int nd = 0;
    for (int r = 0; r < R; ++r) {
      const int tb = F.tbin[(size_t)c * R + r];
      int slot = -1;
      if (tb >= 0) {
        for (int q = 0; q < nd; ++q) if (dbin_s[q] == tb) slot = q;
        if (slot < 0) { slot = nd; dbin_s[nd] = tb; dcnt_s[nd] = 0; dpos_s[nd] = 0; ++nd; }
      }
      tslot_s[r] = slot;                                   // -1: the target sits on a constant column's value
    }
    nd_s = nd;
  }
  __syncthreads();
  const int nd = nd_s;
  const uint32_t n = min(F.ncand[c], (uint32_t)kFastCap);
  const double a = F.a[c], scale = F.scale[c];
  const double* cand = F.cand + (size_t)c * kFastCap;
  uint64_t* keys = n <= (uint32_t)kFastKeysSmem ? key_s : F.keys + (size_t)c * kFastCap;   // CTA-uniform
  auto slot_of = [&](double x) {
    const int b = fast_bin(x, a, scale);
    int slot = 0;
    for (int q = 1; q < nd; ++q) slot = dbin_s[q] == b ? q : slot;
    return slot;
  };
  for (uint32_t i = tid; i < n; i += kBandThreads) atomicAdd(&dcnt_s[slot_of(cand[i])], 1u);
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (int q = 0; q < nd; ++q) { doff_s[q] = run; run += dcnt_s[q]; }
  }
  __syncthreads();
  for (uint32_t i = tid; i < n; i += kBandThreads) {
    const double x = cand[i];
    const int q = slot_of(x);
    keys[doff_s[q] + atomicAdd(&dpos_s[q], 1u)] = key_of(x);
  }
  __syncthreads();
  if (n <= (uint32_t)kFastKeysSmem) {
    // the usual case: a few hundred candidates.  The bins are monotone in the key, so ONE sort of all candidates puts
    // the lists one after the other in slot order of their bins: pad to a power of two with the largest key, sort,
    // index.  (doff_s was laid out in the order the target bins were met; the sorted order is by bin.)
    uint32_t np2 = 32;
    while (np2 < n) np2 <<= 1;
    for (uint32_t i = n + tid; i < np2; i += kBandThreads) key_s[i] = ~0ull;
    band_cta_sort(key_s, (int)np2);
    if (tid < R) {
      const int q = tslot_s[tid];
      if (q < 0) {
        v_s[tid] = a;
      } else {
        uint32_t below = 0;                                // candidates in target bins below this one
        for (int p = 0; p < nd; ++p) below += dbin_s[p] < dbin_s[q] ? dcnt_s[p] : 0u;
        v_s[tid] = value_of(key_s[below + F.trank[(size_t)c * R + tid]]);
      }
    }
  } else
  for (int r = 0; r < R; ++r) {
    const int q = tslot_s[r];                              // CTA-uniform
    if (q < 0) {
      if (tid == 0) v_s[r] = a;
      continue;
    }
    // the partner rank of the pair (k, k + 1) is answered by the same pass when it sits in the same list
    const bool same = (r & 1) == 0 && tslot_s[r + 1] == q;
    uint64_t k2 = 0;
    bool got2 = false;
    const uint64_t k = band_cta_select(keys + doff_s[q], dcnt_s[q], F.trank[(size_t)c * R + r], q_s, hist_s, &st,
                                       same ? F.trank[(size_t)c * R + r + 1] : 0xffffffffu, same ? &k2 : nullptr, &got2);
    if (tid == 0) v_s[r] = value_of(k);
    if (got2) {
      if (tid == 0) v_s[r + 1] = value_of(k2);
      ++r;
    }
  }
  __syncthreads();
  if (tid < tg.n_q) {
    double res = numpy_lerp(v_s[2 * tid], v_s[2 * tid + 1], tg.gamma[tid]);
    if (W.nan[c]) res = __longlong_as_double(0x7ff8000000000000ll);
    out[(size_t)tid * T + c] = res;
  }
}

}  // namespace rvlp
