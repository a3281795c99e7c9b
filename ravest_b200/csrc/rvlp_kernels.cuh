// rvlp_kernels.cuh — the sm_100a kernels of the batched RV log-probability path.
//
// Work decomposition: ONE WARP owns one sample at a time.  The 32 lanes are
// 32 epochs; the warp walks the epoch axis in strides of 32*W (W = 2 epochs per lane in
// flight for ILP), looping over the planets innermost.  Consequences:
//   * every per-(sample, planet) quantity - eccentricity, the solver's iteration counts, the
//     circular-orbit branch - is warp-uniform: no divergence anywhere in the hot loop;
//   * the chi^2 + log-det sum of a sample is 32 lane-partials (fixed epoch order) folded by one
//     xor-butterfly: the bits of out[s] depend on (theta[s], epochs) only, never on S, the
//     grid, the shard or the GPU count;
//   * per-sample constants live in a small per-warp shared-memory record and are read back as
//     broadcasts; the epoch arrays are staged once per CTA into shared memory by a TMA bulk
//     copy (cp.async.bulk -> UBLKCP).
#pragma once
#include <cuda_runtime.h>

#include "rvlp_math.cuh"

namespace rvlp {

#ifndef RVLP_W
#define RVLP_W 4
#endif
#ifndef RVLP_GSS_DIV
#define RVLP_GSS_DIV 2   // guided self-scheduling: a grab is (samples left) / (RVLP_GSS_DIV x warps), see logprob_kernel
#endif
constexpr int kGssUnits = 4096;   // least work (sample x epoch x planet units) in a grab of the guided schedule
#ifndef RVLP_MIN_BLOCKS
#define RVLP_MIN_BLOCKS 2
#endif
#ifndef RVLP_THREADS
#define RVLP_THREADS 256
#endif
constexpr int kThreads = RVLP_THREADS;
constexpr int kWarps = kThreads / 32;
#ifndef RVLP_WP
#define RVLP_WP 2
#endif
// Bit mask: software-pipelined sample evaluation (see sample_chi_pipelined) for samples whose planets all
// take the lite plan (bit 0) / all have e <= 0.97 (bit 1).  Measured on B200 (profiles/r01_sweep5/6.log):
// with the polynomial sincos pipelining helped the mixed-plan high-e workload (+8% on config 4) and cost 9% on
// the all-lite config 3; with the table-based sincos (r01_sweep7.log) the plain loop wins on both: default off.
#ifndef RVLP_PIPELINE
#define RVLP_PIPELINE 0
#endif
#ifndef RVLP_STAGGER_NS        // experiment: delay every other warp pair at kernel start
#define RVLP_STAGGER_NS 0
#endif
constexpr int kW = RVLP_W;        // epochs per lane in flight (ILP) in the generic path
constexpr int kWP = RVLP_WP;      // same, in the software-pipelined paths
constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
constexpr int kPadTo = 32 * (kW / cgcd(kW, kWP) * kWP);   // n_pad granularity: whole lane groups for both paths
constexpr int kG = 4;             // samples whose prologue a warp does together
constexpr int kPlanetRec = 16;    // doubles per planet in the sample record
constexpr double kLog2Pi = 1.8378770664093453;   // np.log(2*np.pi), fit.py:3595

// Device view of a context: descriptor tables + resident epoch arrays (all device pointers).
struct DevProblem {
  int n_planets, par, n_inst, ndim, n_priors, n_hyper, n_model, n_epochs, n_pad;
  int epochs_global;   // 1: too many epochs for shared memory - the kernels read them from global memory (L1 / L2)
  int batch_cap;       // samples a warp can take through the prologue together (>= kG; sizes the per-warp scratch)
  int gss_min;         // smallest grab of K1's guided schedule: about kGssUnits units of work (1 sample at config 3)
  double t0, jacobian, renorm;
  const int32_t* src_col;
  const double* src_const;
  const rvlp_prior* priors;
  const double* epochs;   // [t | vel | err2] x n_pad doubles, then n_pad int32 instrument ids
};

// Fused all-gather (multi-GPU): K1 stores a batch's log-probabilities straight into the gathered [S]
// vector of EVERY rank - p[i] is rank i's buffer, mapped into this process by CUDA IPC; one coalesced NVLink peer
// store per rank and batch - at rows off + s, instead of a local vector that a collective copies afterwards.
constexpr int kMaxPeers = 8;
struct PeerOut {
  int n;                  // 0: not in use
  long long off;          // this rank's first row in the gathered vector
  double* p[kMaxPeers];
};

constexpr int kHdr = 6;           // sample record header doubles
__host__ __device__ inline int sample_rec_doubles(int n_planets, int n_inst) {
  return kHdr + 2 * n_inst + kPlanetRec * n_planets;
}
// sample record: [0] lp  [1] flags  [2] gd  [3] gdd  [4] lhp  [5] sum_k C_k  [6..) gamma[n_inst]  jit2[n_inst]  planets
// planet record: n tp e A B C w K | tol plan(bits) P Kraw e w tp invalid
enum { F_JIT = 1, F_PLANET = 2, F_PRIOR = 4, F_HYPER = 8, F_SKIP = 16,
       F_CLS_LITE = 32,   // every planet 0 < e <= 0.65: pipelined (1 fp32 step, lite fp64 step)
       F_CLS_FULL = 64 }; // every planet 0 < e <= 0.97: pipelined (2 fp32 steps, 4th-order fp64 step)

struct SmemLayout {
  int off_t, off_v, off_e2, off_inst, off_priors, off_srccol, off_srcconst, off_scratch, off_pv, total;
};

__host__ __device__ inline SmemLayout smem_layout(const DevProblem& P) {
  SmemLayout L;
  int o = 0;
  const int ns = P.epochs_global ? 0 : P.n_pad;           // epochs staged in shared memory
  L.off_t = o; o += ns * 8;
  L.off_v = o; o += ns * 8;
  L.off_e2 = o; o += ns * 8;
  L.off_inst = o; o += ns * 4;
  o = (o + 15) & ~15;
  L.off_priors = o; o += P.n_priors * (int)sizeof(rvlp_prior);
  L.off_srcconst = o; o += (P.n_model + P.n_hyper) * 8;
  L.off_srccol = o; o += (P.n_model + P.n_hyper) * 4;
  o = (o + 15) & ~15;
  const int cap = P.batch_cap > kG ? P.batch_cap : kG;
  L.off_scratch = o; o += kWarps * cap * sample_rec_doubles(P.n_planets, P.n_inst) * 8;
  L.off_pv = o; o += kWarps * cap * P.n_priors * 8;       // prior values of a batch, one per (sample, prior)
  L.total = o + 16;   // + mbarrier
  return L;
}

// ------------------------------------------------------------------ TMA bulk copy helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  // try_wait suspends the warp in hardware for up to the hinted time instead of spinning through issue slots
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase), "r"(100000u)
      : "memory");
}

// Stage the resident epoch arrays + descriptor tables into shared memory (once per CTA).
__device__ __forceinline__ void stage_problem(const DevProblem& P, const SmemLayout& L, unsigned char* smem) {
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.total - 16);
  const uint32_t epoch_bytes = (uint32_t)(P.n_pad * 28);   // 3 x 8 + 4 per epoch, n_pad % 64 == 0
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && !P.epochs_global) {
    mbar_expect_tx(bar, epoch_bytes);
    bulk_g2s(smem + L.off_t, P.epochs, epoch_bytes, bar);
  }
  // small tables: plain loads
  {
    const int nw = P.n_priors * (int)(sizeof(rvlp_prior) / 8);
    const double* src = reinterpret_cast<const double*>(P.priors);
    double* dst = reinterpret_cast<double*>(smem + L.off_priors);
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    const int nm = P.n_model + P.n_hyper;
    double* dc = reinterpret_cast<double*>(smem + L.off_srcconst);
    int* di = reinterpret_cast<int*>(smem + L.off_srccol);
    for (int i = threadIdx.x; i < nm; i += blockDim.x) {
      dc[i] = P.src_const[i];
      di[i] = P.src_col[i];
    }
  }
  if (!P.epochs_global) mbar_wait(bar, 0);
  __syncthreads();
}

// Per-call overrides of model parameters (Fitter._resolve_freeze_params, fit.py:2586-2688: `params.update(
// resolved_freeze)` for every sample): entry i pins model parameter idx[i] to val[i] for this launch only by
// patching the CTA's shared-memory copy of the source table.
struct FrozenParams {
  int n;
  int idx[RVLP_MAX_FROZEN];
  double val[RVLP_MAX_FROZEN];
};
__device__ __forceinline__ void apply_frozen(const FrozenParams& F, const SmemLayout& L, unsigned char* smem) {
  if ((int)threadIdx.x < F.n) {
    reinterpret_cast<double*>(smem + L.off_srcconst)[F.idx[threadIdx.x]] = F.val[threadIdx.x];
    reinterpret_cast<int*>(smem + L.off_srccol)[F.idx[threadIdx.x]] = -1;
  }
  __syncthreads();
}

// ------------------------------------------------------------------ per-sample prologue
struct Tables {
  const double* t;
  const double* v;
  const double* e2;
  const int* inst;
  const rvlp_prior* priors;
  const double* src_const;
  const int* src_col;
};

// GE (compile time): the epoch arrays stay in global memory (P.epochs_global) instead of shared memory; a
// template parameter rather than a run-time select so that the common path keeps its LDS loads.
template <bool GE = false>
__device__ __forceinline__ Tables tables_of(const DevProblem& P, const SmemLayout& L, unsigned char* smem) {
  Tables T;
  if (GE) {
    T.t = P.epochs;
    T.v = P.epochs + P.n_pad;
    T.e2 = P.epochs + 2 * (size_t)P.n_pad;
    T.inst = reinterpret_cast<const int*>(P.epochs + 3 * (size_t)P.n_pad);
  } else {
    T.t = reinterpret_cast<const double*>(smem + L.off_t);
    T.v = reinterpret_cast<const double*>(smem + L.off_v);
    T.e2 = reinterpret_cast<const double*>(smem + L.off_e2);
    T.inst = reinterpret_cast<const int*>(smem + L.off_inst);
  }
  T.priors = reinterpret_cast<const rvlp_prior*>(smem + L.off_priors);
  T.src_const = reinterpret_cast<const double*>(smem + L.off_srcconst);
  T.src_col = reinterpret_cast<const int*>(smem + L.off_srccol);
  return T;
}

__device__ __forceinline__ double model_param(const Tables& T, const double* row, int i) {
  const int c = T.src_col[i];
  return c >= 0 ? row[c] : T.src_const[i];
}

// Phase A: one lane per (sample, planet): conversion, validity, Kepler constants.
// Phase B: one lane per sample: jitter check, gamma / jitter^2, priors in order.
__device__ __forceinline__ void sample_prologue(const DevProblem& P, const Tables& T, const double* theta,
                                                int64_t s0, int64_t S, double* scratch, int rec, int lane,
                                                bool with_priors, int nb, double* pv) {
  const int npl = P.n_planets;
  for (int task = lane; task < nb * npl; task += 32) {
    const int g = task / npl, k = task - g * npl;
    const int64_t s = s0 + g;
    if (s >= S) continue;
    const double* row = theta + s * P.ndim;
    double in[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) in[q] = model_param(T, row, 5 * k + q);
    const DefaultPars d = to_default(P.par, in);
    double* pr = scratch + g * rec + kHdr + 2 * P.n_inst + k * kPlanetRec;
    PlanetConst pc = planet_const(d);
    const SolverPlan plan = plan_for(d.e);
    pr[0] = pc.n; pr[1] = pc.tp; pr[2] = pc.e; pr[3] = pc.A; pr[4] = pc.B; pr[5] = pc.C;
    pr[6] = pc.w; pr[7] = pc.K; pr[8] = plan.tol;
    pr[9] = __hiloint2double(plan.n64, plan.n32);
    pr[10] = d.P; pr[11] = d.K; pr[12] = d.e; pr[13] = d.w; pr[14] = d.tp;
    pr[15] = d.invalid ? 1.0 : 0.0;
  }
  __syncwarp();
  if (lane < nb && s0 + lane < S) {
    const int g = lane;
    const double* row = theta + (s0 + g) * P.ndim;
    double* sr = scratch + g * rec;
    const int i_gd = 5 * npl, i_g = i_gd + 2, i_jit = i_g + P.n_inst;
    int flags = 0;
    sr[2] = model_param(T, row, i_gd);
    sr[3] = model_param(T, row, i_gd + 1);
    for (int j = 0; j < P.n_inst; ++j) {
      const double jit = model_param(T, row, i_jit + j);
      if (jit < 0) flags |= F_JIT;                          // fit.py:3465-3468
      sr[kHdr + j] = model_param(T, row, i_g + j);
      sr[kHdr + P.n_inst + j] = jit * jit;                     // fit.py:3654
    }
    const double* planets = sr + kHdr + 2 * P.n_inst;
    double csum = 0.0;
    bool all_lite = npl > 0, all_full = npl > 0;
    for (int k = 0; k < npl; ++k) {
      if (planets[k * kPlanetRec + 15] != 0.0) flags |= F_PLANET;
      csum += planets[k * kPlanetRec + 5];                  // K e cos w of every planet (model.py:170)
      const double e = planets[k * kPlanetRec + 2];
      all_lite = all_lite && (e > 0.0) && (e <= 0.65);
      all_full = all_full && (e > 0.0) && (e <= 0.97);
    }
    sr[5] = csum;
    if (all_lite) flags |= F_CLS_LITE;
    else if (all_full) flags |= F_CLS_FULL;
    if (P.n_hyper) {                                        // gp.py:98-108
      for (int k = 0; k < 4; ++k) {
        const double h = model_param(T, row, P.n_model + k);
        if (!(fabs(h) <= 1.79769313486231570e308) || h <= 0) flags |= F_HYPER;
      }
    }
    sr[1] = __hiloint2double(0, flags);
  }
  __syncwarp();
  if (with_priors) {
    // Priors: one lane per (sample, prior) pair instead of one lane walking a sample's priors one after the other
    // (29 sequential log-pdfs per lane at config 3 were most of the prologue's instructions), values parked in
    // shared memory, then summed per sample IN THE REFERENCE'S ORDER (fit.py:3685-3691) - same bits as before.
    const int np = P.n_priors;
    for (int task = lane; task < nb * np; task += 32) {
      const int g = task / np, j = task - g * np;
      if (s0 + g >= S) continue;
      const double* row = theta + (s0 + g) * P.ndim;
      const double* planets = scratch + g * rec + kHdr + 2 * P.n_inst;
      const rvlp_prior& pr = T.priors[j];
      double x;
      if (pr.target == RVLP_TARGET_COLUMN) x = row[pr.index];
      else x = planets[pr.index * kPlanetRec + 9 + pr.target];   // P K e w tp at [10..14]
      pv[task] = prior_logpdf(pr, x);
    }
    __syncwarp();
  }
  if (lane < nb && s0 + lane < S) {
    const int g = lane;
    double* sr = scratch + g * rec;
    int flags = __double2loint(sr[1]);
    double lp = 0.0, lhp = 0.0;
    if (with_priors) {
      for (int j = 0; j < P.n_priors; ++j) {
        const double v = pv[g * P.n_priors + j];
        if (T.priors[j].is_hyper) lhp += v; else lp += v;
      }
      // a failed Tc->Tp conversion makes a derived Tp NaN, i.e. lp non-finite: fit.py:3478-3482
      if (!(fabs(lp) <= 1.79769313486231570e308)) flags |= F_PRIOR;
      if (!(fabs(lhp) <= 1.79769313486231570e308)) flags |= F_PRIOR;
    }
    sr[0] = lp;
    sr[1] = __hiloint2double(0, flags);
    sr[4] = lhp;
  }
  __syncwarp();
}

// RV of every planet + trend at W epochs per lane (fit.py:3613-3636 / model.py:639-664).
template <int W>
__device__ __forceinline__ void model_rv(const DevProblem& P, const double* sr, const double (&tt)[W],
                                         double (&rv)[W], int only_planet, bool with_trend) {
  const double2* planets = reinterpret_cast<const double2*>(sr + kHdr + 2 * P.n_inst);
  const double c0 = only_planet < 0 ? sr[5]
                                    : (only_planet < P.n_planets ? sr[kHdr + 2 * P.n_inst + only_planet * kPlanetRec + 5] : 0.0);
#pragma unroll
  for (int j = 0; j < W; ++j) rv[j] = c0;
  for (int k = 0; k < P.n_planets; ++k) {
    if (only_planet >= 0 && k != only_planet) continue;
    const double2* pr = planets + k * (kPlanetRec / 2);     // 16-byte aligned: LDS.128 broadcasts
    const double2 a = pr[0], b = pr[1], c = pr[2], d = pr[3], e = pr[4];
    PlanetConst pc;
    pc.n = a.x; pc.tp = a.y; pc.e = b.x; pc.A = b.y; pc.B = c.x; pc.C = c.y; pc.w = d.x; pc.K = d.y;
    SolverPlan plan;
    plan.tol = e.x;
    plan.n32 = __double2loint(e.y);
    plan.n64 = __double2hiint(e.y);
    planet_rv_add<W>(pc, plan, tt, rv);
  }
  if (with_trend) {                                          // model.py:483-509
    const double gd = sr[2], gdd = sr[3];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const double dt = tt[j] - P.t0;
      rv[j] = fma(gdd, dt * dt, fma(gd, dt, rv[j]));         // exact zeros when gd / gdd == 0
    }
  }
}

// Lane-local accumulators of one sample's chi^2 + log-det sum (fixed epoch order per lane).
// sum_i ln var_i is kept as a running product of the variances' mantissas plus an integer exponent sum:
// one log per lane per sample instead of one per epoch.
struct ChiAcc {
  double chi = 0.0, prodm = 1.0, slow = 0.0;
  int exsum = 0, cnt = 0;
};

// Consumes the model RV (planets + trend) at epochs base + j*32 + lane (fit.py:3642-3658).
template <int W>
__device__ __forceinline__ void chi_epilogue(const DevProblem& P, const Tables& T, const double* sr, int base,
                                             int lane, const double (&rv)[W], ChiAcc& a) {
  double var[W], res[W];
  bool odd = false;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const int idx = base + j * 32 + lane;
    const int in = T.inst[idx];
    const double tot = rv[j] + sr[kHdr + in];             // fit.py:3642-3644
    var[j] = T.e2[idx] + sr[kHdr + P.n_inst + in];        // fit.py:3654
    res[j] = tot - T.v[idx];
    const int h = __double2hiint(var[j]);
    const bool normal = (unsigned)(h - 0x00100000) < 0x7fe00000u;   // positive, normal, finite
    const bool live = idx < P.n_epochs;
    odd |= live && !normal;
    if (live && normal) {                                 // fit.py:3655-3658
#if RVLP_OPT & 2
      a.chi = fma(res[j] * res[j], rcp64_3(var[j]), a.chi);
#else
      a.chi = fma(res[j] * res[j], rcp64(var[j]), a.chi);
#endif
      a.prodm *= __hiloint2double((h & 0x000fffff) | 0x3ff00000, __double2loint(var[j]));
      a.exsum += (h >> 20) - 1023;
      a.cnt += 1;
    }
  }
  if (__any_sync(0xffffffffu, odd)) {                     // var == 0, denormal, inf, NaN: IEEE path
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int idx = base + j * 32 + lane;
      const int h = __double2hiint(var[j]);
      if (idx < P.n_epochs && !((unsigned)(h - 0x00100000) < 0x7fe00000u))
        a.slow += res[j] * res[j] / var[j] + (kLog2Pi + log(var[j]));
    }
  }
}

// prodm is a running product of mantissas in [1, 2): it overflows after ~1023 factors per lane (n_epochs >= 32 768,
// i.e. only the epochs-in-global-memory shape).  Folding its exponent into exsum every <= 512 factors keeps it a
// normal number for any epoch count; scaling by a power of two is exact, so ln(prod var) is unchanged.
__device__ __forceinline__ void chi_renorm(ChiAcc& a) {
  const int h = __double2hiint(a.prodm);
  a.exsum += (h >> 20) - 1023;
  a.prodm = __hiloint2double((h & 0x000fffff) | 0x3ff00000, __double2loint(a.prodm));
}

__device__ __forceinline__ double chi_finish(const ChiAcc& a) {
  // prodm is a product of mantissas in [1, 2): positive and normal unless it overflowed (> 1023 epochs per lane)
  const double lm = __double2hiint(a.prodm) < 0x7ff00000 ? log_pos_normal(a.prodm) : log(a.prodm);
  double acc = a.chi + ((double)a.cnt * kLog2Pi + fma((double)a.exsum, 0.6931471805599453, lm)) + a.slow;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

struct HotPlanet { double n, tp, e, A, B, C, tol; };
__device__ __forceinline__ HotPlanet load_hot(const double2* planets, int k) {
  const double2* pr = planets + k * (kPlanetRec / 2);
  const double2 a = pr[0], b = pr[1], c = pr[2], e = pr[4];
  HotPlanet h;
  h.n = a.x; h.tp = a.y; h.e = b.x; h.A = b.y; h.B = c.x; h.C = c.y; h.tol = e.x;
  return h;
}

// Software-pipelined evaluation of one sample whose planets all share a compile-time solver plan.
// The (epoch group, planet) double loop is flattened; in every iteration stage A (fp32 / MUFU pipes) of the
// NEXT (group, planet) is issued next to stage B (fp64 pipe) of the current one.  They are independent, so
// each warp's instruction stream mixes the two pipes instead of alternating long single-pipe phases (which
// leaves the 2-cycle-issue FP64 pipe's gaps unfilled when all warps of an SM run in lock-step).
template <int W, int N32, int N64>
__device__ __forceinline__ void sample_chi_pipelined(const DevProblem& P, const Tables& T, const double* sr, int lane,
                                                     double tol, ChiAcc& acc) {
  const int npl = P.n_planets;
  const int G = P.n_pad / (32 * W);
  const double2* planets = reinterpret_cast<const double2*>(sr + kHdr + 2 * P.n_inst);
  const double c0 = sr[5], gd = sr[2], gdd = sr[3];
  double tt[W], M[W], rv[W];
#pragma unroll
  for (int j = 0; j < W; ++j) tt[j] = T.t[j * 32 + lane];
  HotPlanet pc = load_hot(planets, 0);
  bool bigA = false;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    M[j] = mean_anomaly(pc.n, tt[j], pc.tp);
    bigA |= anomaly_is_big(M[j]);
    rv[j] = c0;
  }
  StarterOut<W> A;
  kepler_stage_a<W, N32>(M, pc.e, N32, A);
  int k = 0, g = 0;
  const int Q = G * npl;
  for (int q = 0; q < Q; ++q) {
    int kn = k + 1, gn = g;
    if (kn == npl) { kn = 0; gn = g + 1; }
    const int gl = gn < G ? gn : G - 1;                    // the last iteration's look-ahead is discarded
    double tn[W], Mn[W];
    const HotPlanet pn = load_hot(planets, kn);
    bool bigN = false;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      tn[j] = T.t[gl * 32 * W + j * 32 + lane];
      Mn[j] = mean_anomaly(pn.n, tn[j], pn.tp);
      bigN |= anomaly_is_big(Mn[j]);
    }
    StarterOut<W> An;
    kepler_stage_a<W, N32>(Mn, pn.e, N32, An);             // next: fp32 / MUFU

    double cE[W], sE[W], dl[W], ri[W];
    kepler_stage_b<W, N64>(A, pc.e, N64, cE, sE, dl, ri);  // current: fp64
    bool bad = bigA;
#pragma unroll
    for (int j = 0; j < W; ++j) bad |= step_rejected(dl[j], tol);
    if (__any_sync(0xffffffffu, bad)) {                    // warp-uniform, rare
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const double Mj = mean_anomaly(pc.n, tt[j], pc.tp);
        if (step_rejected(dl[j], tol) || anomaly_is_big(Mj)) {
          const CosSin cs = kepler_robust(Mj, pc.e);
          cE[j] = cs.c;
          sE[j] = cs.s;
          ri[j] = 1.0 / (1.0 - pc.e * cs.c);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const double u = fma(-sE[j], pc.B, fma(cE[j], pc.A, -pc.C));
      rv[j] = fma(ri[j], u, rv[j]);
    }
    if (k == npl - 1) {                                    // all planets of this epoch group done
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const double dt = tt[j] - P.t0;                    // model.py:483-509
        rv[j] = fma(gdd, dt * dt, fma(gd, dt, rv[j]));
      }
      chi_epilogue<W>(P, T, sr, g * 32 * W, lane, rv, acc);
#pragma unroll
      for (int j = 0; j < W; ++j) rv[j] = c0;
    }
    A = An;
    pc = pn;
    bigA = bigN;
#pragma unroll
    for (int j = 0; j < W; ++j) tt[j] = tn[j];
    k = kn;
    g = gn;
  }
}

// ------------------------------------------------------------------ K1: log-probability
// Two shapes are compiled: <kW = 4 epochs per lane in flight, 2 CTAs/SM> and <2, 3 CTAs/SM>.  Neither wins
// everywhere (profiles/r01_sweep8.log: W = 4 is 2-4 % ahead on the low-eccentricity 5-planet config, W = 2
// is 10 % ahead on the high-eccentricity config and 4 % on 120-epoch data); rvlp_ctx_autotune times both on
// the caller's own rows.  A lane visits its epochs (lane, lane + 32, ...) in ascending order for either W, so
// the choice never changes a bit of the result.
template <int W, int MB, bool GE, bool PEERS = false>
__global__ void __launch_bounds__(kThreads, MB)
logprob_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, double* __restrict__ out,
               double* __restrict__ ll_out, double* __restrict__ lp_out, int nb,
               unsigned long long* __restrict__ next_batch, PeerOut peers) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);
  stage_problem(P, L, smem);
  const Tables T = tables_of<GE>(P, L, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  const int cap = P.batch_cap > kG ? P.batch_cap : kG;
  double* scratch = reinterpret_cast<double*>(smem + L.off_scratch) + warp * cap * rec;
  double* pv = reinterpret_cast<double*>(smem + L.off_pv) + warp * cap * P.n_priors;
  // nb = samples per prologue batch (1..batch_cap): small launches use 1 so that every warp gets a sample, large
  // ones as many as fill the lanes of the (sample, planet) and (sample, prior) phases; the bits of a sample's
  // result do not depend on it.
  const int64_t gw = (int64_t)blockIdx.x * kWarps + warp, nw = (int64_t)gridDim.x * kWarps;
#if RVLP_STAGGER_NS > 0
  if ((warp >> 2) & 1) __nanosleep(RVLP_STAGGER_NS);
#endif

  // Every warp starts on the batch of `nb` samples at gw * nb.  When the launch has several batches per warp the
  // further ones come from a global SAMPLE counter by guided self-scheduling: a grab is `nb` samples while plenty
  // are left and shrinks to (what is left) / (RVLP_GSS_DIV x warps), down to P.gss_min samples (about 4096 units of work: one
  // sample at config 3, no shrinking at all for 120-epoch data, whose prologue wants full batches), so the launch's
  // tail is one sample's work instead of one batch's (6 samples x 5000 units at config 3: 0.15 ms of a 2.7 ms
  // shard, profiles/r02ap_k1_guided_schedule.log).
  // What is left is estimated from the end of the warp's own previous grab - an upper bound, the counter only
  // moves on.  Small launches use a static stride.  Which warp evaluates a sample, and in a batch of what size,
  // never changes its bits.
  const int64_t s_static = nw * nb;                          // the samples dealt out statically
  const int nmin = P.gss_min < 1 ? 1 : (P.gss_min < nb ? P.gss_min : nb);
  int n = nb;
  for (int64_t s0 = gw * nb; s0 < S;) {
    const int n_cur = n;
    const int64_t s_cur = s0;
    if (next_batch) {
      const int64_t seen = s_cur + n_cur > s_static ? s_cur + n_cur : s_static;
      const int64_t share = (S - seen) / (RVLP_GSS_DIV * nw);
      n = share >= nb ? nb : (share < nmin ? nmin : (int)share);
      unsigned long long t = 0;
      if (lane == 0) t = atomicAdd(next_batch, (unsigned long long)n);
      s0 = s_static + (int64_t)__shfl_sync(0xffffffffu, t, 0);
    } else {
      s0 += s_static;
    }
    sample_prologue(P, T, theta, s_cur, S, scratch, rec, lane, true, n_cur, pv);
    for (int g = 0; g < n_cur; ++g) {
      const int64_t s = s_cur + g;
      if (s >= S) break;
      const double* sr = scratch + g * rec;
      const int flags = __double2loint(sr[1]);
      const double lp = sr[0];
      double ll;
      if (flags & F_PLANET) {
        ll = -INFINITY;                                      // fit.py:3625-3627
      } else if ((flags & (F_JIT | F_PRIOR)) && ll_out == nullptr) {
        ll = 0.0;                                            // result is -inf regardless: skip the work
      } else {
        ChiAcc acc;
        if ((RVLP_PIPELINE & 1) && (flags & F_CLS_LITE)) {
          sample_chi_pipelined<kWP, 1, 0>(P, T, sr, lane, 4.0e-6, acc);
        } else if ((RVLP_PIPELINE & 2) && (flags & F_CLS_FULL)) {
          sample_chi_pipelined<kWP, 2, 1>(P, T, sr, lane, 2.5e-4, acc);
        } else if ((RVLP_OPT & 32) && (flags & F_CLS_LITE)) {
          // every planet takes the lite plan: no circular-orbit test, no plan dispatch, three LDS.128 per planet
          const double2* planets = reinterpret_cast<const double2*>(sr + kHdr + 2 * P.n_inst);
          const double c0 = sr[5], gd = sr[2], gdd = sr[3];
          for (int base = 0; base < P.n_pad; base += 32 * W) {
            double tt[W], rv[W];
#pragma unroll
            for (int j = 0; j < W; ++j) {
              tt[j] = T.t[base + j * 32 + lane];
              rv[j] = c0;
            }
            for (int k = 0; k < P.n_planets; ++k) {
              const double2* pr = planets + k * (kPlanetRec / 2);
              const double2 a = pr[0], b = pr[1], c = pr[2];
              planet_rv_add_lite<W>(a.x, a.y, b.x, b.y, c.x, c.y, tt, rv);
            }
#pragma unroll
            for (int j = 0; j < W; ++j) {
              const double dt = tt[j] - P.t0;                  // model.py:483-509
              rv[j] = fma(gdd, dt * dt, fma(gd, dt, rv[j]));
            }
            chi_epilogue<W>(P, T, sr, base, lane, rv, acc);
            if (GE && ((base / (32 * W)) & 127) == 127) chi_renorm(acc);
          }
        } else {
          for (int base = 0; base < P.n_pad; base += 32 * W) {
            double tt[W], rv[W];
#pragma unroll
            for (int j = 0; j < W; ++j) tt[j] = T.t[base + j * 32 + lane];
            model_rv<W>(P, sr, tt, rv, -1, true);
            chi_epilogue<W>(P, T, sr, base, lane, rv, acc);
            if (GE && ((base / (32 * W)) & 127) == 127) chi_renorm(acc);   // <= 128 W = 512 factors between folds
          }
        }
        ll = -0.5 * chi_finish(acc);
      }
      if (lane == 0) {
        double r;
        if (flags & (F_JIT | F_PRIOR | F_HYPER)) {
          r = -INFINITY;                                     // fit.py:3468, 3480-3482
        } else {
          r = ll + lp;                                       // fit.py:3492-3495
          r += P.jacobian;
          r += P.renorm;
        }
        if (out) out[s] = r;
        if (ll_out) ll_out[s] = ll;
        if (lp_out) lp_out[s] = lp;
        if (PEERS) scratch[g * rec] = r;                     // the record's lp slot is free now: park the result
      }
    }
    __syncwarp();
    if (PEERS) {   // (a separate instantiation: the single-GPU kernel carries none of this)
      // the batch's results - consecutive rows of the gathered vector - go to every rank as ONE coalesced store per
      // rank (lane g = sample g): a sixth of the NVLink transactions of per-sample stores at config 3
      const bool live = lane < n_cur && s_cur + lane < S;
      const double r = live ? scratch[lane * rec] : 0.0;
      for (int i = 0; i < peers.n; ++i)
        if (live) peers.p[i][peers.off + s_cur + lane] = r;
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------ K2: RV matrix (fit.py:2690-2824)
__global__ void __launch_bounds__(kThreads, RVLP_MIN_BLOCKS)
rv_matrix_kernel(DevProblem P, const double* __restrict__ theta, int64_t S,
                 const double* __restrict__ times, int64_t T_n, int component, double* __restrict__ out,
                 FrozenParams frozen, unsigned long long* __restrict__ next_batch) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);
  stage_problem(P, L, smem);
  if (frozen.n) apply_frozen(frozen, L, smem);     // kernel-uniform
  const Tables T = tables_of(P, L, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  double* scratch = reinterpret_cast<double*>(smem + L.off_scratch) + warp * kG * rec;
  const int64_t gw = (int64_t)blockIdx.x * kWarps + warp, nw = (int64_t)gridDim.x * kWarps;
  const int only = component >= 0 ? component : (component == RVLP_RV_TREND ? P.n_planets : -1);
  const bool trend = component < 0;

  const int64_t s_static = nw * kG;                          // guided self-scheduling, as in logprob_kernel
  const int nmin = T_n * (only < 0 ? P.n_planets : 1) >= kGssUnits / 4 ? 1 : kG;   // short rows: whole batches only
  int n = kG;
  for (int64_t s0 = gw * kG; s0 < S;) {
    const int n_cur = n;
    const int64_t s_cur = s0;
    if (next_batch) {
      const int64_t seen = s_cur + n_cur > s_static ? s_cur + n_cur : s_static;
      const int64_t share = (S - seen) / (RVLP_GSS_DIV * nw);
      n = share >= kG ? kG : (share < nmin ? nmin : (int)share);
      unsigned long long t = 0;
      if (lane == 0) t = atomicAdd(next_batch, (unsigned long long)n);
      s0 = s_static + (int64_t)__shfl_sync(0xffffffffu, t, 0);
    } else {
      s0 += s_static;
    }
    sample_prologue(P, T, theta, s_cur, S, scratch, rec, lane, false, n_cur, nullptr);
    for (int g = 0; g < n_cur; ++g) {
      const int64_t s = s_cur + g;
      if (s >= S) break;
      const double* sr = scratch + g * rec;
      const double* planets = sr + kHdr + 2 * P.n_inst;
      bool bad = false;
      for (int k = 0; k < P.n_planets; ++k)
        if ((only < 0 || k == only) && planets[k * kPlanetRec + 15] != 0.0) bad = true;
      double* orow = out + s * T_n;
      for (int64_t base = 0; base < T_n; base += 32 * kW) {
        double tt[kW], rv[kW];
        int64_t idx[kW];
#pragma unroll
        for (int j = 0; j < kW; ++j) {
          idx[j] = base + j * 32 + lane;
          tt[j] = times[idx[j] < T_n ? idx[j] : T_n - 1];
        }
        if (bad) {
#pragma unroll
          for (int j = 0; j < kW; ++j) rv[j] = NAN;
        } else {
          model_rv<kW>(P, sr, tt, rv, only, trend);
        }
#pragma unroll
        for (int j = 0; j < kW; ++j)
          if (idx[j] < T_n) orow[idx[j]] = rv[j];
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ K5: walker-position check (row f-3)
// What Fitter.generate_initial_walker_positions_* and run_mcmc do per candidate row before sampling
// (fit.py:692-725, 884-902, 1048-1062; GP twin fit.py:4500-4540, 4950-4990): _validate_astrophysical_validity
// (fit.py:260-293: every parameter finite, each planet converts + validates, jitter >= 0) then a finite
// log-prior on the converted parameters.  One status word per row instead of one LogPosterior per attempt.
__global__ void __launch_bounds__(kThreads, RVLP_MIN_BLOCKS)
walker_check_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, int32_t* __restrict__ status,
                    double* __restrict__ lp_out, double* __restrict__ lhp_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);
  stage_problem(P, L, smem);
  const Tables T = tables_of(P, L, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  const int64_t gw = (int64_t)blockIdx.x * kWarps + warp, nw = (int64_t)gridDim.x * kWarps;
  // rows per prologue batch: the context's capacity (fills the lanes of the planet / prior phases) when every warp still
  // gets a couple of batches, else kG
  const int cap = P.batch_cap > kG ? P.batch_cap : kG;
  const int nbw = (S + cap - 1) / cap >= 2 * nw ? cap : kG;
  double* scratch = reinterpret_cast<double*>(smem + L.off_scratch) + warp * cap * rec;
  const int64_t n_batches = (S + nbw - 1) / nbw;
  for (int64_t b = gw; b < n_batches; b += nw) {
    const int64_t s0 = b * nbw;
    sample_prologue(P, T, theta, s0, S, scratch, rec, lane, true, nbw,
                    reinterpret_cast<double*>(smem + L.off_pv) + warp * cap * P.n_priors);
    if (lane < nbw && s0 + lane < S) {
      const int64_t s = s0 + lane;
      const double* row = theta + s * P.ndim;
      const double* sr = scratch + lane * rec;
      const int flags = __double2loint(sr[1]);
      int st = 0;
      for (int i = 0; i < P.n_model + P.n_hyper; ++i)          // fit.py:262-265 (free and fixed values alike)
        if (!(fabs(model_param(T, row, i)) <= 1.79769313486231570e308)) st |= RVLP_WALKER_NONFINITE;
      if (flags & F_PLANET) st |= RVLP_WALKER_PLANET;          // fit.py:268-276
      if (flags & F_JIT) st |= RVLP_WALKER_JITTER;             // fit.py:289-293
      if (flags & F_HYPER) st |= RVLP_WALKER_HYPER;            // gp.py:82-108
      const double lp = sr[0], lhp = sr[4];
      if (!(fabs(lp) <= 1.79769313486231570e308)) st |= RVLP_WALKER_PRIOR;        // fit.py:717-720
      if (!(fabs(lhp) <= 1.79769313486231570e308)) st |= RVLP_WALKER_HYPERPRIOR;  // fit.py:4534-4537
      status[s] = st;
      if (lp_out) lp_out[s] = lp;
      if (lhp_out) lhp_out[s] = lhp;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ information criteria (fit.py:1361-1554)
// Fitter.calculate_chi2 works backwards from the log-likelihood: penalty = sum_i ln(2 pi (sigma_i^2 + jit_inst(i)^2)),
// chi2 = -2 ll - penalty (fit.py:1485-1500); AICc = 2k - 2 ll + (2k^2 + 2k) / (n - k - 1) (fit.py:1524-1529);
// BIC = k ln n - 2 ll (fit.py:1553-1554).  One warp per row: the penalty depends on the row's jitter values only;
// lanes walk the epochs in a fixed order and fold with one butterfly (bits independent of the grid).
__global__ void __launch_bounds__(256)
info_criteria_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, const double* __restrict__ ll,
                     double* __restrict__ chi2, double* __restrict__ aicc, double* __restrict__ bic, double two_k,
                     double aicc_corr, double k_ln_n) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const double* e2 = P.epochs + 2 * (size_t)P.n_pad;
  const int* inst = reinterpret_cast<const int*>(P.epochs + 3 * (size_t)P.n_pad);
  const int i_jit = 5 * P.n_planets + 2 + P.n_inst;
  for (int64_t s = gw; s < S; s += nw) {
    const double* row = theta + s * P.ndim;
    double pen = 0.0;
    if (chi2) {
      for (int i = lane; i < P.n_epochs; i += 32) {
        const int q = i_jit + inst[i];
        const int c = P.src_col[q];
        const double jit = c >= 0 ? row[c] : P.src_const[q];
        pen += log((2 * PI_D) * (e2[i] + jit * jit));
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pen += __shfl_xor_sync(0xffffffffu, pen, o);
    }
    if (lane == 0) {
      const double l = ll[s];
      if (chi2) chi2[s] = -2 * l - pen;
      if (aicc) aicc[s] = (two_k - 2 * l) + aicc_corr;
      if (bic) bic[s] = k_ln_n - 2 * l;
    }
  }
}

// ------------------------------------------------------------------ small stateless kernels
// model.py:173-243 on raw mean anomalies, one (e, K, w) for the whole array.
__global__ void kepler_rv_kernel(const double* __restrict__ M, int64_t n, double e, double K, double w,
                                 double* __restrict__ rv) {
  DefaultPars d;
  d.P = 2 * PI_D; d.K = K; d.e = e; d.w = w; d.tp = 0.0; d.conv_error = false; d.invalid = false;
  PlanetConst pc = planet_const(d);
  pc.n = 1.0;                      // M = 1.0 * (M - 0.0) is exact
  const SolverPlan plan = plan_for(e);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double tt[1] = {M[i]}, r[1] = {pc.C};
    planet_rv_add<1>(pc, plan, tt, r);
    rv[i] = r[0];
  }
}

// model.py:259-275 + 329-354 for one planet given default-space parameters (validated on host side
// of the ABI via convert kernel); accumulate != 0 adds into rv (Star.radial_velocity, model.py:659-662).
__global__ void planet_rv_kernel(DefaultPars d, const double* __restrict__ t, int64_t n,
                                 double* __restrict__ rv, int accumulate) {
  const PlanetConst pc = planet_const(d);
  const SolverPlan plan = plan_for(d.e);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double tt[1] = {t[i]}, r[1] = {pc.C};
    planet_rv_add<1>(pc, plan, tt, r);
    rv[i] = accumulate ? rv[i] + r[0] : r[0];
  }
}

__global__ void trend_rv_kernel(double gd, double gdd, double t0, const double* __restrict__ t, int64_t n,
                                double* __restrict__ rv, int accumulate) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double dt = t[i] - t0;
    double tr = 0.0;
    if (gd != 0) tr += gd * dt;
    if (gdd != 0) tr += gdd * (dt * dt);
    rv[i] = accumulate ? rv[i] + tr : tr;
  }
}

__global__ void convert_kernel(int par, const double* __restrict__ in, int64_t n, double* __restrict__ out,
                               int32_t* __restrict__ valid) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) v[q] = in[i * 5 + q];
    const DefaultPars d = to_default(par, v);
    out[i * 5 + 0] = d.P; out[i * 5 + 1] = d.K; out[i * 5 + 2] = d.e; out[i * 5 + 3] = d.w;
    out[i * 5 + 4] = d.tp;
    if (valid) valid[i] = d.invalid ? 0 : 1;
  }
}

__global__ void prior_kernel(rvlp_prior pr, const double* __restrict__ x, int64_t n, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = prior_logpdf(pr, x[i]);
}

// Dependent-free DFMA loop: 8 independent chains per thread, 2 flops per DFMA, in the fastest operand form
// measured on B200 (tools/pipe_probe2.cu: DFMA R,R,R,c issues every 2.08 cycles per SMSP; with three register
// operands it drops to one per 3.06 cycles - register-file bandwidth).
// Cross-rank barrier after a launch with PeerOut (one warp; thread i talks to rank i).  The launch before it on the
// stream has completed, so its peer stores are performed; thread i then publishes `epoch` in rank i's flag word for
// this rank (release, system scope) and waits until rank i's epoch has arrived in this rank's own flag block
// (acquire).  Kernels queued after it on the stream may read the gathered vector.  The wait is bounded by
// timeout_ns of the global timer: a rank that never arrives sets flags_self[kMaxPeers] = 1 instead of hanging the GPU.
__global__ void peer_barrier_kernel(PeerOut flags, int my_rank, unsigned long long epoch, unsigned long long timeout_ns) {
  const int i = threadIdx.x;
  if (i >= flags.n) return;
  unsigned long long* theirs = reinterpret_cast<unsigned long long*>(flags.p[i]) + my_rank;
  volatile unsigned long long* mine = reinterpret_cast<unsigned long long*>(flags.p[my_rank]) + i;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long v, t;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      reinterpret_cast<unsigned long long*>(flags.p[my_rank])[kMaxPeers] = 1ull;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  const double y = a + 1e-12 * threadIdx.x;   // a per-thread register multiplier
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, y, b); x1 = fma(x1, y, b); x2 = fma(x2, y, b); x3 = fma(x3, y, b);
      x4 = fma(x4, y, b); x5 = fma(x5, y, b); x6 = fma(x6, y, b); x7 = fma(x7, y, b);
    }
  }
  out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace rvlp
