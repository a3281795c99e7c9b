// rvlp_ws.cuh — K1, warp-specialised shape: logprob_ws_kernel.
//
// Why: the stage timings of logprob_kernel ADD UP (a stage compiled out at a time, c3 at 2e5 samples: skeleton
// 1.54 ms + fp32/MUFU starter 1.34 ms + fp64 stage 2.27 ms = 5.15 ms = the full kernel).  Each stage runs at its
// own pipe's limit (the starter saturates the XU pipe, the fp64 stage the FP64 pipe), but the four warps of an SM
// sub-partition fall into the same stage together, so the two pipes take turns instead of working at once.
//
// Here the stages live in DIFFERENT warps.  A CTA is NP producer warps and NP * NC consumer warps; producer p and
// its consumers p + NP, p + 2 NP, ... sit on the same SM sub-partition (warp id mod 4).  The starter is about a
// quarter of the work but a lone producer warp cannot saturate the XU pipe, hence two producers and four
// consumers per sub-partition (NP = 8, NC = 2, 24 warps, one CTA per SM) with the registers moved from the
// producers to the consumers by setmaxnreg (48 / 104 per thread).
// Each (producer, consumer) LINK is an independent stream of samples.  The consumer owns the stream: ticket,
// per-sample prologue (K4) one batch ahead into the link's second record set, then per (sample, epoch group,
// planet) the fp64 stage, the RV sum, trend, the chi^2 / log-det epilogue, and the store of the sample's result.
// The producer reads the same records and computes, for every (sample, epoch group, planet), the mean anomaly,
// its reduction and the fp32/MUFU starter (planet_starter), written to the link's shared-memory ring as
// {m, E0 | sign} per epoch; it serves its links round-robin, one ring slot each per turn.  (The prologue stays
// out of the producer because its libm slow paths are real calls, and ptxas 12.9 crashes on a call in a region
// whose registers were cut by setmaxnreg.dec.)  Rings and record sets are guarded by mbarriers (32 arrivals:
// every lane publishes or releases its own words).  The XU pipe and the FP64 pipe are then busy at the same
// time by construction.
//
// Same device functions on the same values as logprob_kernel -> the same bits (tested), so rvlp_ctx_autotune
// can choose between the shapes freely.
#pragma once
#include "rvlp_kernels.cuh"

namespace rvlp {

constexpr int kRecSets = 2;               // sample-record sets per link (the producer runs one batch ahead)
#ifndef RVLP_WS_SLOTS
#define RVLP_WS_SLOTS 3
#endif
constexpr int kGws = 2;                   // samples per prologue batch in this shape (record sets cost shared memory)
constexpr int kSlots = RVLP_WS_SLOTS;     // ring depth per link

struct WsSmem {
  int off_recs, off_ring, off_bars, off_batch, total;
};
__host__ __device__ inline int ws_bars_per_link() { return 2 * kSlots + 2 * kRecSets; }
__host__ __device__ inline WsSmem ws_smem(const DevProblem& P, const SmemLayout& L, int W, int NP, int NC) {
  WsSmem G;
  const int links = NP * NC;
  // the record sets take the place of logprob_kernel's per-warp scratch (unused here)
  int o = L.off_scratch;
  G.off_recs = o; o += links * kRecSets * kGws * sample_rec_doubles(P.n_planets, P.n_inst) * 8;
  if (o < L.total) o = L.total;
  o = (o + 15) & ~15;
  G.off_ring = o; o += links * kSlots * W * 32 * 12;       // per epoch: m (8 B) + packed E0 | sign (4 B)
  G.off_bars = o; o += links * ws_bars_per_link() * 8;
  G.off_batch = o; o += links * kRecSets * 8;              // first sample of the batch (< 0: no more work)
  G.total = o;
  return G;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// What a producer remembers about one of its links.
struct WsStream {
  uint32_t set_i, slot_i;
  int64_t b, s0;
  int g, base, k, set;
  bool have, done;
};

// NP producer warps (a multiple of 4: whole warpgroups), NC consumers per producer; RP / RC > 0: registers per
// thread after setmaxnreg (producers give registers back, consumers take them), 0: keep the launch allocation.
template <int W, int NP, int NC, int MB, int RP, int RC, bool GE>
__global__ void __launch_bounds__(32 * NP * (1 + NC), MB)
logprob_ws_kernel(DevProblem P, const double* __restrict__ theta, int64_t S, double* __restrict__ out,
                  double* __restrict__ ll_out, double* __restrict__ lp_out, int nb,
                  unsigned long long* __restrict__ next_batch) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout L = smem_layout(P);
  const WsSmem G = ws_smem(P, L, W, NP, NC);
  constexpr int kProducers = NP;
  constexpr int kLinks = kProducers * NC;
  constexpr int kBars = 2 * kSlots + 2 * kRecSets;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool producer = warp < kProducers;
  for (int i = threadIdx.x; i < kLinks * kBars; i += blockDim.x)
    mbar_init(reinterpret_cast<uint64_t*>(smem + G.off_bars) + i, 32);
  stage_problem(P, L, smem);              // syncs the CTA after the mbarrier inits
  const Tables T = tables_of<GE>(P, L, smem);
  const int rec = sample_rec_doubles(P.n_planets, P.n_inst);
  const int64_t n_batches = (S + nb - 1) / nb;
  const int64_t n_streams = (int64_t)gridDim.x * kLinks;
  const int npl = P.n_planets;
  // link l = c * kProducers + p  (producer p, its c-th consumer = warp kProducers + l)
  auto link_bars = [&](int l) { return reinterpret_cast<uint64_t*>(smem + G.off_bars) + l * kBars; };
  auto link_recs = [&](int l) { return reinterpret_cast<double*>(smem + G.off_recs) + (size_t)l * kRecSets * kGws * rec; };
  auto link_batch = [&](int l) { return reinterpret_cast<long long*>(smem + G.off_batch) + l * kRecSets; };
  auto link_ring = [&](int l) { return smem + G.off_ring + (size_t)l * kSlots * W * 32 * 12; };

  if (producer) {
    if constexpr (RP > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RP));
    WsStream st[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      st[c].set_i = 0; st[c].slot_i = 0; st[c].have = false; st[c].done = false;
      st[c].b = 0;
      st[c].s0 = 0; st[c].g = 0; st[c].base = 0; st[c].k = 0; st[c].set = 0;
    }
    // one turn for link l: advance its stream to the next (sample, epoch group, planet) that needs a starter and
    // publish it (or finish the stream).  Every decision is warp-uniform.
    auto turn = [&](WsStream& s, int l) {
      uint64_t* bars = link_bars(l);
      uint64_t *full = bars, *empty = bars + kSlots, *rec_full = bars + 2 * kSlots, *rec_free = rec_full + kRecSets;
      for (;;) {
        if (!s.have) {                                       // next batch of this link: records from the consumer
          s.set = s.set_i % kRecSets;
          mbar_wait(rec_full + s.set, (s.set_i / kRecSets) & 1);
          ++s.set_i;
          s.s0 = link_batch(l)[s.set];
          if (s.s0 < 0) { s.done = true; return; }
          s.have = true; s.g = 0; s.base = 0; s.k = 0;
        }
        if (s.g >= nb || s.s0 + s.g >= S) {                  // batch done: the consumer may reuse the record set
          mbar_arrive(rec_free + s.set);
          s.have = false;
          continue;
        }
        const double* sr = link_recs(l) + s.set * kGws * rec + s.g * rec;
        const int flags = __double2loint(sr[1]);
        if ((flags & F_PLANET) || ((flags & (F_JIT | F_PRIOR)) && ll_out == nullptr)) {   // no likelihood needed
          ++s.g; s.base = 0; s.k = 0;
          continue;
        }
        if (s.k >= npl) { s.k = 0; s.base += 32 * W; }
        if (s.base >= P.n_pad) { ++s.g; s.base = 0; s.k = 0; continue; }
        const double2* pr = reinterpret_cast<const double2*>(sr + kHdr + 2 * P.n_inst) + s.k * (kPlanetRec / 2);
        const double2 a = pr[0], bq = pr[1], e4 = pr[4];
        ++s.k;
        if (bq.x == 0) continue;                             // circular orbit: no Kepler solve (model.py:239)
        PlanetConst pc;
        pc.n = a.x; pc.tp = a.y; pc.e = bq.x;
        SolverPlan plan;
        plan.tol = e4.x;
        plan.n32 = __double2loint(e4.y);
        plan.n64 = __double2hiint(e4.y);
        double tt[W];
#pragma unroll
        for (int j = 0; j < W; ++j) tt[j] = T.t[s.base + j * 32 + lane];
        StarterOut<W> o;
        planet_starter<W>(pc, plan, tt, o);
        const int slot = s.slot_i % kSlots;
        mbar_wait(empty + slot, ((s.slot_i / kSlots) & 1) ^ 1);
        ++s.slot_i;
        unsigned char* ring = link_ring(l) + slot * W * 32 * 12;
        double* sm = reinterpret_cast<double*>(ring);
        int* se = reinterpret_cast<int*>(ring + W * 32 * 8);
#pragma unroll
        for (int j = 0; j < W; ++j) {
          sm[j * 32 + lane] = o.m[j];
          se[j * 32 + lane] = (__float_as_int(o.Ef[j]) & 0x7fffffff) | o.sign[j];
        }
        mbar_arrive(full + slot);
        return;
      }
    };
    for (bool any = true; any;) {
      any = false;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (!st[c].done) {
          turn(st[c], c * kProducers + warp);
          any = true;
        }
      }
    }
  } else {
    if constexpr (RC > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RC));
    const int l = warp - kProducers;
    uint64_t* bars = link_bars(l);
    uint64_t *full = bars, *empty = bars + kSlots, *rec_full = bars + 2 * kSlots, *rec_free = rec_full + kRecSets;
    double* recs = link_recs(l);
    long long* batch_of = link_batch(l);
    const unsigned char* ring = link_ring(l);
    uint32_t slot_i = 0;
    int64_t b = (int64_t)blockIdx.x * kLinks + l;
    // records of batch i live in set i % 2; the set of batch i + 1 is written before batch i is processed, so the
    // producer can run ahead.  A set is reused once the producer has left the batch that used it (rec_free).
    uint32_t pub_i = 0;
    auto publish_next = [&]() {
      const int set = pub_i % kRecSets;
      mbar_wait(rec_free + set, ((pub_i / kRecSets) & 1) ^ 1);
      ++pub_i;
      const bool more = b < n_batches;
      const int64_t s0n = more ? b * nb : -1;
      if (more) {
        if (next_batch) {                                    // dynamic schedule, as logprob_kernel
          unsigned long long t = 0;
          if (lane == 0) t = atomicAdd(next_batch, 1ull);
          b = n_streams + (int64_t)__shfl_sync(0xffffffffu, t, 0);
        } else {
          b += n_streams;
        }
        sample_prologue(P, T, theta, s0n, S, recs + set * kGws * rec, rec, lane, true, nb);
      }
      if (lane == 0) batch_of[set] = s0n;
      __syncwarp();
      mbar_arrive(rec_full + set);                           // 32 arrivals: each lane's record writes are ordered
      return s0n;
    };
    int64_t s0 = publish_next();
    for (uint32_t set_i = 0; s0 >= 0; ++set_i) {
      const int set = set_i % kRecSets;
      const double* scratch = recs + set * kGws * rec;
      const int64_t s0_next = publish_next();
      for (int g = 0; g < nb; ++g) {
        const int64_t s = s0 + g;
        if (s >= S) break;
        const double* sr = scratch + g * rec;
        const int flags = __double2loint(sr[1]);
        const double lp = sr[0];
        double ll;
        if (flags & F_PLANET) {
          ll = -INFINITY;                                    // fit.py:3625-3627
        } else if ((flags & (F_JIT | F_PRIOR)) && ll_out == nullptr) {
          ll = 0.0;                                          // result is -inf regardless: skip the work
        } else {
          ChiAcc acc;
          const double2* planets = reinterpret_cast<const double2*>(sr + kHdr + 2 * P.n_inst);
          const double c0 = sr[5], gd = sr[2], gdd = sr[3];
          for (int base = 0; base < P.n_pad; base += 32 * W) {
            double tt[W], rv[W];
#pragma unroll
            for (int j = 0; j < W; ++j) {
              tt[j] = T.t[base + j * 32 + lane];
              rv[j] = c0;
            }
            for (int k = 0; k < npl; ++k) {
              const double2* pr = planets + k * (kPlanetRec / 2);
              const double2 a = pr[0], bq = pr[1], c = pr[2], d = pr[3], e4 = pr[4];
              PlanetConst pc;
              pc.n = a.x; pc.tp = a.y; pc.e = bq.x; pc.A = bq.y; pc.B = c.x; pc.C = c.y; pc.w = d.x; pc.K = d.y;
              SolverPlan plan;
              plan.tol = e4.x;
              plan.n32 = __double2loint(e4.y);
              plan.n64 = __double2hiint(e4.y);
              if (pc.e == 0) {                               // circular branch, no producer work
                planet_rv_add<W>(pc, plan, tt, rv);
                continue;
              }
              const int slot = slot_i % kSlots;
              mbar_wait(full + slot, (slot_i / kSlots) & 1);
              const double* sm = reinterpret_cast<const double*>(ring + slot * W * 32 * 12);
              const int* se = reinterpret_cast<const int*>(ring + slot * W * 32 * 12 + W * 32 * 8);
              StarterOut<W> o;
#pragma unroll
              for (int j = 0; j < W; ++j) {
                o.m[j] = sm[j * 32 + lane];
                const int wd = se[j * 32 + lane];
                o.Ef[j] = __int_as_float(wd & 0x7fffffff);
                o.sign[j] = wd & (int)0x80000000;
              }
              mbar_arrive(empty + slot);                     // values are in registers: release the slot
              ++slot_i;
              planet_rv_from_starter<W>(pc, plan, tt, o, rv);
            }
#pragma unroll
            for (int j = 0; j < W; ++j) {                    // model.py:483-509
              const double dt = tt[j] - P.t0;
              rv[j] = fma(gdd, dt * dt, fma(gd, dt, rv[j]));
            }
            chi_epilogue<W>(P, T, sr, base, lane, rv, acc);
          }
          ll = -0.5 * chi_finish(acc);
        }
        if (lane == 0) {
          double r;
          if (flags & (F_JIT | F_PRIOR | F_HYPER)) {
            r = -INFINITY;                                   // fit.py:3468, 3480-3482
          } else {
            r = ll + lp;                                     // fit.py:3492-3495
            r += P.jacobian;
            r += P.renorm;
          }
          if (out) out[s] = r;
          if (ll_out) ll_out[s] = ll;
          if (lp_out) lp_out[s] = lp;
        }
      }
      s0 = s0_next;
    }
  }
}

}  // namespace rvlp
