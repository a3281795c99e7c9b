// rvlp_capi.cu — C ABI (include/ravest_b200.h) over the kernels.  No torch, no C++ types
// across the boundary; errors are integer codes + a thread-local message.
#include <cuda_runtime.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <climits>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <condition_variable>
#include <memory>
#include <vector>

#include "rvlp_bands.cuh"
#include "rvlp_bands_fast.cuh"
#include "rvlp_gp.cuh"
#include "rvlp_gp_batch.cuh"
#include "rvlp_gp_pipe.cuh"
#include "rvlp_gp_smem.cuh"
#include "rvlp_kernels.cuh"

using namespace rvlp;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fail(RVLP_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                  __LINE__);                                                                   \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int grid_for(int device, const void* kernel, int smem_bytes, int64_t want_blocks, int* grid, int threads = kThreads) {
  int sms = 0, per_sm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem_bytes));
  if (per_sm < 1) return fail(RVLP_EUNSUPPORTED, "kernel does not fit on an SM (smem %d B)", smem_bytes);
  int64_t g = (int64_t)sms * per_sm;      // one full wave of resident CTAs, persistent loop inside
  if (want_blocks < g) g = want_blocks;
  if (g < 1) g = 1;
  *grid = (int)g;
  return RVLP_OK;
}

// Per-device one-time upload of the sin/cos grid table (host libm values).
int ensure_tables(int device) {
  static std::atomic<uint64_t> done{0};
  if (device < 64 && (done.load() >> device) & 1) return RVLP_OK;
  std::vector<double> tab(2 * (size_t)kTabN);
  for (int j = 0; j < kTabN; ++j) {
    tab[2 * j] = sin(j / 512.0);
    tab[2 * j + 1] = cos(j / 512.0);
  }
  CUDA_TRY(cudaMemcpyToSymbol(kSinCosTabDev, tab.data(), sizeof(double) * tab.size()));
  if (device < 64) done.fetch_or(1ull << device);
  return RVLP_OK;
}

// Per-device one-time setup of the default stream-ordered memory pool: keep freed blocks (no trim at synchronisation
// points), so that per-call scratch (rvlp_gp_predict_batch) costs no driver allocation in steady state.
int ensure_pool(int device) {
  static std::atomic<uint64_t> done{0};
  if (device < 64 && (done.load() >> device) & 1) return RVLP_OK;
  cudaMemPool_t pool;
  CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long keep = ~0ull;
  CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  if (device < 64) done.fetch_or(1ull << device);
  return RVLP_OK;
}

int simple_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

// K1 shapes (rvlp_kernels.cuh: logprob_kernel<W, MB>); variant 0 is the default until rvlp_ctx_autotune ran
typedef void (*k1_fn)(DevProblem, const double*, int64_t, double*, double*, double*, int, unsigned long long*, PeerOut);
// 0: logprob_kernel<4, 2>   1: logprob_kernel<2, 3>
constexpr int kK1Variants = 2;
struct K1Shape {
  k1_fn fn;
  int threads;
  int workers;      // sample streams per CTA (warps)
  int smem;         // dynamic shared memory
  int max_nb;       // samples per prologue batch
};
static K1Shape k1_shape(int v, const DevProblem& P, const SmemLayout& L, bool peers = false) {
  const bool ge = P.epochs_global != 0;
  K1Shape s{};
  s.threads = kThreads;
  s.workers = kWarps;
  s.smem = L.total;
  s.max_nb = P.batch_cap > kG ? P.batch_cap : kG;
  if (peers) {   // the instantiation with the fused all-gather epilogue (rvlp_logprob_batch_peers)
    if (v == 1) s.fn = ge ? logprob_kernel<2, 3, true, true> : logprob_kernel<2, 3, false, true>;
    else s.fn = ge ? logprob_kernel<kW, RVLP_MIN_BLOCKS, true, true> : logprob_kernel<kW, RVLP_MIN_BLOCKS, false, true>;
    return s;
  }
  if (v == 1) s.fn = ge ? logprob_kernel<2, 3, true> : logprob_kernel<2, 3, false>;
  else s.fn = ge ? logprob_kernel<kW, RVLP_MIN_BLOCKS, true> : logprob_kernel<kW, RVLP_MIN_BLOCKS, false>;
  return s;
}

// Batch-ticket counters for the dynamic schedule of logprob_kernel / rv_matrix_kernel.  A launch takes the next slot of
// the ring; the slot carries an event recorded right after the kernel that used it, and a later launch that lands on
// the same slot first makes ITS stream wait for that event.  Any number of launches may therefore be in flight on any
// number of streams: a counter is never reset or shared while a kernel still reads it.
constexpr int kTicketRing = 256;

struct rvlp_ctx {
  int device = 0;
  DevProblem P{};
  void* d_src_col = nullptr;
  void* d_src_const = nullptr;
  void* d_priors = nullptr;
  void* d_epochs = nullptr;
  int smem_main = 0, smem_gp_pipe = 0, smem_gp_pipe_pred = 0, gp_tile = 0;
  int max_smem = 0;
  int k1 = 0;          // K1 variant in use
  int k1_tuned = 0;    // rvlp_ctx_autotune has run
  // host-buffer path
  double* h_theta = nullptr;
  double* h_out = nullptr;
  double* d_theta = nullptr;
  double* d_out = nullptr;
  int64_t cap_samples = 0;
  unsigned long long* d_tickets = nullptr;   // ring of batch-ticket counters for logprob_kernel's dynamic schedule
  cudaEvent_t ticket_ev[kTicketRing] = {};   // completion of the last kernel that used the slot
  unsigned ticket_next = 0;
  std::mutex ticket_mu;                      // slot hand-out + (wait, memset, launch, record) is one critical section
  cudaEvent_t ev_done[8] = {};               // host-buffer path: one per chunk, created once
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;   // autotune timing
  std::mutex host_mu;                        // the host-buffer path owns the staging buffers: one call at a time
  cudaStream_t stream = nullptr;    // host-buffer path: chunks alternate between two streams so that
  cudaStream_t stream2 = nullptr;   // the H2D copy of chunk i+1 overlaps the kernel of chunk i
};

// Level-synchronous batched Cholesky (rvlp_gp_batch.cuh): one kernel per block column over ALL samples of a chunk.
// The per-sample factors live in a workspace from the device's stream-ordered pool (chunks of at most ~6 GB).
template <bool PRED>
static int launch_gp_batch(rvlp_ctx* c, const double* theta_dev, int64_t S, double* out_dev, double* beta_dev, cudaStream_t st) {
  DevProblem P = c->P;
  P.epochs_global = 1;                                      // these kernels read the epoch arrays through L1
  const int N = P.n_epochs;
  const GpbDims d = gpb_dims(N);
  const size_t per = (gpb_bytes_per_sample(N, P.n_inst) + 255) & ~(size_t)255;
  size_t cap = (size_t)6 << 30;
  if (const char* e = getenv("RVLP_GP_BATCH_MB")) {         // tests: force chunking
    if (atoi(e) > 0) cap = (size_t)atoi(e) << 20;
  }
  int64_t chunk = (int64_t)(cap / per);
  if (chunk < 1) chunk = 1;
  if (chunk > S) chunk = S;
  int rc = ensure_pool(c->device);
  if (rc) return rc;
  static std::atomic<uint64_t> attr_done{0};                // per device
  if (c->device >= 64 || !((attr_done.load() >> c->device) & 1)) {
    CUDA_TRY(cudaFuncSetAttribute(gpb_prologue_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(gpb_prologue_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    CUDA_TRY(cudaFuncSetAttribute(gpb_backsub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    if (c->device < 64) attr_done.fetch_or(1ull << c->device);
  }
  const int smem_pro = smem_layout(P).total;
  if (smem_pro > c->max_smem || N * 8 > c->max_smem)
    return fail(RVLP_EUNSUPPORTED, "GP problem needs %d B of shared memory per CTA (> %d)", smem_pro > N * 8 ? smem_pro : N * 8, c->max_smem);
  unsigned char* slab = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&slab, (size_t)chunk * per + 8192, st));
  struct SlabFree {
    void* p; cudaStream_t st;
    ~SlabFree() { cudaFreeAsync(p, st); }
  } slab_free{slab, st};
  GpbWork w;
  {
    unsigned char* o = slab;
    auto take = [&](size_t bytes) { unsigned char* r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    w.L = (double*)take((size_t)chunk * d.nblk * 256 * 8);
    w.invd = (double*)take((size_t)chunk * d.np * 8);
    w.resid = (double*)take((size_t)chunk * d.np * 8);
    w.cph = (double*)take((size_t)chunk * d.np * 8);
    w.sph = (double*)take((size_t)chunk * d.np * 8);
    w.hyp = (double*)take((size_t)chunk * 4 * 8);
    w.jit2 = (double*)take((size_t)chunk * P.n_inst * 8);
    w.lp = (double*)take((size_t)chunk * 8);
    w.lhp = (double*)take((size_t)chunk * 8);
    w.chi2 = (double*)take((size_t)chunk * 8);
    w.logdet = (double*)take((size_t)chunk * 8);
    w.status = (int*)take((size_t)chunk * 4);
    if ((size_t)(o - slab) > (size_t)chunk * per + 8192) return fail(RVLP_ECUDA, "internal: GP workspace carve overflow");
  }
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  for (int64_t s0 = 0; s0 < S; s0 += chunk) {
    const int64_t n = S - s0 < chunk ? S - s0 : chunk;
    const double* th = theta_dev + s0 * P.ndim;
    int grid = 0;
    if ((rc = grid_for(c->device, (const void*)gpb_prologue_kernel<PRED>, smem_pro, (n + kWarps - 1) / kWarps, &grid))) return rc;
    gpb_prologue_kernel<PRED><<<grid, kThreads, smem_pro, st>>>(P, th, n, w);
    if ((rc = grid_for(c->device, (const void*)gpb_diag0_kernel, 0, (n + kGbWarps - 1) / kGbWarps, &grid, kGbThreads))) return rc;
    gpb_diag0_kernel<<<grid, kGbThreads, 0, st>>>(P, n, w);
    g_launches += 2;
    const int tpw = N >= 300 ? 2 : 1;                            // row blocks per warp of the step kernel (see there)
    for (int J = 0; J + 1 < d.nbr; ++J) {
      const int64_t tasks = n * (int64_t)((d.nbr - 1 - J + tpw - 1) / tpw);   // one warp per (sample, group of row blocks)
      const void* kern = tpw == 2 ? (const void*)gpb_step_kernel<2> : (const void*)gpb_step_kernel<1>;
      if ((rc = grid_for(c->device, kern, 0, (tasks + kGbWarps - 1) / kGbWarps, &grid, kGbThreads))) return rc;
      if (tpw == 2) gpb_step_kernel<2><<<grid, kGbThreads, 0, st>>>(P, n, w, J);
      else gpb_step_kernel<1><<<grid, kGbThreads, 0, st>>>(P, n, w, J);
      ++g_launches;
    }
    int64_t gf = (n + 127) / 128;
    if (gf > sms * 8) gf = sms * 8;
    gpb_finish_kernel<PRED><<<(int)gf, 128, 0, st>>>(P, n, w, out_dev ? out_dev + s0 : nullptr, beta_dev ? beta_dev + s0 * N : nullptr);
    ++g_launches;
    if (PRED) {
      if ((rc = grid_for(c->device, (const void*)gpb_backsub_kernel, N * 8, n, &grid, kGbThreads))) return rc;
      gpb_backsub_kernel<<<grid, kGbThreads, N * 8, st>>>(P, n, w, beta_dev + s0 * N);
      ++g_launches;
    }
    CUDA_TRY(cudaGetLastError());
  }
  return RVLP_OK;
}

// Shared-memory tensor-core Cholesky (rvlp_gp_smem.cuh): one 4-warp CTA per sample in flight; the prologue kernel of
// the batched path supplies residual / phases / hyperparameters / flags through a small per-sample workspace.
template <bool PRED>
static int launch_gp_smem(rvlp_ctx* c, const double* theta_dev, int64_t S, double* out_dev, double* beta_dev, cudaStream_t st) {
  DevProblem P = c->P;
  P.epochs_global = 1;
  const int N = P.n_epochs;
  const GpbDims d = gpb_dims(N);
    const int smem_f = gps_smem_bytes(N, PRED);
  const int smem_pro = smem_layout(P).total;
  if (smem_f > c->max_smem || smem_pro > c->max_smem)
    return fail(RVLP_EUNSUPPORTED, "shared-memory GP kernel: %d epochs need %d B per CTA (> %d)", N, smem_f, c->max_smem);
  int rc = ensure_pool(c->device);
  if (rc) return rc;
  void (*kern)(DevProblem, int64_t, GpbWork, unsigned long long*, double*, double*) = gps_factor_kernel<PRED>;
  CUDA_TRY(cudaFuncSetAttribute(gpb_prologue_kernel<PRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
  const size_t per = ((size_t)(3 * d.np + 4 + P.n_inst + 4) * 8 + 8 + 255) & ~(size_t)255;
  int64_t chunk = (int64_t)(((size_t)2 << 30) / per);
  if (const char* e = getenv("RVLP_GP_BATCH_MB")) {         // tests: force chunking
    if (atoi(e) > 0) chunk = (int64_t)(((size_t)atoi(e) << 20) / per);
  }
  if (chunk < 1) chunk = 1;
  if (chunk > S) chunk = S;
  unsigned char* slab = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&slab, (size_t)chunk * per + 8192, st));
  struct SlabFree {
    void* p; cudaStream_t st;
    ~SlabFree() { cudaFreeAsync(p, st); }
  } slab_free{slab, st};
  GpbWork w{};
  unsigned long long* ticket = nullptr;
  {
    unsigned char* o = slab;
    auto take = [&](size_t bytes) { unsigned char* r = o; o += (bytes + 255) & ~(size_t)255; return r; };
    ticket = (unsigned long long*)take(8);
    w.resid = (double*)take((size_t)chunk * d.np * 8);
    w.cph = (double*)take((size_t)chunk * d.np * 8);
    w.sph = (double*)take((size_t)chunk * d.np * 8);
    w.hyp = (double*)take((size_t)chunk * 4 * 8);
    w.jit2 = (double*)take((size_t)chunk * P.n_inst * 8);
    w.lp = (double*)take((size_t)chunk * 8);
    w.lhp = (double*)take((size_t)chunk * 8);
    w.chi2 = (double*)take((size_t)chunk * 8);
    w.logdet = (double*)take((size_t)chunk * 8);
    w.status = (int*)take((size_t)chunk * 4);
    if ((size_t)(o - slab) > (size_t)chunk * per + 8192) return fail(RVLP_ECUDA, "internal: GP workspace carve overflow");
  }
  for (int64_t s0 = 0; s0 < S; s0 += chunk) {
    const int64_t n = S - s0 < chunk ? S - s0 : chunk;
    int grid = 0;
    if ((rc = grid_for(c->device, (const void*)gpb_prologue_kernel<PRED>, smem_pro, (n + kWarps - 1) / kWarps, &grid))) return rc;
    CUDA_TRY(cudaMemsetAsync(ticket, 0, 8, st));
    gpb_prologue_kernel<PRED><<<grid, kThreads, smem_pro, st>>>(P, theta_dev + s0 * P.ndim, n, w);
    if ((rc = grid_for(c->device, (const void*)kern, smem_f, n, &grid, kGsThreads))) return rc;
    if (const char* e = getenv("RVLP_GP_GRID")) {           // tests / experiments: cap the grid
      if (atoi(e) > 0 && atoi(e) < grid) grid = atoi(e);
    }
    kern<<<grid, kGsThreads, smem_f, st>>>(P, n, w, ticket, out_dev ? out_dev + s0 : nullptr, beta_dev ? beta_dev + s0 * N : nullptr);
    g_launches += 2;
    CUDA_TRY(cudaGetLastError());
  }
  return RVLP_OK;
}

// Which GP implementation serves a call (profiles/r02aa_gp_sweep3.log: the three log-probability paths over 12..200
// epochs and 256..16384 samples; profiles/r02q_gp_crossover.log for the conditioning path).
//   log-probability, 40..208 epochs: the shared-memory tensor-core kernel (rvlp_gp_smem.cuh; 1.3-2.6x the others -
//     four samples per SM to 128 epochs, three to 160, two to 208; profiles/r02ag_gp_sweep4.log), except batches under
//     512 samples below 80 epochs (the pipelined kernel's single launch wins by ~10 %);
//   under 40 epochs: the pipelined kernel; the batched path for >= 4096 samples at <= 32 epochs;
//   beyond (one shared-memory CTA per SM from 209 epochs, no pipelined shape from 220): the batched path from 512
//     samples, the pipelined kernel below while it has a shape.
//   Conditioning (K7; profiles/r02ak_gp_pred_sweep.log): the shared-memory kernel's <PRED> flavour (whole factor kept,
//     three samples per SM) at 40..144 epochs (1.9x the pipelined kernel at 120) and to 208 under 2048 samples; the
//     pipelined kernel below 40 epochs and for small batches under 80; batched beyond.
// RVLP_GP_KERNEL = pipe | smem | batch forces one (tests, experiments).
enum { GP_PIPE = 0, GP_SMEM = 1, GP_BATCH = 2 };
static int gp_choice(const rvlp_ctx* c, int64_t S, bool pred) {
  const bool smem_ok = gps_smem_bytes(c->P.n_epochs, pred) <= c->max_smem;
  if (const char* e = getenv("RVLP_GP_KERNEL")) {
    if (!strcmp(e, "smem") && smem_ok) return GP_SMEM;
    if (!strcmp(e, "batch")) return GP_BATCH;
    if (!strcmp(e, "pipe") && c->gp_tile != 0) return GP_PIPE;
  }
  const int N = c->P.n_epochs;
  if (smem_ok && N >= 40 && N <= (pred ? 144 : 208)) return (S < 512 && N < 80 && c->gp_tile != 0) ? GP_PIPE : GP_SMEM;
  if (smem_ok && pred && N <= 208 && S < 2048) return GP_SMEM;
  if (c->gp_tile == 0) return GP_BATCH;
  if (N >= 140 && S >= 512) return GP_BATCH;
  if (!pred && S >= 4096 && N <= 32) return GP_BATCH;
  return GP_PIPE;
}

extern "C" {

int rvlp_abi_version(void) { return RVLP_ABI_VERSION; }
const char* rvlp_last_error(void) { return g_err.c_str(); }
int64_t rvlp_launch_count(void) { return g_launches.load(); }

int rvlp_ctx_create(const rvlp_desc* d, const double* time, const double* vel, const double* velerr,
                    const int32_t* inst_idx, int64_t n_epochs, int device, rvlp_ctx** out) {
  if (!d || !out) return fail(RVLP_EINVAL, "null descriptor / out pointer");
  if (d->abi_version != RVLP_ABI_VERSION)
    return fail(RVLP_EINVAL, "descriptor ABI version %d != library %d", d->abi_version, RVLP_ABI_VERSION);
  if (d->n_planets < 0 || d->n_inst < 1 || d->ndim < 0 || d->n_priors < 0)
    return fail(RVLP_EINVAL, "bad descriptor sizes");
  if (d->parameterisation < 0 || d->parameterisation > 3) return fail(RVLP_EINVAL, "bad parameterisation id");
  if (d->n_hyper != 0 && d->n_hyper != 4) return fail(RVLP_EINVAL, "n_hyper must be 0 or 4");
  if (n_epochs < 1 || n_epochs > (1 << 24)) return fail(RVLP_EINVAL, "bad epoch count %lld", (long long)n_epochs);
  const int n_model = 5 * d->n_planets + 2 + 2 * d->n_inst;
  const int n_src = n_model + d->n_hyper;
  for (int i = 0; i < n_src; ++i)
    if (d->src_col[i] >= d->ndim) return fail(RVLP_EINVAL, "src_col[%d] = %d out of range", i, d->src_col[i]);
  for (int i = 0; i < d->n_priors; ++i) {
    const rvlp_prior& p = d->priors[i];
    if (p.kind < 0 || p.kind > RVLP_PRIOR_BETA) return fail(RVLP_EINVAL, "prior %d: bad kind", i);
    if (p.target < 0 || p.target > RVLP_TARGET_TP) return fail(RVLP_EINVAL, "prior %d: bad target", i);
    if (p.target == RVLP_TARGET_COLUMN ? (p.index < 0 || p.index >= d->ndim)
                                       : (p.index < 0 || p.index >= d->n_planets))
      return fail(RVLP_EINVAL, "prior %d: index out of range", i);
  }
  for (int64_t i = 0; i < n_epochs; ++i)
    if (inst_idx[i] < 0 || inst_idx[i] >= d->n_inst) return fail(RVLP_EINVAL, "inst_idx[%lld] out of range", (long long)i);

  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(RVLP_EINVAL, "device %d not present (%d visible)", device, ndev);
  DeviceGuard guard(device);
  if (int rc = ensure_tables(device)) return rc;

  rvlp_ctx* c = new rvlp_ctx();
  c->device = device;
  DevProblem& P = c->P;
  P.n_planets = d->n_planets; P.par = d->parameterisation; P.n_inst = d->n_inst; P.ndim = d->ndim;
  P.n_priors = d->n_priors; P.n_hyper = d->n_hyper; P.n_model = n_model; P.n_epochs = (int)n_epochs;
  P.n_pad = (int)((n_epochs + kPadTo - 1) / kPadTo * kPadTo);   // whole lane groups; x28 B is a multiple of 16 for the bulk copy
  P.t0 = d->t0; P.jacobian = d->jacobian; P.renorm = d->renorm;
  // samples per prologue batch: enough to fill the 32 lanes of the (sample, planet) phase, 4 .. 16
  P.batch_cap = P.n_planets > 0 ? 32 / P.n_planets : 16;
  P.batch_cap = P.batch_cap < kG ? kG : (P.batch_cap > 16 ? 16 : P.batch_cap);
  if (P.n_hyper) P.batch_cap = kG;   // GP contexts: K3 / K7 keep their own records; shared memory decides their occupancy
  if (const char* e = getenv("RVLP_BATCH_CAP")) {            // experiments (tools/variant_time.py)
    const int v = atoi(e);
    if (v >= kG && v <= 32) P.batch_cap = v;
  }
  P.n_pad = (int)((n_epochs + kPadTo - 1) / kPadTo * kPadTo);
  while (P.batch_cap > kG && smem_layout(P).total > 74 * 1024) --P.batch_cap;   // keep three CTAs per SM possible
  {
    int64_t units = (int64_t)P.n_pad * (P.n_planets > 0 ? P.n_planets : 1), gu = kGssUnits;
    if (const char* e = getenv("RVLP_GSS_UNITS")) gu = atoll(e) > 0 ? atoll(e) : gu;   // experiments (tools/k1_shard.py)
    P.gss_min = (int)((gu + units - 1) / units);
  }

  // packed, padded epoch block: [t | vel | velerr^2] doubles + int32 instrument ids
  std::vector<unsigned char> blk((size_t)P.n_pad * 28);
  double* ht = reinterpret_cast<double*>(blk.data());
  double* hv = ht + P.n_pad;
  double* he = hv + P.n_pad;
  int32_t* hi = reinterpret_cast<int32_t*>(he + P.n_pad);
  for (int i = 0; i < P.n_pad; ++i) {
    const bool in = i < n_epochs;
    ht[i] = in ? time[i] : time[n_epochs - 1];
    hv[i] = in ? vel[i] : 0.0;
    he[i] = in ? velerr[i] * velerr[i] : 1.0;                // fit.py:3598
    hi[i] = in ? inst_idx[i] : 0;
  }
#define CTX_TRY(expr)                                                                            \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      int rc = fail(RVLP_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));                 \
      rvlp_ctx_destroy(c);                                                                       \
      return rc;                                                                                 \
    }                                                                                            \
  } while (0)
  CTX_TRY(cudaMalloc(&c->d_epochs, blk.size()));
  CTX_TRY(cudaMemcpy(c->d_epochs, blk.data(), blk.size(), cudaMemcpyHostToDevice));
  CTX_TRY(cudaMalloc(&c->d_src_col, sizeof(int32_t) * (size_t)(n_src > 0 ? n_src : 1)));
  CTX_TRY(cudaMemcpy(c->d_src_col, d->src_col, sizeof(int32_t) * (size_t)n_src, cudaMemcpyHostToDevice));
  CTX_TRY(cudaMalloc(&c->d_src_const, sizeof(double) * (size_t)(n_src > 0 ? n_src : 1)));
  CTX_TRY(cudaMemcpy(c->d_src_const, d->src_const, sizeof(double) * (size_t)n_src, cudaMemcpyHostToDevice));
  CTX_TRY(cudaMalloc(&c->d_priors, sizeof(rvlp_prior) * (size_t)(d->n_priors > 0 ? d->n_priors : 1)));
  if (d->n_priors)
    CTX_TRY(cudaMemcpy(c->d_priors, d->priors, sizeof(rvlp_prior) * (size_t)d->n_priors, cudaMemcpyHostToDevice));
  P.epochs = reinterpret_cast<const double*>(c->d_epochs);
  P.src_col = reinterpret_cast<const int32_t*>(c->d_src_col);
  P.src_const = reinterpret_cast<const double*>(c->d_src_const);
  P.priors = reinterpret_cast<const rvlp_prior*>(c->d_priors);

  CTX_TRY(cudaDeviceGetAttribute(&c->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  if (P.n_hyper) {
    // GP contexts: the pipelined register-tile kernels up to 219 epochs; the level-synchronous batched path
    // (rvlp_gp_batch.cuh, factors in a global workspace) for any epoch count up to 16384
    c->gp_tile = gp_tile_for(P.n_epochs);
    if (c->gp_tile && gp_pipe_smem(P, smem_layout(P), c->gp_tile, true).total > c->max_smem) c->gp_tile = 0;
    if (P.n_epochs > 16384) {
      int rc = fail(RVLP_EUNSUPPORTED, "GP problems are limited to 16384 epochs (%d given): the dense factor needs "
                    "%.1f GB per sample in flight", P.n_epochs, gpb_bytes_per_sample(P.n_epochs, P.n_inst) * 1e-9);
      rvlp_ctx_destroy(c);
      return rc;
    }
    if (c->gp_tile == 0) P.epochs_global = 1;   // no kernel of this context stages the epochs in shared memory
  }
  SmemLayout L = smem_layout(P);
  if (L.total > c->max_smem && !P.n_hyper) {   // too many epochs to stage per CTA: leave them in global memory
    P.epochs_global = 1;
    L = smem_layout(P);
  }
  c->smem_main = L.total;
  if (c->smem_main > c->max_smem) {
    int rc = fail(RVLP_EUNSUPPORTED, "problem needs %d B of shared memory per CTA (> %d): too many epochs",
                  c->smem_main, c->max_smem);
    rvlp_ctx_destroy(c);
    return rc;
  }
  for (int v = 0; v < kK1Variants; ++v) {
    CTX_TRY(cudaFuncSetAttribute(k1_shape(v, P, L).fn, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    CTX_TRY(cudaFuncSetAttribute(k1_shape(v, P, L, true).fn, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
  }
  CTX_TRY(cudaFuncSetAttribute(rv_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
  CTX_TRY(cudaFuncSetAttribute(walker_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
  if (P.n_hyper) {
    c->smem_gp_pipe = c->gp_tile ? gp_pipe_smem(P, L, c->gp_tile).total : 0;
    c->smem_gp_pipe_pred = c->gp_tile ? gp_pipe_smem(P, L, c->gp_tile, true).total : 0;
#define RVLP_GP_ATTR(TT)                                                                                                 \
    CTX_TRY(cudaFuncSetAttribute(gp_logprob_pipe_kernel<TT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem)); \
    CTX_TRY(cudaFuncSetAttribute(gp_logprob_pipe_kernel<TT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
    RVLP_GP_ATTR(2) RVLP_GP_ATTR(4) RVLP_GP_ATTR(6) RVLP_GP_ATTR(8) RVLP_GP_ATTR(10)
#undef RVLP_GP_ATTR
    CTX_TRY(cudaFuncSetAttribute(gp_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem));
  }
  CTX_TRY(cudaMalloc((void**)&c->d_tickets, sizeof(unsigned long long) * kTicketRing));
  for (int i = 0; i < kTicketRing; ++i) CTX_TRY(cudaEventCreateWithFlags(&c->ticket_ev[i], cudaEventDisableTiming));
  for (int i = 0; i < 8; ++i) CTX_TRY(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
  CTX_TRY(cudaEventCreate(&c->ev_a));
  CTX_TRY(cudaEventCreate(&c->ev_b));
  CTX_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CTX_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
#undef CTX_TRY
  *out = c;
  return RVLP_OK;
}

void rvlp_ctx_destroy(rvlp_ctx* c) {
  if (!c) return;
  DeviceGuard guard(c->device);
  cudaFree(c->d_epochs);
  cudaFree(c->d_src_col);
  cudaFree(c->d_src_const);
  cudaFree(c->d_priors);
  cudaFree(c->d_theta);
  cudaFree(c->d_out);
  cudaFree(c->d_tickets);
  for (int i = 0; i < kTicketRing; ++i)
    if (c->ticket_ev[i]) cudaEventDestroy(c->ticket_ev[i]);
  for (int i = 0; i < 8; ++i)
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->h_theta) cudaFreeHost(c->h_theta);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  delete c;
}

static int launch_logprob(rvlp_ctx* c, const double* theta, int64_t S, double* out, double* ll, double* lp,
                          cudaStream_t st, const PeerOut* peer_out = nullptr) {
  if (S == 0) return RVLP_OK;
  PeerOut peers{};
  if (peer_out) peers = *peer_out;
  int grid = 0;
  SmemLayout L = smem_layout(c->P);
  K1Shape shape = k1_shape(c->k1, c->P, L, peers.n > 0);
  if (shape.smem > c->max_smem) shape = k1_shape(0, c->P, L, peers.n > 0);   // rings do not fit next to this many epochs
  const k1_fn kern = shape.fn;
  const int smem = shape.smem;
  int rc = grid_for(c->device, (const void*)kern, smem, INT_MAX, &grid, shape.threads);   // full wave
  if (rc) return rc;
  // samples per prologue batch: kG when every resident worker still gets a batch, else 1 (latency of small S)
  // (beyond kG only while every warp still gets four or more batches: the dynamic schedule's tail is one batch)
  int64_t per_warp = S / ((int64_t)grid * shape.workers);
  int nb = (int)(per_warp >= kG ? kG : (per_warp < 1 ? 1 : per_warp));
  if (per_warp / 4 > kG) nb = (int)(per_warp / 4 < shape.max_nb ? per_warp / 4 : shape.max_nb);
  const int64_t want = ((S + nb - 1) / nb + shape.workers - 1) / shape.workers;
  if (want < grid) grid = (int)want;
  if (per_warp >= 2) {                                     // several batches per warp: dynamic schedule
    std::lock_guard<std::mutex> lock(c->ticket_mu);
    const unsigned slot = c->ticket_next++ % kTicketRing;
    unsigned long long* tickets = c->d_tickets + slot;
    CUDA_TRY(cudaStreamWaitEvent(st, c->ticket_ev[slot], 0));   // the slot's previous user (any stream) has finished
    CUDA_TRY(cudaMemsetAsync(tickets, 0, sizeof(unsigned long long), st));
    kern<<<grid, shape.threads, smem, st>>>(c->P, theta, S, out, ll, lp, nb, tickets, peers);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ticket_ev[slot], st));
    return RVLP_OK;
  }
  kern<<<grid, shape.threads, smem, st>>>(c->P, theta, S, out, ll, lp, nb, nullptr, peers);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

// ------------------------------------------------------------------ fused all-gather over peer memory
int rvlp_peer_alloc(int device, int64_t bytes, void** dev_ptr, void* handle64) {
  if (!dev_ptr || !handle64 || bytes <= 0) return fail(RVLP_EINVAL, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DeviceGuard guard(device);
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(RVLP_ECUDA, "rvlp_peer_alloc: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return RVLP_OK;
}

int rvlp_peer_open(int device, const void* handle64, void** dev_ptr) {
  if (!dev_ptr || !handle64) return fail(RVLP_EINVAL, "bad arguments");
  DeviceGuard guard(device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return RVLP_OK;
}

int rvlp_peer_close(int device, void* dev_ptr) {
  DeviceGuard guard(device);
  if (dev_ptr) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return RVLP_OK;
}

int rvlp_peer_free(int device, void* dev_ptr) {
  DeviceGuard guard(device);
  if (dev_ptr) CUDA_TRY(cudaFree(dev_ptr));
  return RVLP_OK;
}

int rvlp_logprob_batch_peers(rvlp_ctx* c, const double* theta_dev, int64_t S, double* const* outs, int32_t n_outs,
                             int64_t row_offset, void* stream) {
  if (!c || S < 0 || !outs || n_outs < 1 || n_outs > kMaxPeers || row_offset < 0 || (S > 0 && !theta_dev))
    return fail(RVLP_EINVAL, "bad arguments");
  if (c->P.n_hyper) return fail(RVLP_EINVAL, "GP context: the fused gather serves rvlp_logprob_batch only");
  DeviceGuard guard(c->device);
  PeerOut po{};
  po.n = n_outs;
  po.off = row_offset;
  for (int i = 0; i < n_outs; ++i) {
    if (!outs[i]) return fail(RVLP_EINVAL, "null output pointer %d", i);
    po.p[i] = outs[i];
  }
  return launch_logprob(c, theta_dev, S, nullptr, nullptr, nullptr, (cudaStream_t)stream, &po);
}

int rvlp_peer_barrier(int device, void* const* flag_blocks, int32_t n_ranks, int32_t my_rank, uint64_t epoch,
                      int64_t timeout_ms, void* stream) {
  if (!flag_blocks || n_ranks < 1 || n_ranks > kMaxPeers || my_rank < 0 || my_rank >= n_ranks || timeout_ms < 1)
    return fail(RVLP_EINVAL, "bad arguments");
  DeviceGuard guard(device);
  PeerOut fl{};
  fl.n = n_ranks;
  for (int i = 0; i < n_ranks; ++i) fl.p[i] = reinterpret_cast<double*>(flag_blocks[i]);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fl, my_rank, (unsigned long long)epoch,
                                                          (unsigned long long)timeout_ms * 1000000ull);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_logprob_batch(rvlp_ctx* c, const double* theta_dev, int64_t S, double* out_dev, void* stream) {
  if (!c || S < 0 || (S > 0 && (!theta_dev || !out_dev))) return fail(RVLP_EINVAL, "bad arguments");
  if (c->P.n_hyper) return fail(RVLP_EINVAL, "GP context: call rvlp_gp_logprob_batch");
  DeviceGuard guard(c->device);
  return launch_logprob(c, theta_dev, S, out_dev, nullptr, nullptr, (cudaStream_t)stream);
}

int rvlp_ctx_set_variant(rvlp_ctx* c, int32_t variant) {
  if (!c || variant < 0 || variant >= kK1Variants) return fail(RVLP_EINVAL, "variant must be 0..%d", kK1Variants - 1);
  c->k1 = variant;
  c->k1_tuned = 1;
  return RVLP_OK;
}

int rvlp_ctx_autotune(rvlp_ctx* c, const double* theta_dev, int64_t S, void* stream, int32_t* chosen) {
  if (!c || S < 0 || (S > 0 && !theta_dev)) return fail(RVLP_EINVAL, "bad arguments");
  if (chosen) *chosen = c->k1;
  if (c->P.n_hyper || S < 4096) return RVLP_OK;             // nothing to choose for GP contexts / tiny batches
  DeviceGuard guard(c->device);
  // Timed on the CALLER's stream: the launches are ordered after whatever produced theta_dev there (timing rows that
  // are still being written would measure garbage - e.g. all-invalid rows skip the likelihood - and latch the wrong
  // shape for the context's lifetime).
  cudaStream_t st = (cudaStream_t)stream;
  // the whole batch up to 2^18 rows: the prologue batch size and the number of batches per warp depend on the row count,
  // so a short prefix can favour the wrong shape (c2 at 1e5 rows flipped between runs with a 65 536-row prefix)
  const int64_t n = S < 262144 ? S : 262144;
  double* scratch = nullptr;
  CUDA_TRY(cudaMalloc((void**)&scratch, sizeof(double) * (size_t)n));
  const int prev = c->k1;
  float med[kK1Variants];
  int rc = RVLP_OK;
  for (int v = 0; v < kK1Variants && rc == RVLP_OK; ++v) {
    c->k1 = v;
    float t[5];
    for (int rep = -1; rep < 5 && rc == RVLP_OK; ++rep) {    // one warm-up, then the MEDIAN of five (a min-of-two once
      cudaEventRecord(c->ev_a, st);                           // mis-picked under a profiler and cost 9 %)
      rc = launch_logprob(c, theta_dev, n, scratch, nullptr, nullptr, st);
      cudaEventRecord(c->ev_b, st);
      if (rc == RVLP_OK && cudaEventSynchronize(c->ev_b) != cudaSuccess)
        rc = fail(RVLP_ECUDA, "autotune launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      float ms = 0;
      if (rc == RVLP_OK) cudaEventElapsedTime(&ms, c->ev_a, c->ev_b);
      if (rep >= 0) t[rep] = ms;
    }
    if (rc == RVLP_OK) {
      std::sort(t, t + 5);
      med[v] = t[2];
    }
  }
  cudaFree(scratch);
  int best_v = 0;
  if (rc == RVLP_OK)
    for (int v = 1; v < kK1Variants; ++v)
      if (med[v] < 0.98f * med[best_v]) best_v = v;          // an alternative shape must win by 2 %
  c->k1 = rc == RVLP_OK ? best_v : prev;
  c->k1_tuned = rc == RVLP_OK;
  if (chosen) *chosen = c->k1;
  return rc;
}

int rvlp_logprob_parts_batch(rvlp_ctx* c, const double* theta_dev, int64_t S, double* ll_dev, double* lp_dev,
                             void* stream) {
  if (!c || S < 0 || (S > 0 && !theta_dev)) return fail(RVLP_EINVAL, "bad arguments");
  if (c->P.n_hyper) return fail(RVLP_EINVAL, "GP context not supported here");
  if (!ll_dev && !lp_dev) return RVLP_OK;
  DeviceGuard guard(c->device);
  // ll_out non-null forces the likelihood to be evaluated even for rows the prior rejects
  return launch_logprob(c, theta_dev, S, nullptr, ll_dev, lp_dev, (cudaStream_t)stream);
}

int rvlp_info_criteria_batch(rvlp_ctx* c, const double* theta_dev, int64_t S, int32_t k_free, double* ll_dev,
                             double* chi2_dev, double* aicc_dev, double* bic_dev, void* stream) {
  if (!c || S < 0 || (S > 0 && (!theta_dev || !ll_dev))) return fail(RVLP_EINVAL, "bad arguments (loglike_dev is required)");
  if (c->P.n_hyper) return fail(RVLP_EINVAL, "GP context not supported here");
  const int64_t k = k_free >= 0 ? k_free : c->P.ndim, n = c->P.n_epochs;
  if (aicc_dev && n - k - 1 == 0) return fail(RVLP_EINVAL, "division by zero");     // Python's ZeroDivisionError, fit.py:1528
  if (S == 0) return RVLP_OK;
  DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = launch_logprob(c, theta_dev, S, nullptr, ll_dev, nullptr, st)) return rc;
  if (!chi2_dev && !aicc_dev && !bic_dev) return RVLP_OK;
  const double corr = aicc_dev ? (double)(2 * k * k + 2 * k) / (double)(n - k - 1) : 0.0;
  int64_t g = (S + 7) / 8;
  if (g > 148 * 16) g = 148 * 16;
  info_criteria_kernel<<<(int)g, 256, 0, st>>>(c->P, theta_dev, S, ll_dev, chi2_dev, aicc_dev, bic_dev, (double)(2 * k), corr,
                                               (double)k * log((double)n));
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

// Pageable -> pinned staging copy.  One thread moves ~8 GB/s, which made the NumPy-in path copy-bound (33 ms against a
// 24 ms kernel at config 3).  A small persistent pool (created on first use, up to 12 threads, never more than three
// quarters of the host's hardware threads) splits a block into 1 MB slices; the caller copies slices too.
// memcpy with non-temporal stores: the destination (the pinned staging buffer) is next read by the DMA engine, not by
// the CPU, so it should not be pulled into the cache first - a regular store's write-allocate makes the copy three
// memory streams instead of two, which is what bounds the host path when eight ranks stage their shards at once.
static void stream_copy(char* dst, const char* src, size_t n) {
#if defined(__x86_64__) && defined(__SSE2__)
  if (n >= 4096) {
    const size_t head = (16 - ((uintptr_t)dst & 15)) & 15;
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 0);
      const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 1);
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 2);
      const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 3);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 0, a);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 1, b);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 2, c);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 3, d);
      src += 64; dst += 64;
    }
    _mm_sfence();
    n -= blocks * 64;
  }
#endif
  memcpy(dst, src, n);
}

class StagePool {
 public:
  static StagePool& get() {
    static StagePool* p = new StagePool();   // leaked on purpose: worker threads must not be joined from a static destructor
    return *p;
  }
  void copy(char* dst, const char* src, size_t nbytes) {
    if (n_workers_ == 0 || nbytes < 4 * kSlice) { stream_copy(dst, src, nbytes); return; }
    auto job = std::make_shared<Job>();      // a late-waking worker keeps ITS job object: it can never touch the next one
    job->dst = dst; job->src = src; job->nbytes = nbytes;
    job->nslices = (nbytes + kSlice - 1) / kSlice;
    job->pending.store((long long)job->nslices);
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = job;
      ++generation_;
    }
    cv_.notify_all();
    run_slices(*job);
    std::unique_lock<std::mutex> lk(job->mu);
    job->done_cv.wait(lk, [&] { return job->pending.load() == 0; });
  }

 private:
  static constexpr size_t kSlice = (size_t)1 << 20;
  struct Job {
    char* dst = nullptr;
    const char* src = nullptr;
    size_t nbytes = 0, nslices = 0;
    std::atomic<size_t> next{0};
    std::atomic<long long> pending{0};
    std::mutex mu;
    std::condition_variable done_cv;
  };
  StagePool() {
    // profiles/r02ac_host_stage.log (config 3, 232 MB pageable, 16 hardware threads): 2 / 4 / 8 / 12 copying threads ->
    // 31 / 25 / 23.5 / 22.7-23.3 ms per call against 21.4 ms device-resident
    unsigned hw = std::thread::hardware_concurrency();
    unsigned n = hw * 3 / 4 > 12 ? 12 : hw * 3 / 4;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {       // one process per GPU (torchrun): share the host's threads
      const int lw = atoi(e);
      if (lw > 1) n = n / (unsigned)lw > 2 ? n / (unsigned)lw : 2;
    }
    if (const char* e = getenv("RVLP_STAGE_THREADS")) n = (unsigned)atoi(e);
    for (unsigned i = 1; i < n; ++i) {
      try {
        std::thread([this] { worker(); }).detach();
        ++n_workers_;
      } catch (...) {
        break;
      }
    }
  }
  static void run_slices(Job& j) {
    for (;;) {
      const size_t i = j.next.fetch_add(1);
      if (i >= j.nslices) break;
      const size_t b = i * kSlice, e = b + kSlice < j.nbytes ? b + kSlice : j.nbytes;
      stream_copy(j.dst + b, j.src + b, e - b);
      if (j.pending.fetch_sub(1) == 1) {
        std::lock_guard<std::mutex> lk(j.mu);
        j.done_cv.notify_all();
      }
    }
  }
  void worker() {
    unsigned long long seen = 0;
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        job = job_;
      }
      run_slices(*job);
    }
  }
  unsigned n_workers_ = 0;
  std::mutex mu_;
  std::condition_variable cv_;
  unsigned long long generation_ = 0;
  std::shared_ptr<Job> job_;
};
constexpr size_t kStageBlock = (size_t)32 << 20;   // sub-block of a pageable chunk (multiple of 8 bytes)
static void staged_copy(double* dst, const double* src, size_t nbytes) {
  StagePool::get().copy(reinterpret_cast<char*>(dst), reinterpret_cast<const char*>(src), nbytes);
}

int rvlp_logprob_batch_host(rvlp_ctx* c, const double* theta_host, int64_t S, double* out_host) {
  if (!c || S < 0 || (S > 0 && (!theta_host || !out_host))) return fail(RVLP_EINVAL, "bad arguments");
  if (S == 0) return RVLP_OK;
  DeviceGuard guard(c->device);
  std::lock_guard<std::mutex> host_lock(c->host_mu);
  if (S > c->cap_samples) {
    cudaFree(c->d_theta); cudaFree(c->d_out);
    if (c->h_theta) cudaFreeHost(c->h_theta);
    if (c->h_out) cudaFreeHost(c->h_out);
    c->d_theta = c->d_out = c->h_theta = c->h_out = nullptr;
    c->cap_samples = 0;
    const size_t nb = sizeof(double) * (size_t)S * (size_t)(c->P.ndim > 0 ? c->P.ndim : 1);
    CUDA_TRY(cudaMalloc((void**)&c->d_theta, nb));
    CUDA_TRY(cudaMalloc((void**)&c->d_out, sizeof(double) * (size_t)S));
    CUDA_TRY(cudaMallocHost((void**)&c->h_theta, nb));
    CUDA_TRY(cudaMallocHost((void**)&c->h_out, sizeof(double) * (size_t)S));
    c->cap_samples = S;
  }
  // already page-locked caller buffers (e.g. torch pin_memory) are used in place; pageable ones are
  // staged through the context's pinned buffers
  auto is_pinned = [](const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  const bool in_pinned = is_pinned(theta_host), out_pinned = is_pinned(out_host);
  const double* src = in_pinned ? theta_host : c->h_theta;
  double* dst = out_pinned ? out_host : c->h_out;
  // chunked, chunks alternating between two streams: the H2D of chunk i+1 (and the staging memcpy) overlaps
  // the kernel of chunk i, the D2H of chunk i overlaps the kernel of chunk i+1
  // Growing chunks (1/64, 1/64, 1/32, 1/16, 1/8, 1/4, 1/2 of the rows): the first staging copy + H2D that nothing can
  // hide is small (3.6 MB at config 3), every later copy is shorter than the kernel it hides behind, and few launches
  // mean few tails.  Results are copied back to a pageable caller buffer chunk by chunk as the chunks finish, so only
  // the last chunk's copy is exposed.
  int64_t bounds[8];
  int nchunks = 0;
  // number of chunks by payload: 7 from 64 MB, fewer below (a chunk under ~1 MB costs more in launches than it hides)
  const size_t total_bytes = sizeof(double) * (size_t)S * (size_t)c->P.ndim;
  int want_chunks = total_bytes >= ((size_t)64 << 20) ? 7 : total_bytes >= ((size_t)16 << 20) ? 5 : total_bytes >= ((size_t)4 << 20) ? 3 : 1;
  if (const char* e = getenv("RVLP_HOST_CHUNKS")) {         // experiments
    const int v = atoi(e);
    if (v >= 1 && v <= 7) want_chunks = v;
  }
  if (S > 4096 && want_chunks > 1) {
    const int64_t parts = (int64_t)1 << (want_chunks - 1);   // 1, 1, 2, 4, ... units
    const int64_t unit = ((S + parts - 1) / parts + 3) & ~(int64_t)3;
    int64_t at = 0;
    for (int i = 0; i < want_chunks && at < S; ++i) {
      at += unit * (i == 0 ? 1 : ((int64_t)1 << (i - 1)));
      bounds[nchunks++] = at < S && i < want_chunks - 1 ? at : S;
    }
  } else {
    bounds[nchunks++] = S;
  }
  cudaEvent_t* done = c->ev_done;          // created once with the context
  int64_t s0 = 0;
  int rc = RVLP_OK;
  for (int ci = 0; ci < nchunks && rc == RVLP_OK; s0 = bounds[ci], ++ci) {
    cudaStream_t st = (ci & 1) ? c->stream2 : c->stream;
    const int64_t n = bounds[ci] - s0;
    const size_t off = (size_t)s0 * (size_t)c->P.ndim;
    const size_t nbytes = sizeof(double) * (size_t)n * (size_t)c->P.ndim;
    // A pageable chunk is staged and sent in sub-blocks: the H2D of block b overlaps the staging memcpy of block b + 1,
    // so a chunk's copy costs ~bytes / (staging rate) instead of bytes / staging + bytes / PCIe - which is what kept
    // the 1/4 and 1/2 chunks' copies longer than the kernels they were meant to hide behind (e2e 24.6 vs 21.4 ms).
    size_t sub = in_pinned ? nbytes : kStageBlock;
    if (const char* e = getenv("RVLP_HOST_STAGE_MB")) {       // experiments
      if (atoi(e) > 0 && !in_pinned) sub = (size_t)atoi(e) << 20;
    }
    bool copy_ok = true;
    for (size_t b0 = 0; b0 < nbytes && copy_ok; b0 += sub) {
      const size_t nb = nbytes - b0 < sub ? nbytes - b0 : sub;
      const size_t o = off + b0 / sizeof(double);
      if (!in_pinned) staged_copy(c->h_theta + o, theta_host + o, nb);
      copy_ok = cudaMemcpyAsync(c->d_theta + o, src + o, nb, cudaMemcpyHostToDevice, st) == cudaSuccess;
    }
    if (!copy_ok) {
      rc = fail(RVLP_ECUDA, "H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
    rc = c->P.n_hyper ? rvlp_gp_logprob_batch(c, c->d_theta + off, n, c->d_out + s0, st)
                      : launch_logprob(c, c->d_theta + off, n, c->d_out + s0, nullptr, nullptr, st);
    if (rc) break;
    if (cudaMemcpyAsync(dst + s0, c->d_out + s0, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaEventRecord(done[ci], st) != cudaSuccess)
      rc = fail(RVLP_ECUDA, "D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (rc == RVLP_OK) {
    s0 = 0;
    for (int ci = 0; ci < nchunks; s0 = bounds[ci], ++ci) {
      if (cudaEventSynchronize(done[ci]) != cudaSuccess) {
        rc = fail(RVLP_ECUDA, "batch failed: %s", cudaGetErrorString(cudaGetLastError()));
        break;
      }
      if (!out_pinned) memcpy(out_host + s0, c->h_out + s0, sizeof(double) * (size_t)(bounds[ci] - s0));
    }
  }
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->stream2);
  return rc;
}

int rvlp_rv_batch(rvlp_ctx* c, const double* theta_dev, int64_t S, const double* times_dev, int64_t T,
                  int32_t component, double* out_dev, void* stream) {
  return rvlp_rv_batch_frozen(c, theta_dev, S, times_dev, T, component, 0, nullptr, nullptr, out_dev, stream);
}

int rvlp_rv_batch_frozen(rvlp_ctx* c, const double* theta_dev, int64_t S, const double* times_dev, int64_t T,
                         int32_t component, int32_t n_frozen, const int32_t* frozen_index,
                         const double* frozen_value, double* out_dev, void* stream) {
  if (!c || S < 0 || T < 0) return fail(RVLP_EINVAL, "bad arguments");
  if (n_frozen < 0 || n_frozen > RVLP_MAX_FROZEN) return fail(RVLP_EINVAL, "n_frozen %d not in [0, %d]", n_frozen, RVLP_MAX_FROZEN);
  if (n_frozen && (!frozen_index || !frozen_value)) return fail(RVLP_EINVAL, "null frozen arrays");
  FrozenParams frozen{};
  frozen.n = n_frozen;
  for (int i = 0; i < n_frozen; ++i) {
    // fit.py:2631-2638: only planet parameters of the active parameterisation can be frozen
    if (frozen_index[i] < 0 || frozen_index[i] >= 5 * c->P.n_planets)
      return fail(RVLP_EINVAL, "frozen_index[%d] = %d is not a planet parameter", i, frozen_index[i]);
    frozen.idx[i] = frozen_index[i];
    frozen.val[i] = frozen_value[i];
  }
  if (component < RVLP_RV_TOTAL || component >= c->P.n_planets) return fail(RVLP_EINVAL, "bad component %d", component);
  if (S == 0 || T == 0) return RVLP_OK;
  if (!theta_dev || !times_dev || !out_dev) return fail(RVLP_EINVAL, "null pointer");
  DeviceGuard guard(c->device);
  int grid = 0;
  const int64_t want = ((S + kG - 1) / kG + kWarps - 1) / kWarps;
  int rc = grid_for(c->device, (const void*)rv_matrix_kernel, c->smem_main, want, &grid);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (want >= 2 * (int64_t)grid) {                           // several batches per warp: dynamic schedule
    std::lock_guard<std::mutex> lock(c->ticket_mu);
    const unsigned slot = c->ticket_next++ % kTicketRing;
    unsigned long long* tickets = c->d_tickets + slot;
    CUDA_TRY(cudaStreamWaitEvent(st, c->ticket_ev[slot], 0));
    CUDA_TRY(cudaMemsetAsync(tickets, 0, sizeof(unsigned long long), st));
    rv_matrix_kernel<<<grid, kThreads, c->smem_main, st>>>(c->P, theta_dev, S, times_dev, T, component, out_dev, frozen,
                                                          tickets);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ticket_ev[slot], st));
    return RVLP_OK;
  }
  rv_matrix_kernel<<<grid, kThreads, c->smem_main, st>>>(c->P, theta_dev, S, times_dev, T, component, out_dev, frozen,
                                                        nullptr);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_walker_check_batch(rvlp_ctx* c, const double* theta_dev, int64_t S, int32_t* status_dev,
                            double* lp_dev, double* lhp_dev, void* stream) {
  if (!c || S < 0 || (S > 0 && (!theta_dev || !status_dev))) return fail(RVLP_EINVAL, "bad arguments");
  if (S == 0) return RVLP_OK;
  DeviceGuard guard(c->device);
  int grid = 0;
  const int64_t want = ((S + kG - 1) / kG + kWarps - 1) / kWarps;
  int rc = grid_for(c->device, (const void*)walker_check_kernel, c->smem_main, want, &grid);
  if (rc) return rc;
  walker_check_kernel<<<grid, kThreads, c->smem_main, (cudaStream_t)stream>>>(c->P, theta_dev, S, status_dev,
                                                                              lp_dev, lhp_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_gp_logprob_batch(rvlp_ctx* c, const double* theta_dev, int64_t S, double* out_dev, void* stream) {
  if (!c || S < 0 || (S > 0 && (!theta_dev || !out_dev))) return fail(RVLP_EINVAL, "bad arguments");
  if (c->P.n_hyper != 4) return fail(RVLP_EINVAL, "context was not created with GP hyperparameters");
  if (S == 0) return RVLP_OK;
  DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int which = gp_choice(c, S, false);
  if (which == GP_BATCH) return launch_gp_batch<false>(c, theta_dev, S, out_dev, nullptr, st);
  if (which == GP_SMEM) return launch_gp_smem<false>(c, theta_dev, S, out_dev, nullptr, st);
  int grid = 0, rc;
  const char* grid_cap = getenv("RVLP_GP_GRID");          // tests / experiments: cap the grid (e.g. 148 = one CTA per SM)
#define RVLP_GP_PIPE(TT)                                                                                    \
  case TT:                                                                                                  \
    rc = grid_for(c->device, (const void*)gp_logprob_pipe_kernel<TT, false>, c->smem_gp_pipe, S, &grid);    \
    if (rc) return rc;                                                                                      \
    if (grid_cap && atoi(grid_cap) > 0 && atoi(grid_cap) < grid) grid = atoi(grid_cap);                     \
    gp_logprob_pipe_kernel<TT, false><<<grid, kThreads, c->smem_gp_pipe, st>>>(c->P, theta_dev, S, out_dev, nullptr); \
    break;
  switch (c->gp_tile) {
    RVLP_GP_PIPE(2) RVLP_GP_PIPE(4) RVLP_GP_PIPE(6) RVLP_GP_PIPE(8) RVLP_GP_PIPE(10)
    default: return fail(RVLP_EUNSUPPORTED, "no GP tile size for %d epochs", c->P.n_epochs);
  }
#undef RVLP_GP_PIPE
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

#ifdef RVLP_GP_TIMING
// experiments only (tools/gp_phase_time.py): read and clear the phase counters of the blocked GP kernel
int rvlp_debug_gp_timing(unsigned long long* out32) {
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(out32, g_gp_timing, 32 * sizeof(unsigned long long)));
  unsigned long long z[32] = {0};
  CUDA_TRY(cudaMemcpyToSymbol(g_gp_timing, z, sizeof(z)));
  return RVLP_OK;
}
int rvlp_debug_gp_pipe_timing(unsigned long long* out64) {
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpyFromSymbol(out64, g_gp_pipe_timing, 64 * sizeof(unsigned long long)));
  unsigned long long z[64] = {0};
  CUDA_TRY(cudaMemcpyToSymbol(g_gp_pipe_timing, z, sizeof(z)));
  return RVLP_OK;
}
#endif

int rvlp_gp_predict_batch(rvlp_ctx* c, const double* theta_dev, int64_t S, const double* times_dev, int64_t T,
                          double* mean_dev, double* chi2_dev, void* stream) {
  if (!c || S < 0 || T < 0 || (S > 0 && !theta_dev)) return fail(RVLP_EINVAL, "bad arguments");
  if (c->P.n_hyper != 4) return fail(RVLP_EINVAL, "context was not created with GP hyperparameters");
  if (T > 0 && (!times_dev || !mean_dev)) return fail(RVLP_EINVAL, "null times / mean pointer");
  if (S == 0 || (T == 0 && !chi2_dev)) return RVLP_OK;
  // N <= 219 epochs: the pipelined register-tile factorisation with the factor kept in shared memory and a blocked back
  // substitution; more epochs: the blocked global-workspace kernel.  Either writes beta = C^-1 r into a per-call
  // scratch; the conditional-mean kernel follows.
  DeviceGuard guard(c->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int N = c->P.n_epochs;
  // beta scratch [S, N]: allocated per call from the device's stream-ordered pool (no host sync in steady state once
  // the pool has grown; concurrent calls on different streams never share it), released on `st` after the mean kernel.
  if (int prc = ensure_pool(c->device)) return prc;
  double* d_beta = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&d_beta, sizeof(double) * (size_t)S * N, st));
  struct BetaFree {
    double* p; cudaStream_t st;
    ~BetaFree() { cudaFreeAsync(p, st); }
  } beta_free{d_beta, st};
  int grid = 0, rc = RVLP_OK;
  const int which = gp_choice(c, S, true);
  if (which == GP_BATCH) {
    if ((rc = launch_gp_batch<true>(c, theta_dev, S, chi2_dev, d_beta, st))) return rc;
  } else if (which == GP_SMEM) {
    if ((rc = launch_gp_smem<true>(c, theta_dev, S, chi2_dev, d_beta, st))) return rc;
  } else {
    const char* grid_cap = getenv("RVLP_GP_GRID");        // tests / experiments: cap the grid
#define RVLP_GP_PRED(TT)                                                                                          \
  case TT:                                                                                                        \
    rc = grid_for(c->device, (const void*)gp_logprob_pipe_kernel<TT, true>, c->smem_gp_pipe_pred, S, &grid);      \
    if (rc) return rc;                                                                                            \
    if (grid_cap && atoi(grid_cap) > 0 && atoi(grid_cap) < grid) grid = atoi(grid_cap);                           \
    gp_logprob_pipe_kernel<TT, true><<<grid, kThreads, c->smem_gp_pipe_pred, st>>>(c->P, theta_dev, S, chi2_dev, d_beta); \
    break;
    switch (c->gp_tile) {
      RVLP_GP_PRED(2) RVLP_GP_PRED(4) RVLP_GP_PRED(6) RVLP_GP_PRED(8) RVLP_GP_PRED(10)
      default: return fail(RVLP_EUNSUPPORTED, "no GP tile size for %d epochs", N);
    }
#undef RVLP_GP_PRED
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
  }
  if (T > 0) {
    const int smem_mean = 4 * ((N + 1) & ~1) * 8;
    if (smem_mean > c->max_smem)
      return fail(RVLP_EUNSUPPORTED, "GP conditional mean needs %d B of shared memory per CTA (> %d): too many epochs",
                  smem_mean, c->max_smem);
    rc = grid_for(c->device, (const void*)gp_mean_kernel, smem_mean, S, &grid);
    if (rc) return rc;
    gp_mean_kernel<<<grid, kThreads, smem_mean, st>>>(c->P, theta_dev, S, d_beta, times_dev, T, mean_dev);
    ++g_launches;
    CUDA_TRY(cudaGetLastError());
  }
  return RVLP_OK;
}

int64_t rvlp_percentile_workspace_bytes(int64_t n_cols, int32_t n_q) {
  if (n_cols < 0 || n_q < 1 || n_q > RVLP_MAX_PERCENTILES) return -1;
  return (int64_t)(band_ws_bytes(n_cols, 2 * n_q) + band_fast_ws_bytes(n_cols, 2 * n_q));
}

int rvlp_percentile_columns(const double* A_dev, int64_t S, int64_t T, const double* q_percent, int32_t n_q,
                            double* out_dev, void* ws_dev, int64_t ws_bytes, int device, void* stream) {
  if (S < 0 || T < 0 || n_q < 1 || n_q > RVLP_MAX_PERCENTILES || !q_percent)
    return fail(RVLP_EINVAL, "bad arguments (n_q must be 1..%d)", RVLP_MAX_PERCENTILES);
  if (S >= ((int64_t)1 << 31)) return fail(RVLP_EUNSUPPORTED, "more than 2^31 - 1 rows");
  for (int i = 0; i < n_q; ++i)
    if (!(q_percent[i] >= 0.0 && q_percent[i] <= 100.0))
      return fail(RVLP_EINVAL, "Percentiles must be in the range [0, 100]");   // numpy's ValueError
  if (T == 0) return RVLP_OK;
  if (!out_dev) return fail(RVLP_EINVAL, "null output");
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 0) {                                   // numpy: NaN (with a RuntimeWarning)
    std::vector<double> nanv((size_t)n_q * T, NAN);
    CUDA_TRY(cudaMemcpyAsync(out_dev, nanv.data(), nanv.size() * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return RVLP_OK;
  }
  if (!A_dev || !ws_dev) return fail(RVLP_EINVAL, "null matrix / workspace");
  const int R = 2 * n_q;
  const int64_t ws_need = (int64_t)(band_ws_bytes(T, R) + band_fast_ws_bytes(T, R));
  if (ws_bytes < ws_need || ((uintptr_t)ws_dev & 255))
    return fail(RVLP_EINVAL, "workspace too small (%lld < %lld bytes) or not 256-byte aligned", (long long)ws_bytes,
                (long long)ws_need);
  // numpy/lib/_function_base_impl.py: q = true_divide(q, 100); method "linear": virtual index = (n - 1) * q
  // (its comment: preferred to _compute_virtual_index(n, q, 1, 1) for rounding); _get_indexes (floor, +1,
  // clamped at the ends); _get_gamma = virtual - previous.
  BandTargets tg{};
  tg.n_q = n_q;
  for (int i = 0; i < n_q; ++i) {
    volatile double q = q_percent[i] / 100.0;
    volatile double virt = (double)(S - 1) * q;
    double prev = floor(virt), next = prev + 1.0;
    if (virt >= (double)(S - 1)) prev = next = -1.0;
    if (virt < 0) prev = next = 0.0;
    volatile double gam = virt - prev;            // with prev = -1 numpy's gamma is virt + 1: harmless, a == b
    const int64_t kp = prev < 0 ? S - 1 : (int64_t)prev, kn = next < 0 ? S - 1 : (int64_t)next;
    tg.k[2 * i] = (uint32_t)kp;
    tg.k[2 * i + 1] = (uint32_t)kn;
    tg.gamma[i] = gam;
  }
  const BandWorkspace W = band_ws_carve(ws_dev, T, R);
  CUDA_TRY(cudaMemsetAsync(W.hist, 0, (size_t)T * R * 256 * 4, st));   // level 0 accumulates into buffer 0
  CUDA_TRY(cudaMemsetAsync(W.mode, 0xff, (size_t)T * 4, st));          // -1: every column streams
  CUDA_TRY(cudaMemsetAsync(W.ncand, 0, (size_t)2 * T * 4, st));        // candidate counters + NaN flags
  // cudaFuncSetAttribute is per DEVICE: one bit per device, as ensure_tables does
  static std::atomic<uint64_t> attr_done{0};
  if (device >= 64 || !((attr_done.load() >> device) & 1)) {
    CUDA_TRY(cudaFuncSetAttribute(band_level_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, band_smem_bytes(1, kMaxTargets)));
    CUDA_TRY(cudaFuncSetAttribute(band_level_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, band_smem_bytes(1, kMaxTargets)));
    CUDA_TRY(cudaFuncSetAttribute(band_level_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, band_smem_bytes(1, kMaxTargets)));
    CUDA_TRY(cudaFuncSetAttribute(band_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinishSmem));
    CUDA_TRY(cudaFuncSetAttribute(band_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  std::max(band_smem_bytes(1, kMaxTargets), kFinishSmem)));
    CUDA_TRY(cudaFuncSetAttribute(band_fast_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kColBlock * (kFastBins + 1) * 4));
    CUDA_TRY(cudaFuncSetAttribute(band_fast_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kColBlock * (kFastBins + 1) * 4));
    CUDA_TRY(cudaFuncSetAttribute(band_fast_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastFinishSmem));
    if (device < 64) attr_done.fetch_or(1ull << device);
  }
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int64_t ncb = (T + kColBlock - 1) / kColBlock;
  // ---- two-pass value-space path (rvlp_bands_fast.cuh): finishes every column whose target bins fit its candidate
  // buffer; the radix path below skips those (a CTA whose columns are all done returns at once)
  const char* fast_env = getenv("RVLP_BANDS_FAST");         // tests: "0" forces the radix path for every column
  if (S >= kFastMinRows && !(fast_env && fast_env[0] == '0')) {
    const BandFastWs F = band_fast_carve(reinterpret_cast<unsigned char*>(ws_dev) + band_ws_bytes(T, R), T, R);
    CUDA_TRY(cudaMemsetAsync(F.hist, 0, (size_t)T * kFastBins * 4, st));
    CUDA_TRY(cudaMemsetAsync(F.ncand, 0, (size_t)T * 4, st));
    band_fast_sample_kernel<<<(unsigned)T, kBandThreads, 0, st>>>(A_dev, S, T, tg, F);
    const int smem_pass = kColBlock * (kFastBins + 1) * 4;
    for (int pass = 0; pass < 2; ++pass) {
      int per_sm = 0;
      if (pass == 0) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, band_fast_pass_kernel<false>, kBandThreads, smem_pass));
      else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, band_fast_pass_kernel<true>, kBandThreads, smem_pass));
      if (per_sm < 1) return fail(RVLP_EUNSUPPORTED, "band kernel does not fit on an SM");
      // Row split: every CTA does the same amount of work, so what counts is how full the last wave of resident CTAs
      // is.  Among the splits that leave >= 2048 rows per CTA (histogram clear / merge stays small) take the one with
      // the best fill - 125 column blocks on 444 slots: 3 slabs fill 84 % of one wave, 7 slabs 98.5 % of two.
      const int64_t slots = (int64_t)sms * per_sm;
      int64_t max_split = (S + 2047) / 2048;
      if (max_split > 4 * slots / ncb + 1) max_split = 4 * slots / ncb + 1;   // at most ~4 waves
      if (max_split > 65535) max_split = 65535;
      if (max_split < 1) max_split = 1;
      int64_t split = 1;
      double best = 0.0;
      for (int64_t sp = 1; sp <= max_split; ++sp) {
        const int64_t ctas = ncb * sp, waves = (ctas + slots - 1) / slots;
        const double fill = (double)ctas / (double)(waves * slots);
        if (fill > best + 1e-9) { best = fill; split = sp; }
      }
      const dim3 grid((unsigned)ncb, (unsigned)split);
      if (pass == 0) {
        band_fast_pass_kernel<false><<<grid, kBandThreads, smem_pass, st>>>(A_dev, S, T, tg, W, F);
        band_fast_plan_kernel<<<(unsigned)((T + kBandWarps - 1) / kBandWarps), kBandThreads, 0, st>>>(T, tg, W, F);
      } else {
        band_fast_pass_kernel<true><<<grid, kBandThreads, smem_pass, st>>>(A_dev, S, T, tg, W, F);
      }
    }
    band_fast_finish_kernel<<<(unsigned)T, kBandThreads, kFastFinishSmem, st>>>(T, tg, W, F, out_dev);
    // what the two passes could not finish: the radix path, one launch (a CTA per column block, all levels inside)
    band_fallback_kernel<<<(unsigned)ncb, kBandThreads, std::max(band_smem_bytes(1, R), kFinishSmem), st>>>(A_dev, S, T, tg, W, out_dev);
    g_launches += 6;
    CUDA_TRY(cudaGetLastError());
    return RVLP_OK;
  }
  for (int level = 0; level <= kLevels; ++level) {
    // Row split: every CTA does the same amount of work, so the grid is sized to ONE wave of resident CTAs
    // (a 1.06-wave grid costs two full rounds); at least 1024 rows per CTA keeps the per-CTA scan / merge small.
    const int smem = band_smem_bytes(level, R);
    int per_sm = 0;
    void (*kern)(const double*, int64_t, int64_t, int, BandTargets, BandWorkspace) =
        band_mode_of(level) == 0 ? band_level_kernel<0> : (band_mode_of(level) == 1 ? band_level_kernel<1> : band_level_kernel<2>);
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBandThreads, smem));
    if (per_sm < 1) return fail(RVLP_EUNSUPPORTED, "band kernel does not fit on an SM");
    int64_t split = (int64_t)sms * per_sm / ncb;
    const int64_t max_split = (S + 1023) / 1024;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    if (split > 65535) split = 65535;
    const dim3 grid((unsigned)ncb, level == kLevels ? 1u : (unsigned)split);
    kern<<<grid, kBandThreads, smem, st>>>(A_dev, S, T, level, tg, W);
    ++g_launches;
  }
  band_finish_kernel<<<(unsigned)((T + kBandWarps - 1) / kBandWarps), kBandThreads, kFinishSmem, st>>>(T, tg, W, out_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_kepler_rv(const double* M_dev, int64_t n, double e, double K, double w, double* rv_dev, int device,
                   void* stream) {
  if (n < 0 || (n > 0 && (!M_dev || !rv_dev))) return fail(RVLP_EINVAL, "bad arguments");
  if (n == 0) return RVLP_OK;
  DeviceGuard guard(device);
  if (int rc = ensure_tables(device)) return rc;
  kepler_rv_kernel<<<simple_grid(n), 256, 0, (cudaStream_t)stream>>>(M_dev, n, e, K, w, rv_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_planet_rv(int32_t par, const double* p5, const double* t_dev, int64_t n, double* rv_dev, int accumulate,
                   int device, void* stream) {
  if (par < 0 || par > 3 || !p5) return fail(RVLP_EINVAL, "bad parameterisation / params");
  const DefaultPars d = to_default(par, p5);
  if (d.invalid) {   // the reference's ValueError messages, param.py:26-82
    if (d.conv_error || d.e < 0) return fail(RVLP_EINVAL, d.e < 0 ? "Invalid eccentricity: %g < 0" : "Invalid eccentricity: %g >= 1.0", d.e);
    if (d.P <= 0) return fail(RVLP_EINVAL, "Invalid period: %g <= 0", d.P);
    if (d.K <= 0) return fail(RVLP_EINVAL, "Invalid semi-amplitude: %g <= 0", d.K);
    if (d.e >= 1.0) return fail(RVLP_EINVAL, "Invalid eccentricity: %g >= 1.0", d.e);
    return fail(RVLP_EINVAL, "Invalid argument of periastron: %g not in [-pi, +pi)", d.w);
  }
  if (n < 0 || (n > 0 && (!t_dev || !rv_dev))) return fail(RVLP_EINVAL, "bad arguments");
  if (n == 0) return RVLP_OK;
  DeviceGuard guard(device);
  if (int rc = ensure_tables(device)) return rc;
  planet_rv_kernel<<<simple_grid(n), 256, 0, (cudaStream_t)stream>>>(d, t_dev, n, rv_dev, accumulate);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_trend_rv(double gd, double gdd, double t0, const double* t_dev, int64_t n, double* rv_dev, int accumulate,
                  int device, void* stream) {
  if (n < 0 || (n > 0 && (!t_dev || !rv_dev))) return fail(RVLP_EINVAL, "bad arguments");
  if (n == 0) return RVLP_OK;
  DeviceGuard guard(device);
  trend_rv_kernel<<<simple_grid(n), 256, 0, (cudaStream_t)stream>>>(gd, gdd, t0, t_dev, n, rv_dev, accumulate);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_convert_to_default(int32_t par, const double* in_dev, int64_t n, double* out_dev, int32_t* valid_dev,
                            int device, void* stream) {
  if (par < 0 || par > 3) return fail(RVLP_EINVAL, "bad parameterisation id");
  if (n < 0 || (n > 0 && (!in_dev || !out_dev))) return fail(RVLP_EINVAL, "bad arguments");
  if (n == 0) return RVLP_OK;
  DeviceGuard guard(device);
  convert_kernel<<<simple_grid(n), 256, 0, (cudaStream_t)stream>>>(par, in_dev, n, out_dev, valid_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_prior_eval(const rvlp_prior* prior, const double* x_dev, int64_t n, double* out_dev, int device,
                    void* stream) {
  if (!prior || prior->kind < 0 || prior->kind > RVLP_PRIOR_BETA) return fail(RVLP_EINVAL, "bad prior");
  if (n < 0 || (n > 0 && (!x_dev || !out_dev))) return fail(RVLP_EINVAL, "bad arguments");
  if (n == 0) return RVLP_OK;
  DeviceGuard guard(device);
  prior_kernel<<<simple_grid(n), 256, 0, (cudaStream_t)stream>>>(*prior, x_dev, n, out_dev);
  ++g_launches;
  CUDA_TRY(cudaGetLastError());
  return RVLP_OK;
}

int rvlp_measure_fp64_peak(int device, int iters, double* flops_per_s, double* ms_out) {
  if (!flops_per_s || iters < 1) return fail(RVLP_EINVAL, "bad arguments");
  DeviceGuard guard(device);
  int sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int grid = sms * 8;
  double* d_out = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d_out, sizeof(double) * (size_t)grid * 256));
  cudaEvent_t a, b;
  CUDA_TRY(cudaEventCreate(&a));
  CUDA_TRY(cudaEventCreate(&b));
  fp64_peak_kernel<<<grid, 256>>>(d_out, iters / 4 + 1, 0.999999, 1e-7);   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CUDA_TRY(cudaEventRecord(a));
    fp64_peak_kernel<<<grid, 256>>>(d_out, iters, 0.999999, 1e-7);
    CUDA_TRY(cudaEventRecord(b));
    CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
    g_launches += 1;
  }
  CUDA_TRY(cudaGetLastError());
  const double flops = 2.0 * 64.0 * (double)iters * (double)grid * 256.0;
  *flops_per_s = flops / (best * 1e-3);
  if (ms_out) *ms_out = best;
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_out);
  return RVLP_OK;
}

}  // extern "C"
