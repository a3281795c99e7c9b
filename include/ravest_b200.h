/*
 * ravest_b200.h — C ABI of the B200-native batched RV log-probability path.
 *
 * The reference (ross-dobson/ravest v0.4.0) is pure Python: it has no FFI for this path.
 * Each entry point below therefore names the Python callable it replaces
 * (paths relative to /root/reference/src/ravest/); INTEGRATION.md shows the ctypes stub a
 * ravest maintainer would add.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every `*_dev` pointer is device memory on the context's device, owned by the caller
 *     (in practice a torch.Tensor's data_ptr()); the library owns only its contexts;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are
 *     stream-ordered, never synchronise the host unless the name says `_host`;
 *   - return value: RVLP_OK or a negative RVLP_E* code, message via rvlp_last_error();
 *     invalid SAMPLES are data (-inf in the output), never errors;
 *   - all floating point is IEEE fp64.
 */
#ifndef RAVEST_B200_H
#define RAVEST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVLP_ABI_VERSION 1

enum {
  RVLP_OK = 0,
  RVLP_EINVAL = -1,   /* bad argument / malformed descriptor            */
  RVLP_ECUDA = -2,    /* CUDA runtime error (see rvlp_last_error)       */
  RVLP_ENOMEM = -3,
  RVLP_EUNSUPPORTED = -4
};

/* param.py:5-10 ALLOWED_PARAMETERISATIONS, in that order */
enum {
  RVLP_PAR_PKEWTP = 0,      /* "P K e w Tp"            */
  RVLP_PAR_PKEWTC = 1,      /* "P K e w Tc"            */
  RVLP_PAR_PKSECTP = 2,     /* "P K secosw sesinw Tp"  */
  RVLP_PAR_PKSECTC = 3      /* "P K secosw sesinw Tc"  */
};

/* prior.py:6 PRIOR_FUNCTIONS, in that order */
enum {
  RVLP_PRIOR_UNIFORM = 0,          /* p = {lower, upper}                 c = {-ln(upper-lower)}            */
  RVLP_PRIOR_ECC_UNIFORM = 1,      /* p = {upper}                        c = {-ln(upper)}                  */
  RVLP_PRIOR_NORMAL = 2,           /* p = {mean, std}                    c = {0.5 ln(2 pi std^2)}          */
  RVLP_PRIOR_TRUNC_NORMAL = 3,     /* p = {mean, std, lower, upper}      c = {-ln sqrt(2pi) - ln std - ln(Phi(b)-Phi(a))} */
  RVLP_PRIOR_HALF_NORMAL = 4,      /* p = {std}                          c = {0.5 ln(2/pi) - ln std}       */
  RVLP_PRIOR_RAYLEIGH = 5,         /* p = {scale}                        c = {-ln scale}                   */
  RVLP_PRIOR_VANEYLEN19 = 6,       /* p = {sigma_normal, sigma_rayleigh, f}  c = {hn const, -ln sigma_rayleigh} */
  RVLP_PRIOR_BETA = 7              /* p = {a, b}                         c = {ln B(a,b)}                   */
};

/* what value a prior is evaluated on (fit.py:3399-3446 _convert_params_for_prior_evaluation) */
enum {
  RVLP_TARGET_COLUMN = 0,   /* theta[:, index]                                                   */
  RVLP_TARGET_P = 1,        /* planet `index`'s default-space P  (converted, param.py:299-362)   */
  RVLP_TARGET_K = 2,
  RVLP_TARGET_E = 3,
  RVLP_TARGET_W = 4,
  RVLP_TARGET_TP = 5
};

typedef struct rvlp_prior {
  int32_t kind;       /* RVLP_PRIOR_*                                                    */
  int32_t target;     /* RVLP_TARGET_*                                                   */
  int32_t index;      /* theta column, or planet number for derived targets              */
  int32_t is_hyper;   /* 1: belongs to the GP hyper-prior sum (fit.py:7884)              */
  double p[4];        /* distribution parameters, see RVLP_PRIOR_*                       */
  double c[2];        /* host-precomputed constants (scipy on the host, as the reference) */
} rvlp_prior;

/*
 * Problem descriptor = everything ravest.fit.LogPosterior.__init__ (fit.py:3234-3304) holds,
 * with all string/dict logic resolved on the host.
 *
 * Model parameters are numbered: planet k's five parameters in Parameterisation.pars order
 * (param.py:151) at 5k..5k+4, then gd, gdd (5n), (5n+1), then g_<inst> for each instrument
 * in np.unique order (fit.py:3591), then jit_<inst> likewise; n_model = 5n + 2 + 2 n_inst.
 * GP problems append gp_amp, gp_lambda_e, gp_lambda_p, gp_period (gp.py:37) as
 * n_model..n_model+3.  src_col[i] >= 0 reads theta[:, src_col[i]]; -1 uses src_const[i]
 * (a fixed parameter, fit.py:3465 `fixed_params | free_params_dict`).
 */
typedef struct rvlp_desc {
  int32_t abi_version;        /* RVLP_ABI_VERSION */
  int32_t n_planets;
  int32_t parameterisation;   /* RVLP_PAR_* */
  int32_t n_inst;
  int32_t ndim;               /* number of theta columns */
  int32_t n_priors;
  int32_t n_hyper;            /* 0 = white-noise likelihood, 4 = quasi-periodic GP */
  int32_t reserved;
  double t0;                  /* trend reference time (model.py:467-509) */
  double jacobian;            /* sum of per-planet log|J| corrections  (fit.py:3370-3397) */
  double renorm;              /* sum of per-planet prior renormalisations                 */
  const int32_t* src_col;     /* [n_model + n_hyper] */
  const double* src_const;    /* [n_model + n_hyper] */
  const rvlp_prior* priors;   /* [n_priors], evaluated and summed in this order (fit.py:3685-3691) */
} rvlp_desc;

typedef struct rvlp_ctx rvlp_ctx;

/* component selector for rvlp_rv_batch */
#define RVLP_RV_TREND (-1)   /* Fitter.calculate_rv_trend_from_samples  (fit.py:2753-2789) */
#define RVLP_RV_TOTAL (-2)   /* Fitter.calculate_rv_total_from_samples  (fit.py:2791-2824) */
                             /* >= 0: that planet, calculate_rv_planet_from_samples (fit.py:2690-2751) */

int rvlp_abi_version(void);
const char* rvlp_last_error(void);

/* Replaces LogPosterior.__init__ / LogLikelihood.__init__ (fit.py:3234-3304, 3535-3598):
 * copies the HOST epoch arrays (time, vel, velerr, per-epoch instrument index) to `device`
 * once; they stay resident for the context's lifetime. */
int rvlp_ctx_create(const rvlp_desc* desc, const double* time, const double* vel,
                    const double* velerr, const int32_t* inst_idx, int64_t n_epochs,
                    int device, rvlp_ctx** out);
void rvlp_ctx_destroy(rvlp_ctx* ctx);

/* Replaces LogPosterior.log_probability (fit.py:3448-3495), batched: theta_dev is [S, ndim]
 * row-major in free_params_names order, out_dev is [S]. -inf/NaN semantics as the reference. */
int rvlp_logprob_batch(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples,
                       double* out_dev, void* stream);

/* Multi-GPU (one process per GPU; SURVEY.md 8e "a single all-gather of log-probs per ensemble half-step"): the
 * same call with the all-gather FUSED into the kernel.  outs[i] (i < n_outs <= 8) is rank i's gathered [S_total]
 * vector, mapped into this process (rvlp_peer_open); every log-probability of this rank's block goes to
 * outs[i][row_offset + s] of every rank by an NVLink peer store.  rvlp_peer_barrier, queued after it on the same
 * stream, publishes this rank's arrival in every rank's flag block and waits (on the device, at most timeout_ms; a
 * timeout sets word [8] of this rank's flag block) until every rank has arrived: what follows on the stream may
 * read the gathered vector.  The reference has no counterpart (emcee's pool maps walkers over processes,
 * fit.py:1069-1075); ravest_b200/dist.py drives it and falls back to one NCCL all-gather when IPC is refused. */
int rvlp_logprob_batch_peers(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples, double* const* outs,
                             int32_t n_outs, int64_t row_offset, void* stream);
int rvlp_peer_barrier(int device, void* const* flag_blocks, int32_t n_ranks, int32_t my_rank, uint64_t epoch,
                      int64_t timeout_ms, void* stream);
/* Device memory that other processes on this box can map: cudaMalloc + a 64-byte CUDA IPC handle (zero-filled);
 * rvlp_peer_open maps another process's block into this one.  Close what was opened, free what was allocated. */
int rvlp_peer_alloc(int device, int64_t bytes, void** dev_ptr, void* handle64);
int rvlp_peer_open(int device, const void* handle64, void** dev_ptr);
int rvlp_peer_close(int device, void* dev_ptr);
int rvlp_peer_free(int device, void* dev_ptr);

/* Optional, synchronous: times the compiled shapes of the log-probability kernel on (up to 2^18 of) the caller's
 * rows - on `stream`, i.e. ordered after whatever wrote theta_dev there - and keeps the fastest for this context
 * (median of five launches each; the alternative shape must win by 2 %); *chosen (may be NULL) gets its index.
 * The choice changes the speed only - every shape produces identical bits (tests/test_gpu_parity.py). */
int rvlp_ctx_autotune(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples, void* stream, int32_t* chosen);

/* Diagnostic: force shape `variant` (0 or 1) of the log-probability kernel for this context. */
int rvlp_ctx_set_variant(rvlp_ctx* ctx, int32_t variant);

/* Same call for HOST buffers (what a NumPy caller such as emcee holds): pinned staging,
 * H2D of theta, kernel, D2H of out, one host synchronisation at the end. */
int rvlp_logprob_batch_host(rvlp_ctx* ctx, const double* theta_host, int64_t n_samples,
                            double* out_host);

/* Diagnostic split of the same evaluation: ll (LogLikelihood.__call__, fit.py:3600-3660) and
 * lp (LogPrior.__call__ on the converted params, fit.py:3672-3691); either pointer may be NULL. */
int rvlp_logprob_parts_batch(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples,
                             double* loglike_dev, double* logprior_dev, void* stream);

/* Replaces Fitter.calculate_log_likelihood / calculate_chi2 / calculate_aicc / calculate_bic (fit.py:1361-1384,
 * 1457-1554), batched over rows of theta (context column order; the fixed parameters come from the descriptor, as
 * build_params_dict does, fit.py:1386-1455).  loglike_dev [S] is required (the others derive from it, exactly as the
 * reference works backwards from LogLikelihood: chi2 = -2 ll - sum ln(2 pi var)); chi2_dev / aicc_dev / bic_dev [S] may
 * be NULL.  k_free = the reference's `self.ndim` (number of free parameters); < 0 uses the context's ndim.
 * RVLP_EINVAL "division by zero" when n_epochs - k - 1 == 0 and aicc is requested (Python raises there). */
int rvlp_info_criteria_batch(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples, int32_t k_free,
                             double* loglike_dev, double* chi2_dev, double* aicc_dev, double* bic_dev, void* stream);

/* Replaces Fitter.calculate_rv_{planet,trend,total}_from_samples (fit.py:2690-2824):
 * out_dev is [S, n_times] row-major; rows whose planet parameters are invalid are NaN
 * (the reference raises ValueError there). No gamma offsets, as in the reference. */
int rvlp_rv_batch(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples,
                  const double* times_dev, int64_t n_times, int32_t component,
                  double* out_dev, void* stream);

/* rvlp_rv_batch with per-call parameter overrides: Fitter._resolve_freeze_params +
 * _calculate_rv_planet_from_samples (fit.py:2586-2688, 2726-2751: `params.update(resolved_freeze)` for every
 * sample).  frozen_index[i] (HOST array) is a model-parameter number as defined at rvlp_desc, frozen_value[i]
 * the value every sample uses instead of its own; n_frozen <= RVLP_MAX_FROZEN. */
#define RVLP_MAX_FROZEN 16
int rvlp_rv_batch_frozen(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples,
                         const double* times_dev, int64_t n_times, int32_t component,
                         int32_t n_frozen, const int32_t* frozen_index, const double* frozen_value,
                         double* out_dev, void* stream);

/* Replaces the per-row checks of Fitter.generate_initial_walker_positions_{random,around_point} and
 * Fitter.run_mcmc (fit.py:692-725, 884-902, 1048-1062; GPFitter twins fit.py:4500-4540, 4750-4790, 4950-4990):
 * _validate_astrophysical_validity (fit.py:260-293) followed by "log-prior of the converted parameters is
 * finite".  status_dev [S] gets 0 for a usable row, else an OR of the bits below; logprior_dev /
 * loghyperprior_dev [S] (either may be NULL) get LogPrior / log_hyperprior. */
enum {
  RVLP_WALKER_NONFINITE = 1,    /* some free or fixed parameter is NaN / inf        fit.py:262-265      */
  RVLP_WALKER_PLANET = 2,       /* conversion or param.py:88-105 validity failed    fit.py:268-276      */
  RVLP_WALKER_JITTER = 4,       /* jit_<inst> < 0                                   fit.py:289-293      */
  RVLP_WALKER_PRIOR = 8,        /* log-prior not finite                             fit.py:717-720      */
  RVLP_WALKER_HYPER = 16,       /* GP hyperparameter not finite or <= 0             gp.py:82-108        */
  RVLP_WALKER_HYPERPRIOR = 32   /* log-hyperprior not finite                        fit.py:4534-4537    */
};
int rvlp_walker_check_batch(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples,
                            int32_t* status_dev, double* logprior_dev, double* loghyperprior_dev,
                            void* stream);

/* Replaces GPLogPosterior.log_probability (fit.py:7836-7901), batched; requires a descriptor
 * with n_hyper == 4. theta columns: free_params_names + free_hyperparams_names (fit.py:4978). */
int rvlp_gp_logprob_batch(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples,
                          double* out_dev, void* stream);

/* Replaces the per-sample loops of GPFitter's posterior predictions (fit.py:6383-6414, 7494-7554:
 * `gp.condition(y = vel - gamma - planets - trend, X_test = times)` -> conditional mean) and
 * GPFitter._compute_gp_chi2 (fit.py:5386-5429).  mean_dev is [S, n_times] row-major (may be NULL with
 * n_times == 0), chi2_dev [S] (may be NULL).  Rows whose planet parameters or hyperparameters are invalid are
 * NaN (the reference raises there).  Requires n_hyper == 4.  A scratch of n_samples x n_epochs doubles (beta = C^-1 r
 * per sample) is taken from the device's stream-ordered pool for the duration of the call (cudaMallocAsync on
 * `stream`): concurrent calls on different streams of one context are independent. */
int rvlp_gp_predict_batch(rvlp_ctx* ctx, const double* theta_dev, int64_t n_samples,
                          const double* times_dev, int64_t n_times, double* mean_dev,
                          double* chi2_dev, void* stream);

/* Replaces `np.percentile(matrix, q, axis=0)` on the per-sample RV matrices (fit.py:2239-2240, 2493-2495;
 * GP twins fit.py:6440-6450): matrix_dev is [n_rows, n_cols] row-major (the layout rvlp_rv_batch writes),
 * q_percent a HOST array of n_q <= RVLP_MAX_PERCENTILES percentiles in [0, 100], out_dev [n_q, n_cols].
 * numpy's default "linear" method, bit-exact (exact order statistics + numpy's _lerp arithmetic); a column
 * holding a NaN gives NaN, as numpy.  workspace_dev: caller-owned scratch of at least
 * rvlp_percentile_workspace_bytes(n_cols, n_q) bytes, 256-byte aligned. */
#define RVLP_MAX_PERCENTILES 8
int64_t rvlp_percentile_workspace_bytes(int64_t n_cols, int32_t n_q);
int rvlp_percentile_columns(const double* matrix_dev, int64_t n_rows, int64_t n_cols,
                            const double* q_percent, int32_t n_q, double* out_dev,
                            void* workspace_dev, int64_t workspace_bytes, int device, void* stream);

/* Replaces ravest.model._compute_rv / _njit_kepler_rv (model.py:173-243) for a batch of
 * default-space orbits: M_dev [n] mean anomalies, one (e, K, w) triple -> rv_dev [n]. */
int rvlp_kepler_rv(const double* M_dev, int64_t n, double e, double K, double w,
                   double* rv_dev, int device, void* stream);

/* Replaces Planet.__init__ + Planet.radial_velocity (model.py:259-275, 329-354): params5 are
 * the five HOST values in Parameterisation.pars order. Returns RVLP_EINVAL (the reference's
 * ValueError) when they fail param.py:88-105, with the reason in rvlp_last_error(). */
int rvlp_planet_rv(int32_t parameterisation, const double* params5, const double* t_dev,
                   int64_t n, double* rv_dev, int accumulate, int device, void* stream);

/* Replaces Trend.radial_velocity (model.py:493-509). */
int rvlp_trend_rv(double gd, double gdd, double t0, const double* t_dev, int64_t n,
                  double* rv_dev, int accumulate, int device, void* stream);

/* Replaces Parameterisation.convert_pars_to_default_parameterisation (param.py:299-362) for a
 * batch: in_dev [n, 5] in pars order -> out_dev [n, 5] = P K e w Tp; valid_dev [n] (may be
 * NULL) gets 1 where validate_default_parameterisation_params (param.py:88-105) passes. */
int rvlp_convert_to_default(int32_t parameterisation, const double* in_dev, int64_t n,
                            double* out_dev, int32_t* valid_dev, int device, void* stream);

/* Replaces the prior callables (prior.py:49-508): x_dev [n] -> out_dev [n]. */
int rvlp_prior_eval(const rvlp_prior* prior, const double* x_dev, int64_t n, double* out_dev,
                    int device, void* stream);

/* Measurement helpers (not part of the reference's surface): dependent-free DFMA loop to
 * measure the fp64 roofline on this device; returns achieved FLOP/s in *flops_per_s. */
int rvlp_measure_fp64_peak(int device, int iters, double* flops_per_s, double* ms);
/* number of kernel launches issued by this library since load (bench `gpu_launches`) */
int64_t rvlp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RAVEST_B200_H */
