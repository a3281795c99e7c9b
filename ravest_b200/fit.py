"""LogLikelihood / LogPrior / LogPosterior (+ GP twins) — the drop-in boundary.

Host mirror of `ravest.fit` (fit.py:3228-3691, 7596-8105): same constructor arguments, same
scalar call conventions (`log_probability(dict) -> float`, `-inf` for invalid values, never
raising for bad values), plus the new batched entry the reference lacks:

    log_probability_batch(theta[S, ndim]) -> [S]

`theta` columns follow `free_params_names` (GP: + `free_hyperparams_names`, fit.py:4978) — this
is emcee's `vectorize=True` contract.  A CUDA fp64 tensor in gives a CUDA tensor out (no host
round trip); a NumPy array in goes through the host-buffer C-ABI call and returns NumPy.
All arithmetic is in the sm_100a kernels; the objects pickle (spawn pools) by dropping their
device context and rebuilding it lazily.
"""
from __future__ import annotations

import logging
import warnings
from typing import Callable, Dict

import numpy as np

from . import _lib
from .descriptor import Descriptor, instrument_indices
from .gp import GPKernel
from .param import Parameterisation
from .prior import Uniform, _Prior


def _par_str(parameterisation) -> str:
    return parameterisation.parameterisation if isinstance(parameterisation, Parameterisation) else str(parameterisation)


def classify_planet_case(letter: str, parameterisation: str, priors: dict, free_params_names) -> str:
    """fit.py:3306-3368."""
    if "secosw" not in parameterisation:
        return "CASE_1"
    if f"secosw_{letter}" not in free_params_names:
        return "CASE_1"
    sc, ss, ek, wk = f"secosw_{letter}", f"sesinw_{letter}", f"e_{letter}", f"w_{letter}"
    if sc in priors and ss in priors:
        a, b = priors[sc], priors[ss]
        if (isinstance(a, Uniform) and isinstance(b, Uniform) and a.lower == -1 and a.upper == 1
                and b.lower == -1 and b.upper == 1):
            return "CASE_2"
        raise NotImplementedError(
            f"Unsupported priors on (secosw_{letter}, sesinw_{letter}): {a!r}, {b!r}. Only Uniform(-1, 1) priors "
            "on (secosw, sesinw) are supported for evidence-correct log-posterior corrections. A separable, "
            "rotationally-symmetric belief about eccentricity can always be re-expressed as a prior on e instead - "
            f"place priors on (e_{letter}, w_{letter}) using one of Ravest's eccentricity priors (HalfNormal, "
            "Rayleigh, VanEylen19Mixture, Beta, EccentricityUniform, TruncatedNormal).")
    elif ek in priors and wk in priors:
        return "CASE_3"
    raise RuntimeError(f"Could not classify log-posterior correction case for planet '{letter}': "
                       "no priors found on either (secosw, sesinw) or (e, w).")


def compute_logprob_corrections(planet_letters, parameterisation: str, priors: dict, free_params_names):
    """fit.py:3370-3397 — (sum of Jacobians, sum of renormalisations, per-planet breakdown)."""
    log_jac = float(np.log(2)) if "secosw" in parameterisation else 0.0      # param.py:428-435
    total_j = total_r = 0.0
    breakdown = {}
    for letter in planet_letters:
        case = classify_planet_case(letter, parameterisation, priors, free_params_names)
        jac = log_jac if case == "CASE_3" else 0.0
        ren = float(np.log(4.0 / np.pi)) if case == "CASE_2" else 0.0
        total_j += jac
        total_r += ren
        breakdown[letter] = {"case": case, "jacobian": jac, "renorm": ren}
        logging.info(f"Planet {letter}: log-posterior correction case {case} (jacobian={jac}, renorm={ren})")
    return total_j, total_r, breakdown


class LogPrior:
    """fit.py:3663-3691 — sum of the priors of the given params, in dict order."""

    def __init__(self, priors: dict) -> None:
        self.priors = priors

    def __call__(self, params: Dict[str, float]) -> float:
        lp = 0
        for name in params:
            lp += self.priors[name](params[name])
        return lp


class _DeviceBacked:
    """Lazy, picklable ownership of an rvlp context."""

    _ctx = None
    device = None          # None = torch's current device at first use

    def _make_descriptor(self) -> Descriptor:
        raise NotImplementedError

    @property
    def ctx(self) -> "_lib.Context":
        if self._ctx is None:
            self._desc = self._make_descriptor()
            idx = instrument_indices(self.instrument, self.unique_instruments)
            self._ctx = _lib.Context(self._desc, self.time, self.vel, self.velerr, idx, device=self.device)
        return self._ctx

    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop("_ctx", None)
        st.pop("_desc", None)
        return st

    def _eval_batch(self, theta):
        torch = _lib._torch()
        if isinstance(theta, torch.Tensor):
            return self.ctx.logprob(theta)
        return self.ctx.logprob_host(np.asarray(theta, dtype=np.float64))


class _SampleMatrices:
    """Posterior-sample matrices, shared by LogPosterior and GPLogPosterior.

    The reference keeps these on `Fitter` and pulls the rows from its emcee sampler; here they take the
    `samples[S, ndim]` block (free_params_names order; GP: + free_hyperparams_names) explicitly and return
    CUDA tensors, so that the S x T matrices stay on the device for the percentile step.
    """

    RV_TREND, RV_TOTAL = -1, -2

    def _columns(self) -> list[str]:
        return list(self.free_params_names) + list(getattr(self, "free_hyperparams_names", []))

    def resolve_freeze_params(self, freeze_params, samples=None, planet_letter=None):
        """fit.py:2586-2688 (_resolve_freeze_params): validate keys, warn, resolve None -> posterior median."""
        if freeze_params is None:
            return None
        valid = {f"{par}_{L}" for par in self.parameterisation.pars for L in self.planet_letters}
        unknown = set(freeze_params) - valid
        if unknown:
            raise ValueError(
                f"Unknown freeze_params key(s): {sorted(unknown)}. Keys must be planet parameters of the active "
                f"parameterisation, i.e. one of {sorted(valid)}.")
        if planet_letter is not None:
            wrong = [k for k in freeze_params if k.rsplit("_", 1)[-1] != planet_letter]
            if wrong:
                warnings.warn(f"freeze_params names parameter(s) for a different planet than '{planet_letter}': "
                              f"{sorted(wrong)}. Freezing is intended for the target planet's parameters "
                              f"(typically P and Tc); check the planet letter.", UserWarning, stacklevel=2)
        fixed_frozen = [k for k in freeze_params if k not in self.free_params_names]
        if fixed_frozen:
            warnings.warn(f"freeze_params names parameter(s) that are already fixed, not free: {sorted(fixed_frozen)}. "
                          f"Freezing only affects parameters that vary across posterior samples, so this has no "
                          f"de-smearing effect (a None value just resolves to the fixed value). Did you mean a free "
                          f"parameter, or pass the wrong name?", UserWarning, stacklevel=2)
        cols = self._columns()
        resolved = {}
        for key, value in freeze_params.items():
            if value is None:
                if key in cols:
                    if samples is None:
                        raise ValueError("samples are needed to resolve a None (posterior median) freeze value")
                    torch = _lib._torch()
                    col = samples[:, cols.index(key)]
                    col = col.detach().cpu().numpy() if isinstance(col, torch.Tensor) else np.asarray(col)
                    resolved[key] = float(np.median(col))
                else:
                    resolved[key] = float(self.fixed_params[key])
            else:
                resolved[key] = float(value)
        return resolved

    def rv_planet_from_samples(self, planet_letter: str, times, samples, freeze_params=None):
        """Fitter.calculate_rv_planet_from_samples (fit.py:2690-2751) -> CUDA tensor [S, len(times)]."""
        frozen = self.resolve_freeze_params(freeze_params, samples, planet_letter)
        return self.ctx.rv_matrix(samples, times, list(self.planet_letters).index(planet_letter), frozen=frozen)

    def rv_trend_from_samples(self, times, samples):
        """Fitter.calculate_rv_trend_from_samples (fit.py:2753-2789)."""
        return self.ctx.rv_matrix(samples, times, self.RV_TREND)

    def rv_total_from_samples(self, times, samples):
        """Fitter.calculate_rv_total_from_samples (fit.py:2791-2824): planets + trend, no gamma."""
        return self.ctx.rv_matrix(samples, times, self.RV_TOTAL)

    def rv_percentile_bands(self, times, samples, component="total", q=(15.85, 50, 84.15), freeze_params=None):
        """The matrix + `np.percentile(..., [15.85, 50, 84.15], axis=0)` pair of fit.py:2235-2240, 2468-2495:
        returns a CUDA tensor [len(q), len(times)]; the S x T matrix never leaves the device."""
        if component == "total":
            m = self.rv_total_from_samples(times, samples)
        elif component == "trend":
            m = self.rv_trend_from_samples(times, samples)
        else:
            m = self.rv_planet_from_samples(component, times, samples, freeze_params)
        return _lib.percentile_columns(m, q)

    # -- walker initialisation checks (row f-3) ----------------------------------------
    def check_walker_positions(self, positions):
        """Per row: (usable, status bits, log_prior) as NumPy arrays - the test each candidate walker goes
        through in fit.py:692-725, 884-902 (astrophysical validity, then a finite log-prior)."""
        st, lp, lhp = self.ctx.walker_check(positions)
        st = st.cpu().numpy()
        return st == 0, st, lp.cpu().numpy(), lhp.cpu().numpy()

    def validate_initial_positions(self, initial_positions) -> None:
        """Fitter.run_mcmc's pre-flight loop (fit.py:1048-1062): raise for the first unusable walker."""
        ok, st, lp, lhp = self.check_walker_positions(initial_positions)
        if ok.all():
            return
        i = int(np.argmin(ok))
        s = int(st[i])
        if s & (_lib.WALKER_NONFINITE | _lib.WALKER_PLANET | _lib.WALKER_JITTER | _lib.WALKER_HYPER):
            why = ("non-finite parameter value" if s & _lib.WALKER_NONFINITE else
                   "planet parameters fail conversion / validity" if s & _lib.WALKER_PLANET else
                   "jitter < 0" if s & _lib.WALKER_JITTER else "GP hyperparameter not finite and > 0")
            raise ValueError(f"Walker {i} has invalid astrophysical parameters: {why}")
        if s & _lib.WALKER_PRIOR:
            raise ValueError(f"Walker {i} is outside prior bounds (log_prior = {lp[i]})")
        raise ValueError(f"Walker {i} is outside hyperprior bounds (log_hyperprior = {lhp[i]})")


class LogLikelihood(_DeviceBacked):
    """fit.py:3529-3660 — white-noise Gaussian log-likelihood of ALL parameters."""

    def __init__(self, time, vel, velerr, instrument, unique_instruments, t0, planet_letters,
                 parameterisation: Parameterisation) -> None:
        self.time, self.vel, self.velerr = time, vel, velerr
        self.instrument, self.unique_instruments, self.t0 = instrument, unique_instruments, t0
        self.planet_letters, self.parameterisation = planet_letters, parameterisation
        self._gamma_keys = [f"g_{i}" for i in self.unique_instruments]
        self._jitter_keys = [f"jit_{i}" for i in self.unique_instruments]

    @property
    def param_names(self) -> list[str]:
        names = []
        for L in self.planet_letters:
            names += [f"{p}_{L}" for p in self.parameterisation.pars]
        return names + ["gd", "gdd"] + self._gamma_keys + self._jitter_keys

    def _make_descriptor(self) -> Descriptor:
        # every model parameter is a column, no priors, no corrections: out = ll + 0
        return Descriptor(self.planet_letters, _par_str(self.parameterisation), {}, {}, self.param_names,
                          self.unique_instruments, self.t0)

    def batch(self, theta):
        """theta[S, n_model] in `param_names` order -> ll[S]."""
        th = _lib.as_cuda_f64(theta)
        ll, _ = self.ctx.logprob_parts(th)
        return ll

    def __call__(self, params: Dict[str, float]) -> float:
        row = np.array([[float(params[n]) for n in self.param_names]])
        return float(self.batch(row)[0])


class LogPosterior(_DeviceBacked, _SampleMatrices):
    """fit.py:3228-3526."""

    def __init__(self, planet_letters, parameterisation: Parameterisation, priors: dict, fixed_params: dict,
                 free_params_names, time, vel, velerr, instrument, unique_instruments, t0) -> None:
        self.planet_letters = planet_letters
        self.parameterisation = parameterisation
        self.priors = priors
        self.fixed_params = fixed_params
        self.free_params_names = free_params_names
        self.time, self.vel, self.velerr = time, vel, velerr
        self.instrument, self.unique_instruments, self.t0 = instrument, unique_instruments, t0
        self.log_likelihood = LogLikelihood(time, vel, velerr, instrument, unique_instruments, t0,
                                            planet_letters, parameterisation)
        self.log_prior = LogPrior(self.priors)
        (self._logprob_jacobian_correction, self._logprob_prior_renorm_correction,
         self._logprob_correction_breakdown) = compute_logprob_corrections(
            planet_letters, _par_str(parameterisation), priors, free_params_names)

    def _make_descriptor(self) -> Descriptor:
        return Descriptor(self.planet_letters, _par_str(self.parameterisation), self.priors, self.fixed_params,
                          self.free_params_names, self.unique_instruments, self.t0,
                          jacobian=self._logprob_jacobian_correction,
                          renorm=self._logprob_prior_renorm_correction)

    # -- the new batched boundary ----------------------------------------------------
    def log_probability_batch(self, theta):
        return self._eval_batch(theta)

    def log_probability_parts_batch(self, theta):
        """(log-likelihood, log-prior) per row, for diagnostics / information criteria."""
        return self.ctx.logprob_parts(_lib.as_cuda_f64(theta))

    # -- information criteria (Fitter.calculate_log_likelihood / chi2 / aicc / bic, fit.py:1361-1554) ----------
    def information_criteria_batch(self, theta) -> dict:
        """{"loglike", "chi2", "aicc", "bic"}: CUDA tensors [S] for rows of free-parameter values (fixed parameters are
        merged in as `Fitter.build_params_dict` does, fit.py:1386-1455); k = number of free parameters (`Fitter.ndim`)."""
        ll, chi2, aicc, bic = self.ctx.info_criteria(theta, k_free=len(self.free_params_names))
        return {"loglike": ll, "chi2": chi2, "aicc": aicc, "bic": bic}

    def _criteria_of(self, params_dict: Dict[str, float]) -> dict:
        row = np.array([[float(params_dict[n]) for n in self.free_params_names]])
        return {k: float(v[0]) for k, v in self.information_criteria_batch(row).items()}

    def calculate_log_likelihood(self, params_dict: Dict[str, float]) -> float:
        """fit.py:1361-1384 (params_dict holds ALL parameters; the fixed ones must equal this object's)."""
        return self._criteria_of(params_dict)["loglike"]

    def calculate_chi2(self, params_dict: Dict[str, float]) -> float:
        """fit.py:1457-1502."""
        return self._criteria_of(params_dict)["chi2"]

    def calculate_aicc(self, params_dict: Dict[str, float]) -> float:
        """fit.py:1504-1530."""
        return self._criteria_of(params_dict)["aicc"]

    def calculate_bic(self, params_dict: Dict[str, float]) -> float:
        """fit.py:1532-1554."""
        return self._criteria_of(params_dict)["bic"]

    # -- the reference's scalar conventions ------------------------------------------
    def log_probability(self, free_params_dict: Dict[str, float]) -> float:
        """fit.py:3448-3495."""
        row = np.array([[float(free_params_dict[n]) for n in self.free_params_names]])
        return float(self.ctx.logprob_host(row)[0])

    def _negative_log_probability_for_MAP(self, free_params_vals) -> float:
        """fit.py:3497-3526."""
        neg = -float(self.ctx.logprob_host(np.asarray(free_params_vals, dtype=np.float64).reshape(1, -1))[0])
        if not np.isfinite(neg):
            return 1e30
        return neg

    def _convert_params_for_prior_evaluation(self, free_params_dict: dict) -> Dict[str, float]:
        """fit.py:3399-3446 (host-side helper used by walker validation, fit.py:1059)."""
        prior_keys, free_keys = set(self.priors.keys()), set(self.free_params_names)
        if prior_keys == free_keys:
            return free_params_dict
        out = {k: v for k, v in free_params_dict.items() if k in prior_keys}
        allp = self.fixed_params | free_params_dict
        for L in self.planet_letters:
            pp = {par: allp[f"{par}_{L}"] for par in self.parameterisation.pars}
            d = self.parameterisation.convert_pars_to_default_parameterisation(pp)
            for k, v in d.items():
                if f"{k}_{L}" in prior_keys:
                    out[f"{k}_{L}"] = v
        return out


class GPLogLikelihood(_DeviceBacked):
    """fit.py:7942-8105 — quasi-periodic GP likelihood of ALL params + hyperparams."""

    def __init__(self, time, vel, velerr, t0, instrument, unique_instruments, planet_letters,
                 parameterisation: Parameterisation, gp_kernel: GPKernel) -> None:
        self.time, self.vel, self.velerr, self.t0 = time, vel, velerr, t0
        self.instrument, self.unique_instruments = instrument, unique_instruments
        self.planet_letters, self.parameterisation, self.gp_kernel = planet_letters, parameterisation, gp_kernel

    @property
    def param_names(self) -> list[str]:
        names = []
        for L in self.planet_letters:
            names += [f"{p}_{L}" for p in self.parameterisation.pars]
        return (names + ["gd", "gdd"] + [f"g_{i}" for i in self.unique_instruments]
                + [f"jit_{i}" for i in self.unique_instruments])

    def _make_descriptor(self) -> Descriptor:
        return Descriptor(self.planet_letters, _par_str(self.parameterisation), {}, {}, self.param_names,
                          self.unique_instruments, self.t0, hyperpriors={}, fixed_hyperparams={},
                          free_hyperparams_names=self.gp_kernel.get_expected_hyperparams())

    def __call__(self, params: Dict[str, float], hyperparams: Dict[str, float]) -> float:
        # fit.py:8062-8105 validates nothing: it squares the jitter (fit.py:8094-8096) and builds the kernel as given
        # (gp.py:145-156: A^2, tau^2 / lambda_e^2, 1 / (2 lambda_p^2), sin^2(pi tau / P) - even in every
        # hyperparameter).  The device entry applies GPLogPosterior's rejections (jit < 0, hyperparameter <= 0:
        # fit.py:7857-7886), so the signs are dropped here: a negative jitter / hyperparameter gives the same finite
        # value as in the reference instead of -inf.  (Zero / non-finite hyperparameters stay -inf; the reference
        # returns NaN or raises inside tinygp there.)
        hyper_names = self.gp_kernel.get_expected_hyperparams()
        names = self.param_names + hyper_names
        both = dict(params) | dict(hyperparams)
        for n in hyper_names + [f"jit_{i}" for i in self.unique_instruments]:
            both[n] = abs(float(both[n]))
        row = np.array([[float(both[n]) for n in names]])
        return float(self.ctx.logprob_host(row)[0])


class GPLogPosterior(_DeviceBacked, _SampleMatrices):
    """fit.py:7596-7939."""

    def __init__(self, planet_letters, parameterisation: Parameterisation, gp_kernel: GPKernel, priors: dict,
                 hyperpriors: dict, fixed_params: dict, fixed_hyperparams: dict, free_params_names,
                 free_hyperparams_names, time, vel, velerr, t0, instrument, unique_instruments) -> None:
        self.planet_letters, self.parameterisation, self.gp_kernel = planet_letters, parameterisation, gp_kernel
        self.priors, self.hyperpriors = priors, hyperpriors
        self.fixed_params, self.fixed_hyperparams = fixed_params, fixed_hyperparams
        self.free_params_names, self.free_hyperparams_names = free_params_names, free_hyperparams_names
        self.time, self.vel, self.velerr, self.t0 = time, vel, velerr, t0
        self.instrument, self.unique_instruments = instrument, unique_instruments
        self.gp_log_likelihood = GPLogLikelihood(time, vel, velerr, t0, instrument, unique_instruments,
                                                 planet_letters, parameterisation, gp_kernel)
        self.log_prior = LogPrior(self.priors)
        self.log_hyperprior = LogPrior(self.hyperpriors)
        (self._logprob_jacobian_correction, self._logprob_prior_renorm_correction,
         self._logprob_correction_breakdown) = compute_logprob_corrections(
            planet_letters, _par_str(parameterisation), priors, free_params_names)

    def _make_descriptor(self) -> Descriptor:
        return Descriptor(self.planet_letters, _par_str(self.parameterisation), self.priors, self.fixed_params,
                          self.free_params_names, self.unique_instruments, self.t0,
                          hyperpriors=self.hyperpriors, fixed_hyperparams=self.fixed_hyperparams,
                          free_hyperparams_names=self.free_hyperparams_names,
                          jacobian=self._logprob_jacobian_correction,
                          renorm=self._logprob_prior_renorm_correction)

    def log_probability_batch(self, theta):
        return self._eval_batch(theta)

    def log_probability(self, combined: Dict[str, float]) -> float:
        """fit.py:7836-7901."""
        names = list(self.free_params_names) + list(self.free_hyperparams_names)
        row = np.array([[float(combined[n]) for n in names]])
        return float(self.ctx.logprob_host(row)[0])

    # -- GP conditioning (row f-4) ------------------------------------------------------
    def gp_mean_from_samples(self, times, samples):
        """The per-sample loop of fit.py:6383-6414 / 7494-7554: GP conditional mean at `times`, conditioned on
        vel - gamma - planets - trend, for every row of samples -> CUDA tensor [S, len(times)]."""
        return self.ctx.gp_predict(samples, times)

    def chi2_batch(self, samples):
        """GPFitter._compute_gp_chi2 (fit.py:5386-5429) for every row: r^T C^-1 r."""
        return self.ctx.gp_predict(samples, np.empty(0), want_chi2=True)[1]

    def _negative_log_probability_for_MAP(self, vals) -> float:
        """fit.py:7903-7939."""
        neg = -float(self.ctx.logprob_host(np.asarray(vals, dtype=np.float64).reshape(1, -1))[0])
        return 1e30 if not np.isfinite(neg) else neg


def from_spec(spec: dict):
    """Build a LogPosterior / GPLogPosterior from a workload spec (ravest_b200/workloads.py)."""
    from . import prior as P
    params = spec["params"]
    free = [k for k, (_, fx) in params.items() if not fx]
    fixed = {k: v for k, (v, fx) in params.items() if fx}
    priors = {k: P.from_tuple(v) for k, v in spec["priors"].items()}
    inst = np.asarray(spec["instrument"])
    args = dict(time=np.ascontiguousarray(spec["time"], dtype=np.float64),
                vel=np.ascontiguousarray(spec["vel"], dtype=np.float64),
                velerr=np.ascontiguousarray(spec["velerr"], dtype=np.float64),
                instrument=inst, unique_instruments=np.unique(inst), t0=spec["t0"])
    par = Parameterisation(spec["parameterisation"])
    if "hyperparams" in spec:
        hp = spec["hyperparams"]
        return GPLogPosterior(list(spec["planet_letters"]), par, GPKernel("Quasiperiodic"), priors,
                              {k: P.from_tuple(v) for k, v in spec["hyperpriors"].items()}, fixed,
                              {k: v for k, (v, fx) in hp.items() if fx}, free,
                              [k for k, (_, fx) in hp.items() if not fx], **args)
    return LogPosterior(list(spec["planet_letters"]), par, priors, fixed, free, **args)
