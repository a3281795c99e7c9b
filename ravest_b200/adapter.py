"""Bind the REFERENCE's own objects to the B200 path.

`ravest.fit.LogPosterior` / `GPLogPosterior` instances (fit.py:3234-3304, 7596-7680) built by an unmodified
`ravest.fit.Fitter` carry `ravest.prior.*` objects (prior.py:9-511) and a `ravest.param.Parameterisation`.
Nothing here imports ravest: the objects are duck-typed by class NAME and by the attributes the reference's
constructors set, so the adapter works with whatever ravest the caller has imported.

    post = from_reference(lp)                 # ravest_b200.fit.LogPosterior / GPLogPosterior twin
    f    = BatchedLogPosterior(lp)            # emcee vectorize=True callable: coords[n, ndim] -> n log-probs
    desc, keep = compile_descriptor(lp)       # raw rvlp_desc for the ctypes stub of INTEGRATION.md §2

The twin shares the reference object's arrays (no copy until `rvlp_ctx_create` uploads them) and the reference's
own correction constants (`_logprob_jacobian_correction`, `_logprob_prior_renorm_correction`, fit.py:3370-3397);
they are cross-checked against this repo's classification so that a drift in either is an error, not a silent bias.
"""
from __future__ import annotations

import numpy as np

from . import fit as _fit
from . import prior as _prior
from .descriptor import Descriptor
from .gp import GPKernel
from .param import Parameterisation

# class name -> constructor attributes, in constructor order (prior.py:39-47, 99-104, 151-155, 215-224, 282-285,
# 338-341, 406-416, 480-486)
_PRIOR_ATTRS = {
    "Uniform": ("lower", "upper"),
    "EccentricityUniform": ("upper",),
    "Normal": ("mean", "std"),
    "TruncatedNormal": ("mean", "std", "lower", "upper"),
    "HalfNormal": ("std",),
    "Rayleigh": ("scale",),
    "VanEylen19Mixture": ("sigma_normal", "sigma_rayleigh", "f"),
    "Beta": ("a", "b"),
}


def convert_prior(p) -> "_prior._Prior":
    """A `ravest.prior.*` instance (or one of this package's) -> the `ravest_b200.prior` record of the same law."""
    if isinstance(p, _prior._Prior):
        return p
    name = type(p).__name__
    attrs = _PRIOR_ATTRS.get(name)
    if attrs is None:
        raise NotImplementedError(
            f"prior {p!r} (class {name}) has no device implementation; supported: {sorted(_PRIOR_ATTRS)}")
    try:
        args = [getattr(p, a) for a in attrs]
    except AttributeError as ex:
        raise TypeError(f"{name} prior object lacks attribute {ex.name!r}; expected {attrs}") from None
    return getattr(_prior, name)(*args)


def _parameterisation(par) -> Parameterisation:
    if isinstance(par, Parameterisation):
        return par
    return Parameterisation(getattr(par, "parameterisation", par))


def _is_gp(lp) -> bool:
    return hasattr(lp, "hyperpriors") and hasattr(lp, "free_hyperparams_names")


def from_reference(lp, check_corrections: bool = True):
    """Build this package's LogPosterior / GPLogPosterior from a reference (or duck-typed) posterior object."""
    par = _parameterisation(lp.parameterisation)
    priors = {k: convert_prior(v) for k, v in lp.priors.items()}          # keeps the dict order (fit.py:3685-3691)
    common = dict(time=np.ascontiguousarray(lp.time, dtype=np.float64),
                  vel=np.ascontiguousarray(lp.vel, dtype=np.float64),
                  velerr=np.ascontiguousarray(lp.velerr, dtype=np.float64),
                  instrument=np.asarray(lp.instrument), unique_instruments=np.asarray(lp.unique_instruments),
                  t0=float(lp.t0))
    fixed = {k: float(v) for k, v in lp.fixed_params.items()}
    if _is_gp(lp):
        kt = getattr(getattr(lp, "gp_kernel", None), "kernel_type", "Quasiperiodic")
        post = _fit.GPLogPosterior(list(lp.planet_letters), par, GPKernel(kt), priors,
                                   {k: convert_prior(v) for k, v in lp.hyperpriors.items()}, fixed,
                                   {k: float(v) for k, v in lp.fixed_hyperparams.items()},
                                   list(lp.free_params_names), list(lp.free_hyperparams_names), **common)
    else:
        post = _fit.LogPosterior(list(lp.planet_letters), par, priors, fixed, list(lp.free_params_names), **common)
    if check_corrections and hasattr(lp, "_logprob_jacobian_correction"):
        ref = (float(lp._logprob_jacobian_correction), float(lp._logprob_prior_renorm_correction))
        ours = (post._logprob_jacobian_correction, post._logprob_prior_renorm_correction)
        if ref != ours:
            raise RuntimeError(f"log-posterior corrections disagree: reference (jacobian, renorm) = {ref}, "
                               f"ravest_b200 = {ours} (fit.py:3370-3397)")
    return post


def compile_descriptor(lp):
    """(rvlp_desc ctypes struct, keep-alive) for a reference posterior object: what INTEGRATION.md's stub passes to
    `rvlp_ctx_create`.  The keep-alive owns the tables the struct points into."""
    post = from_reference(lp)
    desc: Descriptor = post._make_descriptor()
    return desc.pod, desc


class BatchedLogPosterior:
    """`log_prob_fn` for `emcee.EnsembleSampler(nwalkers, ndim, f, vectorize=True)` built from the reference's
    posterior object (replaces the per-walker dict path of fit.py:1067-1075 / 4984-4990).

    `f(coords[n, ndim]) -> n` log-probabilities, columns in `free_params_names` (+ `free_hyperparams_names`) order.
    NumPy in -> NumPy out (host-buffer C-ABI call); CUDA tensor in -> CUDA tensor out.  Picklable: the device
    context is dropped and rebuilt lazily (a CUDA context cannot cross a spawn pool, fit.py:1069).
    """

    def __init__(self, lp, device: int | None = None) -> None:
        self.post = from_reference(lp)
        self.post.device = device
        self.parameter_names = list(self.post.free_params_names) + list(getattr(self.post, "free_hyperparams_names", []))
        self.ndim = len(self.parameter_names)

    def __call__(self, coords):
        if getattr(coords, "ndim", 2) == 1:            # emcee also probes single positions
            return self.post.log_probability_batch(np.asarray(coords, dtype=np.float64).reshape(1, -1))[0]
        return self.post.log_probability_batch(coords)

    def __getattr__(self, name):
        # sample-matrix rows (rv_*_from_samples, rv_percentile_bands, check_walker_positions, gp_mean_from_samples, ...)
        if name in ("post", "__setstate__", "__getstate__"):
            raise AttributeError(name)
        return getattr(self.post, name)

    # the reference's scalar conventions stay available on the same object
    def log_probability(self, free_params_dict) -> float:
        return self.post.log_probability(free_params_dict)

    def _negative_log_probability_for_MAP(self, vals) -> float:
        return self.post._negative_log_probability_for_MAP(vals)
