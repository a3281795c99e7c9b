"""ravest_b200 — B200-native batched RV log-probability path for ravest.

Drop-in for ONE path of ross-dobson/ravest: Kepler solve -> multi-planet RV -> Gaussian (or
quasi-periodic GP) log-likelihood -> priors + parameterisation corrections, evaluated for a
whole batch of samples by hand-written sm_100a CUDA kernels behind a C ABI
(include/ravest_b200.h).  The module layout mirrors the reference's: `model` (Planet / Trend /
Star), `param` (Parameterisation / Parameter), `prior`, `gp` (GPKernel), `fit` (LogLikelihood /
LogPrior / LogPosterior and GP twins).

Importing the package never touches CUDA; the first compute call loads
`csrc/libravest_b200.so` and raises if it (or a GPU) is missing - there is no CPU fallback.
"""
from . import descriptor, dist, fit, gp, model, param, prior, workloads  # noqa: F401
from ._lib import RvlpError, build, launch_count, load  # noqa: F401

__version__ = "0.1.0"
