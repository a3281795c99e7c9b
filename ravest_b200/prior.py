"""Prior distributions — host-side records mirroring `ravest.prior` (prior.py:9-511).

Same class names, constructor arguments, validation errors and `repr` as the reference.
The objects hold no arithmetic of their own: the log-density is evaluated on the GPU
(`rvlp_prior_eval`, or fused into the log-probability kernel's per-sample prologue); this
module only validates arguments and precomputes, once, the normalising constants the kernel
needs (with scipy on the host, as the reference does at construction or inside scipy.stats).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import gammaln, log_ndtr, ndtr

PRIOR_FUNCTIONS = ["Uniform", "EccentricityUniform", "Normal", "TruncatedNormal", "HalfNormal",
                   "Rayleigh", "VanEylen19Mixture", "Beta"]

KIND_ID = {name: i for i, name in enumerate(PRIOR_FUNCTIONS)}


def _log_gauss_mass(a: float, b: float) -> float:
    """log(Phi(b) - Phi(a)) evaluated stably (the quantity scipy.stats.truncnorm.logpdf
    subtracts; prior.py:243-246 calls truncnorm.logpdf)."""
    if b <= 0:
        return float(log_ndtr(b) + math.log1p(-math.exp(log_ndtr(a) - log_ndtr(b))))
    if a > 0:
        return float(log_ndtr(-a) + math.log1p(-math.exp(log_ndtr(-b) - log_ndtr(-a))))
    return float(math.log1p(-ndtr(a) - ndtr(-b)))


class _Prior:
    kind: str = ""

    def _p(self) -> list[float]:
        raise NotImplementedError

    def _c(self) -> list[float]:
        raise NotImplementedError

    def pod(self) -> tuple[int, list[float], list[float]]:
        """(kind id, p[4], c[2]) as laid out in include/ravest_b200.h."""
        p = list(map(float, self._p())) + [0.0] * 4
        c = list(map(float, self._c())) + [0.0] * 2
        return KIND_ID[self.kind], p[:4], c[:2]

    def __call__(self, value: float) -> float:
        """Scalar evaluation (used by the setup-time checks, fit.py:464-479) — runs on the GPU."""
        from . import _lib
        return float(_lib.prior_eval(self, np.asarray([value], dtype=np.float64))[0])

    def logpdf_batch(self, values):
        from . import _lib
        return _lib.prior_eval(self, values)


class Uniform(_Prior):
    """prior.py:9-68 — closed interval [lower, upper]."""
    kind = "Uniform"

    def __init__(self, lower: float, upper: float) -> None:
        if not np.isfinite(lower):
            raise ValueError(f"Lower bound must be finite, got {lower}")
        if not np.isfinite(upper):
            raise ValueError(f"Upper bound must be finite, got {upper}")
        if lower >= upper:
            raise ValueError(f"Lower bound ({lower}) must be less than upper bound ({upper})")
        self.lower = lower
        self.upper = upper

    def _p(self):
        return [self.lower, self.upper]

    def _c(self):
        return [-np.log(self.upper - self.lower)]

    def __repr__(self) -> str:
        return f"Uniform(lower={self.lower}, upper={self.upper})"


class EccentricityUniform(_Prior):
    """prior.py:71-125 — half-open interval [0, upper)."""
    kind = "EccentricityUniform"

    def __init__(self, upper: float) -> None:
        if upper > 1:
            raise ValueError("Upper bound of eccentricity must be less than or equal to 1.")
        if upper <= 0:
            raise ValueError("Upper bound of eccentricity must be greater than 0.")
        self.upper = upper

    def _p(self):
        return [self.upper]

    def _c(self):
        return [-np.log(self.upper)]

    def __repr__(self) -> str:
        return f"EccentricityUniform(upper={self.upper})"


class Normal(_Prior):
    """prior.py:128-174."""
    kind = "Normal"

    def __init__(self, mean: float, std: float) -> None:
        if std <= 0:
            raise ValueError(f"Standard deviation must be positive, got {std}")
        self.mean = mean
        self.std = std
        self._log_norm_const = 0.5 * np.log((self.std ** 2) * 2. * np.pi)

    def _p(self):
        return [self.mean, self.std]

    def _c(self):
        return [self._log_norm_const]

    def __repr__(self) -> str:
        return f"Normal(mean={self.mean}, std={self.std})"


class TruncatedNormal(_Prior):
    """prior.py:177-249 — properly normalised on [lower, upper]."""
    kind = "TruncatedNormal"

    def __init__(self, mean: float, std: float, lower: float, upper: float) -> None:
        if std <= 0:
            raise ValueError("Standard deviation must be positive")
        if lower >= upper:
            raise ValueError("Lower bound must be less than upper bound")
        self.mean = mean
        self.std = std
        self.lower = lower
        self.upper = upper
        self._a = (lower - mean) / std
        self._b = (upper - mean) / std

    def _p(self):
        return [self.mean, self.std, self.lower, self.upper]

    def _c(self):
        return [-math.log(math.sqrt(2 * math.pi)) - _log_gauss_mass(self._a, self._b) - math.log(self.std)]

    def __repr__(self) -> str:
        return f"TruncatedNormal(mean={self.mean}, std={self.std}, lower={self.lower}, upper={self.upper})"


def _halfnorm_const(std: float) -> float:
    return 0.5 * math.log(2.0 / math.pi) - math.log(std)


class HalfNormal(_Prior):
    """prior.py:252-306."""
    kind = "HalfNormal"

    def __init__(self, std: float) -> None:
        if std <= 0:
            raise ValueError(f"Standard deviation must be positive, got {std}")
        self.std = float(std)

    def _p(self):
        return [self.std]

    def _c(self):
        return [_halfnorm_const(self.std)]

    def __repr__(self) -> str:
        return f"HalfNormal(std={self.std})"


class Rayleigh(_Prior):
    """prior.py:309-362 — log p(0) = -inf."""
    kind = "Rayleigh"

    def __init__(self, scale: float) -> None:
        if scale <= 0:
            raise ValueError(f"Scale parameter must be positive, got {scale}")
        self.scale = float(scale)

    def _p(self):
        return [self.scale]

    def _c(self):
        return [-math.log(self.scale)]

    def __repr__(self) -> str:
        return f"Rayleigh(scale={self.scale})"


class VanEylen19Mixture(_Prior):
    """prior.py:365-443 — (1-f) HalfNormal + f Rayleigh."""
    kind = "VanEylen19Mixture"

    def __init__(self, sigma_normal: float, sigma_rayleigh: float, f: float) -> None:
        if sigma_normal <= 0:
            raise ValueError(f"sigma_normal must be positive, got {sigma_normal}")
        if sigma_rayleigh <= 0:
            raise ValueError(f"sigma_rayleigh must be positive, got {sigma_rayleigh}")
        if not (0 <= f <= 1):
            raise ValueError(f"Mixing fraction f must be between 0 and 1, got {f}")
        self.sigma_normal = float(sigma_normal)
        self.sigma_rayleigh = float(sigma_rayleigh)
        self.f = float(f)

    def _p(self):
        return [self.sigma_normal, self.sigma_rayleigh, self.f]

    def _c(self):
        return [_halfnorm_const(self.sigma_normal), -math.log(self.sigma_rayleigh)]

    def __repr__(self) -> str:
        return f"VanEylen19Mixture(sigma_normal={self.sigma_normal}, sigma_rayleigh={self.sigma_rayleigh}, f={self.f})"


class Beta(_Prior):
    """prior.py:446-511."""
    kind = "Beta"

    def __init__(self, a: float, b: float) -> None:
        if not a > 0:
            raise ValueError(f"Value of a > 0 required, got {a}")
        if not b > 0:
            raise ValueError(f"Value of b > 0 required, got {b}")
        self.a = float(a)
        self.b = float(b)
        self._log_beta = gammaln(self.a) + gammaln(self.b) - gammaln(self.a + self.b)

    def _p(self):
        return [self.a, self.b]

    def _c(self):
        return [self._log_beta]

    def __repr__(self) -> str:
        return f"Beta(a={self.a}, b={self.b})"


def from_tuple(t) -> _Prior:
    """('Uniform', lo, hi) -> Uniform(lo, hi); used by workload specs and fixtures."""
    return globals()[t[0]](*t[1:])
