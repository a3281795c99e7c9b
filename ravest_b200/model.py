"""Planet / Trend / Star — host mirror of `ravest.model` (model.py:173-664) over the CUDA path.

`radial_velocity(t)` keeps the reference's contract (numpy in -> numpy out; a CUDA tensor in
-> a CUDA tensor out) and its errors (ValueError from the constructor for invalid orbital
parameters), but the arithmetic runs in the sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .param import Parameterisation


def _is_tensor(x) -> bool:
    try:
        import torch
        return isinstance(x, torch.Tensor)
    except Exception:  # pragma: no cover
        return False


def _njit_kepler_rv(M, e: float, K: float, w: float):
    """model.py:173-213 — Kepler solve + RV for an array of mean anomalies (0 < e < 1)."""
    return _compute_rv(M, e, K, w)


def _compute_rv(M, e: float, K: float, w: float):
    """model.py:216-243 — dispatches the circular shortcut on e == 0 exactly."""
    torch = _lib._torch()
    is_t = _is_tensor(M)
    m = _lib.as_cuda_f64(M).reshape(-1)
    out = torch.empty_like(m)
    _lib.check(_lib.load().rvlp_kepler_rv(m.data_ptr(), m.numel(), float(e), float(K), float(w), out.data_ptr(),
                                          m.device.index, _lib.stream_ptr(m.device.index)))
    return out if is_t else out.cpu().numpy()


class Planet:
    """model.py:246-355."""

    def __init__(self, letter: str, parameterisation: Parameterisation, params: dict) -> None:
        if not (letter.isalpha() and (letter == letter[0] * len(letter))):
            raise ValueError(f"Letter {letter} is not a single alphabet character.")
        self.letter = letter
        self.parameterisation = parameterisation
        self.params = params
        if not set(params.keys()) == set(parameterisation.pars):
            raise ValueError(f"Parameterisation {parameterisation} does not match input params {params}")
        self._p5 = (C.c_double * 5)(*[float(params[p]) for p in parameterisation.pars])
        # validate now (model.py:270-275): zero-length evaluation runs conversion + validity only
        _lib._torch()
        _lib.check(_lib.load().rvlp_planet_rv(parameterisation.id, self._p5, None, 0, None, 0,
                                              _lib.current_device(), None))
        self._rvparams = None

    @property
    def rvparams(self) -> dict:
        if self._rvparams is None:
            self._rvparams = self.parameterisation.convert_pars_to_default_parameterisation(self.params)
        return self._rvparams

    def __repr__(self) -> str:
        return f"Planet(letter={self.letter!r}, parameterisation={self.parameterisation!r}, params={self.params!r})"

    def __str__(self) -> str:
        return f"Planet {self.letter} {self.params}"

    def _rv_into(self, t_dev, out_dev, accumulate: bool) -> None:
        _lib.check(_lib.load().rvlp_planet_rv(self.parameterisation.id, self._p5, t_dev.data_ptr(), t_dev.numel(),
                                              out_dev.data_ptr(), int(accumulate), t_dev.device.index,
                                              _lib.stream_ptr(t_dev.device.index)))

    def radial_velocity(self, t):
        """model.py:329-354."""
        torch = _lib._torch()
        is_t = _is_tensor(t)
        tt = _lib.as_cuda_f64(t).reshape(-1)
        out = torch.empty_like(tt)
        self._rv_into(tt, out, False)
        return out if is_t else out.cpu().numpy()


class Trend:
    """model.py:426-509."""

    def __init__(self, t0: float, params: dict) -> None:
        self.gammadot = params["gd"]
        self.gammadotdot = params["gdd"]
        try:
            self.t0 = float(t0)
        except (TypeError, ValueError) as e:
            raise ValueError(f"t0 must be a numeric value (recommend mean or median of observation times), "
                             f"but got {type(t0).__name__}: {t0}") from e

    def __repr__(self) -> str:
        return f"Trend(params={{'gd': {self.gammadot}, 'gdd': {self.gammadotdot}}}, t0={self.t0:.2f})"

    def _rv_into(self, t_dev, out_dev, accumulate: bool) -> None:
        _lib.check(_lib.load().rvlp_trend_rv(float(self.gammadot), float(self.gammadotdot), self.t0,
                                             t_dev.data_ptr(), t_dev.numel(), out_dev.data_ptr(), int(accumulate),
                                             t_dev.device.index, _lib.stream_ptr(t_dev.device.index)))

    def radial_velocity(self, t):
        torch = _lib._torch()
        is_t = _is_tensor(t)
        tt = _lib.as_cuda_f64(t).reshape(-1)
        out = torch.empty_like(tt)
        self._rv_into(tt, out, False)
        return out if is_t else out.cpu().numpy()


class Instrument:
    """model.py:356-424 (record only)."""

    def __init__(self, name: str, g: float, jit: float) -> None:
        if not isinstance(name, str) or len(name) == 0:
            raise ValueError(f"Instrument name must be a non-empty string, got: {name!r}")
        if jit < 0:
            raise ValueError(f"Jitter must be >= 0, got: {jit}")
        self.name, self.g, self.jit = name, g, jit


class Star:
    """model.py:512-664 — sum of planets (insertion order) + trend."""

    def __init__(self, name: str, mass: float) -> None:
        self.name = name
        self.mass = mass
        self.planets: dict[str, Planet] = {}
        self.instruments: dict[str, Instrument] = {}
        self.num_planets = 0
        if mass <= 0:
            raise ValueError(f"Stellar mass {self.mass} must be greater than zero")

    def add_planet(self, planet: Planet) -> None:
        if planet.letter in self.planets:
            import warnings
            warnings.warn(f"Planet {planet.letter} already exists and will be overwritten", UserWarning, stacklevel=2)
        self.planets[planet.letter] = planet
        self.num_planets = len(self.planets)

    def add_trend(self, trend: Trend) -> None:
        self.trend = trend

    def add_instrument(self, instrument: Instrument) -> None:
        self.instruments[instrument.name] = instrument

    def radial_velocity(self, t):
        """model.py:639-664."""
        torch = _lib._torch()
        is_t = _is_tensor(t)
        tt = _lib.as_cuda_f64(t).reshape(-1)
        out = torch.zeros_like(tt)
        for planet in self.planets.values():
            planet._rv_into(tt, out, True)
        self.trend._rv_into(tt, out, True)
        return out if is_t else out.cpu().numpy()
