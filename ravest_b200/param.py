"""Parameterisation / Parameter — host mirror of `ravest.param` (param.py:13-435, 597-625).

Same names, arguments and error behaviour as the reference for the slice the hot path
needs.  Conversions run on the GPU (`rvlp_convert_to_default`); there is no CPU arithmetic
here beyond argument checks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

ALLOWED_PARAMETERISATIONS = ["P K e w Tp", "P K e w Tc", "P K secosw sesinw Tp", "P K secosw sesinw Tc"]


class Parameter:
    """param.py:597-625 — value, unit (display only), fixed flag."""

    def __init__(self, value: float, unit: str, fixed: bool = False) -> None:
        self.value = value
        self.unit = unit
        self.fixed = fixed

    def __repr__(self) -> str:
        return f"Parameter(value={self.value!r}, unit={self.unit!r}, fixed={self.fixed!r})"

    def __str__(self) -> str:
        return f"Parameter {self.value} {self.unit}"


class Parameterisation:
    """param.py:13-435."""

    def __init__(self, parameterisation: str) -> None:
        if parameterisation not in ALLOWED_PARAMETERISATIONS:
            raise ValueError(f"parameterisation {parameterisation} not recognised. "
                             f"Must be one of {ALLOWED_PARAMETERISATIONS}")
        self.parameterisation = parameterisation
        self.pars = parameterisation.split()
        self.id = ALLOWED_PARAMETERISATIONS.index(parameterisation)

    def __str__(self) -> str:
        return f"Parameterisation: {self.parameterisation}"

    def __repr__(self) -> str:
        return f"Parameterisation({self.parameterisation})"

    def log_jacobian_determinant(self) -> float:
        """param.py:428-435 — ln 2 for the (secosw, sesinw) parameterisations, else 0."""
        if "secosw" in self.parameterisation:
            return float(np.log(2))
        return 0.0

    # ------------------------------------------------------------------ conversions (GPU)
    def convert_batch(self, values):
        """[n, 5] values in `pars` order -> ([n, 5] P K e w Tp, valid[n]) as CUDA tensors."""
        from . import _lib
        torch = _lib._torch()
        x = _lib.as_cuda_f64(values).reshape(-1, 5)
        out = torch.empty_like(x)
        valid = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
        _lib.check(_lib.load().rvlp_convert_to_default(
            self.id, x.data_ptr(), x.shape[0], out.data_ptr(), valid.data_ptr(), x.device.index,
            _lib.stream_ptr(x.device.index)))
        return out, valid

    def convert_pars_to_default_parameterisation(self, inpars: dict) -> dict:
        """param.py:299-362.  Raises ValueError for e outside [0, 1) when a Tc conversion is needed
        (param.py:209), exactly where the reference does."""
        row = [float(inpars[p]) for p in self.pars]
        out, _ = self.convert_batch([row])
        P, K, e, w, tp = (float(v) for v in out[0].cpu())
        if "Tc" in self.pars and (e < 0 or e >= 1.0):
            raise ValueError(f"Invalid eccentricity: {e} < 0" if e < 0 else f"Invalid eccentricity: {e} >= 1.0")
        return {"P": P, "K": K, "e": e, "w": w, "Tp": tp}

    def validate_default_parameterisation_params(self, params_dict: dict) -> None:
        """param.py:88-105."""
        P, K, e, w = (params_dict[k] for k in ("P", "K", "e", "w"))
        if P <= 0:
            raise ValueError(f"Invalid period: {P} <= 0")
        if K <= 0:
            raise ValueError(f"Invalid semi-amplitude: {K} <= 0")
        if e < 0:
            raise ValueError(f"Invalid eccentricity: {e} < 0")
        if e >= 1.0:
            raise ValueError(f"Invalid eccentricity: {e} >= 1.0")
        if not -np.pi <= w < np.pi:
            raise ValueError(f"Invalid argument of periastron: {w} not in [-pi, +pi)")

    def validate_planetary_params(self, params_dict: dict) -> None:
        """param.py:108-126."""
        if self.parameterisation != "P K e w Tp":
            params_dict = self.convert_pars_to_default_parameterisation(params_dict)
        self.validate_default_parameterisation_params(params_dict)
