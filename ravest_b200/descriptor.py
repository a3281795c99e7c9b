"""Compile a LogPosterior's state into the flat POD descriptor of `include/ravest_b200.h`.

This is where every string / dict lookup of the reference's per-call path dies
(`fit.py:3399-3446`, `3465-3468`, `3616-3621`, `3642`, `3652`): parameter names become
column numbers or constants, priors become (kind, target, constants) rows in evaluation
order, the per-planet Jacobian / renormalisation cases become two scalars.
Pure host Python + ctypes; no CUDA needed to build a descriptor.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np

from . import prior as _prior

ABI_VERSION = 1

ALLOWED_PARAMETERISATIONS = ["P K e w Tp", "P K e w Tc", "P K secosw sesinw Tp", "P K secosw sesinw Tc"]
GP_HYPERPARAMS = ["gp_amp", "gp_lambda_e", "gp_lambda_p", "gp_period"]      # gp.py:37
TARGETS = {"P": 1, "K": 2, "e": 3, "w": 4, "Tp": 5}


class PriorPOD(C.Structure):
    _fields_ = [("kind", C.c_int32), ("target", C.c_int32), ("index", C.c_int32),
                ("is_hyper", C.c_int32), ("p", C.c_double * 4), ("c", C.c_double * 2)]


class DescPOD(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_planets", C.c_int32), ("parameterisation", C.c_int32),
                ("n_inst", C.c_int32), ("ndim", C.c_int32), ("n_priors", C.c_int32),
                ("n_hyper", C.c_int32), ("reserved", C.c_int32),
                ("t0", C.c_double), ("jacobian", C.c_double), ("renorm", C.c_double),
                ("src_col", C.POINTER(C.c_int32)), ("src_const", C.POINTER(C.c_double)),
                ("priors", C.POINTER(PriorPOD))]


def make_prior_pod(pr: "_prior._Prior", target: int = 0, index: int = 0, is_hyper: int = 0) -> PriorPOD:
    kind, p, c = pr.pod()
    pod = PriorPOD()
    pod.kind, pod.target, pod.index, pod.is_hyper = kind, target, index, is_hyper
    pod.p[:] = p
    pod.c[:] = c
    return pod


class Descriptor:
    """Owns the ctypes buffers behind an `rvlp_desc` and the name <-> column bookkeeping."""

    def __init__(
        self,
        planet_letters: Sequence[str],
        parameterisation: str,
        priors: Mapping[str, "_prior._Prior"],
        fixed_params: Mapping[str, float],
        free_params_names: Sequence[str],
        unique_instruments: Sequence[str],
        t0: float,
        hyperpriors: Mapping[str, "_prior._Prior"] | None = None,
        fixed_hyperparams: Mapping[str, float] | None = None,
        free_hyperparams_names: Sequence[str] | None = None,
        jacobian: float = 0.0,
        renorm: float = 0.0,
    ) -> None:
        if parameterisation not in ALLOWED_PARAMETERISATIONS:
            raise ValueError(f"parameterisation {parameterisation} not recognised. "
                             f"Must be one of {ALLOWED_PARAMETERISATIONS}")
        self.planet_letters = list(planet_letters)
        self.parameterisation = parameterisation
        self.pars = parameterisation.split()
        self.unique_instruments = [str(i) for i in unique_instruments]
        self.free_params_names = list(free_params_names)
        self.free_hyperparams_names = list(free_hyperparams_names or [])
        self.is_gp = hyperpriors is not None
        columns = self.free_params_names + self.free_hyperparams_names        # fit.py:4978
        self.columns = columns
        col_of = {n: i for i, n in enumerate(columns)}
        if len(col_of) != len(columns):
            raise ValueError("duplicate free parameter names")

        # model parameter numbering (include/ravest_b200.h)
        self.model_names = []
        for L in self.planet_letters:
            self.model_names += [f"{p}_{L}" for p in self.pars]
        self.model_names += ["gd", "gdd"]
        self.model_names += [f"g_{i}" for i in self.unique_instruments]       # fit.py:3591
        self.model_names += [f"jit_{i}" for i in self.unique_instruments]     # fit.py:3592
        names = list(self.model_names) + (GP_HYPERPARAMS if self.is_gp else [])
        fixed = dict(fixed_params)
        fixed.update(fixed_hyperparams or {})
        src_col = np.full(len(names), -1, dtype=np.int32)
        src_const = np.zeros(len(names), dtype=np.float64)
        for i, n in enumerate(names):
            if n in col_of:
                src_col[i] = col_of[n]
            elif n in fixed:
                src_const[i] = float(fixed[n])
            else:
                raise KeyError(f"parameter {n!r} is neither free nor fixed")
        self.src_col, self.src_const = src_col, src_const

        # priors in the order LogPrior.__call__ visits them (fit.py:3399-3446, 3685-3691)
        rows: list[PriorPOD] = []
        prior_keys = set(priors.keys())
        free_keys = set(self.free_params_names)
        if prior_keys == free_keys:
            for n in self.free_params_names:
                rows.append(make_prior_pod(priors[n], 0, col_of[n]))
        else:
            seen: dict[str, int] = {}
            for n in self.free_params_names:
                if n in prior_keys:
                    seen[n] = len(rows)
                    rows.append(make_prior_pod(priors[n], 0, col_of[n]))
            for k, L in enumerate(self.planet_letters):
                for dpar in ("P", "K", "e", "w", "Tp"):
                    key = f"{dpar}_{L}"
                    if key in prior_keys:
                        pod = make_prior_pod(priors[key], TARGETS[dpar], k)
                        if key in seen:
                            rows[seen[key]] = pod      # dict value overwritten, position kept
                        else:
                            seen[key] = len(rows)
                            rows.append(pod)
            missing = prior_keys - set(seen)
            if missing:
                # the reference would sum only the keys it builds; an unmatched prior is never evaluated
                pass
        for n in self.free_hyperparams_names:                                   # fit.py:7884
            if n in (hyperpriors or {}):
                rows.append(make_prior_pod(hyperpriors[n], 0, col_of[n], is_hyper=1))
        self._prior_array = (PriorPOD * max(1, len(rows)))(*rows)
        self.n_priors = len(rows)

        d = DescPOD()
        d.abi_version = ABI_VERSION
        d.n_planets = len(self.planet_letters)
        d.parameterisation = ALLOWED_PARAMETERISATIONS.index(parameterisation)
        d.n_inst = len(self.unique_instruments)
        d.ndim = len(columns)
        d.n_priors = self.n_priors
        d.n_hyper = 4 if self.is_gp else 0
        d.t0 = float(t0)
        d.jacobian = float(jacobian)
        d.renorm = float(renorm)
        d.src_col = src_col.ctypes.data_as(C.POINTER(C.c_int32))
        d.src_const = src_const.ctypes.data_as(C.POINTER(C.c_double))
        d.priors = C.cast(self._prior_array, C.POINTER(PriorPOD))
        self.pod = d
        self.ndim = len(columns)
        self.n_model = len(self.model_names)

    def byref(self):
        return C.byref(self.pod)

    # convenience for tests / oracle wrappers ------------------------------------------
    @classmethod
    def from_spec(cls, spec: dict, jacobian: float | None = None, renorm: float | None = None) -> "Descriptor":
        """Build from a workload/fixture spec (see ravest_b200/workloads.py)."""
        from .fit import compute_logprob_corrections
        params = spec["params"]
        free = [k for k, (_, fx) in params.items() if not fx]
        fixed = {k: v for k, (v, fx) in params.items() if fx}
        priors = {k: _prior.from_tuple(v) for k, v in spec["priors"].items()}
        uniq = np.unique(np.asarray(spec["instrument"]))
        kw = {}
        if "hyperparams" in spec:
            hp = spec["hyperparams"]
            kw = dict(hyperpriors={k: _prior.from_tuple(v) for k, v in spec["hyperpriors"].items()},
                      fixed_hyperparams={k: v for k, (v, fx) in hp.items() if fx},
                      free_hyperparams_names=[k for k, (_, fx) in hp.items() if not fx])
        if jacobian is None:
            jacobian, renorm, _ = compute_logprob_corrections(
                list(spec["planet_letters"]), spec["parameterisation"], priors, free)
        return cls(list(spec["planet_letters"]), spec["parameterisation"], priors, fixed, free,
                   uniq, spec["t0"], jacobian=jacobian, renorm=renorm, **kw)


def instrument_indices(instrument, unique_instruments) -> np.ndarray:
    """fit.py:3585-3586 — per-epoch integer instrument index."""
    idx = {str(inst): i for i, inst in enumerate(unique_instruments)}
    return np.array([idx[str(i)] for i in instrument], dtype=np.int32)
