"""Time the UNMODIFIED reference's own log-probability on the host cores.

TEST / BASELINE INFRASTRUCTURE ONLY - imported by bench.py's `cpu_baseline` leg and `--impl reference` arm, never by
the product.  Runs `ravest.fit.LogPosterior.log_probability(dict)` (fit.py:3448-3495) exactly as emcee drives it with
`parameter_names=` (fit.py:1070-1075): one dict per row, (i) in this process on one core, (ii) mapped over a
`multiprocessing.get_context("spawn").Pool()` - the reference's own parallel mechanism (fit.py:1069-1072), which pickles
the bound method with every chunk.  The reference is imported from oracle/_ref (or /root/reference), see ref_import.py.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

from .ref_import import import_reference, reference_available


def available() -> tuple[bool, str]:
    if not reference_available():
        return False, "oracle/_ref (installed copy of the reference) not present"
    try:
        import numba  # noqa: F401
    except Exception as ex:        # the reference's Kepler solver is @njit
        return False, f"numba not importable: {ex!r}"
    return True, ""


def _init_worker():
    """Pool initializer: the reference sets OMP/MKL threads to 1 at import (fit.py:13-17); import it (with the
    stand-ins for its plotting / sampler imports) before any pickled LogPosterior arrives."""
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    import logging
    logging.disable(logging.CRITICAL)
    import_reference()


def build_logposterior(spec):
    model, param, prior, fit = import_reference()
    params = spec["params"]
    free = [k for k, (_, fx) in params.items() if not fx]
    fixed = {k: v for k, (v, fx) in params.items() if fx}
    inst = np.asarray(spec["instrument"])
    lp = fit.LogPosterior(list(spec["planet_letters"]), param.Parameterisation(spec["parameterisation"]),
                          {k: getattr(prior, p[0])(*p[1:]) for k, p in spec["priors"].items()}, fixed, free,
                          np.asarray(spec["time"], float), np.asarray(spec["vel"], float),
                          np.asarray(spec["velerr"], float), inst, np.unique(inst), spec["t0"])
    return lp, free


def rows_to_dicts(names, theta):
    return [dict(zip(names, (float(x) for x in row))) for row in theta]


class ReferenceRunner:
    """Owns the reference LogPosterior of one problem and (optionally) a spawn pool."""

    def __init__(self, spec, workers: int = 0):
        import logging
        logging.disable(logging.CRITICAL)          # the reference logs one line per planet at construction
        self.lp, self.names = build_logposterior(spec)
        logging.disable(logging.NOTSET)
        self.workers = workers
        self.pool = None
        if workers > 1:
            self.pool = mp.get_context("spawn").Pool(workers, initializer=_init_worker)     # fit.py:1069

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def evaluate(self, theta) -> np.ndarray:
        """All rows through the reference; the pool path is `pool.map(lp.log_probability, dicts)` as emcee does."""
        dicts = rows_to_dicts(self.names, theta)
        if self.pool is not None:
            return np.asarray(self.pool.map(self.lp.log_probability, dicts), dtype=np.float64)
        return np.asarray([self.lp.log_probability(d) for d in dicts], dtype=np.float64)

    def evaluate_serial(self, theta) -> np.ndarray:
        return np.asarray([self.lp.log_probability(d) for d in rows_to_dicts(self.names, theta)], dtype=np.float64)

    def time_serial(self, theta, seconds: float = 3.0) -> tuple[float, int]:
        """(seconds per log-probability on ONE core, rows used): numba JIT warmed on the first rows."""
        dicts = rows_to_dicts(self.names, theta)
        for d in dicts[:3]:
            self.lp.log_probability(d)
        n = 0
        t0 = time.perf_counter()
        while n < len(dicts):
            self.lp.log_probability(dicts[n])
            n += 1
            if n >= 20 and time.perf_counter() - t0 > seconds:
                break
        return (time.perf_counter() - t0) / n, n
