"""ctypes wrapper of oracle/oracle.c (the C restatement; TEST INFRASTRUCTURE ONLY).

Builds `oracle/_build/liboracle.so` with gcc on first use.  Importable only from tests/,
`__graft_entry__` and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ravest_b200.descriptor import DescPOD, Descriptor, PriorPOD, instrument_indices, make_prior_pod

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "oracle.c")
    hdr = os.path.join(HERE, "..", "include", "ravest_b200.h")
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", LIB, src, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"gcc failed: {res.stderr}")
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_prior.restype = C.c_double
        _lib.oracle_prior.argtypes = [C.POINTER(PriorPOD), C.c_double]
    return _lib


def max_threads() -> int:
    return int(load().oracle_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleProblem:
    """The C oracle bound to one problem (same flat descriptor as the CUDA library)."""

    def __init__(self, spec: dict):
        self.desc = Descriptor.from_spec(spec)
        self.t = np.ascontiguousarray(spec["time"], dtype=np.float64)
        self.v = np.ascontiguousarray(spec["vel"], dtype=np.float64)
        self.e = np.ascontiguousarray(spec["velerr"], dtype=np.float64)
        inst = np.asarray(spec["instrument"])
        self.inst = instrument_indices(inst, np.unique(inst))
        self.lib = load()

    def _theta(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        if th.ndim == 1:
            th = th.reshape(1, -1)
        assert th.shape[1] == self.desc.ndim, (th.shape, self.desc.ndim)
        return th

    def logprob(self, theta, nthreads: int = 0) -> np.ndarray:
        th = self._theta(theta)
        out = np.empty(th.shape[0])
        fn = self.lib.oracle_gp_logprob_batch if self.desc.is_gp else self.lib.oracle_logprob_batch
        rc = fn(self.desc.byref(), _p(self.t), _p(self.v), _p(self.e), _p(self.inst), C.c_int64(len(self.t)),
                _p(th), C.c_int64(th.shape[0]), _p(out), C.c_int(nthreads))
        assert rc == 0
        return out

    def parts(self, theta, nthreads: int = 0):
        th = self._theta(theta)
        ll, lp = np.empty(th.shape[0]), np.empty(th.shape[0])
        rc = self.lib.oracle_logprob_parts_batch(self.desc.byref(), _p(self.t), _p(self.v), _p(self.e),
                                                 _p(self.inst), C.c_int64(len(self.t)), _p(th),
                                                 C.c_int64(th.shape[0]), _p(ll), _p(lp), C.c_int(nthreads))
        assert rc == 0
        return ll, lp

    def rv_matrix(self, theta, times, component: int, nthreads: int = 0) -> np.ndarray:
        th = self._theta(theta)
        tt = np.ascontiguousarray(times, dtype=np.float64)
        out = np.empty((th.shape[0], len(tt)))
        rc = self.lib.oracle_rv_batch(self.desc.byref(), _p(th), C.c_int64(th.shape[0]), _p(tt),
                                      C.c_int64(len(tt)), C.c_int32(component), _p(out), C.c_int(nthreads))
        assert rc == 0
        return out


def kepler_rv(M, e: float, K: float, w: float) -> np.ndarray:
    M = np.ascontiguousarray(M, dtype=np.float64)
    out = np.empty_like(M)
    load().oracle_kepler_rv(_p(M), C.c_int64(M.size), C.c_double(e), C.c_double(K), C.c_double(w), _p(out))
    return out


def convert_to_default(par_id: int, values) -> tuple[np.ndarray, np.ndarray]:
    x = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, 5)
    out = np.empty_like(x)
    valid = np.empty(x.shape[0], dtype=np.int32)
    load().oracle_convert_to_default(C.c_int(par_id), _p(x), C.c_int64(x.shape[0]), _p(out), _p(valid))
    return out, valid


def prior(prior_obj, x: float) -> float:
    pod = make_prior_pod(prior_obj)
    return float(load().oracle_prior(C.byref(pod), C.c_double(x)))
