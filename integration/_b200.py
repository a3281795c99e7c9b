# src/ravest/_b200.py  (new file in ravest) - the ctypes stub of INTEGRATION.md §2, kept here verbatim so that
# tests/test_adapter.py and tests/test_gpu_dropin.py execute exactly the text the document shows.
import ctypes as C, numpy as np, os

from ravest_b200.adapter import compile_descriptor          # ravest.fit.LogPosterior -> rvlp_desc

_lib = C.CDLL(os.environ.get("RVLP_LIB", "libravest_b200.so"))   # fails loudly if absent: no silent fallback

class _Prior(C.Structure):                   # include/ravest_b200.h: rvlp_prior
    _fields_ = [("kind", C.c_int32), ("target", C.c_int32), ("index", C.c_int32), ("is_hyper", C.c_int32),
                ("p", C.c_double * 4), ("c", C.c_double * 2)]

class _Desc(C.Structure):                    # include/ravest_b200.h: rvlp_desc
    _fields_ = [("abi_version", C.c_int32), ("n_planets", C.c_int32), ("parameterisation", C.c_int32),
                ("n_inst", C.c_int32), ("ndim", C.c_int32), ("n_priors", C.c_int32), ("n_hyper", C.c_int32),
                ("reserved", C.c_int32), ("t0", C.c_double), ("jacobian", C.c_double), ("renorm", C.c_double),
                ("src_col", C.POINTER(C.c_int32)), ("src_const", C.POINTER(C.c_double)),
                ("priors", C.POINTER(_Prior))]

_lib.rvlp_ctx_create.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int64, C.c_int, C.POINTER(C.c_void_p)]
_lib.rvlp_logprob_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
_lib.rvlp_ctx_destroy.argtypes = [C.c_void_p]
_lib.rvlp_last_error.restype = C.c_char_p

class BatchedLogPosterior:
    """log_prob_fn for emcee.EnsembleSampler(..., vectorize=True): coords[n, ndim] -> n log-probs."""
    def __init__(self, lp, device=0):        # lp: ravest.fit.LogPosterior, already constructed
        self.desc, self._keep = compile_descriptor(lp)
        assert C.sizeof(self.desc) == C.sizeof(_Desc)
        t, v, e = (np.ascontiguousarray(x, dtype=np.float64) for x in (lp.time, lp.vel, lp.velerr))
        idx = np.ascontiguousarray(lp.log_likelihood._instrument_indices, dtype=np.int32)   # fit.py:3585-3586
        h = C.c_void_p()
        rc = _lib.rvlp_ctx_create(C.byref(self.desc), t.ctypes.data, v.ctypes.data, e.ctypes.data,
                                  idx.ctypes.data, len(t), device, C.byref(h))
        if rc: raise RuntimeError(_lib.rvlp_last_error().decode())
        self.h = h
    def __call__(self, coords):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        out = np.empty(len(coords))
        rc = _lib.rvlp_logprob_batch_host(self.h, coords.ctypes.data, len(coords), out.ctypes.data)
        if rc: raise RuntimeError(_lib.rvlp_last_error().decode())
        return out
    def __del__(self):
        if getattr(self, "h", None): _lib.rvlp_ctx_destroy(self.h); self.h = None
