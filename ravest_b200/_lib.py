"""ctypes binding of `csrc/libravest_b200.so` (the C ABI of include/ravest_b200.h).

PyTorch is used only for device memory, streams and torch.distributed: every compute call
goes `torch.Tensor.data_ptr()` -> C ABI -> hand-written sm_100a kernels.  There is NO CPU
fallback: if the shared library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

from .descriptor import DescPOD, PriorPOD, make_prior_pod

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# RVLP_LIB lets tools/kernel_sweep.py time alternative builds of the same sources; it is not a fallback
LIB_PATH = os.environ.get("RVLP_LIB") or os.path.join(CSRC, "libravest_b200.so")
SOURCES = ["rvlp_capi.cu"]
HEADERS = ["rvlp_math.cuh", "rvlp_kernels.cuh", "rvlp_gp.cuh", "rvlp_gp_batch.cuh", "rvlp_gp_smem.cuh", "rvlp_gpcov.cuh", "rvlp_gp_pipe.cuh", "rvlp_bands.cuh", "rvlp_bands_fast.cuh", os.path.join("..", "..", "include", "ravest_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

EXPORTS = [
    "rvlp_abi_version", "rvlp_last_error", "rvlp_ctx_create", "rvlp_ctx_destroy", "rvlp_logprob_batch",
    "rvlp_logprob_batch_host", "rvlp_logprob_parts_batch", "rvlp_rv_batch", "rvlp_gp_logprob_batch",
    "rvlp_kepler_rv", "rvlp_planet_rv", "rvlp_trend_rv", "rvlp_convert_to_default", "rvlp_prior_eval",
    "rvlp_measure_fp64_peak", "rvlp_launch_count", "rvlp_rv_batch_frozen", "rvlp_walker_check_batch",
    "rvlp_gp_predict_batch", "rvlp_percentile_workspace_bytes", "rvlp_percentile_columns", "rvlp_ctx_autotune", "rvlp_ctx_set_variant",
    "rvlp_info_criteria_batch", "rvlp_logprob_batch_peers", "rvlp_peer_barrier", "rvlp_peer_alloc", "rvlp_peer_open",
    "rvlp_peer_close", "rvlp_peer_free",
]
MAX_FROZEN = 16
MAX_PERCENTILES = 8
# rvlp_walker_check_batch status bits (include/ravest_b200.h)
WALKER_NONFINITE, WALKER_PLANET, WALKER_JITTER, WALKER_PRIOR, WALKER_HYPER, WALKER_HYPERPRIOR = 1, 2, 4, 8, 16, 32


class RvlpError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(d) for d in deps)
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + srcs
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RvlpError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise loudly if it is not there (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RvlpError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback).")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    lib.rvlp_abi_version.restype = C.c_int
    lib.rvlp_last_error.restype = C.c_char_p
    lib.rvlp_launch_count.restype = i64
    lib.rvlp_ctx_create.argtypes = [C.POINTER(DescPOD), vp, vp, vp, vp, i64, C.c_int, C.POINTER(vp)]
    lib.rvlp_ctx_destroy.argtypes = [vp]
    lib.rvlp_ctx_destroy.restype = None
    lib.rvlp_logprob_batch.argtypes = [vp, vp, i64, vp, vp]
    lib.rvlp_logprob_batch_host.argtypes = [vp, vp, i64, vp]
    lib.rvlp_logprob_batch_peers.argtypes = [vp, vp, i64, vp, i32, i64, vp]
    lib.rvlp_peer_barrier.argtypes = [C.c_int, vp, i32, i32, C.c_uint64, i64, vp]
    lib.rvlp_peer_alloc.argtypes = [C.c_int, i64, C.POINTER(vp), vp]
    lib.rvlp_peer_open.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.rvlp_peer_close.argtypes = [C.c_int, vp]
    lib.rvlp_peer_free.argtypes = [C.c_int, vp]
    lib.rvlp_logprob_parts_batch.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.rvlp_rv_batch.argtypes = [vp, vp, i64, vp, i64, i32, vp, vp]
    lib.rvlp_gp_logprob_batch.argtypes = [vp, vp, i64, vp, vp]
    lib.rvlp_kepler_rv.argtypes = [vp, i64, dbl, dbl, dbl, vp, C.c_int, vp]
    lib.rvlp_planet_rv.argtypes = [i32, vp, vp, i64, vp, C.c_int, C.c_int, vp]
    lib.rvlp_trend_rv.argtypes = [dbl, dbl, dbl, vp, i64, vp, C.c_int, C.c_int, vp]
    lib.rvlp_convert_to_default.argtypes = [i32, vp, i64, vp, vp, C.c_int, vp]
    lib.rvlp_prior_eval.argtypes = [C.POINTER(PriorPOD), vp, i64, vp, C.c_int, vp]
    lib.rvlp_measure_fp64_peak.argtypes = [C.c_int, C.c_int, C.POINTER(dbl), C.POINTER(dbl)]
    lib.rvlp_ctx_autotune.argtypes = [vp, vp, i64, vp, C.POINTER(i32)]
    lib.rvlp_ctx_set_variant.argtypes = [vp, i32]
    lib.rvlp_info_criteria_batch.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, vp]
    lib.rvlp_rv_batch_frozen.argtypes = [vp, vp, i64, vp, i64, i32, i32, vp, vp, vp, vp]
    lib.rvlp_walker_check_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.rvlp_gp_predict_batch.argtypes = [vp, vp, i64, vp, i64, vp, vp, vp]
    lib.rvlp_percentile_workspace_bytes.argtypes = [i64, i32]
    lib.rvlp_percentile_workspace_bytes.restype = i64
    lib.rvlp_percentile_columns.argtypes = [vp, i64, i64, vp, i32, vp, vp, i64, C.c_int, vp]
    for name in EXPORTS:
        if name not in ("rvlp_last_error", "rvlp_launch_count", "rvlp_ctx_destroy",
                        "rvlp_percentile_workspace_bytes"):
            getattr(lib, name).restype = C.c_int
    lib.rvlp_launch_count.restype = i64
    if lib.rvlp_abi_version() != 1:
        raise RvlpError("libravest_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().rvlp_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise RvlpError(f"rvlp error {rc}: {msg}")


def launch_count() -> int:
    return int(load().rvlp_launch_count())


# ------------------------------------------------------------------ torch plumbing
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RvlpError("no CUDA device visible: ravest_b200 has no CPU fallback")
    return torch


def current_device() -> int:
    return int(_torch().cuda.current_device())


def stream_ptr(device: int | None = None) -> int:
    torch = _torch()
    return int(torch.cuda.current_stream(device).cuda_stream)


def as_cuda_f64(x, device: int | None = None):
    """numpy / list / tensor -> contiguous float64 CUDA tensor (no copy if already one)."""
    torch = _torch()
    dev = torch.device("cuda", current_device() if device is None else device)
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(dev)


def prior_eval(prior, values):
    """prior.py callables, batched on the device; returns numpy for numpy input, tensor for tensor."""
    torch = _torch()
    lib = load()
    is_tensor = isinstance(values, torch.Tensor)
    x = as_cuda_f64(values).reshape(-1)
    out = torch.empty_like(x)
    pod = make_prior_pod(prior)
    check(lib.rvlp_prior_eval(C.byref(pod), x.data_ptr(), x.numel(), out.data_ptr(), x.device.index,
                              stream_ptr(x.device.index)))
    return out if is_tensor else out.cpu().numpy()


def percentile_columns(matrix, q, out=None):
    """`np.percentile(matrix, q, axis=0)` on the device (fit.py:2239-2240, 2493-2495): matrix [S, T] CUDA
    fp64 tensor (or array-like, copied to the device) -> [len(q), T]; tensor in, tensor out."""
    torch = _torch()
    lib = load()
    is_tensor = isinstance(matrix, torch.Tensor)
    A = as_cuda_f64(matrix)
    if A.dim() != 2:
        raise ValueError(f"matrix must be 2-D (samples x times), got shape {tuple(A.shape)}")
    scalar_q = np.ndim(q) == 0
    qh = np.ascontiguousarray(np.atleast_1d(q), dtype=np.float64)
    if qh.ndim != 1 or not 1 <= len(qh) <= MAX_PERCENTILES:
        raise ValueError(f"q must hold 1..{MAX_PERCENTILES} percentiles")
    S, T = A.shape
    dev = A.device.index
    if out is None:
        out = torch.empty((len(qh), T), dtype=torch.float64, device=A.device)
    nbytes = lib.rvlp_percentile_workspace_bytes(T, len(qh))
    ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=A.device)
    check(lib.rvlp_percentile_columns(A.data_ptr(), S, T, qh.ctypes.data, len(qh), out.data_ptr(), ws.data_ptr(),
                                      ws.numel(), dev, stream_ptr(dev)))
    res = out[0] if scalar_q else out
    return res if is_tensor else res.cpu().numpy()


def measure_fp64_peak(device: int = 0, iters: int = 4096) -> tuple[float, float]:
    lib = load()
    _torch()
    f, ms = C.c_double(), C.c_double()
    check(lib.rvlp_measure_fp64_peak(device, iters, C.byref(f), C.byref(ms)))
    return f.value, ms.value


class Context:
    """RAII wrapper of rvlp_ctx: resident epoch arrays + descriptor tables on one device."""

    def __init__(self, desc, time, vel, velerr, inst_idx, device: int | None = None):
        lib = load()
        _torch()
        self.desc = desc
        self.device = current_device() if device is None else int(device)
        t = np.ascontiguousarray(time, dtype=np.float64)
        v = np.ascontiguousarray(vel, dtype=np.float64)
        e = np.ascontiguousarray(velerr, dtype=np.float64)
        ii = np.ascontiguousarray(inst_idx, dtype=np.int32)
        if not (len(t) == len(v) == len(e) == len(ii)):
            raise ValueError("Time, velocity, uncertainty, and instrument arrays must be the same length.")
        self.n_epochs = len(t)
        h = C.c_void_p()
        check(lib.rvlp_ctx_create(desc.byref(), t.ctypes.data, v.ctypes.data, e.ctypes.data, ii.ctypes.data,
                                  len(t), self.device, C.byref(h)))
        self._h = h
        self._lib = lib
        self.k1_variant = None      # set by autotune()

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rvlp_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- device-tensor entry points --------------------------------------------------
    def _theta(self, theta):
        th = as_cuda_f64(theta, self.device)
        if th.dim() == 1:
            th = th.reshape(1, -1)
        if th.dim() != 2 or th.shape[1] != self.desc.ndim:
            raise ValueError(f"theta must have shape (S, {self.desc.ndim}), got {tuple(th.shape)}")
        return th

    AUTOTUNE_MIN_ROWS = 1 << 15

    def autotune(self, theta) -> int:
        """Pick the fastest compiled shape of the log-probability kernel for this problem (synchronous, once)."""
        th = self._theta(theta)
        chosen = C.c_int32(0)
        check(self._lib.rvlp_ctx_autotune(self._h, th.data_ptr(), th.shape[0], stream_ptr(self.device), C.byref(chosen)))
        self.k1_variant = int(chosen.value)
        return self.k1_variant

    def set_variant(self, variant: int) -> None:
        check(self._lib.rvlp_ctx_set_variant(self._h, int(variant)))
        self.k1_variant = int(variant)

    def logprob(self, theta, out=None):
        torch = _torch()
        th = self._theta(theta)
        if (self.k1_variant is None and th.shape[0] >= self.AUTOTUNE_MIN_ROWS and not self.desc.is_gp
                and os.environ.get("RVLP_AUTOTUNE", "1") != "0"):
            self.autotune(th)
        if out is None:
            out = torch.empty(th.shape[0], dtype=torch.float64, device=th.device)
        fn = self._lib.rvlp_gp_logprob_batch if self.desc.is_gp else self._lib.rvlp_logprob_batch
        check(fn(self._h, th.data_ptr(), th.shape[0], out.data_ptr(), stream_ptr(self.device)))
        return out

    def logprob_parts(self, theta):
        torch = _torch()
        th = self._theta(theta)
        ll = torch.empty(th.shape[0], dtype=torch.float64, device=th.device)
        lp = torch.empty_like(ll)
        check(self._lib.rvlp_logprob_parts_batch(self._h, th.data_ptr(), th.shape[0], ll.data_ptr(),
                                                 lp.data_ptr(), stream_ptr(self.device)))
        return ll, lp

    def info_criteria(self, theta, k_free: int = -1):
        """(loglike, chi2, aicc, bic) tensors [S] - rvlp_info_criteria_batch (fit.py:1361-1554)."""
        torch = _torch()
        th = self._theta(theta)
        out = [torch.empty(th.shape[0], dtype=torch.float64, device=th.device) for _ in range(4)]
        rc = self._lib.rvlp_info_criteria_batch(self._h, th.data_ptr(), th.shape[0], int(k_free), out[0].data_ptr(),
                                                out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
                                                stream_ptr(self.device))
        if rc == -1 and b"division by zero" in self._lib.rvlp_last_error():
            raise ZeroDivisionError("division by zero")            # (2k^2 + 2k) / (n - k - 1), fit.py:1528
        check(rc)
        return tuple(out)

    def logprob_host(self, theta_np: np.ndarray, out_np: np.ndarray | None = None) -> np.ndarray:
        """NumPy in, NumPy out through rvlp_logprob_batch_host (H2D + kernel + D2H inside)."""
        th = np.ascontiguousarray(theta_np, dtype=np.float64)
        if th.ndim == 1:
            th = th.reshape(1, -1)
        if th.shape[1] != self.desc.ndim:
            raise ValueError(f"theta must have shape (S, {self.desc.ndim}), got {th.shape}")
        if out_np is None:
            out_np = np.empty(th.shape[0], dtype=np.float64)
        if (self.k1_variant is None and th.shape[0] >= self.AUTOTUNE_MIN_ROWS and not self.desc.is_gp
                and os.environ.get("RVLP_AUTOTUNE", "1") != "0"):
            self.autotune(th[:1 << 18])
        check(self._lib.rvlp_logprob_batch_host(self._h, th.ctypes.data, th.shape[0], out_np.ctypes.data))
        return out_np

    def rv_matrix(self, theta, times, component: int, out=None, frozen: dict | None = None):
        """[S, T] RV of `component` (planet number, RV_TREND = -1, RV_TOTAL = -2) for every row of theta.
        `frozen` maps model-parameter NAMES to values that override every row (fit.py:2726-2751)."""
        torch = _torch()
        th = self._theta(theta)
        tt = as_cuda_f64(times, self.device).reshape(-1)
        if out is None:
            out = torch.empty((th.shape[0], tt.numel()), dtype=torch.float64, device=th.device)
        if frozen:
            idx = np.array([self.desc.model_names.index(k) for k in frozen], dtype=np.int32)
            val = np.array([float(v) for v in frozen.values()], dtype=np.float64)
            check(self._lib.rvlp_rv_batch_frozen(self._h, th.data_ptr(), th.shape[0], tt.data_ptr(), tt.numel(),
                                                 int(component), len(idx), idx.ctypes.data, val.ctypes.data,
                                                 out.data_ptr(), stream_ptr(self.device)))
        else:
            check(self._lib.rvlp_rv_batch(self._h, th.data_ptr(), th.shape[0], tt.data_ptr(), tt.numel(),
                                          int(component), out.data_ptr(), stream_ptr(self.device)))
        return out

    def walker_check(self, theta):
        """(status int32[S], log_prior[S], log_hyperprior[S]) - rvlp_walker_check_batch; status 0 = usable row."""
        torch = _torch()
        th = self._theta(theta)
        st = torch.empty(th.shape[0], dtype=torch.int32, device=th.device)
        lp = torch.empty(th.shape[0], dtype=torch.float64, device=th.device)
        lhp = torch.empty_like(lp)
        check(self._lib.rvlp_walker_check_batch(self._h, th.data_ptr(), th.shape[0], st.data_ptr(), lp.data_ptr(),
                                                lhp.data_ptr(), stream_ptr(self.device)))
        return st, lp, lhp

    def gp_predict(self, theta, times, want_chi2: bool = False):
        """GP conditional mean [S, T] at `times` (and alpha.alpha [S]) for every row - rvlp_gp_predict_batch."""
        torch = _torch()
        th = self._theta(theta)
        tt = as_cuda_f64(times, self.device).reshape(-1)
        mean = torch.empty((th.shape[0], tt.numel()), dtype=torch.float64, device=th.device)
        chi2 = torch.empty(th.shape[0], dtype=torch.float64, device=th.device) if want_chi2 else None
        check(self._lib.rvlp_gp_predict_batch(self._h, th.data_ptr(), th.shape[0], tt.data_ptr(), tt.numel(),
                                              mean.data_ptr(), chi2.data_ptr() if want_chi2 else None,
                                              stream_ptr(self.device)))
        return (mean, chi2) if want_chi2 else mean
