"""Timings of the f-row kernels (RV matrix, percentile bands, walker check, GP conditioning) on one GPU."""
import json, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from ravest_b200 import _lib, fit, workloads


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {}
spec, theta = workloads.make_c2(100_000)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda")
times = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), 1000, dtype=torch.float64, device="cuda")
S, T = len(theta), 1000
m = torch.empty((S, T), dtype=torch.float64, device="cuda")
ms = timeit(lambda: post.ctx.rv_matrix(th, times, -2, out=m))
res["rv_matrix_total c2 1e5x1000"] = {"ms": ms, "evals_per_s": S * T * 2 / ms * 1e3, "write_GBs": S * T * 8 / ms / 1e6}
out = torch.empty((3, T), dtype=torch.float64, device="cuda")
ms = timeit(lambda: _lib.percentile_columns(m, [15.85, 50, 84.15], out=out))
res["percentile_bands 1e5x1000"] = {"ms": ms, "passes": 8, "read_GBs": 8 * S * T * 8 / ms / 1e6, "matrix_MB": S * T * 8 / 1e6}
for S2 in (10_000, 1_000):
    m2 = m[:S2].contiguous()
    ms = timeit(lambda: _lib.percentile_columns(m2, [15.85, 50, 84.15], out=out))
    res[f"percentile_bands {S2}x1000"] = {"ms": ms, "read_GBs": 8 * S2 * T * 8 / ms / 1e6}
if len(sys.argv) > 1 and sys.argv[1] == "bands":
    for k, v in res.items():
        print(k, json.dumps(v))
    sys.exit(0)
spec3, theta3 = workloads.make_c3(1_000_000)
post3 = fit.from_spec(spec3)
th3 = torch.as_tensor(theta3, device="cuda")
ms = timeit(lambda: post3.ctx.walker_check(th3))
res["walker_check c3 1e6 rows"] = {"ms": ms, "rows_per_s": 1e6 / ms * 1e3, "read_GBs": theta3.nbytes / ms / 1e6}
spec5, theta5 = workloads.make_c5(10_000)
post5 = fit.from_spec(spec5)
th5 = torch.as_tensor(theta5, device="cuda")
t5 = torch.linspace(float(spec5["time"].min()), float(spec5["time"].max()), 1000, dtype=torch.float64, device="cuda")
ms = timeit(lambda: post5.ctx.gp_predict(th5, t5), n=3, warm=1)
N = len(spec5["time"])
res["gp_predict c5 1e4 x N=120 x T=1000"] = {"ms": ms, "samples_per_s": 1e4 / ms * 1e3, "kernel_evals_per_s": 1e4 * 1000 * N / ms * 1e3}
ms = timeit(lambda: post5.ctx.gp_predict(th5, spec5["time"]), n=3, warm=1)
res["gp_predict c5 1e4 x N=120 x T=120"] = {"ms": ms, "samples_per_s": 1e4 / ms * 1e3}
ms = timeit(lambda: post5.ctx.logprob(th5), n=3, warm=1)
res["gp_logprob c5 1e4"] = {"ms": ms, "samples_per_s": 1e4 / ms * 1e3}
for k, v in res.items():
    print(k, json.dumps(v))
