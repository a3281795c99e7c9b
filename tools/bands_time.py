"""K6 timing: two-pass value-space path vs the radix path. python tools/bands_time.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import _lib, fit, workloads
spec, theta = workloads.make_c2(100_000)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda")
for T in (1000, 200):
    times = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), T, dtype=torch.float64, device="cuda")
    m = post.ctx.rv_matrix(th, times, -2)
    out = torch.empty((3, T), dtype=torch.float64, device="cuda")
    for mode in ("1", "0"):
        os.environ["RVLP_BANDS_FAST"] = mode
        for _ in range(3): _lib.percentile_columns(m, [15.85, 50, 84.15], out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): _lib.percentile_columns(m, [15.85, 50, 84.15], out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(f"S=1e5 T={T} fast={mode}: {ms:.3f} ms  {m.numel() * 8 / ms / 1e6:.0f} GB/s algorithmic", flush=True)
