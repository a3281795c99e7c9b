"""K6 once per mode for an ncu launch list. python tools/bands_one.py [S] [T]"""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import _lib, fit, workloads
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
spec, theta = workloads.make_c2(S)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda")
times = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), T, dtype=torch.float64, device="cuda")
m = post.ctx.rv_matrix(th, times, -2)
out = torch.empty((3, T), dtype=torch.float64, device="cuda")
for _ in range(3): _lib.percentile_columns(m, [15.85, 50, 84.15], out=out)
torch.cuda.synchronize()
