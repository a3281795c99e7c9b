import sys, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import _lib, fit, workloads
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
spec, theta = workloads.make_c2(S)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda")
times = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), 1000, dtype=torch.float64, device="cuda")
m = post.ctx.rv_matrix(th, times, -2)
out = torch.empty((3, 1000), dtype=torch.float64, device="cuda")
for _ in range(2):
    _lib.percentile_columns(m, [15.85, 50, 84.15], out=out)
torch.cuda.synchronize()
