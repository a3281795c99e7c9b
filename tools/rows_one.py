"""One call of each f-row kernel for ncu captures: python tools/rows_one.py"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import _lib, fit, workloads
spec, theta = workloads.make_c2(100_000)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda")
T = 1000
times = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), T, dtype=torch.float64, device="cuda")
m = torch.empty((len(theta), T), dtype=torch.float64, device="cuda")
post.ctx.rv_matrix(th, times, -2, out=m)
out = torch.empty((3, T), dtype=torch.float64, device="cuda")
_lib.percentile_columns(m, [15.85, 50, 84.15], out=out)
del m
spec3, theta3 = workloads.make_c3(1_000_000)
post3 = fit.from_spec(spec3)
th3 = torch.as_tensor(theta3, device="cuda")
post3.ctx.walker_check(th3)
post3.information_criteria_batch(th3)
del th3
spec5, theta5 = workloads.make_c5(10_000)
post5 = fit.from_spec(spec5)
th5 = torch.as_tensor(theta5, device="cuda")
t5 = torch.linspace(float(spec5["time"].min()), float(spec5["time"].max()), T, dtype=torch.float64, device="cuda")
post5.ctx.gp_predict(th5, t5)
torch.cuda.synchronize(); print("ok")
