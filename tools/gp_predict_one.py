"""One GP conditioning call at the bench shape (1e4 samples x 120 epochs -> 1000 test times) for ncu captures."""
import sys, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
spec, theta = workloads.make_c5(n_samples=10000)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda")
t = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), 1000, dtype=torch.float64, device="cuda")
for _ in range(3): m = post.ctx.gp_predict(th, t)
torch.cuda.synchronize(); print("ok", float(m[5, 7]))
