"""python tools/gp_quick_time.py  - batch GP path timing at a few shapes (used with RVLP_LIB variants)"""
import os, sys, torch
sys.path.insert(0, ".")
os.environ["RVLP_GP_KERNEL"] = os.environ.get("RVLP_GP_KERNEL", "batch")
from ravest_b200 import fit, workloads
out = []
for N, S in ((120, 10000), (256, 4000)):
    spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda"); o = torch.empty(S, dtype=torch.float64, device="cuda")
    for _ in range(2): post.ctx.logprob(th, out=o)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): post.ctx.logprob(th, out=o)
    b.record(); torch.cuda.synchronize()
    out.append(f"N={N}: {a.elapsed_time(b) / 5:.3f} ms")
print(os.path.basename(os.environ.get("RVLP_LIB", "default")), "  ".join(out))
