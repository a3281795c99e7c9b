"""Latency of ONE sample per CTA in the shared-memory GP kernel: S = 148 (one CTA per SM alone), 296, 444 (three per SM).
python tools/gp_smem_lat.py [N]"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, ".")
os.environ["RVLP_GP_KERNEL"] = sys.argv[2] if len(sys.argv) > 2 else "smem"
from ravest_b200 import fit, workloads
N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
for S in (148, 296, 444, 888, 1332, 4440, 10000):
    spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
    theta[:, -1] = np.abs(theta[:, -1])      # no rejected rows
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
    fn = lambda: post.ctx.logprob(th, out=out)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(json.dumps({"N": N, "S": S, "ms": round(ms, 4), "us_per_sample_per_SM": round(ms * 1e3 / (S / 148), 2)}), flush=True)
