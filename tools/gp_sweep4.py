"""batch vs smem at the large-N end: python tools/gp_sweep4.py"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
for S in (1024, 4096, 16384):
    for N in (152, 168, 176, 184, 200, 216, 232):
        row = {"S": S, "N": N}
        for k in ("batch", "smem"):
            os.environ["RVLP_GP_KERNEL"] = k
            spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
            post = fit.from_spec(spec)
            th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
            fn = lambda: post.ctx.logprob(th, out=out)
            for _ in range(2): fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): fn()
            b.record(); torch.cuda.synchronize()
            row[k] = round(a.elapsed_time(b) / 5, 4)
        print(json.dumps(row), flush=True)
