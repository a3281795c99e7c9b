"""pipe / batch / smem GP log-prob paths over epochs and batch size: python tools/gp_sweep3.py"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
for S in (256, 1024, 4096, 16384):
    for N in (12, 16, 24, 30, 40, 48, 64, 80, 96, 120, 136, 144, 152, 160, 176, 200):
        row = {"S": S, "N": N}
        for k in ("pipe", "batch", "smem"):
            os.environ["RVLP_GP_KERNEL"] = k
            spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
            post = fit.from_spec(spec)
            th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
            fn = lambda: post.ctx.logprob(th, out=out)
            try:
                for _ in range(2): fn()
            except Exception as e:
                row[k] = None; continue
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): fn()
            b.record(); torch.cuda.synchronize()
            row[k] = round(a.elapsed_time(b) / 5, 4)
        print(json.dumps(row), flush=True)
