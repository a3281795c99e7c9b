"""pipe vs batch GP paths over the epoch count (large batch): python tools/gp_crossover.py"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for N in (16, 30, 48, 64, 80, 96, 112, 120, 128, 144, 160, 176, 200, 219):
    row = {"N": N, "S": S}
    for k in ("pipe", "batch"):
        os.environ["RVLP_GP_KERNEL"] = k
        spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
        post = fit.from_spec(spec)
        th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
        times = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), N, dtype=torch.float64, device="cuda")
        for what, fn in (("logprob", lambda: post.ctx.logprob(th, out=out)), ("predict", lambda: post.ctx.gp_predict(th, times, want_chi2=True))):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3): fn()
            b.record(); torch.cuda.synchronize()
            row[f"{what}_{k}_ms"] = round(a.elapsed_time(b) / 3, 3)
    print(json.dumps(row), flush=True)
