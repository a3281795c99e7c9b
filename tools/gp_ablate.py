"""Timing of the ablation builds of K3 (results are wrong by design; only the time matters).
usage (GPU box): for k in 1 2 3 4; do RVLP_LIB=$PWD/build_variants/abl$k.so python tools/gp_ablate.py; done"""
import os, sys, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
for N, S in ((120, 10000),):
    spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
    line = [os.path.basename(os.environ.get("RVLP_LIB", "product"))]
    for cap in ("0", "148"):
        os.environ["RVLP_GP_GRID"] = cap
        for _ in range(2): post.ctx.logprob(th, out=out)
        torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): post.ctx.logprob(th, out=out)
        b.record(); torch.cuda.synchronize()
        line.append(f"cap {cap}: {a.elapsed_time(b) / 5:.3f} ms")
    print("  ".join(line), flush=True)
