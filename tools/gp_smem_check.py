"""Shared-memory DMMA GP kernel vs the pipelined / batched ones: agreement and time.  python tools/gp_smem_check.py [S]"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
S = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
Ns = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [7, 8, 16, 30, 57, 64, 100, 120, 128, 136]
for N in Ns:
    row = {"N": N, "S": S}
    res = {}
    for k in ("pipe", "batch", "smem"):
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
        row[f"{k}_ms"] = round(a.elapsed_time(b) / 5, 3)
        res[k] = out.cpu().numpy().copy()
    fin = np.isfinite(res["pipe"])
    row["inf_pattern_equal"] = bool(np.array_equal(np.isneginf(res["pipe"]), np.isneginf(res["smem"])) and np.array_equal(np.isnan(res["pipe"]), np.isnan(res["smem"])))
    row["max_abs_d_smem_pipe"] = float(np.max(np.abs(res["smem"][fin] - res["pipe"][fin]))) if fin.any() else None
    row["max_abs_d_batch_pipe"] = float(np.max(np.abs(res["batch"][fin] - res["pipe"][fin]))) if fin.any() else None
    row["max_abs_logp"] = float(np.max(np.abs(res["pipe"][fin]))) if fin.any() else None
    print(json.dumps(row), flush=True)
