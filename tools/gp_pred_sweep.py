"""K7 (gp_predict: factor + beta + conditional mean at T test times) through pipe / batch / smem: python tools/gp_pred_sweep.py [T]"""
import os, sys, json, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
T = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for S in (1024, 10000):
    for N in (30, 48, 64, 96, 120, 136, 160, 200):
        row = {"S": S, "N": N, "T": T}
        res = {}
        for k in ("pipe", "batch", "smem"):
            os.environ["RVLP_GP_KERNEL"] = k
            spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
            post = fit.from_spec(spec)
            th = torch.as_tensor(theta, device="cuda")
            times = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), max(T, 1), dtype=torch.float64, device="cuda")
            fn = (lambda: post.ctx.gp_predict(th, times, want_chi2=True)) if T else (lambda: post.ctx.gp_predict(th, times[:0], want_chi2=True))
            for _ in range(2): r = fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): r = fn()
            b.record(); torch.cuda.synchronize()
            row[k] = round(a.elapsed_time(b) / 5, 4)
            res[k] = r[1].cpu().numpy()
        ok = np.isfinite(res["pipe"])
        row["max_rel_dchi2_smem_pipe"] = float(np.max(np.abs(res["smem"][ok] - res["pipe"][ok]) / np.abs(res["pipe"][ok])))
        row["nan_equal"] = bool(np.array_equal(np.isnan(res["smem"]), np.isnan(res["pipe"])))
        print(json.dumps(row), flush=True)
