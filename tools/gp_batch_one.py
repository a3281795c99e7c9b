"""One call of the batched GP path for ncu launch lists: python tools/gp_batch_one.py N S [kernel]"""
import os, sys, torch
sys.path.insert(0, ".")
os.environ["RVLP_GP_KERNEL"] = sys.argv[3] if len(sys.argv) > 3 else "batch"
from ravest_b200 import fit, workloads
N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
for _ in range(2): post.ctx.logprob(th, out=out)
torch.cuda.synchronize(); print("ok", float(out[5]))
