"""One K1 launch at the c3 bench shape (200k samples) for ncu captures."""
import sys, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
spec, theta = getattr(workloads, "make_" + name)(S)
post = fit.from_spec(spec)
post.ctx.set_variant(0)
th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
for _ in range(3): post.ctx.logprob(th, out=out)
torch.cuda.synchronize(); print("ok", float(out[5]))
