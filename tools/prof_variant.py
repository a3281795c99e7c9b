import sys, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
v = int(sys.argv[1]); S = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
spec, theta = workloads.make_c3(S)
post = fit.from_spec(spec); post.ctx.set_variant(v)
th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
for _ in range(3): post.ctx.logprob(th, out=out)
torch.cuda.synchronize()
