import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from ravest_b200 import fit, workloads
spec, theta = workloads.make_c5(n_samples=10000)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device='cuda'); out = torch.empty(len(theta), dtype=torch.float64, device='cuda')
for _ in range(3): post.ctx.logprob(th, out=out)
torch.cuda.synchronize(); print("ok", float(out[5]))
