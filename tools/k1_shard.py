"""K1 at the shard sizes of a strong-scaling run of config 3: python tools/k1_shard.py"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
spec, theta = workloads.make_c3(1_000_000)
post = fit.from_spec(spec)
full = torch.as_tensor(theta, device="cuda")
for S in (1_000_000, 500_000, 250_000, 125_000, 62_500):
    th = full[:S].contiguous(); out = torch.empty(S, dtype=torch.float64, device="cuda")
    for _ in range(3): post.ctx.logprob(th, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): post.ctx.logprob(th, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"S={S}: {ms:.3f} ms  {S * 5000 / ms / 1e6:.1f} G units/s  ({1e6 / S * ms:.2f} ms per 1e6)", flush=True)
