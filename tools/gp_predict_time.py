"""Timing + parity of the GP conditioning path (K7, row f-4) on the GPU box."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads


def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for N, S in ((120, 10000), (57, 10000), (170, 4000)):
    spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda")
    t1000 = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), 1000, dtype=torch.float64, device="cuda")
    tN = torch.as_tensor(np.asarray(spec["time"], dtype=np.float64), device="cuda")
    ms1000 = timeit(lambda: post.ctx.gp_predict(th, t1000))
    msN = timeit(lambda: post.ctx.gp_predict(th, tN))
    print(f"N={N} S={S}: T=1000 {ms1000:.3f} ms, T=N {msN:.3f} ms", flush=True)
