"""Latency of emcee-sized batches through the host-buffer API (NumPy in/out) and the device API."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from ravest_b200 import fit, workloads
for name, maker in (("c1 1pl x 153", lambda S: workloads.make_c1(S)), ("c2 2pl x 120", lambda S: workloads.make_c2(S)), ("c3 5pl x 1000", lambda S: workloads.make_c3(S))):
    for S in (16, 32, 128, 1024, 8192):
        spec, theta = maker(S)
        post = fit.from_spec(spec)
        th = torch.as_tensor(theta, device="cuda")
        for _ in range(20): post.log_probability_batch(theta)
        t0 = time.perf_counter(); n = 300
        for _ in range(n): post.log_probability_batch(theta)
        host_us = (time.perf_counter() - t0) / n * 1e6
        for _ in range(20): post.log_probability_batch(th)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): post.log_probability_batch(th)
        torch.cuda.synchronize(); dev_us = (time.perf_counter() - t0) / n * 1e6
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = torch.empty(S, dtype=torch.float64, device="cuda")
        a.record()
        for _ in range(n): post.ctx.logprob(th, out=out)
        b.record(); torch.cuda.synchronize(); k_us = a.elapsed_time(b) / n * 1e3
        print(f"{name} S={S:5d}: numpy API {host_us:8.1f} us/call ({S/host_us*1e6:.3e} logp/s)  device API {dev_us:7.1f} us  kernel-stream {k_us:7.1f} us")
