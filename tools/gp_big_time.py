"""Time the GP log-posterior kernels over the epoch count: pipelined register-tile kernel (N <= 219) vs the blocked
DMMA kernel (rvlp_gp_big.cuh; forced with RVLP_GP_KERNEL=big below 220).  python tools/gp_big_time.py"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads, _lib

peak, _ = _lib.measure_fp64_peak(0, 4096)
res = []
CASES = []
SEL = os.environ.get("GPT_KERNELS", "default,big,batch").split(",")
for N, S in ((30, 20000), (57, 20000), (120, 10000), (120, 1000), (120, 100), (200, 4000), (256, 4000), (512, 2000), (1024, 600), (1024, 32), (2048, 300)):
    for k in SEL:
        CASES.append((N, S, k))
for N, S, force in CASES:
    if force != "default":
        os.environ["RVLP_GP_KERNEL"] = force
    else:
        os.environ.pop("RVLP_GP_KERNEL", None)
    spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N, seed=505)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda")
    out = torch.empty(S, dtype=torch.float64, device="cuda")
    for _ in range(2):
        post.ctx.logprob(th, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        post.ctx.logprob(th, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    flops = N ** 3 / 3.0 + 100.0 * N * (N - 1) / 2      # Cholesky + covariance build (SURVEY 8d weights)
    r = {"shape": os.environ.get("RVLP_GP_BIG_SHAPE", "82"), "N": N, "S": S, "kernel": force, "ms": round(ms, 3), "us_per_sample": round(1e3 * ms / S, 3),
         "logprob_per_s": round(S / ms * 1e3), "chol_tflops": round(S * N ** 3 / 3.0 / ms / 1e9, 2),
         "frac_of_fp64_peak": round(S * flops / (ms * 1e-3) / peak, 3)}
    res.append(r)
    print(json.dumps(r), flush=True)
