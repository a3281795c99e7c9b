"""K1 at the bench shapes for several floors of the guided schedule (RVLP_GSS_UNITS): python tools/k1_gss_sweep.py"""
import os, subprocess, sys
CHILD = r"""
import sys, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
for name, S in (("c3", 1_000_000), ("c3", 125_000), ("c4", 1_000_000), ("c2", 100_000), ("c1", 100_000)):
    spec, theta = getattr(workloads, "make_" + name)(S)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
    res = []
    for v in (0, 1):
        post.ctx.set_variant(v)
        for _ in range(3): post.ctx.logprob(th, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): post.ctx.logprob(th, out=out)
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 10)
    print(f"  {name} S={S}: v0 {res[0]:.3f} ms  v1 {res[1]:.3f} ms", flush=True)
"""
for gu in sys.argv[1:] or ["1", "1024", "4096", "16384", "100000000"]:
    print("RVLP_GSS_UNITS =", gu, flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, RVLP_GSS_UNITS=gu))
