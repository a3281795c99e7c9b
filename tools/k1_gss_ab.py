"""A/B of the guided schedule's floor inside ONE process (contexts created under different RVLP_GSS_UNITS), interleaved."""
import os, sys, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
FLOORS = sys.argv[1:] or ["100000000", "4096"]
for name, S in (("c3", 1_000_000), ("c3", 125_000), ("c4", 1_000_000), ("c2", 100_000)):
    spec, theta = getattr(workloads, "make_" + name)(S)
    th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
    posts = []
    for f in FLOORS:
        os.environ["RVLP_GSS_UNITS"] = f
        p = fit.from_spec(spec); p.ctx.set_variant(0); posts.append(p)
    for p in posts:
        for _ in range(3): p.ctx.logprob(th, out=out)
    torch.cuda.synchronize()
    acc = [[] for _ in posts]
    for rnd in range(5):
        for i, p in enumerate(posts):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): p.ctx.logprob(th, out=out)
            b.record(); torch.cuda.synchronize()
            acc[i].append(a.elapsed_time(b) / 5)
    print(name, S, "  ".join(f"floor {f}: min {min(x):.3f} med {sorted(x)[2]:.3f} ms" for f, x in zip(FLOORS, acc)), flush=True)
