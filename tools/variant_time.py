"""Time every compiled shape of K1 on the bench workloads and check that they agree bit for bit."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
S = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
VARIANTS = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
for name, maker in (("c3", workloads.make_c3), ("c4", workloads.make_c4), ("c2", workloads.make_c2), ("c1", workloads.make_c1)):
    spec, theta = maker(S)
    th = torch.as_tensor(theta, device="cuda")
    units = S * len(spec["time"]) * len(spec["planet_letters"])
    ref = None
    line = [name]
    for v in VARIANTS:
        post = fit.from_spec(spec)
        post.ctx.set_variant(v)
        out = torch.empty(S, dtype=torch.float64, device="cuda")
        for _ in range(2):
            post.ctx.logprob(th, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            post.ctx.logprob(th, out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        got = out.cpu().numpy()
        if ref is None:
            ref = got
        same = np.array_equal(got.view(np.int64), ref.view(np.int64))
        line.append(f"v{v}: {ms:.3f} ms {units / ms / 1e6:.1f} G/s same={same}")
    print("  ".join(line), flush=True)
