"""Host-buffer path (NumPy in / out) per-call time against the number of chunks (RVLP_HOST_CHUNKS), c2 and c3 sizes."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads
for name, S in (("c2", 100_000), ("c2", 1_000_000), ("c1", 100_000), ("c3", 100_000), ("c3", 1_000_000)):
    spec, theta = getattr(workloads, "make_" + name)(S)
    post = fit.from_spec(spec)
    out = np.empty(S)
    line = [f"{name} S={S} ({theta.nbytes / 1e6:.1f} MB)"]
    for ch in ("0", "1", "2", "3", "4", "5", "7"):
        if ch == "0": os.environ.pop("RVLP_HOST_CHUNKS", None)
        else: os.environ["RVLP_HOST_CHUNKS"] = ch
        for _ in range(3): post.ctx.logprob_host(theta, out)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 10
        for _ in range(n): post.ctx.logprob_host(theta, out)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
        line.append(f"{'auto' if ch == '0' else ch}: {dt * 1e3:.3f} ms")
    print("  ".join(line), flush=True)
