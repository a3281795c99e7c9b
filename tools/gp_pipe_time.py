"""Per-warp phase breakdown of the pipelined GP kernel (K3) on the GPU box.
Needs a -DRVLP_GP_TIMING build:  RVLP_LIB=$PWD/build_variants/gptiming.so python tools/gp_pipe_time.py [N] [S]
Cycles (clock64) of lane 0 of each warp of CTA 0, per sample."""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, ".")
from ravest_b200 import fit, workloads, _lib
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 120
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
names = ["wait record", "build tiles", "diag / idle", "barrier A", "trsm + barrier B", "update+final", "produce", "AD+catchup/CU"]
spec, theta = workloads.make_c5(n_samples=S, n_planets=1, n_epochs=N)
post = fit.from_spec(spec)
th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
for cap in ("0", "148"):
    os.environ["RVLP_GP_GRID"] = cap
    for _ in range(2): post.ctx.logprob(th, out=out)
    buf = (C.c_ulonglong * 64)()
    lib.rvlp_debug_gp_pipe_timing(buf)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): post.ctx.logprob(th, out=out)
    b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 3
    lib.rvlp_debug_gp_pipe_timing(buf)
    v = np.array(list(buf), dtype=np.float64).reshape(8, 8)
    grid = 296 if cap == "0" else int(cap)
    ns = 3 * len(range(0, S, grid))
    print(f"N={N} S={S} grid cap {cap}: {ms:.3f} ms  {S / ms * 1e3:.3e} logprob/s; ~{ns / 3:.0f} samples per CTA; cycles per sample:")
    print("   warp " + " ".join(f"{n[:13]:>14s}" for n in names) + f" {'sum':>10s}")
    for w in range(8):
        print(f"   w{w} r{(7 - w) if w < 4 else (w - 4)} " + " ".join(f"{v[w, k] / ns:14.0f}" for k in range(8)) + f" {v[w].sum() / ns:10.0f}")
