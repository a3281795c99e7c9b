"""Phase breakdown of the blocked GP kernel (K3) on the GPU box.
Needs a library built with -DRVLP_GP_TIMING:  RVLP_LIB=$PWD/build_variants/gptiming.so python tools/gp_phase_time.py
Counters are clock64 cycles summed over the samples of CTA 0, read by the thread that owns the last tile."""
import ctypes as C, os, sys, numpy as np, torch
sys.path.insert(0, ".")
import ravest_b200
from ravest_b200 import fit, workloads, _lib
lib = _lib.load()
names = ["prologue", "resid", "cov(own tile)", "wait diag p0 (cov of others)", "update-tail + diag", "trsm phase", "own update",
         "final reduce", "samples", "diag owner section", "trsm section (I = Jt+1)"]
for npl, N, S in ((1, 120, 10000),):
    spec, theta = workloads.make_c5(n_samples=S, n_planets=npl, n_epochs=N)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
    for cap in ("0", "148", "296"):
        os.environ["RVLP_GP_GRID"] = cap
        for _ in range(2): post.ctx.logprob(th, out=out)
        buf = (C.c_ulonglong * 32)()
        lib.rvlp_debug_gp_timing(buf)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): post.ctx.logprob(th, out=out)
        b.record(); torch.cuda.synchronize(); ms = a.elapsed_time(b) / 3
        lib.rvlp_debug_gp_timing(buf)
        v = np.array(list(buf), dtype=np.float64)
        ns = max(v[8], 1.0)
        print(f"N={N} S={S} grid cap {cap}: {ms:.3f} ms  {S / ms * 1e3:.3e} logprob/s; CTA 0 did {ns / 3:.0f} samples per launch")
        tot = v[:8].sum()
        for k, nm in enumerate(names):
            if k == 8: continue
            print(f"   {nm:32s} {v[k] / ns:10.0f} cycles/sample  {100 * v[k] / tot:5.1f} %")
        print(f"   {'total (lap sum)':32s} {tot / ns:10.0f} cycles/sample")
