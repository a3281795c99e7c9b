"""Host-buffer path at a 1/8 shard of config 3 (125k rows, 29 MB): python tools/host_small.py"""
import os, sys, time, subprocess
if len(sys.argv) > 1:
    import numpy as np, torch
    sys.path.insert(0, ".")
    from ravest_b200 import fit, workloads
    spec, theta = workloads.make_c3(1_000_000)
    theta = np.ascontiguousarray(theta[:125_000])
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda"); out = torch.empty(len(theta), dtype=torch.float64, device="cuda")
    for _ in range(3): post.ctx.logprob(th, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): post.ctx.logprob(th, out=out)
    b.record(); torch.cuda.synchronize()
    print(f"threads {os.environ.get('RVLP_STAGE_THREADS')}: device-resident {a.elapsed_time(b) / 20:.3f} ms")
    for chunks in ("1", "2", "3", "5", "7"):
        os.environ["RVLP_HOST_CHUNKS"] = chunks
        for _ in range(3): post.log_probability_batch(theta)
        t0 = time.perf_counter()
        for _ in range(20): post.log_probability_batch(theta)
        print(f"   chunks {chunks}: {(time.perf_counter() - t0) * 50:.3f} ms", flush=True)
else:
    for th in ("1", "2", "4", "12"):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, RVLP_STAGE_THREADS=th))
