"""e2e host-buffer path from pageable NumPy memory at config 3 over staging sub-block size / threads: python tools/host_stage.py"""
import os, sys, time, subprocess
if len(sys.argv) > 1:
    import numpy as np, torch
    sys.path.insert(0, ".")
    from ravest_b200 import fit, workloads
    spec, theta = workloads.make_c3(1_000_000)
    post = fit.from_spec(spec)
    for mb in ("1024", "32", "16", "8", "4"):
        for chunks in ("7", "5"):
            os.environ["RVLP_HOST_STAGE_MB"] = mb; os.environ["RVLP_HOST_CHUNKS"] = chunks
            for _ in range(3): post.log_probability_batch(theta)
            t0 = time.perf_counter()
            for _ in range(10): post.log_probability_batch(theta)
            print(f"threads {os.environ.get('RVLP_STAGE_THREADS')} stage block {mb} MB chunks {chunks}: {(time.perf_counter() - t0) * 100:.3f} ms", flush=True)
else:
    for th in ("2", "4", "8", "12"):
        subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, RVLP_STAGE_THREADS=th))
