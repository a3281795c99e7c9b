#!/usr/bin/env python
"""Time alternative builds of the CUDA library (different -D tuning macros) on the same workloads.

    python tools/kernel_sweep.py build   # here: compile variants into build_variants/
    python tools/kernel_sweep.py run     # on the GPU box: time each variant (+ parity spot check)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "build_variants")
VARIANTS = {
    "opt07": ["-DRVLP_OPT=7"],
    "opt23": ["-DRVLP_OPT=23"],
    "opt55": ["-DRVLP_OPT=55"],
    "opt63": ["-DRVLP_OPT=63"],
    "opt51": ["-DRVLP_OPT=51"],
    "opt53": ["-DRVLP_OPT=53"],
}

CHILD = r"""
import sys, json, time, numpy as np, torch
sys.path.insert(0, %r)
import ravest_b200
from ravest_b200 import fit, workloads
from oracle import oracle_c
res = {}
for name, maker, S in (("c3", workloads.make_c3, 200000), ("c4", workloads.make_c4, 200000), ("c2", workloads.make_c2, 100000)):
    spec, theta = maker(S)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda"); out = torch.empty(S, dtype=torch.float64, device="cuda")
    best = None
    for v in (0, 1):
        post.ctx.set_variant(v)
        for _ in range(3): post.ctx.logprob(th, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): post.ctx.logprob(th, out=out)
        b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) / 5
        res.setdefault(name + "_v", []).append(round(t, 3))
        best = t if best is None else min(best, t)
    ms = best
    units = S * len(spec["time"]) * len(spec["planet_letters"])
    ref = oracle_c.OracleProblem(spec).logprob(theta[:256])
    got = out[:256].cpu().numpy()
    fin = np.isfinite(ref)
    err = float(np.max(np.abs(got[fin] - ref[fin]) / (1e-7 + 2e-13 * np.abs(ref[fin]))))
    ok = bool(np.array_equal(np.isneginf(got), np.isneginf(ref)) and err <= 1.0)
    res[name] = {"ms": ms, "Gunits_s": units / ms / 1e6, "parity_ok": ok, "err_over_tol": err}
print("RESULT " + json.dumps(res))
"""


def build():
    os.makedirs(VDIR, exist_ok=True)
    src = os.path.join(ROOT, "ravest_b200", "csrc", "rvlp_capi.cu")
    procs = []
    for f in os.listdir(VDIR):
        if f.endswith(".so"):
            os.unlink(os.path.join(VDIR, f))
    for name, flags in VARIANTS.items():
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-shared"] + flags + ["-o", out, src]
        procs.append((out, subprocess.Popen(cmd)))
    for out, p in procs:
        assert p.wait() == 0, out
        print("built", out)


def run():
    for f in sorted(os.listdir(VDIR)):
        if not f.endswith(".so"):
            continue
        env = dict(os.environ, RVLP_LIB=os.path.join(VDIR, f))
        r = subprocess.run([sys.executable, "-c", CHILD % ROOT], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(f, "FAILED", r.stderr[-400:])
            continue
        res = json.loads(line[0][7:])
        print(f, " ".join(f"{k}: {v['ms']:.3f} ms {v['Gunits_s']:.1f} G/s ok={v['parity_ok']}({v['err_over_tol']:.2g}) v0/v1={res[k + '_v']}"
                          for k, v in res.items() if not k.endswith("_v")), flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
