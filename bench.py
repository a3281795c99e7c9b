#!/usr/bin/env python
"""bench.py — throughput of the batched RV log-probability path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

A "step" is one pass of the hot path over one batch: S samples x N epochs x n_pl planets through
LogPosterior.log_probability_batch (one kernel launch, plus one all-gather of the S log-probs per
rank when N > 1).  Headline workload: BASELINE config 3 (5 planets, 1000 epochs, 1e6 samples PER
GPU - weak scaling: the sample axis shards with no data-path collective).  Prints ONE JSON line.

  value     (sample x epoch x planet) evaluations/s, whole job, theta resident in HBM
  e2e       same metric through the public API with HOST buffers (H2D of theta + D2H of the
            log-probs inside the timed region)
  roofline  fp64: algorithmic FLOPs (418 per unit at config 3) / CUDA-event time of
            the kernel, against the fp64 FMA peak measured live on this GPU by a dependent-free
            DFMA kernel (MEASURED_PEAKS.json carries HBM and bf16 only)
  cpu_baseline  the C oracle (a port of the reference's algorithm) on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic fp64 FLOPs per (sample x epoch x planet) unit, SURVEY.md §8d table
FLOPS_PER_UNIT = {"c1": 467.0, "c2": 437.0, "c3": 418.0, "c4": 507.0}
# algorithmic HBM bytes per unit: theta row read + one fp64 written per sample
WORKLOADS = {
    "c1": dict(maker="make_c1", samples=100_000, desc="51 Peg b-shaped: 1 planet x 153 epochs, e free"),
    "c2": dict(maker="make_c2", samples=100_000, desc="TOI-544-shaped: 2 planets x 120 epochs"),
    "c3": dict(maker="make_c3", samples=1_000_000, desc="synthetic 5 planets x 1000 epochs"),
    "c4": dict(maker="make_c4", samples=1_000_000, desc="high-e stress: 3 planets x 1000 epochs x 2 instruments, e<=0.97"),
    "c5": dict(maker="make_c5", samples=10_000, desc="K2-229-shaped quasi-periodic GP: 120 epochs"),
}


def make_workload(name: str, samples: int, rank: int = 0):
    from ravest_b200 import workloads
    w = WORKLOADS[name]
    base_seed = {"c1": 101, "c2": 202, "c3": 303, "c4": 404, "c5": 505}[name]
    # rank r of a weak-scaling run draws its own rows; rank 0 sees the single-GPU bytes
    return getattr(workloads, w["maker"])(samples, seed=base_seed + 1000 * rank)


def units_per_step(spec, n_samples: int) -> float:
    return float(n_samples) * len(spec["time"]) * len(spec["planet_letters"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(self.device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(spec, theta, target_seconds: float = 12.0) -> dict:
    """The C oracle (port of the reference's algorithm: Halley from E0=M, libm sin/cos) on all host
    threads, on a bounded prefix of the same theta."""
    from oracle import oracle_c
    orc = oracle_c.OracleProblem(spec)
    cores = oracle_c.max_threads()
    n0 = min(len(theta), 64 * cores)
    t0 = time.perf_counter()
    orc.logprob(theta[:n0])
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(len(theta), max(n0, n0 * target_seconds / dt)))
    t0 = time.perf_counter()
    orc.logprob(theta[:n])
    dt = time.perf_counter() - t0
    u = units_per_step(spec, n)
    return {"value": u / dt, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"first {n} of {len(theta)} samples, all {len(spec['time'])} epochs, {dt:.1f} s wall",
            "logprob_per_s": n / dt}


def run_reference(args) -> None:
    """--impl reference: the reference's CPU implementation of the path.  ravest is pure Python + numba and
    /root/reference does not exist on the GPU box, so this arm times the C port of its algorithm
    (oracle/oracle.c) with all host threads - `kind: port`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_c
    name = args.workload
    n_sample = args.ref_samples
    spec, theta = make_workload(name, n_sample)
    orc = oracle_c.OracleProblem(spec)
    cores = oracle_c.max_threads()
    for _ in range(args.warmup):
        orc.logprob(theta[: max(64, n_sample // 8)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.logprob(theta)
    dt = time.perf_counter() - t0
    u = units_per_step(spec, n_sample) * args.steps
    val = u / dt
    line = {
        "impl": "reference", "metric": "kepler_rv_evals_per_sec", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {WORKLOADS[name]['desc']}", "samples_per_step": n_sample,
                   "epochs": len(spec["time"]), "planets": len(spec["planet_letters"]),
                   "note": "bounded sample of the same workload (same generator, same seed)"},
        "logprob_per_s": n_sample * args.steps / dt,
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port",
                         "sample": f"{n_sample} samples x {args.steps} steps, OpenMP over samples"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_kernel(torch, fn, steps: int, warmup: int, barrier) -> float:
    """CUDA-event time (ms) of `steps` calls on the current stream, max over ranks done by the caller."""
    for _ in range(warmup):
        fn()
    barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    barrier()
    return a.elapsed_time(b)


def sample_matrix_rows(torch, fit, barrier) -> dict:
    """SURVEY.md §8 rows f-1..f-4 on one GPU: per-sample RV matrix (K2), percentile bands (K6), walker checks
    (K5), GP conditioning (K7).  HBM-bound rows are quoted against MEASURED_PEAKS.json's copy bandwidth."""
    from ravest_b200 import _lib, workloads
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm, hbm_src = 6550.0, "fallback (B200_PROFILING.md)"
    res = {"hbm_peak_gbs": hbm, "hbm_peak_source": hbm_src}
    spec, theta = workloads.make_c2(100_000)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda")
    S, T = len(theta), 1000
    times = torch.linspace(float(spec["time"].min()), float(spec["time"].max()), T, dtype=torch.float64, device="cuda")
    m = torch.empty((S, T), dtype=torch.float64, device="cuda")
    ms = time_kernel(torch, lambda: post.ctx.rv_matrix(th, times, -2, out=m), 5, 2, barrier) / 5
    res["f1_rv_matrix"] = {"shape": f"c2 posterior: {S} samples x {T} times x 2 planets", "ms": ms,
                           "evals_per_s": S * T * 2 / ms * 1e3, "write_gbs": S * T * 8 / ms / 1e6}
    out = torch.empty((3, T), dtype=torch.float64, device="cuda")
    ms = time_kernel(torch, lambda: _lib.percentile_columns(m, [15.85, 50, 84.15], out=out), 5, 2, barrier) / 5
    res["f2_percentile_bands"] = {"shape": f"{S} x {T} fp64 matrix ({S * T * 8 / 1e6:.0f} MB), q = [15.85, 50, 84.15]",
                                  "ms": ms, "streaming_passes": 4, "achieved_gbs": 4 * S * T * 8 / ms / 1e6,
                                  "frac_of_hbm_peak": 4 * S * T * 8 / ms / 1e6 / hbm,
                                  "matrix_reads_per_s_gbs": S * T * 8 / ms / 1e6}
    del m
    spec3, theta3 = workloads.make_c3(1_000_000)
    post3 = fit.from_spec(spec3)
    th3 = torch.as_tensor(theta3, device="cuda")
    ms = time_kernel(torch, lambda: post3.ctx.walker_check(th3), 5, 2, barrier) / 5
    res["f3_walker_check"] = {"shape": "c3: 1e6 candidate rows x 29 columns", "ms": ms, "rows_per_s": 1e6 / ms * 1e3,
                              "read_gbs": theta3.nbytes / ms / 1e6}
    del th3, post3
    spec5, theta5 = workloads.make_c5(10_000)
    post5 = fit.from_spec(spec5)
    th5 = torch.as_tensor(theta5, device="cuda")
    t5 = torch.linspace(float(spec5["time"].min()), float(spec5["time"].max()), T, dtype=torch.float64, device="cuda")
    ms = time_kernel(torch, lambda: post5.ctx.gp_predict(th5, t5), 3, 1, barrier) / 3
    res["f4_gp_conditioning"] = {"shape": f"c5: 1e4 samples x {len(spec5['time'])} epochs -> {T} test times", "ms": ms,
                                 "samples_per_s": 1e4 / ms * 1e3}
    return res


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="samples per GPU (default: the workload's)")
    ap.add_argument("--ref-samples", type=int, default=4000)
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads / cpu baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import ravest_b200
    from ravest_b200 import _lib, fit
    from ravest_b200 import dist as rdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ravest_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    ravest_b200.load()
    name = args.workload
    S = args.samples or WORKLOADS[name]["samples"]
    spec, theta = make_workload(name, S, rank)
    post = fit.from_spec(spec)
    th = torch.as_tensor(theta, device="cuda")
    out = torch.empty(S, dtype=torch.float64, device="cuda")
    gathered = torch.empty(S * world, dtype=torch.float64, device="cuda") if world > 1 else None
    ctx = post.ctx
    units = units_per_step(spec, S) * world

    def step():
        ctx.logprob(th, out=out)
        if world > 1:            # every rank (and the host sampler) sees all log-probs: the path's one exchange
            dist.all_gather_into_tensor(gathered, out)

    def kernel_only():
        ctx.logprob(th, out=out)

    # fp64 peak of this GPU (roofline denominator), measured before the timed region
    peak_flops, _ = _lib.measure_fp64_peak(local, 4096)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = ravest_b200.launch_count()
    ms = time_kernel(torch, step, args.steps, args.warmup, barrier)
    launches = ravest_b200.launch_count() - n0 - args.warmup
    clocks = sampler.stop() if sampler else None
    ms_kernel = time_kernel(torch, kernel_only, args.steps, 1, barrier)

    # e2e: NumPy / pinned host buffers in and out through the public API (H2D + kernel + D2H per step)
    theta_pinned = torch.as_tensor(theta).pin_memory()
    theta_host = theta_pinned.numpy()
    host_out = np.empty(S)
    for _ in range(2):
        ctx.logprob_host(theta_host, host_out)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.logprob_host(theta_host, host_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    same = np.array_equal(host_out.view(np.int64), out.cpu().numpy().view(np.int64))

    t = torch.tensor([ms, ms_kernel, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_kernel, e2e_ms = (float(x) for x in t.cpu())

    if rank == 0:
        value = units * args.steps / (ms * 1e-3)
        fpu = FLOPS_PER_UNIT.get(name)
        per_gpu_units_per_s = (units / world) * args.steps / (ms_kernel * 1e-3)
        roofline = None
        if fpu:
            achieved = per_gpu_units_per_s * fpu / 1e12
            traffic = None
            try:      # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this launch shape
                tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(name)
                if tj and tj["samples"] == S:
                    traffic = tj["dram_bytes_per_launch"]
            except Exception:
                traffic = None
            hw = None
            try:      # hardware view of the same launch shape from the committed ncu capture
                hw = json.load(open(os.path.join(ROOT, "profiles", "r01_hw.json"))).get(name)
            except Exception:
                hw = None
            roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_flops / 1e12, "unit": "TFLOP/s",
                        "frac": achieved / (peak_flops / 1e12), "traffic": traffic, "hw": hw,
                        "peak_source": "measured live: dependent-free DFMA kernel (rvlp_measure_fp64_peak), burst",
                        "flops_per_unit": fpu, "kernel": "rvlp::logprob_kernel",
                        "kernel_ms_per_launch": ms_kernel / args.steps,
                        "algorithmic_hbm_bytes_per_launch": S * (theta.shape[1] + 1) * 8,
                        "note": "frac > 1 is expected: the 418-FLOP/unit figure counts the REFERENCE algorithm (4 Halley "
                                "passes with a libm sincos each); this kernel needs ~56 fp64 instructions per unit (fp32 "
                                "starter + one fp64 step). Hardware view: profiles/ (ncu sm__inst_executed_pipe_fp64)"}
        line = {
            "metric": "kepler_rv_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {WORKLOADS[name]['desc']}", "samples_per_gpu": S,
                       "epochs": len(spec["time"]), "planets": len(spec["planet_letters"]), "ndim": int(theta.shape[1]),
                       "parallelism": f"sample-sharded x{world}, 1 all-gather of log-probs per step" if world > 1 else "single GPU",
                       "l2": f"theta is {theta.nbytes / 1e6:.0f} MB per GPU (> 126 MB L2), streamed once per step"},
            "logprob_per_s": S * world * args.steps / (ms * 1e-3),
            "e2e": {"value": units * args.steps / (e2e_ms * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": int(theta.nbytes), "d2h_bytes_per_step": int(S * 8),
                    "ms_per_step": e2e_ms / args.steps, "api": "LogPosterior.log_probability_batch(numpy) -> rvlp_logprob_batch_host",
                    "bit_identical_to_device_path": bool(same)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
        }
        if world == 1 and not args.no_extras:
            line["cpu_baseline"] = cpu_baseline(spec, theta)
            others = {}
            for other in ("c1", "c2", "c4", "c5"):
                if other == name:
                    continue
                try:
                    s2, t2 = make_workload(other, WORKLOADS[other]["samples"])
                    p2 = fit.from_spec(s2)
                    th2 = torch.as_tensor(t2, device="cuda")
                    o2 = torch.empty(len(t2), dtype=torch.float64, device="cuda")
                    m2 = time_kernel(torch, lambda: p2.ctx.logprob(th2, out=o2), 5, 2, barrier)
                    u2 = units_per_step(s2, len(t2)) * 5 / (m2 * 1e-3)
                    entry = {"evals_per_s": u2, "logprob_per_s": len(t2) * 5 / (m2 * 1e-3), "ms_per_step": m2 / 5,
                             "samples": len(t2)}
                    if other in FLOPS_PER_UNIT:
                        entry["roofline_frac"] = u2 * FLOPS_PER_UNIT[other] / peak_flops
                    elif other == "c5":
                        # ~1.36e6 algorithmic fp64 FLOPs per GP log-prob at N = 120, one planet
                        # (mean model 4.9e4 + covariance build 7.1e5 + Cholesky N^3/3 5.8e5 + solve / log-det 2e4)
                        entry["flops_per_logprob"] = 1.36e6
                        entry["roofline_frac"] = entry["logprob_per_s"] * 1.36e6 / peak_flops
                        entry["kernel"] = "rvlp::gp_logprob_pipe_kernel<6, false>"
                    others[other] = entry
                    del p2, th2, o2
                except Exception as ex:       # report, never hide
                    others[other] = {"error": repr(ex)}
            line["other_workloads"] = others
            try:
                line["sample_matrix_rows"] = sample_matrix_rows(torch, fit, barrier)
            except Exception as ex:           # report, never hide
                line["sample_matrix_rows"] = {"error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
