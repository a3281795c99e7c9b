#!/usr/bin/env python
"""bench.py — throughput of the batched RV log-probability path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3] [--scaling strong|weak]

A "step" is one pass of the hot path over one batch: S samples x N epochs x n_pl planets through
LogPosterior.log_probability_batch (one kernel launch per rank; when N > 1 the all-gather of the S log-probs is fused
into that kernel as NVLink peer stores + a one-warp flag barrier, and the NCCL all-gather is timed beside it).
Headline workload: BASELINE config 3 as named - 5 planets, 1000 epochs, **1e6 samples in total, sharded over the N
GPUs** (`"scaling": "strong"`, the product's `ravest_b200.dist.sharded_logprob`).  The weak-scaling point
(1e6 samples PER GPU) is measured in the same run and reported under `weak_scaling`.  Prints ONE JSON line.

  value      (sample x epoch x planet) evaluations/s, whole job, theta resident in HBM
  e2e        same metric through the public API with HOST buffers: a PAGEABLE NumPy array in, NumPy out (what an
             emcee / harmonic caller holds); H2D of theta + D2H of the log-probs inside the timed region.  The
             pinned-buffer variant is reported beside it (`e2e.pinned`).
  roofline   fp64.  `frac` is the HARDWARE fraction: fp64-pipe instructions this kernel executes per unit (from the
             committed ncu capture, profiles/r02_hw.json) x units/s, against the fp64 FMA issue peak measured live on
             this GPU by a dependent-free DFMA kernel.  `algorithmic_frac` is SURVEY.md §8(d)'s contract figure
             (418 reference-algorithm FLOPs per unit at config 3) over the same peak - it exceeds 1 because the
             kernel needs far fewer operations than the reference algorithm, not because work is skipped.
  cpu_baseline  the reference's own `LogPosterior.log_probability` (unmodified ravest from oracle/_ref) on the host
             cores - one core and a spawn pool of all cores (its own mechanism, fit.py:1069) - on a bounded sample of
             the same theta; falls back to the C port (oracle/oracle.c) only if the reference cannot be imported.
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
WORKLOADS = {
    "c1": dict(maker="make_c1", samples=100_000, desc="51 Peg b-shaped: 1 planet x 153 epochs, e free"),
    "c2": dict(maker="make_c2", samples=100_000, desc="TOI-544-shaped: 2 planets x 120 epochs"),
    "c3": dict(maker="make_c3", samples=1_000_000, desc="synthetic 5 planets x 1000 epochs"),
    "c4": dict(maker="make_c4", samples=1_000_000, desc="high-e stress: 3 planets x 1000 epochs x 2 instruments, e<=0.97"),
    "c5": dict(maker="make_c5", samples=10_000, desc="K2-229-shaped quasi-periodic GP: 120 epochs"),
}
BASE_SEED = {"c1": 101, "c2": 202, "c3": 303, "c4": 404, "c5": 505}
METRIC = "kepler_rv_evals_per_sec"


def make_workload(name: str, samples: int, rank: int = 0):
    from ravest_b200 import workloads
    # rank r of a WEAK-scaling run draws its own rows; rank 0 (and every rank of a strong-scaling run) sees the
    # single-GPU bytes
    return getattr(workloads, WORKLOADS[name]["maker"])(samples, seed=BASE_SEED[name] + 1000 * rank)


def units_per_step(spec, n_samples: int) -> float:
    return float(n_samples) * len(spec["time"]) * len(spec["planet_letters"])


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


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
        # the sampler also sees the idle edges of the region: the clock "under load" is the median of the upper half
        hot = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(hot)) if hot else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU baselines
def reference_baseline(spec, theta, seconds: float = 14.0, workers: int | None = None) -> dict | None:
    """The UNMODIFIED reference (oracle/_ref): LogPosterior.log_probability on one core and on a spawn pool of all host
    cores, bounded sample of the same theta.  None when the reference cannot be imported here."""
    from oracle import ref_runner
    ok, why = ref_runner.available()
    if not ok:
        return {"unavailable": why}
    cores = workers or host_cores()
    run = ref_runner.ReferenceRunner(spec, workers=cores)
    try:
        per1, n1 = run.time_serial(theta, seconds=min(3.0, seconds / 4))
        upl = len(spec["time"]) * len(spec["planet_letters"])
        n_warm = min(len(theta), 64 * cores)
        run.evaluate(theta[:n_warm])                                     # spawn + import + numba cache in every worker
        n = int(min(len(theta), max(n_warm, 0.7 * seconds / per1 * cores)))
        t0 = time.perf_counter()
        out = run.evaluate(theta[:n])
        dt = time.perf_counter() - t0
        return {"value": n * upl / dt, "unit": "evals/s", "cores": cores, "kind": "reference",
                "sample": f"first {n} of {len(theta)} samples x all {len(spec['time'])} epochs through the unmodified "
                          f"ravest LogPosterior.log_probability (one dict per row) on a spawn pool of {cores} "
                          f"workers (fit.py:1069), {dt:.1f} s wall",
                "logprob_per_s": n / dt, "us_per_logprob_pool": 1e6 * dt / n,
                "single_core": {"value": upl / per1, "us_per_logprob": 1e6 * per1, "rows": n1},
                "_rows": n, "_out": out}
    finally:
        run.close()


def port_baseline(spec, theta, seconds: float = 6.0) -> dict:
    """The C restatement (oracle/oracle.c) on all host threads: the 'best case CPU' line, and the fallback baseline."""
    from oracle import oracle_c
    orc = oracle_c.OracleProblem(spec)
    cores = host_cores()          # explicit: torchrun exports OMP_NUM_THREADS=1, which must not cripple this arm
    n0 = min(len(theta), 64 * cores)
    t0 = time.perf_counter()
    orc.logprob(theta[:n0], nthreads=cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    n = int(min(len(theta), max(n0, n0 * seconds / dt)))
    t0 = time.perf_counter()
    out = orc.logprob(theta[:n], nthreads=cores)
    dt = time.perf_counter() - t0
    return {"value": units_per_step(spec, n) / dt, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": f"first {n} of {len(theta)} samples, all {len(spec['time'])} epochs, OpenMP x{cores}, {dt:.1f} s wall",
            "logprob_per_s": n / dt, "_rows": n, "_out": out}


def parity_stats(got: np.ndarray, ref: np.ndarray) -> dict:
    """Measured deviation of GPU log-probabilities from a CPU evaluation of the same rows (checker role only)."""
    fin = np.isfinite(ref)
    err = np.abs(got[fin] - ref[fin])
    mag = np.abs(ref[fin])
    small = mag < 1e6                       # rows where ulp(|logp|) << 1e-7, i.e. the bare absolute bound is meaningful
    return {"rows": int(len(ref)), "finite_rows": int(fin.sum()),
            "max_abs_dlogp": float(err.max()) if fin.any() else 0.0,
            "max_abs_dlogp_where_abs_logp_below_1e6": float(err[small].max()) if small.any() else 0.0,
            "max_dlogp_in_ulps_of_logp": float((err / np.spacing(np.maximum(mag, 1.0))).max()) if fin.any() else 0.0,
            "max_abs_logp": float(mag.max()) if fin.any() else 0.0,
            "neg_inf_pattern_equal": bool(np.array_equal(np.isneginf(got), np.isneginf(ref)))}


def cpu_baseline(spec, theta, gpu_out: np.ndarray | None = None) -> dict:
    """cpu_baseline leg (rank 0, N = 1): the reference on the host cores, the C port beside it, and - the oracle in its
    checker role - the measured deviation of the GPU results from both on the rows they evaluated."""
    base = reference_baseline(spec, theta)
    port = port_baseline(spec, theta)
    if base is None or "unavailable" in base:
        why = (base or {}).get("unavailable", "unknown")
        base = dict(port)
        base["note"] = f"reference not importable here ({why}): C port of its algorithm instead"
    else:
        base["c_port"] = {k: v for k, v in port.items() if not k.startswith("_")}
    if gpu_out is not None:
        par = {}
        for tag, b in (("reference" if base["kind"] == "reference" else "port", base), ("port", port)):
            ref, n = b["_out"], b["_rows"]
            got = gpu_out[:n]
            fin = np.isfinite(ref)
            par[tag] = parity_stats(got, ref)
        base["gpu_parity_on_sample"] = par
    for b in (base, port):
        b.pop("_out", None); b.pop("_rows", None)
    return base


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores - the unmodified
    ravest `LogPosterior.log_probability` (installed copy oracle/_ref) under its own spawn pool with all
    cores; each step is a bounded sample of the same workload (same generator and seed).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_runner
    name = args.workload
    cores = host_cores()
    ok, why = ref_runner.available()
    n_sample = args.ref_samples or (max(2000, 1250 * cores) if ok else 20000)
    spec, theta = make_workload(name, n_sample)
    upl = len(spec["time"]) * len(spec["planet_letters"])
    single = None
    if ok:
        run = ref_runner.ReferenceRunner(spec, workers=cores)
        per1, n1 = run.time_serial(theta, seconds=2.0)
        single = {"value": upl / per1, "us_per_logprob": 1e6 * per1, "rows": n1}
        step = lambda rows: run.evaluate(rows)
        kind = "reference"
        how = (f"unmodified ravest LogPosterior.log_probability, one dict per row, spawn pool of {cores} workers "
               f"(fit.py:1069); {n_sample} rows per step")
    else:
        from oracle import oracle_c
        orc = oracle_c.OracleProblem(spec)
        step = lambda rows: orc.logprob(rows, nthreads=cores)
        kind = "port"
        how = f"C port (oracle/oracle.c), OpenMP x{cores}; reference not importable: {why}"
    try:
        for _ in range(max(args.warmup, 1)):
            step(theta[: max(64 * cores, n_sample // 4)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(theta)
        dt = time.perf_counter() - t0
    finally:
        if ok:
            run.close()
    val = upl * n_sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {WORKLOADS[name]['desc']}", "samples_per_step": n_sample,
                   "epochs": len(spec["time"]), "planets": len(spec["planet_letters"]),
                   "note": "bounded sample of the same workload (same generator, same seed); rate metric"},
        "logprob_per_s": n_sample * args.steps / dt,
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": kind, "sample": how,
                         "single_core": single},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU timing
L2_BYTES = 126 << 20
MIN_WARM_MS = 250.0
MIN_WARM_MS_MULTI = 1500.0   # multi-GPU regions (collectives: NCCL channels, proxy threads and peer mappings settle)


def time_kernel(torch, fn, steps: int, warmup: int, barrier, flush=None, agree=None) -> float:
    """CUDA-event time (ms) of `steps` calls on the current stream, max over ranks done by the caller.
    flush: a device buffer larger than L2; when given it is rewritten BETWEEN the timed iterations (each iteration has
    its own event pair, the flush sits outside them) so that no iteration finds its inputs in L2."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    # the W warm-up steps of a 3 ms step are 9 ms of work: not enough for the clocks to settle after the idle set-up
    # phase (the first timed region of a multi-GPU run read 0.3 ms per step high).  Warm up for >= MIN_WARM_MS of work.
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    one = max(time.perf_counter() - t0, 1e-5)
    extra = min(1000, int((MIN_WARM_MS if agree is None else MIN_WARM_MS_MULTI) * 1e-3 / one))
    if agree is not None:                   # fn may hold a collective: every rank runs the same number of calls
        extra = agree(extra)
    for _ in range(extra):
        fn()
    barrier()
    torch.cuda.synchronize()
    import ravest_b200
    l0 = ravest_b200.launch_count()
    if flush is None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        time_kernel.launches = ravest_b200.launch_count() - l0      # this library's kernels inside the timed region
        barrier()
        return a.elapsed_time(b)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.add_(1)                       # read + write 2 x L2 bytes
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    time_kernel.launches = ravest_b200.launch_count() - l0
    barrier()
    return float(sum(a.elapsed_time(b) for a, b in ev))


def time_host(torch, fn, steps: int, warmup: int, barrier) -> float:
    """Wall time (ms) of `steps` synchronous host-buffer calls (each ends with its own device synchronisation)."""
    for _ in range(warmup):
        fn()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return dt * 1e3


def hbm_peak() -> tuple[float, str]:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json"
    except Exception:
        return 6550.0, "fallback (B200_PROFILING.md)"


def load_profile_json(name: str, key: str):
    for rnd in ("r02", "r01"):
        try:
            j = json.load(open(os.path.join(ROOT, "profiles", f"{rnd}_{name}.json"))).get(key)
            if j:
                j = dict(j)
                j["file"] = f"profiles/{rnd}_{name}.json"
                return j
        except Exception:
            pass
    return None


def sample_matrix_rows(torch, fit, barrier) -> dict:
    """SURVEY.md §8 rows f-1..f-4 on one GPU: per-sample RV matrix (K2), percentile bands (K6), walker checks
    (K5), GP conditioning (K7).  HBM-bound rows are quoted against MEASURED_PEAKS.json's copy bandwidth."""
    from ravest_b200 import _lib, workloads
    hbm, hbm_src = hbm_peak()
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
    launches = time_kernel.launches // 5
    res["f2_percentile_bands"] = {"shape": f"{S} x {T} fp64 matrix ({S * T * 8 / 1e6:.0f} MB), q = [15.85, 50, 84.15]",
                                  "ms": ms, "launches_per_call": launches,
                                  "algorithmic_gbs": S * T * 8 / ms / 1e6,
                                  "frac_of_hbm_peak": S * T * 8 / ms / 1e6 / hbm,
                                  "note": "algorithmic bytes = ONE read of the matrix"}
    del m
    spec3, theta3 = workloads.make_c3(1_000_000)
    post3 = fit.from_spec(spec3)
    th3 = torch.as_tensor(theta3, device="cuda")
    ms = time_kernel(torch, lambda: post3.ctx.walker_check(th3), 5, 2, barrier) / 5
    res["f3_walker_check"] = {"shape": "c3: 1e6 candidate rows x 29 columns", "ms": ms, "rows_per_s": 1e6 / ms * 1e3,
                              "read_gbs": theta3.nbytes / ms / 1e6}
    ms = time_kernel(torch, lambda: post3.information_criteria_batch(th3), 3, 1, barrier) / 3
    res["info_criteria"] = {"shape": "c3: log-likelihood, chi2, AICc, BIC for 1e6 rows (fit.py:1361-1554)", "ms": ms,
                            "rows_per_s": 1e6 / ms * 1e3}
    del th3, post3
    spec5, theta5 = workloads.make_c5(10_000)
    post5 = fit.from_spec(spec5)
    th5 = torch.as_tensor(theta5, device="cuda")
    t5 = torch.linspace(float(spec5["time"].min()), float(spec5["time"].max()), T, dtype=torch.float64, device="cuda")
    ms = time_kernel(torch, lambda: post5.ctx.gp_predict(th5, t5), 3, 1, barrier) / 3
    res["f4_gp_conditioning"] = {"shape": f"c5: 1e4 samples x {len(spec5['time'])} epochs -> {T} test times", "ms": ms,
                                 "samples_per_s": 1e4 / ms * 1e3}
    return res


def other_workloads(torch, fit, barrier, name, peak_flops) -> dict:
    """The other BASELINE configs on one GPU (kernel time, roofline fractions) and - oracle as checker, bounded
    subsample - the measured deviation of their results from the C restatement."""
    from oracle import oracle_c
    others = {}
    for other in ("c1", "c2", "c4", "c5"):
        if other == name:
            continue
        try:
            s2, t2 = make_workload(other, WORKLOADS[other]["samples"])
            p2 = fit.from_spec(s2)
            th2 = torch.as_tensor(t2, device="cuda")
            o2 = torch.empty(len(t2), dtype=torch.float64, device="cuda")
            m2 = time_kernel(torch, lambda: p2.ctx.logprob(th2, out=o2), 5, 3, barrier)
            u2 = units_per_step(s2, len(t2)) * 5 / (m2 * 1e-3)
            entry = {"evals_per_s": u2, "logprob_per_s": len(t2) * 5 / (m2 * 1e-3), "ms_per_step": m2 / 5,
                     "samples": len(t2)}
            hw = load_profile_json("hw", other)
            if other in FLOPS_PER_UNIT:
                entry["algorithmic_frac"] = u2 * FLOPS_PER_UNIT[other] / peak_flops
                if hw and "fp64_pipe_instructions_per_32_units" in hw:
                    entry["roofline_frac"] = u2 * hw["fp64_pipe_instructions_per_32_units"] * 2.0 / peak_flops
                    entry["roofline_source"] = hw["file"]
            elif other == "c5":
                # ~1.36e6 algorithmic fp64 FLOPs per GP log-prob at N = 120, one planet
                # (mean model 4.9e4 + covariance build 7.1e5 + Cholesky N^3/3 5.8e5 + solve / log-det 2e4)
                entry["flops_per_logprob"] = 1.36e6
                entry["roofline_frac"] = entry["logprob_per_s"] * 1.36e6 / peak_flops
                entry["kernel"] = ("rvlp::gps_factor_kernel (+ gpb_prologue_kernel): one 4-warp CTA per sample, factor in "
                                   "shared memory as 8x8 fragment-order tiles, DMMA Gram sums and solves (rvlp_gp_smem.cuh)")
                if hw:
                    entry["hw"] = hw
            n_chk = 2000 if other != "c5" else 300
            ref = oracle_c.OracleProblem(s2).logprob(t2[:n_chk], nthreads=host_cores())
            got = o2[:n_chk].cpu().numpy()
            fin = np.isfinite(ref)
            entry["parity_vs_c_oracle"] = parity_stats(got, ref)
            others[other] = entry
            del p2, th2, o2
        except Exception as ex:       # report, never hide
            others[other] = {"error": repr(ex)}
    return others


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=list(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the workload's samples in TOTAL, sharded over the GPUs (BASELINE config 3 as "
                         "named); weak: that many samples PER GPU")
    ap.add_argument("--samples", type=int, default=0, help="samples (total if strong, per GPU if weak); default: the workload's")
    ap.add_argument("--ref-samples", type=int, default=0, help="--impl reference: rows per step (default 1250 per host core)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads / cpu baseline")
    ap.add_argument("--no-clocks", action="store_true", help="experiment: do not run the nvidia-smi sampler during the timed region")
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

    def max_over_ranks(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    def agree(v: int) -> int:
        return int(max_over_ranks(float(v))[0])

    ravest_b200.load()
    name = args.workload
    S = args.samples or WORKLOADS[name]["samples"]
    peak_flops, _ = _lib.measure_fp64_peak(local, 4096)     # roofline denominator, measured before the timed region

    def measure(scaling: str, steps: int, warmup: int, with_clocks: bool):
        """One scaling mode: device-resident `value`, host-buffer `e2e`, cross-rank bit check."""
        if scaling == "strong":
            spec, theta_all = make_workload(name, S, 0)                 # every rank: the rank-0-seed rows
            lo, hi = rdist.shard_bounds(S, world, rank)
            theta = np.ascontiguousarray(theta_all[lo:hi])
            total = S
        else:
            spec, theta = make_workload(name, S, rank)
            theta_all, lo, hi, total = theta, 0, S, S * world
        post = fit.from_spec(spec)
        ctx = post.ctx
        th = torch.as_tensor(theta, device="cuda")
        part = torch.empty(hi - lo, dtype=torch.float64, device="cuda")
        units = units_per_step(spec, total)
        gathered = [None]

        if scaling == "strong":
            def step():            # the product's multi-GPU entry: this rank's block, gather fused into the kernel
                gathered[0] = rdist.sharded_logprob(lambda t: ctx.logprob(t, out=part), th, n_samples=S,
                                                    theta_is_local=True, ctx=ctx, copy=False)

            def step_nccl():       # the same with ONE NCCL all-gather after the kernel (fused=False / no CUDA IPC)
                gathered[0] = rdist.sharded_logprob(lambda t: ctx.logprob(t, out=part), th, n_samples=S,
                                                    theta_is_local=True, fused=False)
        else:
            recv = torch.empty(S * world, dtype=torch.float64, device="cuda") if world > 1 else None

            def step():
                ctx.logprob(th, out=part)
                if world > 1:
                    dist.all_gather_into_tensor(recv, part)
                gathered[0] = recv if world > 1 else part

        ctx.logprob(th, out=part)                                   # first call autotunes (synchronous), outside any timing
        # timing rule: inputs larger than L2, or L2 flushed between the timed iterations
        flush = None
        if theta.nbytes <= L2_BYTES * 1.5:
            flush = torch.zeros(2 * L2_BYTES, dtype=torch.uint8, device="cuda")
        # In a multi-GPU process the FIRST timed region reads 0.3-0.5 ms per step high, whichever gather path runs in
        # it and however long the warm-up (profiles/r02at_bench_n8_order.json: NCCL first 3.25, again 2.83; fused first
        # 3.06, later 2.80).  The protocol (W warm-ups, K timed steps) therefore runs twice; the second pass is reported,
        # the first is kept beside it as `first_pass_ms_per_step`.
        first_pass = time_kernel(torch, step, steps, warmup, barrier, flush, agree) if world > 1 else None
        sampler = ClockSampler(local) if (rank == 0 and with_clocks and not args.no_clocks) else None
        if sampler:
            sampler.start()
        ms = time_kernel(torch, step, steps, warmup, barrier, flush, agree)
        launches = time_kernel.launches
        clocks = sampler.stop() if sampler else None
        ms_kernel = time_kernel(torch, lambda: ctx.logprob(th, out=part), steps, 1, barrier, flush)
        gather = None
        if scaling == "strong" and world > 1:
            pgs = [g for g in getattr(ctx, "_peer_gathers", {}).values()]
            fused = bool(pgs) and all(g is not None for g in pgs)
            ms_nccl = time_kernel(torch, step_nccl, steps, warmup, barrier, flush, agree)
            nccl_out = gathered[0].clone()
            # order check: the two paths once more, interleaved
            ms_again = [time_kernel(torch, f, steps, warmup, barrier, flush, agree) / steps for f in (step, step_nccl)]
            step()                                                   # leave the default path's result in gathered[0]
            timed_out = any(g.timed_out() for g in pgs if g is not None)
            gather = {"how": ("fused into the kernel: K1 stores each batch of log-probs into every rank's vector (CUDA IPC "
                              "peer mappings, NVLink) + a one-warp flag barrier, no collective launch") if fused else
                             "one NCCL all_gather_into_tensor (CUDA IPC refused)",
                      "barrier_timed_out": timed_out,
                      "nccl_all_gather": {"ms_per_step": ms_nccl / steps,
                                          "bit_identical_to_fused_path": bool(torch.equal(nccl_out.view(torch.int64),
                                                                                          gathered[0].view(torch.int64))),
                                          "how": "sharded_logprob(fused=False): one all_gather_into_tensor after the kernel"},
                      "second_pass_ms_per_step": {"fused": ms_again[0], "nccl": ms_again[1]}}
        l2_note = (f"theta is {theta.nbytes / 1e6:.0f} MB per GPU (> 126 MB L2), streamed once per step" if flush is None else
                   f"theta is {theta.nbytes / 1e6:.0f} MB per GPU: L2 flushed between the timed iterations (a {2 * L2_BYTES >> 20} MB "
                   f"buffer rewritten outside the per-iteration event pairs)")

        # cross-GPU bit check: rank 0 evaluates ALL rows on its own GPU and compares with what the ranks gathered
        bit_identical = None
        if scaling == "strong" and world > 1:
            full = gathered[0].cpu().numpy()
            if rank == 0:
                own = ctx.logprob(torch.as_tensor(theta_all, device="cuda")).cpu().numpy()
                bit_identical = bool(np.array_equal(own.view(np.int64), full.view(np.int64)))

        # e2e: host buffers in and out through the public API (H2D + kernel + D2H per step, one sync per call)
        host_out = np.empty(hi - lo)
        pageable = np.array(theta, copy=True)                      # what a NumPy caller holds
        e2e_ms = time_host(torch, lambda: post.log_probability_batch(pageable), steps, 2, barrier)
        same = np.array_equal(post.log_probability_batch(pageable).view(np.int64), part.cpu().numpy().view(np.int64))
        pinned = torch.as_tensor(theta).pin_memory().numpy()
        e2e_pinned_ms = time_host(torch, lambda: ctx.logprob_host(pinned, host_out), steps, 2, barrier)
        ms, ms_kernel, e2e_ms, e2e_pinned_ms = max_over_ranks(ms, ms_kernel, e2e_ms, e2e_pinned_ms)
        if first_pass is not None:
            first_pass = max_over_ranks(first_pass)[0] / steps
        if gather:
            g2 = gather["second_pass_ms_per_step"]
            gather["nccl_all_gather"]["ms_per_step"], g2["fused"], g2["nccl"] = max_over_ranks(
                gather["nccl_all_gather"]["ms_per_step"], g2["fused"], g2["nccl"])
        return dict(spec=spec, theta=theta, theta_all=theta_all, post=post, units=units, total=total, ms=ms,
                    ms_kernel=ms_kernel, e2e_ms=e2e_ms, e2e_pinned_ms=e2e_pinned_ms, launches=launches, clocks=clocks,
                    bit_identical=bit_identical, same=bool(same), rows_local=hi - lo, steps=steps, warmup=warmup,
                    part=part, l2_note=l2_note, gather=gather, first_pass=first_pass)

    primary = measure(args.scaling, args.steps, args.warmup, True)
    other_mode = None
    if world > 1:        # the other scaling curve, fewer steps, same run
        other_mode = measure("weak" if args.scaling == "strong" else "strong", max(5, args.steps // 2), 3, False)

    if rank == 0:
        m = primary
        spec, theta = m["spec"], m["theta"]
        steps = m["steps"]
        value = m["units"] * steps / (m["ms"] * 1e-3)
        fpu = FLOPS_PER_UNIT.get(name)
        local_units = units_per_step(spec, m["rows_local"])
        per_gpu_units_per_s = local_units * steps / (m["ms_kernel"] * 1e-3)
        roofline = None
        if fpu:
            hw = load_profile_json("hw", name)
            traffic = None
            tj = load_profile_json("traffic", name)
            if tj and tj.get("samples") == m["rows_local"]:
                traffic = tj["dram_bytes_per_launch"]
            alg = per_gpu_units_per_s * fpu / 1e12
            roofline = {"bound": "fp64", "peak": peak_flops / 1e12, "unit": "TFLOP/s", "traffic": traffic,
                        "peak_source": "measured live: dependent-free DFMA kernel (rvlp_measure_fp64_peak), burst",
                        "algorithmic_achieved": alg, "algorithmic_frac": alg / (peak_flops / 1e12), "flops_per_unit": fpu,
                        "kernel": "rvlp::logprob_kernel", "kernel_ms_per_launch": m["ms_kernel"] / steps,
                        "algorithmic_hbm_bytes_per_launch": m["rows_local"] * (theta.shape[1] + 1) * 8, "hw": hw,
                        "note": "frac = fp64-pipe instructions the kernel EXECUTES per unit (committed ncu capture) x units/s "
                                "x 2 FLOP, over the measured DFMA peak - the hardware fraction.  algorithmic_frac counts the "
                                "REFERENCE algorithm's 418 FLOP/unit (4 Halley passes with a libm sincos each) and exceeds 1 "
                                "because this kernel needs ~3x fewer fp64 operations (fp32/MUFU starter + one fp64 step)."}
            if hw and "fp64_pipe_instructions_per_32_units" in hw:
                # "per 32 units" = warp instructions per 32 lanes = thread-level instructions per unit
                ach = per_gpu_units_per_s * hw["fp64_pipe_instructions_per_32_units"] * 2.0 / 1e12
                roofline["achieved"] = ach
                roofline["frac"] = ach / (peak_flops / 1e12)
            else:
                roofline["achieved"] = alg
                roofline["frac"] = alg / (peak_flops / 1e12)
                roofline["note"] += "  (no committed ncu capture found: frac falls back to the algorithmic figure)"
        par = (f"{m['total']} samples sharded x{world} (ravest_b200.dist.sharded_logprob), the all-gather of log-probs fused "
               f"into the kernel (see `gather`)"
               if args.scaling == "strong" else f"{S} samples per GPU x{world}, 1 all-gather of log-probs per step")
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": steps, "warmup": m["warmup"], "ms_per_step": m["ms"] / steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {WORKLOADS[name]['desc']}", "samples_total": m["total"],
                       "samples_per_gpu": m["rows_local"], "epochs": len(spec["time"]),
                       "planets": len(spec["planet_letters"]), "ndim": int(theta.shape[1]),
                       "parallelism": par if world > 1 else "single GPU",
                       "l2": m["l2_note"]},
            "logprob_per_s": m["total"] * steps / (m["ms"] * 1e-3),
            "e2e": {"value": m["units"] * steps / (m["e2e_ms"] * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": int(theta.nbytes) * world if args.scaling == "weak" else int(m["theta_all"].nbytes),
                    "d2h_bytes_per_step": int(m["total"] * 8), "ms_per_step": m["e2e_ms"] / steps,
                    "host_memory": "pageable NumPy array (staged through the context's pinned buffers, chunked H2D || kernel || D2H)",
                    "api": "LogPosterior.log_probability_batch(numpy) -> rvlp_logprob_batch_host",
                    "bit_identical_to_device_path": m["same"],
                    "pinned": {"value": m["units"] * steps / (m["e2e_pinned_ms"] * 1e-3), "ms_per_step": m["e2e_pinned_ms"] / steps}},
            "gpu_launches": int(m["launches"]),
            "clocks": m["clocks"],
            "roofline": roofline,
        }
        if m["bit_identical"] is not None:
            line["bit_identical_across_ranks"] = m["bit_identical"]
            if m.get("gather"):
                line["gather"] = m["gather"]
            if m.get("first_pass") is not None:
                line["first_pass_ms_per_step"] = m["first_pass"]
        if other_mode is not None:
            o = other_mode
            key = "weak_scaling" if args.scaling == "strong" else "strong_scaling"
            line[key] = {"value": o["units"] * o["steps"] / (o["ms"] * 1e-3), "ms_per_step": o["ms"] / o["steps"],
                         "samples_total": o["total"], "samples_per_gpu": o["rows_local"], "steps": o["steps"],
                         "e2e_value": o["units"] * o["steps"] / (o["e2e_ms"] * 1e-3)}
            if o["bit_identical"] is not None:
                line["bit_identical_across_ranks"] = o["bit_identical"]
        if world == 1 and not args.no_extras:
            try:
                line["cpu_baseline"] = cpu_baseline(spec, theta, m["part"].cpu().numpy())
            except Exception as ex:           # report, never hide
                line["cpu_baseline"] = {"error": repr(ex)}
            line["other_workloads"] = other_workloads(torch, fit, barrier, name, peak_flops)
            try:
                line["sample_matrix_rows"] = sample_matrix_rows(torch, fit, barrier)
            except Exception as ex:           # report, never hide
                line["sample_matrix_rows"] = {"error": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
