"""CPU-only checks of the host side: descriptor compilation, the C ABI's symbol table, the
solver's convergence (compiled for the host), sharding logic and a world_size-2 gloo run."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, spec_from_json
from ravest_b200 import _lib, dist, prior as P, workloads
from ravest_b200.descriptor import Descriptor


def test_library_loads_and_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    header = open(os.path.join(ROOT, "include", "ravest_b200.h")).read()
    declared = set(re.findall(r"\b(rvlp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rvlp_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf" T {name}\b", out), name
    sass = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ravest_b200 import fit
    spec, theta = workloads.make_c1(8)
    post = fit.from_spec(spec)
    with pytest.raises(_lib.RvlpError):
        post.log_probability_batch(theta)
    with pytest.raises(_lib.RvlpError):
        P.Uniform(0, 1)(0.5)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ravest_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "oracle/" not in src and "liboracle" not in src, f


def test_descriptor_source_table_and_prior_rows():
    c = [x for x in load_golden("logprob_cases") if x["name"] == "2pl case3 + Tp prior"][0]
    spec = spec_from_json(c["spec"])
    d = Descriptor.from_spec(spec)
    assert d.free_params_names == c["free_names"]
    assert d.model_names[:5] == ["P_b", "K_b", "secosw_b", "sesinw_b", "Tc_b"]
    assert d.model_names[10:] == ["gd", "gdd", "g_HARPS", "jit_HARPS"]
    for i, n in enumerate(d.model_names):
        if n in d.free_params_names:
            assert d.src_col[i] == d.free_params_names.index(n)
        else:
            assert d.src_col[i] == -1 and d.src_const[i] == spec["params"][n][0]
    # priors on e, w, Tp are evaluated on converted values, in the reference's dict order (fit.py:3426-3446)
    inv = {v: k for k, v in {"P": 1, "K": 2, "e": 3, "w": 4, "Tp": 5}.items()}
    names = []
    for r in d._prior_array[: d.n_priors]:
        names.append(d.free_params_names[r.index] if r.target == 0
                     else f"{inv[r.target]}_{d.planet_letters[r.index]}")
    direct = [n for n in d.free_params_names if n in spec["priors"]]
    expect = direct + [f"{q}_{L}" for L in "bc" for q in ("e", "w", "Tp")]
    assert names == expect
    # P_b / K_b keep their dict position but read the CONVERTED value (same number), as fit.py:3441-3444 does
    assert [(r.target, r.index) for r in d._prior_array[:2]] == [(1, 0), (2, 0)]
    assert abs(d.pod.jacobian - 2 * np.log(2)) < 1e-15 and d.pod.renorm == 0.0   # two CASE_3 planets


def test_logprob_correction_cases():
    from ravest_b200.fit import compute_logprob_corrections
    U = P.Uniform(-1, 1)
    free = ["P_b", "K_b", "secosw_b", "sesinw_b", "Tc_b"]
    j, r, b = compute_logprob_corrections(["b"], "P K secosw sesinw Tc", {"secosw_b": U, "sesinw_b": U}, free)
    assert (j, b["b"]["case"]) == (0.0, "CASE_2") and abs(r - np.log(4 / np.pi)) < 1e-16
    j, r, b = compute_logprob_corrections(["b"], "P K secosw sesinw Tc",
                                          {"e_b": P.Beta(1, 3), "w_b": P.Uniform(-np.pi, np.pi)}, free)
    assert (r, b["b"]["case"]) == (0.0, "CASE_3") and abs(j - np.log(2)) < 1e-16
    j, r, b = compute_logprob_corrections(["b"], "P K e w Tc", {}, ["P_b"])
    assert (j, r, b["b"]["case"]) == (0.0, 0.0, "CASE_1")
    j, r, b = compute_logprob_corrections(["b"], "P K secosw sesinw Tc", {}, ["P_b"])   # (u, v) fixed
    assert b["b"]["case"] == "CASE_1"
    with pytest.raises(NotImplementedError):
        compute_logprob_corrections(["b"], "P K secosw sesinw Tc",
                                    {"secosw_b": P.Uniform(-0.5, 0.5), "sesinw_b": U}, free)
    with pytest.raises(RuntimeError):
        compute_logprob_corrections(["b"], "P K secosw sesinw Tc", {"P_b": U}, free)


def test_prior_constructor_errors_match_reference():
    for bad in (lambda: P.Uniform(1, 1), lambda: P.Uniform(np.inf, 2), lambda: P.EccentricityUniform(1.5),
                lambda: P.EccentricityUniform(0), lambda: P.Normal(0, 0), lambda: P.TruncatedNormal(0, 1, 2, 1),
                lambda: P.HalfNormal(-1), lambda: P.Rayleigh(0), lambda: P.VanEylen19Mixture(1, 1, 1.5),
                lambda: P.Beta(0, 1)):
        with pytest.raises(ValueError):
            bad()
    assert repr(P.Uniform(0, 1)) == "Uniform(lower=0, upper=1)"
    assert P.PRIOR_FUNCTIONS[3] == "TruncatedNormal"


@pytest.mark.parametrize("mufu", [False, True])
def test_host_compiled_solver_converges(mufu):
    """The solver (rvlp_math.cuh compiled for the host) over a dense (e, M) sweep against an 80-bit
    reference; mufu=True injects the worst-case MUFU.SIN/COS error into the fp32 stage."""
    exe = "/tmp/rvlp_solver_check" + ("_mufu" if mufu else "")
    src = os.path.join(ROOT, "tests", "host", "solver_check.cpp")
    flags = ["-DRVLP_EMULATE_MUFU"] if mufu else []
    subprocess.run(["g++", "-O2", "-std=c++17", "-mfma"] + flags + ["-o", exe, src, "-lm"], check=True)
    res = subprocess.run([exe, "60000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    lines = res.stdout.strip().splitlines()
    assert len(lines) >= 20
    for ln in lines:
        fb = float(re.search(r"fallback=([0-9.]+)%", ln).group(1))
        e = float(re.search(r"e=(\S+)", ln).group(1))
        ratio = float(re.search(r"rv_err/cond=([0-9.]+)", ln).group(1))
        assert ratio < 4.0, ln            # error in units of the problem's own conditioning
        if e <= 0.995:
            assert fb == 0.0, ln          # the safety net is never needed where the fast plans apply


def test_host_compiled_gp_covariance_is_accurate():
    """The branch-free quasi-periodic covariance (rvlp_gpcov.cuh compiled for the host, contraction off so the
    algorithm is tested as written) against a long-double evaluation of gp.py:145-156, plus edge inputs."""
    exe = "/tmp/rvlp_gpcov_check"
    src = os.path.join(ROOT, "tests", "host", "gpcov_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    res = subprocess.run([exe, "400000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert res.stdout.strip().endswith("OK"), res.stdout


def test_host_compiled_fast_log_is_accurate():
    """log_pos_normal (the one log per sample and lane that closes the chi^2 / log-det reduction) against logl."""
    exe = "/tmp/rvlp_log_check"
    src = os.path.join(ROOT, "tests", "host", "log_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-o", exe, src, "-lm"], check=True)
    res = subprocess.run([exe, "1000000"], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK"), res.stdout


def test_shard_bounds_cover_and_align():
    for S in (0, 1, 3, 4, 5, 31, 32, 1000, 100_003):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = dist.shard_bounds(S, world, r)
                assert lo == prev and lo <= hi <= S
                if hi < S:
                    assert hi % dist.ALIGN == 0
                prev = hi
            assert prev == S


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as td
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_c
    spec, theta = workloads.make_multiplanet(2, 40, 101, seed=9, invalid_frac=0.05)
    orc = oracle_c.OracleProblem(spec)
    th = torch.as_tensor(theta)

    def eval_fn(block):                      # CPU stand-in for the CUDA evaluator: tests the plumbing only
        return torch.as_tensor(orc.logprob(block.numpy(), nthreads=1))

    full = dist.sharded_logprob(eval_fn, th)
    lo, hi = dist.shard_bounds(len(theta), world, rank)
    local = dist.sharded_logprob(eval_fn, th[lo:hi], n_samples=len(theta), theta_is_local=True)
    # a `ctx` on CPU tensors / gloo must not take the fused (CUDA IPC + NCCL only) path, whatever the default says
    again = dist.sharded_logprob(eval_fn, th, ctx=object(), fused=True)
    assert torch.equal(again, full) or bool(torch.isnan(full).any())
    q.put((rank, full.numpy(), local.numpy()))
    td.destroy_process_group()


def test_sharded_logprob_gloo_world2():
    import torch.multiprocessing as mp
    from oracle import oracle_c
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec, theta = workloads.make_multiplanet(2, 40, 101, seed=9, invalid_frac=0.05)
    ref = oracle_c.OracleProblem(spec).logprob(theta, nthreads=1)
    for rank, full, local in got:
        assert np.array_equal(full, ref, equal_nan=True)      # bit-identical regardless of the split
        assert np.array_equal(local, ref, equal_nan=True)


def test_workload_shapes():
    spec, theta = workloads.make_c3(n_samples=64, n_epochs=1000)
    assert theta.shape == (64, 29) and len(spec["time"]) == 1000
    spec, theta = workloads.make_c4(n_samples=64)
    assert theta.shape == (64, 21) and len(np.unique(spec["instrument"])) == 2
    spec, theta = workloads.make_c2(n_samples=16)
    assert theta.shape == (16, 12) and len(spec["time"]) == 120
    spec, theta = workloads.make_c5(n_samples=16)
    assert theta.shape[1] == 7 + 4
    a = workloads.make_c3(n_samples=32)[1]
    b = workloads.make_c3(n_samples=32)[1]
    assert np.array_equal(a, b)
