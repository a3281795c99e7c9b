"""The drop-in, shown on the reference's OWN objects on the GPU: an unmodified `ravest.fit.LogPosterior` (imported
from oracle/_ref on the GPU box, /root/reference in the build container) goes through (i) the ctypes stub of
INTEGRATION.md §2 verbatim and (ii) `ravest_b200.adapter.BatchedLogPosterior` inside a stretch-move ensemble loop that
calls it as emcee does with vectorize=True; results are compared with the reference's per-walker dict path
(`lp.log_probability`, fit.py:3448-3495) evaluated live in the same process."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT, assert_logp_close, load_golden, spec_from_json
from oracle.ref_import import reference_available

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not reference_available(), reason="oracle/_ref did not travel (run __graft_entry__.build() in the build container)")


def _stub_module():
    from ravest_b200 import _lib
    os.environ["RVLP_LIB"] = _lib.LIB_PATH
    sp = importlib.util.spec_from_file_location("ravest_b200_stub", os.path.join(ROOT, "integration", "_b200.py"))
    mod = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(mod)
    return mod


@needs_ref
def test_integration_stub_verbatim_on_reference_objects(cuda):
    from helpers import ref_logposterior, ref_logprob_rows
    stub = _stub_module()
    n = 0
    for case in load_golden("logprob_cases") + load_golden("known_answers"):
        spec = spec_from_json(case["spec"])
        lp = ref_logposterior(spec)
        f = stub.BatchedLogPosterior(lp)                    # the documented stub, unchanged
        theta = np.asarray(case["theta"], dtype=np.float64)
        got = f(theta)
        assert isinstance(got, np.ndarray) and got.shape == (len(theta),)
        live = ref_logprob_rows(lp, lp.free_params_names, theta)      # the reference, evaluated here and now
        assert_logp_close(got, live, f"stub vs live reference: {case['name']}")
        if "logprob" in case:
            assert np.array_equal(live, np.asarray(case["logprob"], dtype=np.float64), equal_nan=True), \
                "the live reference no longer reproduces the committed fixture"
        n += len(theta)
    assert n > 500


@needs_ref
@pytest.mark.parametrize("maker,nwalk,steps", [("make_c1", 32, 60), ("make_c2", 64, 40)])
def test_stretch_move_ensemble_vectorized_matches_reference_dict_path(cuda, maker, nwalk, steps):
    """emcee's vectorize=True contract: one call per half-step with the (nwalkers/2, ndim) proposal block.  The same
    seeded ensemble is advanced twice - log-probabilities from the reference's dict path, and from the CUDA path - and
    must make the same accept/reject decisions (the two differ by ~1e-10, the acceptance test by O(1))."""
    from helpers import ref_logposterior, ref_logprob_rows, stretch_move_run
    from ravest_b200 import adapter, workloads
    import ravest_b200
    spec, theta = getattr(workloads, maker)(4096)
    lp = ref_logposterior(spec, via_fitter=True)
    names = list(lp.free_params_names)
    f = adapter.BatchedLogPosterior(lp)
    assert f.parameter_names == names
    # walkers: a tight ball around the best of the workload's rows (finite log-prob by construction)
    base = f(theta)
    best = theta[int(np.nanargmax(np.where(np.isfinite(base), base, -np.inf)))]
    rng = np.random.default_rng(11)
    p0 = best + 1e-4 * np.abs(best) * rng.normal(size=(nwalk, len(best)))
    assert np.isfinite(f(p0)).all()
    n0 = ravest_b200.launch_count()
    chain_g, logp_g, acc_g, calls = stretch_move_run(f, p0, steps, seed=3)
    launches = ravest_b200.launch_count() - n0
    chain_r, logp_r, acc_r, _ = stretch_move_run(lambda x: ref_logprob_rows(lp, names, x), p0, steps, seed=3)
    assert calls == 2 * steps and launches >= calls
    assert acc_g == acc_r and 0.05 < acc_g / (steps * nwalk) < 0.95
    assert np.array_equal(chain_g, chain_r), "the ensembles diverged: an accept/reject decision differed"
    assert_logp_close(logp_g.ravel(), logp_r.ravel(), "chain log-probabilities")


@needs_ref
def test_adapter_scalar_conventions_on_reference_object(cuda):
    from helpers import ref_logposterior
    from ravest_b200 import adapter, workloads
    spec, theta = workloads.make_c2(64)
    lp = ref_logposterior(spec)
    f = adapter.BatchedLogPosterior(lp)
    names = list(lp.free_params_names)
    for row in theta[:8]:
        d = dict(zip(names, map(float, row)))
        a, b = f.log_probability(d), lp.log_probability(d)
        assert (a == b) or abs(a - b) <= 1e-7 + 2e-13 * abs(b)
        na, nb = f._negative_log_probability_for_MAP(list(row)), lp._negative_log_probability_for_MAP(list(row))
        assert (na == nb) or abs(na - nb) <= 1e-7 + 2e-13 * abs(nb)
    import torch
    dev = f(torch.as_tensor(theta, device="cuda"))
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), f(theta), equal_nan=True)
    # the sample-matrix rows are reachable from the same object (INTEGRATION.md "Posterior plots")
    bands = f.rv_percentile_bands(np.linspace(spec["time"].min(), spec["time"].max(), 50), torch.as_tensor(theta, device="cuda"))
    assert tuple(bands.shape) == (3, 50)
