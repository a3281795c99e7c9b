import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# north_star tolerances (BASELINE.json): RV within 1e-9 relative to max(|rv|, K),
# log-probability within 1e-7 absolute.  Rows whose log-probability is astronomically large
# (|logp| up to 6e11 for the deliberately absurd "far" rows of the fixtures) cannot be held to an
# absolute 1e-7 by ANY summation order - ulp(6e11) is 1.2e-4 - so a relative floor of 2e-13
# (a few ulps per accumulated term) is added and documented here.
LOGP_ATOL = 1e-7
LOGP_RTOL = 2e-13
RV_RTOL = 1e-9


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def spec_from_json(spec):
    s = dict(spec)
    s["params"] = {k: (v[0], bool(v[1])) for k, v in spec["params"].items()}
    s["priors"] = {k: tuple(v) for k, v in spec["priors"].items()}
    for k in ("time", "vel", "velerr"):
        s[k] = np.asarray(spec[k], dtype=np.float64)
    s["instrument"] = np.asarray(spec["instrument"])
    return s


def assert_logp_close(got, ref, what=""):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.array_equal(np.isneginf(got), np.isneginf(ref)), f"{what}: -inf pattern differs"
    assert np.array_equal(np.isnan(got), np.isnan(ref)), f"{what}: NaN pattern differs"
    assert np.array_equal(np.isposinf(got), np.isposinf(ref)), f"{what}: +inf pattern differs"
    fin = np.isfinite(ref)
    if fin.any():
        err = np.abs(got[fin] - ref[fin])
        tol = LOGP_ATOL + LOGP_RTOL * np.abs(ref[fin])
        worst = int(np.argmax(err - tol))
        assert np.all(err <= tol), f"{what}: |d|={err[worst]:.3e} at ref={ref[fin][worst]:.6e} (tol {tol[worst]:.1e})"


def assert_rv_close(got, ref, K, what="", rtol=RV_RTOL):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = np.maximum(np.abs(ref), abs(K))
    err = np.abs(got - ref) / scale
    assert np.all(err <= rtol), f"{what}: max rel err {err.max():.3e} > {rtol}"


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test run without a CUDA device")
    import ravest_b200
    ravest_b200.load()      # fails loudly if the extension is missing
    return torch
