"""The drop-in binding on the REFERENCE's own objects (CPU part): a `ravest.fit.LogPosterior` built by the
unmodified reference compiles to the same POD descriptor as the fixture spec it came from, and the ctypes stub of
INTEGRATION.md §2 is the text that runs.  Skipped where the reference is absent (neither /root/reference nor
oracle/_ref)."""
import ctypes as C
import importlib.util
import os
import re
import types

import numpy as np
import pytest

from conftest import ROOT, load_golden, spec_from_json
from oracle.ref_import import reference_available
from ravest_b200 import adapter, prior as P, workloads
from ravest_b200.descriptor import Descriptor, DescPOD, PriorPOD

needs_ref = pytest.mark.skipif(not reference_available(), reason="reference not present (no /root/reference, no oracle/_ref)")


def pod_fingerprint(d: Descriptor):
    pod = d.pod
    n = len(d.src_col)
    rows = [(r.kind, r.target, r.index, r.is_hyper, tuple(r.p), tuple(r.c)) for r in d._prior_array[: d.n_priors]]
    return dict(hdr=(pod.abi_version, pod.n_planets, pod.parameterisation, pod.n_inst, pod.ndim, pod.n_priors,
                     pod.n_hyper, pod.t0, pod.jacobian, pod.renorm),
                src_col=[pod.src_col[i] for i in range(n)], src_const=[pod.src_const[i] for i in range(n)],
                priors=rows, columns=list(d.columns), model=list(d.model_names))


@needs_ref
@pytest.mark.parametrize("fixture", ["logprob_cases", "known_answers"])
def test_reference_logposterior_compiles_to_the_fixture_descriptor(fixture):
    from helpers import ref_logposterior
    n = 0
    for case in load_golden(fixture):
        spec = spec_from_json(case["spec"])
        lp = ref_logposterior(spec)                        # ravest.fit.LogPosterior with ravest.prior.* objects
        assert type(lp).__module__ == "ravest.fit"
        assert all(type(p).__module__ == "ravest.prior" for p in lp.priors.values())
        pod, keep = adapter.compile_descriptor(lp)
        assert isinstance(pod, DescPOD) and C.sizeof(pod) == 80
        assert pod_fingerprint(keep) == pod_fingerprint(Descriptor.from_spec(spec)), case["name"]
        post = adapter.from_reference(lp)
        assert post._logprob_jacobian_correction == lp._logprob_jacobian_correction == case.get("jacobian", lp._logprob_jacobian_correction)
        assert post._logprob_prior_renorm_correction == lp._logprob_prior_renorm_correction
        assert post.time is not None and np.array_equal(post.time, lp.time)
        n += 1
    assert n >= 3


@needs_ref
def test_reference_logposterior_via_fitter_and_all_workloads():
    from helpers import ref_logposterior
    for maker, kw in (("make_c1", {}), ("make_c2", {}), ("make_c3", {}), ("make_c4", {})):
        spec, theta = getattr(workloads, maker)(8, **kw)
        lp = ref_logposterior(spec, via_fitter=True)
        _, keep = adapter.compile_descriptor(lp)
        assert pod_fingerprint(keep) == pod_fingerprint(Descriptor.from_spec(spec)), maker
        assert keep.columns == workloads.free_names(spec)


def test_prior_duck_typing_by_class_name_and_attributes():
    # stand-ins with the reference's class names and attributes, no ravest import needed (runs everywhere)
    def mk(name, **kw):
        return type(name, (), {})().__class__ and _obj(name, kw)

    def _obj(name, kw):
        o = type(name, (), {})()
        o.__dict__.update(kw)
        return o
    cases = [(_obj("Uniform", dict(lower=-1.0, upper=2.0)), P.Uniform(-1.0, 2.0)),
             (_obj("EccentricityUniform", dict(upper=0.9)), P.EccentricityUniform(0.9)),
             (_obj("Normal", dict(mean=1.0, std=2.0)), P.Normal(1.0, 2.0)),
             (_obj("TruncatedNormal", dict(mean=0.1, std=0.2, lower=0.0, upper=1.0)), P.TruncatedNormal(0.1, 0.2, 0.0, 1.0)),
             (_obj("HalfNormal", dict(std=0.3)), P.HalfNormal(0.3)),
             (_obj("Rayleigh", dict(scale=0.26)), P.Rayleigh(0.26)),
             (_obj("VanEylen19Mixture", dict(sigma_normal=0.049, sigma_rayleigh=0.26, f=0.76)), P.VanEylen19Mixture(0.049, 0.26, 0.76)),
             (_obj("Beta", dict(a=0.867, b=3.03)), P.Beta(0.867, 3.03))]
    for duck, ours in cases:
        got = adapter.convert_prior(duck)
        assert type(got) is type(ours) and got.pod() == ours.pod()
    assert adapter.convert_prior(cases[0][1]) is cases[0][1]
    with pytest.raises(NotImplementedError):
        adapter.convert_prior(_obj("LogUniform", dict(lower=1, upper=2)))
    with pytest.raises(TypeError):
        adapter.convert_prior(_obj("Normal", dict(mean=1.0)))


def test_gp_posterior_duck_typed():
    spec, theta = workloads.make_c5(n_samples=4, n_planets=1, n_epochs=20, seed=3)
    ours = Descriptor.from_spec(spec)
    hp = spec["hyperparams"]
    lp = types.SimpleNamespace(
        planet_letters=list(spec["planet_letters"]), parameterisation=types.SimpleNamespace(parameterisation=spec["parameterisation"]),
        gp_kernel=types.SimpleNamespace(kernel_type="Quasiperiodic"),
        priors={k: P.from_tuple(v) for k, v in spec["priors"].items()},
        hyperpriors={k: P.from_tuple(v) for k, v in spec["hyperpriors"].items()},
        fixed_params={k: v for k, (v, fx) in spec["params"].items() if fx},
        fixed_hyperparams={k: v for k, (v, fx) in hp.items() if fx},
        free_params_names=[k for k, (_, fx) in spec["params"].items() if not fx],
        free_hyperparams_names=[k for k, (_, fx) in hp.items() if not fx],
        time=spec["time"], vel=spec["vel"], velerr=spec["velerr"], t0=spec["t0"],
        instrument=np.asarray(spec["instrument"]), unique_instruments=np.unique(np.asarray(spec["instrument"])))
    _, keep = adapter.compile_descriptor(lp)
    assert pod_fingerprint(keep) == pod_fingerprint(ours)
    f = adapter.BatchedLogPosterior(lp)
    assert f.parameter_names == keep.columns and f.ndim == theta.shape[1]


def test_integration_stub_is_the_documented_text_and_loads():
    """INTEGRATION.md §2 shows integration/_b200.py verbatim; the module imports (binding every symbol it names)
    without a GPU."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = open(os.path.join(ROOT, "integration", "_b200.py")).read()
    m = re.search(r"```python\n(# src/ravest/_b200\.py.*?)```", doc, re.S)
    assert m, "INTEGRATION.md lost its stub block"
    assert m.group(1).strip() == stub.strip(), "INTEGRATION.md §2 and integration/_b200.py differ"
    from ravest_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    os.environ["RVLP_LIB"] = _lib.LIB_PATH
    sp = importlib.util.spec_from_file_location("ravest_b200_stub", os.path.join(ROOT, "integration", "_b200.py"))
    mod = importlib.util.module_from_spec(sp)
    sp.loader.exec_module(mod)
    assert C.sizeof(mod._Desc) == C.sizeof(DescPOD) and C.sizeof(mod._Prior) == C.sizeof(PriorPOD) == 64
    assert [f[0] for f in mod._Desc._fields_] == [f[0] for f in DescPOD._fields_]
    assert callable(mod.BatchedLogPosterior)


def test_stretch_move_helper_samples_a_gaussian():
    """The in-repo emcee stand-in (tests/helpers.py) is a correct ensemble sampler: unit Gaussian moments."""
    from helpers import stretch_move_run
    rng = np.random.default_rng(5)
    p0 = rng.normal(size=(40, 3))
    chain, logp, acc, calls = stretch_move_run(lambda x: -0.5 * (x ** 2).sum(axis=1), p0, 1500, seed=9)
    flat = chain[300:].reshape(-1, 3)
    assert calls == 3000 and 0.2 < acc / (1500 * 40) < 0.8
    assert np.all(np.abs(flat.mean(axis=0)) < 0.15) and np.all(np.abs(flat.var(axis=0) - 1.0) < 0.2)
