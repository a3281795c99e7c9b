"""Import the UNMODIFIED reference (ravest v0.4.0) from /root/reference for golden generation.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`ravest_b200/`), `bench.py`,
`__graft_entry__.smoke()` or the `-m gpu` tests imports this module: `/root/reference`
does not exist on the GPU box; the git-ignored installed copy `oracle/_ref/` (oracle/make_ref.py) does travel, and
is what `bench.py --impl reference` / the `cpu_baseline` leg and the live-reference tests import there.  It is used by `tests/golden/make_golden.py` (run in the
build container) to produce the committed fixtures, and by a handful of `not gpu` tests that
are skipped when `/root/reference` is absent.

The reference imports matplotlib / astropy / emcee / corner / jax / tinygp at module top
(`src/ravest/model.py:10-15`, `src/ravest/fit.py:19-30`) although none of them touches the
white-noise arithmetic; they are not installed here, so empty stand-ins are registered.
The GP arithmetic (tinygp/jax) is therefore NOT available through this import.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# 1. the read-only reference tree (build container only); 2. oracle/_ref, the git-ignored installed copy that
# oracle/make_ref.py (run by __graft_entry__.build()) makes so that the unmodified reference travels to the GPU box
CANDIDATES = ["/root/reference/src", os.path.join(_HERE, "_ref")]


def reference_src() -> str | None:
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "ravest", "fit.py")):
            return c
    return None


REFERENCE_SRC = reference_src() or CANDIDATES[0]


def reference_available() -> bool:
    return reference_src() is not None


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def import_reference():
    """Return the reference's (model, param, prior, fit) modules, imported unmodified."""
    if not reference_available():
        raise RuntimeError("the reference is not present on this machine (neither /root/reference nor oracle/_ref)")
    if "ravest.fit" in sys.modules and getattr(sys.modules["ravest"], "_is_reference", False):
        r = sys.modules
        return r["ravest.model"], r["ravest.param"], r["ravest.prior"], r["ravest.fit"]

    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_ravest_ref")

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, item):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "astropy",
                 "astropy.constants", "corner", "emcee", "jax", "jax.numpy", "tinygp",
                 "tinygp.kernels"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].ticker = sys.modules["matplotlib.ticker"]
    for cls in ("MultipleLocator", "AutoLocator", "AutoMinorLocator"):
        setattr(sys.modules["matplotlib.ticker"], cls, _Anything)
    sys.modules["astropy"].constants = sys.modules["astropy.constants"]
    jax = sys.modules["jax"]
    if not hasattr(jax, "config"):
        jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
        jax.jit = lambda f: f
        jax.Array = type("Array", (), {})      # scipy's array-api helpers probe sys.modules['jax'].Array
        jax.numpy = sys.modules["jax.numpy"]
        sys.modules["jax.numpy"].ndarray = object
    tinygp = sys.modules["tinygp"]
    if not hasattr(tinygp, "GaussianProcess"):
        tinygp.GaussianProcess = _Anything
        tinygp.kernels = sys.modules["tinygp.kernels"]
        sys.modules["tinygp.kernels"].Kernel = _Anything

    pkg = types.ModuleType("ravest")
    pkg.__path__ = [os.path.join(reference_src(), "ravest")]
    pkg._is_reference = True
    sys.modules["ravest"] = pkg
    model = importlib.import_module("ravest.model")
    param = importlib.import_module("ravest.param")
    prior = importlib.import_module("ravest.prior")
    fit = importlib.import_module("ravest.fit")
    return model, param, prior, fit
