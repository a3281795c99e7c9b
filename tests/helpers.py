"""Test helpers: reference-object builders (need the reference: /root/reference or oracle/_ref) and a minimal
stretch-move ensemble sampler that calls its log-probability function exactly as emcee does with vectorize=True.

emcee is not installed in this image.  Its public contract (emcee 3.1.6 `EnsembleSampler.compute_log_prob`,
`moves/red_blue.py`, `moves/stretch.py`): the ensemble is split in two halves; for each half, every walker k draws
z = ((a - 1) u + 1)^2 / a, proposes y = c_j + z (x_k - c_j) with c_j a random walker of the OTHER half, and the
whole (n/2, ndim) block of proposals goes through `log_prob_fn` in ONE call when vectorize=True; accept when
ln u' < (ndim - 1) ln z + logp(y) - logp(x_k).
"""
from __future__ import annotations

import numpy as np


def ref_prior(prior_mod, p):
    return getattr(prior_mod, p[0])(*p[1:])


def ref_logposterior(spec, via_fitter: bool = False):
    """A `ravest.fit.LogPosterior` of the unmodified reference for a workload / fixture spec."""
    from oracle.ref_import import import_reference
    model, param, prior, fit = import_reference()
    if via_fitter:                     # the path Fitter.run_mcmc takes (fit.py:1021-1033)
        f = fit.Fitter(list(spec["planet_letters"]), param.Parameterisation(spec["parameterisation"]))
        f.add_data(np.asarray(spec["time"], float), np.asarray(spec["vel"], float),
                   np.asarray(spec["velerr"], float), np.asarray(spec["instrument"]), spec["t0"])
        f.params = {k: param.Parameter(v, "", fixed=fx) for k, (v, fx) in spec["params"].items()}
        f.priors = {k: ref_prior(prior, p) for k, p in spec["priors"].items()}
        return fit.LogPosterior(f.planet_letters, f.parameterisation, f.priors, f.fixed_params_values_dict,
                                f.free_params_names, f.time, f.vel, f.velerr, f.instrument, f.unique_instruments, f.t0)
    params = spec["params"]
    free = [k for k, (_, fx) in params.items() if not fx]
    fixed = {k: v for k, (v, fx) in params.items() if fx}
    inst = np.asarray(spec["instrument"])
    return fit.LogPosterior(list(spec["planet_letters"]), param.Parameterisation(spec["parameterisation"]),
                            {k: ref_prior(prior, p) for k, p in spec["priors"].items()}, fixed, free,
                            np.asarray(spec["time"], float), np.asarray(spec["vel"], float),
                            np.asarray(spec["velerr"], float), inst, np.unique(inst), spec["t0"])


def ref_logprob_rows(lp, names, theta) -> np.ndarray:
    """The reference's own per-walker dict path (what emcee does with parameter_names=, fit.py:1070-1075)."""
    return np.array([float(lp.log_probability(dict(zip(names, (float(x) for x in row))))) for row in theta])


def stretch_move_run(log_prob_fn, p0: np.ndarray, nsteps: int, seed: int, a: float = 2.0):
    """Red-blue stretch move; `log_prob_fn(coords[n, ndim]) -> n` is called once per half-step (vectorize=True).
    Returns (chain[nsteps, nwalkers, ndim], logp[nsteps, nwalkers], n_accepted, n_calls)."""
    rng = np.random.default_rng(seed)
    x = np.array(p0, dtype=np.float64)
    nw, ndim = x.shape
    assert nw % 2 == 0 and nw >= 2 * ndim
    lp = np.asarray(log_prob_fn(x), dtype=np.float64)
    assert lp.shape == (nw,)
    chain = np.empty((nsteps, nw, ndim))
    logps = np.empty((nsteps, nw))
    halves = [np.arange(0, nw // 2), np.arange(nw // 2, nw)]
    acc = calls = 0
    for it in range(nsteps):
        for h in (0, 1):
            S, C = halves[h], halves[1 - h]
            z = ((a - 1.0) * rng.random(len(S)) + 1.0) ** 2 / a
            cj = x[C[rng.integers(len(C), size=len(S))]]
            y = cj + z[:, None] * (x[S] - cj)
            new = np.asarray(log_prob_fn(y), dtype=np.float64)     # ONE call per half-step
            calls += 1
            assert new.shape == (len(S),)
            with np.errstate(invalid="ignore"):
                lnr = (ndim - 1.0) * np.log(z) + new - lp[S]
            ok = np.log(rng.random(len(S))) < lnr
            ok &= np.isfinite(new)
            x[S[ok]] = y[ok]
            lp[S[ok]] = new[ok]
            acc += int(ok.sum())
        chain[it] = x
        logps[it] = lp
    return chain, logps, acc, calls
