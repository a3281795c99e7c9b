"""Generate tests/golden/gp_sklearn.json: GP log-likelihood, conditional mean and chi^2 from scikit-learn.

    python tests/golden/make_gp_sklearn.py      (build container; needs scikit-learn, committed output travels)

tinygp / jax - where the reference's GP arithmetic lives (fit.py:8047-8060, gp.py:145-156, fit.py:7536-7554) - are
absent from this image, so the GP rows are "parity unpinned" against the reference itself.  This fixture pins them to
an INDEPENDENT published implementation of the same model instead of the builder's own restatement:

    sklearn.gaussian_process.GaussianProcessRegressor(
        kernel = ConstantKernel(A^2) * ExpSineSquared(length_scale = 2 lambda_p, periodicity = P_gp) * RBF(lambda_e),
        alpha = sigma_i^2 + jit_inst(i)^2, optimizer = None, normalize_y = False)

    ExpSineSquared: exp(-2 sin^2(pi d / p) / l^2)  ==  exp(-Gamma sin^2(pi d / P_gp)), Gamma = 1 / (2 lambda_p^2)  (gp.py:150-153)
    RBF:            exp(-d^2 / (2 l^2))            ==  tinygp ExpSquared(scale = lambda_e)                          (gp.py:147)

`log_marginal_likelihood()` is the Gaussian log-density of y = vel - mean model (Rasmussen & Williams alg. 2.1), i.e.
`GaussianProcess.log_probability(y)`; `predict(t*)` is the conditional mean `gp.condition(y, X_test).mean`;
`y . alpha_` is chi^2 = y^T C^-1 y (fit.py:5427-5429).  The mean model (planets + trend + gamma) and the priors come
from oracle/oracle_py.py, which is pinned bit-for-bit to the reference on the white-noise fixtures.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle_py  # noqa: E402
from ravest_b200 import workloads  # noqa: E402


def sklearn_gp(pr: "oracle_py.Problem", combined: dict, times: np.ndarray):
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, ExpSineSquared
    allp = pr.fixed | {k: combined[k] for k in pr.free_names}
    allh = pr.fixed_hyper | {k: combined[k] for k in pr.free_hyper}
    mean = pr.mean_model(allp)
    if mean is None:
        return None
    A, le, lp_, Pg = (allh[k] for k in ("gp_amp", "gp_lambda_e", "gp_lambda_p", "gp_period"))
    kernel = (ConstantKernel(A ** 2, constant_value_bounds="fixed")
              * ExpSineSquared(length_scale=2.0 * lp_, periodicity=Pg, length_scale_bounds="fixed", periodicity_bounds="fixed")
              * RBF(length_scale=le, length_scale_bounds="fixed"))
    jit = np.array([allp[f"jit_{i}"] for i in pr.unique])[pr.inst_idx]
    y = pr.vel - mean
    gpr = GaussianProcessRegressor(kernel=kernel, alpha=pr.velerr_sq + jit ** 2, optimizer=None, normalize_y=False)
    gpr.fit(pr.time[:, None], y)
    ll = float(gpr.log_marginal_likelihood())
    mu = gpr.predict(np.asarray(times)[:, None])
    chi2 = float(y @ gpr.alpha_)
    return ll, mu, chi2


def main():
    cases = []
    for n_epochs, n_planets, seed, S in ((30, 1, 11, 24), (120, 1, 12, 24), (120, 2, 13, 20), (57, 2, 14, 20)):
        spec, theta = workloads.make_c5(n_samples=S, n_planets=n_planets, n_epochs=n_epochs, seed=seed)
        pr = oracle_py.Problem(spec)
        names = pr.free_names + pr.free_hyper
        times = np.linspace(spec["time"].min() - 2.0, spec["time"].max() + 2.0, 9)
        full = pr.gp_log_probability_batch(theta)          # -inf pattern + priors + corrections (restatement)
        ll, mean, chi2, logprob = [], [], [], []
        for row, lp_full in zip(theta, full):
            comb = dict(zip(names, map(float, row)))
            res = sklearn_gp(pr, comb, times) if np.isfinite(lp_full) else None
            if res is None:
                ll.append(None); mean.append(None); chi2.append(None); logprob.append(float(lp_full))
                continue
            allp = pr.fixed | {k: comb[k] for k in pr.free_names}
            allh = pr.fixed_hyper | {k: comb[k] for k in pr.free_hyper}
            own_ll = pr.gp_log_likelihood(allp, allh)
            ll.append(res[0]); mean.append(res[1].tolist()); chi2.append(res[2])
            logprob.append(res[0] + (float(lp_full) - own_ll))       # sklearn likelihood + pinned prior / correction part
        cases.append({"name": f"c5 N={n_epochs} planets={n_planets}", "n_epochs": n_epochs, "n_planets": n_planets,
                      "seed": seed, "n_samples": S, "names": names, "theta": theta.tolist(), "times": times.tolist(),
                      "ll": ll, "mean": mean, "chi2": chi2, "logprob": logprob})
    import sklearn
    out = {"generator": "tests/golden/make_gp_sklearn.py", "sklearn_version": sklearn.__version__, "cases": cases}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gp_sklearn.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
