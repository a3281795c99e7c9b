"""CPU restatement (numpy / pure Python) of ravest's RV log-probability path.

TEST INFRASTRUCTURE ONLY — imported by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`; never by the product package.

Parity status: PINNED for the white-noise path (configs 1-4) — checked against golden
vectors produced by the unmodified reference (`tests/golden/*.json`, made by
`tests/golden/make_golden.py`), which include the reference's own published known answers
(rv1/rv2 vectors, notebook MAP log-posteriors).  The GP path (config 5) is "parity
unpinned": its arithmetic lives in tinygp 0.3.0 / jax 0.8.3 (`poetry.lock:5191`, `1827`),
absent from /root/reference and from this image; `gp_log_likelihood` restates the published
dense-Cholesky Gaussian log-density at the reference's call sites
(`src/ravest/gp.py:145-156`, `src/ravest/fit.py:8058-8060`).

Everything works at the level of the reference's own objects (names, dicts), one sample at
a time, deliberately independent of the product's flat descriptor compiler.
All `file:line` citations are relative to /root/reference/src/ravest/.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import gammaln, logsumexp, xlog1py, xlogy
from scipy.stats import halfnorm, rayleigh, truncnorm

PARS = {
    "P K e w Tp": ["P", "K", "e", "w", "Tp"],
    "P K e w Tc": ["P", "K", "e", "w", "Tc"],
    "P K secosw sesinw Tp": ["P", "K", "secosw", "sesinw", "Tp"],
    "P K secosw sesinw Tc": ["P", "K", "secosw", "sesinw", "Tc"],
}


class InvalidParams(ValueError):
    pass


# ----------------------------------------------------------------------------- model.py
def solve_kepler(Mi: float, e: float) -> tuple[float, float]:
    """model.py:23-70 — Halley from E0=M, tol 1.48e-8, ≤50 iterations; returns (cosE, sinE)."""
    tol = 1.48e-08
    Ei = Mi
    sin_E = cos_E = 0.0
    for _ in range(50):
        sin_E = math.sin(Ei)
        cos_E = math.cos(Ei)
        f = Ei - e * sin_E - Mi
        fp = 1.0 - e * cos_E
        fpp = e * sin_E
        E_new = Ei - f / (fp - (f * fpp) / (2.0 * fp))
        if abs(E_new - Ei) < tol:
            sin_E = math.sin(E_new)
            cos_E = math.cos(E_new)
            break
        Ei = E_new
    return cos_E, sin_E


def kepler_rv(M: np.ndarray, e: float, K: float, w: float) -> np.ndarray:
    """model.py:173-243 — `_compute_rv`: circular shortcut for e == 0 else the scalar loop."""
    M = np.asarray(M, dtype=np.float64)
    if e == 0:
        return K * (np.cos(M + w) + e * np.cos(w))                      # model.py:242
    sqrt_1me2 = math.sqrt(1.0 - e * e)                                  # model.py:203-206
    cos_w = math.cos(w)
    sin_w = math.sin(w)
    e_cos_w = e * cos_w
    rv = np.empty(M.shape[0])
    for i in range(M.shape[0]):
        cos_E, sin_E = solve_kepler(float(M[i]), e)
        denom = 1.0 - e * cos_E                                         # model.py:119-121
        cos_f = (cos_E - e) / denom
        sin_f = sqrt_1me2 * sin_E / denom
        rv[i] = K * (cos_f * cos_w - sin_f * sin_w + e_cos_w)           # model.py:170
    return rv


# ----------------------------------------------------------------------------- param.py
def _validate_eccentricity(e: float) -> None:
    """param.py:59-63."""
    if e < 0:
        raise InvalidParams("e < 0")
    if e >= 1.0:
        raise InvalidParams("e >= 1")


def convert_tc_to_tp(tc: float, P: float, e: float, w: float) -> float:
    """param.py:198-215."""
    theta_tc = (np.pi / 2) - w
    _validate_eccentricity(e)
    E = 2 * np.arctan(np.sqrt((1 - e) / (1 + e)) * np.tan(theta_tc / 2))
    Mc = E - (e * np.sin(E))
    return tc - (P / (2 * np.pi)) * Mc


def to_default(parameterisation: str, p: dict) -> dict:
    """param.py:299-362 — convert one planet's params to `P K e w Tp`."""
    if parameterisation == "P K e w Tp":
        return {"P": p["P"], "K": p["K"], "e": p["e"], "w": p["w"], "Tp": p["Tp"]}
    if parameterisation == "P K e w Tc":
        tp = convert_tc_to_tp(p["Tc"], p["P"], p["e"], p["w"])
        return {"P": p["P"], "K": p["K"], "e": p["e"], "w": p["w"], "Tp": tp}
    e = p["secosw"] ** 2 + p["sesinw"] ** 2                             # param.py:232-233
    w = np.arctan2(p["sesinw"], p["secosw"])
    if parameterisation == "P K secosw sesinw Tp":
        return {"P": p["P"], "K": p["K"], "e": e, "w": w, "Tp": p["Tp"]}
    if parameterisation == "P K secosw sesinw Tc":
        tp = convert_tc_to_tp(p["Tc"], p["P"], e, w)
        return {"P": p["P"], "K": p["K"], "e": e, "w": w, "Tp": tp}
    raise ValueError(parameterisation)


def validate_default(d: dict) -> None:
    """param.py:17-105 — P>0, K>0, 0<=e<1, -pi<=w<pi (NaN passes every test)."""
    if d["P"] <= 0:
        raise InvalidParams("P")
    if d["K"] <= 0:
        raise InvalidParams("K")
    _validate_eccentricity(d["e"])
    if not -np.pi <= d["w"] < np.pi:
        raise InvalidParams("w")


def planet_rv(parameterisation: str, p: dict, t: np.ndarray) -> np.ndarray:
    """model.py:259-275 + 329-354 — Planet(...).radial_velocity(t)."""
    d = to_default(parameterisation, p)
    validate_default(d)
    n = 2 * np.pi / d["P"]                                              # model.py:302
    M = n * (np.asarray(t, dtype=np.float64) - d["Tp"])                 # model.py:327
    return kepler_rv(M, d["e"], d["K"], d["w"])


def trend_rv(gd: float, gdd: float, t: np.ndarray, t0: float) -> np.ndarray:
    """model.py:483-509."""
    t = np.asarray(t, dtype=np.float64)
    rv = 0
    rv = rv + (np.zeros(len(t)) if gd == 0 else gd * (t - t0))
    rv = rv + (np.zeros(len(t)) if gdd == 0 else gdd * ((t - t0) ** 2))
    return rv


# ----------------------------------------------------------------------------- prior.py
def prior_logpdf(kind: str, args, x: float) -> float:
    """prior.py:49-65, 106-122, 158-171, 230-246, 287-303, 343-359, 418-440, 490-508."""
    if kind == "Uniform":
        lo, hi = args
        return -np.inf if (x < lo or x > hi) else -np.log(hi - lo)
    if kind == "EccentricityUniform":
        (hi,) = args
        return -np.inf if (x < 0 or x >= hi) else -np.log(hi)
    if kind == "Normal":
        mu, sd = args
        return -0.5 * ((x - mu) / sd) ** 2 - 0.5 * np.log((sd ** 2) * 2.0 * np.pi)
    if kind == "TruncatedNormal":
        mu, sd, lo, hi = args
        if x < lo or x > hi:
            return -np.inf
        return truncnorm.logpdf(x, (lo - mu) / sd, (hi - mu) / sd, loc=mu, scale=sd)
    if kind == "HalfNormal":
        (sd,) = args
        return -np.inf if x < 0.0 else halfnorm.logpdf(x, scale=float(sd))
    if kind == "Rayleigh":
        (sc,) = args
        if x < 0.0:
            return -np.inf
        with np.errstate(divide="ignore"):
            return rayleigh.logpdf(x, scale=float(sc))
    if kind == "VanEylen19Mixture":
        sn, sr, f = (float(a) for a in args)
        if x < 0.0:
            return -np.inf
        with np.errstate(divide="ignore"):
            lh = halfnorm.logpdf(x, scale=sn)
            lr = rayleigh.logpdf(x, scale=sr)
            return logsumexp([lh, lr], b=[1 - f, f])
    if kind == "Beta":
        a, b = (float(v) for v in args)
        if x < 0.0 or x > 1.0:
            return -np.inf
        lb = gammaln(a) + gammaln(b) - gammaln(a + b)
        return xlogy(a - 1, x) + xlog1py(b - 1, -x) - lb
    raise ValueError(kind)


# ----------------------------------------------------------------------------- fit.py
class Problem:
    """Holds what `LogPosterior.__init__` holds (fit.py:3234-3304)."""

    def __init__(self, spec: dict):
        self.spec = spec
        self.letters = list(spec["planet_letters"])
        self.parameterisation = spec["parameterisation"]
        self.pars = PARS[self.parameterisation]
        self.params = dict(spec["params"])
        self.priors = {k: (v[0], tuple(v[1:])) for k, v in spec["priors"].items()}
        self.free_names = [k for k, (_, fx) in self.params.items() if not fx]
        self.fixed = {k: v for k, (v, fx) in self.params.items() if fx}
        self.time = np.ascontiguousarray(spec["time"], dtype=np.float64)
        self.vel = np.ascontiguousarray(spec["vel"], dtype=np.float64)
        self.velerr = np.ascontiguousarray(spec["velerr"], dtype=np.float64)
        self.instrument = np.asarray(spec["instrument"])
        self.unique = np.unique(self.instrument)                        # fit.py:113
        self.t0 = spec["t0"]
        idx = {inst: i for i, inst in enumerate(self.unique)}           # fit.py:3585-3586
        self.inst_idx = np.array([idx[i] for i in self.instrument])
        self.velerr_sq = self.velerr ** 2                               # fit.py:3598
        self.log_2pi = np.log(2 * np.pi)
        # GP extras
        self.hyperparams = dict(spec.get("hyperparams", {}))
        self.hyperpriors = {k: (v[0], tuple(v[1:])) for k, v in spec.get("hyperpriors", {}).items()}
        self.free_hyper = [k for k, (_, fx) in self.hyperparams.items() if not fx]
        self.fixed_hyper = {k: v for k, (v, fx) in self.hyperparams.items() if fx}
        self.jacobian, self.renorm = self._corrections()

    # fit.py:3306-3397
    def _corrections(self) -> tuple[float, float]:
        log_jac = np.log(2) if "secosw" in self.parameterisation else 0.0   # param.py:428-435
        tj = tr = 0.0
        for L in self.letters:
            case = "CASE_1"
            if log_jac != 0.0 and f"secosw_{L}" in self.free_names:
                if f"secosw_{L}" in self.priors and f"sesinw_{L}" in self.priors:
                    a, b = self.priors[f"secosw_{L}"], self.priors[f"sesinw_{L}"]
                    if a[0] == "Uniform" and b[0] == "Uniform" and a[1] == (-1, 1) and b[1] == (-1, 1):
                        case = "CASE_2"
                    else:
                        raise NotImplementedError("unsupported priors on (secosw, sesinw)")
                elif f"e_{L}" in self.priors and f"w_{L}" in self.priors:
                    case = "CASE_3"
                else:
                    raise RuntimeError("cannot classify planet")
            tj += log_jac if case == "CASE_3" else 0.0
            tr += np.log(4.0 / np.pi) if case == "CASE_2" else 0.0
        return tj, tr

    # fit.py:3600-3660
    def log_likelihood(self, params: dict) -> float:
        rv_total = np.zeros(len(self.time))
        for L in self.letters:
            pp = {par: params[f"{par}_{L}"] for par in self.pars}
            try:
                rv_total += planet_rv(self.parameterisation, pp, self.time)
            except InvalidParams:
                return -np.inf
        rv_total += trend_rv(params["gd"], params["gdd"], self.time, self.t0)
        gam = np.array([params[f"g_{i}"] for i in self.unique])
        rv_total += gam[self.inst_idx]
        jit = np.array([params[f"jit_{i}"] for i in self.unique])[self.inst_idx]
        var = self.velerr_sq + jit ** 2
        with np.errstate(all="ignore"):
            penalty = self.log_2pi + np.log(var)
            chi2 = (rv_total - self.vel) ** 2 / var
            return float(-0.5 * np.sum(chi2 + penalty))

    # fit.py:3399-3446
    def params_for_prior(self, free: dict) -> dict:
        if set(self.priors) == set(self.free_names):
            return free
        out = {k: v for k, v in free.items() if k in self.priors}
        allp = self.fixed | free
        for L in self.letters:
            pp = {par: allp[f"{par}_{L}"] for par in self.pars}
            d = to_default(self.parameterisation, pp)
            for k, v in d.items():
                if f"{k}_{L}" in self.priors:
                    out[f"{k}_{L}"] = v
        return out

    def log_prior(self, params_for_prior: dict) -> float:
        """fit.py:3672-3691."""
        lp = 0
        for k, v in params_for_prior.items():
            lp += prior_logpdf(self.priors[k][0], self.priors[k][1], v)
        return lp

    # fit.py:3448-3495
    def log_probability(self, free: dict) -> float:
        allp = self.fixed | free
        for inst in self.unique:
            if allp[f"jit_{inst}"] < 0:
                return -np.inf
        try:
            lp = self.log_prior(self.params_for_prior(free))
        except InvalidParams:
            return -np.inf
        if not np.isfinite(lp):
            return -np.inf
        ll = self.log_likelihood(allp)
        logprob = ll + lp
        logprob += self.jacobian
        logprob += self.renorm
        return logprob

    def log_probability_batch(self, theta: np.ndarray) -> np.ndarray:
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        out = np.empty(theta.shape[0])
        for i, row in enumerate(theta):
            out[i] = self.log_probability(dict(zip(self.free_names, (float(x) for x in row))))
        return out

    # ------------------------------------------------------------------ GP (parity unpinned)
    def mean_model(self, params: dict):
        """fit.py:7994-8043."""
        rv_total = np.zeros(len(self.time))
        for L in self.letters:
            pp = {par: params[f"{par}_{L}"] for par in self.pars}
            try:
                rv_total += planet_rv(self.parameterisation, pp, self.time)
            except InvalidParams:
                return None
        rv_total += trend_rv(params["gd"], params["gdd"], self.time, self.t0)
        gam = np.array([params[f"g_{i}"] for i in self.unique])
        return rv_total + gam[self.inst_idx]

    def gp_log_likelihood(self, params: dict, hyper: dict) -> float:
        """fit.py:8062-8105 with tinygp's quasi-periodic kernel restated from gp.py:145-156:
        k(tau) = A^2 exp(-Gamma sin^2(pi |tau| / P_gp)) exp(-tau^2 / (2 lambda_e^2)),
        Gamma = 1/(2 lambda_p^2); dense Cholesky log-density."""
        mean = self.mean_model(params)
        if mean is None or not np.isfinite(mean).all():
            return -np.inf
        A, le, lp_, Pg = (hyper[k] for k in ("gp_amp", "gp_lambda_e", "gp_lambda_p", "gp_period"))
        gamma = 1 / (2 * lp_ ** 2)
        tau = self.time[:, None] - self.time[None, :]
        C = A ** 2 * np.exp(-gamma * np.sin(np.pi * np.abs(tau) / Pg) ** 2) * np.exp(-0.5 * (tau / le) ** 2)
        jit = np.array([params[f"jit_{i}"] for i in self.unique])[self.inst_idx]
        C = C + np.diag(self.velerr_sq + jit ** 2)
        try:
            Lc = np.linalg.cholesky(C)
        except np.linalg.LinAlgError:
            return float("nan")
        from scipy.linalg import solve_triangular
        alpha = solve_triangular(Lc, self.vel - mean, lower=True)
        n = len(self.time)
        return float(-0.5 * alpha @ alpha - np.sum(np.log(np.diag(Lc))) - 0.5 * n * np.log(2 * np.pi))

    def gp_log_probability(self, combined: dict) -> float:
        """fit.py:7836-7901."""
        free = {k: combined[k] for k in self.free_names}
        fh = {k: combined[k] for k in self.free_hyper}
        allp = self.fixed | free
        for inst in self.unique:
            if allp[f"jit_{inst}"] < 0:
                return -np.inf
        allh = self.fixed_hyper | fh
        for k in ("gp_amp", "gp_lambda_e", "gp_lambda_p", "gp_period"):       # gp.py:98-108
            if not np.isfinite(allh[k]) or allh[k] <= 0:
                return -np.inf
        try:
            lp = self.log_prior(self.params_for_prior(free))
        except InvalidParams:
            return -np.inf
        if not np.isfinite(lp):
            return -np.inf
        lhp = 0
        for k, v in fh.items():
            lhp += prior_logpdf(self.hyperpriors[k][0], self.hyperpriors[k][1], v)
        if not np.isfinite(lhp):
            return -np.inf
        ll = self.gp_log_likelihood(allp, allh)
        logprob = ll + lp + lhp
        logprob += self.jacobian
        logprob += self.renorm
        return logprob

    def gp_log_probability_batch(self, theta: np.ndarray) -> np.ndarray:
        names = self.free_names + self.free_hyper                            # fit.py:4978
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        return np.array([self.gp_log_probability(dict(zip(names, (float(x) for x in r)))) for r in theta])

    # ------------------------------------------------------------------ RV matrices (row f-1)
    def build_params(self, row) -> dict:
        return self.fixed | dict(zip(self.free_names, (float(x) for x in row)))

    def rv_matrix(self, theta: np.ndarray, times: np.ndarray, component: str, frozen: dict | None = None) -> np.ndarray:
        """fit.py:2726-2824 — component: planet letter, 'trend' or 'total'; `frozen` is the resolved
        freeze_params mapping applied to every sample (`params.update(resolved_freeze)`, fit.py:2743-2745)."""
        theta = np.atleast_2d(theta)
        out = np.zeros((theta.shape[0], len(times)))
        for i, row in enumerate(theta):
            p = self.build_params(row)
            if frozen:
                p.update(frozen)
            acc = np.zeros(len(times))
            if component in ("trend", "total"):
                acc = acc + trend_rv(p["gd"], p["gdd"], times, self.t0)
            for L in self.letters:
                if component == L or component == "total":
                    acc = acc + planet_rv(self.parameterisation, {q: p[f"{q}_{L}"] for q in self.pars}, times)
            out[i] = acc
        return out

    # ------------------------------------------------------------------ information criteria (Appendix B.10)
    def information_criteria(self, row) -> tuple[float, float, float, float]:
        """fit.py:1361-1384, 1457-1554: (log-likelihood, chi2, AICc, BIC) for one row of free values."""
        params = self.build_params(row)
        ll = self.log_likelihood(params)
        var = np.zeros_like(self.velerr_sq)
        for j, inst in enumerate(self.unique):                                   # fit.py:1493-1497
            mask = self.inst_idx == j
            var[mask] = self.velerr_sq[mask] + params[f"jit_{inst}"] ** 2
        penalty = np.sum(np.log(2 * np.pi * var))                                # fit.py:1499
        chi2 = -2 * ll - penalty                                                 # fit.py:1500
        k, n = len(self.free_names), len(self.time)
        aicc = (2 * k - 2 * ll) + (2 * k ** 2 + 2 * k) / (n - k - 1)             # fit.py:1526-1529
        bic = k * np.log(n) - 2 * ll                                             # fit.py:1554
        return float(ll), float(chi2), float(aicc), float(bic)

    # ------------------------------------------------------------------ walker checks (row f-3)
    def walker_stage(self, row) -> tuple[str, float | None]:
        """What fit.py:1048-1062 (and the retry loops fit.py:692-725, 884-902) decide for one candidate row:
        'astro' = _validate_astrophysical_validity raised (fit.py:260-293), 'prior' = non-finite log-prior,
        'ok' otherwise (with the log-prior)."""
        free = dict(zip(self.free_names, (float(x) for x in row)))
        allp = self.fixed | free
        if any(not np.isfinite(v) for v in allp.values()):                 # fit.py:262-265
            return "astro", None
        for L in self.letters:                                             # fit.py:268-276, param.py:107-126
            pp = {par: allp[f"{par}_{L}"] for par in self.pars}
            try:
                validate_default(to_default(self.parameterisation, pp))
            except InvalidParams:
                return "astro", None
        for inst in self.unique:                                           # fit.py:282-293
            if allp[f"jit_{inst}"] < 0:
                return "astro", None
        try:
            lp = self.log_prior(self.params_for_prior(free))               # fit.py:1057-1058
        except InvalidParams:
            return "astro", None
        if not np.isfinite(lp):
            return "prior", None
        return "ok", float(lp)

    # ------------------------------------------------------------------ GP conditioning (row f-4; parity unpinned)
    def _gp_system(self, params: dict, hyper: dict):
        A, le, lp_, Pg = (hyper[k] for k in ("gp_amp", "gp_lambda_e", "gp_lambda_p", "gp_period"))
        gamma = 1 / (2 * lp_ ** 2)

        def kern(t1, t2):
            tau = t1[:, None] - t2[None, :]
            return A ** 2 * np.exp(-gamma * np.sin(np.pi * np.abs(tau) / Pg) ** 2) * np.exp(-0.5 * (tau / le) ** 2)

        jit = np.array([params[f"jit_{i}"] for i in self.unique])[self.inst_idx]
        C = kern(self.time, self.time) + np.diag(self.velerr_sq + jit ** 2)         # fit.py:6399, 7531
        rv = np.zeros(len(self.time))
        for L in self.letters:
            rv += planet_rv(self.parameterisation, {par: params[f"{par}_{L}"] for par in self.pars}, self.time)
        rv += trend_rv(params["gd"], params["gdd"], self.time, self.t0)
        gam = np.array([params[f"g_{i}"] for i in self.unique])[self.inst_idx]
        resid = (self.vel - gam) - rv                                                # fit.py:7543-7550
        return kern, C, resid

    def gp_predict(self, combined: dict, times: np.ndarray) -> tuple[np.ndarray, float]:
        """fit.py:7494-7554 (`gp.condition(y=residuals, X_test=times)` mean, zero GP mean function) and
        fit.py:5386-5429 (chi2 = alpha.alpha): mu* = K(t*, t) C^-1 r, restated with a dense Cholesky."""
        from scipy.linalg import cho_solve, solve_triangular
        allp = self.fixed | {k: combined[k] for k in self.free_names}
        allh = self.fixed_hyper | {k: combined[k] for k in self.free_hyper}
        kern, C, resid = self._gp_system(allp, allh)
        Lc = np.linalg.cholesky(C)
        alpha = solve_triangular(Lc, resid, lower=True)
        beta = cho_solve((Lc, True), resid)
        return kern(np.asarray(times, dtype=np.float64), self.time) @ beta, float(alpha @ alpha)
