"""Synthetic workloads of the shapes `BASELINE.json:configs` names.

A *problem spec* is a plain dict mirroring the arguments a user gives the reference's
`Fitter` (`/root/reference/src/ravest/fit.py:51-195`):

    planet_letters, parameterisation (str), params {name: (value, fixed)} in the user's
    insertion order, priors {name: (kind, *args)}, time / vel / velerr (float64 arrays),
    instrument (array of str), t0; GP problems add hyperparams / hyperpriors.

The generators only build numpy inputs (no CUDA): `bench.py`, the parity tests and
`tests/golden/make_golden.py` all draw the same bytes from the same seeds, so every GPU
count and the CPU oracle see identical inputs.
"""
from __future__ import annotations

import numpy as np

PARS = {
    "P K e w Tp": ["P", "K", "e", "w", "Tp"],
    "P K e w Tc": ["P", "K", "e", "w", "Tc"],
    "P K secosw sesinw Tp": ["P", "K", "secosw", "sesinw", "Tp"],
    "P K secosw sesinw Tc": ["P", "K", "secosw", "sesinw", "Tc"],
}

LETTERS = "bcdefghij"


def free_names(spec) -> list[str]:
    return [k for k, (_, fixed) in spec["params"].items() if not fixed]


def _tc_from_tp(tp, P, e, w):
    """Inverse of the reference's convert_tc_to_tp (`param.py:159-196`), generator use only."""
    theta = np.pi / 2 - w
    E = 2 * np.arctan(np.sqrt((1 - e) / (1 + e)) * np.tan(theta / 2))
    return (E - e * np.sin(E)) * (P / (2 * np.pi)) + tp


def _kepler_rv_np(t, P, K, e, w, tp):
    """Plain numpy Keplerian RV used only to synthesise `vel` (not a parity oracle)."""
    M = 2 * np.pi / P * (t - tp)
    E = M + e * np.sin(M)
    for _ in range(60):
        E = E - (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
    cf = (np.cos(E) - e) / (1 - e * np.cos(E))
    sf = np.sqrt(1 - e * e) * np.sin(E) / (1 - e * np.cos(E))
    return K * (cf * np.cos(w) - sf * np.sin(w) + e * np.cos(w))


def make_multiplanet(
    n_planets: int,
    n_epochs: int,
    n_samples: int,
    seed: int,
    parameterisation: str = "P K secosw sesinw Tc",
    instruments: tuple[str, ...] = ("HARPS",),
    e_range: tuple[float, float] = (0.0, 0.6),
    p_range: tuple[float, float] = (1.5, 400.0),
    t_span: float = 1500.0,
    invalid_frac: float = 1e-3,
    periastron_frac: float = 0.0,
    prior_style: str = "mixed",
    fixed: tuple[str, ...] = (),
):
    """C3/C4-style workload: i.i.d. theta rows drawn from broad distributions.

    Returns (spec, theta[S, ndim] float64 C-contiguous).  Tp/Tc always lies inside the
    observing window so |M| stays ≲ 1e4 rad (SURVEY.md §7 "tolerance vs conditioning").
    """
    rng = np.random.default_rng(seed)
    pars = PARS[parameterisation]
    letters = LETTERS[:n_planets]
    N, S = n_epochs, n_samples

    t = np.sort(rng.uniform(0.0, t_span, N))
    sig = rng.uniform(0.5, 2.0, N)
    n_inst = len(instruments)
    # alternating blocks of epochs per instrument
    block = max(1, N // (4 * n_inst))
    inst_idx = (np.arange(N) // block) % n_inst
    inst_names = np.array(sorted(instruments))
    instrument = inst_names[inst_idx]

    # truth used to synthesise the velocities
    truth_P = np.exp(rng.uniform(np.log(p_range[0]), np.log(p_range[1]), n_planets))
    truth_K = rng.uniform(0.5, 20.0, n_planets)
    truth_e = rng.uniform(e_range[0], e_range[1], n_planets)
    truth_w = rng.uniform(-np.pi, np.pi, n_planets)
    truth_tp = rng.uniform(0, truth_P)
    vel = np.zeros(N)
    for k in range(n_planets):
        vel += _kepler_rv_np(t, truth_P[k], truth_K[k], truth_e[k], truth_w[k], truth_tp[k])
    truth_g = rng.uniform(-5, 5, n_inst)
    vel += truth_g[inst_idx] + rng.normal(0, sig)
    t0 = float(np.mean(t))

    # per-sample draws
    cols: dict[str, np.ndarray] = {}
    for k, L in enumerate(letters):
        P = np.exp(rng.uniform(np.log(p_range[0]), np.log(p_range[1]), S))
        K = rng.uniform(0.5, 20.0, S)
        e = rng.uniform(e_range[0], e_range[1], S)
        w = rng.uniform(-np.pi, np.pi, S)
        tp = rng.uniform(0.0, 1.0, S) * P + 0.25 * t_span
        if periastron_frac > 0:
            # force a slice of rows to e in [0.96, 0.97] with an epoch within 1e-3 P of periastron
            m = rng.uniform(0, 1, S) < periastron_frac
            e = np.where(m, rng.uniform(0.96, 0.97, S), e)
            pick = t[rng.integers(0, N, S)]
            tp = np.where(m, pick - rng.uniform(-1e-3, 1e-3, S) * P, tp)
        cols[f"P_{L}"] = P
        cols[f"K_{L}"] = K
        if "secosw" in pars:
            cols[f"secosw_{L}"] = np.sqrt(e) * np.cos(w)
            cols[f"sesinw_{L}"] = np.sqrt(e) * np.sin(w)
        else:
            cols[f"e_{L}"] = e
            cols[f"w_{L}"] = w
        if "Tc" in pars:
            cols[f"Tc_{L}"] = _tc_from_tp(tp, P, e, w)
        else:
            cols[f"Tp_{L}"] = tp
    cols["gd"] = rng.normal(0, 1e-3, S)
    cols["gdd"] = rng.normal(0, 1e-6, S)
    for name in inst_names:
        cols[f"g_{name}"] = rng.uniform(-5, 5, S)
    for name in inst_names:
        cols[f"jit_{name}"] = rng.uniform(0, 3, S)

    # params dict in the reference's conventional order (planets, g, gd, gdd, jit)
    order: list[str] = []
    for L in letters:
        order += [f"{p}_{L}" for p in pars]
    order += [f"g_{n}" for n in inst_names] + ["gd", "gdd"] + [f"jit_{n}" for n in inst_names]
    params = {}
    for name in order:
        params[name] = (float(np.median(cols[name])), name in fixed)
    # initial values must be valid for the reference's setter: use the first sample's planets
    for name in order:
        if name.split("_")[0] in ("P", "K", "e", "w", "secosw", "sesinw", "Tc", "Tp"):
            params[name] = (float(cols[name][0]), name in fixed)

    priors = {}
    for name in order:
        if name in fixed:
            continue
        base = name.split("_")[0]
        if base == "P":
            priors[name] = ("Uniform", p_range[0] * 0.5, p_range[1] * 2.0)
        elif base == "K":
            priors[name] = ("Uniform", 0.0, 50.0) if prior_style == "uniform" else ("HalfNormal", 25.0)
        elif base in ("secosw", "sesinw"):
            priors[name] = ("Uniform", -1.0, 1.0)
        elif base == "e":
            priors[name] = ("EccentricityUniform", 0.99) if prior_style == "uniform" else ("Beta", 0.867, 3.03)
        elif base == "w":
            priors[name] = ("Uniform", -np.pi, np.pi)
        elif base in ("Tc", "Tp"):
            priors[name] = ("Uniform", -2.0 * p_range[1], t_span + 2.0 * p_range[1])
        elif base == "g":
            priors[name] = ("Uniform", -10.0, 10.0) if prior_style == "uniform" else ("Normal", 0.0, 10.0)
        elif name == "gd":
            priors[name] = ("Normal", 0.0, 0.01)
        elif name == "gdd":
            priors[name] = ("Normal", 0.0, 1e-4)
        elif base == "jit":
            priors[name] = ("Uniform", 0.0, 5.0) if prior_style == "uniform" else ("HalfNormal", 3.0)

    names = [n for n in order if n not in fixed]
    theta = np.ascontiguousarray(np.stack([cols[n] for n in names], axis=1))

    # a fixed fraction of invalid rows to exercise the -inf paths
    n_bad = int(round(invalid_frac * S))
    if n_bad:
        bad_rows = rng.choice(S, n_bad, replace=False)
        jit_cols = [i for i, n in enumerate(names) if n.startswith("jit_")]
        k_cols = [i for i, n in enumerate(names) if n.startswith("K_")]
        u_cols = [i for i, n in enumerate(names) if n.startswith(("secosw_", "e_"))]
        g_cols = [i for i, n in enumerate(names) if n.startswith("g_")]
        for j, r in enumerate(bad_rows):
            kind = j % 4
            if kind == 0 and jit_cols:
                theta[r, jit_cols[0]] = -0.5
            elif kind == 1 and u_cols:
                theta[r, u_cols[j % len(u_cols)]] = 1.25
            elif kind == 2 and k_cols:
                theta[r, k_cols[j % len(k_cols)]] = -1.0
            elif g_cols:
                theta[r, g_cols[0]] = 1e3

    spec = {
        "planet_letters": list(letters),
        "parameterisation": parameterisation,
        "params": params,
        "priors": priors,
        "time": t,
        "vel": vel,
        "velerr": sig,
        "instrument": instrument,
        "t0": t0,
    }
    return spec, theta


def make_c1(n_samples: int = 32, seed: int = 101, circular: bool = False):
    """51-Peg-b-shaped: 1 planet, 153 epochs, 1 instrument (ELODIE), `P K e w Tc`."""
    fixed = ("e_b", "w_b", "gd", "gdd", "jit_ELODIE") if circular else ("gd", "gdd")
    spec, theta = make_multiplanet(
        1, 153, n_samples, seed, parameterisation="P K e w Tc", instruments=("ELODIE",),
        e_range=(0.0, 0.4), p_range=(2.0, 10.0), t_span=4200.0, prior_style="uniform",
        fixed=fixed)
    if circular:
        p = dict(spec["params"])
        p["e_b"] = (0.0, True)
        p["w_b"] = (float(np.pi / 2), True)
        p["gd"] = (0.0, True)
        p["gdd"] = (0.0, True)
        p["jit_ELODIE"] = (0.0, True)
        spec["params"] = p
    return spec, theta


def make_c2(n_samples: int = 100_000, seed: int = 202):
    """TOI-544-shaped: 2 planets, 120 epochs, single instrument, `P K secosw sesinw Tc`."""
    return make_multiplanet(
        2, 120, n_samples, seed, parameterisation="P K secosw sesinw Tc",
        instruments=("HARPS",), e_range=(0.0, 0.5), p_range=(1.5, 60.0), t_span=140.0,
        fixed=("gd", "gdd"))


def make_c3(n_samples: int = 1_000_000, seed: int = 303, n_epochs: int = 1000):
    """Synthetic 5-planet system, 1000 epochs, 1 instrument, everything free (ndim 29)."""
    return make_multiplanet(5, n_epochs, n_samples, seed)


def make_c4(n_samples: int = 1_000_000, seed: int = 404, n_epochs: int = 1000):
    """High-eccentricity stress: 3 planets, 2 instruments, e in U(0.6, 0.97) + periastron rows."""
    return make_multiplanet(
        3, n_epochs, n_samples, seed, instruments=("ESPRESSO", "HARPS"),
        e_range=(0.6, 0.97), periastron_frac=0.01)


def make_c5(n_samples: int = 10_000, seed: int = 505, n_planets: int = 1, n_epochs: int = 120):
    """K2-229-shaped quasi-periodic GP problem (adds hyperparams / hyperpriors)."""
    spec, theta = make_multiplanet(
        n_planets, n_epochs, n_samples, seed, parameterisation="P K secosw sesinw Tc",
        instruments=("HARPS",), e_range=(0.0, 0.5), p_range=(0.5, 40.0), t_span=110.0,
        fixed=("gd", "gdd"))
    rng = np.random.default_rng(seed + 1)
    S = n_samples
    hyper = {
        "gp_amp": rng.uniform(1, 15, S),
        "gp_lambda_e": rng.uniform(10, 100, S),
        "gp_lambda_p": rng.uniform(0.2, 1.5, S),
        "gp_period": rng.uniform(10, 30, S),
    }
    spec["hyperparams"] = {k: (float(v[0]), False) for k, v in hyper.items()}
    spec["hyperpriors"] = {
        "gp_amp": ("Uniform", 0.0, 50.0),
        "gp_lambda_e": ("Uniform", 1.0, 500.0),
        "gp_lambda_p": ("Uniform", 0.05, 5.0),
        "gp_period": ("Normal", 18.0, 10.0),
    }
    theta = np.ascontiguousarray(np.concatenate([theta, np.stack(list(hyper.values()), axis=1)], axis=1))
    n_bad = max(1, S // 1000)
    bad = rng.choice(S, n_bad, replace=False)
    theta[bad[::2], -1] = -1.0          # gp_period <= 0 -> -inf
    return spec, theta
