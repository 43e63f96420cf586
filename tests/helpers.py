"""Shared helpers for parity tests: seeded synthetic inputs per SURVEY.md §8d and comparison rules."""
import numpy as np


def golden_array(values):
    return np.array([np.nan if v is None else float(v) for v in values], dtype=float)


def loguniform(rng, lo, hi, n):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n))


def arnett_bolometric_prior(rng, n):
    """priors/arnett_bolometric.prior."""
    return dict(f_nickel=loguniform(rng, 1e-3, 1, n), mej=loguniform(rng, 1e-4, 100, n),
                vej=loguniform(rng, 1e3, 1e5, n), kappa=rng.uniform(0.05, 2, n),
                kappa_gamma=loguniform(rng, 1e-4, 1e4, n))


def slsn_prior(rng, n, zmax=3.0):
    """priors/slsn.prior."""
    return dict(redshift=rng.uniform(1e-6, zmax, n), p0=rng.uniform(1, 10, n), bp=loguniform(rng, 0.1, 10, n),
                mass_ns=rng.uniform(1.1, 2.2, n), theta_pb=rng.uniform(0, 1.57, n),
                mej=loguniform(rng, 0.1, 100, n), vej=loguniform(rng, 1e3, 1e5, n), kappa=rng.uniform(0.05, 2, n),
                kappa_gamma=loguniform(rng, 1e-4, 1e4, n), temperature_floor=loguniform(rng, 1e3, 1e5, n))


def arnett_prior(rng, n, zmax=1.0):
    d = arnett_bolometric_prior(rng, n)
    d.update(redshift=rng.uniform(1e-3, zmax, n), temperature_floor=loguniform(rng, 1e3, 1e5, n))
    return d


def config1_time():
    """examples/bolometric_luminosity_fitting.py:11-12"""
    return np.append(np.geomspace(0.5, 50, 20), np.linspace(50, 400, 80))


def config2_epochs(rng):
    """ZTF g/r/i, 100 epochs per band, t = sort(U(1, 300 d)); frequencies from tables/filters.csv:65-67."""
    nus = {"ztfg": 6.272e14, "ztfr": 4.675e14, "ztfi": 3.813e14}
    t = np.concatenate([rng.uniform(1, 300, 100) for _ in nus])
    nu = np.concatenate([np.full(100, v) for v in nus.values()])
    order = np.argsort(t, kind="stable")
    return t[order], nu[order]


def diffusion_condition(time_src, tau_d):
    """Relative noise floor of the reference's own result: eps * (t/tau_d)^2 from the cancellation in
    t_i^2 - t_e^2 (interaction_processes.py:57).  Used to widen rtol for extreme samples."""
    return 4 * np.finfo(float).eps * (np.asarray(time_src) / tau_d) ** 2


def assert_close(actual, expected, rtol, extra_rtol=0.0, what=""):
    actual = np.asarray(actual, dtype=float)
    expected = np.asarray(expected, dtype=float)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    nan_a, nan_e = np.isnan(actual), np.isnan(expected)
    assert np.array_equal(nan_a, nan_e), f"{what}: NaN pattern differs ({nan_a.sum()} vs {nan_e.sum()})"
    ok = ~nan_e
    tol = (rtol + extra_rtol) * np.abs(expected)
    diff = np.abs(actual - expected)
    with np.errstate(invalid="ignore"):
        bad = ok & ~(diff <= tol) & ~((actual == expected))
    if bad.any():
        i = np.argmax(np.where(bad, diff / np.maximum(np.abs(expected), 1e-300), 0))
        raise AssertionError(f"{what}: {bad.sum()} / {bad.size} outside rtol={rtol}; worst idx {i}: "
                             f"{actual.flat[i]!r} vs {expected.flat[i]!r}")
