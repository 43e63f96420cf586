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


LSST_NU = dict(lsstu=8.152e14, lsstg=6.273e14, lsstr=4.825e14, lssti=3.983e14, lsstz=3.454e14, lssty=3.083e14)


def config3_epochs(n=200):
    """t = geomspace(0.1, 30 d, 200) cycling LSST u,g,r,i,z,y
    (effective frequencies from redback/tables/filters.csv:54-59)."""
    t = np.geomspace(0.1, 30, n)
    nus = list(LSST_NU.values())
    nu = np.array([nus[i % 6] for i in range(n)])
    return t, nu


def one_component_prior(rng, n):
    """priors/one_component_kilonova_model.prior"""
    return dict(redshift=rng.uniform(1e-6, 0.1, n), mej=rng.uniform(0.01, 0.05, n), vej=rng.uniform(0.1, 0.5, n),
                kappa=rng.uniform(1, 30, n), temperature_floor=loguniform(rng, 100, 6000, n))


def metzger_prior(rng, n):
    """priors/metzger_kilonova_model.prior"""
    return dict(redshift=rng.uniform(1e-3, 0.1, n), mej=rng.uniform(0.01, 0.05, n), vej=rng.uniform(0.1, 0.5, n),
                beta=rng.uniform(3, 5, n), kappa=rng.uniform(1, 30, n))


def engine_atan_condition(t_src_seconds):
    """Noise floor of the reference's own kilonova engine term (0.5 - arctan((t-1.3)/0.11)/pi)**1.3
    (kilonova_models.py:1878): one ulp of arctan near pi/2 (2.2e-16) is a relative error
    1.3 * 2.2e-16 * pi * x of the luminosity, x = (t-1.3)/0.11."""
    x = np.maximum((np.asarray(t_src_seconds) - 1.3) / 0.11, 1.0)
    return 1.3 * 2.2e-16 * np.pi * x


def tophat_prior(rng, n):
    """priors/tophat_redback.prior."""
    return dict(redshift=rng.uniform(0.01, 3, n), thv=np.arccos(rng.uniform(0, 1, n)), loge0=rng.uniform(44, 54, n),
                thc=rng.uniform(0.01, 0.1, n), logn0=rng.uniform(-5, 2, n), p=rng.uniform(2, 3, n),
                logepse=rng.uniform(-5, 0, n), logepsb=rng.uniform(-5, 0, n), g0=rng.uniform(100, 2000, n),
                xiN=rng.uniform(0, 1, n))


def config4_epochs(n_per=250):
    """radio + X-ray, 250 epochs each, t = geomspace(0.1, 300 d)."""
    t = np.concatenate([np.geomspace(0.1, 300, n_per)] * 2)
    nu = np.concatenate([np.full(n_per, 5e9), np.full(n_per, 2e17)])
    order = np.argsort(t, kind="stable")
    return t[order], nu[order]


def csm_prior(rng, n, zmax=1.0):
    """priors/csm_interaction.prior"""
    return dict(redshift=rng.uniform(1e-3, zmax, n), mej=loguniform(rng, 1, 100, n),
                csm_mass=loguniform(rng, 1, 100, n), vej=loguniform(rng, 1e3, 1e5, n), eta=rng.uniform(0, 2, n),
                rho=loguniform(rng, 1e-14, 1e-12, n), kappa=rng.uniform(0.05, 2, n), r0=rng.uniform(50, 700, n),
                temperature_floor=loguniform(rng, 100, 1e4, n))


def mergernova_prior(rng, n, kind="basic"):
    """priors/basic_mergernova.prior, general_mergernova.prior, general_mergernova_thermalisation.prior"""
    p = dict(redshift=rng.uniform(1e-6, 0.1, n), mej=rng.uniform(1e-4, 0.1, n), beta=rng.uniform(0.1, 0.7, n),
             ejecta_radius=loguniform(rng, 1e9, 1e10, n), kappa=rng.uniform(1, 30, n),
             n_ism=loguniform(rng, 1e-4, 1, n))
    if kind == "basic":
        p.update(p0=loguniform(rng, 0.7, 1e4, n), bp=loguniform(rng, 1e-4, 1e4, n), mass_ns=rng.uniform(1.1, 2.2, n),
                 theta_pb=rng.uniform(0.01, 1.57, n))
    else:
        p.update(l0=loguniform(rng, 1e40, 1e50, n), tau_sd=loguniform(rng, 10, 1e5, n), nn=rng.uniform(2, 7, n))
    if kind == "thermalisation":
        p["kappa_gamma"] = loguniform(rng, 1e-4, 1e4, n)
    else:
        p["thermalisation_efficiency"] = rng.uniform(0.1, 1, n)
    return p


def metzger_magnetar_prior(rng, n, kind="basic"):
    """priors/metzger_magnetar_driven_kilonova_model.prior, general_metzger_magnetar_driven(_thermalisation).prior"""
    p = dict(redshift=rng.uniform(1e-6, 0.1, n), mej=rng.uniform(1e-2, 0.05, n), vej=rng.uniform(0.05, 0.3, n),
             beta=rng.uniform(1.5, 8, n), kappa_r=rng.uniform(1, 30, n))
    if kind == "basic":
        p.update(p0=loguniform(rng, 0.7, 1e2, n), bp=loguniform(rng, 1e-2, 1e2, n), mass_ns=rng.uniform(1.1, 2.2, n),
                 theta_pb=rng.uniform(0.01, 1.57, n))
    else:
        p.update(l0=loguniform(rng, 1e40, 1e48, n), tau_sd=loguniform(rng, 10, 1e5, n), nn=rng.uniform(2, 7, n))
    if kind == "thermalisation":
        p["kappa_gamma"] = loguniform(rng, 1e-4, 1e4, n)
    else:
        p["thermalisation_efficiency"] = rng.uniform(0.1, 1, n)
    return p


def realistic_bandpass_table(lo, hi, seed=0):
    """A transmission table shaped like a real survey filter file (sncosmo's lsst* / ztf* data are not available
    offline): NON-UNIFORM wavelength sampling (fine at the edges, coarse on the plateau), exactly-zero wings of
    several samples on both sides, asymmetric edges, a sloped plateau with ripples and a narrow absorption dip.
    These are the cases in which sncosmo's `integration_grid` / linear `Bandpass.__call__` differ from a top-hat."""
    rng = np.random.default_rng(seed)
    w = np.concatenate([np.arange(lo - 320.0, lo - 60.0, 52.0), np.arange(lo - 60.0, lo + 110.0, 2.5),
                        np.arange(lo + 110.0, hi - 120.0, 23.0 + 4.0 * rng.uniform()),
                        np.arange(hi - 120.0, hi + 90.0, 3.3), np.arange(hi + 90.0, hi + 420.0, 41.0)])
    w = np.unique(np.round(w, 3))
    rise = 1.0 / (1.0 + np.exp(-(w - lo) / 9.0))
    fall = 1.0 / (1.0 + np.exp((w - hi) / 17.0))
    plateau = 0.62 + 0.2 * (w - lo) / (hi - lo) + 0.015 * np.sin((w - lo) / 37.0)
    dip_at = lo + 0.63 * (hi - lo)
    dip = 1.0 - 0.35 * np.exp(-0.5 * ((w - dip_at) / 6.0) ** 2)
    tr = rise * fall * plateau * dip
    tr[(w < lo - 60.0) | (w > hi + 90.0)] = 0.0          # zero-padded wings
    return w, tr
