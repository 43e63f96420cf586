"""Magnetar-driven relativistic ejecta: basic_mergernova, general_mergernova,
general_mergernova_thermalisation -- device vs the oracle (which reproduces the reference's own
_ejecta_dynamics_and_interaction bit for bit, tests/test_verbatim_fixtures.py)."""
import numpy as np
import pytest

import helpers as h
from oracle import restated as r

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def rb():
    import redback_b200
    return redback_b200


def _exp_condition(t, z, nu, dyn):
    """The reference evaluates exp(x) - 1 (magnetar_driven_ejecta_models.py:172): one rounding of exp(x) is a relative
    error eps / x of the flux for small x = h nu / (k T' D)."""
    from scipy.interpolate import interp1d
    t_grid, temp_g, _, dop_g = dyn
    ts = t * 86400 / (1 + z)
    x = r.planck * nu * (1 + z) / (r.boltzmann_constant * interp1d(t_grid, temp_g)(ts) * interp1d(t_grid, dop_g)(ts))
    return 4 * 2.3e-16 / np.minimum(x, 1.0)


@pytest.mark.parametrize("kind", ["basic", "general", "thermalisation"])
def test_prior_draws(rb, kind):
    rng = np.random.default_rng({"basic": 91, "general": 92, "thermalisation": 93}[kind])
    t = np.concatenate([np.geomspace(0.01, 300, 40), np.geomspace(0.02, 200, 30)])     # unsorted on purpose
    nu = np.concatenate([np.full(40, 4.6e14), np.where(np.arange(30) % 2 == 0, 5e9, 2.4e17)])
    N = 40
    p = h.mergernova_prior(rng, N, kind)
    name = dict(basic="basic_mergernova", general="general_mergernova",
                thermalisation="general_mergernova_thermalisation")[kind]
    fn = getattr(rb.magnetar_driven_ejecta_models, name)
    kw = dict(frequency=nu, output_format="flux_density")
    if kind == "general":
        kw["pair_cascade_switch"] = False           # non-default value
    out = fn(t, params_batch=p, **kw)
    y = out[7] * 1.05
    sigma = 0.1 * np.abs(out[7]) + 1e-30
    logl = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=kw).log_likelihood_batch(p)
    tg = r.get_optimal_time_array(1e-4, 1e8, 500)
    for i in range(N):
        kwi = {k: v[i] for k, v in p.items()}
        z = kwi.pop("redshift")
        okw = dict(frequency=nu)
        if kind == "general":
            okw["pair_cascade_switch"] = False
        ref = getattr(r, name)(t, z, **kwi, **okw)
        lum_fn = (lambda tt: r.basic_magnetar(tt, kwi["p0"], kwi["bp"], kwi["mass_ns"], kwi["theta_pb"])) \
            if kind == "basic" else (lambda tt: r.magnetar_only(tt, kwi["l0"], kwi["tau_sd"], kwi["nn"]))
        dyn = r.ejecta_dynamics_and_interaction(
            tg, kwi["mej"], kwi["beta"], kwi["ejecta_radius"], kwi["kappa"], kwi["n_ism"], lum_fn(tg),
            kind != "basic" and kind != "general", kind == "thermalisation",
            thermalisation_efficiency=kwi.get("thermalisation_efficiency"), kappa_gamma=kwi.get("kappa_gamma"))
        extra = _exp_condition(t, z, nu, (tg, dyn[1], dyn[2], dyn[3]))
        h.assert_close(out[i], ref, RTOL, extra, what=f"{name} sample {i}")
        ref_l = r.gaussian_likelihood(y, ref, sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-8 * abs(ref_l), (i, logl[i], ref_l)


def test_reference_golden_vectors(rb):
    """The reference's own regression vectors (test/reference_results/all_models_reference.pkl.gz) for the three
    mergernova models, through the drop-in functions."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "all_models_subset.json")) as fh:
        golden = json.load(fh)
    for name in ("basic_mergernova", "general_mergernova", "general_mergernova_thermalisation"):
        e = golden["magnetar." + name]
        t = np.array(e["time"])
        for params, res in zip(e["params"], e["results"]):
            out = getattr(rb.magnetar_driven_ejecta_models, name)(t, **params, **e["eval_kwargs"])
            # exp(x) - 1 of the reference is ill-conditioned for small x (see _exp_condition): x > 1e-3 here
            h.assert_close(out, h.golden_array(res), 1e-9, 1e-12, what=name)


def test_trapped_magnetar(rb):
    """trapped_magnetar (magnetar_driven_ejecta_models.py:477-573): time in seconds, 'flux' and 'luminosity' outputs;
    the reference's golden vectors and prior draws against the oracle."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "all_models_subset.json")) as fh:
        e = json.load(fh)["magnetar.trapped_magnetar"]
    t = np.array(e["time"])
    fn = rb.magnetar_driven_ejecta_models.trapped_magnetar
    for params, res in zip(e["params"], e["results"]):
        out = fn(t, **params, **e["eval_kwargs"])
        h.assert_close(out, h.golden_array(res), 1e-9, 1e-12, what="trapped_magnetar golden")
    rng = np.random.default_rng(95)
    ts = np.geomspace(10.0, 1e6, 30)                      # seconds
    nu = np.full(30, 2.4e17)
    N = 16
    p = h.mergernova_prior(rng, N, "general")
    flux = fn(ts, params_batch=p, frequency=nu, output_format="flux", photon_index=1.8)
    lum = fn(ts, params_batch={k: v for k, v in p.items() if k != "redshift"}, frequency=nu, output_format="luminosity")
    y = flux[2] * 1.1
    sigma = 0.2 * np.abs(flux[2]) + 1e-30
    kw = dict(frequency=nu, output_format="flux", photon_index=1.8)
    logl = rb.GaussianLikelihood(ts, y, sigma, fn, kwargs=kw).log_likelihood_batch(p)
    for i in range(N):
        kwi = {k: v[i] for k, v in p.items()}
        z = kwi.pop("redshift")
        ref_f = r.trapped_magnetar(ts, z, frequency=nu, output_format="flux", photon_index=1.8, **kwi)
        ref_l = r.trapped_magnetar(ts, 0.0, frequency=nu, output_format="luminosity", **kwi)
        h.assert_close(flux[i], ref_f, 1e-9, 1e-11, what=f"trapped flux {i}")
        h.assert_close(lum[i], ref_l, 1e-9, 1e-11, what=f"trapped luminosity {i}")
        want = r.gaussian_likelihood(y, ref_f, sigma)
        assert abs(logl[i] - want) <= 1e-6 + 1e-8 * abs(want)
    with pytest.raises(KeyError):
        fn(ts, params_batch=p, frequency=nu, output_format="flux")          # photon_index is required (:518)
    with pytest.raises(ValueError):
        fn(ts, params_batch=p, frequency=nu, output_format="flux_density")


@pytest.mark.parametrize("kind", ["basic", "general", "thermalisation"])
def test_metzger_magnetar_driven(rb, kind):
    """metzger_magnetar_driven_kilonova_model / general_metzger_magnetar_driven(_thermalisation)
    (magnetar_driven_ejecta_models.py:576-905): the reference's golden vectors and prior draws against the oracle.
    The thermalisation variant's efficiency 1 - exp(-x), x ~ 1e-15, is quantised in units of 1.1e-16 in the reference
    itself (see tests/test_oracle_golden.py): it is compared at the reference's own 1e-5."""
    import json
    import os
    name = dict(basic="metzger_magnetar_driven_kilonova_model", general="general_metzger_magnetar_driven",
                thermalisation="general_metzger_magnetar_driven_thermalisation")[kind]
    rtol = 1e-5 if kind == "thermalisation" else 1e-9
    fn = getattr(rb.magnetar_driven_ejecta_models, name)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "all_models_subset.json")) as fh:
        e = json.load(fh)["magnetar." + name]
    t = np.array(e["time"])
    for params, res in zip(e["params"], e["results"]):
        out = fn(t, **params, **e["eval_kwargs"])
        h.assert_close(out, h.golden_array(res), max(rtol, 2e-7), what=name + " golden")   # Wien-tail values x 200
    rng = np.random.default_rng({"basic": 111, "general": 112, "thermalisation": 113}[kind])
    t = np.geomspace(0.05, 60, 30)
    nu = np.where(np.arange(30) % 3 == 0, 3e14, 6e14)
    N = 12
    p = h.metzger_magnetar_prior(rng, N, kind)
    kw = dict(frequency=nu, output_format="flux_density")
    out = fn(t, params_batch=p, **kw)
    y = out[1] * 1.04
    sigma = 0.1 * np.abs(out[1]) + 1e-30
    logl = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=kw).log_likelihood_batch(p)
    for i in range(N):
        kwi = {k: v[i] for k, v in p.items()}
        z = kwi.pop("redshift")
        ref = getattr(r, name)(t, z, frequency=nu, **kwi)
        # h nu / k T amplifies the ~1e-15 agreement of the temperatures in the Wien tail
        with np.errstate(all="ignore"):
            amp = 1e-13 * np.maximum(1.0, -np.log(np.maximum(np.abs(ref), 1e-300) / np.nanmax(np.abs(ref))))
        h.assert_close(out[i], ref, rtol, amp, what=f"{name} sample {i}")
        ref_l = r.gaussian_likelihood(y, ref, sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + max(rtol, 1e-8) * abs(ref_l), (i, logl[i], ref_l)


def test_metzger_magnetar_all_layers(rb):
    """magnetar_heating='all_layers' (magnetar_driven_ejecta_models.py:698-704): every shell is heated by the engine.
    The oracle's march is pinned to the reference's own source for both settings
    (tests/golden/metzger_magnetar_reference_draws.json)."""
    rng = np.random.default_rng(115)
    t = np.geomspace(0.05, 60, 30)
    nu = np.where(np.arange(30) % 3 == 0, 3e14, 6e14)
    N = 8
    for kind, name in (("general", "general_metzger_magnetar_driven"),
                       ("thermalisation", "general_metzger_magnetar_driven_thermalisation")):
        p = h.metzger_magnetar_prior(rng, N, kind)
        fn = getattr(rb.magnetar_driven_ejecta_models, name)
        kw = dict(frequency=nu, output_format="flux_density", magnetar_heating="all_layers")
        out = fn(t, params_batch=p, **kw)
        first = fn(t, params_batch=p, frequency=nu, output_format="flux_density")
        assert not np.allclose(out, first, rtol=1e-3, equal_nan=True)
        rtol = 1e-5 if kind == "thermalisation" else 1e-9
        for i in range(N):
            kwi = {k: v[i] for k, v in p.items()}
            z = kwi.pop("redshift")
            ref = getattr(r, name)(t, z, frequency=nu, magnetar_heating="all_layers", **kwi)
            with np.errstate(all="ignore"):
                amp = 1e-13 * np.maximum(1.0, -np.log(np.maximum(np.abs(ref), 1e-300) / np.nanmax(np.abs(ref))))
            h.assert_close(out[i], ref, rtol, amp, what=f"{name} all_layers sample {i}")
    with pytest.raises(ValueError):
        fn(t, params_batch=p, frequency=nu, output_format="flux_density", magnetar_heating="some_layers")


def test_t0_base_model_mergernova(rb, obands=None):
    """phase_models.t0_base_model (phase_models.py:19-59) over the mergernova models: the per-sample merger epoch is a
    shift inside mn_kernel.  t0_base_model(t, t0) must equal the plain model (checked against the oracle above) at
    t[t >= t0] - t0, return 0 (1000 mag) before t0, and the fused log L must include the placeholder epochs."""
    rng = np.random.default_rng(97)
    mjd0 = 60000.0
    t = mjd0 + np.sort(rng.uniform(0.2, 40, 36))
    nu = np.where(np.arange(t.size) % 2 == 0, 4.6e14, 6.3e14)
    N = 10
    p = h.mergernova_prior(rng, N, "general")
    p["t0"] = mjd0 + rng.uniform(-5, 10, N)
    fn = rb.phase_models.t0_base_model
    plain = rb.magnetar_driven_ejecta_models.general_mergernova
    kw = dict(base_model="general_mergernova", frequency=nu, output_format="flux_density")
    out = fn(t, params_batch=p, **kw)
    y = np.abs(out[3]) * 1.05 + 1e-8
    sigma = 0.1 * y
    logl = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=kw).log_likelihood_batch(p)
    for i in range(N):
        kwi = {k: float(v[i]) for k, v in p.items()}
        t0 = kwi.pop("t0")
        good = t - t0 >= 0
        assert np.all(out[i][~good] == 0.0)
        want = plain(t[good] - t0, frequency=nu[good], output_format="flux_density", **kwi)
        h.assert_close(out[i][good], want, 1e-12, what=f"t0(general_mergernova) sample {i}")
        ref_l = r.gaussian_likelihood(y, out[i], sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-9 * abs(ref_l), (i, logl[i], ref_l)
    rb.bands.register_synthetic_lsst()
    bands = np.array(list(h.LSST_NU)[:6] * 6)
    mag = fn(t, params_batch=p, base_model="general_mergernova", bands=bands, output_format="magnitude")
    for i in range(N):
        kwi = {k: float(v[i]) for k, v in p.items()}
        t0 = kwi.pop("t0")
        good = t - t0 >= 0
        assert np.all(mag[i][~good] == 1000.0)
        want = plain(t[good] - t0, bands=bands[good], output_format="magnitude", **kwi)
        ok = np.isfinite(want) & (want < 60)
        assert np.max(np.abs(mag[i][good][ok] - want[ok])) < 1e-8


def test_t0_base_model_metzger_magnetar(rb):
    """t0_base_model over the magnetar-driven Metzger shell models (rb_mk_config.t0): same checks as for the
    mergernova models."""
    rng = np.random.default_rng(98)
    mjd0 = 60000.0
    t = mjd0 + np.sort(rng.uniform(0.2, 30, 24))
    nu = np.where(np.arange(t.size) % 2 == 0, 4.6e14, 6.3e14)
    N = 6
    p = h.metzger_magnetar_prior(rng, N, "general")
    p["t0"] = mjd0 + rng.uniform(-5, 8, N)
    fn = rb.phase_models.t0_base_model
    plain = rb.magnetar_driven_ejecta_models.general_metzger_magnetar_driven
    kw = dict(base_model="general_metzger_magnetar_driven", frequency=nu, output_format="flux_density")
    out = fn(t, params_batch=p, **kw)
    y = np.abs(out[2]) * 1.05 + 1e-8
    sigma = 0.1 * y
    logl = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=kw).log_likelihood_batch(p)
    for i in range(N):
        kwi = {k: float(v[i]) for k, v in p.items()}
        t0 = kwi.pop("t0")
        good = t - t0 >= 0
        assert np.all(out[i][~good] == 0.0)
        want = plain(t[good] - t0, frequency=nu[good], output_format="flux_density", **kwi)
        h.assert_close(out[i][good], want, 1e-12, what=f"t0(general_metzger_magnetar_driven) sample {i}")
        ref_l = r.gaussian_likelihood(y, out[i], sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-9 * abs(ref_l), (i, logl[i], ref_l)
    rb.bands.register_synthetic_lsst()
    bands = np.array(list(h.LSST_NU)[:6] * 4)
    mag = fn(t, params_batch=p, base_model="general_metzger_magnetar_driven", bands=bands, output_format="magnitude")
    for i in range(N):
        kwi = {k: float(v[i]) for k, v in p.items()}
        t0 = kwi.pop("t0")
        good = t - t0 >= 0
        assert np.all(mag[i][~good] == 1000.0)
        want = plain(t[good] - t0, bands=bands[good], output_format="magnitude", **kwi)
        ok = np.isfinite(want) & (want < 60)
        assert np.max(np.abs(mag[i][good][ok] - want[ok])) < 1e-8


def test_evolving_magnetar_variants(rb):
    """general_mergernova_evolution (:422-474) and general_metzger_magnetar_driven_evolution (:908-956): the spin-down
    ODE of magnetar_models._evolving_gw_and_em_magnetar runs on device; golden vectors + a small batch."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "all_models_subset.json")) as fh:
        golden = json.load(fh)
    mod = rb.magnetar_driven_ejecta_models
    for name, rtol in (("general_mergernova_evolution", 1e-9), ("general_metzger_magnetar_driven_evolution", 1e-5)):
        e = golden["magnetar." + name]
        t = np.array(e["time"])
        for params, res in zip(e["params"], e["results"]):
            out = getattr(mod, name)(t, **params, **e["eval_kwargs"])
            h.assert_close(out, h.golden_array(res), rtol, 1e-9, what=name + " golden")
    rng = np.random.default_rng(121)
    N = 8
    ev = dict(logbint=rng.uniform(14, 17, N), logbext=rng.uniform(13.5, 16, N), p0=h.loguniform(rng, 7e-4, 1e-2, N),
              chi0=rng.uniform(0, 1.5, N), radius=rng.uniform(10, 13, N), logmoi=rng.uniform(44.5, 45.5, N),
              kappa_gamma=h.loguniform(rng, 1e-2, 1e3, N))
    t = np.geomspace(0.05, 40, 24)
    nu = np.where(np.arange(24) % 2 == 0, 3e14, 6e14)
    base = h.mergernova_prior(rng, N, "general")
    p = dict({k: base[k] for k in ("redshift", "mej", "beta", "ejecta_radius", "kappa", "n_ism")}, **ev)
    out = mod.general_mergernova_evolution(t, params_batch=p, frequency=nu, output_format="flux_density")
    mk = h.metzger_magnetar_prior(rng, N, "general")
    q = dict({k: mk[k] for k in ("redshift", "mej", "vej", "beta", "kappa_r")}, **ev)
    out2 = mod.general_metzger_magnetar_driven_evolution(t, params_batch=q, frequency=nu, output_format="flux_density")
    for i in range(N):
        kwi = {k: v[i] for k, v in p.items()}
        z = kwi.pop("redshift")
        ref = r.general_mergernova_evolution(t, z, frequency=nu, **kwi)
        h.assert_close(out[i], ref, 1e-9, 1e-9, what=f"general_mergernova_evolution sample {i}")
        kwi = {k: v[i] for k, v in q.items()}
        z = kwi.pop("redshift")
        ref = r.general_metzger_magnetar_driven_evolution(t, z, frequency=nu, **kwi)
        h.assert_close(out2[i], ref, 1e-5, what=f"general_metzger_magnetar_driven_evolution sample {i}")


def test_scalar_call_kwargs_and_errors(rb):
    t = np.geomspace(0.05, 100, 12)
    kw = dict(mej=0.02, beta=0.3, ejecta_radius=3e9, kappa=5.0, n_ism=1e-2, l0=1e47, tau_sd=1e3, nn=3.0,
              thermalisation_efficiency=0.4)
    out = rb.magnetar_driven_ejecta_models.general_mergernova(t, 0.05, frequency=4.6e14, output_format="flux_density",
                                                              ejecta_albedo=0.3, pair_cascade_fraction=0.2, **kw)
    ref = r.general_mergernova(t, 0.05, frequency=np.full(12, 4.6e14), ejecta_albedo=0.3, pair_cascade_fraction=0.2,
                               **kw)
    h.assert_close(out, ref, RTOL, what="general_mergernova kwargs")
    with pytest.raises(ValueError):
        rb.magnetar_driven_ejecta_models.general_mergernova(np.array([0.0, 1.0]), 0.05, frequency=4.6e14,
                                                            output_format="flux_density", **kw)
    with pytest.raises(KeyError):
        rb.magnetar_driven_ejecta_models.general_mergernova(t, 0.05, output_format="magnitude", **kw)     # no bands
    with pytest.raises(KeyError):
        rb.magnetar_driven_ejecta_models.general_mergernova(t, 0.05, output_format="flux_density", **kw)
    far = rb.magnetar_driven_ejecta_models.general_mergernova(np.array([10.0, 2000.0]), 0.05, frequency=4.6e14,
                                                              output_format="flux_density", **kw)
    assert np.isfinite(far[0]) and np.isnan(far[1])          # 2000 d > 1e8 s: interp1d raises in the reference
