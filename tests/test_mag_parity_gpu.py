"""GPU parity: magnitude / bandpass-flux branch (device spectral grid -> not-a-knot bicubic spline -> band
integrals) vs the oracle built on scipy's RectBivariateSpline (the spline sncosmo.TimeSeriesSource uses).

PARITY UNPINNED against the reference itself: sncosmo is not installable here and the reference has no golden
vectors for this branch; bandpasses are documented synthetic top-hats.

Tolerance: |d mag| <= 1e-9 + 1e-9*|mag|, widened where the *spline itself* is ill-conditioned: the fit is done
in linear flux, so a band whose flux is f_band / f_peak of the brightest part of the spectral grid carries a
relative noise ~ eps * f_peak / f_band in both implementations (Wien tail in the blue bands of cool sources)."""
import numpy as np
import pytest

import helpers as h
from oracle import restated as r

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import redback_b200
    redback_b200.bands.register_synthetic_lsst()
    return redback_b200


@pytest.fixture(scope="module")
def obands(rb):
    return {n: r.Bandpass(n, b.wave, b.trans) for n, b in
            ((n, rb.bands.get_bandpass(n)) for n in rb.bands.SYNTHETIC_LSST_EDGES)}


def _check(mag, ref, what, atol=2e-9, floor_mag=None):
    mag, ref = np.asarray(mag), np.asarray(ref)
    assert mag.shape == ref.shape
    bad_nan = np.isnan(mag) != np.isnan(ref)
    assert not bad_nan.any(), (what, "NaN pattern")
    ok = ~np.isnan(ref)
    tol = atol + 1e-9 * np.abs(ref)
    if floor_mag is not None:
        # flux far below the brightest band of the same sample: spline round-off floor (see module docstring)
        tol = tol + 2.3e-16 * 10 ** (0.4 * (ref - floor_mag)) * 1e3
    diff = np.abs(mag - ref)
    assert np.all(diff[ok] <= tol[ok]), (what, float(np.max(diff[ok])), int(np.argmax(np.where(ok, diff, 0))))


def test_host_band_tables_match_oracle(rb, obands):
    for n, ob in obands.items():
        b = rb.bands.get_bandpass(n)
        w, dw = b.integration_grid()
        ow, odw = r.integration_grid(ob.minwave(), ob.maxwave(), 5.0)
        assert np.array_equal(w, ow) and dw == odw
        assert b.zpbandflux_ab() == ob.zpbandflux_ab()
        assert np.array_equal(b(w), ob(w))


@pytest.mark.parametrize("model", ["arnett", "basic_magnetar_powered", "slsn", "type_1a"])
def test_supernova_magnitudes(rb, obands, model):
    rng = np.random.default_rng(21)
    t = np.sort(rng.uniform(3, 150, 48))
    names = np.array(list(h.LSST_NU)[:6] * 8)
    N = 24
    if model in ("arnett", "type_1a"):
        p = h.arnett_prior(rng, N, zmax=0.5)
        fn = getattr(rb.supernova_models, model)
    else:
        p = h.slsn_prior(rng, N, zmax=0.5)
        fn = getattr(rb.supernova_models, model)
    # keep the photosphere in the optical so that every LSST band has signal
    p["temperature_floor"] = h.loguniform(rng, 3e3, 1.2e4, N)
    out = fn(t, params_batch=p, bands=names, output_format="magnitude")
    assert out.shape == (N, t.size)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.sn_magnitude(model, t, z, bands=names, bandpasses=obands, **kw)
        _check(out[i], ref, f"{model} sample {i}", floor_mag=np.nanmin(ref))


@pytest.mark.parametrize("n_lambda,n_epochs", [(57, 48), (300, 48), (131, 1500), (200, 7), (9, 12), (16, 12), (33, 12)])
def test_photometry_kernel_odd_shapes(rb, obands, n_lambda, n_epochs):
    """Shapes that exercise every branch of the photometry kernel's work split: wavelength grids that are not a
    multiple of the warp / segment size, more columns than threads (two columns per thread in the time-axis passes),
    more epochs than one hand-over chunk, fewer epochs than warps."""
    rng = np.random.default_rng(1000 + n_lambda + n_epochs)
    t = np.sort(rng.uniform(3, 150, n_epochs))
    names = np.array((list(h.LSST_NU)[:6] * (n_epochs // 6 + 1))[:n_epochs])
    lam = np.geomspace(1500, 25000, n_lambda)
    N = 5
    p = h.arnett_prior(rng, N, zmax=0.3)
    p["temperature_floor"] = h.loguniform(rng, 3e3, 1.2e4, N)
    out = rb.supernova_models.arnett(t, params_batch=p, bands=names, output_format="magnitude", lambda_array=lam)
    for i in range(N if n_epochs < 1000 else 2):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.sn_magnitude("arnett", t, z, bands=names, bandpasses=obands, lambda_array=lam, **kw)
        _check(out[i], ref, f"arnett NL={n_lambda} E={n_epochs} sample {i}", floor_mag=np.nanmin(ref))


def test_csm_interaction_magnitudes(rb, obands):
    """csm_interaction magnitude branch (supernova_models.py:2093-2111): grid get_optimal_time_array(0.1, 500, 300)."""
    rng = np.random.default_rng(27)
    t = np.sort(rng.uniform(3, 150, 36))
    names = np.array(list(h.LSST_NU)[:6] * 6)
    N = 10
    p = h.csm_prior(rng, N, zmax=0.5)
    p["temperature_floor"] = h.loguniform(rng, 3e3, 1.2e4, N)
    out = rb.supernova_models.csm_interaction(t, params_batch=p, bands=names, output_format="magnitude")
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.sn_magnitude("csm_interaction", t, z, bands=names, bandpasses=obands, **kw)
        _check(out[i], ref, f"csm_interaction sample {i}", floor_mag=np.nanmin(ref))


@pytest.mark.parametrize("kind", ["basic", "thermalisation"])
def test_mergernova_magnitudes(rb, obands, kind):
    """_processing_other_formats of the mergernova models (magnetar_driven_ejecta_models.py:198-238)."""
    rng = np.random.default_rng(29)
    t = np.geomspace(0.3, 40, 30)
    names = np.array(list(h.LSST_NU)[:6] * 5)
    N = 8
    p = h.mergernova_prior(rng, N, kind)
    name = "basic_mergernova" if kind == "basic" else "general_mergernova_thermalisation"
    out = getattr(rb.magnetar_driven_ejecta_models, name)(t, params_batch=p, bands=names, output_format="magnitude")
    assert out.shape == (N, t.size)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.mergernova_magnitude(name, t, z, bands=names, bandpasses=obands, **kw)
        _check(out[i], ref, f"{name} sample {i}", atol=2e-8, floor_mag=np.nanmin(ref))


def test_metzger_magnetar_driven_magnitudes(rb, obands):
    """Magnitude branch of the magnetar-driven Metzger shell models (_processing_other_formats, plain blackbody); the
    last grid point of the reference has T = NaN (0/0), which the spectra clean-up replaces -- same on device."""
    rng = np.random.default_rng(31)
    t = np.geomspace(0.3, 30, 24)
    names = np.array(list(h.LSST_NU)[:6] * 4)
    N = 6
    p = h.metzger_magnetar_prior(rng, N, "general")
    name = "general_metzger_magnetar_driven"
    out = getattr(rb.magnetar_driven_ejecta_models, name)(t, params_batch=p, bands=names, output_format="magnitude")
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.metzger_magnetar_magnitude(name, t, z, bands=names, bandpasses=obands, **kw)
        _check(out[i], ref, f"{name} sample {i}", atol=2e-8, floor_mag=np.nanmin(ref))


def test_kilonova_magnitudes_config3(rb, obands):
    """BASELINE configs[2]: one_component_kilonova_model in LSST ugrizy magnitudes."""
    t = np.geomspace(0.3, 20, 36)
    names = np.array(list(h.LSST_NU)[:6] * 6)
    rng = np.random.default_rng(22)
    N = 16
    p = h.one_component_prior(rng, N)
    fn = rb.kilonova_models.one_component_kilonova_model
    out = fn(t, params_batch=p, bands=names, output_format="magnitude")
    y = r.kn_magnitude("one_component", t, 0.05, bands=names, bandpasses=obands, mej=0.03, vej=0.2, kappa=5.,
                       temperature_floor=3000.)
    like = rb.GaussianLikelihood(t, y, 0.1, fn, kwargs=dict(bands=names, output_format="magnitude"))
    logl = like.log_likelihood_batch(p)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.kn_magnitude("one_component", t, z, bands=names, bandpasses=obands, **kw)
        _check(out[i], ref, f"kn sample {i}", floor_mag=np.nanmin(ref))
        ref_l = r.gaussian_likelihood(y, ref, 0.1)
        assert abs(logl[i] - ref_l) <= 1e-5 + 1e-8 * abs(ref_l), (i, logl[i], ref_l)
    # metzger + flux output
    pm = h.metzger_prior(rng, 6)
    outm = rb.kilonova_models.metzger_kilonova_model(t, params_batch=pm, bands=names, output_format="magnitude")
    outf = rb.kilonova_models.metzger_kilonova_model(t, params_batch=pm, bands=names, output_format="flux")
    ref_flux = np.array([rb.bands.REFERENCE_FLUX[n] for n in names])
    for i in range(6):
        kw = {k: v[i] for k, v in pm.items()}
        z = kw.pop("redshift")
        ref = r.kn_magnitude("metzger", t, z, bands=names, bandpasses=obands, **kw)
        _check(outm[i], ref, f"metzger sample {i}", floor_mag=np.nanmin(ref))
        # 'flux' is the same band integral: its relative tolerance carries the same round-off floor as the magnitudes
        rtol_i = 1e-7 + 2.3e-13 * 10 ** (0.4 * (ref - np.nanmin(ref)))
        assert np.all(np.abs(outf[i] - r.bandpass_magnitude_to_flux(ref, ref_flux))
                      <= rtol_i * r.bandpass_magnitude_to_flux(ref, ref_flux))


def test_multi_component_kilonova_magnitudes(rb, obands):
    """two_component_kilonova_model / three_component_kilonova_model, magnitude and flux branches
    (kilonova_models.py:1166-1211, 1277-1323): summed blackbody spectra on the shared grid."""
    names = np.array(list(h.LSST_NU)[:6] * 5)
    rng = np.random.default_rng(31)
    N = 8
    ref_flux = np.array([rb.bands.REFERENCE_FLUX[n] for n in names])
    for ncomp, t_hi, res, gmax in ((2, 5.0, 500, 86400 * 6), (3, 20.0, 300, 7e6)):
        t = np.geomspace(0.3, t_hi, 30)
        comps = [h.one_component_prior(rng, N) for _ in range(ncomp)]
        p = dict(redshift=comps[0]["redshift"])
        for c, comp in enumerate(comps, start=1):
            for k in ("mej", "vej", "kappa", "temperature_floor"):
                p[f"{k}_{c}"] = comp[k]
        fn = getattr(rb.kilonova_models, {2: "two", 3: "three"}[ncomp] + "_component_kilonova_model")
        out = fn(t, params_batch=p, bands=names, output_format="magnitude")
        outf = fn(t, params_batch=p, bands=names, output_format="flux")
        for i in range(N):
            cl = [(comp["mej"][i], comp["vej"][i], comp["kappa"][i], comp["temperature_floor"][i]) for comp in comps]
            ref = r.multi_component_kn_magnitude(t, p["redshift"][i], cl, bands=names, bandpasses=obands,
                                                 dense_resolution=res, grid_t_max=gmax)
            _check(out[i], ref, f"{ncomp}-component sample {i}", floor_mag=np.nanmin(ref))
            np.testing.assert_allclose(outf[i], r.bandpass_magnitude_to_flux(ref, ref_flux), rtol=1e-7)
        # three identical components = the one-component model, 2.5 log10(3) mag brighter (same grid)
        if ncomp == 3:
            single = rb.kilonova_models.one_component_kilonova_model(
                t, params_batch=dict(redshift=p["redshift"], mej=p["mej_1"], vej=p["vej_1"], kappa=p["kappa_1"],
                                     temperature_floor=p["temperature_floor_1"]),
                bands=names, output_format="magnitude", dense_resolution=300)
            same = dict(p)
            for c in (2, 3):
                for k in ("mej", "vej", "kappa", "temperature_floor"):
                    same[f"{k}_{c}"] = p[f"{k}_1"]
            out3 = fn(t, params_batch=same, bands=names, output_format="magnitude")
            ok = np.isfinite(single) & (single < 60)
            assert np.max(np.abs(out3[ok] - (single[ok] - 2.5 * np.log10(3.0)))) < 1e-9


def test_general_magnetar_driven_supernova_magnitudes(rb, obands):
    """Magnitude / flux branch of general_magnetar_driven_supernova (supernova_models.py:2585-2637): CutoffBlackbody
    spectra on the 1998-point dynamics grid."""
    rng = np.random.default_rng(57)
    t = np.sort(rng.uniform(2, 200, 24))
    names = np.array(list(h.LSST_NU)[:6] * 4)
    N = 4
    p = dict(redshift=rng.uniform(0.01, 0.4, N), mej=h.loguniform(rng, 1, 30, N), E_sn=h.loguniform(rng, 1e50, 3e51, N),
             kappa=rng.uniform(0.05, 0.3, N), l0=h.loguniform(rng, 1e44, 1e47, N), tau_sd=h.loguniform(rng, 1e4, 1e7, N),
             nn=rng.uniform(2.2, 6, N), kappa_gamma=h.loguniform(rng, 1e-2, 1e2, N), f_nickel=rng.uniform(0, 0.3, N),
             temperature_floor=h.loguniform(rng, 3e3, 1e4, N))
    fn = rb.supernova_models.general_magnetar_driven_supernova
    out = fn(t, params_batch=p, bands=names, output_format="magnitude")
    outf = fn(t, params_batch=p, bands=names, output_format="flux")
    ref_flux = np.array([rb.bands.REFERENCE_FLUX[n] for n in names])
    y = out[0] + 0.1
    logl = rb.GaussianLikelihood(t, y, 0.1, fn, kwargs=dict(bands=names, output_format="magnitude")).log_likelihood_batch(p)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.general_magnetar_driven_supernova_magnitude(t, z, bands=names, bandpasses=obands, **kw)
        _check(out[i], ref, f"general_magnetar_driven_supernova sample {i}", atol=2e-8, floor_mag=np.nanmin(ref))
        np.testing.assert_allclose(outf[i], r.bandpass_magnitude_to_flux(ref, ref_flux), rtol=1e-7)
        ref_l = r.gaussian_likelihood(y, ref, 0.1)
        assert abs(logl[i] - ref_l) <= 1e-4 + 1e-7 * abs(ref_l), (i, logl[i], ref_l)


def test_errors(rb):
    t = np.array([1., 2., 3.])
    base = dict(f_nickel=0.1, mej=1., vej=5e3, kappa=0.1, kappa_gamma=0.1, temperature_floor=3e3)
    with pytest.raises(KeyError):
        rb.supernova_models.arnett(t, 0.1, output_format="magnitude", **base)           # no bands
    with pytest.raises(KeyError):
        rb.supernova_models.arnett(t, 0.1, output_format="magnitude", bands="nosuchband", **base)
    with pytest.raises(ValueError):
        rb.supernova_models.arnett(t, 0.1, output_format="magnitude", bands="lsstr",
                                   lambda_array=np.geomspace(6000, 60000, 50), **base)    # band outside grid
