"""output_format='spectra' on device (supernova_models.py:1019-1023, kilonova_models.py:1596-1599,
magnetar_driven_ejecta_models.py:231-234): the namedtuple (time, lambdas, spectra) per sample against the oracle's
spectra, i.e. the array the oracle's magnitude branch hands to its spline (captured before the clean-up)."""
import contextlib

import numpy as np
import pytest

import helpers as h
from oracle import restated as r

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def capture_spectra():
    """The oracle's magnitude functions end in bandmag_from_spectra(time, time_eval, spectra, lambda_array, ...):
    return those arguments instead of integrating them."""
    keep = r.bandmag_from_spectra
    r.bandmag_from_spectra = lambda time, time_eval, spectra, lam, bands, bandpasses, **kw: (
        np.asarray(time_eval), np.asarray(lam), np.asarray(spectra))
    try:
        yield
    finally:
        r.bandmag_from_spectra = keep


def _close(got, ref, what, rtol=1e-9, row_rtol=None):
    """`row_rtol` [n_t]: the documented conditioning of the reference's own diffusion integral at late grid times
    (helpers.diffusion_condition, eps (t / tau_d)^2), added per row."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    tol = rtol if row_rtol is None else rtol + np.asarray(row_rtol)[:, None]
    with np.errstate(all="ignore"):
        same = (got == ref) | (np.isnan(got) & np.isnan(ref)) | (np.abs(got - ref) <= tol * np.abs(ref))
        # Entries below 1e-280: the reference forms them in erg/s/Hz/cm^2, 1e9-3e14 below these units, where they are
        # subnormal (a few significant bits) or flush to zero; the device works in output units.  Both must be tiny.
        tiny = np.abs(ref) < 1e-280
        same |= tiny & (np.abs(got) < 1e-270)
        assert np.array_equal(np.isfinite(got), np.isfinite(ref)), (what, "finite pattern")
        assert not np.any((got == 0) & ~tiny), (what, "zero pattern")
    assert same.all(), (what, int((~same).sum()), float(np.nanmax(np.where(same, 0, np.abs(got - ref) / np.abs(ref)))))


@pytest.mark.parametrize("model", ["arnett", "slsn", "type_1a", "csm_interaction"])
def test_supernova_spectra(model):
    import redback_b200 as rb
    rng = np.random.default_rng(2)
    N = 6
    t = np.linspace(5, 100, 7)
    if model == "csm_interaction":
        p = h.csm_prior(rng, N, zmax=0.5)
    elif model == "slsn":
        p = h.slsn_prior(rng, N, zmax=0.5)
    else:
        p = h.arnett_prior(rng, N, zmax=0.5)
    out = getattr(rb.supernova_models, model)(t, params_batch=p, output_format="spectra")
    assert out.spectra.shape == (N, 300, 100) and out.time.shape == (N, 300) and out.lambdas.shape == (100,)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        tt, lam, spec = r.sn_magnitude(model, t, z, spectra_only=True, **kw)
        np.testing.assert_allclose(out.time[i], tt, rtol=1e-15)
        np.testing.assert_array_equal(out.lambdas, lam)
        cond = None
        if model != "csm_interaction":
            tau = np.sqrt(2.0 * r.solar_mass / (13.7 * r.speed_of_light * r.km_cgs) * kw["kappa"] * kw["mej"]
                          / kw["vej"]) / 86400
            cond = h.diffusion_condition(tt / (1 + z), tau)
        _close(out.spectra[i], spec, f"{model} sample {i}", rtol=5e-9 if model == "csm_interaction" else 1e-9,
               row_rtol=cond)
    # scalar call: plain arrays, as the reference returns them
    kw = {k: float(v[0]) for k, v in p.items()}
    one = getattr(rb.supernova_models, model)(t, output_format="spectra", **kw)
    np.testing.assert_array_equal(one.spectra, out.spectra[0])
    np.testing.assert_array_equal(one.time, out.time[0])


@pytest.mark.parametrize("model", ["one_component", "metzger"])
def test_kilonova_spectra(model):
    import redback_b200 as rb
    rng = np.random.default_rng(4)
    N = 5
    t = np.geomspace(0.3, 25, 9)
    p = h.one_component_prior(rng, N) if model == "one_component" else h.metzger_prior(rng, N)
    fn = rb.kilonova_models.one_component_kilonova_model if model == "one_component" else \
        rb.kilonova_models.metzger_kilonova_model
    out = fn(t, params_batch=p, output_format="spectra", lambda_array=np.geomspace(500, 40000, 60))
    assert out.spectra.shape == (N, 500, 60)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        tt, lam, spec = r.kn_magnitude(model, t, z, spectra_only=True, lambda_array=np.geomspace(500, 40000, 60), **kw)
        g = tt.size                                    # np.unique drops the junction duplicates: g <= 500
        np.testing.assert_allclose(out.time[i, :g], tt, rtol=1e-14)
        assert np.all(np.isnan(out.time[i, g:])) and np.all(np.isnan(out.spectra[i, g:]))
        _close(out.spectra[i, :g], spec, f"{model} sample {i}", rtol=2e-9)
    kw = {k: float(v[1]) for k, v in p.items()}
    one = fn(t, output_format="spectra", lambda_array=np.geomspace(500, 40000, 60), **kw)
    assert one.spectra.shape[0] == one.time.size and not np.isnan(one.time).any()
    np.testing.assert_array_equal(one.spectra, out.spectra[1, :one.time.size])


def test_two_component_kilonova_spectra():
    import redback_b200 as rb
    t = np.geomspace(0.5, 5, 6)
    kw = dict(mej_1=0.03, vej_1=0.1, temperature_floor_1=2500.0, kappa_1=3.0, mej_2=0.01, vej_2=0.25,
              temperature_floor_2=4000.0, kappa_2=0.5)
    out = rb.kilonova_models.two_component_kilonova_model(t, redshift=0.02, output_format="spectra", **kw)
    comps = [(kw["mej_1"], kw["vej_1"], kw["kappa_1"], kw["temperature_floor_1"]),           # the oracle's tuple order
             (kw["mej_2"], kw["vej_2"], kw["kappa_2"], kw["temperature_floor_2"])]
    with capture_spectra():
        tt, lam, spec = r.multi_component_kn_magnitude(t, 0.02, comps, bands=None, bandpasses=None,
                                                       dense_resolution=500, grid_t_max=86400 * 6)
    np.testing.assert_allclose(out.time, tt * 86400.0, rtol=1e-14)
    _close(out.spectra, spec, "two_component", rtol=2e-9)


def test_mergernova_and_extinction_spectra():
    import redback_b200 as rb
    rng = np.random.default_rng(8)
    t = np.geomspace(0.2, 30, 8)
    p = h.mergernova_prior(rng, 4, kind="basic")
    out = rb.magnetar_driven_ejecta_models.basic_mergernova(t, params_batch=p, output_format="spectra")
    for i in range(4):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        with capture_spectra():
            tt, lam, spec = r.mergernova_magnitude("basic_mergernova", t, z, bands=None, bandpasses=None, **kw)
        np.testing.assert_allclose(out.time[i], tt * 86400.0, rtol=1e-14)
        _close(out.spectra[i], spec, f"basic_mergernova sample {i}", rtol=1e-8)
    # the reference's extinction wrapper consumes exactly this output (extinction_models.py:223-235): spectra with the
    # attenuation applied on device against perform_extinction of the plain ones
    pa = h.arnett_prior(rng, 3, zmax=0.4)
    plain = rb.supernova_models.arnett(t, params_batch=pa, output_format="spectra")
    dim = rb.extinction_models.extinction_with_supernova_base_model(
        t, base_model="arnett", params_batch=dict(pa, av_host=np.array([0.0, 0.6, 1.9])), output_format="spectra",
        av_mw=0.15, host_law="odonnell94")
    for i, av in enumerate((0.0, 0.6, 1.9)):
        ref = r.perform_extinction(plain.spectra[i], plain.lambdas, av, 3.1, av_mw=0.15, host_law="odonnell94",
                                   redshift=pa["redshift"][i])
        _close(dim.spectra[i], ref, f"extinction sample {i}", rtol=1e-12)
