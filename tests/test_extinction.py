"""Dust extinction wrappers (redback/transient_models/extinction_models.py).

CPU part: the oracle's laws against the `extinction` package's documented example values (the package is a third-party
dependency, absent here) and the papers' fixed points; the product's host-side Fitzpatrick spline table against the
oracle; the properties the reference's own test checks (test/extinction_test.py:92-141); the wrapper's error behaviour.
GPU part: device flux-density and magnitude branches against the oracle."""
import numpy as np
import pytest

import helpers as h
from oracle import restated as r

WAVE_DOC = np.array([2000.0, 4000.0, 8000.0])


def test_laws_reproduce_the_documented_example_values():
    # README / API documentation of kbarbary/extinction (a_v = 1, r_v = 3.1), printed to 8 decimals
    np.testing.assert_allclose(r.extinction_ccm89(WAVE_DOC, 1.0, 3.1), [2.84252644, 1.4645557, 0.59748901], atol=5e-9)
    np.testing.assert_allclose(r.extinction_fitzpatrick99(WAVE_DOC, 1.0, 3.1), [2.76225609, 1.42325373, 0.55333671],
                               atol=5e-9)


def test_laws_fixed_points_from_the_papers():
    v = np.array([1e4 / 1.82])
    for law in (r.extinction_ccm89, r.extinction_odonnell94):
        assert abs(law(v, 0.7, 2.3)[0] - 0.7) < 1e-15                  # a(1.82) = 1, b(1.82) = 0
    # Calzetti 2000: k(V = 5500 A) = R_V = 4.05 to the paper's rounding
    assert abs(r.extinction_calzetti00(np.array([5500.0]), 1.0)[0] - 1.0) < 2e-3
    # the two Calzetti branches meet at 0.63 micron, the CCM89 pieces at 1.1 and 3.3 / micron
    eps = 1e-9
    for law, x0 in ((r.extinction_calzetti00, 1 / 0.63), (r.extinction_ccm89, 1.1), (r.extinction_ccm89, 3.3),
                    (r.extinction_odonnell94, 3.3)):
        lo, hi = law(np.array([1e4 / (x0 - eps)]), 1.0, 3.1)[0], law(np.array([1e4 / (x0 + eps)]), 1.0, 3.1)[0]
        assert abs(lo - hi) < 2e-2, (law.__name__, x0, lo, hi)
    # Fitzpatrick 1999 passes through its anchors: A(lambda -> inf) = 0 and the UV curve joins at 2700 A
    assert abs(r.extinction_fitzpatrick99(np.array([1e12]), 1.0, 3.1)[0]) < 1e-6
    j = r.extinction_fitzpatrick99(np.array([2700.0 * (1 + 1e-12), 2700.0 * (1 - 1e-12)]), 1.0, 3.1)
    assert abs(j[0] - j[1]) < 1e-9


def test_reference_properties_of_perform_extinction():
    """test/extinction_test.py:92-141 restated on the oracle."""
    ang = np.linspace(3000, 8000, 20)
    flux = np.ones(20)
    kw = dict(redshift=0.1)
    np.testing.assert_array_equal(r.perform_extinction(flux, ang, 0.0, 3.1, av_mw=0.0, **kw), flux)
    host = r.perform_extinction(flux, ang, 1.0, 3.1, **kw)
    assert np.all(host < flux)
    mw = r.perform_extinction(flux, ang, 0.0, 3.1, av_mw=1.0, **kw)
    assert np.all(mw < flux)
    both = r.perform_extinction(flux, ang, 0.5, 3.1, av_mw=0.5, **kw)
    assert np.all(both <= r.perform_extinction(flux, ang, 0.5, 3.1, **kw))
    assert np.array_equal(flux, np.ones(20))
    with pytest.raises(ValueError):
        r.perform_extinction(flux, ang, 1.0, 3.1, host_law="bad_law", **kw)
    with pytest.raises(ValueError):
        r.perform_extinction(flux, ang, 0.0, 3.1, av_mw=1.0, mw_law="bad_law", **kw)
    # the 10-mag cap (:137-139): far-UV rest wavelengths at a modest A_V are left unextinguished
    capped = r.perform_extinction(np.ones(2), np.array([300.0, 5000.0]), 3.0, 3.1, host_law="ccm89", redshift=0.0)
    assert capped[0] == 1.0 and capped[1] < 0.1


@pytest.mark.parametrize("rv", [2.3, 3.1, 4.05, 5.5])
def test_host_spline_table_matches_the_oracle(rv):
    from redback_b200 import extinction_models as em
    tab = em.f99_spline_table(rv)
    wave = np.geomspace(2701.0, 2e6, 4000)
    x = 1e4 / wave
    i = np.clip(np.searchsorted(tab[:9], x, side="right") - 1, 0, 7)
    c = tab[9:].reshape(8, 4)[i]
    dx = x - tab[i]
    k = ((c[:, 0] * dx + c[:, 1]) * dx + c[:, 2]) * dx + c[:, 3]
    np.testing.assert_allclose((k + rv) / rv, r.extinction_fitzpatrick99(wave, 1.0, rv), rtol=0, atol=1e-13)


def test_wrapper_errors_like_the_reference():
    """test/extinction_test.py:45-84, :181-186 (no GPU needed: the failures precede any launch)."""
    import redback_b200 as rb
    em = rb.extinction_models
    assert em._get_correct_function(rb.supernova_models.arnett) is rb.supernova_models.arnett
    assert em._get_correct_function("arnett", "supernova") is em._get_correct_function("arnett")
    for args in (("nonexistent_model", "supernova"), ("nonexistent_model", None), ("arnett", "invalid_type"),
                 (42, "supernova")):
        with pytest.raises(ValueError):
            em._get_correct_function(*args)
    t = np.linspace(1, 100, 10)
    kw = dict(redshift=0.1, f_nickel=0.1, mej=1.0, output_format="flux_density", frequency=3e14, kappa=0.1,
              kappa_gamma=10.0, temperature_floor=3000.0, vej=10000.0)
    with pytest.raises(ValueError):
        em.extinction_with_supernova_base_model(t, av_host=0.5, base_model="not_a_model", **kw)
    with pytest.raises(ValueError):
        em.extinction_with_supernova_base_model(t, av_host=0.5, base_model="arnett", host_law="bad_law", **kw)
    with pytest.raises(NotImplementedError):
        em.extinction_with_supernova_base_model(t, av_host=0.5, base_model="arnett", host_law="fm07", **kw)
    for fn in em.MODELS:
        assert rb.all_models_dict[fn.__name__] is fn


# ---------------------------------------------------------------------------------------------------------------------
LAWS = ["fitzpatrick99", "ccm89", "odonnell94", "calzetti00"]


@pytest.mark.gpu
def test_zero_extinction_is_the_base_model_and_scalar_call():
    """test/extinction_test.py:150-176, :196-213: av_host = 0 reproduces the bare model; shape; finiteness; dimming."""
    import redback_b200 as rb
    t = np.linspace(1, 100, 10)
    kw = dict(redshift=0.1, f_nickel=0.1, mej=1.0, output_format="flux_density", frequency=3e14, kappa=0.1,
              kappa_gamma=10.0, temperature_floor=3000.0, vej=10000.0)
    base = rb.supernova_models.arnett(t, **kw)
    z0 = rb.extinction_models.extinction_with_supernova_base_model(t, av_host=0.0, base_model="arnett", **kw)
    np.testing.assert_allclose(z0, base, rtol=1e-10)
    fn0 = rb.extinction_models.extinction_with_function(t, av_host=0.0, base_model=rb.supernova_models.arnett, **kw)
    np.testing.assert_allclose(fn0, base, rtol=1e-10)
    dim = rb.extinction_models.extinction_with_supernova_base_model(t, av_host=0.5, base_model="arnett", **kw)
    assert dim.shape == t.shape and np.all(np.isfinite(dim)) and np.all(dim <= base)
    ref = r.perform_extinction(r.arnett(t, **kw), r.nu_to_lambda(np.full(t.size, 3e14)), 0.5, 3.1, redshift=0.1)
    np.testing.assert_allclose(dim, ref, rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("host_law,mw_law", [(a, b) for a in LAWS for b in ("fitzpatrick99", "ccm89")])
def test_flux_density_branch_against_the_oracle(host_law, mw_law):
    """slsn multiband flux density with sampled av_host / redshift, radio to far-UV frequencies (every piece of every law,
    incl. the 10-mag cap at the far-UV end)."""
    import redback_b200 as rb
    rng = np.random.default_rng(5)
    N, E = 300, 60
    t = np.sort(rng.uniform(2, 200, E))
    freq = 2.99792458e18 / np.geomspace(700.0, 4e4, E)[rng.permutation(E)]
    p = h.slsn_prior(rng, N, zmax=1.5)
    p["av_host"] = rng.uniform(0, 3, N)
    p["av_host"][:5] = 0.0
    fixed = dict(output_format="flux_density", frequency=freq, av_mw=0.31, rv_host=2.7, rv_mw=3.1, host_law=host_law,
                 mw_law=mw_law)
    out = rb.extinction_models.extinction_with_supernova_base_model(t, base_model="slsn", params_batch=p, **fixed)
    assert out.shape == (N, E)
    # the base model's own parity is tests/test_sn_parity_gpu.py's business: isolate the attenuation factor
    base = rb.supernova_models.slsn(t, params_batch={k: v for k, v in p.items() if k != "av_host"},
                                    output_format="flux_density", frequency=freq)
    lam = r.nu_to_lambda(freq)
    for i in range(N):
        ref = r.perform_extinction(base[i], lam, p["av_host"][i], 2.7, av_mw=0.31, rv_mw=3.1, host_law=host_law,
                                   mw_law=mw_law, redshift=p["redshift"][i])
        h.assert_close(out[i], ref, 1e-12, what=f"{host_law}/{mw_law} sample {i}")
    assert np.array_equal(out[:5] < base[:5], np.ones((5, E), bool) & (base[:5] > 0))      # av_host = 0: Milky Way only


@pytest.mark.gpu
@pytest.mark.parametrize("host_law", LAWS)
def test_magnitude_branch_against_the_oracle(host_law):
    """The reference attenuates the spectra and then runs the bandpass photometry (:223-251); here the photometry kernel
    carries the factor per wavelength column."""
    import redback_b200 as rb
    rb.bands.register_synthetic_lsst()
    obands = {n: r.Bandpass(n, b.wave, b.trans) for n, b in
              ((n, rb.bands.get_bandpass(n)) for n in rb.bands.SYNTHETIC_LSST_EDGES)}
    rng = np.random.default_rng(17)
    N = 12
    t = np.sort(rng.uniform(3, 120, 36))
    names = np.array(list(h.LSST_NU)[:6] * 6)
    p = h.arnett_prior(rng, N, zmax=0.5)
    p["temperature_floor"] = h.loguniform(rng, 4e3, 1.2e4, N)
    p["av_host"] = rng.uniform(0, 2, N)
    p["av_host"][0] = 0.0
    out = rb.extinction_models.extinction_with_supernova_base_model(
        t, base_model="arnett", params_batch=p, bands=names, output_format="magnitude", av_mw=0.2, host_law=host_law)
    plain = rb.supernova_models.arnett(t, params_batch={k: v for k, v in p.items() if k != "av_host"}, bands=names,
                                       output_format="magnitude")
    assert out.shape == (N, t.size) and np.all(out >= plain - 1e-12)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        av, z = kw.pop("av_host"), kw.pop("redshift")

        def dim(spectra, lam, av=av, z=z):
            return r.perform_extinction(spectra, lam, av, 3.1, av_mw=0.2, host_law=host_law, redshift=z)
        ref = r.sn_magnitude("arnett", t, z, bands=names, bandpasses=obands, spectra_transform=dim, **kw)
        ok = np.isfinite(ref) & (ref < np.nanmin(ref) + 25)
        assert np.max(np.abs(out[i][ok] - ref[ok])) < 5e-9, (i, np.max(np.abs(out[i][ok] - ref[ok])))


@pytest.mark.gpu
def test_kilonova_and_afterglow_wrappers():
    import redback_b200 as rb
    rng = np.random.default_rng(3)
    t = np.sort(rng.uniform(0.5, 20, 30))
    freq = 2.99792458e18 / rng.uniform(3000, 20000, t.size)
    N = 40
    p = dict(redshift=rng.uniform(0.01, 0.3, N), mej=rng.uniform(0.01, 0.05, N), vej=rng.uniform(0.1, 0.3, N),
             kappa=rng.uniform(1, 10, N), av_host=rng.uniform(0, 1.5, N))
    out = rb.extinction_models.extinction_with_kilonova_base_model(
        t, base_model="one_component_kilonova_model", params_batch=p, frequency=freq, output_format="flux_density",
        temperature_floor=4000.0)
    for i in range(0, N, 4):
        base = r.one_component_kilonova_model(t, p["redshift"][i], p["mej"][i], p["vej"][i], p["kappa"][i],
                                              frequency=freq, temperature_floor=4000.0)
        ref = r.perform_extinction(base, r.nu_to_lambda(freq), p["av_host"][i], 3.1, redshift=p["redshift"][i])
        np.testing.assert_allclose(out[i], ref, rtol=2e-9)
    # dust-to-gas wrapper: av = 10**lognh / (factor 1e21), the afterglow base model
    ta = np.geomspace(0.1, 50, 12)
    fa = np.where(np.arange(12) % 2 == 0, 4.6e14, 2.4e17)
    akw = dict(redshift=0.5, thv=0.1, loge0=52.0, thc=0.08, logn0=0.0, p=2.3, logepse=-1.0, logepsb=-2.0, g0=300.0,
               xiN=1.0, frequency=fa, output_format="flux_density", base_model="tophat_redback")
    got = rb.extinction_models.extinction_afterglow_galactic_dust_to_gas_ratio(ta, lognh=21.3, factor=2.21, **akw)
    base = rb.afterglow_models.tophat_redback(ta, **{k: v for k, v in akw.items() if k != "base_model"})
    ref = r.perform_extinction(base, r.nu_to_lambda(fa), 10 ** 21.3 / 2.21e21, 3.1, redshift=0.5)
    np.testing.assert_allclose(got, ref, rtol=1e-12)


@pytest.mark.gpu
def test_t0_with_extinction():
    """phase_models._t0_with_extinction (:57-86): a sampled explosion epoch inside the supernova kernel, a scalar one
    shifted on the host for the afterglow; epochs before t0 -> 0 (5000 mag)."""
    import redback_b200 as rb
    rng = np.random.default_rng(11)
    N, E = 30, 40
    mjd = np.sort(rng.uniform(59000, 59120, E))
    freq = 2.99792458e18 / rng.uniform(3500, 9000, E)
    p = h.arnett_prior(rng, N, zmax=0.8)
    p["t0"] = rng.uniform(58990, 59030, N)
    p["av_host"] = rng.uniform(0, 2, N)
    out = rb.phase_models.t0_supernova_extinction(mjd, base_model="arnett", params_batch=p, frequency=freq,
                                                  output_format="flux_density", av_mw=0.1, mw_law="ccm89")
    plain = rb.phase_models.t0_base_model(mjd, base_model="arnett", frequency=freq, output_format="flux_density",
                                          params_batch={k: v for k, v in p.items() if k != "av_host"})
    for i in range(N):
        ref = r.perform_extinction(plain[i], r.nu_to_lambda(freq), p["av_host"][i], 3.1, av_mw=0.1, mw_law="ccm89",
                                   redshift=p["redshift"][i])
        h.assert_close(out[i], ref, 1e-12, what=f"sample {i}")
        assert np.all(out[i][mjd < p["t0"][i]] == 0.0)
    # magnitudes: the placeholder is 5000 (phase_models.py:81)
    rb.bands.register_synthetic_lsst()
    names = np.array(list(h.LSST_NU)[:6] * 7)[:E]
    mag = rb.phase_models.t0_supernova_extinction(mjd, base_model="arnett", params_batch=p, bands=names,
                                                  output_format="magnitude")
    ext = rb.extinction_models.extinction_with_supernova_base_model
    for i in range(0, N, 7):
        before = mjd < p["t0"][i]
        assert np.all(mag[i][before] == 5000.0)
        kw = {k: v[i] for k, v in p.items() if k not in ("t0",)}
        ref = ext(mjd[~before] - p["t0"][i], base_model="arnett", bands=names[~before], output_format="magnitude", **kw)
        ok = np.isfinite(ref) & (ref < np.nanmin(ref) + 25)
        assert np.max(np.abs(mag[i][~before][ok] - ref[ok])) < 1e-8
    # afterglow: scalar t0 on the host
    ta = 59000.0 + np.geomspace(0.1, 50, 12) - 0.5
    fa = np.where(np.arange(12) % 2 == 0, 4.6e14, 2.4e17)
    akw = dict(redshift=0.5, thv=0.1, loge0=52.0, thc=0.08, logn0=0.0, p=2.3, logepse=-1.0, logepsb=-2.0, g0=300.0,
               xiN=1.0, output_format="flux_density", base_model="tophat_redback")
    got = rb.phase_models.t0_afterglow_extinction(ta, t0=59000.0, av_host=0.7, frequency=fa, **akw)
    good = ta - 59000.0 >= 0
    assert np.all(got[~good] == 0) and (~good).sum() > 0
    ref = rb.extinction_models.extinction_with_afterglow_base_model(ta[good] - 59000.0, av_host=0.7, frequency=fa[good],
                                                                    **akw)
    np.testing.assert_array_equal(got[good], ref)


@pytest.mark.gpu
def test_t0_base_model_scalar_t0_for_models_without_a_fused_epoch():
    """phase_models.t0_base_model (:19-59) around an afterglow and the CSM model: a scalar t0 is shifted on the host
    (epochs before t0 -> 0), a sampled t0 is refused for these base models."""
    import redback_b200 as rb
    ta = 59000.0 + np.geomspace(0.1, 50, 12) - 0.5
    fa = np.where(np.arange(12) % 2 == 0, 4.6e14, 2.4e17)
    akw = dict(redshift=0.5, thv=0.1, loge0=52.0, thc=0.08, logn0=0.0, p=2.3, logepse=-1.0, logepsb=-2.0, g0=300.0,
               xiN=1.0, output_format="flux_density")
    got = rb.phase_models.t0_base_model(ta, t0=59000.0, base_model="tophat_redback", frequency=fa, **akw)
    good = ta - 59000.0 >= 0
    assert (~good).sum() > 0 and np.all(got[~good] == 0)
    ref = rb.afterglow_models.tophat_redback(ta[good] - 59000.0, frequency=fa[good], **akw)
    np.testing.assert_array_equal(got[good], ref)
    rng = np.random.default_rng(9)
    p = h.csm_prior(rng, 5, zmax=0.3)
    tc = 58000.0 + np.linspace(-3, 60, 10)
    out = rb.phase_models.t0_base_model(tc, t0=58000.0, base_model="csm_interaction", params_batch=p,
                                        frequency=np.full(10, 5e14), output_format="flux_density", temperature_floor=5000.0)
    assert out.shape == (5, 10) and np.all(out[:, tc < 58000.0] == 0)
    with pytest.raises(NotImplementedError):
        rb.phase_models.t0_base_model(tc, base_model="csm_interaction", params_batch=dict(p, t0=np.full(5, 58000.0)),
                                      frequency=np.full(10, 5e14), output_format="flux_density", temperature_floor=5000.0)
