"""GPU parity: fused kilonova kernels vs the CPU oracle (one-component and Metzger models).

Tolerance: rtol 1e-9 on flux densities, widened by the documented conditioning of the reference's own
engine term (helpers.engine_atan_condition) times the blackbody's temperature sensitivity; |d logL| < 1e-6
or 1e-9 relative."""
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


def _tol_extra(t_obs, z, nu, temp_floor=100.0):
    # d ln F / d ln L <= 0.5 + (h nu / k T)/4 with T >= a few 1e3 K at these epochs; bound it generously
    amp = 0.5 + 0.25 * (6.626e-27 * nu * (1 + z)) / (1.381e-16 * 1500.0)
    return amp * h.engine_atan_condition(t_obs * 86400 / (1 + z))


def test_golden_vectors(rb, golden):
    for name, fn in [("one_component_kilonova_model", rb.kilonova_models.one_component_kilonova_model),
                     ("metzger_kilonova_model", rb.kilonova_models.metzger_kilonova_model),
                     ("two_component_kilonova_model", rb.kilonova_models.two_component_kilonova_model),
                     ("three_component_kilonova_model", rb.kilonova_models.three_component_kilonova_model)]:
        e = golden["kilonova." + name]
        t = np.array(e["time"])
        for params, res in zip(e["params"], e["results"]):
            out = fn(t, **dict(params, **e["eval_kwargs"]))
            h.assert_close(out, h.golden_array(res), RTOL, what=name)


def test_ejecta_relation_models(rb, golden):
    """one_component_ejecta_relation(_projection) and one_component_nsbh_ejecta_relation (kilonova_models.py:1309-1366,
    1470-1493): the reference's golden vectors through the scalar call, and a batch through params_batch."""
    for name in ("one_component_ejecta_relation", "one_component_ejecta_relation_projection",
                 "one_component_nsbh_ejecta_relation"):
        e = golden["kilonova." + name]
        t = np.array(e["time"])
        fn = getattr(rb.kilonova_models, name)
        for params, res in zip(e["params"], e["results"]):
            out = fn(t, **params, **e["eval_kwargs"])
            h.assert_close(out, h.golden_array(res), RTOL, what=name)
        batch = {k: np.array([p[k] for p in e["params"]]) for k in e["params"][0]}
        kw = dict(e["eval_kwargs"])
        batch["redshift"] = np.full(len(e["params"]), kw.pop("redshift"))
        outb = fn(t, params_batch=batch, **kw)
        for i, res in enumerate(e["results"]):
            h.assert_close(outb[i], h.golden_array(res), RTOL, what=name + " batch")
        assert name in rb.all_models_dict


def test_two_component_ejecta_relation_models(rb):
    """two_component_bns_ejecta_relation / two_component_nsbh_ejecta_relation (kilonova_models.py:1369-1397, 1496-1527):
    batch and scalar calls against the oracle (whose ejecta relations are pinned to the reference's own classes)."""
    rng = np.random.default_rng(123)
    t = np.geomspace(0.3, 5.0, 20)
    nu = np.where(np.arange(t.size) % 2 == 0, 6.3e14, 4.7e14)
    N = 6
    common = dict(redshift=rng.uniform(0.005, 0.1, N), zeta=rng.uniform(0.05, 0.5, N), vej_2=rng.uniform(0.05, 0.2, N),
                  kappa_1=rng.uniform(1, 10, N), kappa_2=rng.uniform(0.5, 3, N), tf_1=rng.uniform(1e3, 4e3, N),
                  tf_2=rng.uniform(1e3, 4e3, N))
    cases = (("two_component_bns_ejecta_relation",
              dict(mass_1=rng.uniform(1.35, 1.8, N), mass_2=rng.uniform(1.1, 1.35, N), lambda_1=rng.uniform(100, 600, N),
                   lambda_2=rng.uniform(300, 1500, N), mtov=rng.uniform(2.0, 2.3, N))),
             ("two_component_nsbh_ejecta_relation",
              dict(mass_bh=rng.uniform(3, 6, N), mass_ns=rng.uniform(1.2, 1.6, N), chi_bh=rng.uniform(0.2, 0.9, N),
                   lambda_ns=rng.uniform(300, 1500, N))))
    for name, own in cases:
        p = dict(common, **own)
        fn = getattr(rb.kilonova_models, name)
        out = fn(t, params_batch=p, frequency=nu, output_format="flux_density")
        for i in range(N):
            kwi = {k: float(v[i]) for k, v in p.items()}
            z = kwi.pop("redshift")
            ref = getattr(r, name)(t, z, frequency=nu, **kwi)
            h.assert_close(out[i], ref, RTOL, _tol_extra(t, z, nu), what=f"{name} sample {i}")
        one = fn(t, frequency=nu, output_format="flux_density", **{k: float(v[0]) for k, v in p.items()})
        np.testing.assert_allclose(one, out[0], rtol=1e-12)
        assert name in rb.all_models_dict


def test_t0_base_model_kilonova(rb):
    """phase_models.t0_base_model (phase_models.py:19-59) over the kilonova models: per-sample merger epochs in one
    launch; epochs before t0 return 0 (1000 mag); each sample's adaptive grid follows ITS valid epochs."""
    rng = np.random.default_rng(91)
    mjd0 = 59000.0
    t = mjd0 + np.sort(rng.uniform(0.5, 25, 30))
    nu = np.where(np.arange(t.size) % 2 == 0, 6.3e14, 4.7e14)
    N = 8
    p = h.one_component_prior(rng, N)
    p["t0"] = mjd0 + rng.uniform(-3, 8, N)
    p["t0"][1] = t[-1] - 1e-2                              # only the last epoch is valid
    fn = rb.phase_models.t0_base_model
    kw = dict(base_model="one_component_kilonova_model", frequency=nu, output_format="flux_density")
    out = fn(t, params_batch=p, **kw)
    y = np.where(np.isfinite(out[0]), out[0], 0.0) * 1.02
    sigma = 0.1 * np.abs(y) + 1e-6
    logl = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=kw).log_likelihood_batch(p)
    for i in range(N):
        kwi = {k: v[i] for k, v in p.items()}
        z, t0 = kwi.pop("redshift"), kwi.pop("t0")
        ref = r.t0_base_model(t, t0, lambda tt, frequency: r.one_component_kilonova_model(tt, z, frequency=frequency, **kwi),
                              frequency=nu)
        assert np.all(out[i][t - t0 < 0] == 0.0)
        h.assert_close(out[i], ref, RTOL, _tol_extra(np.maximum(t - t0, 1e-3), z, nu), what=f"t0(one_component) {i}")
        ref_l = r.gaussian_likelihood(y, ref, sigma)
        if np.isfinite(ref_l):
            assert abs(logl[i] - ref_l) <= 1e-6 + 1e-8 * abs(ref_l), (i, logl[i], ref_l)
    # Metzger model, and the magnitude branch (1000 mag before t0)
    pm = h.metzger_prior(rng, 4)
    pm["t0"] = mjd0 + rng.uniform(-2, 5, 4)
    outm = fn(t, params_batch=pm, base_model="metzger_kilonova_model", frequency=nu, output_format="flux_density")
    for i in range(4):
        kwi = {k: v[i] for k, v in pm.items()}
        z, t0 = kwi.pop("redshift"), kwi.pop("t0")
        ref = r.t0_base_model(t, t0, lambda tt, frequency: r.metzger_kilonova_model(tt, z, frequency=frequency, **kwi),
                              frequency=nu)
        h.assert_close(outm[i], ref, RTOL, _tol_extra(np.maximum(t - t0, 1e-3), z, nu), what=f"t0(metzger) {i}")
    rb.bands.register_synthetic_lsst()
    bands = np.array(list(h.LSST_NU)[:6] * 5)
    mag = fn(t, params_batch=p, base_model="one_component_kilonova_model", bands=bands, output_format="magnitude")
    plain = rb.kilonova_models.one_component_kilonova_model
    for i in range(N):
        kwi = {k: float(v[i]) for k, v in p.items()}
        t0 = kwi.pop("t0")
        good = t - t0 >= 0
        assert np.all(mag[i][~good] == 1000.0)
        want = plain(t[good] - t0, bands=bands[good], output_format="magnitude", **kwi)
        ok = np.isfinite(want) & (want < 60)
        assert np.max(np.abs(mag[i][good][ok] - want[ok])) < 1e-8


def test_config3_one_component_flux_density(rb):
    """BASELINE configs[2] shape (flux_density branch): 200 epochs cycling LSST ugrizy, prior draws."""
    t, nu = h.config3_epochs()
    rng = np.random.default_rng(3)
    N = 200
    p = h.one_component_prior(rng, N)
    fn = rb.kilonova_models.one_component_kilonova_model
    out = fn(t, params_batch=p, frequency=nu, output_format="flux_density")
    assert out.shape == (N, t.size)
    y = r.one_component_kilonova_model(t, 0.05, 0.03, 0.3, 10.0, frequency=nu, temperature_floor=2000.0)
    sigma = 0.1 * np.abs(y) + 1e-9
    like = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=dict(frequency=nu, output_format="flux_density"))
    logl = like.log_likelihood_batch(p)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.one_component_kilonova_model(t, z, frequency=nu, **kw)
        h.assert_close(out[i], ref, RTOL, _tol_extra(t, z, nu), what=f"sample {i}")
        ref_l = r.gaussian_likelihood(y, ref, sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-8 * abs(ref_l), (i, logl[i], ref_l)


def test_one_component_fp32_variant(rb):
    """dtype='float32' (rb_kn_config.precision = RB_PREC_F32): the smooth per-grid-point factors in FP32.  North-star
    tolerance of the FP32 path: 1e-5 relative on the model (0.001 mag); NaN pattern as in FP64."""
    t, nu = h.config3_epochs()
    rng = np.random.default_rng(31)
    N = 2000
    p = h.one_component_prior(rng, N)
    fn = rb.kilonova_models.one_component_kilonova_model
    f64 = fn(t, params_batch=p, frequency=nu, output_format="flux_density")
    f32 = fn(t, params_batch=p, frequency=nu, output_format="flux_density", dtype="float32")
    assert np.array_equal(np.isnan(f32), np.isnan(f64))
    # The FP32 factors carry ~3e-7 into L_bol, i.e. ~1e-7 into T; the blackbody amplifies a temperature error by
    # x = h nu / k T, which is O(1) around the peak of a light curve and > 100 deep in its Wien tail.  So: 1e-5 within a
    # factor 1e4 of each sample's peak flux (10 mag), and the north star's 0.001 mag (9.2e-4) on every finite epoch.
    ok = np.isfinite(f64) & (f64 > 0)
    with np.errstate(all="ignore"):
        rel = np.where(ok, np.abs(f32 - f64) / f64, 0.0)
        peak = np.nanmax(np.where(ok, f64, 0.0), axis=1, keepdims=True)
    bright = ok & (f64 >= 1e-4 * peak)
    assert rel[bright].max() < 1e-5, rel[bright].max()
    assert rel.max() < 9.2e-4, rel.max()
    assert rel.max() > 1e-9                       # the variant really ran (not the FP64 path under another name)
    for i in range(0, N, 200):                    # and against the oracle directly
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.one_component_kilonova_model(t, z, frequency=nu, **kw)
        good = np.isfinite(ref) & (ref >= 1e-4 * np.nanmax(ref))
        assert np.max(np.abs(f32[i][good] - ref[good]) / ref[good]) < 1e-5
    with pytest.raises(NotImplementedError):
        rb.kilonova_models.metzger_kilonova_model(t, params_batch=h.metzger_prior(rng, 4), frequency=nu,
                                                  output_format="flux_density", dtype="float32")


def test_metzger_flux_density(rb):
    t, nu = h.config3_epochs(60)
    rng = np.random.default_rng(4)
    N = 48
    p = h.metzger_prior(rng, N)
    fn = rb.kilonova_models.metzger_kilonova_model
    out = fn(t, params_batch=p, frequency=nu, output_format="flux_density")
    out_noprec = fn(t, params_batch=p, frequency=nu, output_format="flux_density", neutron_precursor_switch=False)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.metzger_kilonova_model(t, z, frequency=nu, **kw)
        h.assert_close(out[i], ref, RTOL, _tol_extra(t, z, nu), what=f"sample {i}")
        ref2 = r.metzger_kilonova_model(t, z, frequency=nu, neutron_precursor_switch=False, **kw)
        h.assert_close(out_noprec[i], ref2, RTOL, _tol_extra(t, z, nu), what=f"no-precursor sample {i}")


def test_scalar_call_and_small_resolution(rb):
    t = np.array([0.5, 1.0, 2.0, 4.0, 8.0])
    kw = dict(frequency=4.8e14, output_format="flux_density", temperature_floor=3000.0, dense_resolution=120)
    out = rb.kilonova_models.one_component_kilonova_model(t, 0.02, 0.02, 0.2, 5.0, **kw)
    ref = r.one_component_kilonova_model(t, 0.02, 0.02, 0.2, 5.0, frequency=np.full(5, 4.8e14),
                                         temperature_floor=3000.0, dense_resolution=120)
    h.assert_close(out, ref, RTOL, 1e-9)
    with pytest.raises(ValueError):
        rb.kilonova_models.one_component_kilonova_model(np.array([0.0, 1.0]), 0.02, 0.02, 0.2, 5.0, **kw)


def test_multi_component_batches(rb):
    """two / three_component_kilonova_model (kilonova_models.py:1113-1260): components summed on device, likelihood
    through the stand-alone reduction."""
    rng = np.random.default_rng(57)
    t = np.geomspace(0.2, 5.0, 25)                    # the two-component grid ends at 6 d (source frame)
    nu = np.array(list(h.LSST_NU.values()))[np.arange(t.size) % 6]
    N = 10
    p = dict(redshift=rng.uniform(0.005, 0.1, N))
    for c in (1, 2, 3):
        one = h.one_component_prior(rng, N)
        for k in ("mej", "vej", "kappa", "temperature_floor"):
            p[f"{k}_{c}"] = one[k]
    for fn, ofn, ncomp in ((rb.kilonova_models.two_component_kilonova_model, r.two_component_kilonova_model, 2),
                           (rb.kilonova_models.three_component_kilonova_model, r.three_component_kilonova_model, 3)):
        pc = {k: v for k, v in p.items() if k == "redshift" or int(k[-1]) <= ncomp}
        kw = dict(frequency=nu, output_format="flux_density")
        out = fn(t, params_batch=pc, **kw)
        y = out[0] * 1.02
        sigma = 0.1 * np.abs(out[0]) + 1e-12
        logl = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=kw).log_likelihood_batch(pc)
        for i in range(N):
            kwi = {k: v[i] for k, v in pc.items()}
            z = kwi.pop("redshift")
            ref = ofn(t, z, frequency=nu, **kwi)
            h.assert_close(out[i], ref, RTOL, _tol_extra(t, z, nu), what=f"{fn.__name__} sample {i}")
            ref_l = r.gaussian_likelihood(y, ref, sigma)
            assert abs(logl[i] - ref_l) <= 1e-6 + 1e-8 * abs(ref_l), (i, logl[i], ref_l)
    late = rb.kilonova_models.two_component_kilonova_model(
        np.array([1.0, 9.0]), **{k: float(v[0]) for k, v in p.items() if k == "redshift" or int(k[-1]) <= 2},
        frequency=4e14, output_format="flux_density")
    assert np.isfinite(late[0]) and np.isnan(late[1])        # 9 d is beyond the 6 d grid: interp1d raises in the reference
