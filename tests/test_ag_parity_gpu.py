"""GPU parity: redback-native tophat / Gaussian afterglow kernels vs the CPU oracle and the reference's own
golden file test/reference_results/tophat_redback_ref_data.ecsv (tests/golden/tophat_redback_ref.json).

The reference checks the ECSV with np.allclose defaults (rtol 1e-5, atol 1e-8); the file itself deviates from
today's reference code by 3.4e-8, so that comparison uses rtol 1e-6.  Against the oracle
(which reproduces the reference's current code to 1e-15) the tolerance is rtol 1e-9."""
import json
import os

import numpy as np
import pytest

import helpers as h
from oracle import restated as r

pytestmark = pytest.mark.gpu
RTOL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rb():
    import redback_b200
    return redback_b200


def test_reference_ecsv_golden(rb):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "tophat_redback_ref.json")))
    t = np.array(g["time"])
    rows = g["rows"]
    names = ["redshift", "thv", "loge0", "thc", "logn0", "p", "logepse", "logepsb", "g0", "xiN"]
    batch = {k: np.array([row[k] for row in rows]) for k in names}
    out = rb.afterglow_models.tophat_redback(t, params_batch=batch, frequency=2e17, output_format="flux_density")
    ref = np.array([row["results"] for row in rows])
    assert out.shape == ref.shape == (108, 100)
    np.testing.assert_allclose(out, ref, rtol=1e-6, atol=0)


def test_config4_tophat_prior_draws(rb):
    """BASELINE configs[3] shape: radio + X-ray epochs, prior draws (res follows the reference's rule)."""
    t, nu = h.config4_epochs(40)
    rng = np.random.default_rng(7)
    N = 96
    p = h.tophat_prior(rng, N)
    # make sure on-axis / high-resolution samples are present
    p["thv"][:8] = p["thc"][:8] * rng.uniform(0, 1.5, 8)
    fn = rb.afterglow_models.tophat_redback
    out = fn(t, params_batch=p, frequency=nu, output_format="flux_density")
    y = r.redback_afterglow(t, 0.5, 0.1, 52.0, 0.08, -1.0, 2.3, -1.0, -2.0, 300.0, 0.5, frequency=nu)
    sigma = 0.1 * np.abs(y) + 1e-12
    like = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=dict(frequency=nu, output_format="flux_density"))
    logl = like.log_likelihood_batch(p)
    res_hist = {}
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        res = r.ag_default_res(kw["thv"], kw["thc"], kw["g0"])
        res_hist[res] = res_hist.get(res, 0) + 1
        ref = r.redback_afterglow(t, z, frequency=nu, **kw)
        h.assert_close(out[i], ref, RTOL, what=f"sample {i} (res {res})")
        ref_l = r.gaussian_likelihood(y, ref, sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-9 * abs(ref_l), (i, logl[i], ref_l)
    assert max(res_hist) > 10, res_hist


@pytest.mark.parametrize("k,pval,expansion", [(0, 1.8, 1), (2, 2.0, 1), (0, 2.6, 0), (2, 2.4, 1)])
def test_variants(rb, k, pval, expansion):
    t = np.geomspace(0.1, 300, 30)
    nu = np.where(np.arange(30) % 2 == 0, 5e9, 2e17)
    kw = dict(thv=0.2, loge0=52., thc=0.08, logn0=-1., p=pval, logepse=-1., logepsb=-2., g0=300., xiN=0.5)
    out = rb.afterglow_models.tophat_redback(t, 0.5, frequency=nu, output_format="flux_density", k=k,
                                             expansion=expansion, res=14, **kw)
    ref = r.redback_afterglow(t, 0.5, frequency=nu, k=k, expansion=expansion, res=14, **kw)
    h.assert_close(out, ref, RTOL, what=f"tophat k={k} p={pval}")
    outg = rb.afterglow_models.gaussian_redback(t, 0.5, thj=0.25, frequency=nu, output_format="flux_density", k=k,
                                                expansion=expansion, **kw)
    refg = r.redback_afterglow(t, 0.5, frequency=nu, method_id=2, thj=0.25, k=k, expansion=expansion, **kw)
    h.assert_close(outg, refg, RTOL, what=f"gaussian k={k} p={pval}")


@pytest.mark.parametrize("name,method_id,ss,aa", [("twocomponent_redback", 5, 0.01, 4), ("powerlaw_redback", 3, 3, -3),
                                                  ("alternativepowerlaw_redback", 4, 3, 3),
                                                  ("doublegaussian_redback", 6, 0.1, 0.5)])
def test_other_jet_structures(rb, name, method_id, ss, aa):
    """get_structure_numba methods 3-6 (base_models.py:114-146) through their *_redback model functions."""
    t = np.geomspace(0.1, 300, 24)
    nu = np.where(np.arange(24) % 2 == 0, 5e9, 2e17)
    rng = np.random.default_rng(17)
    N = 12
    p = h.tophat_prior(rng, N)
    p["thj"] = p["thc"] * rng.uniform(1.5, 4.0, N)
    p["thv"][:4] = p["thc"][:4] * rng.uniform(0, 1.5, 4)
    out = getattr(rb.afterglow_models, name)(t, params_batch=p, frequency=nu, output_format="flux_density")
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.redback_afterglow(t, z, frequency=nu, method_id=method_id, ss=ss, aa=aa, **kw)
        h.assert_close(out[i], ref, RTOL, what=f"{name} sample {i}")


def test_monochromatic_ab_magnitude(rb):
    """output_format='magnitude' (base_models.py:944-946): AB magnitude of the mJy flux density, also inside the
    fused likelihood."""
    t = np.geomspace(0.1, 100, 20)
    nu = np.full(20, 4.6e14)
    rng = np.random.default_rng(23)
    N = 6
    p = h.tophat_prior(rng, N)
    out = rb.afterglow_models.tophat_redback(t, params_batch=p, frequency=nu, output_format="magnitude")
    y = out[2] + 0.1
    logl = rb.GaussianLikelihood(t, y, 0.2, rb.afterglow_models.tophat_redback,
                                 kwargs=dict(frequency=nu, output_format="magnitude")).log_likelihood_batch(p)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.calc_ABmag_from_flux_density(r.redback_afterglow(t, z, frequency=nu, **kw))
        np.testing.assert_allclose(out[i], ref, rtol=0, atol=2.5 / np.log(10) * RTOL)
        ref_l = r.gaussian_likelihood(y, ref, 0.2)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-9 * abs(ref_l)
    assert abs(r.calc_ABmag_from_flux_density(3631e3) - 0.0) < 1e-4     # 3631 Jy is AB = 0 to 1e-4 mag
    with pytest.raises(NotImplementedError):
        rb.afterglow_models.tophat_redback(t, params_batch=p, frequency=nu, output_format="flux")


def test_errors(rb):
    t = np.array([1., 2.])
    kw = dict(thv=0.2, loge0=52., thc=0.08, logn0=-1., p=2.2, logepse=-1., logepsb=-2., g0=300., xiN=0.5)
    with pytest.raises(ValueError):
        rb.afterglow_models.tophat_redback(t, 0.5, frequency=1e9, output_format="flux_density", k=1, **kw)
    out = rb.afterglow_models.tophat_redback(t, 0.5, frequency=1e9, output_format="flux_density", **dict(kw, p=3.6))
    assert np.all(np.isnan(out))  # the reference raises ValueError("p is out of range"); flagged as NaN per sample


def test_tophat_fp32_variant():
    """dtype='float32' (rb_ag_config.precision = RB_PREC_F32): the cell-step synchrotron spectrum through single-precision
    hardware transcendentals on FP64-reduced arguments.  North-star tolerance of the FP32 path: 1e-5 on the model."""
    import redback_b200 as rb
    rng = np.random.default_rng(41)
    t, nu = h.config4_epochs(40)
    N = 400
    p = h.tophat_prior(rng, N)
    fn = rb.afterglow_models.tophat_redback
    f64 = fn(t, params_batch=p, frequency=nu, output_format="flux_density")
    f32 = fn(t, params_batch=p, frequency=nu, output_format="flux_density", dtype="float32")
    assert np.array_equal(np.isnan(f32), np.isnan(f64))
    ok = np.isfinite(f64) & (f64 > 0)
    rel = np.abs(f32[ok] - f64[ok]) / f64[ok]
    assert rel.max() < 1e-5, rel.max()
    assert rel.max() > 1e-12                      # the variant really ran
    y = f64[0] * 1.02
    kw = dict(frequency=nu, output_format="flux_density")
    l64 = rb.GaussianLikelihood(t, y, 0.1 * np.abs(y) + 1e-12, fn, kwargs=kw).log_likelihood_batch(p)
    l32 = rb.GaussianLikelihood(t, y, 0.1 * np.abs(y) + 1e-12, fn, kwargs=dict(kw, dtype="float32")).log_likelihood_batch(p)
    assert np.all(np.abs(l32 - l64) <= 1e-4 + 1e-4 * np.abs(l64))
