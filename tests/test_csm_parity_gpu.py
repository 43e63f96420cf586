"""CSM-interaction path: csm_interaction(_bolometric) = _csm_engine + CSMDiffusion (cumulative)
+ TemperatureFloor + Blackbody, device vs the reference's own outputs (tests/golden/csm_reference_draws.json) and vs
the oracle on prior draws."""
import json
import os

import numpy as np
import pytest

import helpers as h
from oracle import restated as r

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-9


@pytest.fixture(scope="module")
def rb():
    import redback_b200
    return redback_b200


def test_reference_fixture_bolometric(rb):
    with open(os.path.join(ROOT, "tests", "golden", "csm_reference_draws.json")) as fh:
        fx = json.load(fh)
    t_all = np.array(fx["time"])
    for row in fx["draws"]:
        t = t_all[row["skip"]:]
        p = dict(row["params"])
        p.pop("temperature_floor")
        out = rb.supernova_models.csm_interaction_bolometric(t, **p)
        h.assert_close(out, np.array(row["lbol"]), RTOL, what="csm_interaction_bolometric vs reference")


def test_reference_fixture_quadrature_method_and_csm_nickel(rb):
    """csm_diffusion_method='quadrature' (interaction_processes.py:199-227) and csm_nickel(_bolometric)
    (supernova_models.py:2120-2232) against outputs of the reference's own sources; flux density and the fused
    likelihood of csm_nickel against the oracle restatement (pinned to the same fixture on the CPU)."""
    with open(os.path.join(ROOT, "tests", "golden", "csm_reference_draws.json")) as fh:
        fx = json.load(fh)
    t_all = np.array(fx["time"])
    rows = [row for row in fx["draws"] if "lbol_quadrature" in row]
    assert len(rows) >= 4
    for row in rows:
        t = t_all[row["skip"]:]
        p = dict(row["params"])
        p.pop("temperature_floor")
        out = rb.supernova_models.csm_interaction_bolometric(t, csm_diffusion_method="quadrature", **p)
        h.assert_close(out, np.array(row["lbol_quadrature"]), RTOL, what="CSMDiffusion quadrature vs reference")
        c = {k: v for k, v in p.items() if k != "vej"}
        out = rb.supernova_models.csm_nickel_bolometric(t, **c, **row["nickel_params"])
        h.assert_close(out, np.array(row["csm_nickel_lbol"]), RTOL, what="csm_nickel_bolometric vs reference")
        out = rb.supernova_models.csm_nickel_bolometric(t, csm_diffusion_method="legacy", csm_diffusion_timesteps=1000,
                                                        **c, **row["nickel_params"])
        h.assert_close(out, np.array(row["csm_nickel_lbol_quadrature"]), RTOL, what="csm_nickel legacy vs reference")
    # batched flux density + fused likelihood
    t = t_all[6:]
    nu = np.where(np.arange(t.size) % 2 == 0, 6.3e14, 3.9e14)
    names = ("mej", "csm_mass", "eta", "rho", "kappa", "r0")
    batch = {k: np.array([row["params"][k] for row in rows]) for k in names}
    for k in ("f_nickel", "ek", "kappa_gamma"):
        batch[k] = np.array([row["nickel_params"][k] for row in rows])
    batch["redshift"] = np.linspace(0.01, 0.3, len(rows))
    batch["temperature_floor"] = np.array([row["params"]["temperature_floor"] for row in rows])
    fn = rb.supernova_models.csm_nickel
    kw = dict(frequency=nu, output_format="flux_density")
    out = fn(t, params_batch=batch, **kw)
    y, sigma = out[1] * 1.03, 0.05 * np.abs(out[1]) + 1e-12
    logl = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=kw).log_likelihood_batch(batch)
    for i in range(len(rows)):
        kwi = {k: float(v[i]) for k, v in batch.items()}
        z = kwi.pop("redshift")
        ref = r.csm_nickel(t, z, frequency=nu, **kwi)
        h.assert_close(out[i], ref, RTOL, what=f"csm_nickel sample {i}")
        ref_l = r.gaussian_likelihood(y, ref, sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-9 * abs(ref_l), (i, logl[i], ref_l)


def test_prior_draws_flux_density_and_likelihood(rb):
    rng = np.random.default_rng(81)
    tb = np.sort(rng.uniform(0.5, 250, 40))
    t = np.concatenate([tb, tb, tb[:20] + 0.25])          # multi-band: repeated and interleaved epochs, unsorted
    nu = np.concatenate([np.full(40, 6.3e14), np.full(40, 4.7e14), np.full(20, 3.9e14)])
    N = 60
    p = h.csm_prior(rng, N)
    fn = rb.supernova_models.csm_interaction
    kw = dict(frequency=nu, output_format="flux_density")
    out = fn(t, params_batch=p, **kw)
    y = out[5] * (1 + 0.05 * rng.standard_normal(t.size))
    sigma = 0.05 * np.abs(out[5]) + 1e-12
    logl = rb.GaussianLikelihood(t, y, sigma, fn, kwargs=kw).log_likelihood_batch(p)
    for i in range(N):
        kwi = {k: v[i] for k, v in p.items()}
        z = kwi.pop("redshift")
        ref = r.csm_interaction(t, z, frequency=nu, **kwi)
        h.assert_close(out[i], ref, RTOL, what=f"csm_interaction sample {i}")
        ref_l = r.gaussian_likelihood(y, ref, sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-9 * abs(ref_l), (i, logl[i], ref_l)
    # scalar call == batch row; early epochs (< 0.1 d) move the dense grid start (supernova_models.py:2028)
    te = np.array([0.02, 0.05, 0.3, 2.0, 30.0])
    kwi = {k: float(v[3]) for k, v in p.items()}
    z = kwi.pop("redshift")
    one = fn(te, z, frequency=5e14, output_format="flux_density", **kwi)
    h.assert_close(one, r.csm_interaction(te, z, frequency=np.full(5, 5e14), **kwi), RTOL, what="early epochs")


def test_kwargs_and_errors(rb):
    t = np.geomspace(1, 300, 25)
    base = dict(mej=10.0, csm_mass=20.0, vej=6000.0, eta=0.7, rho=3e-13, kappa=0.34, r0=200.0)
    out = rb.supernova_models.csm_interaction_bolometric(t, nn=9.5, delta=0, efficiency=0.3, dense_resolution=600,
                                                         **base)
    ref = r.csm_interaction_bolometric(t, nn=9.5, delta=0, efficiency=0.3, dense_resolution=600, **base)
    h.assert_close(out, ref, RTOL, what="nn / delta / efficiency / dense_resolution")
    with pytest.raises(ValueError):
        rb.supernova_models.csm_interaction_bolometric(t, nn=15, **base)          # outside the (n, s) table
    with pytest.raises(ValueError):
        rb.supernova_models.csm_interaction_bolometric(np.array([0.0, 1.0]), **base)
    with pytest.raises(ValueError):
        rb.supernova_models.csm_interaction_bolometric(t, csm_diffusion_method="simpson", **base)
    bad = rb.supernova_models.csm_interaction_bolometric(t, **dict(base, eta=2.5))  # reference raises; NaN per sample
    assert np.all(np.isnan(bad))
