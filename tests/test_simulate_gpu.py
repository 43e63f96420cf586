"""GPU: batched SimulateTransientWithCadence (optical) -- model magnitudes, noise model, SNR cut."""
import numpy as np
import pytest
import torch

import helpers as h
from oracle import restated as r

pytestmark = pytest.mark.gpu

LIMITING = dict(lsstu=23.9, lsstg=25.0, lsstr=24.7, lssti=24.0, lsstz=23.3, lssty=22.1)


@pytest.fixture(scope="module")
def rb():
    import redback_b200
    redback_b200.bands.register_synthetic_lsst()
    return redback_b200


@pytest.mark.parametrize("noise_type", ["limiting_mag", "gaussian", "gaussianmodel"])
def test_noise_and_cuts_match_oracle(rb, noise_type):
    rng = np.random.default_rng(5)
    N, E = 64, 60
    names = np.array(list(LIMITING)[:6] * 10)
    mags = rng.uniform(18, 27, (N, E))
    mags[0, 0] = np.nan
    z = rng.standard_normal((N, E))
    lim = np.array([LIMITING[n] for n in names])
    ref = np.array([rb.bands.REFERENCE_FLUX[n] for n in names])
    out = rb.simulate.add_optical_noise_and_cuts(torch.as_tensor(mags, device="cuda"), names, lim, noise_type,
                                                 standard_normal=torch.as_tensor(z))
    exp = r.optical_noise_and_cuts(mags, ref, lim, z, noise_type)
    for key in ("magnitude", "magnitude_error", "snr"):
        got = out[key].cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(exp[key])), key
        np.testing.assert_allclose(got, exp[key], rtol=1e-12, equal_nan=True, err_msg=key)
    assert np.array_equal(out["detected"].cpu().numpy(), exp["detected"])
    if noise_type == "limiting_mag":
        np.testing.assert_allclose(out["flux"].cpu().numpy(), exp["flux"], rtol=1e-12, equal_nan=True)


def test_population_with_cadence(rb):
    """Config-5 shape in miniature: Arnett population on an LSST-like cadence with limiting-magnitude noise."""
    t = np.arange(2.0, 92.0, 3.0)
    names = np.array([list(LIMITING)[i % 6] for i in range(t.size)])
    lim = np.array([LIMITING[n] for n in names])
    rng = np.random.default_rng(9)
    N = 200
    p = h.arnett_prior(rng, N, zmax=0.3)
    p["temperature_floor"] = h.loguniform(rng, 3e3, 1e4, N)
    out = rb.simulate.simulate_population_with_cadence(rb.supernova_models.arnett, p, t, names, lim, seed=3)
    mm = out["model_magnitude"].cpu().numpy()
    assert mm.shape == (N, t.size)
    ref = np.array([rb.bands.REFERENCE_FLUX[n] for n in names])
    # deterministic parts against the oracle, given the model magnitudes
    exp = r.optical_noise_and_cuts(mm, ref, lim, np.zeros_like(mm))
    np.testing.assert_allclose(out["snr"].cpu().numpy(), exp["snr"], rtol=1e-12)
    assert np.array_equal(out["detected"].cpu().numpy(), exp["detected"])
    # noise statistics: (flux - model_flux) / flux_error ~ N(0, 1)
    pull = ((out["flux"].cpu().numpy() - r.bandpass_magnitude_to_flux(mm, ref)) / out["flux_error"].cpu().numpy())
    pull = pull[np.isfinite(pull)]
    assert abs(pull.mean()) < 0.05 and abs(pull.std() - 1) < 0.05
    again = rb.simulate.simulate_population_with_cadence(rb.supernova_models.arnett, p, t, names, lim, seed=3)
    assert torch.equal(again["flux"].nan_to_num(0.0), out["flux"].nan_to_num(0.0))   # seeded => reproducible


def test_population_light_curves(rb):
    """PopulationSynthesizer.simulate_population's per-event light-curve loop (simulate_transients.py:2027-2036) as
    one batched call; non-model columns of the population table are ignored."""
    import numpy as np
    import helpers as h
    from oracle import restated as r
    rng = np.random.default_rng(5)
    N = 9
    p = h.arnett_prior(rng, N, zmax=0.3)
    table = dict(p, ra=rng.uniform(0, 360, N), dec=rng.uniform(-90, 90, N), t0_mjd_transient=rng.uniform(6e4, 6.1e4, N))
    times = np.linspace(1, 100, 40)
    out = rb.simulate.population_light_curves(rb.supernova_models.arnett, table, times,
                                              model_kwargs=dict(frequency=4.7e14, output_format="flux_density"))
    assert out["flux"].shape == (N, 40) and np.array_equal(out["times"], times)
    for i in range(N):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        h.assert_close(out["flux"][i], r.arnett(times, z, frequency=np.full(40, 4.7e14), **kw), 1e-9, what=f"event {i}")
