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


def test_device_philox_variates_match_restatement(rb):
    """The kernel's own N(0,1) stream = the oracle's Philox4x32-10 restatement (pinned to Random123's known answers);
    the noise model replayed on the host with those variates reproduces every output."""
    rng = np.random.default_rng(6)
    N, E, base, seed = 37, 60, 12345, 2**40 + 17
    names = np.array(list(LIMITING)[:6] * 10)
    lim = np.array([LIMITING[n] for n in names])
    ref = np.array([rb.bands.REFERENCE_FLUX[n] for n in names])
    mags = rng.uniform(18, 27, (N, E))
    out = rb.simulate.add_optical_noise_and_cuts(torch.as_tensor(mags, device="cuda"), names, lim, seed=seed,
                                                 sample_base=base, return_variates=True)
    z_dev = out["standard_normal"].cpu().numpy()
    z_ref = r.philox_standard_normal(seed, (base * E + np.arange(N * E)).reshape(N, E))
    np.testing.assert_allclose(z_dev, z_ref, rtol=0, atol=4e-15)      # log / cos differ by an ulp between libm and CUDA
    exp = r.optical_noise_and_cuts(mags, ref, lim, z_dev)
    np.testing.assert_allclose(out["magnitude"].cpu().numpy(), exp["magnitude"], rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(out["flux"].cpu().numpy(), exp["flux"], rtol=1e-12)
    assert np.array_equal(out["detected"].cpu().numpy(), exp["detected"])


def test_population_streaming_is_chunk_invariant(rb):
    """Chunked + streamed-to-host evaluation (the 10^7-transient mode) returns exactly what the one-shot call returns,
    for any chunk size: the variates are keyed by the global element index."""
    t = np.arange(2.0, 62.0, 3.0)
    names = np.array([list(LIMITING)[i % 6] for i in range(t.size)])
    lim = np.array([LIMITING[n] for n in names])
    rng = np.random.default_rng(10)
    N = 333
    p = h.arnett_prior(rng, N, zmax=0.3)
    p["temperature_floor"] = h.loguniform(rng, 3e3, 1e4, N)
    whole = rb.simulate.simulate_population_with_cadence(rb.supernova_models.arnett, p, t, names, lim, seed=8)
    got = {k: np.full((N, t.size), np.nan) for k in ("magnitude", "magnitude_error", "model_magnitude")}
    got["detected"] = np.zeros((N, t.size), dtype=np.uint8)
    seen = []

    def sink(lo, hi, arrays):
        seen.append((lo, hi))
        for k in got:
            got[k][lo:hi] = arrays[k]

    assert rb.simulate.simulate_population_with_cadence(rb.supernova_models.arnett, p, t, names, lim, seed=8, chunk=50,
                                                        sink=sink, outputs=("magnitude", "magnitude_error")) is None
    assert sorted(seen) == [(a, min(a + 50, N)) for a in range(0, N, 50)]
    for k in ("magnitude", "magnitude_error", "model_magnitude"):
        assert np.array_equal(got[k], whole[k].cpu().numpy(), equal_nan=True), k
    assert np.array_equal(got["detected"].astype(bool), whole["detected"].cpu().numpy())
    # a rank that owns samples [100, 200) of the same population reproduces that slice
    sub = {k: v[100:200] for k, v in p.items()}
    part = rb.simulate.simulate_population_with_cadence(rb.supernova_models.arnett, sub, t, names, lim, seed=8,
                                                        sample_base=100)
    assert np.array_equal(part["magnitude"].cpu().numpy(), whole["magnitude"][100:200].cpu().numpy(), equal_nan=True)


def test_survey_observations_with_per_transient_pointings(rb):
    """SimulateOpticalTransient._make_observation_single (simulate_transients.py:1096-1140) for a small population with
    ragged pointing lists: empty lists, pointings before the explosion (phase <= 0: clamped by the spline, flagged
    undetected), unequal counts, optional source noise.  Model magnitudes against the oracle; the noise stage replayed
    by the oracle with the device's own variates; results independent of the chunk size."""
    from redback_b200.bands import REFERENCE_FLUX
    ob = {n: r.Bandpass(n, b.wave, b.trans) for n, b in
          ((n, rb.bands.get_bandpass(n)) for n in rb.bands.SYNTHETIC_LSST_EDGES)}
    rng = np.random.default_rng(17)
    N = 41
    counts = rng.integers(0, 60, N)
    counts[3] = 0
    counts[7] = 1
    off = np.concatenate([[0], np.cumsum(counts)])
    t0 = rng.uniform(60000, 60100, N)
    p = h.arnett_prior(rng, N, zmax=0.3)
    p["temperature_floor"] = h.loguniform(rng, 3e3, 1e4, N)
    names_all = list(LIMITING)
    mjd = np.concatenate([np.sort(t0[i] + rng.uniform(-5.0, 150.0, counts[i])) for i in range(N)])
    filt = np.array([names_all[k] for k in rng.integers(0, 6, off[-1])])
    depth = np.array([LIMITING[b] for b in filt]) + rng.normal(0, 0.3, off[-1])
    for source_noise in (False, True):
        out = rb.simulate.simulate_survey_observations(rb.supernova_models.arnett, p, t0, off, mjd, filt, depth, seed=21,
                                                       add_source_noise=source_noise, return_variates=True)
        small = rb.simulate.simulate_survey_observations(rb.supernova_models.arnett, p, t0, off, mjd, filt, depth,
                                                         seed=21, add_source_noise=source_noise, chunk=7)
        for k in ("model_magnitude", "magnitude", "e_magnitude", "flux", "flux_error"):
            assert np.array_equal(out[k], small[k], equal_nan=True), k
        assert np.array_equal(out["detected"], small["detected"])
        zref = r.philox_standard_normal(21, np.arange(off[-1]))
        np.testing.assert_allclose(out["standard_normal"], zref, rtol=0, atol=4e-15)
        for i in range(N):
            sl = slice(off[i], off[i + 1])
            if counts[i] == 0:
                continue
            kw = {k: v[i] for k, v in p.items()}
            z = kw.pop("redshift")
            phase = mjd[sl] - t0[i]
            ref_mag = r.sn_magnitude("arnett", phase, z, bands=filt[sl], bandpasses=ob, **kw)
            assert np.max(np.abs(out["model_magnitude"][sl] - ref_mag)) < 5e-9, i
            rf = np.array([REFERENCE_FLUX[b] for b in filt[sl]])
            exp = r.survey_observation_single(ref_mag, phase, rf, depth[sl], out["standard_normal"][sl],
                                              add_source_noise=source_noise)
            np.testing.assert_allclose(out["flux"][sl], exp["flux"], rtol=1e-8)
            np.testing.assert_allclose(out["flux_error"][sl], exp["flux_error"], rtol=1e-8)
            np.testing.assert_allclose(out["magnitude"][sl], exp["magnitude"], rtol=0, atol=1e-6, equal_nan=True)
            np.testing.assert_allclose(out["e_magnitude"][sl], exp["e_magnitude"], rtol=1e-8, equal_nan=True)
            assert np.array_equal(out["detected"][sl], exp["detected"]), i
            assert not out["detected"][sl][phase <= 0].any()
        assert np.array_equal(out["time_days"], mjd - np.repeat(t0, counts))
