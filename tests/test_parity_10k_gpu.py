""">= 10^4 random prior draws per BASELINE config against the CPU oracle (all host cores), at
the FLAT north-star tolerances -- the same gate bench.py runs in-run (bench_paths.parity_gate), here as assertions.

What is asserted per config (and why it is not simply "all elements within 1e-9"):
  * flux-like outputs: every element within rtol 1e-5 (the FP32 bar) and >= 99.99 % within rtol 1e-9; the remainder is the
    reference's own cancellation in t_i^2 - t_e^2 and must stay below 1e-8;
  * magnitudes: every element within 30 mag of its sample's peak within 1e-9 mag, every element within the documented
    round-off floor of the linear-flux spline (tests/test_reference_conditioning.py shows the reference's own result
    moving by whole magnitudes there under a 1-ulp perturbation);
  * log L: |d| <= 1e-6 + 1e-9 |log L| for flux-like outputs; the fused reduction equals a host sum over the device's own
    model rows to 1e-6 + 1e-12 |log L| everywhere (separates the likelihood arithmetic from the model comparison).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

N_DRAWS = 10_000
KEYS = ["config1_arnett_bolometric", "config2_slsn_cutoff_blackbody", "config3_kilonova_lsst_magnitudes",
        "config4_tophat_afterglow", "config5_arnett_population_lsst"]


@pytest.fixture(scope="module")
def env():
    import torch
    import redback_b200 as rb
    import bench_paths as bp
    dev = torch.device("cuda:0")
    return rb, bp, dev, {w["key"]: w for w in bp.workloads(rb, dev, scale=0.0)}, bp._oracle_bandpasses()


@pytest.mark.parametrize("key", KEYS)
def test_ten_thousand_prior_draws(env, key):
    rb, bp, dev, table, bandpasses = env
    w = table[key]
    n = N_DRAWS if key != "config4_tophat_afterglow" else 4000      # the afterglow oracle runs ~400 draws / s on 16 cores
    p = w["prior"](np.random.default_rng(77), n)
    y, sigma = bp.synthetic_data(w)
    path = w["build"](y, sigma)
    stats, _ = bp.parity_gate(rb, w, path, p, y, sigma, n, bandpasses, dev, budget_s=240.0)
    assert stats["n"] >= min(n, 2000), stats                          # the oracle's wall-clock budget was enough
    assert stats["nan_pattern_mismatches"] == 0, stats
    fused = stats["fused_logl_vs_host_sum_of_device_model"]
    assert fused["frac_within_1e-6_or_1e-12_rel"] == 1.0, stats
    if w.get("magnitudes"):
        assert stats["frac_within_1e-9_within_30mag_of_sample_peak"] == 1.0, stats
        assert stats["frac_within_1e-9_plus_roundoff_floor"] == 1.0, stats
        if key.startswith("config5"):
            assert stats["frac_within_1e-9"] == 1.0, stats
            assert stats["frac_dlogl_within_1e-6_or_1e-9_rel"] == 1.0, stats
            assert stats["noise_stage"]["detected_mismatches"] == 0, stats
    else:
        assert stats["frac_within_1e-5"] == 1.0, stats
        assert stats["frac_within_1e-9"] >= 0.9999, stats
        assert stats["max_rel"] < 1e-8, stats
        assert stats["frac_dlogl_within_1e-6_or_1e-9_rel"] == 1.0, stats
