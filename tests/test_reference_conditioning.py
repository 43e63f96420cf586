"""How well-conditioned is the REFERENCE's magnitude branch?  (CPU only; oracle = scipy FITPACK like sncosmo.)

`get_correct_output_format_from_spectra` (sed.py:971-1009) fits the bicubic spline in LINEAR flux.  A band-epoch whose
flux is 10^(-0.4 dm) of the sample's brightest one is a difference of O(1) spline terms at the 1e-16 level of the
brightest value, so its magnitude carries a round-off of ~ eps * 10^(0.4 dm) in the reference's own result.  For the
one_component_kilonova_model prior (cool sources, blue bands: magnitudes of 100+) this is visible: a 1-ulp perturbation of
the input spectra moves the oracle's own magnitudes by > 1e-9 -- up to whole magnitudes -- but only far down the Wien
tail.  The parity gates (bench_paths.parity_stats, tests/test_mag_parity_gpu.py) therefore hold every element within
30 mag of its sample's peak to the flat 1e-9 and allow the documented floor below that.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import helpers as h  # noqa: E402
import bench_paths as bp  # noqa: E402
from oracle import restated as r  # noqa: E402


def test_magnitudes_of_the_reference_algorithm_are_roundoff_limited_only_far_below_peak(monkeypatch):
    bps = bp._oracle_bandpasses()
    t3, _ = h.config3_epochs()
    b3 = np.array([bp.LSST[i % 6] for i in range(t3.size)])
    p = h.one_component_prior(np.random.default_rng(5), 12)
    prng = np.random.default_rng(11)
    clean = r.clean_spectra

    def one_ulp(values):
        v = clean(values)
        return v * (1 + prng.integers(-1, 2, size=v.shape) * 1.1e-16)

    moved, brightest_moved = 0, np.inf
    for i in range(12):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        with np.errstate(all="ignore"):
            monkeypatch.setattr(r, "clean_spectra", clean)
            m0 = r.kn_magnitude("one_component", t3, z, bands=b3, bandpasses=bps, **kw)
            monkeypatch.setattr(r, "clean_spectra", one_ulp)
            m1 = r.kn_magnitude("one_component", t3, z, bands=b3, bandpasses=bps, **kw)
        ok = np.isfinite(m0) & np.isfinite(m1)
        d = np.abs(m1 - m0)[ok]
        dm = (m0 - np.nanmin(m0))[ok]
        # within 30 mag of the peak the reference's result is stable to far better than the 1e-9 gate ...
        assert d[dm <= 30.0].max() < 1e-10
        # ... and the documented floor covers every element
        assert np.all(d <= 1e-9 + 2.3e-13 * 10 ** (0.4 * dm))
        bad = d > 1e-9
        moved += int(bad.sum())
        if bad.any():
            brightest_moved = min(brightest_moved, dm[bad].min())
    # the instability is real for this prior (not an artefact of the device path) and sits > 60 mag below the peak
    assert moved > 0
    assert brightest_moved > 60.0
