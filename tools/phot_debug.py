"""Developer aid: evaluate the config-3 kilonova magnitudes (and the config-5 arnett magnitudes) for a few thousand
prior draws on the GPU and save inputs + outputs for an offline comparison with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h  # noqa: E402
import redback_b200 as rb  # noqa: E402
import bench_paths as bp  # noqa: E402


def mode_counts(reset=True):
    """rb_phot_mode_counts: samples per clean-up mode of the photometry kernel since the last reset."""
    import ctypes
    buf = (ctypes.c_uint64 * 8)()
    rb._capi.lib().rb_phot_mode_counts(buf, 1 if reset else 0)
    names = ["all_valid", "single_pass", "single_pass+mask", "two_pass", "refused:negative", "refused:exponent",
             "refused:zero", "refused:magnitude"]
    return dict(zip(names, [int(x) for x in buf]))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    rb.bands.register_synthetic_lsst()
    out = {}
    mode_counts()
    t3, _ = h.config3_epochs()
    b3 = np.array([bp.LSST[i % 6] for i in range(t3.size)])
    p = h.one_component_prior(np.random.default_rng(77), n)
    out["kn_mag"] = rb.kilonova_models.one_component_kilonova_model(t3, params_batch=p, bands=b3, output_format="magnitude")
    print("kilonova modes", mode_counts())
    for k, v in p.items():
        out["kn_" + k] = v
    t5, b5, _ = bp.config5_cadence()
    q = h.arnett_prior(np.random.default_rng(77), n, zmax=0.5)
    out["sn_mag"] = rb.supernova_models.arnett(t5, params_batch=q, bands=b5, output_format="magnitude")
    print("arnett modes", mode_counts())
    for k, v in q.items():
        out["sn_" + k] = v
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez(os.path.join(ROOT, "gpurun_out", "phot_debug.npz"), **out)
    print("saved", n)


if __name__ == "__main__":
    main()
