"""Throughput of the secondary fused paths (kilonova, afterglow, supernova variants) on one GPU.

    python tools/bench_paths.py [--quick]

Prints one JSON line per workload: sample x epoch likelihood evaluations per second with inputs resident in
HBM, timed with CUDA events (3 warm-up + 3 timed launches).  Not the headline bench (bench.py is)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers as h  # noqa: E402
from redback_b200 import _capi  # noqa: E402
from redback_b200 import bands as rb_bands  # noqa: E402
from redback_b200.engine import (AfterglowPath, KilonovaMagPath, KilonovaPath, MergernovaPath,  # noqa: E402
                                 MetzgerMagnetarPath, SupernovaMagPath, SupernovaPath)


def timed(path, params, steps=3, warmup=3):
    for _ in range(warmup):
        path.evaluate(params, want_model=False, want_logl=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        _, logl = path.evaluate(params, want_model=False, want_logl=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, logl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    scale = 0.1 if args.quick else 1.0
    rng = np.random.default_rng(0)
    jobs = []
    # config 1: arnett_bolometric, E = 100
    t1 = h.config1_time()
    jobs.append(("arnett_bolometric+GaussianLikelihood E=100 (config 1 shape)", int(2_000_000 * scale),
                 lambda n: SupernovaPath(_capi.ENGINE_NICKEL, _capi.SED_BOLOMETRIC, t1, y=np.full(100, 1e42), sigma=1e41),
                 lambda n: h.arnett_bolometric_prior(rng, n), 4.2e3))
    t2, nu2 = h.config2_epochs(np.random.default_rng(0))
    jobs.append(("slsn CutoffBlackbody multiband E=300 (config 2, default sed)", int(2_000_000 * scale),
                 lambda n: SupernovaPath(_capi.ENGINE_MAGNETAR, _capi.SED_CUTOFF_BLACKBODY, t2, frequency=nu2,
                                         y=np.full(300, 1e-3), sigma=1e-4),
                 lambda n: h.slsn_prior(rng, n), 4.1e3))
    jobs.append(("csm_interaction Blackbody multiband E=300 (config 2 epochs)", int(2_000_000 * scale),
                 lambda n: SupernovaPath(_capi.ENGINE_CSM, _capi.SED_BLACKBODY, t2, frequency=nu2,
                                         y=np.full(300, 1e-3), sigma=1e-4,
                                         csm=dict(nn=12, delta=1, efficiency=0.5)),
                 lambda n: h.csm_prior(rng, n, zmax=3.0), None))
    t3, nu3 = h.config3_epochs()
    jobs.append(("one_component_kilonova_model flux_density E=200 (config 3 epochs)", int(1_000_000 * scale),
                 lambda n: KilonovaPath(_capi.KN_ONE_COMPONENT, t3, frequency=nu3, y=np.full(200, 1e-3), sigma=1e-4),
                 lambda n: h.one_component_prior(rng, n), None))
    jobs.append(("one_component_kilonova_model flux_density E=200, FP32 smooth factors", int(1_000_000 * scale),
                 lambda n: KilonovaPath(_capi.KN_ONE_COMPONENT, t3, frequency=nu3, y=np.full(200, 1e-3), sigma=1e-4,
                                        dtype="float32"),
                 lambda n: h.one_component_prior(rng, n), None))
    jobs.append(("metzger_kilonova_model flux_density E=200", int(200_000 * scale),
                 lambda n: KilonovaPath(_capi.KN_METZGER, t3, frequency=nu3, y=np.full(200, 1e-3), sigma=1e-4),
                 lambda n: h.metzger_prior(rng, n), None))
    jobs.append(("general_mergernova flux_density E=200 (config 3 epochs)", int(1_000_000 * scale),
                 lambda n: MergernovaPath(_capi.MN_MAGNETAR_ONLY, t3, frequency=nu3, y=np.full(200, 1e-3), sigma=1e-4,
                                          pair_cascade_switch=True),
                 lambda n: h.mergernova_prior(rng, n, "general"), None))
    jobs.append(("general_metzger_magnetar_driven flux_density E=200 (config 3 epochs)", int(200_000 * scale),
                 lambda n: MetzgerMagnetarPath(_capi.MN_MAGNETAR_ONLY, t3, frequency=nu3, y=np.full(200, 1e-3),
                                               sigma=1e-4),
                 lambda n: h.metzger_magnetar_prior(rng, n, "general"), None))
    t4, nu4 = h.config4_epochs()
    jobs.append(("tophat_redback radio+X-ray E=500 (config 4 shape)", int(200_000 * scale),
                 lambda n: AfterglowPath(_capi.AG_TOPHAT, t4, frequency=nu4, y=np.full(500, 1e-3), sigma=1e-4),
                 lambda n: h.tophat_prior(rng, n), None))
    jobs.append(("tophat_redback radio+X-ray E=500 (config 4 shape), FP32 spectrum transcendentals", int(200_000 * scale),
                 lambda n: AfterglowPath(_capi.AG_TOPHAT, t4, frequency=nu4, y=np.full(500, 1e-3), sigma=1e-4,
                                         dtype="float32"),
                 lambda n: h.tophat_prior(rng, n), None))
    rb_bands.register_synthetic_lsst()
    names6 = list(h.LSST_NU)
    b3 = np.array([names6[i % 6] for i in range(200)])
    jobs.append(("one_component_kilonova_model LSST ugrizy magnitudes E=200 (config 3)", int(100_000 * scale),
                 lambda n: KilonovaMagPath(_capi.KN_ONE_COMPONENT, t3, b3, y=np.full(200, 22.0), sigma=0.1),
                 lambda n: h.one_component_prior(rng, n), 4.0e4))
    t5 = np.arange(1.0, 151.0, 0.5)     # LSST-like cadence over 150 d, E = 300 (config 5 shape)
    b5 = np.array([names6[i % 6] for i in range(300)])
    jobs.append(("arnett LSST ugrizy magnitudes E=300 (config 5 shape, no noise model)", int(200_000 * scale),
                 lambda n: SupernovaMagPath(_capi.ENGINE_NICKEL, _capi.SED_BLACKBODY, t5, b5, y=np.full(300, 22.0),
                                            sigma=0.1),
                 lambda n: h.arnett_prior(rng, n, zmax=0.5), 2.0e4))
    for name, n, mk, prior, flop_eq in jobs:
        if args.only and args.only not in name:
            continue
        path = mk(n)
        p = prior(n)
        dev = path.pack(p, n)
        ms, logl = timed(path, dev)
        rate = n * path.E / (ms * 1e-3)
        line = {"workload": name, "samples": n, "epochs": path.E, "ms": ms, "units_per_s": rate,
                "finite_logl_frac": float(torch.isfinite(logl).double().mean().item())}
        if flop_eq:
            line["tflop_eq_per_s"] = rate * flop_eq / 1e12
        if "tophat" in name:
            res = np.array([max(10, min(int((2 - 10 * max(a - b, 0)) * b * c), 100))
                            for a, b, c in zip(p["thv"], p["thc"], p["g0"])])
            vals, counts = np.unique(res, return_counts=True)
            line["res_histogram"] = {int(v): int(c) for v, c in zip(vals, counts) if c * 200 > n or v in (10, 100)}
            line["mean_res"] = float(res.mean())
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
