"""Secondary workloads of bench.py: every BASELINE.json config next to the headline (config 2 / Blackbody / FP64).

Each workload is one fused model + Gaussian-likelihood path evaluated over a batch of prior draws of the reference's
own prior file, timed on the device with CUDA events (L2 flushed between timed launches), with
  * `roofline`  -- SURVEY.md §8d's algorithmic flop-eq per sample x epoch against the FP64 DFMA peak measured in-run,
  * `clocks`    -- nvidia-smi samples taken during this workload's own timed region,
  * `parity`    -- the same prior draws (>= 10^4 by default) evaluated by the CPU oracle on the host cores in this
                   run, compared at the FLAT north-star tolerances (rtol 1e-9 on the model, |d log L| <= 1e-6),
  * `cpu_port`  -- the oracle's own rate measured while it produced the parity reference.
`oracle/` is used here only as the checker / CPU baseline (never on the product path).
"""
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
import sys  # noqa: E402
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import helpers as h  # noqa: E402  (tests/helpers.py: the seeded synthetic inputs of SURVEY.md §8d)

ZTF_NU = (6.272e14, 4.675e14, 3.813e14)  # redback/tables/filters.csv:65-67 (ztfg, ztfr, ztfi)
LSST = ["lsstu", "lsstg", "lsstr", "lssti", "lsstz", "lssty"]
LSST_5SIGMA = dict(lsstu=23.9, lsstg=25.0, lsstr=24.7, lssti=24.0, lsstz=23.3, lssty=22.1)


def config2_epochs():
    return h.config2_epochs(np.random.default_rng(0))


def config5_cadence():
    """bands ugrizy in rotation, cadence 3 d per band, duration 150 d -> E = 300 (simulate_transients.py:168-176)."""
    t = np.arange(1.0, 151.0, 0.5)
    bands = np.array([LSST[i % 6] for i in range(t.size)])
    lim = np.array([LSST_5SIGMA[b] for b in bands])
    return t, bands, lim


# ------------------------------------------------------------------------------------------------------------
# oracle side (CPU): module-level state so that forked workers see the job without pickling it
# ------------------------------------------------------------------------------------------------------------
_JOB = {}


def _oracle_bandpasses():
    from oracle import restated as r
    from redback_b200 import bands as rb_bands
    rb_bands.register_synthetic_lsst()
    return {n: r.Bandpass(n, b.wave, b.trans) for n, b in
            ((n, rb_bands.get_bandpass(n)) for n in rb_bands.SYNTHETIC_LSST_EDGES)}


def oracle_model(kind, kw, ctx):
    """One sample of workload `kind` through the CPU oracle: the model row [E]."""
    from oracle import restated as r
    kw = dict(kw)
    if kind == "arnett_bolometric":
        return r.arnett_bolometric(ctx["t"], **kw)
    z = kw.pop("redshift")
    if kind in ("slsn_blackbody", "slsn_f32"):
        return r.slsn(ctx["t"], z, frequency=ctx["nu"], sed="blackbody", **kw)
    if kind == "slsn_cutoff":
        return r.slsn(ctx["t"], z, frequency=ctx["nu"], sed="cutoff_blackbody", **kw)
    if kind == "kn_mag":
        return r.kn_magnitude("one_component", ctx["t"], z, bands=ctx["bands"], bandpasses=ctx["bandpasses"], **kw)
    if kind == "tophat":
        return r.redback_afterglow(ctx["t"], z, frequency=ctx["nu"], **kw)
    if kind in ("arnett_mag", "arnett_population"):
        return r.sn_magnitude("arnett", ctx["t"], z, bands=ctx["bands"], bandpasses=ctx["bandpasses"], **kw)
    raise KeyError(kind)


def _oracle_worker(span):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import restated as r
    lo, hi = span
    job = _JOB
    ctx, params, kind = job["ctx"], job["params"], job["kind"]
    E = ctx["t"].size
    model = np.empty((hi - lo, E))
    logl = np.empty(hi - lo)
    t0 = time.perf_counter()
    with np.errstate(all="ignore"):
        for i in range(lo, hi):
            kw = {k: (v[i] if np.ndim(v) else v) for k, v in params.items()}
            m = oracle_model(kind, kw, ctx)
            model[i - lo] = m
            logl[i - lo] = r.gaussian_likelihood(ctx["y"], m, ctx["sigma"]) if ctx.get("y") is not None else 0.0
    return lo, model, logl, time.perf_counter() - t0


def run_oracle(kind, ctx, params, n, cores=None, budget_s=None):
    """Oracle model rows and log L for the first `n` samples on `cores` processes.  Returns
    (model[n_done, E], logl[n_done], n_done, core_seconds, wall_seconds).  With `budget_s` the tail of the sample is
    dropped once the wall-clock budget is spent (n_done < n is reported, never hidden)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    _JOB.clear()
    _JOB.update(kind=kind, ctx=ctx, params=params)
    step = max(1, min(64, n // (cores * 4) or 1))
    spans = [(a, min(a + step, n)) for a in range(0, n, step)]
    E = ctx["t"].size
    model = np.full((n, E), np.nan)
    logl = np.full(n, np.nan)
    done = np.zeros(n, dtype=bool)
    core_s = 0.0
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        for lo, m, l, dt in pool.imap_unordered(_oracle_worker, spans):
            model[lo:lo + len(l)] = m
            logl[lo:lo + len(l)] = l
            done[lo:lo + len(l)] = True
            core_s += dt
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                pool.terminate()
                break
    wall = time.perf_counter() - t0
    n_done = int(np.argmin(done)) if not done.all() else n      # keep the contiguous prefix
    return model[:n_done], logl[:n_done], n_done, core_s * (n_done / max(1, int(done.sum()))), wall


def parity_stats(gpu_model, gpu_logl, ref_model, ref_logl, magnitudes=False):
    """Agreement at the FLAT north-star tolerances (no condition-number widening)."""
    a, b = np.asarray(gpu_model, dtype=float), np.asarray(ref_model, dtype=float)
    nan_mismatch = int(np.sum(np.isnan(a) != np.isnan(b)))
    ok = np.isfinite(a) & np.isfinite(b)
    with np.errstate(all="ignore"):
        if magnitudes:
            err = np.abs(a - b)
            within = err <= 1e-9 + 1e-9 * np.abs(b)
            within5 = err <= 1e-5
        else:
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
            within = (err <= 1e-9) | (a == b)
            within5 = (err <= 1e-5) | (a == b)
    same_inf = (~ok) & (a == b)
    counted = ok | same_inf
    err_ok = err[ok]
    extra = {}
    if magnitudes:
        # The spline is fitted in linear flux: a band-epoch whose flux is 10^(-0.4 dm) of the sample's brightest one
        # carries a relative round-off ~ eps * 1e3 * 10^(0.4 dm) in the REFERENCE's own result (Wien tail of cool
        # sources: magnitudes of 100+), so a second fraction is reported with that floor added to the tolerance.
        with np.errstate(all="ignore"):
            floor = np.nanmin(np.where(ok, b, np.inf), axis=-1, keepdims=True)
            cond = err <= 1e-9 + 1e-9 * np.abs(b) + 2.3e-13 * 10 ** (0.4 * (b - floor))
            faint = (b - floor) > 9.0
            # the same statement without a model of the floor: every element within 30 mag (a flux ratio of 1e-12) of
            # its sample's brightest epoch at the flat tolerance, and how far down the first disagreement sits.  The
            # oracle's own FITPACK result moves by up to several mag under a 1-ulp perturbation of the spectra, but only
            # from ~80 mag below the peak (tests/test_reference_conditioning.py).
            near = ok & ((b - floor) <= 30.0)
            fails = ok & ~within
        extra = {"frac_within_1e-9_plus_roundoff_floor": float((cond & ok | same_inf).sum() / max(1, counted.sum())),
                 "frac_elements_more_than_9mag_below_sample_peak": float((faint & ok).sum() / max(1, ok.sum())),
                 "elements_within_30mag_of_sample_peak": int(near.sum()),
                 "frac_within_1e-9_within_30mag_of_sample_peak": float((within & near).sum() / max(1, near.sum())),
                 "max_abs_dmag_within_30mag_of_sample_peak": float(err[near].max()) if near.any() else 0.0,
                 "brightest_disagreement_mag_below_sample_peak":
                     float((b - floor)[fails].min()) if fails.any() else None}
    out = {"n": int(a.shape[0]), "elements": int(counted.sum()), **extra,
           ("max_abs_dmag" if magnitudes else "max_rel"): float(err_ok.max()) if err_ok.size else 0.0,
           "frac_within_1e-9": float((within & ok | same_inf).sum() / max(1, counted.sum())),
           "frac_within_1e-5": float((within5 & ok | same_inf).sum() / max(1, counted.sum())),
           "nan_pattern_mismatches": nan_mismatch}
    if gpu_logl is not None:
        la, lb = np.asarray(gpu_logl, dtype=float), np.nan_to_num(np.asarray(ref_logl, dtype=float))
        d = np.abs(la - lb)
        out.update({"max_abs_dlogl": float(d.max()), "frac_dlogl_within_1e-6": float(np.mean(d <= 1e-6)),
                    "frac_dlogl_within_1e-6_or_1e-9_rel": float(np.mean(d <= 1e-6 + 1e-9 * np.abs(lb))),
                    "median_abs_logl": float(np.median(np.abs(lb)))})
    return out


def synthetic_data(w):
    """Data of a workload: the DEVICE model at the workload's fixed truth + seeded noise (10 % of the model for flux-like
    outputs, 0.1 mag for magnitudes)."""
    probe = w["build"](None, None)
    y_model, _ = probe.evaluate(probe.pack(w["truth"], 1))
    y_model = y_model[0].cpu().numpy()
    nrng = np.random.default_rng(4)
    if w.get("magnitudes"):
        sigma = np.full(y_model.shape, 0.1)
        y = y_model + nrng.normal(0, 0.1, y_model.shape)
    else:
        sigma = 0.1 * np.abs(y_model) + 1e-300
        y = y_model + nrng.normal(0, 1, y_model.shape) * sigma
    return np.where(np.isfinite(y), y, 0.0), sigma


def parity_gate(rb, w, path, p, y, sigma, par_n, bandpasses, dev, budget_s=None):
    """The first `par_n` prior draws of `p` through the device path (model rows and fused log L) and through the CPU
    oracle on the host cores; returns (parity statistics at the flat north-star tolerances, CPU-port rate)."""
    import torch
    ctx = dict(w["ctx"], y=y, sigma=sigma)
    if "bands" in ctx:
        ctx["bandpasses"] = bandpasses
    E = path.E
    head = {k: v[:par_n] for k, v in p.items()}
    gm, gl = path.evaluate(path.pack(head, par_n), want_model=True, want_logl=True)
    gm, gl = gm.cpu().numpy(), gl.cpu().numpy()
    rm, rl, n_done, core_s, wall = run_oracle(w["kind"], ctx, head, par_n, budget_s=budget_s)
    stats = parity_stats(gm[:n_done], gl[:n_done], rm, rl, magnitudes=bool(w.get("magnitudes")))
    stats["requested"] = par_n
    # the fused reduction against a host sum over the device's OWN model rows: separates the likelihood
    # arithmetic from the (for magnitudes: ill-conditioned, see parity_stats) model comparison
    with np.errstate(all="ignore"):
        own = np.nan_to_num(np.sum(-0.5 * ((y - gm[:n_done]) / sigma) ** 2 - 0.5 * np.log(2 * np.pi * sigma ** 2), axis=1))
    dd = np.abs(gl[:n_done] - own)
    stats["fused_logl_vs_host_sum_of_device_model"] = {
        "max_abs": float(dd.max()), "frac_within_1e-6_or_1e-12_rel": float(np.mean(dd <= 1e-6 + 1e-12 * np.abs(own)))}
    if w.get("population"):
        # noise + cut replayed by the oracle on ITS magnitudes with the device's own variates
        from oracle import restated as r
        lim = ctx["limiting_mag"]
        mags = torch.as_tensor(gm[:n_done], device=dev)
        dev_out = rb.simulate.add_optical_noise_and_cuts(mags, ctx["bands"], lim, seed=99, return_variates=True)
        ref = np.array([rb.bands.REFERENCE_FLUX[b] for b in ctx["bands"]])
        exp = r.optical_noise_and_cuts(rm, ref, lim, dev_out["standard_normal"].cpu().numpy())
        om = dev_out["magnitude"].cpu().numpy()
        both = np.isfinite(om) & np.isfinite(exp["magnitude"])
        stats["noise_stage"] = {
            "max_abs_dmag_observed": float(np.max(np.abs(om - exp["magnitude"])[both])) if both.any() else 0.0,
            "nan_pattern_mismatches": int(np.sum(np.isnan(om) != np.isnan(exp["magnitude"]))),
            "detected_mismatches": int(np.sum(dev_out["detected"].cpu().numpy() != exp["detected"]))}
    cores = os.cpu_count() or 1
    cpu_port = {"units_per_s_all_cores": n_done * E / wall, "units_per_s_per_core": n_done * E / core_s if core_s else None,
                "cores": cores, "kind": "port",
                "sample": f"{n_done} of the same prior draws x {E} epochs, {wall:.1f} s wall"}
    return stats, cpu_port


# ------------------------------------------------------------------------------------------------------------
# workload table
# ------------------------------------------------------------------------------------------------------------
def workloads(rb, dev, scale=1.0):
    """[(key, description, n_samples, flop_eq | None, build(dev, y, sigma) -> path, prior(rng, n), ctx, kind)]"""
    from redback_b200 import _capi
    from redback_b200.engine import AfterglowPath, KilonovaMagPath, SupernovaMagPath, SupernovaPath
    rb.bands.register_synthetic_lsst()
    t1 = h.config1_time()
    t2, nu2 = config2_epochs()
    t3, _ = h.config3_epochs()
    b3 = np.array([LSST[i % 6] for i in range(t3.size)])
    t4, nu4 = h.config4_epochs()
    t5, b5, lim5 = config5_cadence()
    n = lambda x: max(2048, int(x * scale))
    W = []
    W.append(dict(key="config1_arnett_bolometric", config=1, kind="arnett_bolometric", n=n(4_000_000), flop_eq=4.2e3,
                  desc="arnett_bolometric + GaussianLikelihood, E=100 (examples/bolometric_luminosity_fitting.py epochs)",
                  ctx=dict(t=t1), prior=h.arnett_bolometric_prior, truth=dict(f_nickel=0.15, mej=10., vej=5000.,
                                                                               kappa=0.07, kappa_gamma=0.1),
                  build=lambda y, s, **k: SupernovaPath(_capi.ENGINE_NICKEL, _capi.SED_BOLOMETRIC, t1, y=y, sigma=s,
                                                        device=dev)))
    slsn_truth = dict(redshift=0.1, p0=2.0, bp=1.0, mass_ns=1.4, theta_pb=0.5, mej=5.0, vej=8000.0, kappa=0.1,
                      kappa_gamma=0.1, temperature_floor=5000.0)
    W.append(dict(key="config2_slsn_cutoff_blackbody", config=2, kind="slsn_cutoff", n=n(4_000_000), flop_eq=4.1e3,
                  desc="slsn multiband flux_density, ZTF g/r/i x 100, the reference's default CutoffBlackbody SED",
                  ctx=dict(t=t2, nu=nu2), prior=h.slsn_prior, truth=slsn_truth,
                  build=lambda y, s, **k: SupernovaPath(_capi.ENGINE_MAGNETAR, _capi.SED_CUTOFF_BLACKBODY, t2,
                                                        frequency=nu2, y=y, sigma=s, device=dev)))
    W.append(dict(key="config2_slsn_blackbody_fp32", config=2, kind="slsn_f32", n=n(4_000_000), flop_eq=3.8e3,
                  desc="slsn multiband flux_density, Blackbody, FP32 diffusion quadrature + FP64 epilogue",
                  ctx=dict(t=t2, nu=nu2), prior=h.slsn_prior, truth=slsn_truth, fp32=True,
                  build=lambda y, s, **k: SupernovaPath(_capi.ENGINE_MAGNETAR, _capi.SED_BLACKBODY, t2, frequency=nu2,
                                                        y=y, sigma=s, device=dev, dtype="float32")))
    W.append(dict(key="config3_kilonova_lsst_magnitudes", config=3, kind="kn_mag", n=n(400_000), flop_eq=4.0e4,
                  desc="one_component_kilonova_model, LSST ugrizy magnitudes (synthetic smoothed top-hat bandpasses), "
                       "E=200", magnitudes=True,
                  ctx=dict(t=t3, bands=b3), prior=h.one_component_prior,
                  truth=dict(redshift=0.05, mej=0.03, vej=0.2, kappa=5.0, temperature_floor=2500.0),
                  build=lambda y, s, **k: KilonovaMagPath(_capi.KN_ONE_COMPONENT, t3, b3, y=y, sigma=s, device=dev)))
    W.append(dict(key="config4_tophat_afterglow", config=4, kind="tophat", n=n(400_000), flop_eq=None,
                  desc="tophat_redback radio (5 GHz) + X-ray (2e17 Hz) flux_density, E=500, res by the reference rule",
                  ctx=dict(t=t4, nu=nu4), prior=h.tophat_prior,
                  truth=dict(redshift=0.5, thv=0.1, loge0=52.0, thc=0.08, logn0=0.0, p=2.3, logepse=-1.0,
                             logepsb=-2.0, g0=500.0, xiN=1.0),
                  build=lambda y, s, **k: AfterglowPath(_capi.AG_TOPHAT, t4, frequency=nu4, y=y, sigma=s, device=dev)))
    W.append(dict(key="config5_arnett_population_lsst", config=5, kind="arnett_population", n=n(400_000), flop_eq=2.0e4,
                  desc="simulate_transients population: arnett LSST ugrizy magnitudes on a 3-day cadence over 150 d "
                       "(E=300) + limiting-magnitude noise + SNR cut, outputs streamed to the host in chunks",
                  magnitudes=True, population=True,
                  ctx=dict(t=t5, bands=b5, limiting_mag=lim5), prior=lambda rng, m: h.arnett_prior(rng, m, zmax=0.5),
                  truth=dict(redshift=0.1, f_nickel=0.15, mej=10., vej=5000., kappa=0.07, kappa_gamma=0.1,
                             temperature_floor=5000.0),
                  build=lambda y, s, **k: SupernovaMagPath(_capi.ENGINE_NICKEL, _capi.SED_BLACKBODY, t5, b5, y=y,
                                                           sigma=s, device=dev)))
    return W


def afterglow_flop_eq(params, E):
    """per sample 5.4e4 C + 8.8e4 R flop-eq with the realised res (C = R * N_rot)."""
    from redback_b200.distributed import afterglow_cost
    cost, res = afterglow_cost(params["thv"], params["thc"], params["g0"])
    vals, counts = np.unique(res, return_counts=True)
    hist = {int(v): int(c) for v, c in zip(vals, counts) if c * 200 > res.size or v in (10, 100)}
    return float(cost.mean() / E), {"mean_res": float(res.mean()), "histogram": hist}
