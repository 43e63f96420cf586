#!/usr/bin/env python
"""Headline benchmark: sample x epoch model+likelihood evaluations per second (BASELINE.json).

Workload (BASELINE.json configs[1]): magnetar-powered SLSN (`slsn`) multiband
flux_density with TemperatureFloor + blackbody SED, ZTF g/r/i x 100 epochs (E = 300), 10^7 prior draws
per GPU, fused with the Gaussian log-likelihood.  One step = one pass over the whole batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--samples S] [--sed blackbody|cutoff]
    python bench.py --impl reference ...      # the reference algorithm (CPU oracle port) on host cores

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events with inputs resident in HBM; `e2e` is the
same metric through the host-buffer entry (pinned host params -> H2D -> kernel -> D2H of log L).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sample_x_epoch_likelihood_evals_per_sec"
UNIT = "sample*epoch/s"
ZTF_NU = (6.272e14, 4.675e14, 3.813e14)  # redback/tables/filters.csv:65-67 (ztfg, ztfr, ztfi)
# algorithmic work per sample x epoch in flop-eq = F + 20 T
FLOP_EQ = {"blackbody": 3.8e3, "cutoff": 4.1e3}
TRUTH = dict(redshift=0.1, p0=2.0, bp=1.0, mass_ns=1.4, theta_pb=0.5, mej=5.0, vej=8000.0, kappa=0.1,
             kappa_gamma=0.1, temperature_floor=5000.0)


def epochs(seed=0):
    rng = np.random.default_rng(seed)
    t = np.concatenate([rng.uniform(1, 300, 100) for _ in ZTF_NU])
    nu = np.concatenate([np.full(100, v) for v in ZTF_NU])
    order = np.argsort(t, kind="stable")
    return t[order], nu[order]


def prior_rows_numpy(rng, n):
    """priors/slsn.prior as SoA rows in C-ABI order (include/redback_b200.h RB_SN_*)."""
    lu = lambda lo, hi: np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    return np.stack([rng.uniform(1e-6, 3, n), rng.uniform(1, 10, n), lu(0.1, 10), rng.uniform(1.1, 2.2, n),
                     rng.uniform(0, 1.57, n), lu(0.1, 100), lu(1e3, 1e5), rng.uniform(0.05, 2, n), lu(1e-4, 1e4),
                     lu(1e3, 1e5)])


def shard_bounds(n_total, rank, world):
    """Contiguous shard [lo, hi) of `n_total` samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def prior_rows_torch(n, device, seed):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = lambda lo, hi: lo + (hi - lo) * torch.rand(n, dtype=torch.float64, device=device, generator=g)
    lu = lambda lo, hi: torch.exp(u(np.log(lo), np.log(hi)))
    return torch.stack([u(1e-6, 3), u(1, 10), lu(0.1, 10), u(1.1, 2.2), u(0, 1.57), lu(0.1, 100), lu(1e3, 1e5),
                        u(0.05, 2), lu(1e-4, 1e4), lu(1e3, 1e5)]).contiguous()


# ----------------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference algorithm on the host cores (only place oracle/ is executed)
# ----------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    rows, t, nu, y, sigma, sed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import restated as r
    names = ["p0", "bp", "mass_ns", "theta_pb", "mej", "vej", "kappa", "kappa_gamma", "temperature_floor"]
    t0 = time.perf_counter()
    acc = 0.0
    for i in range(rows.shape[1]):
        kw = dict(zip(names, rows[1:, i]))
        model = r.slsn(t, rows[0, i], frequency=nu, sed=sed, **kw)
        acc += r.gaussian_likelihood(y, model, sigma)
    return time.perf_counter() - t0, acc


def cpu_reference_rate(sed, per_core, seed, cores=None):
    """Units/s of the oracle port on `cores` processes (one per core), `per_core` samples each."""
    import multiprocessing as mp
    from oracle import restated as r
    cores = cores or os.cpu_count() or 1
    t, nu = epochs()
    osed = "blackbody" if sed == "blackbody" else "cutoff_blackbody"
    truth = dict(TRUTH)
    z = truth.pop("redshift")
    y = r.slsn(t, z, frequency=nu, sed=osed, **truth)
    sigma = 0.1 * np.abs(y) + 1e-9
    rng = np.random.default_rng(seed)
    jobs = [(prior_rows_numpy(rng, per_core), t, nu, y, sigma, osed) for _ in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(j[0][:, :2],) + j[1:] for j in jobs])  # warm import / caches
        t0 = time.perf_counter()
        pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    units = cores * per_core * t.size
    return units / wall, cores, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = max(8, args.cpu_samples_per_core)
    for _ in range(args.warmup):
        cpu_reference_rate(args.sed, 4, 1, cores)
    walls, rates = [], []
    for s in range(args.steps):
        rate, cores, wall = cpu_reference_rate(args.sed, per_core, 100 + s, cores)
        rates.append(rate)
        walls.append(wall)
    value = statistics.mean(rates)
    sample = f"{cores * per_core} prior draws x 300 epochs per step ({per_core} per core), same priors/epochs as the GPU arm"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(walls), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, cores * per_core),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, samples, total=None):
    return {"workload": f"slsn multiband flux_density (ZTF g/r/i x 100 epochs), TemperatureFloor + "
                        f"{'Blackbody' if args.sed == 'blackbody' else 'CutoffBlackbody'} SED, fused GaussianLikelihood",
            "samples_per_gpu": samples, "samples_total": total if total is not None else samples, "epochs": 300,
            "bands": ["ztfg", "ztfr", "ztfi"],
            "priors": "priors/slsn.prior incl. redshift U(1e-6,3); d_L integrated on device",
            "l2": "inputs_exceed_l2" if samples * 80 > 126e6 else "small_inputs", "sharding": "samples"}


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons, power = [], [], set(), []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def _events(n):
    import torch
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


class L2Flush:
    """Write a buffer larger than the 126 MB L2 between timed launches (outside the event brackets)."""

    def __init__(self, dev):
        import torch
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.v = 0

    def __call__(self):
        self.v = (self.v + 1) & 0xFF
        self.buf.fill_(self.v)


def _allmax(x, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def time_launches(fn, dev, world, flush, min_seconds=1.5, max_steps=50, warmup=3, sampler=None):
    """Mean milliseconds per call of `fn` (device-timed, one event pair per call, L2 flushed in between), max over
    ranks; the clock sampler runs during the timed region only."""
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    probe = _events(1)[0]
    probe[0].record(); fn(); probe[1].record()
    torch.cuda.synchronize(dev)
    est = _allmax(max(probe[0].elapsed_time(probe[1]), 1e-3), dev, world)
    steps = int(min(max_steps, max(3, np.ceil(min_seconds * 1e3 / est))))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler is not None:
        sampler.start()
    ev = _events(steps)
    for a, b in ev:
        flush()
        a.record(); fn(); b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if sampler is not None else None
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    return _allmax(ms, dev, world), steps, clocks


def run_paths(args, dev, rank, world, local, peak_tflops, lib):
    """Every other BASELINE config (bench_paths.workloads) with its own clock sample, roofline fraction and -- at N = 1
    -- an in-run parity gate against the CPU oracle on the host cores."""
    import torch
    import redback_b200 as rb
    import bench_paths as bp
    from redback_b200 import distributed as rd
    flush = L2Flush(dev)
    out = []
    do_parity = world == 1 and rank == 0 and args.parity_samples > 0
    bandpasses = bp._oracle_bandpasses() if do_parity else None
    for w in bp.workloads(rb, dev, args.paths_scale):
        if args.only and args.only not in w["key"]:
            continue
        n = w["n"]
        rng = np.random.default_rng(77)                      # the parity prefix is the same on every rank
        par_n = min(args.parity_samples, n) if do_parity else 0
        rng_r = np.random.default_rng(1000 + 17 * rank + w["config"])
        p = w["prior"](rng_r, n)
        if par_n:
            head = w["prior"](rng, par_n)
            p = {k: np.concatenate([head[k], v[par_n:]]) for k, v in p.items()}
        y, sigma = bp.synthetic_data(w)
        path = w["build"](y, sigma)
        E = path.E
        order = None
        if w["kind"] == "tophat" and world > 1:
            # cost-balanced partition of the GLOBAL batch (res spans 10..100 => 100x work spread per sample)
            gp = w["prior"](np.random.default_rng(2024), n * world)
            cost, _ = rd.afterglow_cost(gp["thv"], gp["thc"], gp["g0"])
            part = rd.balanced_partition(cost, world) if args.ag_partition == "balanced" else \
                [np.arange(*rd.shard_bounds(n * world, r, world)) for r in range(world)]
            order = part[rank]
            p = {k: v[order] for k, v in gp.items()}
            n = len(order)
        params = path.pack(p, n)
        logl = torch.empty(n, dtype=torch.float64, device=dev)
        line = {"key": w["key"], "config": w["config"], "workload": w["desc"], "samples_per_gpu": n, "epochs": E,
                "dtype": "f32-quadrature/f64-epilogue" if w.get("fp32") else "f64", "l2": "flushed between launches"}
        sampler = ClockSampler(local) if rank == 0 else None
        if w.get("population"):
            ctx = w["ctx"]
            lim = ctx["limiting_mag"]

            def fn():
                mags, _ = path.evaluate(params, want_model=True, want_logl=False)
                rb.simulate.add_optical_noise_and_cuts(mags, ctx["bands"], lim, seed=99,
                                                       outputs=("magnitude", "magnitude_error"))
        else:
            def fn():
                path.evaluate(params, want_model=False, want_logl=True, out_logl=logl)
        l0 = lib.rb_launch_count()
        ms, steps, clocks = time_launches(fn, dev, world, flush, args.path_seconds, sampler=sampler)
        line["gpu_launches_per_step"] = int((lib.rb_launch_count() - l0) // (steps + 4))
        rate = world * n * E / (ms * 1e-3)
        line.update(units_per_s=rate, ms_per_step=ms, steps=steps, clocks=clocks)
        flop_eq = w["flop_eq"]
        if w["kind"] == "tophat":
            flop_eq, line["res"] = bp.afterglow_flop_eq(p, E)
            line["partition"] = args.ag_partition if world > 1 else "single GPU"
        line["roofline"] = {"bound": "fp64", "flop_eq_per_unit": flop_eq, "achieved": rate / world * flop_eq / 1e12,
                            "peak": peak_tflops, "unit": "TFLOP/s",
                            "frac": rate / world * flop_eq / 1e12 / peak_tflops if peak_tflops else None}
        if w.get("population"):
            # end to end: host parameter table -> chunks -> magnitudes -> noise + cut -> streamed to pinned host memory
            host_p = {k: np.ascontiguousarray(v) for k, v in p.items()}
            seen = [0, 0]

            def sink(lo, hi, arrays):
                seen[0] += hi - lo
                seen[1] += int(arrays["detected"].sum())
            def population():
                rb.simulate.simulate_population_with_cadence(
                    rb.supernova_models.arnett, host_p, ctx["t"], ctx["bands"], lim, seed=99, sink=sink, chunk=65536,
                    outputs=("magnitude", "magnitude_error"), sample_base=rank * n, device=dev)
            population()                     # warm-up: pins the staging ring, builds the band tables
            seen[0] = seen[1] = 0
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            population()
            wall = _allmax(time.perf_counter() - t0, dev, world)
            line["e2e"] = {"units_per_s": world * n * E / wall, "host_bytes_out_per_sample": E * (8 * 3 + 1),
                           "transients": seen[0], "detected_observations": seen[1],
                           "what": "simulate_population_with_cadence(sink=...), 65536-transient chunks, D2H of "
                                   "magnitude, error, model magnitude, detected flag through a pinned double buffer; one warm-up call"}
        if par_n:
            line["parity"], line["cpu_port"] = bp.parity_gate(rb, w, path, p, y, sigma, par_n, bandpasses, dev,
                                                              budget_s=args.parity_budget)
        out.append(line)
        del path, params, logl
        torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import redback_b200 as rb
    from redback_b200 import _capi
    from redback_b200 import distributed as rd
    from redback_b200.engine import SupernovaPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- redback_b200 has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()
    if args.scaling == "strong":
        lo, hi = rd.shard_bounds(args.samples, rank, world)     # `--samples` is the TOTAL batch (BASELINE: 10^7)
        n = hi - lo
    else:
        n = args.samples                                        # per GPU
    n_total = args.samples if args.scaling == "strong" else world * n
    t, nu = epochs()
    E = t.size
    sed_id = _capi.SED_BLACKBODY if args.sed == "blackbody" else _capi.SED_CUTOFF_BLACKBODY

    # data: our own model at a fixed truth + 10 % noise
    mk = SupernovaPath(_capi.ENGINE_MAGNETAR, sed_id, t, frequency=nu, device=dev)
    y_model, _ = mk.evaluate(mk.pack(TRUTH, 1))
    y_model = y_model[0].cpu().numpy()
    rng = np.random.default_rng(0)
    y = y_model + rng.normal(0, 0.1 * np.abs(y_model))
    sigma = 0.1 * np.abs(y_model) + 1e-12
    path = SupernovaPath(_capi.ENGINE_MAGNETAR, sed_id, t, frequency=nu, y=y, sigma=sigma, device=dev,
                         dtype=args.dtype)

    params = prior_rows_torch(n, dev, seed=1000 + rank)     # resident in HBM before the timed region
    # library-level multi-GPU step (redback_b200.distributed): the local shard in `chunks` pieces, the all-gather of
    # piece i on a side stream under the kernels of piece i+1
    chunks = args.gather_chunks if world > 1 else 1
    sizes = [rd.shard_bounds(n, c, chunks) for c in range(chunks)]
    ev_probe = rd.ShardedEvaluator(n, lambda a, b, out: None, device=dev, chunks=chunks)
    csz = ev_probe.chunk
    blocks = [params[:, a:min(a + csz, n)].contiguous() for a in range(0, n, csz)] if world > 1 else [params]
    if world > 1:
        del params

    def evaluate(a, b, out):
        path.evaluate(blocks[a // csz] if world > 1 else blocks[0], want_model=False, want_logl=True, out_logl=out)

    sharded = rd.ShardedEvaluator(n, evaluate, device=dev, chunks=chunks)
    del ev_probe, sizes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        sharded.log_likelihood()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.rb_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        sharded.log_likelihood()
        ev[i + 1].record()
    barrier()
    launches = lib.rb_launch_count() - launches0
    total_ms = _allmax(ev[0].elapsed_time(ev[-1]), dev, world)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = n_total * E / (ms_per_step * 1e-3)
    logl = sharded.local[:n].clone()
    finite = bool(torch.isfinite(logl).all().item())

    # kernel-only duration (no collective) for the roofline: the same launches without the gather, timed per launch
    kev = _events(max(3, min(args.steps, 5)))
    tmp = torch.empty(n, dtype=torch.float64, device=dev)
    for a, b in kev:
        a.record()
        for bi, blk in enumerate(blocks):
            path.evaluate(blk, want_model=False, want_logl=True, out_logl=tmp[bi * csz: bi * csz + blk.shape[1]])
        b.record()
    barrier()
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)

    # ---- end to end through the public entry points, inputs in ordinary (pageable) host memory ----------------
    host_rows = torch.cat(blocks, dim=1).cpu().numpy() if world > 1 else blocks[0].cpu().numpy()
    names = ["redshift", "p0", "bp", "mass_ns", "theta_pb", "mej", "vej", "kappa", "kappa_gamma", "temperature_floor"]
    batch = {k: np.ascontiguousarray(host_rows[i]) for i, k in enumerate(names)}     # a params_batch dict
    like = rb.GaussianLikelihood(t, y, sigma, rb.supernova_models.slsn,
                                 kwargs=dict(frequency=nu, output_format="flux_density", dtype=args.dtype,
                                             sed="Blackbody" if args.sed == "blackbody" else "CutoffBlackbody",
                                             device=dev))
    like.log_likelihood_batch({k: v[:4096] for k, v in batch.items()})               # warm-up (path + pipeline set-up)
    like.log_likelihood_batch(batch)
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_logl = like.log_likelihood_batch(batch)
    e2e_s = _allmax((time.perf_counter() - t0) / e2e_steps, dev, world)
    e2e_value = n_total * E / e2e_s
    same = bool(np.array_equal(host_logl, logl.cpu().numpy()))
    # the C ABI's host entry as a non-Python caller would bind it (ctypes, plain pointers)
    cfg = _capi.sn_config(_capi.ENGINE_MAGNETAR, sed_id, precision=1 if args.dtype == "float32" else 0)
    rows_c = np.ascontiguousarray(host_rows)
    out_c = np.empty(n)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _capi.check(lib.rb_sn_eval_f64_host(C.byref(cfg), n, E, vp(rows_c), vp(t), vp(nu), None, vp(y), vp(sigma), None,
                                            None, vp(out_c)))
    capi_s = _allmax((time.perf_counter() - t0) / e2e_steps, dev, world)
    same_c = bool(np.array_equal(out_c, host_logl))
    del like, batch, rows_c, host_rows

    peak, pms = C.c_double(0), C.c_double(0)
    _capi.check(lib.rb_measure_fp64_peak(1 << 16, C.byref(peak), C.byref(pms)))
    paths = None
    if args.paths == "on" or (args.paths == "auto"):
        del blocks, sharded, tmp
        torch.cuda.empty_cache()
        paths = run_paths(args, dev, rank, world, local, peak.value, lib)

    if rank == 0:
        flop_eq = FLOP_EQ[args.sed]
        achieved = n * E * flop_eq / (kernel_ms * 1e-3) / 1e12
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        hbm_bytes = n * (10 * 8 + 8)
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
            "frac": achieved / peak.value if peak.value else None,
            # dram__bytes_read+write of rb::sn_kernel from the committed ncu --set full capture, scaled to this launch
            "traffic": 136.5 * n, "traffic_source": "ncu capture at 200000 samples, scaled per sample "
                                                    "(profiles/r2l_sn_headline_ncu_summary.txt: 27.3 MB read, writes absorbed by L2)",
            "kernel": "rb::sn_kernel", "kernel_ms": kernel_ms,
            "flop_eq_per_unit": flop_eq,
            "peak_source": "DFMA micro-benchmark measured live in this run (rb_measure_fp64_peak); nominal 37.2",
            "hbm": {"achieved_gbs": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": peaks.get("hbm_gbs", 6650.0),
                    "peak_source": "measured" if peaks else "fallback", "algorithmic_bytes_per_launch": hbm_bytes},
        }
        cpu = None
        if not args.no_cpu_baseline:
            rate, cores, wall = cpu_reference_rate(args.sed, args.cpu_samples_per_core, 7)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cores * args.cpu_samples_per_core} prior draws x {E} epochs, one process per core, "
                             f"{wall:.1f} s wall"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if args.dtype == "float64" else "f32-quadrature/f64-epilogue", "data": "synthetic",
            "config": workload_config(args, n, n_total), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 10 * 8),
                    "d2h_bytes_per_step": int(n * 8), "matches_resident_run": same,
                    "entry": "redback_b200.GaussianLikelihood.log_likelihood_batch(params_batch = dict of pageable numpy "
                             "arrays) -> rb_sn_eval_f64_host_rows (chunked two-stream pipeline, pinned staging in the "
                             "library)",
                    "capi_host": {"value": n_total * E / capi_s, "entry": "rb_sn_eval_f64_host via ctypes, pageable "
                                  "numpy block [10][N]", "matches": same_c}},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "all_logl_finite": finite,
            "multi_gpu": {"api": "redback_b200.distributed.ShardedEvaluator", "gather_chunks": chunks,
                          "collective": "all_gather_into_tensor of log L per chunk on a side stream"} if world > 1
            else None,
            "paths": paths,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--samples", type=int, default=10_000_000,
                    help="prior draws per GPU (--scaling weak) or in total (--scaling strong)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak")
    ap.add_argument("--gather-chunks", type=int, default=4, help="pieces of the local shard whose all-gather overlaps "
                                                                 "the next piece's kernels (N > 1)")
    ap.add_argument("--paths", choices=["auto", "on", "off"], default="auto",
                    help="also run every other BASELINE config (bench_paths.py); parity gates at N = 1 only")
    ap.add_argument("--paths-scale", type=float, default=1.0, help="scale the secondary workloads' sample counts")
    ap.add_argument("--path-seconds", type=float, default=1.5, help="timed region per secondary workload")
    ap.add_argument("--parity-samples", type=int, default=10_000, help="prior draws per config checked against the "
                                                                        "CPU oracle in-run (BASELINE.md section 5.5)")
    ap.add_argument("--parity-budget", type=float, default=75.0, help="wall-clock cap per config for the oracle [s]")
    ap.add_argument("--ag-partition", choices=["balanced", "contiguous"], default="balanced")
    ap.add_argument("--only", default="", help="restrict the secondary workloads to keys containing this string")
    ap.add_argument("--sed", choices=["blackbody", "cutoff"], default="blackbody")
    ap.add_argument("--dtype", choices=["float64", "float32"], default="float64",
                    help="float32: FP32 diffusion quadrature (RB_PREC_F32); the headline number is float64")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--cpu-samples-per-core", type=int, default=8000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
