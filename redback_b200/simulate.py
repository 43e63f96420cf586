"""Batched counterpart of redback.simulate_transients for optical surveys: ONE call evaluates a population of N
transients, adds the reference's noise model and applies its SNR cut, all on device.

* `simulate_population_with_cadence` -- SimulateTransientWithCadence in optical mode
  (redback/simulate_transients.py:136-610): every transient is observed on the same cadence table.
* `simulate_survey_observations` -- SimulateOpticalTransient._make_observation_single / _make_observations_for_population
  (:1096-1178): every transient has its OWN list of pointings (time, filter, 5-sigma depth), given in CSR form.
* `population_light_curves` -- the light-curve loop of PopulationSynthesizer.simulate_population (:2027-2036).

The reference loops over transients in Python (:665-667, 1168-1178, 2029-2036); here the model is a `params_batch`
call and the noise / cut is one elementwise kernel (csrc/noise_kernel.cu) whose normal variates come from a
counter-based Philox generator keyed by (seed, global element index): a population can be processed in chunks, in
any order and on any number of GPUs with identical results.  Large populations are streamed: outputs of chunk i are
copied to pinned host memory (or handed to a `sink`) while chunk i+1 is computed, so 10^7 transients never need
N x E device or host arrays at once.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi, _fused, bands as _bands
from .engine import _device, _ptr, as_device_f64, batch_size

_F64 = torch.float64
OUTPUT_KEYS = ("magnitude", "magnitude_error", "flux", "flux_error", "snr", "detected")


def _reference_flux(name):
    b = _bands.get_bandpass(name)
    if b.reference_flux is None:
        raise KeyError(f"Band {name} is not defined in filters.csv!")   # utils.py:869
    return b.reference_flux


def add_optical_noise_and_cuts(model_magnitude, bands, limiting_mag=None, noise_type="limiting_mag",
                               snr_threshold=5.0, apply_snr_cut=True, seed=None, standard_normal=None,
                               sample_base=0, outputs=OUTPUT_KEYS, return_variates=False):
    """_add_optical_noise_and_cuts (simulate_transients.py:457-527) for a (N, E) device tensor of model magnitudes.

    The N(0,1) draws come from the device Philox stream (`seed`, element index (sample_base + n) * E + e) unless
    `standard_normal` (an (N, E) tensor) is given.  `outputs` selects which arrays are materialised.  Returns a dict of
    device tensors (+ 'standard_normal' when `return_variates`)."""
    if noise_type not in _capi.NOISE_TYPES:
        raise ValueError(f"unknown noise_type {noise_type!r}")
    mm = model_magnitude.contiguous()
    N, E = mm.shape
    dev = mm.device
    names = np.broadcast_to(np.asarray(bands), (E,))
    ref = lim = None
    if noise_type == "limiting_mag":
        if limiting_mag is None:
            raise ValueError("limiting_mag is required for noise_type='limiting_mag'")
        ref = as_device_f64(np.array([_reference_flux(n) for n in names.tolist()], dtype=np.float64), dev)
        lim = as_device_f64(np.broadcast_to(np.asarray(limiting_mag, dtype=np.float64), (E,)).copy(), dev)
    want = set(outputs) | {"magnitude", "magnitude_error", "detected"}
    out = {k: torch.empty((N, E), dtype=_F64, device=dev) for k in OUTPUT_KEYS[:5] if k in want}
    out["detected"] = torch.empty((N, E), dtype=torch.uint8, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        if standard_normal is not None:
            z = standard_normal.to(device=dev, dtype=_F64).contiguous()
            if z.shape != (N, E):
                raise ValueError("standard_normal must have shape (N, E)")
            if "snr" not in out:
                out["snr"] = torch.empty((N, E), dtype=_F64, device=dev)
            _capi.check(_capi.lib().rb_optical_noise_f64(
                _capi.NOISE_TYPES[noise_type], N, E, _ptr(mm), _ptr(ref), _ptr(lim), _ptr(z), float(snr_threshold),
                1 if apply_snr_cut else 0, _ptr(out["magnitude"]), _ptr(out["magnitude_error"]), _ptr(out.get("flux")),
                _ptr(out.get("flux_error")), _ptr(out["snr"]), _ptr(out["detected"]), stream))
        else:
            zout = torch.empty((N, E), dtype=_F64, device=dev) if return_variates else None
            _capi.check(_capi.lib().rb_optical_noise_philox_f64(
                _capi.NOISE_TYPES[noise_type], N, E, int(sample_base), C.c_uint64(0 if seed is None else int(seed)),
                _ptr(mm), _ptr(ref), _ptr(lim), float(snr_threshold), 1 if apply_snr_cut else 0,
                _ptr(out["magnitude"]), _ptr(out["magnitude_error"]), _ptr(out.get("flux")),
                _ptr(out.get("flux_error")), _ptr(out.get("snr")), _ptr(out["detected"]), _ptr(zout), stream))
            if return_variates:
                out["standard_normal"] = zout
    out["detected"] = out["detected"].bool()
    return out


class _ModelRunner:
    """A model function bound to fixed epochs / kwargs: the device path is built once, parameter chunks stream through."""

    def __init__(self, model, times, kwargs):
        from .likelihoods import FUSED
        if model not in FUSED:
            raise NotImplementedError(f"{getattr(model, '__name__', model)} has no fused device path")
        self.kwargs = dict(kwargs)
        self.spec = FUSED[model].for_kwargs(self.kwargs)
        self.path = self.spec.build(np.asarray(times, dtype=np.float64), self.kwargs)

    def __call__(self, params_chunk):
        params, n = _fused.collect({}, self.kwargs, params_chunk, self.spec.needed, self.spec.defaults)
        params = self.spec.transform(params)
        dev = self.path.pack(params, n)
        model, _ = self.path.evaluate(dev, dl=_fused.dl_from_kwargs(self.kwargs, self.path, n),
                                      t0=_fused.t0_tensor(params, self.path, n),
                                      aux=_fused.aux_tensor(params, self.path, n))
        return model


def _slice(params_batch, lo, hi, n):
    return {k: (v[lo:hi] if (isinstance(v, torch.Tensor) and v.dim() > 0 and v.numel() == n) or
                (not isinstance(v, torch.Tensor) and np.ndim(v) > 0 and len(v) == n) else v)
            for k, v in params_batch.items()}


_PINNED = {}     # (slot, output name) -> pinned staging buffer, kept across calls: page-locking ~1 GB costs ~0.5 s


class HostRing:
    """Two pinned host buffers per output and two copy streams: chunk i is copied device->host while chunk i+1 is
    computed; `consume(lo, hi, {name: ndarray view})` is called once a chunk has landed.  The pinned buffers are
    process-wide and reused by later calls (one population simulation at a time per process)."""

    def __init__(self, device, consume):
        self.device = device
        self.consume = consume
        self.slots = [dict(stream=torch.cuda.Stream(device), event=None, buf=_PINNED.setdefault(k, {}), span=None,
                           keep=None) for k in range(2)]
        self.i = 0

    def _flush(self, slot):
        if slot["span"] is None:
            return
        slot["event"].synchronize()
        lo, hi = slot["span"]
        self.consume(lo, hi, {k: b[: hi - lo].numpy() for k, b in slot["buf"].items()})
        slot["span"] = slot["keep"] = None

    def push(self, lo, hi, tensors):
        slot = self.slots[self.i % 2]
        self.i += 1
        self._flush(slot)
        cur = torch.cuda.current_stream(self.device)
        slot["stream"].wait_stream(cur)
        with torch.cuda.stream(slot["stream"]):
            for k, t in tensors.items():
                buf = slot["buf"].get(k)
                if buf is None or buf.shape[0] < t.shape[0] or buf.shape[1:] != t.shape[1:] or buf.dtype != t.dtype:
                    buf = slot["buf"][k] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                buf[: t.shape[0]].copy_(t, non_blocking=True)
            slot["event"] = torch.cuda.Event()
            slot["event"].record(slot["stream"])
        slot["span"] = (lo, hi)
        slot["keep"] = tensors          # the device tensors stay alive until the copy has run

    def finish(self):
        for k in (self.i % 2, (self.i + 1) % 2):
            self._flush(self.slots[k])


def simulate_population_with_cadence(model, params_batch, times, bands, limiting_mag=None, model_kwargs=None,
                                     noise_type="limiting_mag", snr_threshold=5.0, apply_snr_cut=True, seed=1234,
                                     device=None, chunk=None, sink=None, outputs=OUTPUT_KEYS, sample_base=0):
    """Population version of SimulateTransientWithCadence (optical mode): `model` is a redback_b200 model
    function, `params_batch` {name: array[N]}, (`times`, `bands`, `limiting_mag`) the shared cadence
    (time_since_t0 in days).

    Without `sink` the result is the noise dict plus 'model_magnitude' as (N, E) device tensors (the population is
    still evaluated in chunks).  With `sink(lo, hi, {name: ndarray})` every chunk's selected `outputs` (plus
    'model_magnitude') are streamed to the host through a pinned double buffer and handed over as numpy views that are
    valid during the call; nothing of size N x E is kept -- this is the 10^7-transient mode.  `sample_base` offsets the
    Philox element index (rank r of a sharded population passes its first global sample index)."""
    dev = _device(device)
    kw = dict(model_kwargs or {})
    kw.update(bands=np.asarray(bands), output_format="magnitude", device=dev)
    times = np.asarray(times, dtype=np.float64)
    n = batch_size(params_batch)
    if n is None:
        raise ValueError("params_batch must contain at least one array")
    E = times.size
    if chunk is None:
        chunk = max(1, min(n, (1 << 28) // (8 * E * (3 + len(outputs)))))
    runner = _ModelRunner(model, times, kw)
    ring = HostRing(dev, sink) if sink is not None else None
    parts = []
    with torch.cuda.device(dev):
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            mags = runner(_slice(params_batch, lo, hi, n))
            out = add_optical_noise_and_cuts(mags, bands, limiting_mag, noise_type, snr_threshold, apply_snr_cut, seed,
                                             sample_base=sample_base + lo, outputs=outputs)
            out = {k: v for k, v in out.items() if k in outputs or k == "detected"}
            out["model_magnitude"] = mags
            if ring is not None:
                ring.push(lo, hi, {k: (v.to(torch.uint8) if v.dtype == torch.bool else v) for k, v in out.items()})
            else:
                parts.append(out)
        if ring is not None:
            ring.finish()
            return None
    return {k: torch.cat([p[k] for p in parts]) for k in parts[0]}


def simulate_survey_observations(model, params_batch, t0_mjd, pointing_offsets, exp_mjd, filters, five_sigma_depth,
                                 model_kwargs=None, snr_threshold=5.0, add_source_noise=False, source_noise=0.02,
                                 seed=1234, device=None, chunk=16384, sample_base=0, return_variates=False):
    """SimulateOpticalTransient._make_observations_for_population (simulate_transients.py:1096-1178) for N transients
    in one pass: transient n was pointed at by rows pointing_offsets[n] : pointing_offsets[n + 1] of the flat pointing
    table (`exp_mjd`, `filters` = band names, `five_sigma_depth`, sorted by time within a transient like the
    reference's `sort_values(by=['expMJD'])`), and exploded at `t0_mjd[n]`.  The model's spectral grid, the bicubic
    spline, the band integrals at the transient's own phases, the noise draws and the detection flags are evaluated
    on device; results come back as flat arrays aligned with the pointing table (the columns of the reference's
    observation DataFrame).  Supernova-family models (the magnitude branch of arnett, type_1a, slsn, ...)."""
    dev = _device(device)
    off = np.asarray(pointing_offsets, dtype=np.int64)
    n = off.size - 1
    exp_mjd = np.asarray(exp_mjd, dtype=np.float64)
    depth = np.asarray(five_sigma_depth, dtype=np.float64)
    names = np.asarray(filters)
    uniq = sorted(set(names.tolist()))
    index = {b: i for i, b in enumerate(uniq)}
    band_flat = np.array([index[b] for b in names.tolist()], dtype=np.int32)
    ref = np.array([_reference_flux(b) for b in uniq], dtype=np.float64)
    t0 = np.broadcast_to(np.asarray(t0_mjd, dtype=np.float64), (n,))
    kw = dict(model_kwargs or {})
    kw.update(bands=np.array(uniq), output_format="magnitude", device=dev)
    runner = _ModelRunner(model, np.arange(1.0, len(uniq) + 1.0), kw)
    if not hasattr(runner.path, "evaluate_ragged"):
        raise NotImplementedError("per-transient pointings are available for the supernova-family magnitude paths")
    total = int(off[-1])
    cols = {k: np.full(total, np.nan) for k in ("model_magnitude", "magnitude", "e_magnitude", "flux", "flux_error")}
    cols["detected"] = np.zeros(total, dtype=bool)
    counts = np.diff(off).astype(np.int32)
    ref_dev = as_device_f64(ref, dev)
    lib = _capi.lib()
    with torch.cuda.device(dev):
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            cnt = counts[lo:hi]
            emax = max(1, int(cnt.max()) if cnt.size else 1)
            # padded [nc, emax] tables from the CSR rows
            col = np.arange(emax)[None, :]
            valid = col < cnt[:, None]
            src = np.where(valid, off[lo:hi, None] + col, 0)
            phase = np.where(valid, exp_mjd[src] - t0[lo:hi, None], 1.0)
            bidx = np.where(valid, band_flat[src], 0).astype(np.int32)
            dep = np.where(valid, depth[src], 0.0)
            nc = hi - lo
            params, _ = _fused.collect({}, runner.kwargs, _slice(params_batch, lo, hi, n), runner.spec.needed,
                                       runner.spec.defaults)
            params = runner.spec.transform(params)
            pdev = runner.path.pack(params, nc)
            phase_d = torch.from_numpy(np.ascontiguousarray(phase)).to(dev)
            bidx_d = torch.from_numpy(np.ascontiguousarray(bidx)).to(dev)
            cnt_d = torch.from_numpy(np.ascontiguousarray(cnt)).to(dev)
            dep_d = torch.from_numpy(np.ascontiguousarray(dep)).to(dev)
            start_d = torch.from_numpy(np.ascontiguousarray(off[lo:hi])).to(dev)
            mags = runner.path.evaluate_ragged(pdev, phase_d, bidx_d, cnt_d,
                                               dl=_fused.dl_from_kwargs(runner.kwargs, runner.path, nc))
            out = {k: torch.empty((nc, emax), dtype=_F64, device=dev) for k in ("magnitude", "e_magnitude", "flux",
                                                                                "flux_error")}
            det = torch.empty((nc, emax), dtype=torch.uint8, device=dev)
            zout = torch.empty((nc, emax), dtype=_F64, device=dev) if return_variates else None
            _capi.check(lib.rb_survey_noise_philox_f64(
                nc, emax, int(sample_base), C.c_uint64(int(seed)), _ptr(mags), _ptr(phase_d), _ptr(bidx_d),
                _ptr(cnt_d), _ptr(start_d), _ptr(dep_d), _ptr(ref_dev), float(source_noise) if add_source_noise else 0.0,
                float(snr_threshold), _ptr(out["magnitude"]), _ptr(out["e_magnitude"]), _ptr(out["flux"]),
                _ptr(out["flux_error"]), _ptr(det), _ptr(zout), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
            flat = slice(int(off[lo]), int(off[hi]))
            if return_variates:
                cols.setdefault("standard_normal", np.full(total, np.nan))[flat] = zout.cpu().numpy()[valid]
            cols["model_magnitude"][flat] = mags.cpu().numpy()[valid]
            for k, v in out.items():
                cols[k][flat] = v.cpu().numpy()[valid]
            cols["detected"][flat] = det.cpu().numpy()[valid].astype(bool)
    ref_per_row = ref[band_flat]
    cols.update(time=exp_mjd, band=names, time_days=exp_mjd - np.repeat(t0, counts), limiting_magnitude=depth,
                flux_limit=10.0 ** (depth / (-2.5)) * ref_per_row, offsets=off)
    return cols


def population_light_curves(model, parameters, times=None, model_kwargs=None, device_output=False):
    """The light-curve loop of PopulationSynthesizer.simulate_population (simulate_transients.py:2027-2036) as one
    batched call: `parameters` is the population table ({name: array[N]} or a pandas DataFrame as returned by
    generate_population), `times` defaults to np.linspace(0, 100, 100) like the reference.  Columns that are not
    parameters of `model` (sky position, event time, ...) are ignored.  Returns {'times': [E], 'flux': [N, E]}."""
    times = np.linspace(0, 100, 100) if times is None else np.asarray(times, dtype=np.float64)
    table = parameters.to_dict("list") if hasattr(parameters, "to_dict") else dict(parameters)
    batch = {}
    for k, v in table.items():          # the model picks the columns it needs; non-numeric ones cannot be parameters
        try:
            batch[k] = np.asarray(v, dtype=np.float64)
        except (TypeError, ValueError):
            continue
    kw = dict(model_kwargs or {})
    if device_output:
        kw["device_output"] = True
    return {"times": times, "flux": model(times, params_batch=batch, **kw)}
