"""Batched counterpart of redback.simulate_transients.SimulateTransientWithCadence for optical cadences
(redback/simulate_transients.py:136-610): ONE call evaluates a population of N transients on a shared cadence
(times, bands, limiting magnitudes), adds the reference's noise model and applies its SNR cut, all on device.

The reference loops over transients in Python (simulate_transients.py:665-667, 1168-1178, 2029-2036); here the
model is a `params_batch` call and the noise / cut is one elementwise kernel (csrc/noise_kernel.cu).
"""
import ctypes as C

import numpy as np
import torch

from . import _capi, bands as _bands
from .engine import _device, _ptr, as_device_f64


def add_optical_noise_and_cuts(model_magnitude, bands, limiting_mag=None, noise_type="limiting_mag",
                               snr_threshold=5.0, apply_snr_cut=True, seed=None, standard_normal=None):
    """_add_optical_noise_and_cuts (simulate_transients.py:457-527) for a (N, E) device tensor of model magnitudes.

    `standard_normal`: optional (N, E) tensor of N(0,1) draws (for reproducibility / parity tests); otherwise drawn
    on device from a Philox generator seeded with `seed`.  Returns a dict of device tensors:
    magnitude, magnitude_error, flux, flux_error, snr, detected, model_flux."""
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
        ref_np = np.array([_reference_flux(n) for n in names.tolist()], dtype=np.float64)
        ref = as_device_f64(ref_np, dev)
        lim = as_device_f64(np.broadcast_to(np.asarray(limiting_mag, dtype=np.float64), (E,)).copy(), dev)
    if standard_normal is None:
        gen = torch.Generator(device=dev)
        gen.manual_seed(0 if seed is None else int(seed))
        z = torch.randn((N, E), dtype=torch.float64, device=dev, generator=gen)
    else:
        z = standard_normal.to(device=dev, dtype=torch.float64).contiguous()
        if z.shape != (N, E):
            raise ValueError("standard_normal must have shape (N, E)")
    out = {k: torch.empty((N, E), dtype=torch.float64, device=dev)
           for k in ("magnitude", "magnitude_error", "flux", "flux_error", "snr")}
    out["detected"] = torch.empty((N, E), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().rb_optical_noise_f64(
            _capi.NOISE_TYPES[noise_type], N, E, _ptr(mm), _ptr(ref), _ptr(lim), _ptr(z), float(snr_threshold),
            1 if apply_snr_cut else 0, _ptr(out["magnitude"]), _ptr(out["magnitude_error"]), _ptr(out["flux"]),
            _ptr(out["flux_error"]), _ptr(out["snr"]), _ptr(out["detected"]),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    out["detected"] = out["detected"].bool()
    return out


def _reference_flux(name):
    b = _bands.get_bandpass(name)
    if b.reference_flux is None:
        raise KeyError(f"Band {name} is not defined in filters.csv!")   # utils.py:869
    return b.reference_flux


def simulate_population_with_cadence(model, params_batch, times, bands, limiting_mag=None, model_kwargs=None,
                                     noise_type="limiting_mag", snr_threshold=5.0, apply_snr_cut=True, seed=1234,
                                     device=None):
    """Population version of SimulateTransientWithCadence (optical mode): `model` is a redback_b200 model
    function, `params_batch` {name: array[N]}, (`times`, `bands`, `limiting_mag`) the shared cadence
    (time_since_t0 in days).  Returns the noise dict plus 'model_magnitude'."""
    dev = _device(device)
    kw = dict(model_kwargs or {})
    kw.update(bands=np.asarray(bands), output_format="magnitude", device_output=True, device=dev)
    mags = model(np.asarray(times, dtype=np.float64), params_batch=params_batch, **kw)
    out = add_optical_noise_and_cuts(mags, bands, limiting_mag, noise_type, snr_threshold, apply_snr_cut, seed)
    out["model_magnitude"] = mags
    return out


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
