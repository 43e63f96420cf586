"""Batched drop-ins for redback.transient_models.extinction_models (extinction_models.py:172-459): a base model with
host-galaxy and Milky-Way dust extinction applied.

    extinction_with_function / _supernova_base_model / _kilonova_base_model / _tde_base_model /
    _magnetar_driven_base_model / _afterglow_base_model(time, av_host, **kwargs)
    extinction_afterglow_galactic_dust_to_gas_ratio(time, lognh, factor=2.21, **kwargs)

kwargs as in the reference: `base_model` (name or function), every parameter of the base model incl. `redshift`,
`av_mw` (0), `rv_host` / `rv_mw` (3.1), `host_law` / `mw_law` ('fitzpatrick99'), `output_format`; plus the new
`params_batch` ({name: array[N]}; `av_host`, `redshift` and any base-model parameter may be sampled).

* `output_format='flux_density'` (:194-221): the base model's (N, E) flux densities are attenuated in place on device
  (`rb_extinction_apply_f64`): host extinction at lambda / (1 + z), Milky Way at lambda, lambda = c / frequency.
* `'magnitude'` / `'flux'` (:223-251): the reference attenuates the model SPECTRA per observer-frame wavelength and then
  runs the bandpass photometry; here the photometry kernel multiplies its per-column factors by the same attenuation
  (`rb_photometry.extinction`), so nothing of size N_t x N_lambda is formed.

The laws are the third-party `extinction` package's (not installable here): ccm89, odonnell94, calzetti00 and
fitzpatrick99 are restated from the papers (csrc/ext_kernel.cu) and anchored on the package's documented example values
(tests/test_extinction.py); 'fm07' is refused.
"""
import numpy as np
import torch

from . import _capi, _fused
from .engine import batch_size

_LAWS = {"ccm89": _capi.EXT_CCM89, "odonnell94": _capi.EXT_ODONNELL94, "calzetti00": _capi.EXT_CALZETTI00,
         "fitzpatrick99": _capi.EXT_FITZPATRICK99}
_ALL_LAWS = ("fitzpatrick99", "fm07", "calzetti00", "odonnell94", "ccm89")

# Fitzpatrick (1999): anchor points of the optical / infrared spline [1 / micron]
F99_XKNOTS = 1e4 / np.array([np.inf, 26500.0, 12200.0, 6000.0, 5470.0, 4670.0, 4110.0, 2700.0, 2600.0])


def f99_uv_k(x, rv):
    """Fitzpatrick & Massa (1990) ultraviolet curve with Fitzpatrick's (1999) parameters (x < 5.9: no far-UV term)."""
    x0, gamma, c3 = 4.596, 0.99, 3.23
    c2 = -0.824 + 4.717 / rv
    c1 = 2.030 - 3.007 * c2
    d = x ** 2 / ((x ** 2 - x0 ** 2) ** 2 + x ** 2 * gamma ** 2)
    return c1 + c2 * x + c3 * d


def f99_kknots(rv):
    rv2 = rv * rv
    k = np.empty(9)
    k[0] = -rv
    k[1] = 0.26469 * rv / 3.1 - rv
    k[2] = 0.82925 * rv / 3.1 - rv
    k[3] = -0.422809 + 1.00270 * rv + 2.13572e-04 * rv2 - rv
    k[4] = -5.13540e-02 + 1.00216 * rv - 7.35778e-05 * rv2 - rv
    k[5] = 0.700127 + 1.00184 * rv - 3.32598e-05 * rv2 - rv
    k[6] = 1.19456 + 1.01707 * rv - 5.46959e-03 * rv2 + 7.97809e-04 * rv2 * rv - 4.45636e-05 * rv2 * rv2 - rv
    k[7:] = f99_uv_k(F99_XKNOTS[7:], rv)
    return k


def f99_spline_table(rv):
    """9 breakpoints + 8 x 4 piecewise-cubic coefficients (highest power first, in x - breakpoint) of the NATURAL cubic
    spline (second derivative zero at both ends) through Fitzpatrick's anchor points: the table of rb_extinction.f99_*.
    The natural end condition is the one that reproduces the `extinction` package's documented values
    (fitzpatrick99([2000, 4000, 8000] A, 1.0, 3.1) = [2.76225609, 1.42325373, 0.55333671]); scipy's not-a-knot splrep
    differs from them in the fourth digit."""
    x, y = F99_XKNOTS, f99_kknots(rv)
    n = x.size
    h = np.diff(x)
    # second derivatives m[0] = m[n-1] = 0; tridiagonal system for the interior ones
    a = np.zeros((n, n))
    b = np.zeros(n)
    a[0, 0] = a[-1, -1] = 1.0
    for i in range(1, n - 1):
        a[i, i - 1], a[i, i], a[i, i + 1] = h[i - 1], 2.0 * (h[i - 1] + h[i]), h[i]
        b[i] = 6.0 * ((y[i + 1] - y[i]) / h[i] - (y[i] - y[i - 1]) / h[i - 1])
    m = np.linalg.solve(a, b)
    out = np.empty(41)
    out[:9] = x
    for i in range(n - 1):
        out[9 + 4 * i: 13 + 4 * i] = [(m[i + 1] - m[i]) / (6.0 * h[i]), m[i] / 2.0,
                                      (y[i + 1] - y[i]) / h[i] - h[i] * (2.0 * m[i] + m[i + 1]) / 6.0, y[i]]
    return out


def _extinction_struct(kwargs, av_mw):
    host_law, mw_law = kwargs.pop("host_law", "fitzpatrick99"), kwargs.pop("mw_law", "fitzpatrick99")
    rv_host, rv_mw = kwargs.pop("rv_host", 3.1), kwargs.pop("rv_mw", 3.1)
    if host_law not in _ALL_LAWS:                                           # extinction_models.py:112-114
        raise ValueError(f"Unknown host extinction law: {host_law}. Available: {list(_ALL_LAWS)}")
    if mw_law not in _ALL_LAWS:
        raise ValueError(f"Unknown MW extinction law: {mw_law}. Available: {list(_ALL_LAWS)}")
    for law in (host_law, mw_law):
        if law not in _LAWS:
            raise NotImplementedError(f"extinction law '{law}' is not restated on device (available: {sorted(_LAWS)})")
    if np.ndim(rv_host) or np.ndim(rv_mw) or np.ndim(av_mw):
        raise NotImplementedError("rv_host, rv_mw and av_mw must be scalars in the fused B200 path")
    ext = _capi.Extinction()
    ext.host_law, ext.mw_law = _LAWS[host_law], _LAWS[mw_law]
    # calzetti00 is called without R_V in the reference (:131-133): the package default 4.05
    ext.rv_host = 4.05 if host_law == "calzetti00" else float(rv_host)
    ext.rv_mw = 4.05 if mw_law == "calzetti00" else float(rv_mw)
    ext.av_mw = float(av_mw)
    ext.f99_host[:] = f99_spline_table(ext.rv_host).tolist()
    ext.f99_mw[:] = f99_spline_table(ext.rv_mw).tolist()
    return ext


_MODEL_MODULES = {"supernova": "supernova_models", "afterglow": "afterglow_models",
                  "magnetar_driven": "magnetar_driven_ejecta_models", "tde": "tde_models", "kilonova": "kilonova_models",
                  "integrated_flux_afterglow": "afterglow_models"}


def _get_correct_function(base_model, model_type=None):
    """extinction_models._get_correct_function (:17-75) over the fused models of this package."""
    import redback_b200 as rb
    if callable(base_model):
        return base_model
    if not isinstance(base_model, str):
        raise ValueError("base_model must be a string name or a callable function")
    if model_type is None:
        if base_model in rb.all_models_dict:
            return rb.all_models_dict[base_model]
        raise ValueError(f"Model '{base_model}' not found in any module. Ensure the model name is correct and the "
                         "module is loaded.")
    if model_type not in _MODEL_MODULES:
        raise ValueError(f"Unknown model_type '{model_type}'. Available types: {list(_MODEL_MODULES)}")
    module = getattr(rb, _MODEL_MODULES[model_type])
    fn = getattr(module, base_model, None)
    if fn is None or not callable(fn):
        raise ValueError(f"Model '{base_model}' not found in '{_MODEL_MODULES[model_type]}'")
    return fn


def _evaluate_extinction_model(time, av_host, av_mw=0.0, model_type=None, params_batch=None, **kwargs):
    """extinction_models._evaluate_extinction_model (:172-251)."""
    base_model = kwargs["base_model"]
    if kwargs["base_model"] in ["thin_shell_supernova", "homologous_expansion_supernova"]:
        kwargs["base_model"] = kwargs.get("submodel", "arnett_bolometric")               # :188-189
    else:
        kwargs.pop("base_model")
    ext = _extinction_struct(kwargs, av_mw)
    batch = dict(params_batch) if params_batch is not None else {}
    av = batch.pop("av_host", av_host)
    if av is None:
        raise KeyError("av_host")
    z = batch["redshift"] if "redshift" in batch else kwargs["redshift"]                  # KeyError like the reference (:100)
    N = batch_size(dict(batch, av_host=av))
    function = kwargs.pop("_function", None) or _get_correct_function(base_model=base_model, model_type=model_type)
    device_output = kwargs.pop("device_output", False)
    if N is not None and not batch:
        # only av_host is sampled: the base model still has to produce N rows
        batch = {"redshift": np.full(N, float(z))}
        kwargs = {k: v for k, v in kwargs.items() if k != "redshift"}
    call = dict(kwargs)
    if batch:
        call["params_batch"] = batch
    if kwargs["output_format"] == "flux_density":
        model = function(time, device_output=True, **call)
        if not isinstance(model, torch.Tensor):
            raise NotImplementedError(f"{getattr(function, '__name__', function)} has no device output")
        scalar = model.dim() == 1
        m2 = model.reshape(1, -1) if scalar else model
        n = m2.shape[0]
        dev = m2.device

        def column(v):
            if isinstance(v, torch.Tensor):
                return v.to(device=dev, dtype=torch.float64).reshape(-1).expand(n).contiguous()
            return torch.from_numpy(np.array(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)))).to(dev)
        freq = kwargs["frequency"]
        f = np.asarray(freq.detach().cpu().numpy() if isinstance(freq, torch.Tensor) else freq, dtype=np.float64)
        if f.ndim == 0:
            f = np.full(m2.shape[1], float(f))                                            # :196-197
        f_dev = torch.from_numpy(np.ascontiguousarray(f)).to(dev)
        z_dev, av_dev = column(z), column(av)
        m2 = m2.contiguous()
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().rb_extinction_apply_f64(
                _capi.C.byref(ext), n, m2.shape[1], f_dev.data_ptr(), z_dev.data_ptr(), av_dev.data_ptr(), m2.data_ptr(),
                _capi.C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        out = m2[0] if scalar else m2
        return out if device_output else out.cpu().numpy()
    state = dict(struct=ext, av_host=av, applied=False)
    out = function(time, device_output=device_output, _extinction=state, **call)
    if not state["applied"]:
        raise NotImplementedError(f"{getattr(function, '__name__', function)}: extinction of the magnitude / flux "
                                  "branch is not fused for this base model")
    return out


def extinction_with_function(time, av_host=None, params_batch=None, **kwargs):
    """redback extinction_models.extinction_with_function (:257-274)."""
    return _evaluate_extinction_model(time, av_host, model_type=None, params_batch=params_batch, **kwargs)


def extinction_with_supernova_base_model(time, av_host=None, params_batch=None, **kwargs):
    """redback extinction_models.extinction_with_supernova_base_model (:277-294)."""
    return _evaluate_extinction_model(time, av_host, model_type="supernova", params_batch=params_batch, **kwargs)


def extinction_with_kilonova_base_model(time, av_host=None, params_batch=None, **kwargs):
    """redback extinction_models.extinction_with_kilonova_base_model (:297-314)."""
    return _evaluate_extinction_model(time, av_host, model_type="kilonova", params_batch=params_batch, **kwargs)


def extinction_with_tde_base_model(time, av_host=None, params_batch=None, **kwargs):
    """redback extinction_models.extinction_with_tde_base_model (:317-334)."""
    return _evaluate_extinction_model(time, av_host, model_type="tde", params_batch=params_batch, **kwargs)


def extinction_with_magnetar_driven_base_model(time, av_host=None, params_batch=None, **kwargs):
    """redback extinction_models.extinction_with_magnetar_driven_base_model (:357-374)."""
    return _evaluate_extinction_model(time, av_host, model_type="magnetar_driven", params_batch=params_batch, **kwargs)


def extinction_with_afterglow_base_model(time, av_host=None, params_batch=None, **kwargs):
    """redback extinction_models.extinction_with_afterglow_base_model (:417-434)."""
    return _evaluate_extinction_model(time, av_host, model_type="afterglow", params_batch=params_batch, **kwargs)


def extinction_afterglow_galactic_dust_to_gas_ratio(time, lognh=None, factor=2.21, params_batch=None, **kwargs):
    """redback extinction_models.extinction_afterglow_galactic_dust_to_gas_ratio (:438-459): av = 10**lognh / (factor 1e21)."""
    batch = dict(params_batch) if params_batch is not None else {}
    lognh = batch.pop("lognh", lognh)
    if lognh is None:
        raise KeyError("lognh")
    lg = lognh.detach().cpu().numpy() if isinstance(lognh, torch.Tensor) else np.asarray(lognh, dtype=np.float64)
    av = 10 ** lg / (factor * 1e21)
    if np.ndim(av):
        batch["av_host"] = av
        return extinction_with_afterglow_base_model(time, None, params_batch=batch, **kwargs)
    return extinction_with_afterglow_base_model(time, float(av), params_batch=batch or None, **kwargs)


MODELS = (extinction_with_function, extinction_with_supernova_base_model, extinction_with_kilonova_base_model,
          extinction_with_tde_base_model, extinction_with_magnetar_driven_base_model,
          extinction_with_afterglow_base_model, extinction_afterglow_galactic_dust_to_gas_ratio)
