"""Glue shared by the model modules and the likelihoods: how a redback-style model function maps onto a
fused device path (which parameters it needs, how to build the path from kwargs)."""
import numpy as np
import torch

from . import _capi
from .engine import batch_size


class FusedSpec:
    def __init__(self, name, needed, build, validate=None, transform=None, resolver=None, defaults=None):
        self.defaults = dict(defaults or {})  # parameters that are optional kwargs in the reference
        self._resolver = resolver            # kwargs -> FusedSpec | None: a variant selected by kwargs
        self.name = name
        self.needed = tuple(needed)          # parameter names, in any order
        self._build = build                  # (time, kwargs, y, sigma, sigma_mode) -> path
        self._validate = validate            # kwargs -> None (raises like the reference)
        self._transform = transform          # {name: value} -> {name: value}: derived parameters (e.g. ek -> vej)

    def for_kwargs(self, kwargs):
        alt = self._resolver(kwargs) if self._resolver is not None else None
        return self if alt is None else alt

    def transform(self, params):
        return params if self._transform is None else self._transform(params)

    def validate(self, kwargs):
        if self._validate is not None:
            self._validate(kwargs)

    def build(self, time, kwargs, y=None, sigma=None, sigma_mode=_capi.SIGMA_FIXED):
        self.validate(kwargs)
        return self._build(time, kwargs, y, sigma, sigma_mode)

    def cache_key(self, kwargs):
        """Key of the device-resident path built from `kwargs`: scalars by value, arrays (frequency, bands,
        lambda_array, ...) by a digest of their CONTENT, so in-place edits and recycled ids cannot alias."""
        return (self.name, tuple((k, content_key(kwargs[k])) for k in sorted(kwargs)))


def content_key(v):
    """Hashable stand-in for a kwarg / data value: the value itself for scalars, a content digest for arrays."""
    if isinstance(v, (int, float, str, bool, type(None))):
        return v
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    if isinstance(v, (np.ndarray, list, tuple)):
        arr = np.asarray(v)
        if arr.dtype == object:
            return ("obj", tuple(content_key(x) for x in arr.reshape(-1).tolist()))
        import hashlib
        return (arr.shape, arr.dtype.str, hashlib.blake2b(np.ascontiguousarray(arr).tobytes(), digest_size=12).digest())
    return ("id", id(v))


def collect(named, kwargs, params_batch, needed, defaults=None):
    """Merge positional/kwargs scalars and params_batch arrays into {name: value} for `needed`."""
    params = {}
    batch = dict(params_batch) if params_batch is not None else {}
    for name in needed:
        if name in batch:
            params[name] = batch[name]
        elif named.get(name) is not None:
            params[name] = named[name]
        elif name in kwargs:
            params[name] = kwargs[name]
        elif defaults is not None and name in defaults:
            params[name] = defaults[name]
        else:
            raise KeyError(name)  # same failure mode as the reference (missing kwarg)
    return params, batch_size(params)


def finish(model, N, kwargs):
    if kwargs.get("device_output", False):
        return model if N is not None else model[0]
    out = model.cpu().numpy()
    return out if N is not None else out[0]


def cosmology_from_kwargs(kwargs):
    cosmo = kwargs.get("cosmology")
    if cosmo is None or isinstance(cosmo, _capi.Cosmology):
        return cosmo
    try:  # astropy-like FlatLambdaCDM object
        H0 = getattr(cosmo.H0, "value", cosmo.H0)
        Tcmb0 = getattr(cosmo.Tcmb0, "value", cosmo.Tcmb0)
        m_nu = [float(getattr(m, "value", m)) for m in np.atleast_1d(getattr(cosmo.m_nu, "value", cosmo.m_nu))]
        m_nu = sorted(m_nu, reverse=True)[:3] + [0.0] * (3 - len(m_nu))
        return _capi.flat_lcdm(float(H0), float(cosmo.Om0), float(Tcmb0), float(cosmo.Neff), m_nu)
    except AttributeError as exc:
        raise NotImplementedError("cosmology must be a flat LambdaCDM-like object") from exc


def dl_from_kwargs(kwargs, path, N):
    """Optional precomputed luminosity distance [cm] (kwarg 'dl', scalar or array[N]); default: integrate
    the cosmology on device from the sampled redshift."""
    dl = kwargs.get("dl")
    if dl is None:
        return None
    n = N if N is not None else 1
    if isinstance(dl, torch.Tensor):
        return dl.to(device=path.device, dtype=torch.float64).reshape(-1).expand(n).contiguous()
    arr = np.array(np.broadcast_to(np.asarray(dl, dtype=np.float64), (n,)))
    return torch.from_numpy(arr).to(path.device)


def output_mode(kwargs, name, photometric=True):
    """kwargs['output_format'] -> 'flux_density' | 'magnitude' | 'flux' | 'spectra' (sed.py:971-1009).  'spectra' is the
    namedtuple (time, lambdas, spectra) evaluated on device (run -> engine.evaluate_spectra, no bands needed);
    'sncosmo_source' returns an sncosmo object in the reference and is not available."""
    fmt = kwargs.get("output_format")
    if fmt is None:
        raise KeyError("output_format")
    if fmt == "flux_density":
        return fmt
    if fmt in ("magnitude", "flux"):
        if not photometric:
            raise NotImplementedError(f"{name}: output_format='{fmt}' is not fused on device")
        if "bands" not in kwargs:
            raise KeyError("bands")
        return fmt
    if fmt == "spectra" and photometric:
        return fmt
    if fmt in ("spectra", "sncosmo_source"):
        raise NotImplementedError(f"{name}: output_format='{fmt}' is not fused on device")
    raise ValueError("Output format must be 'flux', 'magnitude', 'sncosmo_source', or 'flux_density'")


def flux_density_only(kwargs, name):
    output_mode(kwargs, name, photometric=False)


def t0_tensor(params, path, N):
    """Per-sample explosion epoch (phase_models.t0_base_model) as a device tensor [N], or None."""
    if "t0" not in params:
        return None
    n = N if N is not None else 1
    t0 = params["t0"]
    if isinstance(t0, torch.Tensor):
        return t0.to(device=path.device, dtype=torch.float64).reshape(-1).expand(n).contiguous()
    return torch.from_numpy(np.array(np.broadcast_to(np.asarray(t0, dtype=np.float64), (n,)))).to(path.device)


def aux_tensor(params, path, N):
    """The path's extra per-sample parameter (AUX_PARAM, e.g. f_nickel of magnetar_nickel) as a device tensor [N]."""
    name = getattr(path, "AUX_PARAM", None)
    if name is None:
        return None
    return t0_tensor({"t0": params[name]}, path, N)


def run_spectra(path, dev, params, N, kwargs, ext):
    """output_format='spectra': namedtuple('output', ['time', 'lambdas', 'spectra']) as in the reference
    (supernova_models.py:1019-1023).  Batched calls return time [N, n_t] and spectra [N, n_t, n_lambda]; where a sample's
    grid is shorter than n_t (kilonova grids, np.unique) the remaining rows are NaN; a scalar call is trimmed to its grid."""
    from collections import namedtuple
    if not hasattr(path, "phot") or "t0" in params:
        raise NotImplementedError("output_format='spectra' is not fused for this model")
    if ext is not None:
        path.set_extinction(ext["struct"], t0_tensor({"t0": ext["av_host"]}, path, N))
        ext["applied"] = True
    try:
        tt, spec, count = path.evaluate_spectra(dev, dl=dl_from_kwargs(kwargs, path, N), aux=aux_tensor(params, path, N))
    finally:
        if ext is not None:
            path.set_extinction(None, None)
    lam = np.array(path._keep[0])
    output = namedtuple("output", ["time", "lambdas", "spectra"])
    if N is None:
        g = int(count[0])
        tt, spec = tt[0, :g], spec[0, :g]
    if kwargs.get("device_output", False):
        return output(time=tt, lambdas=lam, spectra=spec)
    return output(time=tt.cpu().numpy(), lambdas=lam, spectra=spec.cpu().numpy())


def run(spec, time, named, kwargs, params_batch):
    spec = spec.for_kwargs(kwargs)
    params, N = collect(named, kwargs, params_batch, spec.needed, spec.defaults)
    params = spec.transform(params)
    spectra = kwargs.get("output_format") == "spectra"
    if spectra:
        kwargs = dict(kwargs, bands=None)              # the reference reads no bands on this branch
    path = spec.build(time, kwargs)
    dev = path.pack(params, N if N is not None else 1)
    ext = kwargs.get("_extinction")
    if spectra:
        return run_spectra(path, dev, params, N, kwargs, ext)
    if ext is not None:
        # extinction_models: the magnitude / flux branch attenuates the spectra inside the photometry kernel
        path.set_extinction(ext["struct"], t0_tensor({"t0": ext["av_host"]}, path, N))
        ext["applied"] = True
    try:
        model, _ = path.evaluate(dev, dl=dl_from_kwargs(kwargs, path, N), t0=t0_tensor(params, path, N),
                                 aux=aux_tensor(params, path, N))
    finally:
        if ext is not None:
            path.set_extinction(None, None)
    return finish(model, N, kwargs)
