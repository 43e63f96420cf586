"""Batched drop-ins for redback's diffusion-powered supernova models.

Same names, positional parameters and kwargs as redback/transient_models/supernova_models.py
(arnett_bolometric :949-969, arnett :971-1028, basic_magnetar_powered_bolometric :1447-1469,
basic_magnetar_powered :1471-1530, slsn_bolometric :1532-1549, slsn :1552-1624) plus the new
`params_batch` entry point: a dict {name: array[N]} (numpy or torch) of sampled parameters.  Any model
parameter that is not in `params_batch` is taken from the call (positional / kwargs) as a scalar.
With `params_batch` the result has shape (N, E); without it the scalar call returns (E,), as in redback.
"""
import numpy as np
import torch

from . import _capi
from .engine import SupernovaPath, batch_size

_ALLOWED_OPERATORS = {
    "interaction_process": ("Diffusion",),
    "photosphere": ("TemperatureFloor",),
    "sed": ("Blackbody", "CutoffBlackbody"),
}
_SED_IDS = {"Blackbody": _capi.SED_BLACKBODY, "CutoffBlackbody": _capi.SED_CUTOFF_BLACKBODY}


def _operator_name(kind, value, default):
    """kwargs 'interaction_process' / 'photosphere' / 'sed' (supernova_models.py:993-995).  Only the
    default operator classes are fused on device; anything else raises (no CPU fallback)."""
    if value is None and kind != "interaction_process":
        return default
    name = value if isinstance(value, str) else getattr(value, "__name__", None)
    if value is None or name not in _ALLOWED_OPERATORS[kind]:
        raise NotImplementedError(
            f"{kind}={value!r} is not available in the fused B200 path (supported: "
            f"{_ALLOWED_OPERATORS[kind]}); redback_b200 has no CPU fallback")
    return name


def _collect(named, kwargs, params_batch, needed):
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
        else:
            raise KeyError(name)  # same failure mode as the reference (missing kwarg)
    N = batch_size(params)
    return params, N


def _finish(model, N, kwargs):
    if kwargs.get("device_output", False):
        return model if N is not None else model[0]
    out = model.cpu().numpy()
    return out if N is not None else out[0]


_DIFFUSION = ("mej", "vej", "kappa", "kappa_gamma")


def _run(engine, sed_id, engine_params, time, named, kwargs, params_batch, redshift_needed):
    needed = list(engine_params) + list(_DIFFUSION)
    if redshift_needed:
        needed = ["redshift"] + needed + ["temperature_floor"]
    params, N = _collect(named, kwargs, params_batch, needed)
    path = SupernovaPath(engine, sed_id, time, frequency=kwargs.get("frequency") if redshift_needed else None,
                         dense_resolution=kwargs.get("dense_resolution", 1000),
                         cutoff_wavelength=kwargs.get("cutoff_wavelength", 3000),
                         absorption_index=kwargs.get("absorption_index", 1.0),
                         cosmology=_cosmology(kwargs), device=kwargs.get("device"))
    dev = path.pack(params, N if N is not None else 1)
    model, _ = path.evaluate(dev, dl=_dl(kwargs, path, N))
    return _finish(model, N, kwargs)


def _cosmology(kwargs):
    cosmo = kwargs.get("cosmology")
    if cosmo is None or isinstance(cosmo, _capi.Cosmology):
        return cosmo
    # astropy-like FlatLambdaCDM object
    try:
        H0 = getattr(cosmo.H0, "value", cosmo.H0)
        Tcmb0 = getattr(cosmo.Tcmb0, "value", cosmo.Tcmb0)
        m_nu = [float(getattr(m, "value", m)) for m in np.atleast_1d(getattr(cosmo.m_nu, "value", cosmo.m_nu))]
        m_nu = sorted(m_nu, reverse=True)[:3] + [0.0] * (3 - len(m_nu))
        return _capi.flat_lcdm(float(H0), float(cosmo.Om0), float(Tcmb0), float(cosmo.Neff), m_nu)
    except AttributeError as exc:
        raise NotImplementedError("cosmology must be a flat LambdaCDM-like object") from exc


def _dl(kwargs, path, N):
    """Optional precomputed luminosity distance [cm] (kwarg 'dl', scalar or array[N]); default: integrate
    the cosmology on device from the sampled redshift."""
    dl = kwargs.get("dl")
    if dl is None:
        return None
    n = N if N is not None else 1
    if isinstance(dl, torch.Tensor):
        return dl.to(device=path.device, dtype=torch.float64).reshape(-1).expand(n).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.broadcast_to(np.asarray(dl, dtype=np.float64), (n,)))).to(path.device)


def _flux_density_only(kwargs, name):
    fmt = kwargs.get("output_format")
    if fmt is None:
        raise KeyError("output_format")
    if fmt != "flux_density":
        if fmt in ("magnitude", "flux", "spectra", "sncosmo_source"):
            raise NotImplementedError(f"{name}: output_format='{fmt}' is not fused on device yet")
        raise ValueError("Output format must be 'flux', 'magnitude', 'sncosmo_source', or 'flux_density'")


def arnett_bolometric(time, f_nickel=None, mej=None, params_batch=None, **kwargs):
    """redback supernova_models.arnett_bolometric (:949-969); returns erg/s."""
    _operator_name("interaction_process", kwargs.get("interaction_process", "Diffusion"), "Diffusion")
    return _run(_capi.ENGINE_NICKEL, _capi.SED_BOLOMETRIC, ("f_nickel",), time,
                dict(f_nickel=f_nickel, mej=mej), kwargs, params_batch, False)


def arnett(time, redshift=None, f_nickel=None, mej=None, params_batch=None, **kwargs):
    """redback supernova_models.arnett (:971-1028), output_format='flux_density'; returns mJy."""
    _flux_density_only(kwargs, "arnett")
    _operator_name("interaction_process", kwargs.get("interaction_process", "Diffusion"), "Diffusion")
    _operator_name("photosphere", kwargs.get("photosphere"), "TemperatureFloor")
    sed = _operator_name("sed", kwargs.get("sed"), "Blackbody")
    return _run(_capi.ENGINE_NICKEL, _SED_IDS[sed], ("f_nickel",), time,
                dict(redshift=redshift, f_nickel=f_nickel, mej=mej), kwargs, params_batch, True)


_MAGNETAR = ("p0", "bp", "mass_ns", "theta_pb")


def basic_magnetar_powered_bolometric(time, p0=None, bp=None, mass_ns=None, theta_pb=None, params_batch=None,
                                      **kwargs):
    """redback supernova_models.basic_magnetar_powered_bolometric (:1447-1469); returns erg/s."""
    _operator_name("interaction_process", kwargs.get("interaction_process", "Diffusion"), "Diffusion")
    return _run(_capi.ENGINE_MAGNETAR, _capi.SED_BOLOMETRIC, _MAGNETAR, time,
                dict(p0=p0, bp=bp, mass_ns=mass_ns, theta_pb=theta_pb), kwargs, params_batch, False)


def slsn_bolometric(time, p0=None, bp=None, mass_ns=None, theta_pb=None, params_batch=None, **kwargs):
    """redback supernova_models.slsn_bolometric (:1532-1549)."""
    return basic_magnetar_powered_bolometric(time, p0, bp, mass_ns, theta_pb, params_batch=params_batch, **kwargs)


def _magnetar_flux_density(name, default_sed, time, redshift, p0, bp, mass_ns, theta_pb, params_batch, kwargs):
    _flux_density_only(kwargs, name)
    _operator_name("interaction_process", kwargs.get("interaction_process", "Diffusion"), "Diffusion")
    _operator_name("photosphere", kwargs.get("photosphere"), "TemperatureFloor")
    sed = _operator_name("sed", kwargs.get("sed"), default_sed)
    return _run(_capi.ENGINE_MAGNETAR, _SED_IDS[sed], _MAGNETAR, time,
                dict(redshift=redshift, p0=p0, bp=bp, mass_ns=mass_ns, theta_pb=theta_pb), kwargs,
                params_batch, True)


def basic_magnetar_powered(time, redshift=None, p0=None, bp=None, mass_ns=None, theta_pb=None,
                           params_batch=None, **kwargs):
    """redback supernova_models.basic_magnetar_powered (:1471-1530); default sed = Blackbody; mJy."""
    return _magnetar_flux_density("basic_magnetar_powered", "Blackbody", time, redshift, p0, bp, mass_ns,
                                  theta_pb, params_batch, kwargs)


def slsn(time, redshift=None, p0=None, bp=None, mass_ns=None, theta_pb=None, params_batch=None, **kwargs):
    """redback supernova_models.slsn (:1552-1624); default sed = CutoffBlackbody (3000 A, index 1); mJy."""
    return _magnetar_flux_density("slsn", "CutoffBlackbody", time, redshift, p0, bp, mass_ns, theta_pb,
                                  params_batch, kwargs)


# (engine id, engine parameter names, needs redshift/SED stage, default sed) -- used by likelihoods.py to
# route GaussianLikelihood.log_likelihood_batch through the fused kernel without an (N, E) intermediate.
FUSED = {
    arnett_bolometric: (_capi.ENGINE_NICKEL, ("f_nickel",), False, None),
    arnett: (_capi.ENGINE_NICKEL, ("f_nickel",), True, "Blackbody"),
    basic_magnetar_powered_bolometric: (_capi.ENGINE_MAGNETAR, _MAGNETAR, False, None),
    slsn_bolometric: (_capi.ENGINE_MAGNETAR, _MAGNETAR, False, None),
    basic_magnetar_powered: (_capi.ENGINE_MAGNETAR, _MAGNETAR, True, "Blackbody"),
    slsn: (_capi.ENGINE_MAGNETAR, _MAGNETAR, True, "CutoffBlackbody"),
}
