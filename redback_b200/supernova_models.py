"""Batched drop-ins for redback's diffusion-powered supernova models.

Same names, positional parameters and kwargs as redback/transient_models/supernova_models.py
(arnett_bolometric :949-969, arnett :971-1028, basic_magnetar_powered_bolometric :1447-1469,
basic_magnetar_powered :1471-1530, slsn_bolometric :1532-1549, slsn :1552-1624) plus the new
`params_batch` entry point: a dict {name: array[N]} (numpy or torch) of sampled parameters.  Any model
parameter that is not in `params_batch` is taken from the call (positional / kwargs) as a scalar.
With `params_batch` the result has shape (N, E); without it the scalar call returns (E,), as in redback.
"""
from . import _capi
from ._fused import FusedSpec, cosmology_from_kwargs, output_mode, run
from .engine import SupernovaMagPath, SupernovaPath

_ALLOWED_OPERATORS = {
    "interaction_process": ("Diffusion", "CSMDiffusion", "AsphericalDiffusion", "Viscous"),
    "photosphere": ("TemperatureFloor",),
    "sed": ("Blackbody", "CutoffBlackbody"),
}
_SED_IDS = {"Blackbody": _capi.SED_BLACKBODY, "CutoffBlackbody": _capi.SED_CUTOFF_BLACKBODY}
_DIFFUSION = ("mej", "vej", "kappa", "kappa_gamma")
_MAGNETAR = ("p0", "bp", "mass_ns", "theta_pb")


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


_LINE_DEFAULTS = dict(line_wavelength=7.5e3, line_width=500, line_time=50, line_duration=25, line_amplitude=0.3)


def _line_kwargs(kwargs):
    """sed.Line parameters of type_1a (supernova_models.py:2255-2259): per-call scalars on device."""
    import numpy as np
    out = {}
    for key, default in _LINE_DEFAULTS.items():
        val = kwargs.get(key, default)
        if np.ndim(val) != 0:
            raise NotImplementedError(f"{key} must be a scalar in the fused B200 path")
        out[key] = float(val)
    return out


def _csm_kwargs(kwargs):
    """_csm_engine / CSMDiffusion keyword arguments (supernova_models.py:1934-1936, interaction_processes.py:154-155)."""
    method = kwargs.get("csm_diffusion_method", "cumulative")
    if method not in ("cumulative", "quadrature", "legacy"):
        raise ValueError("csm_diffusion_method must be 'cumulative' or 'quadrature'")          # interaction_processes.py:168
    if "dense_time_min" in kwargs or "stop_time" in kwargs:
        raise NotImplementedError("dense_time_min / stop_time overrides are not fused on device")
    timesteps = kwargs.get("csm_diffusion_timesteps", 3000)
    if method != "cumulative" and (int(timesteps) != timesteps or int(timesteps) % 2 or not 4 <= timesteps <= 6000):
        raise NotImplementedError("csm_diffusion_timesteps must be an even integer in [4, 6000] on device")
    return dict(nn=kwargs.get("nn", 12), delta=kwargs.get("delta", 1), efficiency=kwargs.get("efficiency", 0.5),
                quadrature=method != "cumulative", timesteps=int(timesteps))


def _area_ratio(kwargs):
    """interaction_process=AsphericalDiffusion (interaction_processes.py:66-131): area_projection / area_reference from
    the kwargs (per-call scalars), None for the plain Diffusion."""
    import numpy as np
    value = kwargs.get("interaction_process")
    name = value if isinstance(value, str) else getattr(value, "__name__", None)
    if name != "AsphericalDiffusion":
        return None
    proj, ref = kwargs["area_projection"], kwargs["area_reference"]      # KeyError like the reference's constructor
    if np.ndim(proj) != 0 or np.ndim(ref) != 0:
        raise NotImplementedError("area_projection / area_reference must be scalars in the fused B200 path")
    return float(proj) / float(ref)


_USES_MEJ = (_capi.ENGINE_NICKEL, _capi.ENGINE_MAGNETAR_NICKEL, _capi.ENGINE_FALLBACK_NICKEL)


def _ip_name(kwargs):
    value = kwargs.get("interaction_process")
    return value if isinstance(value, str) else getattr(value, "__name__", None)


def _spec(name, engine, engine_params, flux_density, default_sed, fixed_sed=None, direct=False, _variant=False,
          _viscous=False, csm_nickel=False):
    """direct: the model has no interaction process at all (sn_fallback); otherwise `interaction_process=None`
    selects the direct variant of the same spec (engine luminosity -> photosphere -> SED, :961-968), and
    `interaction_process=Viscous` (interaction_processes.py:229-273; kwarg / sampled parameter `t_viscous`) the viscous
    variant: the opacities are not needed then, `mej` only where the engine itself uses it."""
    is_csm = engine == _capi.ENGINE_CSM
    default_ip = "CSMDiffusion" if is_csm else "Diffusion"
    needed = list(engine_params) + list(_DIFFUSION)
    if is_csm:
        needed = list(engine_params)
    if direct or _variant or _viscous:
        needed = list(engine_params) + (["mej"] if engine in _USES_MEJ and "mej" not in engine_params else []) + \
            (["vej"] if flux_density or not _viscous else []) + (["t_viscous"] if _viscous else [])
    if flux_density:
        needed = ["redshift"] + needed + ["temperature_floor"]

    def validate(kwargs):
        if not (direct or _variant or _viscous):
            ip_name = _operator_name("interaction_process", kwargs.get("interaction_process", default_ip), default_ip)
            if ip_name == "AsphericalDiffusion" and not is_csm:
                _area_ratio(kwargs)
            elif ip_name != default_ip:
                raise NotImplementedError(f"{name}: only interaction_process={default_ip} is fused on device")
        if is_csm:
            _csm_kwargs(kwargs)
        if flux_density:
            output_mode(kwargs, name)
            _operator_name("photosphere", kwargs.get("photosphere"), "TemperatureFloor")
            if fixed_sed is None:
                _operator_name("sed", kwargs.get("sed"), default_sed)

    def build(time, kwargs, y, sigma, sigma_mode):
        sed_id = _capi.SED_BOLOMETRIC
        line = None
        csm = dict(_csm_kwargs(kwargs), nickel=csm_nickel) if is_csm else None
        no_ip = direct or _variant
        area_ratio = None if no_ip or is_csm or _viscous else _area_ratio(kwargs)
        if flux_density:
            if fixed_sed is not None:     # type_1a hard-wires CutoffBlackbody + Line (supernova_models.py:2268-2275)
                sed_id, line = fixed_sed, _line_kwargs(kwargs)
            else:
                sed_id = _SED_IDS[_operator_name("sed", kwargs.get("sed"), default_sed)]
            mode = output_mode(kwargs, name)
            if mode != "flux_density":
                return SupernovaMagPath(engine, sed_id, time, kwargs["bands"], output=mode,
                                        lambda_array=kwargs.get("lambda_array"), y=y, sigma=sigma,
                                        sigma_mode=sigma_mode, dense_resolution=kwargs.get("dense_resolution", 1000),
                                        cutoff_wavelength=kwargs.get("cutoff_wavelength", 3000),
                                        absorption_index=kwargs.get("absorption_index", 1.0),
                                        cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"),
                                        line=line, csm=csm, no_interaction=no_ip, area_ratio=area_ratio,
                                        viscous=_viscous)
        return SupernovaPath(engine, sed_id, time, frequency=kwargs.get("frequency") if flux_density else None,
                             y=y, sigma=sigma, sigma_mode=sigma_mode,
                             dense_resolution=kwargs.get("dense_resolution", 1000),
                             cutoff_wavelength=kwargs.get("cutoff_wavelength", 3000),
                             absorption_index=kwargs.get("absorption_index", 1.0),
                             cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"), line=line,
                             dtype=kwargs.get("dtype", "float64"), csm=csm, no_interaction=no_ip,
                             area_ratio=area_ratio, viscous=_viscous)

    resolver = None
    if not (direct or _variant or _viscous or is_csm):
        variant = _spec(name + "[no interaction]", engine, engine_params, flux_density, default_sed, fixed_sed,
                        _variant=True)
        viscous = _spec(name + "[Viscous]", engine, engine_params, flux_density, default_sed, fixed_sed, _viscous=True)

        def resolver(kwargs):
            if "interaction_process" in kwargs and kwargs["interaction_process"] is None:
                return variant
            return viscous if _ip_name(kwargs) == "Viscous" else None
    return FusedSpec(name, needed, build, validate, resolver=resolver)


_ARNETT_BOLO = _spec("arnett_bolometric", _capi.ENGINE_NICKEL, ("f_nickel",), False, None)
_ARNETT = _spec("arnett", _capi.ENGINE_NICKEL, ("f_nickel",), True, "Blackbody")
_MAGNETAR_BOLO = _spec("basic_magnetar_powered_bolometric", _capi.ENGINE_MAGNETAR, _MAGNETAR, False, None)
_MAGNETAR_BB = _spec("basic_magnetar_powered", _capi.ENGINE_MAGNETAR, _MAGNETAR, True, "Blackbody")
_SLSN = _spec("slsn", _capi.ENGINE_MAGNETAR, _MAGNETAR, True, "CutoffBlackbody")
_MAGNETAR_NICKEL = _spec("magnetar_nickel", _capi.ENGINE_MAGNETAR_NICKEL, ("f_nickel",) + _MAGNETAR, True, "Blackbody")
_TYPE_1A = _spec("type_1a", _capi.ENGINE_NICKEL, ("f_nickel",), True, None, fixed_sed=_capi.SED_CUTOFF_LINE)


_CSM = ("mej", "csm_mass", "vej", "eta", "rho", "kappa", "r0")
_CSM_BOLO = _spec("csm_interaction_bolometric", _capi.ENGINE_CSM, _CSM, False, None)
_CSM_BB = _spec("csm_interaction", _capi.ENGINE_CSM, _CSM, True, "Blackbody")


def csm_interaction_bolometric(time, mej=None, csm_mass=None, vej=None, eta=None, rho=None, kappa=None, r0=None,
                               params_batch=None, **kwargs):
    """redback supernova_models.csm_interaction_bolometric (:2001-2050): _csm_engine on the adaptive dense grid +
    CSMDiffusion (cumulative); kwargs nn, delta, efficiency, dense_resolution; returns erg/s."""
    named = dict(mej=mej, csm_mass=csm_mass, vej=vej, eta=eta, rho=rho, kappa=kappa, r0=r0)
    return run(_CSM_BOLO, time, named, kwargs, params_batch)


def csm_interaction(time, redshift=None, mej=None, csm_mass=None, vej=None, eta=None, rho=None, kappa=None, r0=None,
                    params_batch=None, **kwargs):
    """redback supernova_models.csm_interaction (:2054-2111); TemperatureFloor + Blackbody by default; mJy / mag."""
    named = dict(redshift=redshift, mej=mej, csm_mass=csm_mass, vej=vej, eta=eta, rho=rho, kappa=kappa, r0=r0)
    return run(_CSM_BB, time, named, kwargs, params_batch)


# csm_nickel (supernova_models.py:2120-2232): the CSM shock + the nickel-cobalt engine diffused through the ejecta and the
# optically thick CSM; vej from the kinetic energy, sqrt(2 ek / mej) (:2140); kappa_gamma feeds Diffusion
_CSM_NI = ("mej", "f_nickel", "csm_mass", "vej", "eta", "rho", "kappa", "r0", "kappa_gamma")
_CSM_NI_BOLO = _spec("csm_nickel_bolometric", _capi.ENGINE_CSM, _CSM_NI, False, None, csm_nickel=True)
_CSM_NI_BB = _spec("csm_nickel", _capi.ENGINE_CSM, _CSM_NI, True, "Blackbody", csm_nickel=True)


def _with_ek(spec):
    needed = tuple(n for n in spec.needed if n != "vej") + ("ek",)
    return FusedSpec(spec.name, needed, spec._build, spec._validate, transform=lambda p: _velocity_from_ek((2.0, 1.0))(p))


def csm_nickel_bolometric(time, mej=None, f_nickel=None, csm_mass=None, ek=None, eta=None, rho=None, kappa=None, r0=None,
                          params_batch=None, **kwargs):
    """redback supernova_models.csm_nickel_bolometric (:2120-2165); kwargs kappa_gamma, csm_diffusion_method, nn, delta,
    efficiency, dense_resolution; returns erg/s."""
    named = dict(mej=mej, f_nickel=f_nickel, csm_mass=csm_mass, ek=ek, eta=eta, rho=rho, kappa=kappa, r0=r0)
    return run(_CSM_NI_BOLO_EK, time, named, kwargs, params_batch)


def csm_nickel(time, redshift=None, mej=None, f_nickel=None, csm_mass=None, ek=None, eta=None, rho=None, kappa=None,
               r0=None, params_batch=None, **kwargs):
    """redback supernova_models.csm_nickel (:2168-2232); TemperatureFloor + Blackbody; mJy / mag."""
    named = dict(redshift=redshift, mej=mej, f_nickel=f_nickel, csm_mass=csm_mass, ek=ek, eta=eta, rho=rho, kappa=kappa,
                 r0=r0)
    return run(_CSM_NI_BB_EK, time, named, kwargs, params_batch)


def arnett_bolometric(time, f_nickel=None, mej=None, params_batch=None, **kwargs):
    """redback supernova_models.arnett_bolometric (:949-969); returns erg/s."""
    return run(_ARNETT_BOLO, time, dict(f_nickel=f_nickel, mej=mej), kwargs, params_batch)


def arnett(time, redshift=None, f_nickel=None, mej=None, params_batch=None, **kwargs):
    """redback supernova_models.arnett (:971-1028), output_format='flux_density'; returns mJy."""
    return run(_ARNETT, time, dict(redshift=redshift, f_nickel=f_nickel, mej=mej), kwargs, params_batch)


def type_1c(time, redshift=None, f_nickel=None, mej=None, pp=None, params_batch=None, **kwargs):
    """redback supernova_models.type_1c (:2317-2352), output_format='flux_density': arnett with the Blackbody SED plus
    the epoch-independent Synchrotron SED (sed.py:703-751; kwargs nu_max = 1e9 Hz, f0 = 1e-26, source radius 1e13 cm
    as in the reference's call).  The synchrotron term is a closed form per (sample, epoch) added in place to the fused arnett
    output by `rb_synchrotron_add_f64`; the likelihood of this model goes through the stand-alone reduction."""
    import numpy as np
    import torch
    from . import engine
    name = "type_1c"
    if output_mode(kwargs, name, photometric=False) != "flux_density":
        raise NotImplementedError(f"{name}: output_format='{kwargs['output_format']}' is not fused on device")
    if kwargs.get("sed") is not None and _operator_name("sed", kwargs.get("sed"), "Blackbody") != "Blackbody":
        raise NotImplementedError(f"{name}: the reference always uses sed.Blackbody + sed.Synchrotron")
    batch = None if params_batch is None else {k: v for k, v in params_batch.items() if k != "pp"}
    want_device = kwargs.get("device_output", False)
    inner = dict(kwargs, device_output=True, sed=None)
    fd = arnett(time, redshift, f_nickel, mej, params_batch=batch, **inner)          # [E] or [N, E], mJy
    dev = fd.device
    as_col = (lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64) if not isinstance(v, torch.Tensor) else v,
                                        dtype=torch.float64, device=dev).reshape(-1, 1))
    src = params_batch if params_batch is not None else dict(redshift=redshift, pp=pp)
    if src.get("pp") is None or src.get("redshift") is None:
        raise TypeError(f"{name}() missing required parameter 'pp' or 'redshift'")
    z, p = as_col(src["redshift"]), as_col(src["pp"])
    dl = kwargs.get("dl")
    dl = engine.luminosity_distance(z.reshape(-1), cosmology_from_kwargs(kwargs), device=dev).reshape(-1, 1) \
        if dl is None else as_col(dl)
    out = fd.reshape(-1, fd.shape[-1]).contiguous()
    n, e = out.shape
    nu = torch.as_tensor(np.broadcast_to(np.asarray(kwargs["frequency"], dtype=np.float64), (e,)).copy(), device=dev)
    col = lambda v: v.reshape(-1).expand(n).contiguous()
    z, p, dl = col(z), col(p), col(dl)
    nu_max, f0, radius = float(kwargs.get("nu_max", 1e9)), float(kwargs.get("f0", 1e-26)), 1e13
    with torch.cuda.device(dev):                                                      # sed.py:703-751, :2343-2352
        _capi.check(_capi.lib().rb_synchrotron_add_f64(
            n, e, nu.data_ptr(), z.data_ptr(), p.data_ptr(), dl.data_ptr(), nu_max, f0, radius, out.data_ptr(),
            _capi.C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    out = out if params_batch is not None else out[0]
    return out if want_device else out.cpu().numpy()


def basic_magnetar_powered_bolometric(time, p0=None, bp=None, mass_ns=None, theta_pb=None, params_batch=None,
                                      **kwargs):
    """redback supernova_models.basic_magnetar_powered_bolometric (:1447-1469); returns erg/s."""
    return run(_MAGNETAR_BOLO, time, dict(p0=p0, bp=bp, mass_ns=mass_ns, theta_pb=theta_pb), kwargs, params_batch)


def slsn_bolometric(time, p0=None, bp=None, mass_ns=None, theta_pb=None, params_batch=None, **kwargs):
    """redback supernova_models.slsn_bolometric (:1532-1549)."""
    return run(_MAGNETAR_BOLO, time, dict(p0=p0, bp=bp, mass_ns=mass_ns, theta_pb=theta_pb), kwargs, params_batch)


def basic_magnetar_powered(time, redshift=None, p0=None, bp=None, mass_ns=None, theta_pb=None,
                           params_batch=None, **kwargs):
    """redback supernova_models.basic_magnetar_powered (:1471-1530); default sed = Blackbody; mJy."""
    return run(_MAGNETAR_BB, time, dict(redshift=redshift, p0=p0, bp=bp, mass_ns=mass_ns, theta_pb=theta_pb),
               kwargs, params_batch)


def slsn(time, redshift=None, p0=None, bp=None, mass_ns=None, theta_pb=None, params_batch=None, **kwargs):
    """redback supernova_models.slsn (:1552-1624); default sed = CutoffBlackbody (3000 A, index 1); mJy."""
    return run(_SLSN, time, dict(redshift=redshift, p0=p0, bp=bp, mass_ns=mass_ns, theta_pb=theta_pb),
               kwargs, params_batch)


_MAGNETAR_ONLY_BOLO = _spec("general_magnetar_slsn_bolometric", _capi.ENGINE_MAGNETAR_ONLY, ("l0", "tsd", "nn"), False, None)
_MAGNETAR_ONLY = _spec("general_magnetar_slsn", _capi.ENGINE_MAGNETAR_ONLY, ("l0", "tsd", "nn"), True, "Blackbody")
_EXP_PL = ("lbol_0", "alpha_1", "alpha_2", "tpeak_d")
_EXP_PL_BOLO = _spec("exponential_powerlaw_bolometric", _capi.ENGINE_EXP_POWERLAW, _EXP_PL, False, None)
_EXP_PL_BB = _spec("sn_exponential_powerlaw", _capi.ENGINE_EXP_POWERLAW, _EXP_PL, True, "Blackbody")


def general_magnetar_slsn_bolometric(time, l0=None, tsd=None, nn=None, params_batch=None, **kwargs):
    """redback supernova_models.general_magnetar_slsn_bolometric (:2379-2401): magnetar_only engine + Diffusion."""
    return run(_MAGNETAR_ONLY_BOLO, time, dict(l0=l0, tsd=tsd, nn=nn), kwargs, params_batch)


def general_magnetar_slsn(time, redshift=None, l0=None, tsd=None, nn=None, params_batch=None, **kwargs):
    """redback supernova_models.general_magnetar_slsn (:2404-2460); default sed = Blackbody; mJy / mag."""
    return run(_MAGNETAR_ONLY, time, dict(redshift=redshift, l0=l0, tsd=tsd, nn=nn), kwargs, params_batch)


def exponential_powerlaw_bolometric(time, lbol_0=None, alpha_1=None, alpha_2=None, tpeak_d=None, params_batch=None,
                                    **kwargs):
    """redback supernova_models.exponential_powerlaw_bolometric (:357-381)."""
    named = dict(lbol_0=lbol_0, alpha_1=alpha_1, alpha_2=alpha_2, tpeak_d=tpeak_d)
    return run(_EXP_PL_BOLO, time, named, kwargs, params_batch)


def sn_exponential_powerlaw(time, redshift=None, lbol_0=None, alpha_1=None, alpha_2=None, tpeak_d=None,
                            params_batch=None, **kwargs):
    """redback supernova_models.sn_exponential_powerlaw (:501-560); default sed = Blackbody."""
    named = dict(redshift=redshift, lbol_0=lbol_0, alpha_1=alpha_1, alpha_2=alpha_2, tpeak_d=tpeak_d)
    return run(_EXP_PL_BB, time, named, kwargs, params_batch)


_FALLBACK = _spec("sn_fallback", _capi.ENGINE_FALLBACK, ("logl1", "tr"), True, "Blackbody", direct=True)
_FALLBACK_NI = _spec("sn_nickel_fallback", _capi.ENGINE_FALLBACK_NICKEL, ("mej", "f_nickel", "logl1", "tr"), True,
                     "Blackbody", direct=True)


def sn_fallback(time, redshift=None, logl1=None, tr=None, params_batch=None, **kwargs):
    """redback supernova_models.sn_fallback (:383-438): fallback luminosity l1 t^(-5/3) (flat before tr) ->
    TemperatureFloor -> Blackbody, no interaction process; kwargs vej, temperature_floor."""
    return run(_FALLBACK, time, dict(redshift=redshift, logl1=logl1, tr=tr), kwargs, params_batch)


def sn_nickel_fallback(time, redshift=None, mej=None, f_nickel=None, logl1=None, tr=None, params_batch=None, **kwargs):
    """redback supernova_models.sn_nickel_fallback (:441-498): fallback + nickel-cobalt decay, no interaction process."""
    named = dict(redshift=redshift, mej=mej, f_nickel=f_nickel, logl1=logl1, tr=tr)
    return run(_FALLBACK_NI, time, named, kwargs, params_batch)


def magnetar_nickel(time, redshift=None, f_nickel=None, mej=None, p0=None, bp=None, mass_ns=None, theta_pb=None,
                    params_batch=None, **kwargs):
    """redback supernova_models.magnetar_nickel (:1628-1711): nickel-cobalt decay + magnetar spin-down through one
    Diffusion; default sed = Blackbody; mJy / mag."""
    named = dict(redshift=redshift, f_nickel=f_nickel, mej=mej, p0=p0, bp=bp, mass_ns=mass_ns, theta_pb=theta_pb)
    return run(_MAGNETAR_NICKEL, time, named, kwargs, params_batch)


def type_1a(time, redshift=None, f_nickel=None, mej=None, params_batch=None, **kwargs):
    """redback supernova_models.type_1a (:2229-2314): nickel engine, CutoffBlackbody + absorption/emission
    Line; kwargs line_wavelength, line_width, line_time, line_duration, line_amplitude (scalars); mJy / mag."""
    return run(_TYPE_1A, time, dict(redshift=redshift, f_nickel=f_nickel, mej=mej), kwargs, params_batch)


# model function -> FusedSpec: lets GaussianLikelihood.log_likelihood_batch route through the fused kernel
# without materialising the (N, E) model matrix.
_CSM_NI_BOLO_EK, _CSM_NI_BB_EK = None, None      # bound below, once _velocity_from_ek exists

FUSED = {
    arnett_bolometric: _ARNETT_BOLO, arnett: _ARNETT,
    basic_magnetar_powered_bolometric: _MAGNETAR_BOLO, slsn_bolometric: _MAGNETAR_BOLO,
    basic_magnetar_powered: _MAGNETAR_BB, slsn: _SLSN, type_1a: _TYPE_1A,
    csm_interaction_bolometric: _CSM_BOLO, csm_interaction: _CSM_BB, magnetar_nickel: _MAGNETAR_NICKEL,
    general_magnetar_slsn_bolometric: _MAGNETAR_ONLY_BOLO, general_magnetar_slsn: _MAGNETAR_ONLY,
    exponential_powerlaw_bolometric: _EXP_PL_BOLO, sn_exponential_powerlaw: _EXP_PL_BB,
    sn_fallback: _FALLBACK, sn_nickel_fallback: _FALLBACK_NI,
}


# ---- kinetic-energy parameterisations (supernova_models.py:1708-1905): vej derived from ek and mej ---------------
_SOLAR_MASS, _KM = 1.988409870698051e33, 100000.0
_EK_BASES = {"arnett_bolometric": (arnett_bolometric, arnett),
             "basic_magnetar_powered_bolometric": (basic_magnetar_powered_bolometric, basic_magnetar_powered),
             "slsn_bolometric": (slsn_bolometric, basic_magnetar_powered),    # the wrapper's default sed is Blackbody
             "general_magnetar_slsn_bolometric": (general_magnetar_slsn_bolometric, general_magnetar_slsn),
             "exponential_powerlaw_bolometric": (exponential_powerlaw_bolometric, sn_exponential_powerlaw)}


def _velocity_from_ek(factor):
    def transform(params):
        import numpy as np
        import torch
        params = dict(params)
        ek, mej = params.pop("ek"), params["mej"]
        if isinstance(ek, torch.Tensor) or isinstance(mej, torch.Tensor):
            ek, mej = torch.as_tensor(ek, dtype=torch.float64), torch.as_tensor(mej, dtype=torch.float64)
            params["vej"] = torch.sqrt(factor[0] * ek / (factor[1] * mej * _SOLAR_MASS)) / _KM
        else:
            params["vej"] = np.sqrt(factor[0] * np.asarray(ek, dtype=np.float64)
                                    / (factor[1] * np.asarray(mej, dtype=np.float64) * _SOLAR_MASS)) / _KM
        return params
    return transform


_CSM_NI_BOLO_EK, _CSM_NI_BB_EK = _with_ek(_CSM_NI_BOLO), _with_ek(_CSM_NI_BB)
FUSED[csm_nickel_bolometric] = _CSM_NI_BOLO_EK
FUSED[csm_nickel] = _CSM_NI_BB_EK


class _EkSpec:
    """FusedSpec of the base model with `vej` replaced by `ek` (resolved from kwargs['base_model'])."""

    def __init__(self, name, factor, full):
        self.name, self.factor, self.full = name, factor, full

    def for_kwargs(self, kwargs):
        base_model = kwargs.get("base_model", "arnett_bolometric")
        key = base_model if isinstance(base_model, str) else getattr(base_model, "__name__", None)
        if key not in _EK_BASES:
            raise ValueError("Please choose a different base model")     # supernova_models.py:1724-1726
        base = FUSED[_EK_BASES[key][1 if self.full else 0]]
        needed = tuple(n for n in base.needed if n != "vej") + ("ek",)

        def strip(kw):
            return {k: v for k, v in kw.items() if k != "base_model"}
        return FusedSpec(f"{self.name}:{key}", needed,
                         lambda time, kw, y, sigma, sigma_mode: base.build(time, strip(kw), y=y, sigma=sigma,
                                                                           sigma_mode=sigma_mode),
                         lambda kw: base.validate(strip(kw)), _velocity_from_ek(self.factor))


_HOMOLOGOUS_BOLO = _EkSpec("homologous_expansion_supernova_model_bolometric", (10.0, 3.0), False)
_HOMOLOGOUS = _EkSpec("homologous_expansion_supernova", (10.0, 3.0), True)
_THIN_SHELL_BOLO = _EkSpec("thin_shell_supernova_model_bolometric", (2.0, 1.0), False)
_THIN_SHELL = _EkSpec("thin_shell_supernova", (2.0, 1.0), True)


def homologous_expansion_supernova_model_bolometric(time, mej=None, ek=None, params_batch=None, **kwargs):
    """redback supernova_models.homologous_expansion_supernova_model_bolometric (:1708-1741):
    vej = sqrt(10 ek / (3 mej)); kwargs base_model (default 'arnett_bolometric') + its parameters."""
    return run(_HOMOLOGOUS_BOLO.for_kwargs(kwargs), time, dict(mej=mej, ek=ek), kwargs, params_batch)


def thin_shell_supernova_model_bolometric(time, mej=None, ek=None, params_batch=None, **kwargs):
    """redback supernova_models.thin_shell_supernova_model_bolometric (:1744-1779): vej = sqrt(2 ek / mej)."""
    return run(_THIN_SHELL_BOLO.for_kwargs(kwargs), time, dict(mej=mej, ek=ek), kwargs, params_batch)


def homologous_expansion_supernova(time, redshift=None, mej=None, ek=None, params_batch=None, **kwargs):
    """redback supernova_models.homologous_expansion_supernova (:1782-1844); default sed = Blackbody."""
    return run(_HOMOLOGOUS.for_kwargs(kwargs), time, dict(redshift=redshift, mej=mej, ek=ek), kwargs, params_batch)


def thin_shell_supernova(time, redshift=None, mej=None, ek=None, params_batch=None, **kwargs):
    """redback supernova_models.thin_shell_supernova (:1847-1905); default sed = Blackbody."""
    return run(_THIN_SHELL.for_kwargs(kwargs), time, dict(redshift=redshift, mej=mej, ek=ek), kwargs, params_batch)


FUSED.update({homologous_expansion_supernova_model_bolometric: _HOMOLOGOUS_BOLO,
              homologous_expansion_supernova: _HOMOLOGOUS,
              thin_shell_supernova_model_bolometric: _THIN_SHELL_BOLO, thin_shell_supernova: _THIN_SHELL})


# ---- general_magnetar_driven_supernova (supernova_models.py:2464-2638): relativistic ejecta dynamics -------------------
def _gmds_spec(name, flux_density):
    from .engine import MergernovaPath
    needed = (("redshift",) if flux_density else ()) + ("mej", "E_sn", "kappa", "l0", "tau_sd", "nn", "kappa_gamma",
                                                         "f_nickel") + (("temperature_floor",) if flux_density else ())

    def validate(kwargs):
        if flux_density:
            fmt = kwargs.get("output_format")
            if fmt is None:
                raise KeyError("output_format")
            if fmt == "dynamics_output":
                raise NotImplementedError(f"{name}: output_format='dynamics_output' returns Python objects in the reference")
            output_mode(kwargs, name)
            _operator_name("photosphere", kwargs.get("photosphere"), "TemperatureFloor")
            _operator_name("sed", kwargs.get("sed"), "CutoffBlackbody")
        elif kwargs.get("output_format") == "dynamics_output":
            raise NotImplementedError(f"{name}: output_format='dynamics_output' returns Python objects in the reference")

    def build(time, kwargs, y, sigma, sigma_mode):
        sn = dict(sed=_SED_IDS[_operator_name("sed", kwargs.get("sed"), "CutoffBlackbody")] if flux_density
                  else _capi.SED_BLACKBODY,
                  cutoff_wavelength=kwargs.get("cutoff_wavelength", 3000), absorption_index=kwargs.get("absorption_index", 1.0))
        mode = output_mode(kwargs, name) if flux_density else "flux_density"
        if mode != "flux_density":                                               # :2585-2637
            from .engine import MergernovaMagPath
            return MergernovaMagPath(_capi.MN_MAGNETAR_ONLY, time, kwargs["bands"], output=mode,
                                     lambda_array=kwargs.get("lambda_array"), y=y, sigma=sigma, sigma_mode=sigma_mode,
                                     use_gamma_ray_opacity=True,
                                     pair_cascade_switch=kwargs.get("pair_cascade_switch", False),
                                     ejecta_albedo=kwargs.get("ejecta_albedo", 0.5),
                                     pair_cascade_fraction=kwargs.get("pair_cascade_fraction", 0.05),
                                     cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"), supernova=sn)
        return MergernovaPath(_capi.MN_MAGNETAR_ONLY, time, frequency=kwargs.get("frequency") if flux_density else None,
                              y=y, sigma=sigma, sigma_mode=sigma_mode, use_gamma_ray_opacity=True,
                              pair_cascade_switch=kwargs.get("pair_cascade_switch", False),
                              ejecta_albedo=kwargs.get("ejecta_albedo", 0.5),
                              pair_cascade_fraction=kwargs.get("pair_cascade_fraction", 0.05),
                              cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"),
                              output=_capi.MN_OUT_SUPERNOVA if flux_density else _capi.MN_OUT_LBOL, supernova=sn)

    def transform(params):
        import numpy as np
        import torch
        params = dict(params)
        e_sn, mej = params.pop("E_sn"), params["mej"]
        if isinstance(e_sn, torch.Tensor) or isinstance(mej, torch.Tensor):
            e_sn, mej = torch.as_tensor(e_sn, dtype=torch.float64), torch.as_tensor(mej, dtype=torch.float64)
            params["beta"] = torch.sqrt(e_sn / (0.5 * mej * _SOLAR_MASS)) / 29979245800.0
        else:
            params["beta"] = np.sqrt(np.asarray(e_sn, dtype=np.float64)
                                     / (0.5 * np.asarray(mej, dtype=np.float64) * _SOLAR_MASS)) / 29979245800.0   # :2478
        params["ejecta_radius"] = 1.0e11                                                                            # :2479
        params["n_ism"] = 1.0e-5
        return params

    return FusedSpec(name, needed, build, validate, transform, defaults={"f_nickel": 0})   # :82 of the dynamics


_GMDS_BOLO = _gmds_spec("general_magnetar_driven_supernova_bolometric", False)
_GMDS = _gmds_spec("general_magnetar_driven_supernova", True)


def general_magnetar_driven_supernova_bolometric(time, mej=None, E_sn=None, kappa=None, l0=None, tau_sd=None, nn=None,
                                                 kappa_gamma=None, params_batch=None, **kwargs):
    """redback supernova_models.general_magnetar_driven_supernova_bolometric (:2464-2507): relativistic ejecta dynamics
    (nickel heating, gamma-ray leakage) driven by a magnetar; kwargs f_nickel, pair_cascade_switch; erg/s."""
    named = dict(mej=mej, E_sn=E_sn, kappa=kappa, l0=l0, tau_sd=tau_sd, nn=nn, kappa_gamma=kappa_gamma)
    return run(_GMDS_BOLO, time, named, kwargs, params_batch)


def general_magnetar_driven_supernova(time, redshift=None, mej=None, E_sn=None, kappa=None, l0=None, tau_sd=None,
                                      nn=None, kappa_gamma=None, params_batch=None, **kwargs):
    """redback supernova_models.general_magnetar_driven_supernova (:2510-2637), flux_density and magnitude / flux:
    TemperatureFloor on the dynamics grid with the time-dependent ejecta velocity, CutoffBlackbody by default."""
    named = dict(redshift=redshift, mej=mej, E_sn=E_sn, kappa=kappa, l0=l0, tau_sd=tau_sd, nn=nn,
                 kappa_gamma=kappa_gamma)
    return run(_GMDS, time, named, kwargs, params_batch)


FUSED.update({general_magnetar_driven_supernova_bolometric: _GMDS_BOLO, general_magnetar_driven_supernova: _GMDS})
