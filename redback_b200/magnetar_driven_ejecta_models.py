"""Batched drop-ins for redback's magnetar-driven relativistic-ejecta ("mergernova") models
(redback/transient_models/magnetar_driven_ejecta_models.py: basic_mergernova :275-321, general_mergernova :324-371,
general_mergernova_thermalisation :374-419), output_format='flux_density' | 'magnitude' | 'flux'.  Same signatures as the reference plus
`params_batch` (see supernova_models.py)."""
from . import _capi
from ._fused import FusedSpec, cosmology_from_kwargs, output_mode, run
from .engine import MergernovaMagPath, MergernovaPath, MetzgerMagnetarMagPath, MetzgerMagnetarPath

_EJECTA = ("redshift", "mej", "beta", "ejecta_radius", "kappa", "n_ism")


def _spec(name, engine, engine_params, last, pair_cascade_default, gamma_ray):
    needed = _EJECTA + tuple(engine_params) + (last,)

    def validate(kwargs):
        output_mode(kwargs, name)
        if not kwargs.get("use_r_process", True):
            raise NotImplementedError(f"{name}: use_r_process=False (nickel heating) is not fused on device")

    def build(time, kwargs, y, sigma, sigma_mode):
        mode = output_mode(kwargs, name)
        if mode != "flux_density":
            return MergernovaMagPath(engine, time, kwargs["bands"], output=mode,
                                     lambda_array=kwargs.get("lambda_array"), y=y, sigma=sigma, sigma_mode=sigma_mode,
                                     use_gamma_ray_opacity=gamma_ray,
                                     pair_cascade_switch=kwargs.get("pair_cascade_switch", pair_cascade_default),
                                     ejecta_albedo=kwargs.get("ejecta_albedo", 0.5),
                                     pair_cascade_fraction=kwargs.get("pair_cascade_fraction", 0.05),
                                     cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"))
        return MergernovaPath(engine, time, frequency=kwargs.get("frequency"), y=y, sigma=sigma, sigma_mode=sigma_mode,
                              use_gamma_ray_opacity=gamma_ray,
                              pair_cascade_switch=kwargs.get("pair_cascade_switch", pair_cascade_default),
                              ejecta_albedo=kwargs.get("ejecta_albedo", 0.5),
                              pair_cascade_fraction=kwargs.get("pair_cascade_fraction", 0.05),
                              cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"))

    return FusedSpec(name, needed, build, validate)


_BASIC = _spec("basic_mergernova", _capi.MN_BASIC_MAGNETAR, ("p0", "bp", "mass_ns", "theta_pb"),
               "thermalisation_efficiency", False, False)
_GENERAL = _spec("general_mergernova", _capi.MN_MAGNETAR_ONLY, ("l0", "tau_sd", "nn"),
                 "thermalisation_efficiency", True, False)
_GENERAL_TH = _spec("general_mergernova_thermalisation", _capi.MN_MAGNETAR_ONLY, ("l0", "tau_sd", "nn"),
                    "kappa_gamma", True, True)
_GENERAL_EV = _spec("general_mergernova_evolution", _capi.MN_EVOLVING,
                    ("logbint", "logbext", "p0", "chi0", "radius", "logmoi"), "kappa_gamma", True, True)


def basic_mergernova(time, redshift=None, mej=None, beta=None, ejecta_radius=None, kappa=None, n_ism=None, p0=None,
                     bp=None, mass_ns=None, theta_pb=None, thermalisation_efficiency=None, params_batch=None,
                     **kwargs):
    """redback magnetar_driven_ejecta_models.basic_mergernova (:275-321); kwargs pair_cascade_switch (False),
    ejecta_albedo, pair_cascade_fraction, frequency; returns mJy."""
    named = dict(redshift=redshift, mej=mej, beta=beta, ejecta_radius=ejecta_radius, kappa=kappa, n_ism=n_ism, p0=p0,
                 bp=bp, mass_ns=mass_ns, theta_pb=theta_pb, thermalisation_efficiency=thermalisation_efficiency)
    return run(_BASIC, time, named, kwargs, params_batch)


def general_mergernova(time, redshift=None, mej=None, beta=None, ejecta_radius=None, kappa=None, n_ism=None, l0=None,
                       tau_sd=None, nn=None, thermalisation_efficiency=None, params_batch=None, **kwargs):
    """redback magnetar_driven_ejecta_models.general_mergernova (:324-371); pair_cascade_switch defaults to True."""
    named = dict(redshift=redshift, mej=mej, beta=beta, ejecta_radius=ejecta_radius, kappa=kappa, n_ism=n_ism, l0=l0,
                 tau_sd=tau_sd, nn=nn, thermalisation_efficiency=thermalisation_efficiency)
    return run(_GENERAL, time, named, kwargs, params_batch)


def general_mergernova_thermalisation(time, redshift=None, mej=None, beta=None, ejecta_radius=None, kappa=None,
                                      n_ism=None, l0=None, tau_sd=None, nn=None, kappa_gamma=None, params_batch=None,
                                      **kwargs):
    """redback magnetar_driven_ejecta_models.general_mergernova_thermalisation (:374-419): thermalisation efficiency
    1 - exp(-3 kappa_gamma M / (4 pi v^2 t^2)) per step."""
    named = dict(redshift=redshift, mej=mej, beta=beta, ejecta_radius=ejecta_radius, kappa=kappa, n_ism=n_ism, l0=l0,
                 tau_sd=tau_sd, nn=nn, kappa_gamma=kappa_gamma)
    return run(_GENERAL_TH, time, named, kwargs, params_batch)


_SHELLS = ("redshift", "mej", "vej", "beta", "kappa_r")


def _metzger_spec(name, engine, engine_params, last, gamma_ray, resolution=300):
    needed = _SHELLS + tuple(engine_params) + (last,)

    def validate(kwargs):
        output_mode(kwargs, name)
        if kwargs.get("magnetar_heating", "first_layer") not in ("first_layer", "all_layers"):
            # the reference runs neither branch of :698 / :709 for any other value and returns a zero light curve
            raise ValueError(f"{name}: magnetar_heating must be 'first_layer' or 'all_layers'")

    def build(time, kwargs, y, sigma, sigma_mode):
        mode = output_mode(kwargs, name)
        if mode != "flux_density":
            return MetzgerMagnetarMagPath(engine, time, kwargs["bands"], output=mode,
                                          lambda_array=kwargs.get("lambda_array"), y=y, sigma=sigma,
                                          sigma_mode=sigma_mode, use_gamma_ray_opacity=gamma_ray,
                                          pair_cascade_switch=kwargs.get("pair_cascade_switch", True),
                                          ejecta_albedo=kwargs.get("ejecta_albedo", 0.5),
                                          pair_cascade_fraction=kwargs.get("pair_cascade_fraction", 0.01),
                                          neutron_precursor_switch=kwargs.get("neutron_precursor_switch", True),
                                          vmax=kwargs.get("vmax", 0.7), cosmology=cosmology_from_kwargs(kwargs),
                                          device=kwargs.get("device"), resolution=resolution,
                                          magnetar_heating=kwargs.get("magnetar_heating", "first_layer"))
        return MetzgerMagnetarPath(engine, time, frequency=kwargs.get("frequency"), y=y, sigma=sigma,
                                   sigma_mode=sigma_mode, use_gamma_ray_opacity=gamma_ray,
                                   pair_cascade_switch=kwargs.get("pair_cascade_switch", True),
                                   ejecta_albedo=kwargs.get("ejecta_albedo", 0.5),
                                   pair_cascade_fraction=kwargs.get("pair_cascade_fraction", 0.01),
                                   neutron_precursor_switch=kwargs.get("neutron_precursor_switch", True),
                                   vmax=kwargs.get("vmax", 0.7), cosmology=cosmology_from_kwargs(kwargs),
                                   device=kwargs.get("device"), resolution=resolution,
                                   magnetar_heating=kwargs.get("magnetar_heating", "first_layer"))

    return FusedSpec(name, needed, build, validate)


_EVOLVING = ("logbint", "logbext", "p0", "chi0", "radius", "logmoi")
_MK_BASIC = _metzger_spec("metzger_magnetar_driven_kilonova_model", _capi.MN_BASIC_MAGNETAR,
                          ("p0", "bp", "mass_ns", "theta_pb"), "thermalisation_efficiency", False)
_MK_GENERAL = _metzger_spec("general_metzger_magnetar_driven", _capi.MN_MAGNETAR_ONLY, ("l0", "tau_sd", "nn"),
                            "thermalisation_efficiency", False)
_MK_GENERAL_TH = _metzger_spec("general_metzger_magnetar_driven_thermalisation", _capi.MN_MAGNETAR_ONLY,
                               ("l0", "tau_sd", "nn"), "kappa_gamma", True)


_MK_EVOLUTION = _metzger_spec("general_metzger_magnetar_driven_evolution", _capi.MN_EVOLVING, _EVOLVING, "kappa_gamma",
                              True, resolution=500)


def general_metzger_magnetar_driven_evolution(time, redshift=None, mej=None, vej=None, beta=None, kappa_r=None,
                                              logbint=None, logbext=None, p0=None, chi0=None, radius=None, logmoi=None,
                                              kappa_gamma=None, params_batch=None, **kwargs):
    """redback magnetar_driven_ejecta_models.general_metzger_magnetar_driven_evolution (:908-956): the magnetar spins
    down under EM + GW torques (magnetar_models._evolving_gw_and_em_magnetar); p0 in seconds, radius in km."""
    named = dict(redshift=redshift, mej=mej, vej=vej, beta=beta, kappa_r=kappa_r, logbint=logbint, logbext=logbext,
                 p0=p0, chi0=chi0, radius=radius, logmoi=logmoi, kappa_gamma=kappa_gamma)
    return run(_MK_EVOLUTION, time, named, kwargs, params_batch)


def metzger_magnetar_driven_kilonova_model(time, redshift=None, mej=None, vej=None, beta=None, kappa_r=None, p0=None,
                                           bp=None, mass_ns=None, theta_pb=None, thermalisation_efficiency=None,
                                           params_batch=None, **kwargs):
    """redback magnetar_driven_ejecta_models.metzger_magnetar_driven_kilonova_model (:760-806): Metzger (2017) shell
    model with a basic_magnetar heating the innermost shell; kwargs pair_cascade_switch (True), ejecta_albedo,
    pair_cascade_fraction (0.01), neutron_precursor_switch, vmax; mJy."""
    named = dict(redshift=redshift, mej=mej, vej=vej, beta=beta, kappa_r=kappa_r, p0=p0, bp=bp, mass_ns=mass_ns,
                 theta_pb=theta_pb, thermalisation_efficiency=thermalisation_efficiency)
    return run(_MK_BASIC, time, named, kwargs, params_batch)


def general_metzger_magnetar_driven(time, redshift=None, mej=None, vej=None, beta=None, kappa_r=None, l0=None,
                                    tau_sd=None, nn=None, thermalisation_efficiency=None, params_batch=None, **kwargs):
    """redback magnetar_driven_ejecta_models.general_metzger_magnetar_driven (:809-856)."""
    named = dict(redshift=redshift, mej=mej, vej=vej, beta=beta, kappa_r=kappa_r, l0=l0, tau_sd=tau_sd, nn=nn,
                 thermalisation_efficiency=thermalisation_efficiency)
    return run(_MK_GENERAL, time, named, kwargs, params_batch)


def general_metzger_magnetar_driven_thermalisation(time, redshift=None, mej=None, vej=None, beta=None, kappa_r=None,
                                                   l0=None, tau_sd=None, nn=None, kappa_gamma=None, params_batch=None,
                                                   **kwargs):
    """redback magnetar_driven_ejecta_models.general_metzger_magnetar_driven_thermalisation (:859-905)."""
    named = dict(redshift=redshift, mej=mej, vej=vej, beta=beta, kappa_r=kappa_r, l0=l0, tau_sd=tau_sd, nn=nn,
                 kappa_gamma=kappa_gamma)
    return run(_MK_GENERAL_TH, time, named, kwargs, params_batch)


def general_mergernova_evolution(time, redshift=None, mej=None, beta=None, ejecta_radius=None, kappa=None, n_ism=None,
                                 logbint=None, logbext=None, p0=None, chi0=None, radius=None, logmoi=None,
                                 kappa_gamma=None, params_batch=None, **kwargs):
    """redback magnetar_driven_ejecta_models.general_mergernova_evolution (:422-474): evolving GW + EM magnetar."""
    named = dict(redshift=redshift, mej=mej, beta=beta, ejecta_radius=ejecta_radius, kappa=kappa, n_ism=n_ism,
                 logbint=logbint, logbext=logbext, p0=p0, chi0=chi0, radius=radius, logmoi=logmoi,
                 kappa_gamma=kappa_gamma)
    return run(_GENERAL_EV, time, named, kwargs, params_batch)


def _trapped_spec(output):
    needed = (("redshift",) if output == _capi.MN_OUT_TRAPPED_FLUX else ()) + _EJECTA[1:] + \
        ("l0", "tau_sd", "nn", "thermalisation_efficiency")

    def build(time, kwargs, y, sigma, sigma_mode):
        return MergernovaPath(_capi.MN_MAGNETAR_ONLY, time, frequency=kwargs.get("frequency"), y=y, sigma=sigma,
                              sigma_mode=sigma_mode, cosmology=cosmology_from_kwargs(kwargs),
                              device=kwargs.get("device"), output=output,
                              photon_index=kwargs.get("photon_index", 2.0) if output == _capi.MN_OUT_TRAPPED_FLUX
                              else 2.0)

    return FusedSpec("trapped_magnetar", needed, build)


_TRAPPED_LUM = _trapped_spec(_capi.MN_OUT_TRAPPED_LUMINOSITY)
_TRAPPED_FLUX = _trapped_spec(_capi.MN_OUT_TRAPPED_FLUX)


class _TrappedSpec:
    def for_kwargs(self, kwargs):
        fmt = kwargs["output_format"]                           # KeyError like the reference (:568)
        if fmt == "luminosity":
            return _TRAPPED_LUM
        if fmt == "flux":
            if "photon_index" not in kwargs:
                raise KeyError("photon_index")                  # positional argument of _trapped_magnetar_flux (:518)
            return _TRAPPED_FLUX
        raise ValueError("trapped_magnetar: output_format must be 'luminosity' or 'flux'")


_TRAPPED = _TrappedSpec()


def trapped_magnetar(time, redshift=None, mej=None, beta=None, ejecta_radius=None, kappa=None, n_ism=None, l0=None,
                     tau_sd=None, nn=None, thermalisation_efficiency=None, params_batch=None, **kwargs):
    """redback magnetar_driven_ejecta_models.trapped_magnetar (:549-573): `time` in SECONDS; output_format
    'luminosity' (source frame, redshift unused) or 'flux' (needs kwargs frequency, photon_index)."""
    named = dict(redshift=redshift, mej=mej, beta=beta, ejecta_radius=ejecta_radius, kappa=kappa, n_ism=n_ism, l0=l0,
                 tau_sd=tau_sd, nn=nn, thermalisation_efficiency=thermalisation_efficiency)
    return run(_TRAPPED.for_kwargs(kwargs), time, named, kwargs, params_batch)


FUSED = {trapped_magnetar: _TRAPPED, metzger_magnetar_driven_kilonova_model: _MK_BASIC,
         general_metzger_magnetar_driven: _MK_GENERAL, general_metzger_magnetar_driven_thermalisation: _MK_GENERAL_TH,
         general_metzger_magnetar_driven_evolution: _MK_EVOLUTION, general_mergernova_evolution: _GENERAL_EV,
         basic_mergernova: _BASIC, general_mergernova: _GENERAL, general_mergernova_thermalisation: _GENERAL_TH}
