"""Batched drop-ins for redback's one-component and Metzger kilonova models
(redback/transient_models/kilonova_models.py:1537-1603 and :1895-1962), output_format='flux_density'.
Same signatures as the reference plus `params_batch` (see supernova_models.py)."""
from . import _capi
from ._fused import FusedSpec, cosmology_from_kwargs, output_mode, run
from .engine import KilonovaMagPath, KilonovaPath


def _spec(name, model, needed):
    def validate(kwargs):
        output_mode(kwargs, name)

    def build(time, kwargs, y, sigma, sigma_mode):
        mode = output_mode(kwargs, name)
        if mode != "flux_density":
            return KilonovaMagPath(model, time, kwargs["bands"], output=mode, lambda_array=kwargs.get("lambda_array"),
                                   y=y, sigma=sigma, sigma_mode=sigma_mode,
                                   dense_resolution=kwargs.get("dense_resolution", 500),
                                   neutron_precursor_switch=kwargs.get("neutron_precursor_switch", True),
                                   vmax=kwargs.get("vmax", 0.7), cosmology=cosmology_from_kwargs(kwargs),
                                   device=kwargs.get("device"))
        return KilonovaPath(model, time, frequency=kwargs.get("frequency"), y=y, sigma=sigma, sigma_mode=sigma_mode,
                            dense_resolution=kwargs.get("dense_resolution", 500),
                            neutron_precursor_switch=kwargs.get("neutron_precursor_switch", True),
                            vmax=kwargs.get("vmax", 0.7), cosmology=cosmology_from_kwargs(kwargs),
                            device=kwargs.get("device"))

    return FusedSpec(name, needed, build, validate)


_ONE = _spec("one_component_kilonova_model", _capi.KN_ONE_COMPONENT,
             ("redshift", "mej", "vej", "kappa", "temperature_floor"))
_METZGER = _spec("metzger_kilonova_model", _capi.KN_METZGER, ("redshift", "mej", "vej", "beta", "kappa"))


def one_component_kilonova_model(time, redshift=None, mej=None, vej=None, kappa=None, params_batch=None, **kwargs):
    """redback kilonova_models.one_component_kilonova_model (:1537-1603); kwargs temperature_floor
    (default 4000 K, :1868), frequency, output_format='flux_density'; returns mJy."""
    kwargs.setdefault("temperature_floor", 4000)
    return run(_ONE, time, dict(redshift=redshift, mej=mej, vej=vej, kappa=kappa), kwargs, params_batch)


def metzger_kilonova_model(time, redshift=None, mej=None, vej=None, beta=None, kappa=None, params_batch=None,
                           **kwargs):
    """redback kilonova_models.metzger_kilonova_model (:1895-1962); kwargs neutron_precursor_switch, vmax,
    frequency, output_format='flux_density'; returns mJy."""
    return run(_METZGER, time, dict(redshift=redshift, mej=mej, vej=vej, beta=beta, kappa=kappa), kwargs,
               params_batch)


FUSED = {one_component_kilonova_model: _ONE, metzger_kilonova_model: _METZGER}
