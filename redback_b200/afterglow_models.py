"""Batched drop-ins for redback's native structured-jet afterglows
(redback/transient_models/afterglow_models/base_models.py: tophat_redback :885-946, gaussian_redback :948-1010),
output_format='flux_density'.  Same signatures as the reference plus `params_batch`."""
from . import _capi
from ._fused import FusedSpec, cosmology_from_kwargs, flux_density_only, run
from .engine import AfterglowPath

_COMMON = ("redshift", "thv", "loge0", "thc", "logn0", "p", "logepse", "logepsb", "g0", "xiN")


def _spec(name, jet, needed):
    def validate(kwargs):
        flux_density_only(kwargs, name)

    def build(time, kwargs, y, sigma, sigma_mode):
        return AfterglowPath(jet, time, frequency=kwargs.get("frequency"), y=y, sigma=sigma, sigma_mode=sigma_mode,
                             steps=kwargs.get("steps", 250), res=kwargs.get("res"), k=kwargs.get("k", 0),
                             expansion=kwargs.get("expansion", 1), a1=kwargs.get("a1", 1),
                             cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"))

    return FusedSpec(name, needed, build, validate)


_TOPHAT = _spec("tophat_redback", _capi.AG_TOPHAT, _COMMON)
_GAUSSIAN = _spec("gaussian_redback", _capi.AG_GAUSSIAN, _COMMON[:4] + ("thj",) + _COMMON[4:])


def tophat_redback(time, redshift=None, thv=None, loge0=None, thc=None, logn0=None, p=None, logepse=None,
                   logepsb=None, g0=None, xiN=None, params_batch=None, **kwargs):
    """redback afterglow_models.tophat_redback (:885-946); kwargs frequency, res, steps, k, a1, expansion;
    returns mJy."""
    named = dict(redshift=redshift, thv=thv, loge0=loge0, thc=thc, logn0=logn0, p=p, logepse=logepse,
                 logepsb=logepsb, g0=g0, xiN=xiN)
    return run(_TOPHAT, time, named, kwargs, params_batch)


def gaussian_redback(time, redshift=None, thv=None, loge0=None, thc=None, thj=None, logn0=None, p=None,
                     logepse=None, logepsb=None, g0=None, xiN=None, params_batch=None, **kwargs):
    """redback afterglow_models.gaussian_redback (:948-1010); returns mJy."""
    named = dict(redshift=redshift, thv=thv, loge0=loge0, thc=thc, thj=thj, logn0=logn0, p=p, logepse=logepse,
                 logepsb=logepsb, g0=g0, xiN=xiN)
    return run(_GAUSSIAN, time, named, kwargs, params_batch)


FUSED = {tophat_redback: _TOPHAT, gaussian_redback: _GAUSSIAN}
