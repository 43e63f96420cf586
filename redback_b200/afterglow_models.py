"""Batched drop-ins for redback's native structured-jet afterglows
(redback/transient_models/afterglow_models/base_models.py: tophat_redback :885-946, gaussian_redback :948-1010),
output_format='flux_density' | 'magnitude'.  Same signatures as the reference plus `params_batch`."""
from . import _capi
from ._fused import FusedSpec, cosmology_from_kwargs, run
from .engine import AfterglowPath

_COMMON = ("redshift", "thv", "loge0", "thc", "logn0", "p", "logepse", "logepsb", "g0", "xiN")


def _spec(name, jet, needed, ss=0.01, aa=0.5):
    def validate(kwargs):
        # base_models.py:942-946: 'flux_density' (mJy) or 'magnitude' (monochromatic AB); anything else -> None there
        fmt = kwargs.get("output_format")
        if fmt is None:
            raise KeyError("output_format")
        if fmt not in ("flux_density", "magnitude"):
            raise NotImplementedError(f"{name}: output_format='{fmt}' is not defined for this model")

    def build(time, kwargs, y, sigma, sigma_mode):
        return AfterglowPath(jet, time, frequency=kwargs.get("frequency"), y=y, sigma=sigma, sigma_mode=sigma_mode,
                             steps=kwargs.get("steps", 250), res=kwargs.get("res"), k=kwargs.get("k", 0),
                             expansion=kwargs.get("expansion", 1), a1=kwargs.get("a1", 1),
                             cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"),
                             ss=kwargs.get("ss", ss), aa=kwargs.get("aa", aa),
                             ab_magnitude=kwargs["output_format"] == "magnitude", dtype=kwargs.get("dtype", "float64"))

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


_WITH_THJ = _COMMON[:4] + ("thj",) + _COMMON[4:]
# defaults of the extra structure parameters: base_models.py:1060-1061, 1129-1130, 1197-1198, 1265-1266
_TWO_COMPONENT = _spec("twocomponent_redback", _capi.AG_TWO_COMPONENT, _WITH_THJ, ss=0.01, aa=4)
_POWERLAW = _spec("powerlaw_redback", _capi.AG_POWERLAW, _WITH_THJ, ss=3, aa=-3)
_POWERLAW_ALT = _spec("alternativepowerlaw_redback", _capi.AG_POWERLAW_ALT, _WITH_THJ, ss=3, aa=3)
_DOUBLE_GAUSSIAN = _spec("doublegaussian_redback", _capi.AG_DOUBLE_GAUSSIAN, _WITH_THJ, ss=0.1, aa=0.5)


def _with_thj(spec, doc):
    def model(time, redshift=None, thv=None, loge0=None, thc=None, thj=None, logn0=None, p=None, logepse=None,
              logepsb=None, g0=None, xiN=None, params_batch=None, **kwargs):
        named = dict(redshift=redshift, thv=thv, loge0=loge0, thc=thc, thj=thj, logn0=logn0, p=p, logepse=logepse,
                     logepsb=logepsb, g0=g0, xiN=xiN)
        return run(spec, time, named, kwargs, params_batch)
    model.__name__ = model.__qualname__ = spec.name
    model.__doc__ = doc
    return model


twocomponent_redback = _with_thj(_TWO_COMPONENT, "redback afterglow_models.twocomponent_redback (:1013-1078); "
                                                 "kwargs ss (0.01), aa (4); returns mJy.")
powerlaw_redback = _with_thj(_POWERLAW, "redback afterglow_models.powerlaw_redback (:1081-1147); kwargs ss (3), "
                                        "aa (-3); returns mJy.")
alternativepowerlaw_redback = _with_thj(_POWERLAW_ALT, "redback afterglow_models.alternativepowerlaw_redback "
                                                       "(:1150-1215); kwargs ss (3), aa (3); returns mJy.")
doublegaussian_redback = _with_thj(_DOUBLE_GAUSSIAN, "redback afterglow_models.doublegaussian_redback (:1218-1283); "
                                                     "kwargs ss (0.1), aa (0.5); returns mJy.")

FUSED = {tophat_redback: _TOPHAT, gaussian_redback: _GAUSSIAN, twocomponent_redback: _TWO_COMPONENT,
         powerlaw_redback: _POWERLAW, alternativepowerlaw_redback: _POWERLAW_ALT,
         doublegaussian_redback: _DOUBLE_GAUSSIAN}
