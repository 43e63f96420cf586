"""Batched drop-ins for redback's analytic fallback TDE model (redback/transient_models/tde_models.py:
_analytic_fallback :16-27, tde_analytical_bolometric :682-701, tde_analytical :704-760): fallback luminosity through
Diffusion -> TemperatureFloor -> CutoffBlackbody, on the supernova kernel with its own engine instantiation."""
from . import _capi
from ._fused import run
from .supernova_models import _spec

_TDE_BOLO = _spec("tde_analytical_bolometric", _capi.ENGINE_TDE_FALLBACK, ("l0", "t_0_turn"), False, None)
_TDE = _spec("tde_analytical", _capi.ENGINE_TDE_FALLBACK, ("l0", "t_0_turn"), True, "CutoffBlackbody")


def tde_analytical_bolometric(time, l0=None, t_0_turn=None, params_batch=None, **kwargs):
    """redback tde_models.tde_analytical_bolometric (:682-701); returns erg/s."""
    return run(_TDE_BOLO, time, dict(l0=l0, t_0_turn=t_0_turn), kwargs, params_batch)


def tde_analytical(time, redshift=None, l0=None, t_0_turn=None, params_batch=None, **kwargs):
    """redback tde_models.tde_analytical (:704-760); default sed = CutoffBlackbody (3000 A); mJy / mag."""
    return run(_TDE, time, dict(redshift=redshift, l0=l0, t_0_turn=t_0_turn), kwargs, params_batch)


FUSED = {tde_analytical_bolometric: _TDE_BOLO, tde_analytical: _TDE}
