"""Batched drop-in for redback.transient_models.phase_models.t0_base_model (phase_models.py:19-59): the explosion
epoch `t0` (MJD) is a sampled parameter, `time` is in MJD, epochs before t0 return 0 (1000 for magnitudes).

On device the shift is a per-sample offset inside the fused kernels (`rb_sn_config.t0`, `rb_kn_config.t0`), so a batch
with N different explosion epochs is still one launch.  Supported base models: the diffusion-powered supernova family
(arnett, basic_magnetar_powered, slsn, type_1a, magnetar_nickel and their bolometric variants) the one-component
and Metzger kilonova models, the mergernova and the magnetar-driven Metzger shell models.
"""
import numpy as np

from . import kilonova_models as _kn
from . import magnetar_driven_ejecta_models as _mn
from . import supernova_models as _sn
from ._fused import FusedSpec, run

_BASE = {name: getattr(_sn, name) for name in ("arnett_bolometric", "arnett", "basic_magnetar_powered_bolometric",
                                               "basic_magnetar_powered", "slsn_bolometric", "slsn", "type_1a",
                                               "magnetar_nickel")}
_SPECS = {fn: _sn.FUSED[fn] for fn in _BASE.values()}
# the one-component and Metzger kilonova models (rb_kn_config.t0): each sample's adaptive grid follows its valid epochs
for _name in ("one_component_kilonova_model", "metzger_kilonova_model"):
    _BASE[_name] = getattr(_kn, _name)
    _SPECS[_BASE[_name]] = _kn.FUSED[_BASE[_name]]
# the mergernova and magnetar-driven Metzger shell models (rb_mn_config.t0, rb_mk_config.t0; fixed model grids)
for _name in ("basic_mergernova", "general_mergernova", "general_mergernova_thermalisation",
              "general_mergernova_evolution", "metzger_magnetar_driven_kilonova_model",
              "general_metzger_magnetar_driven", "general_metzger_magnetar_driven_thermalisation",
              "general_metzger_magnetar_driven_evolution"):
    _BASE[_name] = getattr(_mn, _name)
    _SPECS[_BASE[_name]] = _mn.FUSED[_BASE[_name]]


def _check_sorted(time):
    # the reference returns concatenate((fake, real)) (phase_models.py:58): the epochs before t0 first.  That equals
    # the input order only for ascending times, which is what is supported here.
    t = np.asarray(time, dtype=np.float64)
    if t.ndim != 1 or np.any(np.diff(t) < 0):
        raise NotImplementedError("t0_base_model on device needs a 1-D time array in ascending order")
    return t


class _T0Spec:
    """Resolves the base model's FusedSpec from kwargs['base_model'] and adds `t0` to the needed parameters."""

    def for_kwargs(self, kwargs):
        base_model = kwargs["base_model"]                       # KeyError like the reference (phase_models.py:28)
        name = base_model if isinstance(base_model, str) else getattr(base_model, "__name__", None)
        if name not in _BASE:
            raise NotImplementedError(f"t0_base_model: base_model={base_model!r} is not fused on device "
                                      f"(supported: {sorted(_BASE)})")
        base = _SPECS[_BASE[name]]

        def build(time, kw, y, sigma, sigma_mode):
            kw = {k: v for k, v in kw.items() if k != "base_model"}
            return base.build(_check_sorted(time), kw, y=y, sigma=sigma, sigma_mode=sigma_mode)

        return FusedSpec("t0_base_model:" + name, base.needed + ("t0",), build,
                         lambda kw: base.validate({k: v for k, v in kw.items() if k != "base_model"}))


_SPEC = _T0Spec()


def t0_base_model(time, t0=None, params_batch=None, **kwargs):
    """redback phase_models.t0_base_model (:19-59); kwargs base_model + everything the base model needs."""
    return run(_SPEC.for_kwargs(kwargs), time, dict(t0=t0), kwargs, params_batch)


FUSED = {t0_base_model: _SPEC}
