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


def _host_shift(time, t0, function, pad, params_batch, kwargs):
    """Scalar explosion epoch for a base model without a fused t0: time - t0 on the host, epochs before t0 -> `pad`
    (phase_models.py:36-58; frequency / bands sliced like :44-49)."""
    import torch
    t = _check_sorted(time) - float(t0)
    good = t >= 0.0
    kw = dict(kwargs)
    for key in ("frequency", "bands"):
        if isinstance(kw.get(key), np.ndarray):
            kw[key] = kw[key][good]
    if params_batch is not None:
        kw["params_batch"] = params_batch
    real = function(t[good], **kw)
    n_bad = int((~good).sum())
    if isinstance(real, torch.Tensor):
        return torch.cat((real.new_full(real.shape[:-1] + (n_bad,), pad), real), dim=-1)
    return np.concatenate((np.full(np.shape(real)[:-1] + (n_bad,), pad), real), axis=-1)


def t0_base_model(time, t0=None, params_batch=None, **kwargs):
    """redback phase_models.t0_base_model (:19-59); kwargs base_model + everything the base model needs.  Base models
    with a fused explosion epoch (module docstring) take `t0` per sample; any other fused model (afterglows, CSM
    interaction, ...) a scalar `t0`, shifted on the host."""
    base_model = kwargs["base_model"]                            # KeyError like the reference (:28)
    name = base_model if isinstance(base_model, str) else getattr(base_model, "__name__", None)
    if name not in _BASE:
        import redback_b200 as rb
        batch = dict(params_batch) if params_batch is not None else {}
        t0 = batch.pop("t0", t0)
        function = rb.all_models_dict.get(name)
        if function is None or function is t0_base_model or t0 is None or np.ndim(t0):
            raise NotImplementedError(f"t0_base_model: base_model={base_model!r} is not fused on device with a sampled "
                                      f"t0 (per-sample t0: {sorted(_BASE)}; scalar t0: every fused model)")
        kw = {k: v for k, v in kwargs.items() if k != "base_model"}
        return _host_shift(time, t0, function, 1000.0 if kwargs["output_format"] == "magnitude" else 0.0, batch or None, kw)
    return run(_SPEC.for_kwargs(kwargs), time, dict(t0=t0), kwargs, params_batch)


FUSED = {t0_base_model: _SPEC}


def _t0_with_extinction(time, t0, av_host, model_type, params_batch=None, **kwargs):
    """phase_models._t0_with_extinction (:57-86): the extinction wrapper of `model_type` evaluated at time - t0, epochs
    before t0 -> 0 (5000 for magnitudes, where t0_base_model uses 1000).  Base models with a fused explosion epoch
    (see t0_base_model) take a sampled t0 per sample; the others (afterglows) a scalar t0, shifted on the host."""
    from . import extinction_models as _em
    import torch
    base = _em._get_correct_function(kwargs["base_model"], model_type)
    name = getattr(base, "__name__", None)
    magnitude = kwargs["output_format"] == "magnitude"
    batch = dict(params_batch) if params_batch is not None else {}
    t0 = batch.pop("t0", t0)
    if t0 is None:
        raise KeyError("t0")
    if _BASE.get(name) is base:
        def shifted(time, **kw):
            return t0_base_model(time, base_model=name, **kw)
        if np.ndim(t0):
            batch["t0"] = t0
        else:
            kwargs["t0"] = float(t0)
        out = _em._evaluate_extinction_model(time, av_host, model_type=model_type, params_batch=batch or None,
                                             _function=shifted, **kwargs)
        if magnitude:       # exactly 1000.0 is t0_base_model's placeholder (phase_models.py:54 vs :81)
            out = torch.where(out == 1000.0, 5000.0, out) if isinstance(out, torch.Tensor) else \
                np.where(out == 1000.0, 5000.0, out)
        return out
    if np.ndim(t0):
        raise NotImplementedError(f"t0 per sample is not fused for base_model={name!r}; pass a scalar t0")
    def extinguished(t, **kw):
        return _em._evaluate_extinction_model(t, av_host, model_type=model_type, **kw)
    # (the reference passes frequency / bands unsliced, which only works when no epoch is cut)
    return _host_shift(time, t0, extinguished, 5000.0 if magnitude else 0.0, batch or None, kwargs)


def t0_afterglow_extinction(time, t0=None, av_host=None, params_batch=None, **kwargs):
    """redback phase_models.t0_afterglow_extinction (:89-107)."""
    return _t0_with_extinction(time, t0, av_host, "afterglow", params_batch=params_batch, **kwargs)


def t0_supernova_extinction(time, t0=None, av_host=None, params_batch=None, **kwargs):
    """redback phase_models.t0_supernova_extinction (:110-128)."""
    return _t0_with_extinction(time, t0, av_host, "supernova", params_batch=params_batch, **kwargs)


def t0_kilonova_extinction(time, t0=None, av_host=None, params_batch=None, **kwargs):
    """redback phase_models.t0_kilonova_extinction (:131-149)."""
    return _t0_with_extinction(time, t0, av_host, "kilonova", params_batch=params_batch, **kwargs)


def t0_tde_extinction(time, t0=None, av_host=None, params_batch=None, **kwargs):
    """redback phase_models.t0_tde_extinction (:152-173)."""
    return _t0_with_extinction(time, t0, av_host, "tde", params_batch=params_batch, **kwargs)


def t0_magnetar_driven_extinction(time, t0=None, av_host=None, params_batch=None, **kwargs):
    """redback phase_models.t0_magnetar_driven_extinction (:176-194)."""
    return _t0_with_extinction(time, t0, av_host, "magnetar_driven", params_batch=params_batch, **kwargs)


EXTINCTION_MODELS = (t0_afterglow_extinction, t0_supernova_extinction, t0_kilonova_extinction, t0_tde_extinction,
                     t0_magnetar_driven_extinction)
