"""Batched drop-ins for redback's one-component and Metzger kilonova models
(redback/transient_models/kilonova_models.py:1537-1603 and :1895-1962), output_format='flux_density'.
Same signatures as the reference plus `params_batch` (see supernova_models.py)."""
from . import _capi
from ._fused import FusedSpec, cosmology_from_kwargs, output_mode, run
from .engine import KilonovaMagPath, KilonovaMultiMagPath, KilonovaPath


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
                            device=kwargs.get("device"), dtype=kwargs.get("dtype", "float64"))

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


def _multi_component(name, time, redshift, components, params_batch, kwargs, dense_resolution, grid_t_max):
    """Sum of one-component models on a shared time grid (kilonova_models.py:1113-1260, flux_density branch): one
    launch of the one-component kernel per component, the flux densities are added on device."""
    import numpy as np
    from ._fused import collect, dl_from_kwargs, finish
    if kwargs.get("output_format") is None:
        raise KeyError("output_format")
    mode = output_mode(kwargs, name)
    names = ["redshift"] + [f"{k}_{c}" for c in range(1, len(components) + 1)
                            for k in ("mej", "vej", "kappa", "temperature_floor")]
    named = dict(redshift=redshift)
    for c, comp in enumerate(components, start=1):
        named.update({f"{k}_{c}": v for k, v in zip(("mej", "vej", "temperature_floor", "kappa"), comp)})
    params, N = collect(named, kwargs, params_batch, names)
    if mode != "flux_density":
        # kilonova_models.py:1166-1211, 1277-1323: the components' spectra are summed before the band integration
        path = KilonovaMultiMagPath(len(components), time, kwargs.get("bands") if mode == "spectra" else kwargs["bands"],
                                    grid_t_max, output=mode,
                                    lambda_array=kwargs.get("lambda_array"),
                                    dense_resolution=kwargs.get("dense_resolution", dense_resolution),
                                    cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"))
        rows = {}
        for c in range(1, len(components) + 1):
            rows[f"redshift_{c}"] = params["redshift"]
            for k in ("mej", "vej", "kappa", "temperature_floor"):
                rows[f"{k}_{c}"] = params[f"{k}_{c}"]
        n = N if N is not None else 1
        if mode == "spectra":                                                    # :1190-1194, :1301-1305
            from ._fused import run_spectra
            return run_spectra(path, path.pack(rows, n), {}, N, kwargs, None)
        model, _ = path.evaluate(path.pack(rows, n), dl=dl_from_kwargs(kwargs, path, N))
        return finish(model, N, kwargs)
    path = KilonovaPath(_capi.KN_ONE_COMPONENT, time, frequency=kwargs.get("frequency"),
                        dense_resolution=kwargs.get("dense_resolution", dense_resolution),
                        cosmology=cosmology_from_kwargs(kwargs), device=kwargs.get("device"), grid_t_max=grid_t_max)
    dl = dl_from_kwargs(kwargs, path, N)
    total = None
    for c in range(1, len(components) + 1):
        comp = dict(redshift=params["redshift"], mej=params[f"mej_{c}"], vej=params[f"vej_{c}"],
                    kappa=params[f"kappa_{c}"], temperature_floor=params[f"temperature_floor_{c}"])
        model, _ = path.evaluate(path.pack(comp, N if N is not None else 1), dl=dl)
        total = model if total is None else total.add_(model)
    return finish(total, N, kwargs)


def two_component_kilonova_model(time, redshift=None, mej_1=None, vej_1=None, temperature_floor_1=None, kappa_1=None,
                                 mej_2=None, vej_2=None, temperature_floor_2=None, kappa_2=None, params_batch=None,
                                 **kwargs):
    """redback kilonova_models.two_component_kilonova_model (:1214-1323), flux_density and magnitude / flux; grid end
    86400*6 s, 500 points by default.  (Not fused with the likelihood: log_likelihood_batch reduces the summed model matrix.)"""
    comps = [(mej_1, vej_1, temperature_floor_1, kappa_1), (mej_2, vej_2, temperature_floor_2, kappa_2)]
    return _multi_component("two_component_kilonova_model", time, redshift, comps, params_batch, kwargs, 500,
                            86400 * 6)


def three_component_kilonova_model(time, redshift=None, mej_1=None, vej_1=None, temperature_floor_1=None, kappa_1=None,
                                   mej_2=None, vej_2=None, temperature_floor_2=None, kappa_2=None, mej_3=None,
                                   vej_3=None, temperature_floor_3=None, kappa_3=None, params_batch=None, **kwargs):
    """redback kilonova_models.three_component_kilonova_model (:1113-1211), flux_density and magnitude / flux; 300 grid
    points by default."""
    comps = [(mej_1, vej_1, temperature_floor_1, kappa_1), (mej_2, vej_2, temperature_floor_2, kappa_2),
             (mej_3, vej_3, temperature_floor_3, kappa_3)]
    return _multi_component("three_component_kilonova_model", time, redshift, comps, params_batch, kwargs, 300, 7e6)


def _via_ejecta_relation(name, time, redshift, kappa, relation_args, default_relation, params_batch, kwargs):
    """kilonova_models.py:1309-1366, 1470-1493: (mej, vej) from the ejecta relation of the gravitational-wave source
    parameters (kwarg `ejecta_relation` replaces the class, as in the reference), then one_component_kilonova_model."""
    import numpy as np
    kwargs = dict(kwargs)
    relation = kwargs.pop("ejecta_relation", default_relation)
    src = dict(relation_args)
    if params_batch is not None:
        missing = [k for k in src if k not in params_batch]
        if missing:
            raise KeyError(f"{name}: params_batch lacks {missing}")
        src = {k: params_batch[k] for k in src}
        with np.errstate(all="ignore"):        # out-of-range draws give NaN light curves, as in the reference
            rel = relation(**src)
        batch = {k: v for k, v in params_batch.items() if k not in src}
        batch.update(mej=np.asarray(rel.ejecta_mass, dtype=np.float64), vej=np.asarray(rel.ejecta_velocity, dtype=np.float64))
        return one_component_kilonova_model(time, params_batch=batch, **kwargs)
    with np.errstate(all="ignore"):
        rel = relation(**src)
    return one_component_kilonova_model(time, redshift, float(rel.ejecta_mass), float(rel.ejecta_velocity), kappa, **kwargs)


def one_component_ejecta_relation(time, redshift=None, mass_1=None, mass_2=None, lambda_1=None, lambda_2=None,
                                  kappa=None, params_batch=None, **kwargs):
    """redback kilonova_models.one_component_ejecta_relation (:1309-1336): Dietrich & Ujevic fits without projection."""
    from .ejecta_relations import OneComponentBNSNoProjection
    return _via_ejecta_relation("one_component_ejecta_relation", time, redshift, kappa,
                                dict(mass_1=mass_1, mass_2=mass_2, lambda_1=lambda_1, lambda_2=lambda_2),
                                OneComponentBNSNoProjection, params_batch, kwargs)


def one_component_ejecta_relation_projection(time, redshift=None, mass_1=None, mass_2=None, lambda_1=None,
                                             lambda_2=None, kappa=None, params_batch=None, **kwargs):
    """redback kilonova_models.one_component_ejecta_relation_projection (:1339-1366)."""
    from .ejecta_relations import OneComponentBNSProjection
    return _via_ejecta_relation("one_component_ejecta_relation_projection", time, redshift, kappa,
                                dict(mass_1=mass_1, mass_2=mass_2, lambda_1=lambda_1, lambda_2=lambda_2),
                                OneComponentBNSProjection, params_batch, kwargs)


def one_component_nsbh_ejecta_relation(time, redshift=None, mass_bh=None, mass_ns=None, chi_bh=None, lambda_ns=None,
                                       kappa=None, params_batch=None, **kwargs):
    """redback kilonova_models.one_component_nsbh_ejecta_relation (:1470-1493): Kawaguchi et al. fits."""
    from .ejecta_relations import OneComponentNSBH
    return _via_ejecta_relation("one_component_nsbh_ejecta_relation", time, redshift, kappa,
                                dict(mass_bh=mass_bh, mass_ns=mass_ns, chi_bh=chi_bh, lambda_ns=lambda_ns),
                                OneComponentNSBH, params_batch, kwargs)


def _two_component_via_relation(name, time, redshift, relation_args, rest, default_relation, params_batch, kwargs):
    """kilonova_models.py:1369-1397, 1496-1527: dynamical and disk-wind ejecta masses and the dynamical velocity from
    the ejecta relation, then two_component_kilonova_model (tf_1 / tf_2 are its temperature floors)."""
    import numpy as np
    kwargs = dict(kwargs)
    relation = kwargs.pop("ejecta_relation", default_relation)
    src, other = dict(relation_args), dict(rest, redshift=redshift)
    if params_batch is not None:
        missing = [k for k in list(src) + list(other) if k not in params_batch]
        if missing:
            raise KeyError(f"{name}: params_batch lacks {missing}")
        src = {k: params_batch[k] for k in src}
        other = {k: params_batch[k] for k in other}
    with np.errstate(all="ignore"):                  # out-of-range draws give NaN light curves, as in the reference
        rel = relation(**src)
    comp = dict(mej_1=rel.dynamical_mej, vej_1=rel.ejecta_velocity, temperature_floor_1=other["tf_1"],
                kappa_1=other["kappa_1"], mej_2=rel.disk_wind_mej, vej_2=other["vej_2"],
                temperature_floor_2=other["tf_2"], kappa_2=other["kappa_2"])
    if params_batch is not None:
        batch = {k: np.asarray(v, dtype=np.float64) for k, v in comp.items()}
        batch["redshift"] = other["redshift"]
        extra = {k: v for k, v in params_batch.items() if k not in src and k not in other}
        return two_component_kilonova_model(time, params_batch=dict(extra, **batch), **kwargs)
    return two_component_kilonova_model(time, other["redshift"], **{k: float(v) for k, v in comp.items()}, **kwargs)


def two_component_bns_ejecta_relation(time, redshift=None, mass_1=None, mass_2=None, lambda_1=None, lambda_2=None,
                                      mtov=None, zeta=None, vej_2=None, kappa_1=None, kappa_2=None, tf_1=None,
                                      tf_2=None, params_batch=None, **kwargs):
    """redback kilonova_models.two_component_bns_ejecta_relation (:1369-1397)."""
    from .ejecta_relations import TwoComponentBNS
    return _two_component_via_relation(
        "two_component_bns_ejecta_relation", time, redshift,
        dict(mass_1=mass_1, mass_2=mass_2, lambda_1=lambda_1, lambda_2=lambda_2, mtov=mtov, zeta=zeta),
        dict(vej_2=vej_2, kappa_1=kappa_1, kappa_2=kappa_2, tf_1=tf_1, tf_2=tf_2), TwoComponentBNS, params_batch, kwargs)


def two_component_nsbh_ejecta_relation(time, redshift=None, mass_bh=None, mass_ns=None, chi_bh=None, lambda_ns=None,
                                       zeta=None, vej_2=None, kappa_1=None, kappa_2=None, tf_1=None, tf_2=None,
                                       params_batch=None, **kwargs):
    """redback kilonova_models.two_component_nsbh_ejecta_relation (:1496-1527)."""
    from .ejecta_relations import TwoComponentNSBH
    return _two_component_via_relation(
        "two_component_nsbh_ejecta_relation", time, redshift,
        dict(mass_bh=mass_bh, mass_ns=mass_ns, chi_bh=chi_bh, lambda_ns=lambda_ns, zeta=zeta),
        dict(vej_2=vej_2, kappa_1=kappa_1, kappa_2=kappa_2, tf_1=tf_1, tf_2=tf_2), TwoComponentNSBH, params_batch, kwargs)


FUSED = {one_component_kilonova_model: _ONE, metzger_kilonova_model: _METZGER}
