"""Import selected reference source files *verbatim, where they lie* under /root/reference.

TEST INFRASTRUCTURE ONLY, and only usable in the build container (the GPU box has no
/root/reference).  `import redback` fails here because astropy / bilby / sncosmo are not installed,
but `redback/interaction_processes.py`, `redback/photosphere.py` and
`redback/transient_models/afterglow_models/base_models.py` only need a handful of names, which are
provided as stub modules in sys.modules.  Nothing is copied; the files are executed from their
original location.  Used to validate oracle/restated.py and to generate tests/golden/ fixtures.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "redback"))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _load(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Return a namespace with the reference's own Diffusion / TemperatureFloor / afterglow code."""
    if not available():
        raise RuntimeError("reference tree not present")
    from . import restated as r
    if "redback" not in sys.modules or not hasattr(sys.modules["redback"], "_oracle_stub"):
        pkg = _stub("redback", _oracle_stub=True, __path__=[])
        const = _stub(
            "redback.constants",
            speed_of_light=r.speed_of_light, planck=r.planck, boltzmann_constant=r.boltzmann_constant,
            sigma_sb=r.sigma_sb, solar_mass=r.solar_mass, proton_mass=r.proton_mass,
            electron_mass=r.electron_mass, sigma_T=r.sigma_T, km_cgs=r.km_cgs, day_to_s=r.day_to_s,
            angstrom_cgs=r.angstrom_cgs, speed_of_light_si=r.speed_of_light_si,
            radiation_constant=r.sigma_sb * 4 / r.speed_of_light, ev_to_erg=1.60218e-12,
            qe=4.80320425e-10, au_cgs=14959787070000.0, solar_radius=69570000000.0,
            graviational_constant=6.674299999999999e-08, mev_cgs=1.602176634e-06,
            ang_to_hz=2.9979245799999995e+18, jansky=1e-26)
        const.__all__ = [k for k in const.__dict__ if not k.startswith("_")]
        pkg.constants = const
    ns = types.SimpleNamespace()
    ns.interaction_processes = _load("redback.interaction_processes", "redback/interaction_processes.py")
    ns.photosphere = _load("redback.photosphere", "redback/photosphere.py")
    return ns


def load_afterglow():
    """Import afterglow_models/base_models.py verbatim (numba kernels + RedbackAfterglows)."""
    load()
    import logging
    from . import restated as r

    def _unavailable(*a, **k):
        raise RuntimeError("not available in the oracle shim")

    def citation_wrapper(_):
        def deco(f):
            return f
        return deco

    _stub("redback.utils", logger=logging.getLogger("redback-oracle"), citation_wrapper=citation_wrapper,
          calc_ABmag_from_flux_density=_unavailable, lambda_to_nu=r.lambda_to_nu, bands_to_frequency=_unavailable,
          calc_kcorrected_properties=r.calc_kcorrected_properties, nu_to_lambda=r.nu_to_lambda)
    _stub("redback.sed", get_correct_output_format_from_spectra=_unavailable)
    _load("redback.wrappers", "redback/wrappers.py")

    class _Val:
        def __init__(self, v):
            self.value = v

        @property
        def cgs(self):
            return self

    class _Cosmo:
        def luminosity_distance(self, z):
            return _Val(r.cosmo.luminosity_distance_cm(z))

    _stub("astropy")
    _stub("astropy.cosmology", Planck18=_Cosmo())
    _stub("astropy.units")
    sys.modules["astropy"].units = sys.modules["astropy.units"]
    _stub("redback.transient_models", __path__=[])
    _stub("redback.transient_models.afterglow_models", __path__=[])
    return _load("redback.transient_models.afterglow_models.base_models",
                 "redback/transient_models/afterglow_models/base_models.py")


def load_functions(relpath, names, namespace):
    """Compile the named top-level functions / classes of a reference source file *from where it lies* (decorators dropped)
    and define them in `namespace`.  For functions whose module cannot be imported as a whole here
    (utils.py / supernova_models.py need astropy, sncosmo, ...)."""
    import ast
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    picked = [node for node in tree.body if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names]
    missing = set(names) - {node.name for node in picked}
    if missing:
        raise RuntimeError(f"{relpath}: no top-level function(s) / class(es) {sorted(missing)}")
    for node in picked:
        node.decorator_list = []
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), namespace)
    return namespace


def load_csm():
    """The reference's own _csm_engine / csm_interaction_bolometric / get_csm_properties /
    get_optimal_time_array (lru_cache dropped) + CSMDiffusion, executed from /root/reference."""
    from collections import namedtuple
    import numpy as np
    from scipy.interpolate import RegularGridInterpolator
    from . import restated as r
    ns = load()
    glb = dict(np=np, namedtuple=namedtuple, RegularGridInterpolator=RegularGridInterpolator,
               dirname=os.path.join(REFERENCE_ROOT, "redback"), solar_mass=r.solar_mass, au_cgs=r.au_cgs,
               km_cgs=r.km_cgs, day_to_s=r.day_to_s, ip=ns.interaction_processes)
    load_functions("redback/utils.py", ["get_csm_properties", "get_optimal_time_array",
                                         "_get_optimal_time_array_cached"], glb)
    load_functions("redback/transient_models/supernova_models.py", ["_csm_engine", "csm_interaction_bolometric",
                                                                     "csm_nickel_bolometric", "_nickelcobalt_engine"], glb)
    out = types.SimpleNamespace(**{k: glb[k] for k in ("get_csm_properties", "get_optimal_time_array", "_csm_engine",
                                                       "csm_interaction_bolometric", "csm_nickel_bolometric")})
    out.photosphere = ns.photosphere
    out.interaction_processes = ns.interaction_processes
    return out


def load_mergernova():
    """The reference's own _ejecta_dynamics_and_interaction, basic_magnetar, magnetar_only,
    velocity_from_lorentz_factor and get_optimal_time_array, executed from /root/reference."""
    from collections import namedtuple
    import numpy as np
    from . import restated as r
    load()
    glb = dict(np=np, namedtuple=namedtuple, solar_mass=r.solar_mass, speed_of_light=r.speed_of_light,
               proton_mass=r.proton_mass, radiation_constant=r.sigma_sb * 4 / r.speed_of_light, sigma_sb=r.sigma_sb)
    load_functions("redback/utils.py", ["velocity_from_lorentz_factor", "get_optimal_time_array",
                                         "_get_optimal_time_array_cached"], glb)
    load_functions("redback/transient_models/magnetar_models.py", ["basic_magnetar", "magnetar_only"], glb)
    load_functions("redback/transient_models/magnetar_driven_ejecta_models.py", ["_ejecta_dynamics_and_interaction"], glb)
    return types.SimpleNamespace(**{k: glb[k] for k in ("_ejecta_dynamics_and_interaction", "basic_magnetar",
                                                       "magnetar_only", "get_optimal_time_array")})


def load_metzger_magnetar():
    """The reference's own _general_metzger_magnetar_driven_kilonova_model (with the utils.py helpers it calls),
    magnetar_only and get_optimal_time_array, executed from /root/reference."""
    from collections import namedtuple
    import numpy as np
    from scipy.interpolate import RegularGridInterpolator, interp1d
    from . import restated as r
    load()
    glb = dict(np=np, namedtuple=namedtuple, RegularGridInterpolator=RegularGridInterpolator, interp1d=interp1d,
               solar_mass=r.solar_mass, speed_of_light=r.speed_of_light, day_to_s=r.day_to_s, sigma_sb=r.sigma_sb)
    load_functions("redback/utils.py", ["interpolated_barnes_and_kasen_thermalisation_efficiency",
                                         "electron_fraction_from_kappa", "get_optimal_time_array",
                                         "_get_optimal_time_array_cached"], glb)
    load_functions("redback/transient_models/magnetar_models.py", ["magnetar_only"], glb)
    load_functions("redback/transient_models/magnetar_driven_ejecta_models.py",
                   ["_general_metzger_magnetar_driven_kilonova_model"], glb)
    return types.SimpleNamespace(**{k: glb[k] for k in ("_general_metzger_magnetar_driven_kilonova_model",
                                                       "magnetar_only", "get_optimal_time_array")})


def load_ejecta_relations():
    """The reference's own ejecta-relation classes and helper fits (redback/ejecta_relations.py imports astropy at
    module level; the classes themselves only need numpy), executed from /root/reference."""
    import numpy as np
    glb = dict(np=np)
    names = ["OneComponentBNSNoProjection", "OneComponentBNSProjection", "TwoComponentBNS", "TwoComponentNSBH",
             "OneComponentNSBH", "calc_compactness_from_lambda", "calc_baryonic_mass", "calc_vrho", "calc_vz"]
    load_functions("redback/ejecta_relations.py", names, glb)
    return types.SimpleNamespace(**{k: glb[k] for k in names})
