"""CPU: the C-ABI library builds/loads and exports every symbol include/redback_b200.h declares."""
import ctypes
import os
import re

from redback_b200 import _build, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "redback_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _capi.lib()
    names = _declared()
    assert len(names) >= 10
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_capi.EXPORTS), set(names) ^ set(_capi.EXPORTS)


def test_host_side_calls_without_gpu():
    lib = _capi.lib()
    assert b"sm_100a" in lib.rb_version()
    cfg = _capi.SnConfig()
    assert lib.rb_sn_config_default(ctypes.byref(cfg)) == 0
    assert cfg.dense_resolution == 1000 and cfg.cutoff_wavelength == 3000.0
    # Planck18 derived densities (astropy: Ogamma0 = 5.4020e-5, Ode0 = 0.68885)
    assert abs(cfg.cosmology.Ogamma0 - 5.402015e-5) < 1e-9
    assert abs(cfg.cosmology.Ode0 - 0.688847) < 1e-5
    # argument validation happens before any CUDA call
    assert lib.rb_sn_eval_f64(ctypes.byref(cfg), 1, 0, *([None] * 9), None) == _capi.RB_ERR_INVALID
    cfg.sed = _capi.SED_CUTOFF_BLACKBODY
    cfg.absorption_index = 4.5
    one = ctypes.c_void_p(8)
    assert lib.rb_sn_eval_f64(ctypes.byref(cfg), 1, 3, one, one, one, None, None, None, None, one, None, None) \
        == _capi.RB_ERR_INVALID
    assert b"absorption_index" in lib.rb_last_error()


def test_library_is_in_tree():
    assert _build.LIB_PATH.startswith(ROOT)


def test_ctypes_struct_layouts_match_the_library():
    """The ctypes mirrors must have the C structs' sizes (a config_default memset would overrun a short mirror), and the
    defaults must land in the right fields (offsets)."""
    from redback_b200 import bands
    lib = _capi.lib()
    mirrors = [_capi.Cosmology, _capi.UpperLimits, _capi.SnConfig, _capi.KnConfig, _capi.AgConfig, _capi.MnConfig,
               _capi.MkConfig, bands.Photometry, _capi.Extinction]
    for which, cls in enumerate(mirrors):
        assert lib.rb_sizeof(which) == ctypes.sizeof(cls), (cls.__name__, lib.rb_sizeof(which), ctypes.sizeof(cls))
    assert lib.rb_sizeof(99) == -1
    sn = _capi.sn_config(_capi.ENGINE_NICKEL, _capi.SED_BLACKBODY)
    assert (sn.csm_nn, sn.csm_delta, sn.csm_efficiency, sn.line_amplitude, sn.no_interaction) == (12.0, 1.0, 0.5, 0.3, 0)
    kn = _capi.kn_config(_capi.KN_ONE_COMPONENT)
    assert kn.grid_t_max == 7e6 and kn.vmax == 0.7
    mn = _capi.mn_config(_capi.MN_BASIC_MAGNETAR)
    assert (mn.resolution, mn.photon_index, mn.ejecta_albedo) == (500, 2.0, 0.5)
    mk = _capi.mk_config(_capi.MN_BASIC_MAGNETAR)
    assert (mk.resolution, mk.pair_cascade_fraction, mk.vmax, mk.pair_cascade_switch) == (300, 0.01, 0.7, 1)
    ag = _capi.ag_config(_capi.AG_TOPHAT)
    assert (ag.steps, ag.ab_magnitude) == (250, 0)


def test_integration_stub_is_generated_from_the_bindings():
    """INTEGRATION.md section 2 shows the structures a foreign caller mirrors; the text between the stub markers must be
    exactly what tools/gen_integration_stub.py produces from redback_b200/_capi.py, and must execute."""
    import ctypes
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_integration_stub", os.path.join(root, "tools",
                                                                                       "gen_integration_stub.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    committed = doc[doc.index(gen.BEGIN) + len(gen.BEGIN):doc.index(gen.END)]
    assert committed == gen.stub()
    # the class definitions of the stub, executed on their own, have the library's layout
    code = committed.split("```python\n")[1].split("assert lib.rb_sizeof")[0]
    code = "\n".join(line for line in code.splitlines() if not line.startswith("lib"))
    ns = {}
    exec(code, ns)
    from redback_b200 import _capi
    assert ctypes.sizeof(ns["SnConfig"]) == _capi.lib().rb_sizeof(2) == ctypes.sizeof(_capi.SnConfig)
