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
