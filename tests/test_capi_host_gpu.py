"""GPU: the host-buffer C-ABI entry points (plain numpy pointers, no torch) agree with the device-pointer path."""
import ctypes as C

import numpy as np
import pytest

import helpers as h
from oracle import restated as r

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def test_sn_host_entry_point():
    from redback_b200 import _capi
    lib = _capi.lib()
    rng = np.random.default_rng(3)
    t, nu = h.config2_epochs(rng)
    N, E = 50, t.size
    p = h.slsn_prior(rng, N)
    rows = np.zeros((_capi.SN_NPARAM, N))
    for k, v in p.items():
        rows[_capi.SN_ROWS[k]] = v
    rows = np.ascontiguousarray(rows)
    y = r.slsn(t, 0.2, 2.0, 1.0, 1.4, 0.5, mej=5.0, vej=8000.0, kappa=0.1, kappa_gamma=0.1, temperature_floor=5e3,
               frequency=nu)
    sigma = 0.1 * np.abs(y) + 1e-9
    cfg = _capi.sn_config(_capi.ENGINE_MAGNETAR, _capi.SED_CUTOFF_BLACKBODY)
    model = np.empty((N, E))
    logl = np.empty(N)
    rc = lib.rb_sn_eval_f64_host(C.byref(cfg), N, E, _p(rows), _p(t), _p(nu), None, _p(y), _p(sigma), None,
                                 _p(model), _p(logl))
    assert rc == 0, lib.rb_last_error()
    for i in range(0, N, 7):
        kw = {k: v[i] for k, v in p.items()}
        z = kw.pop("redshift")
        ref = r.slsn(t, z, frequency=nu, **kw)
        tau = np.sqrt(2.0 * r.solar_mass / (13.7 * r.speed_of_light * r.km_cgs) * kw["kappa"] * kw["mej"] / kw["vej"]) / 86400
        h.assert_close(model[i], ref, 1e-9, h.diffusion_condition(t / (1 + z), tau), what=f"sample {i}")
        ref_l = r.gaussian_likelihood(y, ref, sigma)
        assert abs(logl[i] - ref_l) <= 1e-6 + 1e-9 * abs(ref_l)
    # error reporting through the ABI
    assert lib.rb_sn_eval_f64_host(C.byref(cfg), N, E, _p(rows), _p(t), None, None, None, None, None, _p(model), None) \
        == _capi.RB_ERR_INVALID
    assert b"freq_obs" in lib.rb_last_error()


def test_kn_and_ag_host_entry_points():
    from redback_b200 import _capi
    lib = _capi.lib()
    rng = np.random.default_rng(4)
    t, nu = h.config3_epochs(30)
    N = 8
    p = h.one_component_prior(rng, N)
    rows = np.zeros((_capi.KN_NPARAM, N))
    for k, v in p.items():
        rows[_capi.KN_ROWS[k]] = v
    cfg = _capi.kn_config(_capi.KN_ONE_COMPONENT)
    model = np.empty((N, t.size))
    rc = lib.rb_kn_eval_f64_host(C.byref(cfg), N, t.size, _p(np.ascontiguousarray(rows)), _p(t), _p(nu), None, None,
                                 None, None, _p(model), None)
    assert rc == 0, lib.rb_last_error()
    kw = {k: v[0] for k, v in p.items()}
    z = kw.pop("redshift")
    h.assert_close(model[0], r.one_component_kilonova_model(t, z, frequency=nu, **kw), 1e-9, 1e-8)

    ta, nua = h.config4_epochs(10)
    pa = h.tophat_prior(rng, N)
    rows = np.zeros((_capi.AG_NPARAM, N))
    for k, v in pa.items():
        rows[_capi.AG_ROWS[k]] = v
    acfg = _capi.ag_config(_capi.AG_TOPHAT)
    amodel = np.empty((N, ta.size))
    rc = lib.rb_ag_eval_f64_host(C.byref(acfg), N, ta.size, _p(np.ascontiguousarray(rows)), _p(ta), _p(nua), None, None,
                                 None, None, _p(amodel), None)
    assert rc == 0, lib.rb_last_error()
    kw = {k: v[1] for k, v in pa.items()}
    z = kw.pop("redshift")
    h.assert_close(amodel[1], r.redback_afterglow(ta, z, frequency=nua, **kw), 1e-9)
