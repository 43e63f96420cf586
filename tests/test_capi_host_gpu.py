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


def _slsn_problem(rng, N):
    from redback_b200 import _capi
    t, nu = h.config2_epochs(rng)
    p = h.slsn_prior(rng, N)
    y = r.slsn(t, 0.2, 2.0, 1.0, 1.4, 0.5, mej=5.0, vej=8000.0, kappa=0.1, kappa_gamma=0.1, temperature_floor=5e3,
               frequency=nu)
    sigma = 0.1 * np.abs(y) + 1e-9
    rows = np.zeros((_capi.SN_NPARAM, N))
    for k, v in p.items():
        rows[_capi.SN_ROWS[k]] = v
    return t, nu, p, np.ascontiguousarray(rows), y, sigma


def test_sn_host_pipeline_million_samples_bitwise():
    """A 10^6-sample host call (pageable numpy buffers, several chunks, ragged last chunk) returns bit-for-bit what
    the device-pointer entry returns, for the block form, the row-pointer form and repeated calls (buffer reuse)."""
    import torch
    from redback_b200 import _capi
    from redback_b200.engine import SupernovaPath
    lib = _capi.lib()
    rng = np.random.default_rng(11)
    N = 1_000_003
    t, nu, p, rows, y, sigma = _slsn_problem(rng, N)
    E = t.size
    path = SupernovaPath(_capi.ENGINE_MAGNETAR, _capi.SED_BLACKBODY, t, frequency=nu, y=y, sigma=sigma)
    _, ref = path.evaluate(torch.from_numpy(rows).cuda(), want_model=False, want_logl=True)
    ref = ref.cpu().numpy()
    cfg = _capi.sn_config(_capi.ENGINE_MAGNETAR, _capi.SED_BLACKBODY)
    for rep in range(2):
        logl = np.full(N, -1.0)
        rc = lib.rb_sn_eval_f64_host(C.byref(cfg), N, E, _p(rows), _p(t), _p(nu), None, _p(y), _p(sigma), None, None,
                                     _p(logl))
        assert rc == 0, lib.rb_last_error()
        assert np.array_equal(logl, ref)
    # row-pointer form with one constant row (kappa fixed) through the Python path object
    q = dict(p)
    q["kappa"] = 0.123
    rows2 = rows.copy()
    rows2[_capi.SN_ROWS["kappa"]] = 0.123
    _, ref2 = path.evaluate(torch.from_numpy(rows2).cuda(), want_model=False, want_logl=True)
    _, got2 = path.evaluate_host(q, N)
    assert np.array_equal(got2, ref2.cpu().numpy())
    # model output: smaller chunks (96 MB cap), a few thousand samples are enough
    M = 70_001
    sub = {k: (v[:M] if np.ndim(v) else v) for k, v in q.items()}
    model_h, logl_h = path.evaluate_host(sub, M, want_model=True)
    model_d, logl_d = path.evaluate(torch.from_numpy(np.ascontiguousarray(rows2[:, :M])).cuda(), want_model=True,
                                    want_logl=True)
    assert np.array_equal(model_h, model_d.cpu().numpy(), equal_nan=True)
    assert np.array_equal(logl_h, logl_d.cpu().numpy())


def test_likelihood_batch_host_route_and_attribute_edits():
    """GaussianLikelihood.log_likelihood_batch with a dict of numpy arrays goes through the host pipeline and equals
    the device route; editing y / sigma / kwargs['frequency'] (by assignment or in place) is honoured on the next call."""
    import torch
    import redback_b200 as rb
    rng = np.random.default_rng(12)
    N = 3000
    t, nu, p, rows, y, sigma = _slsn_problem(rng, N)
    kw = dict(frequency=nu.copy(), output_format="flux_density")
    like = rb.GaussianLikelihood(t, y.copy(), sigma.copy(), rb.supernova_models.slsn, kwargs=kw)
    a = like.log_likelihood_batch(p)
    b = like.log_likelihood_batch({k: torch.from_numpy(v).cuda() for k, v in p.items()})
    assert np.array_equal(a, b)

    def fresh(y_, s_, nu_):
        other = rb.GaussianLikelihood(t, y_, s_, rb.supernova_models.slsn,
                                      kwargs=dict(frequency=nu_, output_format="flux_density"))
        return other.log_likelihood_batch(p)

    like.y = y * 1.5                                   # assignment
    assert np.array_equal(like.log_likelihood_batch(p), fresh(y * 1.5, sigma, nu))
    like.y[::2] *= 0.5                                 # in-place edit
    y2 = y * 1.5
    y2[::2] *= 0.5
    assert np.array_equal(like.log_likelihood_batch(p), fresh(y2, sigma, nu))
    like.sigma = sigma * 2.0                           # public setter
    assert np.array_equal(like.log_likelihood_batch(p), fresh(y2, sigma * 2.0, nu))
    like.kwargs["frequency"][:] = nu[::-1]             # in-place edit of a kwarg array
    assert np.array_equal(like.log_likelihood_batch(p), fresh(y2, sigma * 2.0, nu[::-1].copy()))
    assert not np.array_equal(a, like.log_likelihood_batch(p))


def test_second_device_in_one_process():
    """cudaFuncSetAttribute / the pool threshold are per device: evaluating on cuda:1 after cuda:0 must work."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from redback_b200 import _capi
    from redback_b200.engine import SupernovaPath
    rng = np.random.default_rng(13)
    N = 500
    t, nu, p, rows, y, sigma = _slsn_problem(rng, N)
    out = []
    for dev in ("cuda:0", "cuda:1"):
        path = SupernovaPath(_capi.ENGINE_MAGNETAR, _capi.SED_BLACKBODY, t, frequency=nu, y=y, sigma=sigma, device=dev)
        _, logl = path.evaluate(torch.from_numpy(rows).to(dev), want_model=False, want_logl=True)
        out.append(logl.cpu().numpy())
        with torch.cuda.device(dev):
            _, lh = path.evaluate_host(p, N)
        assert np.array_equal(lh, out[-1])
    assert np.array_equal(out[0], out[1])
