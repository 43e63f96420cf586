"""Fixtures produced by the reference's OWN source files (imported verbatim in the build container by
oracle/make_golden.py::dump_verbatim): interaction_processes.Diffusion + photosphere.TemperatureFloor and the six
redback-native afterglow models on seeded prior draws.  CPU: the oracle reproduces them; GPU: so does the device."""
import json
import os

import numpy as np
import pytest

import helpers as h
from oracle import restated as r

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METHOD = dict(tophat_redback=1, gaussian_redback=2, powerlaw_redback=3, alternativepowerlaw_redback=4,
              twocomponent_redback=5, doublegaussian_redback=6)
EXTRA = dict(tophat_redback=(0.01, 0.5), gaussian_redback=(0.01, 0.5), twocomponent_redback=(0.01, 4),
             powerlaw_redback=(3, -3), alternativepowerlaw_redback=(3, 3), doublegaussian_redback=(0.1, 0.5))


@pytest.fixture(scope="module")
def fx():
    with open(os.path.join(ROOT, "tests", "golden", "verbatim_reference_draws.json")) as fh:
        return json.load(fh)


def _bolometric(kind, t, p):
    common = dict(mej=p["mej"], vej=p["vej"], kappa=p["kappa"], kappa_gamma=p["kappa_gamma"])
    if kind == "nickel":
        return ("arnett_bolometric", dict(f_nickel=p["f_nickel"], **common))
    return ("basic_magnetar_powered_bolometric", dict(p0=p["p0"], bp=p["bp"], mass_ns=p["mass_ns"],
                                                      theta_pb=p["theta_pb"], **common))


def test_oracle_reproduces_verbatim_reference(fx):
    t = np.array(fx["diffusion_time"])
    for row in fx["diffusion"]:
        name, kw = _bolometric(row["params"]["engine"], t, row["params"])
        out = getattr(r, name)(t, **kw)
        np.testing.assert_allclose(out, row["lbol"], rtol=1e-14)
        T, R = r.temperature_floor(t, out, row["params"]["vej"], row["params"]["temperature_floor"])
        np.testing.assert_allclose(T, row["temperature"], rtol=1e-13)
        np.testing.assert_allclose(R, row["radius"], rtol=1e-13)
    ta, fr = np.array(fx["afterglow_time"]), np.array(fx["afterglow_frequency"])
    for row in fx["afterglow"][:6]:
        ss, aa = EXTRA[row["model"]]
        out = r.redback_afterglow(ta, row["redshift"], frequency=fr, method_id=METHOD[row["model"]], k=row["k"],
                                  ss=ss, aa=aa, **row["params"])
        np.testing.assert_allclose(out, row["results"], rtol=1e-12)


@pytest.mark.gpu
def test_device_reproduces_verbatim_reference(fx):
    import redback_b200 as rb
    t = np.array(fx["diffusion_time"])
    for row in fx["diffusion"]:
        name, kw = _bolometric(row["params"]["engine"], t, row["params"])
        out = getattr(rb.supernova_models, name)(t, **kw)
        h.assert_close(out, np.array(row["lbol"]), 1e-9, h.diffusion_condition(t, row["tau_d"]), what=name)
    ta, fr = np.array(fx["afterglow_time"]), np.array(fx["afterglow_frequency"])
    for row in fx["afterglow"]:
        out = getattr(rb.afterglow_models, row["model"])(ta, row["redshift"], frequency=fr,
                                                          output_format="flux_density", k=row["k"], **row["params"])
        h.assert_close(out, np.array(row["results"]), 1e-9, what=row["model"])


@pytest.fixture(scope="module")
def csm_fx():
    with open(os.path.join(ROOT, "tests", "golden", "csm_reference_draws.json")) as fh:
        return json.load(fh)


def test_oracle_reproduces_reference_csm(csm_fx):
    """_csm_engine + CSMDiffusion + TemperatureFloor of the reference's own sources (make_golden.py::dump_csm)."""
    t_all = np.array(csm_fx["time"])
    for row in csm_fx["draws"]:
        t = t_all[row["skip"]:]
        p = dict(row["params"])
        tfl = p.pop("temperature_floor")
        eng, r_ph, m_th = r.csm_engine(t, **p)
        np.testing.assert_allclose(eng, row["engine_lbol"], rtol=1e-14)
        assert abs(r_ph / row["r_photosphere"] - 1) < 1e-14 and abs(m_th / row["mass_csm_threshold"] - 1) < 1e-14
        out = r.csm_interaction_bolometric(t, **p)
        if "lbol_quadrature" in row:      # the legacy quadrature method and csm_nickel (both methods)
            np.testing.assert_allclose(r.csm_interaction_bolometric(t, csm_diffusion_method="quadrature", **p),
                                       row["lbol_quadrature"], rtol=1e-13)
            c = {k: v for k, v in p.items() if k != "vej"}
            np.testing.assert_allclose(r.csm_nickel_bolometric(t, **c, **row["nickel_params"]), row["csm_nickel_lbol"],
                                       rtol=1e-13)
            np.testing.assert_allclose(
                r.csm_nickel_bolometric(t, csm_diffusion_method="legacy", csm_diffusion_timesteps=1000, **c,
                                        **row["nickel_params"]), row["csm_nickel_lbol_quadrature"], rtol=1e-13)
        np.testing.assert_allclose(out, row["lbol"], rtol=1e-14)
        T, R = r.temperature_floor(t, out, p["vej"], tfl)
        np.testing.assert_allclose(T, row["temperature"], rtol=1e-13)
        np.testing.assert_allclose(R, row["radius"], rtol=1e-13)


def test_oracle_reproduces_reference_mergernova():
    """_ejecta_dynamics_and_interaction of the reference's own source (make_golden.py::dump_mergernova)."""
    with open(os.path.join(ROOT, "tests", "golden", "mergernova_reference_draws.json")) as fh:
        fx = json.load(fh)
    tg = r.get_optimal_time_array(1e-4, 1e8, 500)
    assert tg.size == fx["n_grid"]
    pick = slice(None, None, fx["stride"])
    for row in fx["draws"]:
        if row["kind"] == "basic":
            lum = r.basic_magnetar(tg, **row["engine"])
        else:
            lum = r.magnetar_only(tg, row["engine"]["l0"], row["engine"]["tau_sd"], row["engine"]["nn"])
        out = r.ejecta_dynamics_and_interaction(tg, magnetar_luminosity=lum,
                                                pair_cascade_switch=row["pair_cascade_switch"],
                                                use_gamma_ray_opacity=row["kind"] == "thermalisation",
                                                **row["ejecta"], **row["extra"])
        for got, key in zip(out, ("bolometric_luminosity", "comoving_temperature", "radius", "doppler_factor")):
            np.testing.assert_allclose(got[pick], row[key], rtol=1e-14, err_msg=key)


def test_oracle_reproduces_reference_metzger_magnetar_shells():
    """_general_metzger_magnetar_driven_kilonova_model of the reference's own source, both settings of
    magnetar_heating (make_golden.py::dump_metzger_magnetar)."""
    with open(os.path.join(ROOT, "tests", "golden", "metzger_magnetar_reference_draws.json")) as fh:
        fx = json.load(fh)
    tg = r.get_optimal_time_array(1e-4, 1e7, 300)
    assert tg.size == fx["n_grid"]
    pick = slice(None, None, fx["stride"])
    assert {row["magnetar_heating"] for row in fx["draws"]} == {"first_layer", "all_layers"}
    for row in fx["draws"]:
        e = row["engine"]
        ej = dict(row["ejecta"])
        out = r.metzger_magnetar_driven_internal(
            tg, ej["mej"], ej["vej"], ej["beta"], ej["kappa"], r.magnetar_only(tg, e["l0"], e["tau_sd"], e["nn"]),
            row["use_gamma_ray_opacity"], magnetar_heating=row["magnetar_heating"],
            neutron_precursor_switch=row["neutron_precursor_switch"], pair_cascade_switch=row["pair_cascade_switch"],
            **row["extra"])
        for got, key in zip(out, ("bolometric_luminosity", "temperature", "r_photosphere")):
            np.testing.assert_allclose(got[pick], row[key], rtol=1e-13, err_msg=f"{key} {row['magnetar_heating']}")


@pytest.fixture(scope="module")
def asph_fx():
    with open(os.path.join(ROOT, "tests", "golden", "aspherical_reference_draws.json")) as fh:
        return json.load(fh)


def test_oracle_reproduces_reference_aspherical_diffusion(asph_fx):
    t = np.array(asph_fx["time"])
    for row in asph_fx["draws"]:
        p = row["params"]
        dense = np.linspace(0, t[-1] + 100, 1000)
        lum = r.nickelcobalt_engine(dense, p["f_nickel"], p["mej"])
        tau, out = r.aspherical_diffusion(t, dense, lum, p["kappa"], p["kappa_gamma"], p["mej"], p["vej"],
                                          p["area_projection"], p["area_reference"])
        assert abs(tau / row["tau_d"] - 1) < 1e-15
        np.testing.assert_allclose(out, row["lbol"], rtol=1e-14)


@pytest.mark.gpu
def test_device_reproduces_reference_aspherical_diffusion(asph_fx):
    """interaction_process='AsphericalDiffusion' (interaction_processes.py:66-131) through arnett_bolometric, and the
    same factor inside the flux-density path and the fused likelihood."""
    import redback_b200 as rb
    t = np.array(asph_fx["time"])
    for row in asph_fx["draws"]:
        p = dict(row["params"])
        out = rb.supernova_models.arnett_bolometric(t, interaction_process="AsphericalDiffusion", **p)
        h.assert_close(out, np.array(row["lbol"]), 1e-9, h.diffusion_condition(t, row["tau_d"]), what="aspherical")
    p = dict(asph_fx["draws"][0]["params"])
    nu = np.full(t.size, 4.7e14)
    kw = dict(frequency=nu, output_format="flux_density", interaction_process="AsphericalDiffusion",
              area_projection=p.pop("area_projection"), area_reference=p.pop("area_reference"))
    batch = {k: np.array([v, v * 1.1]) for k, v in p.items()}
    batch.update(redshift=np.array([0.05, 0.1]), temperature_floor=np.array([4000.0, 5000.0]))
    flux = rb.supernova_models.arnett(t, params_batch=batch, **kw)
    plain = rb.supernova_models.arnett(t, params_batch=batch, frequency=nu, output_format="flux_density")
    assert np.all(np.isfinite(flux)) and not np.allclose(flux, plain)
    like = rb.GaussianLikelihood(t, flux[0], 0.1 * flux[0], rb.supernova_models.arnett, kwargs=kw)
    ll = like.log_likelihood_batch(batch)
    want = float(np.sum(-np.log(2 * np.pi * (0.1 * flux[0]) ** 2) / 2))
    assert abs(ll[0] - want) <= 1e-6 + 1e-9 * abs(want)
    with pytest.raises(KeyError):
        rb.supernova_models.arnett_bolometric(t, interaction_process="AsphericalDiffusion",
                                              **{k: v for k, v in row["params"].items() if k != "area_reference"})


@pytest.fixture(scope="module")
def visc_fx():
    with open(os.path.join(ROOT, "tests", "golden", "viscous_reference_draws.json")) as fh:
        return json.load(fh)


def test_oracle_reproduces_reference_viscous(visc_fx):
    """interaction_processes.Viscous (:229-273) and Diffusion's treatment of negative / unsorted epochs (:47, :63), both
    against outputs of the reference's own classes (oracle/make_golden.py::dump_viscous)."""
    t, tn = np.array(visc_fx["time"]), np.array(visc_fx["time_negative_unsorted"])
    for row in visc_fx["draws"]:
        p = row["params"]
        dense = np.linspace(0, t[-1] + 100, 1000)
        lum_ni = r.nickelcobalt_engine(dense, p["f_nickel"], p["mej"])
        lum_mag = r.basic_magnetar(dense * 86400.0, p["p0"], p["bp"], p["mass_ns"], p["theta_pb"])
        for key, lum in (("nickel", lum_ni), ("magnetar", lum_mag)):
            np.testing.assert_allclose(r.viscous(t, dense, lum, p["t_viscous"]), row[key], rtol=1e-14)
            np.testing.assert_allclose(r.viscous(tn, dense, lum, p["t_viscous"]), row[key + "_negative_unsorted"],
                                       rtol=1e-14)
        d = row["diffusion_params"]
        np.testing.assert_allclose(r.diffusion(tn, dense, lum_ni, d["kappa"], d["kappa_gamma"], p["mej"], d["vej"])[1],
                                   row["diffusion_negative_unsorted"], rtol=1e-14)


@pytest.mark.gpu
def test_device_reproduces_reference_viscous(visc_fx):
    """interaction_process='Viscous' through arnett_bolometric / basic_magnetar_powered_bolometric (scalar calls and a
    params_batch), epochs before the dense grid and unsorted epochs through Viscous and Diffusion, the flux-density
    branch and the fused likelihood."""
    import redback_b200 as rb
    t, tn = np.array(visc_fx["time"]), np.array(visc_fx["time_negative_unsorted"])
    rows = visc_fx["draws"]
    for row in rows:
        p = row["params"]
        out = rb.supernova_models.arnett_bolometric(t, p["f_nickel"], p["mej"], interaction_process="Viscous",
                                                    t_viscous=p["t_viscous"])
        np.testing.assert_allclose(out, row["nickel"], rtol=1e-9)
        out = rb.supernova_models.basic_magnetar_powered_bolometric(
            tn, p["p0"], p["bp"], p["mass_ns"], p["theta_pb"], interaction_process="Viscous", t_viscous=p["t_viscous"])
        np.testing.assert_allclose(out, row["magnetar_negative_unsorted"], rtol=1e-9)
        d = row["diffusion_params"]
        out = rb.supernova_models.arnett_bolometric(tn, p["f_nickel"], p["mej"], **d)
        np.testing.assert_allclose(out, row["diffusion_negative_unsorted"], rtol=1e-9)
    batch = {k: np.array([row["params"][k] for row in rows]) for k in ("f_nickel", "mej", "t_viscous")}
    out = rb.supernova_models.arnett_bolometric(tn, params_batch=batch, interaction_process="Viscous")
    np.testing.assert_allclose(out, [row["nickel_negative_unsorted"] for row in rows], rtol=1e-9)
    # flux density + fused likelihood against the oracle restatement (pinned above)
    nu = np.where(np.arange(t.size) % 2 == 0, 6.272e14, 3.813e14)
    p = rows[0]["params"]
    kw = dict(frequency=nu, output_format="flux_density", interaction_process="Viscous", t_viscous=p["t_viscous"],
              vej=7000.0, temperature_floor=4500.0)
    out = rb.supernova_models.basic_magnetar_powered(t, 0.07, p["p0"], p["bp"], p["mass_ns"], p["theta_pb"], **kw)
    ref = r.basic_magnetar_powered_viscous(t, 0.07, p["p0"], p["bp"], p["mass_ns"], p["theta_pb"], frequency=nu,
                                           vej=7000.0, t_viscous=p["t_viscous"], temperature_floor=4500.0)
    np.testing.assert_allclose(out, ref, rtol=1e-9)
    like = rb.GaussianLikelihood(t, ref * 1.02, 0.1 * ref, rb.supernova_models.basic_magnetar_powered, kwargs=kw)
    pb = dict(redshift=np.array([0.07, 0.07]), p0=np.array([p["p0"], 2 * p["p0"]]), bp=np.full(2, p["bp"]),
              mass_ns=np.full(2, p["mass_ns"]), theta_pb=np.full(2, p["theta_pb"]))
    logl = like.log_likelihood_batch(pb)
    ref_l = r.gaussian_likelihood(ref * 1.02, ref, 0.1 * ref)
    assert abs(logl[0] - ref_l) <= 1e-6 + 1e-9 * abs(ref_l)
    with pytest.raises(KeyError):
        rb.supernova_models.arnett_bolometric(t, 0.1, 1.0, interaction_process="Viscous")      # t_viscous missing
