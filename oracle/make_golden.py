"""Extract the reference's own golden vectors for the hot path into small JSON fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py

Sources (read-only, never executed as code):
  * /root/reference/test/reference_results/all_models_reference.pkl.gz
    (generator: test/reference_results/generate_all_model_ref_data.py:53-143; checked by
    test/test_all_models_regression.py:75 at rtol 1e-5 / atol 1e-12)
  * /root/reference/test/reference_results/tophat_redback_ref_data.ecsv
    (generator: test/reference_results/make_afterglow_ref_data.py:14-26; checked by
    test/test_on_ref_results.py:57-69 with np.allclose defaults)

The pickle is read with a restricted Unpickler: only numpy array reconstruction is allowed and
astropy Quantity objects are mapped onto a plain ndarray subclass, so nothing from the pickle runs.
"""
import gzip
import io
import json
import os
import pickle
import sys

import numpy as np

REF = "/root/reference/test/reference_results"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class _Quantity(np.ndarray):
    def __setstate__(self, state):
        nd_state, _own = state
        super().__setstate__(nd_state)


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __setstate__(self, state):
        pass


class _Restricted(pickle.Unpickler):
    _ALLOWED = {
        ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "_reconstruct"),
        ("numpy", "dtype"),
        ("numpy", "ndarray"),
        ("numpy._core.numeric", "_frombuffer"),
        ("numpy._core.multiarray", "scalar"),
        ("numpy.core.multiarray", "scalar"),
    }

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            mod = __import__(module, fromlist=[name])
            return getattr(mod, name)
        if module.startswith("astropy.units.quantity") and name == "Quantity":
            return _Quantity
        if module.startswith("astropy"):
            return _Inert
        raise pickle.UnpicklingError(f"blocked {module}.{name}")


KEEP = [
    "supernova.arnett_bolometric", "supernova.arnett",
    "supernova.basic_magnetar_powered_bolometric", "supernova.basic_magnetar_powered",
    "supernova.slsn_bolometric", "supernova.slsn", "supernova.type_1a", "supernova.type_1c",
    "supernova.magnetar_nickel",
    "kilonova.one_component_kilonova_model", "kilonova.metzger_kilonova_model",
    "magnetar.basic_magnetar",
    "magnetar.basic_mergernova", "magnetar.general_mergernova", "magnetar.general_mergernova_thermalisation",
    "magnetar.trapped_magnetar", "magnetar.magnetar_only",
    "supernova.general_magnetar_slsn", "supernova.sn_exponential_powerlaw", "supernova.exponential_powerlaw_bolometric",
    "supernova.sn_fallback", "supernova.sn_nickel_fallback", "supernova.homologous_expansion_supernova",
    "supernova.thin_shell_supernova", "tde.tde_analytical", "tde.tde_analytical_bolometric",
    "kilonova.two_component_kilonova_model", "kilonova.three_component_kilonova_model",
    "kilonova.one_component_ejecta_relation", "kilonova.one_component_ejecta_relation_projection",
    "kilonova.one_component_nsbh_ejecta_relation",
    "supernova.sn_fallback", "supernova.sn_nickel_fallback",
    "magnetar.metzger_magnetar_driven_kilonova_model", "magnetar.general_metzger_magnetar_driven",
    "magnetar.general_metzger_magnetar_driven_thermalisation",
    "magnetar.general_mergernova_evolution", "magnetar.general_metzger_magnetar_driven_evolution",
    "supernova.general_magnetar_driven_supernova", "supernova.general_magnetar_driven_supernova_bolometric",
]


def _tolist(x):
    arr = np.asarray(x, dtype=float)
    return [None if not np.isfinite(v) and np.isnan(v) else (repr(float(v)) if not np.isfinite(v) else float(v))
            for v in arr.ravel()]


def dump_pkl():
    with gzip.open(os.path.join(REF, "all_models_reference.pkl.gz"), "rb") as fh:
        data = _Restricted(io.BytesIO(fh.read())).load()
    keys = sorted(data.keys())
    print(len(keys), "keys in pickle")
    out = {}
    for k in keys:
        if k not in KEEP:
            continue
        entry = data[k]
        md = entry["metadata"]
        out[k] = {
            "time": [float(v) for v in np.asarray(md["time_array"], dtype=float)],
            "eval_kwargs": {kk: (float(vv) if isinstance(vv, (int, float, np.floating)) else vv)
                            for kk, vv in md["eval_kwargs"].items()},
            "params": [{pk: float(pv) for pk, pv in p.items()} for p in entry["params"]],
            "results": [_tolist(r) for r in entry["results"]],
        }
        print(k, list(entry["params"][0].keys()))
    with open(os.path.join(OUT, "all_models_subset.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def dump_ecsv():
    rows = []
    header = None
    with open(os.path.join(REF, "tophat_redback_ref_data.ecsv")) as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            if header is None:
                header = line.split()
                continue
            rows.append(line.rstrip("\n"))
    out = []
    for r in rows:
        cells = r.split(" ")
        rec = {}
        for name, cell in zip(header, cells):
            if cell.startswith("["):
                rec[name] = json.loads(cell)
            else:
                try:
                    rec[name] = float(cell)
                except ValueError:
                    rec[name] = cell
        out.append(rec)
    time = out[0]["time"]
    assert all(o["time"] == time for o in out)
    doc = {"time": time, "rows": [{k: v for k, v in o.items() if k != "time"} for o in out]}
    with open(os.path.join(OUT, "tophat_redback_ref.json"), "w") as fh:
        json.dump(doc, fh)
    print("ecsv rows:", len(out), "epochs:", len(time))


def dump_verbatim():
    """Outputs of the reference's OWN source files, imported verbatim through oracle/refshim.py (numba JIT
    disabled = the pure-Python path the reference's CI tests, .github/workflows/python-app.yml:88), on seeded
    prior draws: interaction_processes.Diffusion + photosphere.TemperatureFloor, and the redback-native afterglows."""
    os.environ["NUMBA_DISABLE_JIT"] = "1"
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import warnings
    warnings.simplefilter("ignore")
    from oracle import refshim, restated as r
    ns = refshim.load()
    bm = refshim.load_afterglow()
    rng = np.random.default_rng(2024)
    doc = {"diffusion": [], "afterglow": []}
    t = np.sort(rng.uniform(0.5, 300, 40))
    for i in range(16):
        prm = dict(mej=float(10 ** rng.uniform(-1, 2)), vej=float(10 ** rng.uniform(3, 5)),
                   kappa=float(rng.uniform(0.05, 2)), kappa_gamma=float(10 ** rng.uniform(-4, 4)))
        dense = np.linspace(0, t[-1] + 100, 1000)
        if i % 2 == 0:
            eng = dict(engine="nickel", f_nickel=float(10 ** rng.uniform(-3, 0)))
            lum = r.nickelcobalt_engine(dense, eng["f_nickel"], prm["mej"])
        else:
            eng = dict(engine="magnetar", p0=float(rng.uniform(1, 10)), bp=float(10 ** rng.uniform(-1, 1)),
                       mass_ns=float(rng.uniform(1.1, 2.2)), theta_pb=float(rng.uniform(0, 1.57)))
            lum = r.basic_magnetar(dense * 86400, eng["p0"], eng["bp"], eng["mass_ns"], eng["theta_pb"])
        d = ns.interaction_processes.Diffusion(t, dense, lum, prm["kappa"], prm["kappa_gamma"], prm["mej"], prm["vej"])
        tfl = float(10 ** rng.uniform(3, 4.5))
        ph = ns.photosphere.TemperatureFloor(t, d.new_luminosity, prm["vej"], tfl)
        doc["diffusion"].append(dict(params=dict(prm, temperature_floor=tfl, **eng), tau_d=float(d.tau_d),
                                     lbol=[float(v) for v in d.new_luminosity],
                                     temperature=[float(v) for v in ph.photosphere_temperature],
                                     radius=[float(v) for v in ph.r_photosphere]))
    doc["diffusion_time"] = [float(v) for v in t]
    ta = np.geomspace(0.1, 300, 24)
    fr = np.where(np.arange(24) % 2 == 0, 5e9, 2e17)
    models = [("tophat_redback", {}), ("gaussian_redback", {}), ("twocomponent_redback", {}), ("powerlaw_redback", {}),
              ("alternativepowerlaw_redback", {}), ("doublegaussian_redback", {})]
    for j in range(12):
        name, _ = models[j % len(models)]
        kw = dict(thv=float(np.arccos(rng.uniform(0.6, 1))), loge0=float(rng.uniform(48, 53)),
                  thc=float(rng.uniform(0.03, 0.1)), logn0=float(rng.uniform(-3, 1)), p=float(rng.uniform(1.6, 3)),
                  logepse=float(rng.uniform(-3, -0.5)), logepsb=float(rng.uniform(-4, -1)), g0=float(rng.uniform(100, 600)),
                  xiN=float(rng.uniform(0.1, 1)))
        if j < 3:
            kw["thv"] = kw["thc"] * float(rng.uniform(0, 1.2))     # on-axis: resolution above 10
        if name != "tophat_redback":
            kw["thj"] = kw["thc"] * float(rng.uniform(1.5, 3))
        z = float(rng.uniform(0.05, 2))
        k = 2 if j % 4 == 3 else 0
        out = getattr(bm, name)(ta, z, frequency=fr, output_format="flux_density", k=k, **kw)
        doc["afterglow"].append(dict(model=name, redshift=z, k=k, params=kw, results=[float(v) for v in out]))
    doc["afterglow_time"] = [float(v) for v in ta]
    doc["afterglow_frequency"] = [float(v) for v in fr]
    with open(os.path.join(OUT, "verbatim_reference_draws.json"), "w") as fh:
        json.dump(doc, fh)
    print("verbatim fixtures:", len(doc["diffusion"]), "diffusion,", len(doc["afterglow"]), "afterglow")


def dump_aspherical():
    """interaction_processes.AsphericalDiffusion of the reference's own source on seeded draws."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import warnings
    warnings.simplefilter("ignore")
    from oracle import refshim, restated as r
    ns = refshim.load()
    rng = np.random.default_rng(2027)
    t = np.sort(rng.uniform(0.5, 250, 30))
    doc = {"time": [float(v) for v in t], "draws": []}
    for i in range(8):
        prm = dict(mej=float(10 ** rng.uniform(-1, 1.5)), vej=float(10 ** rng.uniform(3.3, 4.5)),
                   kappa=float(rng.uniform(0.05, 1)), kappa_gamma=float(10 ** rng.uniform(-2, 2)),
                   f_nickel=float(10 ** rng.uniform(-2, 0)), area_projection=float(rng.uniform(0.3, 2.0)),
                   area_reference=float(rng.uniform(0.5, 1.5)))
        dense = np.linspace(0, t[-1] + 100, 1000)
        lum = r.nickelcobalt_engine(dense, prm["f_nickel"], prm["mej"])
        d = ns.interaction_processes.AsphericalDiffusion(t, dense, lum, prm["kappa"], prm["kappa_gamma"], prm["mej"],
                                                         prm["vej"], prm["area_projection"], prm["area_reference"])
        doc["draws"].append(dict(params=prm, tau_d=float(d.tau_d), lbol=[float(v) for v in d.new_luminosity]))
    with open(os.path.join(OUT, "aspherical_reference_draws.json"), "w") as fh:
        json.dump(doc, fh)
    print("aspherical fixtures:", len(doc["draws"]))


def dump_viscous():
    """interaction_processes.Viscous of the reference's own source on seeded draws: nickel-cobalt and magnetar engines on
    the dense grid of the calling model; includes an epoch before the grid and unsorted epochs."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import warnings
    warnings.simplefilter("ignore")
    from oracle import refshim, restated as r
    ns = refshim.load()
    rng = np.random.default_rng(2031)
    t = np.sort(rng.uniform(0.5, 250, 30))
    t_neg = np.concatenate(([-3.0, -0.5], t[:10][::-1], t[10:]))      # negative and unsorted epochs; time[-1] unchanged
    doc = {"time": [float(v) for v in t], "time_negative_unsorted": [float(v) for v in t_neg], "draws": []}
    for i in range(8):
        prm = dict(mej=float(10 ** rng.uniform(-1, 1.5)), f_nickel=float(10 ** rng.uniform(-2, 0)),
                   t_viscous=float(10 ** rng.uniform(-2, 2)), p0=float(rng.uniform(1, 10)),
                   bp=float(10 ** rng.uniform(-1, 1)), mass_ns=float(rng.uniform(1.1, 2.2)),
                   theta_pb=float(rng.uniform(0.05, 1.57)))
        dense = np.linspace(0, t[-1] + 100, 1000)
        lum_ni = r.nickelcobalt_engine(dense, prm["f_nickel"], prm["mej"])
        lum_mag = r.basic_magnetar(dense * 86400.0, prm["p0"], prm["bp"], prm["mass_ns"], prm["theta_pb"])
        row = dict(params=prm)
        for key, lum in (("nickel", lum_ni), ("magnetar", lum_mag)):
            row[key] = [float(v) for v in ns.interaction_processes.Viscous(t, dense, lum, prm["t_viscous"]).new_luminosity]
            row[key + "_negative_unsorted"] = [float(v) for v in ns.interaction_processes.Viscous(
                t_neg, dense, lum, prm["t_viscous"]).new_luminosity]
        # Diffusion with the same awkward epochs (interaction_processes.py:47, :63)
        dprm = dict(kappa=float(rng.uniform(0.05, 1)), kappa_gamma=float(10 ** rng.uniform(-2, 2)),
                    vej=float(10 ** rng.uniform(3.3, 4.5)))
        row["diffusion_params"] = dprm
        row["diffusion_negative_unsorted"] = [float(v) for v in ns.interaction_processes.Diffusion(
            t_neg, dense, lum_ni, dprm["kappa"], dprm["kappa_gamma"], prm["mej"], dprm["vej"]).new_luminosity]
        doc["draws"].append(row)
    with open(os.path.join(OUT, "viscous_reference_draws.json"), "w") as fh:
        json.dump(doc, fh)
    print("viscous fixtures:", len(doc["draws"]))


def dump_csm():
    """_csm_engine + CSMDiffusion (+ TemperatureFloor) of the reference's own sources on seeded prior draws
    (priors/csm_interaction.prior ranges)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import warnings
    warnings.simplefilter("ignore")
    from oracle import refshim
    ns = refshim.load_csm()
    rng = np.random.default_rng(2025)
    t = np.sort(np.concatenate([rng.uniform(0.03, 5, 6), rng.uniform(5, 400, 34)]))
    doc = {"time": [float(v) for v in t], "draws": []}
    for i in range(16):
        prm = dict(mej=float(10 ** rng.uniform(0, 2)), csm_mass=float(10 ** rng.uniform(0, 2)),
                   vej=float(10 ** rng.uniform(3, 5)), eta=float(rng.uniform(0, 2)),
                   rho=float(10 ** rng.uniform(-14, -12)), kappa=float(rng.uniform(0.05, 2)),
                   r0=float(rng.uniform(50, 700)))
        tt = t if i % 4 else t[6:]          # with and without epochs below the 0.1 d default grid start
        lbol = ns.csm_interaction_bolometric(tt, **prm)
        eng = ns._csm_engine(time=tt, **prm)
        tfl = float(10 ** rng.uniform(2, 4))
        ph = ns.photosphere.TemperatureFloor(tt, lbol, prm["vej"], tfl)
        extra = {}
        if i < 6:
            # the legacy quadrature method and csm_nickel (both methods) on the first draws
            extra["lbol_quadrature"] = [float(v) for v in ns.csm_interaction_bolometric(
                tt, csm_diffusion_method="quadrature", **prm)]
            nprm = dict(f_nickel=float(10 ** rng.uniform(-2, 0)), ek=float(10 ** rng.uniform(50.5, 52)),
                        kappa_gamma=float(10 ** rng.uniform(-2, 2)))
            cprm = {k: v for k, v in prm.items() if k != "vej"}
            extra["nickel_params"] = nprm
            extra["csm_nickel_lbol"] = [float(v) for v in ns.csm_nickel_bolometric(tt, **cprm, **nprm)]
            extra["csm_nickel_lbol_quadrature"] = [float(v) for v in ns.csm_nickel_bolometric(
                tt, csm_diffusion_method="legacy", csm_diffusion_timesteps=1000, **cprm, **nprm)]
        doc["draws"].append(dict(params=dict(prm, temperature_floor=tfl), skip=0 if i % 4 else 6, **extra,
                                 r_photosphere=float(eng.r_photosphere),
                                 mass_csm_threshold=float(eng.mass_csm_threshold),
                                 engine_lbol=[float(v) for v in eng.lbol], lbol=[float(v) for v in lbol],
                                 temperature=[float(v) for v in ph.photosphere_temperature],
                                 radius=[float(v) for v in ph.r_photosphere]))
    with open(os.path.join(OUT, "csm_reference_draws.json"), "w") as fh:
        json.dump(doc, fh)
    print("csm fixtures:", len(doc["draws"]))


def dump_mergernova():
    """_ejecta_dynamics_and_interaction of the reference's own source on seeded prior draws
    (priors/basic_mergernova.prior, general_mergernova(_thermalisation).prior); every 7th grid point is stored."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import warnings
    warnings.simplefilter("ignore")
    from oracle import refshim
    ns = refshim.load_mergernova()
    rng = np.random.default_rng(2026)
    tg = ns.get_optimal_time_array(1e-4, 1e8, 500)
    doc = {"stride": 7, "n_grid": int(tg.size), "draws": []}
    for i in range(12):
        ej = dict(mej=float(rng.uniform(1e-4, 0.1)), beta=float(rng.uniform(0.1, 0.7)),
                  ejecta_radius=float(10 ** rng.uniform(9, 10)), kappa=float(rng.uniform(1, 30)),
                  n_ism=float(10 ** rng.uniform(-4, 0)))
        kind = ("basic", "general", "thermalisation")[i % 3]
        if kind == "basic":
            eng = dict(p0=float(10 ** rng.uniform(np.log10(0.7), 4)), bp=float(10 ** rng.uniform(-4, 4)),
                       mass_ns=float(rng.uniform(1.1, 2.2)), theta_pb=float(rng.uniform(0.05, 1.57)))
            lum = ns.basic_magnetar(tg, **eng)
        else:
            eng = dict(l0=float(10 ** rng.uniform(40, 50)), tau_sd=float(10 ** rng.uniform(1, 5)),
                       nn=float(rng.uniform(2, 7)))
            lum = ns.magnetar_only(tg, eng["l0"], eng["tau_sd"], eng["nn"])
        extra = dict(kappa_gamma=float(10 ** rng.uniform(-4, 4))) if kind == "thermalisation" else \
            dict(thermalisation_efficiency=float(rng.uniform(0.1, 1)))
        pcs = bool(i % 2)
        out = ns._ejecta_dynamics_and_interaction(time=tg, magnetar_luminosity=lum, pair_cascade_switch=pcs,
                                                  use_gamma_ray_opacity=kind == "thermalisation", **ej, **extra)
        pick = slice(None, None, 7)
        doc["draws"].append(dict(kind=kind, ejecta=ej, engine=eng, extra=extra, pair_cascade_switch=pcs,
                                 bolometric_luminosity=[float(v) for v in out.bolometric_luminosity[pick]],
                                 comoving_temperature=[float(v) for v in out.comoving_temperature[pick]],
                                 radius=[float(v) for v in out.radius[pick]],
                                 doppler_factor=[float(v) for v in out.doppler_factor[pick]]))
    with open(os.path.join(OUT, "mergernova_reference_draws.json"), "w") as fh:
        json.dump(doc, fh)
    print("mergernova fixtures:", len(doc["draws"]))


def dump_metzger_magnetar():
    """_general_metzger_magnetar_driven_kilonova_model of the reference's own source with both settings of
    magnetar_heating on seeded draws (priors/general_metzger_magnetar_driven*.prior ranges); every 5th grid point."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import warnings
    warnings.simplefilter("ignore")
    from oracle import refshim
    ns = refshim.load_metzger_magnetar()
    rng = np.random.default_rng(2027)
    tg = ns.get_optimal_time_array(1e-4, 1e7, 300)
    doc = {"stride": 5, "n_grid": int(tg.size), "draws": []}
    for i in range(8):
        ej = dict(mej=float(rng.uniform(1e-3, 0.05)), vej=float(rng.uniform(0.05, 0.3)), beta=float(rng.uniform(1.5, 8)),
                  kappa=float(rng.uniform(1, 30)))
        eng = dict(l0=float(10 ** rng.uniform(42, 48)), tau_sd=float(10 ** rng.uniform(2, 5)), nn=float(rng.uniform(2, 6)))
        gamma = bool(i % 2)
        extra = dict(kappa_gamma=float(10 ** rng.uniform(-2, 2))) if gamma else \
            dict(thermalisation_efficiency=float(rng.uniform(0.1, 1)))
        heating = ("first_layer", "all_layers")[(i // 2) % 2]
        lum = ns.magnetar_only(tg, eng["l0"], eng["tau_sd"], eng["nn"])
        out = ns._general_metzger_magnetar_driven_kilonova_model(
            time=tg, magnetar_luminosity=lum, use_gamma_ray_opacity=gamma, magnetar_heating=heating,
            neutron_precursor_switch=bool(i % 3), pair_cascade_switch=bool((i + 1) % 2), **ej, **extra)
        pick = slice(None, None, 5)
        doc["draws"].append(dict(ejecta=ej, engine=eng, extra=extra, use_gamma_ray_opacity=gamma,
                                 magnetar_heating=heating, neutron_precursor_switch=bool(i % 3),
                                 pair_cascade_switch=bool((i + 1) % 2),
                                 bolometric_luminosity=[float(v) for v in out.bolometric_luminosity[pick]],
                                 temperature=[float(v) for v in out.temperature[pick]],
                                 r_photosphere=[float(v) for v in out.r_photosphere[pick]]))
    with open(os.path.join(OUT, "metzger_magnetar_reference_draws.json"), "w") as fh:
        json.dump(doc, fh)
    print("metzger magnetar fixtures:", len(doc["draws"]))


def dump_ejecta_relations():
    """The reference's own ejecta-relation classes (redback/ejecta_relations.py) on seeded draws."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import refshim
    ns = refshim.load_ejecta_relations()
    rng = np.random.default_rng(2028)
    n = 24
    bns = dict(mass_1=rng.uniform(1.3, 2.0, n), mass_2=rng.uniform(1.0, 1.3, n), lambda_1=rng.uniform(50, 800, n),
               lambda_2=rng.uniform(200, 3000, n))
    two = dict(mtov=rng.uniform(2.0, 2.4, n), zeta=rng.uniform(0.05, 0.5, n))
    nsbh = dict(mass_bh=rng.uniform(3, 10, n), mass_ns=rng.uniform(1.1, 1.8, n), chi_bh=rng.uniform(-0.5, 0.9, n),
                lambda_ns=rng.uniform(100, 2000, n))
    doc = {"bns": {k: v.tolist() for k, v in bns.items()}, "two": {k: v.tolist() for k, v in two.items()},
           "nsbh": {k: v.tolist() for k, v in nsbh.items()}, "out": {}}
    with np.errstate(all="ignore"):
        for name in ("OneComponentBNSNoProjection", "OneComponentBNSProjection"):
            rel = getattr(ns, name)(**bns)
            doc["out"][name] = dict(ejecta_mass=np.asarray(rel.ejecta_mass).tolist(),
                                    ejecta_velocity=np.asarray(rel.ejecta_velocity).tolist())
        rel = ns.TwoComponentBNS(**bns, **two)
        doc["out"]["TwoComponentBNS"] = dict(dynamical_mej=np.asarray(rel.dynamical_mej).tolist(),
                                             disk_wind_mej=np.asarray(rel.disk_wind_mej).tolist(),
                                             ejecta_velocity=np.asarray(rel.ejecta_velocity).tolist())
        rel = ns.TwoComponentNSBH(**nsbh, zeta=two["zeta"])
        doc["out"]["TwoComponentNSBH"] = dict(dynamical_mej=np.asarray(rel.dynamical_mej).tolist(),
                                              disk_wind_mej=np.asarray(rel.disk_wind_mej).tolist(),
                                              ejecta_velocity=np.asarray(rel.ejecta_velocity).tolist())
        rel = ns.OneComponentNSBH(**nsbh)
        doc["out"]["OneComponentNSBH"] = dict(ejecta_mass=np.asarray(rel.ejecta_mass).tolist(),
                                              ejecta_velocity=np.asarray(rel.ejecta_velocity).tolist())
    with open(os.path.join(OUT, "ejecta_relations_reference_draws.json"), "w") as fh:
        json.dump(doc, fh)
    print("ejecta relation fixtures:", n)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    dump_pkl()
    dump_ecsv()
    dump_verbatim()
    dump_aspherical()
    dump_viscous()
    dump_csm()
    dump_mergernova()
    dump_metzger_magnetar()
    dump_ejecta_relations()
