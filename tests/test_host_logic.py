"""CPU: host-side logic of the drop-in layer that needs no GPU -- parameter collection, kwargs-selected spec variants,
host re-parameterisations, the library's host functions."""
import numpy as np
import pytest

from oracle import restated as r
from redback_b200 import _capi, _fused, likelihoods, phase_models
from redback_b200 import magnetar_driven_ejecta_models as mn
from redback_b200 import supernova_models as sn


def test_optimal_time_array_host_function_matches_reference_restatement():
    # utils.get_optimal_time_array without user times (utils.py:79-93): three-segment and single-segment branches
    for args in ((1e-4, 1e8, 500), (1e-4, 1e7, 300), (1e0, 1e8, 2000), (0.1, 3000, 300), (0.1, 500, 300), (1e-2, 7e6, 100)):
        got = _capi.optimal_time_array(*args)
        want = r.get_optimal_time_array(*args)
        assert got.size == want.size, args
        np.testing.assert_allclose(got, want, rtol=4e-16, atol=0, err_msg=str(args))
        assert got[0] == args[0] and got[-1] == args[1]
    with pytest.raises(ValueError):
        _capi.optimal_time_array(0.0, 1.0, 10)


def test_collect_precedence_defaults_and_missing():
    named = dict(redshift=0.1, mej=None)
    kwargs = dict(mej=2.0, vej=5000.0)
    batch = dict(vej=np.array([1.0, 2.0]), kappa=np.array([0.1, 0.2]))
    params, n = _fused.collect(named, kwargs, batch, ("redshift", "mej", "vej", "kappa", "f_nickel"), {"f_nickel": 0})
    assert n == 2 and params["redshift"] == 0.1 and params["mej"] == 2.0 and params["f_nickel"] == 0
    assert params["vej"] is batch["vej"]                        # params_batch wins over kwargs
    with pytest.raises(KeyError):
        _fused.collect(named, kwargs, batch, ("redshift", "temperature_floor"))
    with pytest.raises(ValueError):
        _fused.collect({}, {}, dict(a=np.zeros(2), b=np.zeros(3)), ("a", "b"))


def test_spec_variants_selected_by_kwargs():
    base = sn.FUSED[sn.arnett]
    assert base.for_kwargs({}) is base
    direct = base.for_kwargs(dict(interaction_process=None))
    assert direct is not base and "kappa" not in direct.needed and "vej" in direct.needed
    t0 = phase_models.FUSED[phase_models.t0_base_model].for_kwargs(dict(base_model="slsn"))
    assert t0.needed[-1] == "t0" and "p0" in t0.needed
    with pytest.raises(NotImplementedError):
        phase_models.FUSED[phase_models.t0_base_model].for_kwargs(dict(base_model="tophat_redback"))
    ek = sn.FUSED[sn.thin_shell_supernova].for_kwargs(dict(base_model="slsn_bolometric"))
    assert "ek" in ek.needed and "vej" not in ek.needed and "p0" in ek.needed
    with pytest.raises(ValueError):
        sn.FUSED[sn.homologous_expansion_supernova].for_kwargs(dict(base_model="type_1c_bolometric"))
    trapped = mn.FUSED[mn.trapped_magnetar]
    assert "redshift" in trapped.for_kwargs(dict(output_format="flux", photon_index=2.0)).needed
    assert "redshift" not in trapped.for_kwargs(dict(output_format="luminosity")).needed
    with pytest.raises(KeyError):
        trapped.for_kwargs(dict(output_format="flux"))
    assert set(likelihoods.FUSED) >= {sn.arnett, sn.csm_interaction, mn.basic_mergernova, phase_models.t0_base_model}


def test_host_reparameterisations_match_the_reference_formulas():
    rng = np.random.default_rng(3)
    mej, ek = rng.uniform(1, 20, 5), 10 ** rng.uniform(50, 52, 5)
    for name, thin in (("homologous_expansion_supernova", False), ("thin_shell_supernova", True)):
        spec = sn.FUSED[getattr(sn, name)].for_kwargs({})
        out = spec.transform(dict(mej=mej, ek=ek, f_nickel=0.1))
        assert "ek" not in out
        assert np.array_equal(out["vej"], r.velocity_from_kinetic_energy(mej, ek, thin))      # bit-identical
    gm = sn.FUSED[sn.general_magnetar_driven_supernova].transform(dict(mej=mej, E_sn=ek))
    assert np.array_equal(gm["beta"], np.sqrt(ek / (0.5 * mej * r.solar_mass)) / r.speed_of_light)
    assert gm["ejecta_radius"] == 1.0e11 and gm["n_ism"] == 1.0e-5


def test_ejecta_relations_match_the_reference_formulas():
    """redback_b200.ejecta_relations (host side of the *_ejecta_relation kilonova models) against the oracle's
    independent restatement of redback/ejecta_relations.py, vectorised over prior draws."""
    from redback_b200 import ejecta_relations as ejr
    rng = np.random.default_rng(4)
    m1, m2 = rng.uniform(1.3, 2.0, 32), rng.uniform(1.0, 1.3, 32)
    l1, l2 = rng.uniform(50, 800, 32), rng.uniform(200, 3000, 32)
    for cls, ref in ((ejr.OneComponentBNSNoProjection, r.bns_ejecta_no_projection),
                     (ejr.OneComponentBNSProjection, r.bns_ejecta_projection)):
        rel = cls(m1, m2, l1, l2)
        mej, vej = ref(m1, m2, l1, l2)
        np.testing.assert_allclose(rel.ejecta_mass, mej, rtol=1e-14)
        np.testing.assert_allclose(rel.ejecta_velocity, vej, rtol=1e-14)
    mbh, mns, chi, lam = rng.uniform(3, 10, 32), rng.uniform(1.1, 1.8, 32), rng.uniform(-0.5, 0.9, 32), rng.uniform(100, 2000, 32)
    rel = ejr.OneComponentNSBH(mbh, mns, chi, lam)
    mej, vej = r.nsbh_ejecta(mbh, mns, chi, lam)
    np.testing.assert_allclose(rel.ejecta_mass, mej, rtol=1e-12, atol=1e-15)   # max(a + b + c, 0) of cancelling terms
    np.testing.assert_allclose(rel.ejecta_velocity, vej, rtol=1e-14)
    one = ejr.OneComponentBNSNoProjection(1.4, 1.3, 400.0, 500.0)      # scalars work like the reference's classes
    assert np.ndim(one.ejecta_mass) == 0 and 0 < float(one.ejecta_mass) < 0.1


def test_ejecta_relations_pinned_to_the_reference_classes():
    """The oracle's restatement and the product's host module against the reference's own classes executed from their
    source (tests/golden/ejecta_relations_reference_draws.json, oracle/make_golden.py::dump_ejecta_relations)."""
    import json
    import os
    from redback_b200 import ejecta_relations as ejr
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "ejecta_relations_reference_draws.json")) as fh:
        fx = json.load(fh)
    bns = {k: np.array(v) for k, v in fx["bns"].items()}
    two = {k: np.array(v) for k, v in fx["two"].items()}
    nsbh = {k: np.array(v) for k, v in fx["nsbh"].items()}
    close = lambda a, b: np.testing.assert_allclose(a, np.array(b, dtype=float), rtol=1e-12, atol=1e-15)
    for name, ofn in (("OneComponentBNSNoProjection", r.bns_ejecta_no_projection),
                      ("OneComponentBNSProjection", r.bns_ejecta_projection)):
        mej, vej = ofn(**bns)
        rel = getattr(ejr, name)(**bns)
        for got_m, got_v in ((mej, vej), (rel.ejecta_mass, rel.ejecta_velocity)):
            close(got_m, fx["out"][name]["ejecta_mass"])
            close(got_v, fx["out"][name]["ejecta_velocity"])
    want = fx["out"]["TwoComponentBNS"]
    rel = ejr.TwoComponentBNS(**bns, **two)
    for dyn, wind, vej in (r.bns_two_component_ejecta(**bns, **two), (rel.dynamical_mej, rel.disk_wind_mej, rel.ejecta_velocity)):
        close(dyn, want["dynamical_mej"]); close(wind, want["disk_wind_mej"]); close(vej, want["ejecta_velocity"])
    want = fx["out"]["TwoComponentNSBH"]
    rel = ejr.TwoComponentNSBH(**nsbh, zeta=two["zeta"])
    for dyn, wind, vej in (r.nsbh_two_component_ejecta(**nsbh, zeta=two["zeta"]),
                           (rel.dynamical_mej, rel.disk_wind_mej, rel.ejecta_velocity)):
        close(dyn, want["dynamical_mej"]); close(wind, want["disk_wind_mej"]); close(vej, want["ejecta_velocity"])
    want = fx["out"]["OneComponentNSBH"]
    rel = ejr.OneComponentNSBH(**nsbh)
    for mej, vej in (r.nsbh_ejecta(**nsbh), (rel.ejecta_mass, rel.ejecta_velocity)):
        close(mej, want["ejecta_mass"]); close(vej, want["ejecta_velocity"])


def test_t0_needs_ascending_times():
    phase_models._check_sorted(np.array([1.0, 2.0, 2.0, 3.0]))
    with pytest.raises(NotImplementedError):
        phase_models._check_sorted(np.array([1.0, 3.0, 2.0]))


def test_upper_limit_likelihood_validation_without_gpu():
    x = np.arange(5.0)
    y = np.array([0.5, 1.2, 2.1, 3.5, 4.8])

    def func(x, a, **kwargs):
        return a * x
    like = likelihoods.GaussianLikelihoodWithUpperLimits(x, y, 0.1, func, detections=[1, 1, 1, 0, 0],
                                                         upper_limit_sigma=np.array([2.0, 3.0]))
    np.testing.assert_array_equal(like.get_upper_limit_sigma_values(), [2.0, 3.0])
    level, mag = like._upper_limit_levels()
    np.testing.assert_array_equal(level, [0, 0, 0, 2.0, 3.0])
    assert mag is False and like.summary()["upper_limits"] == 2
    like.data_mode = "mag"
    assert like.data_mode == "magnitude" and like._upper_limit_levels()[1] is True
    with pytest.raises(ValueError):
        like.data_mode = "counts"
    with pytest.raises(ValueError):
        likelihoods.GaussianLikelihoodWithUpperLimits(x, np.array([0.5, 1.2, 2.1, np.nan, 4.8]), 0.1, func,
                                                      detections=[1, 1, 1, 0, 0])
