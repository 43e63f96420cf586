"""ctypes binding of the C ABI in include/redback_b200.h.

The shared library is the product; there is no CPU fallback.  If it is missing it is built in-tree
with nvcc; if that is impossible, importing any compute entry point raises.
"""
import ctypes as C
import os

from . import _build

RB_OK, RB_ERR_INVALID, RB_ERR_UNSUPPORTED, RB_ERR_CUDA = 0, -1, -2, -3
ENGINE_NICKEL, ENGINE_MAGNETAR, ENGINE_CSM, ENGINE_MAGNETAR_NICKEL = 0, 1, 2, 3
ENGINE_MAGNETAR_ONLY, ENGINE_EXP_POWERLAW, ENGINE_TDE_FALLBACK, ENGINE_FALLBACK, ENGINE_FALLBACK_NICKEL = 4, 5, 6, 7, 8
PREC_F64, PREC_F32 = 0, 1
NOISE_TYPES = dict(limiting_mag=0, gaussian=1, gaussianmodel=2)
SED_BOLOMETRIC, SED_BLACKBODY, SED_CUTOFF_BLACKBODY, SED_CUTOFF_LINE = 0, 1, 2, 3
SIGMA_FIXED, SIGMA_SAMPLED, SIGMA_QUADRATURE, SIGMA_FRACTIONAL, SIGMA_SYSTEMATIC = 0, 1, 2, 3, 4

SN_ROWS = dict(redshift=0, f_nickel=1, p0=1, bp=2, mass_ns=3, theta_pb=4, mej=5, vej=6, kappa=7,
               kappa_gamma=8, temperature_floor=9,
               csm_mass=1, eta=2, rho=3, r0=4,     # RB_ENGINE_CSM reuses the engine rows
               l0=1, tsd=2, nn=3,                  # RB_ENGINE_MAGNETAR_ONLY
               lbol_0=1, alpha_1=2, alpha_2=3, tpeak_d=4,   # RB_ENGINE_EXP_POWERLAW
               t_0_turn=2,                         # RB_ENGINE_TDE_FALLBACK (l0 in row 1)
               logl1=1, tr=2)                      # RB_ENGINE_FALLBACK(_NICKEL); f_nickel of the latter: SN_ROWS_FALLBACK
SN_NPARAM = 10


class Cosmology(C.Structure):
    _fields_ = [("hubble_distance_cm", C.c_double), ("Om0", C.c_double), ("Ogamma0", C.c_double),
                ("Ode0", C.c_double), ("neff_per_nu", C.c_double), ("nu_y", C.c_double * 3),
                ("n_massive", C.c_int), ("n_massless", C.c_int)]


class UpperLimits(C.Structure):
    _fields_ = [("sigma_level", C.c_void_p), ("magnitude_mode", C.c_int)]


class SnConfig(C.Structure):
    _fields_ = [("engine", C.c_int), ("sed", C.c_int), ("sigma_mode", C.c_int), ("dense_resolution", C.c_int),
                ("cutoff_wavelength", C.c_double), ("absorption_index", C.c_double), ("cosmology", Cosmology),
                ("line_wavelength", C.c_double), ("line_width", C.c_double), ("line_time", C.c_double),
                ("line_duration", C.c_double), ("line_amplitude", C.c_double), ("precision", C.c_int),
                ("upper_limits", UpperLimits), ("csm_nn", C.c_double), ("csm_delta", C.c_double),
                ("csm_efficiency", C.c_double), ("t0", C.c_void_p), ("f_nickel", C.c_void_p),
                ("no_interaction", C.c_int), ("aspherical", C.c_int), ("area_ratio", C.c_double),
                ("viscous", C.c_int), ("t_first_valid", C.c_double), ("csm_quadrature", C.c_int),
                ("csm_timesteps", C.c_int), ("negative_epochs", C.c_int)]


EXT_CCM89, EXT_ODONNELL94, EXT_CALZETTI00, EXT_FITZPATRICK99 = 1, 2, 3, 4


class Extinction(C.Structure):
    """rb_extinction (include/redback_b200.h)"""
    _fields_ = [("host_law", C.c_int), ("mw_law", C.c_int), ("rv_host", C.c_double), ("rv_mw", C.c_double),
                ("av_mw", C.c_double), ("f99_host", C.c_double * 41), ("f99_mw", C.c_double * 41)]


KN_ONE_COMPONENT, KN_METZGER = 0, 1
KN_ROWS = dict(redshift=0, mej=1, vej=2, kappa=3, temperature_floor=4, beta=4)
KN_NPARAM = 5


class KnConfig(C.Structure):
    _fields_ = [("model", C.c_int), ("sigma_mode", C.c_int), ("dense_resolution", C.c_int),
                ("neutron_precursor_switch", C.c_int), ("vmax", C.c_double), ("cosmology", Cosmology),
                ("upper_limits", UpperLimits), ("grid_t_max", C.c_double), ("t0", C.c_void_p), ("precision", C.c_int)]


AG_TOPHAT, AG_GAUSSIAN, AG_POWERLAW, AG_POWERLAW_ALT, AG_TWO_COMPONENT, AG_DOUBLE_GAUSSIAN = 1, 2, 3, 4, 5, 6
AG_ROWS = dict(redshift=0, thv=1, loge0=2, thc=3, thj=4, logn0=5, p=6, logepse=7, logepsb=8, g0=9, xiN=10)
AG_NPARAM = 11


class AgConfig(C.Structure):
    _fields_ = [("jet", C.c_int), ("sigma_mode", C.c_int), ("steps", C.c_int), ("res", C.c_int), ("k", C.c_int),
                ("expansion", C.c_int), ("a1", C.c_double), ("cosmology", Cosmology), ("ss", C.c_double),
                ("aa", C.c_double), ("upper_limits", UpperLimits), ("ab_magnitude", C.c_int), ("precision", C.c_int)]


MN_BASIC_MAGNETAR, MN_MAGNETAR_ONLY, MN_EVOLVING = 0, 1, 2
MN_OUT_FLUX_DENSITY, MN_OUT_TRAPPED_LUMINOSITY, MN_OUT_TRAPPED_FLUX, MN_OUT_SUPERNOVA, MN_OUT_LBOL = 0, 1, 2, 3, 4
MN_ROWS = dict(redshift=0, mej=1, beta=2, ejecta_radius=3, kappa=4, n_ism=5, p0=6, l0=6, bp=7, tau_sd=7, mass_ns=8,
               nn=8, theta_pb=9, thermalisation_efficiency=10, kappa_gamma=10)
MN_NPARAM = 13
# RB_MN_EVOLVING reuses rows 6-9 and adds two
# general_magnetar_driven_supernova: magnetar_only rows + f_nickel / temperature_floor (RB_MN_F_NICKEL, ..._FLOOR)
MN_ROWS_SUPERNOVA = dict(redshift=0, mej=1, beta=2, ejecta_radius=3, kappa=4, n_ism=5, l0=6, tau_sd=7, nn=8, f_nickel=9,
                         kappa_gamma=10, temperature_floor=11)
MN_ROWS_EVOLVING = dict(redshift=0, mej=1, beta=2, ejecta_radius=3, kappa=4, n_ism=5, logbint=6, logbext=7, p0=8, chi0=9,
                        kappa_gamma=10, radius=11, logmoi=12)


class MnConfig(C.Structure):
    _fields_ = [("engine", C.c_int), ("sigma_mode", C.c_int), ("use_gamma_ray_opacity", C.c_int),
                ("pair_cascade_switch", C.c_int), ("ejecta_albedo", C.c_double), ("pair_cascade_fraction", C.c_double),
                ("resolution", C.c_int), ("cosmology", Cosmology), ("upper_limits", UpperLimits), ("output", C.c_int),
                ("photon_index", C.c_double), ("use_r_process", C.c_int), ("grid_t_min", C.c_double),
                ("grid_t_max", C.c_double), ("sed", C.c_int), ("cutoff_wavelength", C.c_double),
                ("absorption_index", C.c_double), ("t0", C.c_void_p)]


MK_ROWS = dict(redshift=0, mej=1, vej=2, beta=3, kappa_r=4, p0=5, l0=5, bp=6, tau_sd=6, mass_ns=7, nn=7, theta_pb=8,
               thermalisation_efficiency=9, kappa_gamma=9)
MK_NPARAM = 12
MK_ROWS_EVOLVING = dict(redshift=0, mej=1, vej=2, beta=3, kappa_r=4, logbint=5, logbext=6, p0=7, chi0=8, kappa_gamma=9,
                        radius=10, logmoi=11)


class MkConfig(C.Structure):
    _fields_ = [("engine", C.c_int), ("sigma_mode", C.c_int), ("use_gamma_ray_opacity", C.c_int),
                ("pair_cascade_switch", C.c_int), ("ejecta_albedo", C.c_double), ("pair_cascade_fraction", C.c_double),
                ("neutron_precursor_switch", C.c_int), ("vmax", C.c_double), ("resolution", C.c_int),
                ("cosmology", Cosmology), ("upper_limits", UpperLimits), ("heat_all_layers", C.c_int), ("t0", C.c_void_p)]


class RedbackB200Error(RuntimeError):
    pass


_lib = None
_dp = C.c_void_p  # device or host pointer to double

EXPORTS = {
    "rb_last_error": (C.c_char_p, []),
    "rb_version": (C.c_char_p, []),
    "rb_source_hash": (C.c_char_p, []),
    "rb_launch_count": (C.c_int64, []),
    "rb_phot_mode_counts": (C.c_int, [C.POINTER(C.c_uint64), C.c_int]),
    "rb_cosmology_flat_lcdm": (C.c_int, [C.c_double, C.c_double, C.c_double, C.c_double,
                                         C.POINTER(C.c_double), C.POINTER(Cosmology)]),
    "rb_cosmology_planck18": (C.c_int, [C.POINTER(Cosmology)]),
    "rb_sn_config_default": (C.c_int, [C.POINTER(SnConfig)]),
    "rb_sn_eval_f64": (C.c_int, [C.POINTER(SnConfig), C.c_int64, C.c_int64] + [_dp] * 9 + [C.c_void_p]),
    "rb_sn_eval_f64_host": (C.c_int, [C.POINTER(SnConfig), C.c_int64, C.c_int64] + [_dp] * 9),
    "rb_sn_eval_f64_host_rows": (C.c_int, [C.POINTER(SnConfig), C.c_int64, C.c_int64, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_double)] + [_dp] * 8),
    "rb_kn_config_default": (C.c_int, [C.POINTER(KnConfig)]),
    "rb_kn_eval_f64": (C.c_int, [C.POINTER(KnConfig), C.c_int64, C.c_int64] + [_dp] * 9 + [C.c_void_p]),
    "rb_kn_eval_f64_host": (C.c_int, [C.POINTER(KnConfig), C.c_int64, C.c_int64] + [_dp] * 9),
    "rb_kn_eval_f64_host_rows": (C.c_int, [C.POINTER(KnConfig), C.c_int64, C.c_int64, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_double)] + [_dp] * 8),
    "rb_ag_config_default": (C.c_int, [C.POINTER(AgConfig)]),
    "rb_ag_eval_f64": (C.c_int, [C.POINTER(AgConfig), C.c_int64, C.c_int64] + [_dp] * 9 + [C.c_void_p]),
    "rb_ag_eval_f64_host": (C.c_int, [C.POINTER(AgConfig), C.c_int64, C.c_int64] + [_dp] * 9),
    "rb_ag_eval_f64_host_rows": (C.c_int, [C.POINTER(AgConfig), C.c_int64, C.c_int64, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_double)] + [_dp] * 8),
    "rb_mn_config_default": (C.c_int, [C.POINTER(MnConfig)]),
    "rb_mn_eval_f64": (C.c_int, [C.POINTER(MnConfig), C.c_int64, C.c_int64] + [_dp] * 9 + [C.c_void_p]),
    "rb_mn_eval_f64_host": (C.c_int, [C.POINTER(MnConfig), C.c_int64, C.c_int64] + [_dp] * 9),
    "rb_mn_eval_f64_host_rows": (C.c_int, [C.POINTER(MnConfig), C.c_int64, C.c_int64, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_double)] + [_dp] * 8),
    "rb_optimal_time_array": (C.c_int, [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "rb_mk_config_default": (C.c_int, [C.POINTER(MkConfig)]),
    "rb_mk_eval_f64": (C.c_int, [C.POINTER(MkConfig), C.c_int64, C.c_int64] + [_dp] * 9 + [C.c_void_p]),
    "rb_mk_eval_f64_host": (C.c_int, [C.POINTER(MkConfig), C.c_int64, C.c_int64] + [_dp] * 9),
    "rb_mk_eval_f64_host_rows": (C.c_int, [C.POINTER(MkConfig), C.c_int64, C.c_int64, C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_double)] + [_dp] * 8),
    "rb_sn_mag_eval_f64": (C.c_int, [C.POINTER(SnConfig), C.c_void_p, C.c_int64, C.c_int64] + [_dp] * 8 + [C.c_void_p]),
    "rb_kn_mag_eval_f64": (C.c_int, [C.POINTER(KnConfig), C.c_void_p, C.c_int64, C.c_int64] + [_dp] * 8 + [C.c_void_p]),
    "rb_kn_multi_mag_eval_f64": (C.c_int, [C.POINTER(KnConfig), C.c_void_p, C.c_int, C.c_int64, C.c_int64] + [_dp] * 8
                                 + [C.c_void_p]),
    "rb_mn_mag_eval_f64": (C.c_int, [C.POINTER(MnConfig), C.c_void_p, C.c_int64, C.c_int64] + [_dp] * 8 + [C.c_void_p]),
    "rb_mk_mag_eval_f64": (C.c_int, [C.POINTER(MkConfig), C.c_void_p, C.c_int64, C.c_int64] + [_dp] * 8 + [C.c_void_p]),
    "rb_optical_noise_f64": (C.c_int, [C.c_int, C.c_int64, C.c_int64] + [_dp] * 4 + [C.c_double, C.c_int] + [_dp] * 6
                             + [C.c_void_p]),
    "rb_optical_noise_philox_f64": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_uint64] + [_dp] * 3
                                    + [C.c_double, C.c_int] + [_dp] * 7 + [C.c_void_p]),
    "rb_sn_mag_eval_ragged_f64": (C.c_int, [C.POINTER(SnConfig), C.c_void_p, C.c_int64, C.c_int64] + [_dp] * 6
                                  + [C.c_void_p]),
    "rb_survey_noise_philox_f64": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_uint64] + [_dp] * 7
                                   + [C.c_double, C.c_double] + [_dp] * 6 + [C.c_void_p]),
    "rb_luminosity_distance_f64": (C.c_int, [C.POINTER(Cosmology), C.c_int64, _dp, _dp, C.c_void_p]),
    "rb_gaussian_logl_f64": (C.c_int, [C.c_int, C.c_int64, C.c_int64] + [_dp] * 4 + [C.POINTER(UpperLimits), _dp,
                                       C.c_void_p]),
    "rb_mixture_logl_f64": (C.c_int, [C.c_int64, C.c_int64] + [_dp] * 6 + [C.c_void_p]),
    "rb_sizeof": (C.c_int64, [C.c_int]),
    "rb_extinction_apply_f64": (C.c_int, [C.POINTER(Extinction), C.c_int64, C.c_int64, _dp, _dp, _dp, _dp, C.c_void_p]),
    "rb_synchrotron_add_f64": (C.c_int, [C.c_int64, C.c_int64, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, _dp,
                                         C.c_void_p]),
    "rb_measure_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}


def lib():
    """Load (building if necessary) libredback_b200.so and declare its signatures."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("REDBACK_B200_LIB")   # alternate build of the same ABI (kernel experiments)
    if not path:
        path = _build.build()       # no-op when the library's compiled-in source digest matches the sources
    handle = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        if not hasattr(handle, name) and os.environ.get("REDBACK_B200_LIB"):
            continue                # an older experimental build may lack newer entry points
        fn = getattr(handle, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if not os.environ.get("REDBACK_B200_LIB"):
        got = handle.rb_source_hash().decode().replace("RBHASH:", "")
        if got != _build.source_hash():
            raise RedbackB200Error(f"{path} was built from other sources (digest {got}); rebuild with "
                                   "python -m redback_b200._build --force")
    # the ctypes mirrors below are passed by reference: a layout mismatch would be silent memory corruption
    for which, cls in ((0, Cosmology), (1, UpperLimits), (2, SnConfig), (3, KnConfig), (4, AgConfig), (5, MnConfig),
                       (6, MkConfig), (8, Extinction)):
        if handle.rb_sizeof(which) != C.sizeof(cls):
            raise RedbackB200Error(f"struct layout mismatch for {cls.__name__}: library {handle.rb_sizeof(which)} "
                                   f"bytes, binding {C.sizeof(cls)} bytes")
    _lib = handle
    return _lib


def check(rc):
    if rc == RB_OK:
        return
    msg = lib().rb_last_error().decode()
    if rc == RB_ERR_INVALID:
        raise ValueError(msg)
    if rc == RB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RedbackB200Error(msg)


def sn_config(engine, sed, sigma_mode=SIGMA_FIXED, dense_resolution=1000, cutoff_wavelength=3000.0,
              absorption_index=1.0, cosmology=None, line=None, precision=0, csm=None, no_interaction=False,
              area_ratio=None, viscous=False):
    cfg = SnConfig()
    check(lib().rb_sn_config_default(C.byref(cfg)))
    cfg.engine, cfg.sed, cfg.sigma_mode = engine, sed, sigma_mode
    cfg.dense_resolution = int(dense_resolution)
    cfg.cutoff_wavelength = float(cutoff_wavelength)
    cfg.absorption_index = float(absorption_index)
    if cosmology is not None:
        cfg.cosmology = cosmology
    if line is not None:
        for key, val in line.items():
            setattr(cfg, key, float(val))
    cfg.precision = int(precision)
    cfg.no_interaction = int(bool(no_interaction))
    cfg.viscous = int(bool(viscous))
    if area_ratio is not None:
        cfg.aspherical, cfg.area_ratio = 1, float(area_ratio)
    if csm is not None:
        cfg.csm_nn, cfg.csm_delta, cfg.csm_efficiency = (float(csm[k]) for k in ("nn", "delta", "efficiency"))
        cfg.csm_quadrature = int(bool(csm.get("quadrature", False)))
        cfg.csm_timesteps = int(csm.get("timesteps", 3000))
    return cfg


def kn_config(model, sigma_mode=SIGMA_FIXED, dense_resolution=500, neutron_precursor_switch=True, vmax=0.7,
              cosmology=None, grid_t_max=7e6, precision=0):
    cfg = KnConfig()
    check(lib().rb_kn_config_default(C.byref(cfg)))
    cfg.model, cfg.sigma_mode = model, sigma_mode
    cfg.dense_resolution = int(dense_resolution)
    cfg.neutron_precursor_switch = 1 if neutron_precursor_switch else 0
    cfg.vmax = float(vmax)
    cfg.grid_t_max = float(grid_t_max)
    cfg.precision = int(precision)
    if cosmology is not None:
        cfg.cosmology = cosmology
    return cfg


def ag_config(jet, sigma_mode=SIGMA_FIXED, steps=250, res=0, k=0, expansion=1, a1=1.0, cosmology=None, ss=0.01,
              aa=0.5, ab_magnitude=False, precision=0):
    cfg = AgConfig()
    check(lib().rb_ag_config_default(C.byref(cfg)))
    cfg.jet, cfg.sigma_mode = jet, sigma_mode
    cfg.steps, cfg.res, cfg.k = int(steps), int(res), int(k)
    cfg.expansion = 1 if expansion else 0
    cfg.a1 = float(a1)
    cfg.ss, cfg.aa = float(ss), float(aa)
    cfg.ab_magnitude = int(bool(ab_magnitude))
    cfg.precision = int(precision)
    if cosmology is not None:
        cfg.cosmology = cosmology
    return cfg


def mn_config(engine, sigma_mode=SIGMA_FIXED, use_gamma_ray_opacity=False, pair_cascade_switch=False,
              ejecta_albedo=0.5, pair_cascade_fraction=0.05, cosmology=None, output=0, photon_index=2.0,
              supernova=None):
    """supernova: dict(sed, cutoff_wavelength, absorption_index) selects the general_magnetar_driven_supernova set-up
    (nickel heating, grid 1e0..1e8 s with 2000 points requested)."""
    cfg = MnConfig()
    check(lib().rb_mn_config_default(C.byref(cfg)))
    cfg.engine, cfg.sigma_mode = engine, sigma_mode
    cfg.use_gamma_ray_opacity = int(bool(use_gamma_ray_opacity))
    cfg.pair_cascade_switch = int(bool(pair_cascade_switch))
    cfg.ejecta_albedo, cfg.pair_cascade_fraction = float(ejecta_albedo), float(pair_cascade_fraction)
    cfg.output, cfg.photon_index = int(output), float(photon_index)
    if supernova is not None:
        cfg.use_r_process, cfg.grid_t_min, cfg.grid_t_max, cfg.resolution = 0, 1e0, 1e8, 2000
        cfg.sed = int(supernova.get("sed", SED_CUTOFF_BLACKBODY))
        cfg.cutoff_wavelength = float(supernova.get("cutoff_wavelength", 3000.0))
        cfg.absorption_index = float(supernova.get("absorption_index", 1.0))
    if cosmology is not None:
        cfg.cosmology = cosmology
    return cfg


def mk_config(engine, sigma_mode=SIGMA_FIXED, use_gamma_ray_opacity=False, pair_cascade_switch=True,
              ejecta_albedo=0.5, pair_cascade_fraction=0.01, neutron_precursor_switch=True, vmax=0.7, cosmology=None,
              resolution=300, magnetar_heating="first_layer"):
    if magnetar_heating not in ("first_layer", "all_layers"):
        raise ValueError("magnetar_heating must be 'first_layer' or 'all_layers'")
    cfg = MkConfig()
    check(lib().rb_mk_config_default(C.byref(cfg)))
    cfg.engine, cfg.sigma_mode = engine, sigma_mode
    cfg.use_gamma_ray_opacity = int(bool(use_gamma_ray_opacity))
    cfg.pair_cascade_switch = int(bool(pair_cascade_switch))
    cfg.ejecta_albedo, cfg.pair_cascade_fraction = float(ejecta_albedo), float(pair_cascade_fraction)
    cfg.neutron_precursor_switch = int(bool(neutron_precursor_switch))
    cfg.vmax = float(vmax)
    cfg.resolution = int(resolution)
    cfg.heat_all_layers = int(magnetar_heating == "all_layers")
    if cosmology is not None:
        cfg.cosmology = cosmology
    return cfg


def optimal_time_array(t_min, t_max, resolution):
    """utils.get_optimal_time_array without user times, through the library's host function."""
    import numpy as np
    out = (C.c_double * int(resolution))()
    n = C.c_int(0)
    check(lib().rb_optimal_time_array(float(t_min), float(t_max), int(resolution), out, C.byref(n)))
    return np.array(out[:n.value])


def flat_lcdm(H0, Om0, Tcmb0=2.7255, Neff=3.046, m_nu=(0.06, 0.0, 0.0)):
    out = Cosmology()
    arr = (C.c_double * 3)(*[float(m) for m in m_nu])
    check(lib().rb_cosmology_flat_lcdm(float(H0), float(Om0), float(Tcmb0), float(Neff), arr, C.byref(out)))
    return out
