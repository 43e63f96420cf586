"""CPU oracle: numpy restatement of redback's light-curve + Gaussian-likelihood hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``redback_b200/`` may import this module; it exists so that
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` have an independent statement of what the reference computes.

Every function cites the reference file:line it follows (paths relative to /root/reference).  The
arithmetic is float64 numpy/scipy, one parameter sample per call, exactly like the reference.

Pinning (see tests/test_oracle_golden.py):
  * supernova / kilonova / magnetar models are checked against the reference's own golden vectors
    test/reference_results/all_models_reference.pkl.gz (extracted to tests/golden/ by
    oracle/make_golden.py) at the reference's tolerance and much tighter;
  * Diffusion / TemperatureFloor are additionally checked against the reference's *own source files*
    imported verbatim through stub modules (oracle/refshim.py) when /root/reference is present;
  * GaussianLikelihood closed forms from test/likelihood_test.py:63-69,148-153;
  * TemperatureFloor known answer from test/photosphere_test.py:24-29.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.integrate import quad
from scipy.interpolate import interp1d

# --------------------------------------------------------------------------------------------------
# constants  (redback/constants.py:1-23 -> astropy CODATA-2018 cgs values)
# --------------------------------------------------------------------------------------------------
speed_of_light = 29979245800.0
planck = 6.62607015e-27
boltzmann_constant = 1.380649e-16
sigma_sb = 5.6703744191844314e-05
solar_mass = 1.988409870698051e33
proton_mass = 1.67262192369e-24
electron_mass = 9.1093837015e-28
sigma_T = 6.6524587321e-25
km_cgs = 100000.0
day_to_s = 86400
angstrom_cgs = 1e-08
speed_of_light_si = 299792458.0

# --------------------------------------------------------------------------------------------------
# cosmology: astropy Planck18 = FlatLambdaCDM(H0=67.66, Om0=0.30966, Tcmb0=2.7255, Neff=3.046,
# m_nu=[0,0,0.06] eV).  astropy is not installed here, so this restates its published algorithm
# (astropy/cosmology/flrw/base.py: nu_relative_density, inv_efunc; flrw/lambdacdm.py FlatLambdaCDM)
# and is anchored by every flux_density golden vector (d_L enters as d_L**-2).
# Call site: supernova_models.py:997 `cosmology.luminosity_distance(redshift).cgs.value`.
# --------------------------------------------------------------------------------------------------
_MPC_CM = 3.0856775814913674e24
_G_SI = 6.6743e-11
_KB_EV_K = 8.617333262145179e-05


class Planck18:
    H0 = 67.66
    Om0 = 0.30966
    Tcmb0 = 2.7255
    Neff = 3.046
    m_nu = (0.0, 0.0, 0.06)

    def __init__(self):
        H0_s = self.H0 * 1e5 / _MPC_CM                      # 1/s
        rho_crit = 3.0 * H0_s ** 2 / (8.0 * math.pi * _G_SI) * 1e-3   # g/cm^3 (G in SI -> kg/m^3 -> g/cm^3)
        a_B_c2 = 4.0 * (sigma_sb * 1e-3) / (speed_of_light_si ** 3) * 1e-3  # g/cm^3/K^4
        self.Ogamma0 = a_B_c2 * self.Tcmb0 ** 4 / rho_crit
        self.Tnu0 = 0.7137658555036082 * self.Tcmb0
        nnu = int(math.floor(self.Neff))
        self.neff_per_nu = self.Neff / nnu
        massive = [m for m in self.m_nu if m > 0]
        self.nmassless = nnu - len(massive)
        self.nu_y = np.array(massive) / (_KB_EV_K * self.Tnu0)
        self.Onu0 = self.Ogamma0 * self.nu_relative_density(0.0)
        self.Ode0 = 1.0 - self.Om0 - self.Ogamma0 - self.Onu0
        self.hubble_distance_mpc = speed_of_light_si / 1e3 / self.H0

    def nu_relative_density(self, z):
        prefac = 0.22710731766
        p, invp, k = 1.83, 0.54644808743, 0.3173
        curr = self.nu_y / (1.0 + z)
        rel = (1.0 + (k * curr) ** p) ** invp
        return prefac * self.neff_per_nu * (rel.sum() + self.nmassless)

    def inv_efunc(self, z):
        zp1 = 1.0 + z
        Or = self.Ogamma0 * (1.0 + self.nu_relative_density(z))
        return (zp1 ** 3 * (Or * zp1 + self.Om0) + self.Ode0) ** -0.5

    def luminosity_distance_cm(self, z):
        integral = quad(self.inv_efunc, 0.0, float(z), epsabs=0.0, epsrel=1e-13, limit=200)[0]
        return (1.0 + z) * self.hubble_distance_mpc * integral * _MPC_CM


cosmo = Planck18()


# --------------------------------------------------------------------------------------------------
# utilities (redback/utils.py)
# --------------------------------------------------------------------------------------------------
def calc_kcorrected_properties(frequency, redshift, time):
    """utils.py:373-384"""
    return frequency * (1 + redshift), time / (1 + redshift)


def lambda_to_nu(wavelength):
    """utils.py:357-362"""
    return speed_of_light_si / (wavelength * 1.e-10)


def nu_to_lambda(frequency):
    """utils.py:365-370"""
    return 1.e10 * (speed_of_light_si / frequency)


# --------------------------------------------------------------------------------------------------
# engines
# --------------------------------------------------------------------------------------------------
def nickelcobalt_engine(time, f_nickel, mej):
    """supernova_models.py:564-579"""
    return f_nickel * mej * (6.45e43 * np.exp(-time / 8.8) + 1.45e43 * np.exp(-time / 111.3))


def basic_magnetar(time, p0, bp, mass_ns, theta_pb):
    """magnetar_models.py:171-185 (time in seconds)"""
    with np.errstate(all="ignore"):
        erot = 2.6e52 * (mass_ns / 1.4) ** (3. / 2.) * p0 ** (-2)
        tp = 1.3e5 * bp ** (-2) * p0 ** 2 * (mass_ns / 1.4) ** (3. / 2.) * (np.sin(theta_pb)) ** (-2)
        return 2. * erot / tp / (1. + 2. * time / tp) ** 2


# --------------------------------------------------------------------------------------------------
# interaction process: Arnett diffusion  (interaction_processes.py:33-65)
# --------------------------------------------------------------------------------------------------
def diffusion(time, dense_times, luminosity, kappa, kappa_gamma, mej, vej):
    """Return (tau_diff [days], L_bol[E]).  Follows interaction_processes.py:33-65 step by step."""
    time = np.asarray(time, dtype=float)
    with np.errstate(all="ignore"):
        diffusion_constant = 2.0 * solar_mass / (13.7 * speed_of_light * km_cgs)          # :36
        trapping_constant = 3.0 * solar_mass / (4 * np.pi * km_cgs ** 2)                   # :37
        tau_diff = np.sqrt(diffusion_constant * kappa * mej / vej) / day_to_s              # :39
        trap_coeff = (trapping_constant * kappa_gamma * mej / (vej ** 2)) / day_to_s ** 2  # :40
        t_end = dense_times[-1]
        tb = max(0.0, np.min(dense_times))                                                  # :42-43
        interp = interp1d(dense_times, luminosity, copy=False, assume_sorted=True)         # :44
        uniq = np.unique(time[(time >= tb) & (time <= t_end)])                              # :46
        lu = len(uniq)
        lsp = np.logspace(np.log10(tau_diff / t_end) - 3, 0, 50)                            # :49-50
        xm = np.unique(np.concatenate((lsp, 1 - lsp)))                                      # :51
        nodes = np.clip(tb + (uniq.reshape(lu, 1) - tb) * xm, tb, t_end)                    # :53
        te2 = nodes[:, -1] ** 2                                                             # :55
        lum = interp(nodes)                                                                 # :56
        integrand = lum * nodes * np.exp((nodes ** 2 - te2.reshape(lu, 1)) / tau_diff ** 2)  # :57
        integrand[np.isnan(integrand)] = 0.                                                 # :58
        out = np.trapezoid(integrand, nodes, axis=1)                                        # :60
        out *= -2.0 * np.expm1(-trap_coeff / te2) / tau_diff ** 2                           # :61
        return tau_diff, out[np.searchsorted(uniq, time)]                                   # :63


def viscous(time, dense_times, luminosity, t_viscous):
    """interaction_processes.Viscous.convert_input_luminosity (:229-273), step by step.  Returns L_out[E]."""
    time = np.asarray(time, dtype=float)
    with np.errstate(all="ignore"):
        t_end = dense_times[-1]
        tb = max(0.0, min(dense_times))                                                     # :252-253
        interp = interp1d(dense_times, luminosity, copy=False, assume_sorted=True)         # :254
        uniq = np.unique(time[(time >= tb) & (time <= t_end)])                              # :256
        lu = len(uniq)
        lsp = np.logspace(np.log10(t_viscous / t_end) - 3, 0, 500)                          # :259-260
        xm = np.unique(np.concatenate((lsp, 1 - lsp)))                                      # :261
        nodes = np.clip(tb + (uniq.reshape(lu, 1) - tb) * xm, tb, t_end)                    # :263
        tes = nodes[:, -1]                                                                  # :265
        lum = interp(nodes)                                                                 # :266
        integrand = lum * np.exp((nodes - tes.reshape(lu, 1)) / t_viscous)                  # :267
        integrand[np.isnan(integrand)] = 0.0                                                # :268
        out = np.trapezoid(integrand, nodes, axis=1) / t_viscous                            # :270
        return out[np.searchsorted(uniq, time)]                                             # :272


def arnett_bolometric_viscous(time, f_nickel, mej, *, t_viscous, dense_resolution=1000, **_):
    """supernova_models.arnett_bolometric (:949-969) called with interaction_process=ip.Viscous, t_viscous=..."""
    time = np.asarray(time, dtype=float)
    dense_times = np.linspace(0, time[-1] + 100, dense_resolution)
    return viscous(time, dense_times, nickelcobalt_engine(dense_times, f_nickel, mej), t_viscous)


def basic_magnetar_powered_viscous(time, redshift, p0, bp, mass_ns, theta_pb, *, frequency, vej, t_viscous,
                                   temperature_floor, dl=None, dense_resolution=1000, **_):
    """supernova_models.basic_magnetar_powered (:1471-1509, flux_density) with interaction_process=ip.Viscous."""
    def bolometric(t):
        dense_times = np.linspace(0, t[-1] + 100, dense_resolution)
        return viscous(t, dense_times, basic_magnetar(dense_times * day_to_s, p0, bp, mass_ns, theta_pb), t_viscous)
    return _sn_flux_density(time, redshift, bolometric, "blackbody", frequency=frequency, vej=vej,
                            temperature_floor=temperature_floor, dl=dl)


# --------------------------------------------------------------------------------------------------
# photosphere  (photosphere.py:56-111)
# --------------------------------------------------------------------------------------------------
def aspherical_diffusion(time, dense_times, luminosity, kappa, kappa_gamma, mej, vej, area_projection, area_reference):
    """interaction_processes.AsphericalDiffusion (:66-131): Diffusion's quadrature times the viewing-angle factor
    (:124-125).  Returns (tau_diff, new_luminosity)."""
    tau_diff, lum = diffusion(time, dense_times, luminosity, kappa, kappa_gamma, mej, vej)
    time = np.asarray(time, dtype=float)
    with np.errstate(all="ignore"):
        lum = lum * (1 + 1.4 * (2 + time / tau_diff / 0.59) / (1 + np.exp(time / tau_diff / 0.59))
                     * (area_projection / area_reference - 1))
    return tau_diff, lum


def temperature_floor(time, luminosity, vej, temperature_floor):
    """Return (T_phot[E], R_phot[E]).  photosphere.py:87-111"""
    with np.errstate(all="ignore"):
        radius_constant = km_cgs * day_to_s
        stef_constant = 4 * np.pi * sigma_sb
        r2 = (radius_constant * vej * time) ** 2
        rec2 = luminosity / (stef_constant * temperature_floor ** 4)
        mask = r2 <= rec2
        r = rec2 ** 0.5
        r[mask] = r2[mask] ** 0.5
        temp = np.zeros(len(time))
        temp[mask] = (luminosity[mask] / (stef_constant * r2[mask])) ** 0.25
        temp[~mask] = temperature_floor
        return temp, r


# --------------------------------------------------------------------------------------------------
# SEDs  (sed.py)
# --------------------------------------------------------------------------------------------------
_ERG_S_CM2_HZ_TO_MJY = 1e26


def blackbody_to_flux_density(temperature, r_photosphere, dl, frequency):
    """sed.py:199-222; returns erg/s/Hz/cm^2 (plain floats instead of astropy Quantity)."""
    with np.errstate(all="ignore"):
        num = 2 * np.pi * planck * frequency ** 3 * r_photosphere ** 2
        denom = dl ** 2 * speed_of_light ** 2
        frac = 1. / np.expm1((planck * frequency) / (boltzmann_constant * temperature))
        return num / denom * frac


def cutoff_blackbody_mjy(time, temperature, luminosity, r_photosphere, frequency, dl,
                         cutoff_wavelength=3000., absorption_index=1.0):
    """sed.py:301-426 + _SED.flux_density (sed.py:225-251).  `frequency` is [E] (or scalar) for the
    flux_density branch, or [N_lambda, 1] for the spectral-grid branch.  Returns mJy."""
    from scipy.special import gammainc, gammaincc, gamma as gamma_func
    time = np.asarray(time, dtype=float)
    X_CONST = planck * speed_of_light / boltzmann_constant
    FLUX_CONST = 4 * np.pi * 2 * np.pi * planck * speed_of_light ** 2 * angstrom_cgs
    if isinstance(frequency, (float, int)):
        frequency = np.array([frequency] * len(time))                                      # :233-234
    uniq = np.unique(time)
    uniq_is = np.searchsorted(time, uniq, sorter=np.argsort(time))                         # :329-331
    lam_c = cutoff_wavelength * angstrom_cgs
    alpha = absorption_index
    if alpha >= 4.0:
        raise ValueError("absorption_index >= 4 is not supported")                         # :396-400
    with np.errstate(all="ignore"):
        # _set_norm, :386-421
        norms = luminosity[uniq_is] / (FLUX_CONST / angstrom_cgs * r_photosphere[uniq_is] ** 2
                                       * temperature[uniq_is])
        tp = temperature[uniq_is]
        kT_over_hc = tp / X_CONST
        n_arr = np.arange(1, 11, dtype=float)
        x_c = 1.0 / (lam_c * kT_over_hc)
        n_x_c = np.outer(x_c, n_arr)
        above = kT_over_hc ** 4 * gamma_func(4.0) * np.sum(gammainc(4.0, n_x_c) / n_arr ** 4, axis=1)
        s = 4.0 - alpha
        below = (1.0 / lam_c ** alpha) * kT_over_hc ** s * gamma_func(s) * \
            np.sum(gammaincc(s, n_x_c) / n_arr ** s, axis=1)
        norms = norms / ((above + below) / tp)
        # wavelength property, :345-352
        if len(frequency) == 1:
            frequency = np.ones(len(time)) * frequency
        wavelength = nu_to_lambda(frequency) * angstrom_cgs
        if len(wavelength) != len(time):
            wavelength = np.tile(wavelength.T, (len(time), 1)).T
        mask = wavelength < lam_c
        # _set_sed, :362-384
        if np.size(wavelength) != np.size(time):
            R = np.tile(r_photosphere, (len(frequency), 1))
            T = np.tile(temperature, (len(frequency), 1))
        else:
            R, T = r_photosphere, temperature
        sed = np.where(mask,
                       FLUX_CONST * (R ** 2 * (wavelength / lam_c) ** alpha / wavelength ** 5)
                       / np.expm1(X_CONST / wavelength / T),
                       FLUX_CONST * (R ** 2 / wavelength ** 5) / np.expm1(X_CONST / wavelength / T))
        sed = sed * norms[np.searchsorted(uniq, time)]
        # _SED.flux_density, :239-251
        fd = sed / (4 * np.pi * dl ** 2)
        fd = fd * nu_to_lambda(frequency)
        fd = fd / frequency
        return fd * _ERG_S_CM2_HZ_TO_MJY


# --------------------------------------------------------------------------------------------------
# supernova models (supernova_models.py)
# --------------------------------------------------------------------------------------------------
def arnett_bolometric(time, f_nickel, mej, *, vej, kappa, kappa_gamma, dense_resolution=1000, **_):
    """supernova_models.py:949-969 with the default interaction_process=Diffusion."""
    time = np.asarray(time, dtype=float)
    dense_times = np.linspace(0, time[-1] + 100, dense_resolution)
    dense_lbols = nickelcobalt_engine(dense_times, f_nickel, mej)
    return diffusion(time, dense_times, dense_lbols, kappa, kappa_gamma, mej, vej)[1]


def basic_magnetar_powered_bolometric(time, p0, bp, mass_ns, theta_pb, *, mej, vej, kappa, kappa_gamma,
                                      dense_resolution=1000, **_):
    """supernova_models.py:1447-1469 (slsn_bolometric :1532-1549 is the same function)."""
    time = np.asarray(time, dtype=float)
    dense_times = np.linspace(0, time[-1] + 100, dense_resolution)
    dense_lbols = basic_magnetar(dense_times * day_to_s, p0, bp, mass_ns, theta_pb)
    return diffusion(time, dense_times, dense_lbols, kappa, kappa_gamma, mej, vej)[1]


slsn_bolometric = basic_magnetar_powered_bolometric


def _sn_flux_density(time, redshift, bolometric, sed, *, frequency, vej, temperature_floor, dl=None,
                     cutoff_wavelength=3000., absorption_index=1.0):
    """flux_density branch shared by arnett (:999-1007), basic_magnetar_powered (:1501-1509) and
    slsn (:1587-1597).  Returns mJy[E]."""
    time = np.asarray(time, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    frequency, time = calc_kcorrected_properties(np.asarray(frequency, dtype=float), redshift, time)
    lbol = bolometric(time)
    temp, rad = globals()["temperature_floor"](time, lbol, vej, temperature_floor)
    with np.errstate(all="ignore"):
        if sed == "blackbody":
            fd = blackbody_to_flux_density(temp, rad, dl, frequency) * _ERG_S_CM2_HZ_TO_MJY
        elif sed == "cutoff_blackbody":
            freq = frequency if np.ndim(frequency) else float(frequency)
            fd = cutoff_blackbody_mjy(time, temp, lbol, rad, freq, dl, cutoff_wavelength, absorption_index)
        else:
            raise ValueError(sed)
        return fd * (1 + redshift)


def arnett(time, redshift, f_nickel, mej, *, frequency, vej, kappa, kappa_gamma, temperature_floor,
           dl=None, **kw):
    """supernova_models.py:971-1007 (flux_density branch)."""
    return _sn_flux_density(
        time, redshift,
        lambda t: arnett_bolometric(t, f_nickel, mej, vej=vej, kappa=kappa, kappa_gamma=kappa_gamma, **kw),
        "blackbody", frequency=frequency, vej=vej, temperature_floor=temperature_floor, dl=dl)


def synchrotron_mjy(frequency, dl, pp, nu_max, source_radius=1e13, f0=1e-26):
    """sed.Synchrotron (sed.py:703-751) through _SED.flux_density (:238-251): mJy at source-frame frequencies."""
    frequency = np.atleast_1d(np.asarray(frequency, dtype=float))
    sed = np.zeros(frequency.shape)
    mask = frequency < nu_max
    f_max = f0 * source_radius ** 2 * nu_max ** 2.5
    sed[mask] = f0 * source_radius ** 2 * (frequency[mask] / nu_max) ** 2.5 * angstrom_cgs / speed_of_light \
        * frequency[mask] ** 2
    sed[~mask] = f_max * (frequency[~mask] / nu_max) ** (-(pp - 1.) / 2.) * angstrom_cgs / speed_of_light \
        * frequency[~mask] ** 2
    fd = sed / (4 * np.pi * dl ** 2)
    fd = fd * nu_to_lambda(frequency)
    fd = fd / frequency
    return fd * _ERG_S_CM2_HZ_TO_MJY


def type_1c(time, redshift, f_nickel, mej, pp, *, frequency, nu_max=1e9, f0=1e-26, dl=None, **kw):
    """supernova_models.type_1c (:2317-2352), flux_density: arnett with Blackbody + Synchrotron."""
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    time = np.asarray(time, dtype=float)
    freq = np.broadcast_to(np.asarray(frequency, dtype=float), time.shape)
    nu_src, _ = calc_kcorrected_properties(freq, redshift, time)
    return arnett(time, redshift, f_nickel, mej, frequency=frequency, dl=dl, **kw) \
        + synchrotron_mjy(nu_src, dl, pp, nu_max, f0=f0) * (1 + redshift)


def basic_magnetar_powered(time, redshift, p0, bp, mass_ns, theta_pb, *, frequency, mej, vej, kappa,
                           kappa_gamma, temperature_floor, dl=None, **kw):
    """supernova_models.py:1471-1509 (flux_density branch)."""
    return _sn_flux_density(
        time, redshift,
        lambda t: basic_magnetar_powered_bolometric(t, p0, bp, mass_ns, theta_pb, mej=mej, vej=vej,
                                                    kappa=kappa, kappa_gamma=kappa_gamma, **kw),
        "blackbody", frequency=frequency, vej=vej, temperature_floor=temperature_floor, dl=dl)


def slsn(time, redshift, p0, bp, mass_ns, theta_pb, *, frequency, mej, vej, kappa, kappa_gamma,
         temperature_floor, dl=None, sed="cutoff_blackbody", cutoff_wavelength=3000.,
         absorption_index=1.0, **kw):
    """supernova_models.py:1552-1597 (flux_density branch; default sed = CutoffBlackbody)."""
    return _sn_flux_density(
        time, redshift,
        lambda t: basic_magnetar_powered_bolometric(t, p0, bp, mass_ns, theta_pb, mej=mej, vej=vej,
                                                    kappa=kappa, kappa_gamma=kappa_gamma, **kw),
        sed, frequency=frequency, vej=vej, temperature_floor=temperature_floor, dl=dl,
        cutoff_wavelength=cutoff_wavelength, absorption_index=absorption_index)


def _diffused(engine, time, *, mej, vej, kappa, kappa_gamma, dense_resolution=1000):
    """engine(dense_times) -> Diffusion on np.linspace(0, time[-1] + 100, dense_resolution) (the pattern of
    supernova_models.py:373-379, 2392-2400, tde_models.py:693-700)."""
    dense = np.linspace(0, time[-1] + 100, dense_resolution)
    with np.errstate(all="ignore"):
        lum = engine(dense)
    return diffusion(time, dense, lum, kappa, kappa_gamma, mej, vej)[1]


def exponential_powerlaw(time, a_1, alpha_1, alpha_2, tpeak):
    """phenomenological_models.exponential_powerlaw (:743-754)."""
    return a_1 * (1 - np.exp(-time / tpeak)) ** alpha_1 * (time / tpeak) ** (-alpha_2)


def analytic_fallback(time, l0, t_0):
    """tde_models._analytic_fallback (:16-27)."""
    mask = time - t_0 > 0.
    lbol = np.zeros(len(time))
    lbol[mask] = l0 / (time[mask] * 86400) ** (5. / 3.)
    lbol[~mask] = l0 / (t_0 * 86400) ** (5. / 3.)
    return lbol


def general_magnetar_slsn_bolometric(time, l0, tsd, nn, **dkw):
    """supernova_models.general_magnetar_slsn_bolometric (:2379-2401)."""
    return _diffused(lambda d: magnetar_only(d * day_to_s, l0, tsd * day_to_s, nn), np.asarray(time, dtype=float), **dkw)


def exponential_powerlaw_bolometric(time, lbol_0, alpha_1, alpha_2, tpeak_d, **dkw):
    """supernova_models.exponential_powerlaw_bolometric (:357-381)."""
    return _diffused(lambda d: exponential_powerlaw(d, lbol_0, alpha_1, alpha_2, tpeak_d),
                     np.asarray(time, dtype=float), **dkw)


def tde_analytical_bolometric(time, l0, t_0_turn, **dkw):
    """tde_models.tde_analytical_bolometric (:682-701)."""
    return _diffused(lambda d: analytic_fallback(d, l0, t_0_turn), np.asarray(time, dtype=float), **dkw)


def _diffused_flux_density(bolometric, sed, time, redshift, *, frequency, mej, vej, kappa, kappa_gamma,
                           temperature_floor, dl=None, dense_resolution=1000, **sed_kw):
    return _sn_flux_density(
        time, redshift,
        lambda t: bolometric(t, mej=mej, vej=vej, kappa=kappa, kappa_gamma=kappa_gamma, dense_resolution=dense_resolution),
        sed, frequency=frequency, vej=vej, temperature_floor=temperature_floor, dl=dl, **sed_kw)


def general_magnetar_slsn(time, redshift, l0, tsd, nn, **kw):
    """supernova_models.general_magnetar_slsn (:2404-2440), flux_density, Blackbody."""
    return _diffused_flux_density(lambda t, **d: general_magnetar_slsn_bolometric(t, l0, tsd, nn, **d), "blackbody",
                                  time, redshift, **kw)


def sn_exponential_powerlaw(time, redshift, lbol_0, alpha_1, alpha_2, tpeak_d, **kw):
    """supernova_models.sn_exponential_powerlaw (:501-560), flux_density, Blackbody."""
    return _diffused_flux_density(
        lambda t, **d: exponential_powerlaw_bolometric(t, lbol_0, alpha_1, alpha_2, tpeak_d, **d), "blackbody", time,
        redshift, **kw)


def tde_analytical(time, redshift, l0, t_0_turn, **kw):
    """tde_models.tde_analytical (:704-760), flux_density, CutoffBlackbody (cutoff_wavelength 3000 A by default)."""
    return _diffused_flux_density(lambda t, **d: tde_analytical_bolometric(t, l0, t_0_turn, **d), "cutoff_blackbody",
                                  time, redshift, **kw)


def fallback_lbol(time, logl1, tr):
    """phenomenological_models.fallback_lbol (:690-702)."""
    l1 = 10 ** logl1
    time = np.asarray(time, dtype=float) * 86400
    tr = tr * 86400
    lbol = l1 * time ** (-5. / 3.)
    lbol[time < tr] = l1 * tr ** (-5. / 3.)
    return lbol


def sn_fallback(time, redshift, logl1, tr, *, frequency, vej, temperature_floor, dl=None, **_):
    """supernova_models.sn_fallback (:383-412): no interaction process."""
    return _sn_flux_density(time, redshift, lambda t: fallback_lbol(t, logl1, tr), "blackbody", frequency=frequency,
                            vej=vej, temperature_floor=temperature_floor, dl=dl)


def sn_nickel_fallback(time, redshift, mej, f_nickel, logl1, tr, *, frequency, vej, temperature_floor, dl=None, **_):
    """supernova_models.sn_nickel_fallback (:441-470)."""
    return _sn_flux_density(time, redshift,
                            lambda t: fallback_lbol(t, logl1, tr) + nickelcobalt_engine(t, f_nickel, mej), "blackbody",
                            frequency=frequency, vej=vej, temperature_floor=temperature_floor, dl=dl)


def arnett_no_interaction(time, redshift, f_nickel, mej, *, frequency, vej, temperature_floor, dl=None, **_):
    """supernova_models.arnett with interaction_process=None (:961-968, 999-1007): the raw engine luminosity."""
    return _sn_flux_density(time, redshift, lambda t: nickelcobalt_engine(t, f_nickel, mej), "blackbody",
                            frequency=frequency, vej=vej, temperature_floor=temperature_floor, dl=dl)


def velocity_from_kinetic_energy(mej, ek, thin_shell=False):
    """supernova_models.py:1727 (homologous expansion) / :1763 (thin shell), km/s."""
    if thin_shell:
        return np.sqrt(2.0 * ek / (mej * solar_mass)) / km_cgs
    return np.sqrt(10.0 * ek / (3.0 * mej * solar_mass)) / km_cgs


def homologous_expansion_supernova(time, redshift, mej, ek, *, thin_shell=False, **kw):
    """supernova_models.homologous_expansion_supernova (:1782-1844) / thin_shell_supernova (:1847-1905) with the
    default base model arnett_bolometric."""
    return arnett(time, redshift, kw.pop("f_nickel"), mej, vej=velocity_from_kinetic_energy(mej, ek, thin_shell), **kw)


def t0_base_model(time, t0, base, output_format="flux_density", **kw):
    """phase_models.t0_base_model (phase_models.py:19-59): `base(transient_time, frequency=..., bands=...)` is one of
    the restated models; astropy's `(Time(time, 'mjd') - Time(t0, 'mjd')).to(day)` is the double-double difference of
    the two MJDs, i.e. time - t0 rounded once."""
    time = np.asarray(time, dtype=float) - t0
    good = time >= 0.0
    kw = dict(kw)
    for key in ("frequency", "bands"):
        if key in kw and isinstance(kw[key], np.ndarray):
            kw[key] = kw[key][good]
    real = base(time[good], **kw)
    fake = np.zeros(int((~good).sum())) + (1000 if output_format == "magnitude" else 0)
    return np.concatenate((fake, real))


def magnetar_nickel(time, redshift, f_nickel, mej, p0, bp, mass_ns, theta_pb, *, frequency, vej, kappa, kappa_gamma,
                    temperature_floor, dl=None, dense_resolution=1000, **_):
    """supernova_models.magnetar_nickel (:1628-1681), flux_density: nickel + magnetar engines through one Diffusion."""
    def bolometric(t):
        dense = np.linspace(0, t[-1] + 100, dense_resolution)
        lum = nickelcobalt_engine(dense, f_nickel, mej)
        lum += basic_magnetar(dense * day_to_s, p0, bp, mass_ns, theta_pb)                     # :1669-1670
        return diffusion(t, dense, lum, kappa, kappa_gamma, mej, vej)[1]
    return _sn_flux_density(time, redshift, bolometric, "blackbody", frequency=frequency, vej=vej,
                            temperature_floor=temperature_floor, dl=dl)


def gaussian_likelihood_fractional(y, model, sigma_i):
    """GaussianLikelihoodWithFractionalNoise.log_likelihood, likelihoods.py:821-851."""
    with np.errstate(all="ignore"):
        full = np.sqrt(np.asarray(sigma_i) ** 2. * np.asarray(model) ** 2)
    return float(np.nan_to_num(gaussian_log_likelihood(np.asarray(y) - model, full)))


def gaussian_likelihood_systematic(y, model, sigma_i, sigma):
    """GaussianLikelihoodWithSystematicNoise.log_likelihood, likelihoods.py:883-913."""
    with np.errstate(all="ignore"):
        full = np.sqrt(np.asarray(sigma_i) ** 2. + np.asarray(model) ** 2 * sigma ** 2.)
    return float(np.nan_to_num(gaussian_log_likelihood(np.asarray(y) - model, full)))


def line_sed_cgs(time, luminosity, frequency, base_flux_cgs, dl, line_wavelength=7.5e3, line_width=500,
                 line_time=50, line_duration=25, line_amplitude=0.3):
    """sed.Line._set_sed (sed.py:859-909).  `frequency`: (1, n_freq) row for the flux_density branch or (n_freq, 1)
    column for the grid branch; base flux density in erg/s/cm^2/Hz, broadcastable."""
    time_vals = np.atleast_1d(time).reshape(-1, 1).T
    lum_vals = np.atleast_1d(luminosity).reshape(-1, 1).T
    with np.errstate(all="ignore"):
        amplitude_time = line_amplitude * np.exp(-0.5 * ((time_vals - line_time) / line_duration) ** 2)
        flux_modified = base_flux_cgs * (1 - amplitude_time)
        amplitude_scaled = amplitude_time * lum_vals / (4 * np.pi * dl ** 2)
        amplitude_scaled = amplitude_scaled / (line_width * np.sqrt(2 * np.pi))
        wavelength = nu_to_lambda(frequency)
        line_profile = np.exp(-0.5 * ((wavelength - line_wavelength) / line_width) ** 2)
        wavelength_cm = wavelength * angstrom_cgs
        return flux_modified + amplitude_scaled * line_profile * (wavelength_cm ** 2) / speed_of_light


def type_1a(time, redshift, f_nickel, mej, *, frequency, vej, kappa, kappa_gamma, temperature_floor, dl=None,
            cutoff_wavelength=3000, absorption_index=1.0, line_wavelength=7.5e3, line_width=500, line_time=50,
            line_duration=25, line_amplitude=0.3, **kw):
    """supernova_models.py:2229-2278 (flux_density branch).  Returns mJy[E]."""
    time = np.asarray(time, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    frequency = np.asarray(frequency, dtype=float)
    if frequency.ndim == 0:
        frequency = np.full(time.shape, float(frequency))
    frequency, t = calc_kcorrected_properties(frequency, redshift, time)
    lbol = arnett_bolometric(t, f_nickel, mej, vej=vej, kappa=kappa, kappa_gamma=kappa_gamma, **kw)
    temp, rad = globals()["temperature_floor"](t, lbol, vej, temperature_floor)
    with np.errstate(all="ignore"):
        base = cutoff_blackbody_mjy(t, temp, lbol, rad, frequency, dl, cutoff_wavelength, absorption_index) * 1e-26
        fd = line_sed_cgs(t, lbol, frequency.reshape(1, -1), base, dl, line_wavelength, line_width, line_time,
                          line_duration, line_amplitude).flatten()
        return fd * _ERG_S_CM2_HZ_TO_MJY * (1 + redshift)


# --------------------------------------------------------------------------------------------------
# likelihoods (likelihoods.py)
# --------------------------------------------------------------------------------------------------
def gaussian_log_likelihood(res, sigma):
    """likelihoods.py:229-231"""
    with np.errstate(all="ignore"):
        return np.sum(- (res / sigma) ** 2 / 2 - np.log(2 * np.pi * sigma ** 2) / 2)


def gaussian_likelihood(y, model, sigma):
    """GaussianLikelihood.log_likelihood, likelihoods.py:221-227 (incl. np.nan_to_num)."""
    return float(np.nan_to_num(gaussian_log_likelihood(np.asarray(y) - model, sigma)))


def normal_cdf(x):
    """GaussianLikelihoodWithUpperLimits._normal_cdf, likelihoods.py:366-375."""
    from scipy.special import erf
    return 0.5 * (1.0 + erf(x / np.sqrt(2)))


def upper_limit_log_likelihood(observed_ul, model_ul, ul_sigma_levels, data_mode="flux"):
    """GaussianLikelihoodWithUpperLimits._upper_limit_log_likelihood, likelihoods.py:377-419, on the upper-limit
    points only."""
    observed_ul = np.asarray(observed_ul, dtype=float)
    model_ul = np.asarray(model_ul, dtype=float)
    if observed_ul.size == 0:
        return 0.0
    if data_mode == "magnitude":
        sigma_measurement = (2.5 / np.log(10.0)) / ul_sigma_levels
        cdf_values = 1.0 - normal_cdf((observed_ul - model_ul) / sigma_measurement)
    else:
        sigma_measurement = observed_ul / ul_sigma_levels
        cdf_values = normal_cdf((observed_ul - model_ul) / sigma_measurement)
    epsilon = 1e-30
    cdf_values = np.clip(cdf_values, epsilon, 1.0 - epsilon)
    return np.sum(np.log(cdf_values))


def gaussian_likelihood_upper_limits(y, model, sigma, detections, upper_limit_sigma=3.0, data_mode="flux"):
    """GaussianLikelihoodWithUpperLimits.log_likelihood, likelihoods.py:448-468."""
    y, model = np.asarray(y, dtype=float), np.asarray(model, dtype=float)
    det = np.asarray(detections, dtype=bool)
    levels = np.broadcast_to(np.asarray(upper_limit_sigma, dtype=float), y.shape)[~det] \
        if np.ndim(upper_limit_sigma) == 0 or len(upper_limit_sigma) == len(y) else np.asarray(upper_limit_sigma)
    det_ll = 0.0
    if np.any(det):
        sig = sigma if np.isscalar(sigma) else np.asarray(sigma)[det]
        det_ll = gaussian_log_likelihood((y - model)[det], sig)
    with np.errstate(all="ignore"):
        ul_ll = upper_limit_log_likelihood(y[~det], model[~det], levels, data_mode)
    return float(np.nan_to_num(det_ll + ul_ll))


def mixture_gaussian_likelihood(y, model, sigma, sigma_out, alpha):
    """MixtureGaussianLikelihood.log_likelihood, likelihoods.py:542-617."""
    res = np.asarray(y, dtype=float) - model
    with np.errstate(all="ignore"):
        logp_in = -0.5 * np.log(2 * np.pi) - np.log(sigma) - 0.5 * (res / sigma) ** 2
        logp_out = -0.5 * np.log(2 * np.pi) - np.log(sigma_out) - 0.5 * (res / sigma_out) ** 2
        return float(np.sum(np.logaddexp(np.log(alpha) + logp_in, np.log(1 - alpha) + logp_out)))


def mixture_outlier_posteriors(y, model, sigma, sigma_out, alpha):
    """MixtureGaussianLikelihood.calculate_outlier_posteriors, likelihoods.py:637-665."""
    r_ = np.asarray(y, dtype=float) - model
    pin = (1 / (np.sqrt(2 * np.pi) * sigma)) * np.exp(-0.5 * (r_ / sigma) ** 2)
    pout = (1 / (np.sqrt(2 * np.pi) * sigma_out)) * np.exp(-0.5 * (r_ / sigma_out) ** 2)
    den = alpha * pin + (1 - alpha) * pout
    with np.errstate(all="ignore"):
        return np.where(den > 0, (1 - alpha) * pout / den, 0.0)


def gaussian_likelihood_quadrature(y, model, sigma_i, sigma):
    """GaussianLikelihoodQuadratureNoise.log_likelihood, likelihoods.py:767-790."""
    with np.errstate(all="ignore"):
        full = np.sqrt(np.asarray(sigma_i) ** 2. + sigma ** 2.)
    return float(np.nan_to_num(gaussian_log_likelihood(np.asarray(y) - model, full)))


# --------------------------------------------------------------------------------------------------
# kilonova: one-component model (kilonova_models.py:1537-1603, 1853-1893)
# --------------------------------------------------------------------------------------------------
def get_optimal_time_array(t_min, t_max, resolution, user_times=None):
    """utils.py:42-145 (the lru_cache is irrelevant for the values)."""
    if user_times is not None:
        user_times = np.atleast_1d(user_times)
        user_min, user_max = np.min(user_times), np.max(user_times)
        eval_min = max(t_min, user_min * 0.5)
        eval_max = min(t_max, user_max * 2.0)
        n_before = int(resolution * 0.15)
        n_user = int(resolution * 0.70)
        n_after = resolution - n_before - n_user
        if eval_min > t_min:
            segment_before = np.geomspace(t_min, eval_min, n_before)
        else:
            segment_before = np.array([t_min])
            n_user += n_before
        segment_user = np.geomspace(eval_min, eval_max, n_user)
        if eval_max < t_max:
            segment_after = np.geomspace(eval_max, t_max, n_after)
        else:
            segment_after = np.array([t_max])
            n_user += n_after
            segment_user = np.geomspace(eval_min, eval_max, n_user)
        return np.unique(np.concatenate([segment_before, segment_user, segment_after]))
    log_range = np.log10(t_max) - np.log10(t_min)
    if log_range > 6:
        n_early = int(resolution * 0.25)
        n_mid = int(resolution * 0.50)
        n_late = resolution - n_early - n_mid
        trans1 = t_min * (t_max / t_min) ** 0.25
        trans2 = t_min * (t_max / t_min) ** 0.75
        return np.unique(np.concatenate([np.geomspace(t_min, trans1, n_early), np.geomspace(trans1, trans2, n_mid),
                                         np.geomspace(trans2, t_max, n_late)]))
    return np.geomspace(t_min, t_max, resolution)


# --------------------------------------------------------------------------------------------------
# CSM interaction: _csm_engine (supernova_models.py:1910-1997), CSMDiffusion
# (interaction_processes.py:133-229), csm_interaction(_bolometric) (supernova_models.py:2001-2111)
# --------------------------------------------------------------------------------------------------
au_cgs = 14959787070000.0
_CSM_TABLE = None


def _csm_table():
    """redback/tables/csm_table.txt as committed by tools/gen_csm_table.py (tests/golden/csm_table.json)."""
    global _CSM_TABLE
    if _CSM_TABLE is None:
        import json
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "csm_table.json")
        with open(path) as fh:
            doc = json.load(fh)
        _CSM_TABLE = {k: np.asarray(v) for k, v in doc.items()}
    return _CSM_TABLE


def get_csm_properties(nn, eta):
    """utils.get_csm_properties (utils.py:288-315): bilinear RegularGridInterpolator on (n, s); out-of-range raises."""
    from scipy.interpolate import RegularGridInterpolator
    tab = _csm_table()
    out = {}
    for key in ("aa", "bf", "br"):
        out[key] = RegularGridInterpolator((tab["n"], tab["s"]), tab[key])([nn, eta])[0]
    return out["aa"], out["bf"], out["br"]


def csm_engine(time, mej, csm_mass, vej, eta, rho, kappa, r0, *, delta=1, nn=12, efficiency=0.5):
    """_csm_engine (supernova_models.py:1910-1997).  time: source-frame days.  Returns (lbol, r_photosphere,
    mass_csm_threshold) in cgs."""
    time = np.asarray(time, dtype=float)
    m_ej = mej * solar_mass
    m_csm = csm_mass * solar_mass
    r_in = r0 * au_cgs
    v = vej * km_cgs
    e_sn = 3. * v ** 2 * m_ej / 10.
    t_i = 1.
    a_c, b_f, b_r = get_csm_properties(nn, eta)
    with np.errstate(all="ignore"):
        q = rho * r_in ** eta                                                                       # :1942
        r_csm = ((3.0 - eta) / (4.0 * np.pi * q) * m_csm + r_in ** (3.0 - eta)) ** (1.0 / (3.0 - eta))
        r_ph = abs((-2.0 * (1.0 - eta) / (3.0 * kappa * q) + r_csm ** (1.0 - eta)) ** (1.0 / (1.0 - eta)))
        m_th = np.abs(4.0 * np.pi * q / (3.0 - eta) * (r_ph ** (3.0 - eta) - r_in ** (3.0 - eta)))   # :1949-1950
        g_n = (1.0 / (4.0 * np.pi * (nn - delta)) * (2.0 * (5.0 - delta) * (nn - 5.0) * e_sn) ** ((nn - 3.) / 2.0)
               / ((3.0 - delta) * (nn - 3.0) * m_ej) ** ((nn - 5.0) / 2.0))                         # :1953-1955
        pw = (nn - eta) / ((nn - 3.0) * (3.0 - eta))
        t_fs = (abs((3.0 - eta) * q ** ((3.0 - nn) / (nn - eta)) * (a_c * g_n) ** ((eta - 3.0) / (nn - eta))
                    / (4.0 * np.pi * b_f ** (3.0 - eta))) ** pw * m_th ** pw)                        # :1958-1962
        t_rs = (v / (b_r * (a_c * g_n / q) ** (1.0 / (nn - eta)))
                * (1.0 - (3.0 - nn) * m_ej / (4.0 * np.pi * v ** (3.0 - nn) * g_n)) ** (1.0 / (3.0 - nn))
                ) ** ((nn - eta) / (eta - 3.0))                                                     # :1965-1970
        ts = time * day_to_s
        alpha = (2.0 * nn + 6.0 * eta - nn * eta - 15.) / (nn - eta)
        l_fs = (2.0 * np.pi / (nn - eta) ** 3 * g_n ** ((5.0 - eta) / (nn - eta)) * q ** ((nn - 5.0) / (nn - eta))
                * (nn - 3.0) ** 2 * (nn - 5.0) * b_f ** (5.0 - eta) * a_c ** ((5.0 - eta) / (nn - eta))
                * (ts + t_i) ** alpha)                                                              # :1975-1977
        l_rs = (2.0 * np.pi * (a_c * g_n / q) ** ((5.0 - nn) / (nn - eta)) * b_r ** (5.0 - nn) * g_n
                * ((3.0 - eta) / (nn - eta)) ** 3 * (ts + t_i) ** alpha)                            # :1979-1980
        l_fs = np.where(t_fs - ts > 0, l_fs, 0.0)
        l_rs = np.where(t_rs - ts > 0, l_rs, 0.0)
    return efficiency * (l_fs + l_rs), r_ph, m_th


def csm_diffusion(time, dense_times, luminosity, kappa, r_photosphere, mass_csm_threshold):
    """CSMDiffusion._convert_input_luminosity_cumulative (interaction_processes.py:176-197, the default method)."""
    time = np.asarray(time, dtype=float)
    dense_times = np.asarray(dense_times, dtype=float)
    t0 = kappa * mass_csm_threshold / (4. * np.pi ** 3. / 9. * speed_of_light * r_photosphere) / day_to_s   # :171-173
    inside = time[(time >= dense_times[0]) & (time <= dense_times[-1])]
    knots = np.unique(np.concatenate((dense_times, inside)))
    lum = np.interp(knots, dense_times, luminosity)
    out = np.zeros_like(knots)
    with np.errstate(all="ignore"):
        dt = np.diff(knots)
        slope = np.diff(lum) / dt
        decay = np.exp(-dt / t0)
        fill = -np.expm1(-dt / t0)
        for i in range(dt.size):                                                                   # :190-195
            out[i + 1] = out[i] * decay[i] + lum[i] * fill[i] + slope[i] * (dt[i] - t0 * fill[i])
    return np.interp(time, knots, out)


def csm_diffusion_quadrature(time, dense_times, luminosity, kappa, r_photosphere, mass_csm_threshold, timesteps=3000):
    """CSMDiffusion._convert_input_luminosity_quadrature (interaction_processes.py:199-227; csm_diffusion_method =
    'quadrature' | 'legacy')."""
    time = np.asarray(time, dtype=float)
    dense_times = np.asarray(dense_times, dtype=float)
    t0 = kappa * mass_csm_threshold / (4. * np.pi ** 3. / 9. * speed_of_light * r_photosphere) / day_to_s   # :171-173
    with np.errstate(all="ignore"):
        tb = max(0.0, min(dense_times))                                                             # :205-206
        uniq = np.unique(time[(time >= tb) & (time <= dense_times[-1])])                            # :207
        lu = len(uniq)
        num = int(round(timesteps / 2.0))                                                           # :210
        log_start = min(np.log10(t0 / dense_times[-1]) + -3, 0.0)                                   # :211
        lsp = np.logspace(log_start, 0, num)
        xm = np.unique(np.concatenate((lsp, 1 - lsp)))                                              # :213
        xm = xm[(xm >= 0.0) & (xm <= 1.0)]
        xm = np.unique(np.concatenate(([0.0], xm, [1.0])))                                          # :215
        nodes = tb + (uniq.reshape(lu, 1) - tb) * xm                                                # :217
        lum = np.interp(nodes.ravel(), dense_times, luminosity).reshape(nodes.shape)                # :220
        integrand = lum * np.exp((nodes - uniq.reshape(lu, 1)) / t0)                                # :221
        integrand[np.isnan(integrand)] = 0.0
        out = np.trapezoid(integrand, nodes, axis=1) / t0                                           # :224-225
        return out[np.searchsorted(uniq, time)]


def csm_interaction_bolometric(time, mej, csm_mass, vej, eta, rho, kappa, r0, *, dense_resolution=1000,
                               csm_diffusion_method="cumulative", csm_diffusion_timesteps=3000, **kw):
    """supernova_models.py:2001-2050 with the default interaction process."""
    time = np.atleast_1d(np.asarray(time, dtype=float))
    dense = get_optimal_time_array(min(0.1, np.min(time)), np.max(time) + 100, dense_resolution, user_times=time)
    lbol, r_ph, m_th = csm_engine(dense, mej, csm_mass, vej, eta, rho, kappa, r0, **kw)
    if csm_diffusion_method in ("quadrature", "legacy"):
        return csm_diffusion_quadrature(time, dense, lbol, kappa, r_ph, m_th, csm_diffusion_timesteps)
    return csm_diffusion(time, dense, lbol, kappa, r_ph, m_th)


def csm_nickel_bolometric(time, mej, f_nickel, csm_mass, ek, eta, rho, kappa, r0, *, kappa_gamma, dense_resolution=1000,
                          csm_diffusion_method="cumulative", csm_diffusion_timesteps=3000, **kw):
    """supernova_models.csm_nickel_bolometric (:2120-2165): CSM shock + nickel diffused through mej + the optically
    thick CSM, both on the adaptive grid."""
    time = np.atleast_1d(np.asarray(time, dtype=float))
    vej = np.sqrt(2.0 * ek / (mej * solar_mass)) / km_cgs                                           # :2140
    dense = get_optimal_time_array(min(0.01, np.min(time)), np.max(time) + 100, dense_resolution, user_times=time)
    lbol, r_ph, m_th = csm_engine(dense, mej, csm_mass, vej, eta, rho, kappa, r0, **kw)
    mej_eff = mej + m_th / solar_mass                                                               # :2148
    nickel = diffusion(time, dense, nickelcobalt_engine(dense, f_nickel, mej), kappa, kappa_gamma, mej_eff, vej)[1]
    if csm_diffusion_method in ("quadrature", "legacy"):
        csm = csm_diffusion_quadrature(time, dense, lbol, kappa, r_ph, m_th, csm_diffusion_timesteps)
    else:
        csm = csm_diffusion(time, dense, lbol, kappa, r_ph, m_th)
    return nickel + csm                                                                             # :2165


def csm_nickel(time, redshift, mej, f_nickel, csm_mass, ek, eta, rho, kappa, r0, *, frequency, temperature_floor,
               kappa_gamma, dl=None, **kw):
    """supernova_models.csm_nickel (:2168-2200, flux_density): TemperatureFloor + Blackbody, mJy."""
    vej = np.sqrt(2.0 * ek / (mej * solar_mass)) / km_cgs
    return _sn_flux_density(
        time, redshift,
        lambda t: csm_nickel_bolometric(t, mej, f_nickel, csm_mass, ek, eta, rho, kappa, r0, kappa_gamma=kappa_gamma, **kw),
        "blackbody", frequency=frequency, vej=vej, temperature_floor=temperature_floor, dl=dl)


def csm_interaction(time, redshift, mej, csm_mass, vej, eta, rho, kappa, r0, *, frequency, temperature_floor,
                    dl=None, **kw):
    """supernova_models.py:2054-2092 (flux_density, TemperatureFloor + Blackbody defaults), mJy."""
    return _sn_flux_density(
        time, redshift,
        lambda t: csm_interaction_bolometric(t, mej, csm_mass, vej, eta, rho, kappa, r0, **kw),
        "blackbody", frequency=frequency, vej=vej, temperature_floor=temperature_floor, dl=dl)


# --------------------------------------------------------------------------------------------------
# Magnetar-driven relativistic ejecta: _ejecta_dynamics_and_interaction
# (magnetar_driven_ejecta_models.py:14-151), _comoving_blackbody_to_flux_density (:153-174),
# _process_flux_density (:240-272), basic_mergernova / general_mergernova(_thermalisation) (:275-420)
# --------------------------------------------------------------------------------------------------
radiation_constant = sigma_sb * 4 / speed_of_light       # constants.py:14


def magnetar_only(time, l0, tau, nn):
    """magnetar_models.magnetar_only (:134-144)."""
    return l0 * (1. + time / tau) ** ((1. + nn) / (1. - nn))


def ejecta_dynamics_and_interaction(time, mej, beta, ejecta_radius, kappa, n_ism, magnetar_luminosity,
                                    pair_cascade_switch, use_gamma_ray_opacity, *, thermalisation_efficiency=None,
                                    kappa_gamma=None, ejecta_albedo=0.5, pair_cascade_fraction=0.05,
                                    use_r_process=True, f_nickel=0, return_lorentz_factor=False):
    """_ejecta_dynamics_and_interaction (magnetar_driven_ejecta_models.py:14-151).
    time: source-frame seconds.  Returns (bolometric_luminosity, comoving_temperature, radius, doppler_factor)
    [+ lorentz_factor]."""
    m = mej * solar_mass
    n = len(time)
    out_l, out_t, out_r, out_d, out_g = (np.zeros(n) for _ in range(5))
    e_int = 0.5 * beta ** 2 * m * speed_of_light ** 2
    vol = (4 / 3) * np.pi * ejecta_radius ** 3
    gam = 1 / np.sqrt(1 - beta ** 2)
    rad = ejecta_radius
    d_gam = d_rad = d_vol = d_e = 0.0
    with np.errstate(all="ignore"):
        for i in range(n):
            beta = np.sqrt(1 - 1 / gam ** 2)                                                   # :65
            dop = 1 / (gam * (1 - beta))
            if i > 0:
                dt = time[i] - time[i - 1]
                gam = gam + d_gam * dt
                rad = rad + d_rad * dt
                vol = vol + d_vol * dt
                e_int = e_int + d_e * dt
            swept = (4 / 3) * np.pi * rad ** 3 * n_ism * proton_mass
            pressure = e_int / (3 * vol)
            t_com = dop * time[i]
            dvdt_com = 4 * np.pi * rad ** 2 * beta * speed_of_light
            denom = (1 / 2) - (1 / 3.141592654) * np.arctan((t_com - 1.3) / 0.11)              # :78
            if use_r_process:
                l_rad = (4 * 10 ** 49 * (m / (2 * 10 ** 33) * 10 ** 2) * denom ** 1.3)          # :80
            else:                                                                              # :82-84
                nickel_mass = f_nickel * m / solar_mass
                l_rad = nickel_mass * (6.45e43 * np.exp(-t_com / (8.8 * 86400)) + 1.45e43 * np.exp(-t_com / (111.3 * 86400)))
            tau = kappa * (m / vol) * (rad / gam)
            if tau <= 1:
                l_emit = (e_int * speed_of_light) / (rad / gam)
                temp = (e_int / (radiation_constant * vol)) ** (1. / 4.)
            else:
                l_emit = (e_int * speed_of_light) / (tau * rad / gam)
                temp = (e_int / (radiation_constant * vol * tau)) ** (1. / 4.)
            vej = speed_of_light * np.sqrt(1 - 1 / gam ** 2)                                   # utils.py:1524
            if use_gamma_ray_opacity:
                eff = 1 - np.exp(-(3 * kappa_gamma * m / (4 * np.pi * vej ** 2)) * time[i] ** -2)   # :102-104
            else:
                eff = thermalisation_efficiency
            d_rad = (beta * speed_of_light) / (1 - beta)
            d_swept = 4 * np.pi * rad ** 2 * n_ism * proton_mass * d_rad
            dedt = eff * magnetar_luminosity[i] + dop ** 2 * l_rad - dop ** 2 * l_emit
            de_com = eff * dop ** (-2) * magnetar_luminosity[i] + l_rad - l_emit - pressure * dvdt_com
            d_vol = dvdt_com * dop
            d_e = de_com * dop
            d_gam = (dedt - gam * dop * de_com - (gam ** 2 - 1) * speed_of_light ** 2 * d_swept) / (
                m * speed_of_light ** 2 + e_int + 2 * gam * swept * speed_of_light ** 2)
            out_g[i], out_l[i], out_t[i], out_r[i], out_d[i] = gam, l_emit * dop ** 2, temp, rad, dop
        if pair_cascade_switch:                                                                # :131-135
            v0 = ((1 / out_g) ** 2 + 1) ** 0.5 * speed_of_light
            tlife = ((0.6 / (1 - ejecta_albedo)) * (pair_cascade_fraction / 0.1) ** 0.5
                     * (magnetar_luminosity / 1.0e45) ** 0.5 * (v0 / (0.3 * speed_of_light)) ** (0.5)
                     * (time / 86400) ** (-0.5))
            out_l = out_l / (1.0 + tlife)
            out_t = (out_l / (4.0 * np.pi * out_r ** (2.0) * sigma_sb)) ** (0.25)
    if return_lorentz_factor:
        return out_l, out_t, out_r, out_d, out_g
    return out_l, out_t, out_r, out_d


def mergernova_flux_density(time, redshift, dynamics, *, frequency, dl=None):
    """_process_flux_density (:240-263) with use_relativistic_blackbody=True + _comoving_blackbody_to_flux_density
    (:153-174); `dynamics` = (time_temp, comoving_temperature, radius, doppler_factor).  mJy."""
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    t_grid, temp_g, rad_g, dop_g = dynamics
    frequency, t = calc_kcorrected_properties(np.asarray(frequency, dtype=float), redshift,
                                              np.asarray(time, dtype=float) * day_to_s)
    temp = interp1d(t_grid, temp_g)(t)
    rad = interp1d(t_grid, rad_g)(t)
    dop = interp1d(t_grid, dop_g)(t)
    with np.errstate(all="ignore"):
        num = 2 * np.pi * planck * frequency ** 3 * rad ** 2
        den = dl ** 2 * speed_of_light ** 2 * dop ** 2
        frac = 1. / (np.exp((planck * frequency) / (boltzmann_constant * temp * dop)) - 1)
        fd = num / den * frac * (1 + redshift)                                                   # :262
        return fd * _ERG_S_CM2_HZ_TO_MJY * (1 + redshift)                                        # :272


graviational_constant = 6.674299999999999e-08     # constants.py:20 (astropy G in cgs)


def evolving_gw_and_em_magnetar(time, bint, bext, p0, chi0, radius, moi):
    """magnetar_models._evolving_gw_and_em_magnetar (:187-224): Edot_d on `time` (explicit Euler of omega)."""
    epsilon_b = -3e-4 * (bint / bext) ** 2 * (bext / 1e16) ** 2
    omega_0 = 2.0 * np.pi / p0
    dt = time[1:] - time[:-1]
    omega = np.zeros_like(time)
    chi = chi0
    omega[0] = omega_0
    for i in range(len(time) - 1):
        omega[i + 1] = omega[i] + dt[i] * (
            -(bext ** 2 * radius ** 6 / (moi * speed_of_light ** 3)) * omega[i] ** 3 * (1. + np.sin(chi)) - (
                2.0 * graviational_constant * moi * epsilon_b ** 2 / (5.0 * speed_of_light ** 5)) * omega[i] ** 5
            * np.sin(chi) ** 2 * (1.0 + 15.0 * np.sin(chi) ** 2))
    return (bext ** 2 * radius ** 6 / (4 * speed_of_light ** 3)) * omega ** 4 * (1.0 + np.sin(chi) ** 2)


def _evolving(logbint, logbext, p0, chi0, radius, logmoi):
    return lambda t: evolving_gw_and_em_magnetar(t, 10 ** logbint, 10 ** logbext, p0, chi0, radius * km_cgs, 10 ** logmoi)


def general_mergernova_evolution(time, redshift, mej, beta, ejecta_radius, kappa, n_ism, logbint, logbext, p0, chi0,
                                 radius, logmoi, kappa_gamma, *, frequency, pair_cascade_switch=True, dl=None, **kw):
    """general_mergernova_evolution (:422-474), flux_density."""
    return _mergernova(time, redshift, _evolving(logbint, logbext, p0, chi0, radius, logmoi), mej, beta, ejecta_radius,
                       kappa, n_ism, frequency=frequency, pair_cascade_switch=pair_cascade_switch, dl=dl,
                       use_gamma_ray_opacity=True, kappa_gamma=kappa_gamma, **kw)


def _magnetar_driven_supernova_dynamics(mej, E_sn, kappa, l0, tau_sd, nn, kappa_gamma, f_nickel, pair_cascade_switch):
    """Shared part of general_magnetar_driven_supernova(_bolometric) (supernova_models.py:2476-2488, 2560-2570)."""
    t_grid = get_optimal_time_array(1e0, 1e8, 2000)
    beta = np.sqrt(E_sn / (0.5 * mej * solar_mass)) / speed_of_light
    lbol, _, _, _, gam = ejecta_dynamics_and_interaction(
        t_grid, mej, beta, 1.0e11, kappa, 1.0e-5, magnetar_only(t_grid, l0, tau_sd, nn), pair_cascade_switch, True,
        kappa_gamma=kappa_gamma, use_r_process=False, f_nickel=f_nickel, return_lorentz_factor=True)
    with np.errstate(all="ignore"):
        vej = speed_of_light * np.sqrt(1 - 1 / gam ** 2) / km_cgs                                # utils.py:1524
    return t_grid, lbol, vej


def general_magnetar_driven_supernova_bolometric(time, mej, E_sn, kappa, l0, tau_sd, nn, kappa_gamma, *, f_nickel=0,
                                                 pair_cascade_switch=False, **_):
    """supernova_models.general_magnetar_driven_supernova_bolometric (:2464-2507): erg/s at source-frame days."""
    t_grid, lbol, _ = _magnetar_driven_supernova_dynamics(mej, E_sn, kappa, l0, tau_sd, nn, kappa_gamma, f_nickel,
                                                          pair_cascade_switch)
    return interp1d(t_grid, y=lbol)(np.asarray(time, dtype=float) * day_to_s)


def general_magnetar_driven_supernova(time, redshift, mej, E_sn, kappa, l0, tau_sd, nn, kappa_gamma, *, frequency,
                                      temperature_floor, f_nickel=0, pair_cascade_switch=False,
                                      cutoff_wavelength=3000, absorption_index=1.0, dl=None, **_):
    """supernova_models.general_magnetar_driven_supernova (:2510-2584), flux_density, CutoffBlackbody: the photosphere
    is evaluated on the dynamics grid with the time-dependent ejecta velocity, then T, R, L are interpolated."""
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    frequency, t = calc_kcorrected_properties(np.asarray(frequency, dtype=float), redshift, np.asarray(time, dtype=float))
    t_grid, lbol_g, vej = _magnetar_driven_supernova_dynamics(mej, E_sn, kappa, l0, tau_sd, nn, kappa_gamma, f_nickel,
                                                              pair_cascade_switch)
    days = t_grid / day_to_s
    with np.errstate(all="ignore"):
        temp_g, rad_g = temperature_floor_grid = globals()["temperature_floor"](days, lbol_g, vej, temperature_floor)
        temp, rad, lbol = (interp1d(days, y=v)(t) for v in (temp_g, rad_g, lbol_g))
        freq = frequency if np.ndim(frequency) else float(frequency)
        fd = cutoff_blackbody_mjy(t, temp, lbol, rad, freq, dl, cutoff_wavelength, absorption_index)
    return fd * (1 + redshift)


def trapped_magnetar(time, redshift, mej, beta, ejecta_radius, kappa, n_ism, l0, tau_sd, nn,
                     thermalisation_efficiency, *, frequency, output_format="flux", photon_index=2.0, dl=None):
    """magnetar_driven_ejecta_models.trapped_magnetar (:477-573).  time: seconds."""
    time = np.asarray(time, dtype=float)
    frequency = np.asarray(frequency, dtype=float)
    if output_format == "flux":
        frequency, time = calc_kcorrected_properties(frequency, redshift, time)                  # :537
    t_grid = get_optimal_time_array(1e-4, 1e8, 500)
    m = mej * solar_mass
    # optical depth series (:85): recomputed from the dynamics outputs is not possible, so march again with tau kept
    lum_g = magnetar_only(t_grid, l0, tau_sd, nn)
    _, temp_g, rad_g, dop_g = ejecta_dynamics_and_interaction(
        t_grid, mej, beta, ejecta_radius, kappa, n_ism, lum_g, False, False,
        thermalisation_efficiency=thermalisation_efficiency)
    tau_g = _ejecta_optical_depth(t_grid, mej, beta, ejecta_radius, kappa, n_ism, lum_g, thermalisation_efficiency)
    temp, rad, dop = (interp1d(t_grid, y=v)(time) for v in (temp_g, rad_g, dop_g))
    optical_depth = interp1d(t_grid, y=tau_g)(time)
    with np.errstate(all="ignore"):
        num = 8 * np.pi ** 2 * planck * frequency ** 4 * rad ** 2                               # :190-195
        den = speed_of_light ** 2 * dop ** 2
        l_ej = num / den * (1. / (np.exp((planck * frequency) / (boltzmann_constant * temp * dop)) - 1))
        lum = np.exp(-optical_depth) * magnetar_only(time, l0, tau_sd, nn) + l_ej                 # :509-511
    if output_format == "luminosity":
        return lum
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    return lum / (4 * np.pi * dl ** 2 * (1. + redshift) ** (photon_index - 2))                    # :544-545


def _ejecta_optical_depth(time, mej, beta, ejecta_radius, kappa, n_ism, magnetar_luminosity, thermalisation_efficiency):
    """The `tau` series of _ejecta_dynamics_and_interaction (:85, :124): the same march as
    ejecta_dynamics_and_interaction, returning kappa (m / V') (R / gamma) per step."""
    m = mej * solar_mass
    n = len(time)
    out = np.zeros(n)
    e_int = 0.5 * beta ** 2 * m * speed_of_light ** 2
    vol = (4 / 3) * np.pi * ejecta_radius ** 3
    gam = 1 / np.sqrt(1 - beta ** 2)
    rad = ejecta_radius
    d_gam = d_rad = d_vol = d_e = 0.0
    with np.errstate(all="ignore"):
        for i in range(n):
            beta = np.sqrt(1 - 1 / gam ** 2)
            dop = 1 / (gam * (1 - beta))
            if i > 0:
                dt = time[i] - time[i - 1]
                gam, rad, vol, e_int = gam + d_gam * dt, rad + d_rad * dt, vol + d_vol * dt, e_int + d_e * dt
            swept = (4 / 3) * np.pi * rad ** 3 * n_ism * proton_mass
            pressure = e_int / (3 * vol)
            t_com = dop * time[i]
            dvdt_com = 4 * np.pi * rad ** 2 * beta * speed_of_light
            denom = (1 / 2) - (1 / 3.141592654) * np.arctan((t_com - 1.3) / 0.11)
            l_rad = (4 * 10 ** 49 * (m / (2 * 10 ** 33) * 10 ** 2) * denom ** 1.3)
            tau = kappa * (m / vol) * (rad / gam)
            l_emit = (e_int * speed_of_light) / (rad / gam) if tau <= 1 else \
                (e_int * speed_of_light) / (tau * rad / gam)
            eff = thermalisation_efficiency
            d_rad = (beta * speed_of_light) / (1 - beta)
            d_swept = 4 * np.pi * rad ** 2 * n_ism * proton_mass * d_rad
            dedt = eff * magnetar_luminosity[i] + dop ** 2 * l_rad - dop ** 2 * l_emit
            de_com = eff * dop ** (-2) * magnetar_luminosity[i] + l_rad - l_emit - pressure * dvdt_com
            d_vol = dvdt_com * dop
            d_e = de_com * dop
            d_gam = (dedt - gam * dop * de_com - (gam ** 2 - 1) * speed_of_light ** 2 * d_swept) / (
                m * speed_of_light ** 2 + e_int + 2 * gam * swept * speed_of_light ** 2)
            out[i] = tau
    return out


def _mergernova(time, redshift, mag_lum_fn, mej, beta, ejecta_radius, kappa, n_ism, *, frequency, pair_cascade_switch,
                dl=None, **dyn):
    t_grid = get_optimal_time_array(1e-4, 1e8, 500)
    lum, temp, rad, dop = ejecta_dynamics_and_interaction(t_grid, mej, beta, ejecta_radius, kappa, n_ism,
                                                          mag_lum_fn(t_grid), pair_cascade_switch, **dyn)
    return mergernova_flux_density(time, redshift, (t_grid, temp, rad, dop), frequency=frequency, dl=dl)


def basic_mergernova(time, redshift, mej, beta, ejecta_radius, kappa, n_ism, p0, bp, mass_ns, theta_pb,
                     thermalisation_efficiency, *, frequency, pair_cascade_switch=False, dl=None, **kw):
    """magnetar_driven_ejecta_models.basic_mergernova (:275-321), flux_density."""
    return _mergernova(time, redshift, lambda t: basic_magnetar(t, p0, bp, mass_ns, theta_pb), mej, beta,
                       ejecta_radius, kappa, n_ism, frequency=frequency, pair_cascade_switch=pair_cascade_switch,
                       dl=dl, use_gamma_ray_opacity=False, thermalisation_efficiency=thermalisation_efficiency, **kw)


def general_mergernova(time, redshift, mej, beta, ejecta_radius, kappa, n_ism, l0, tau_sd, nn,
                       thermalisation_efficiency, *, frequency, pair_cascade_switch=True, dl=None, **kw):
    """general_mergernova (:324-371), flux_density."""
    return _mergernova(time, redshift, lambda t: magnetar_only(t, l0, tau_sd, nn), mej, beta, ejecta_radius, kappa,
                       n_ism, frequency=frequency, pair_cascade_switch=pair_cascade_switch, dl=dl,
                       use_gamma_ray_opacity=False, thermalisation_efficiency=thermalisation_efficiency, **kw)


def general_mergernova_thermalisation(time, redshift, mej, beta, ejecta_radius, kappa, n_ism, l0, tau_sd, nn,
                                      kappa_gamma, *, frequency, pair_cascade_switch=True, dl=None, **kw):
    """general_mergernova_thermalisation (:374-419), flux_density."""
    return _mergernova(time, redshift, lambda t: magnetar_only(t, l0, tau_sd, nn), mej, beta, ejecta_radius, kappa,
                       n_ism, frequency=frequency, pair_cascade_switch=pair_cascade_switch, dl=dl,
                       use_gamma_ray_opacity=True, kappa_gamma=kappa_gamma, **kw)


_BK_V = np.asarray([0.1, 0.2, 0.3, 0.4])
_BK_M = np.asarray([1.e-3, 5.e-3, 1.e-2, 5.e-2, 1.e-1])
_BK_A = np.asarray([[2.01, 4.52, 8.16, 16.3], [0.81, 1.9, 3.2, 5.0], [0.56, 1.31, 2.19, 3.0],
                    [.27, .55, .95, 2.0], [0.20, 0.39, 0.65, 0.9]])
_BK_B = np.asarray([[0.28, 0.62, 1.19, 2.4], [0.19, 0.28, 0.45, 0.65], [0.17, 0.21, 0.31, 0.45],
                    [0.10, 0.13, 0.15, 0.17], [0.06, 0.11, 0.12, 0.12]])
_BK_D = np.asarray([[1.12, 1.39, 1.52, 1.65], [0.86, 1.21, 1.39, 1.5], [0.74, 1.13, 1.32, 1.4],
                    [0.6, 0.9, 1.13, 1.25], [0.63, 0.79, 1.04, 1.5]])


def barnes_kasen_thermalisation(mej, vej):
    """utils.py:1161-1187: three bilinear RegularGridInterpolators with linear extrapolation."""
    from scipy.interpolate import RegularGridInterpolator
    out = []
    for arr in (_BK_A, _BK_B, _BK_D):
        f = RegularGridInterpolator((_BK_M, _BK_V), arr, bounds_error=False, fill_value=None)
        out.append(f([mej, vej])[0])
    return tuple(out)


def one_component_kilonova_internal(time, mej, vej, kappa, temperature_floor=4000):
    """kilonova_models.py:1853-1893.  time: source-frame seconds.  Returns (L_bol, T, R)."""
    from scipy.integrate import cumulative_trapezoid
    with np.errstate(all="ignore"):
        tdays = time / day_to_s
        av, bv, dv = barnes_kasen_thermalisation(mej, vej)
        e_th = 0.36 * (np.exp(-av * tdays) + np.log1p(2.0 * bv * tdays ** dv) / (2.0 * bv * tdays ** dv))
        t0, sig, beta = 1.3, 0.11, 13.7
        v0 = vej * speed_of_light
        m0 = mej * solar_mass
        tdiff = np.sqrt(2.0 * kappa * (m0) / (beta * v0 * speed_of_light))
        lum_in = 4.0e18 * (m0) * (0.5 - np.arctan((time - t0) / sig) / np.pi) ** 1.3
        integrand = lum_in * e_th * (time / tdiff) * np.exp(time ** 2 / tdiff ** 2)
        lbol = np.zeros(len(time))
        lbol[1:] = cumulative_trapezoid(integrand, time)
        lbol[0] = lbol[1]
        lbol = lbol * np.exp(-time ** 2 / tdiff ** 2) / tdiff
        temperature = (lbol / (4.0 * np.pi * sigma_sb * v0 ** 2 * time ** 2)) ** 0.25
        r_photosphere = (lbol / (4.0 * np.pi * sigma_sb * temperature_floor ** 4)) ** 0.5
        mask = temperature <= temperature_floor
        temperature[mask] = temperature_floor
        mask = np.logical_not(mask)
        r_photosphere[mask] = v0 * time[mask]
        return lbol, temperature, r_photosphere


def one_component_kilonova_model(time, redshift, mej, vej, kappa, *, frequency, temperature_floor=4000,
                                 dense_resolution=500, dl=None, **_):
    """kilonova_models.py:1537-1577 (flux_density branch).  time: observer-frame days.  Returns mJy[E]."""
    time = np.asarray(time, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    t_src_s = time * day_to_s / (1. + redshift)
    time_temp = get_optimal_time_array(1e-2, 7e6, dense_resolution, user_times=t_src_s)
    _, temperature, r_photosphere = one_component_kilonova_internal(time_temp, mej, vej, kappa, temperature_floor)
    with np.errstate(all="ignore"):
        t = time * day_to_s
        temp_func = interp1d(time_temp, y=temperature)
        rad_func = interp1d(time_temp, y=r_photosphere)
        frequency, t = calc_kcorrected_properties(np.asarray(frequency, dtype=float), redshift, t)
        fd = blackbody_to_flux_density(temp_func(t), rad_func(t), dl, frequency)
        return fd * _ERG_S_CM2_HZ_TO_MJY * (1 + redshift)


def compactness_from_lambda(lam):
    """ejecta_relations.py:435-444"""
    return 0.371 - 0.0391 * np.log(lam) + 0.001056 * np.log(lam) ** 2


def baryonic_mass(mass, compactness):
    """ejecta_relations.py:473-482"""
    return mass * (1 + 0.8857853174243745 * compactness ** 1.2082383572002926)


def bns_ejecta_no_projection(mass_1, mass_2, lambda_1, lambda_2):
    """OneComponentBNSNoProjection (ejecta_relations.py:5-65) -> (mej [Msun], vej [c])."""
    c1, c2 = compactness_from_lambda(lambda_1), compactness_from_lambda(lambda_2)
    vej = -0.3090 * (mass_1 / mass_2) * (1 + -1.879 * c1) + -0.3090 * (mass_2 / mass_1) * (1 + -1.879 * c2) + 0.657
    log10_mej = -0.0719 * (mass_1 * (1 - 2 * c1) / c1 + mass_2 * (1 - 2 * c2) / c2) + 0.2116 * \
        (mass_1 * (mass_2 / mass_1) ** -2.905 + mass_2 * (mass_1 / mass_2) ** -2.905) + -2.42
    return 10 ** log10_mej, vej


def bns_ejecta_projection(mass_1, mass_2, lambda_1, lambda_2):
    """OneComponentBNSProjection (ejecta_relations.py:88-147) with calc_vrho / calc_vz (:484-525)."""
    c1, c2 = compactness_from_lambda(lambda_1), compactness_from_lambda(lambda_2)
    q12, q21 = mass_1 / mass_2, mass_2 / mass_1
    vrho = (q12 * (1.0 + -2.67385 * c1) + q21 * (1.0 + -2.67385 * c2)) * -0.219479 + 0.444836
    vz = (q12 * (1.0 + -1.00757 * c1) + q21 * (1.0 + -1.00757 * c2)) * -0.315585 + 0.63808
    mb1, mb2 = baryonic_mass(mass_1, c1), baryonic_mass(mass_2, c2)
    tmp1 = ((mb1 * (q21 ** (1.0 / 3.0)) * (1.0 - 2.0 * c1) / c1) + (mb2 * (q12 ** (1.0 / 3.0)) * (1.0 - 2.0 * c2) / c2)) \
        * -1.35695
    tmp2 = (mb1 * (q21 ** -2.5484) + mb2 * (q12 ** -2.5484)) * 6.11252
    tmp3 = (mb1 * (1.0 - mass_1 / mb1) + mb2 * (1.0 - mass_2 / mb2)) * -49.43355
    return np.maximum(tmp1 + tmp2 + tmp3 + 16.1144, 0) / 1000.0, (vrho ** 2.0 + vz ** 2.0) ** 0.5


def nsbh_ejecta(mass_bh, mass_ns, chi_bh, lambda_ns):
    """OneComponentNSBH (ejecta_relations.py:380-432)."""
    q = mass_bh / mass_ns
    z1 = 1 + ((1 - chi_bh * chi_bh) ** (1 / 3.0)) * (((1 + chi_bh) ** (1 / 3.0)) + (1 - chi_bh) ** (1 / 3.0))
    z2 = (3 * chi_bh * chi_bh + z1 * z1) ** (1 / 2.0)
    risco = 3 + z2 - np.sign(chi_bh) * ((3 - z1) * (3 + z1 + 2 * z2)) ** (1 / 2.0)
    comp = compactness_from_lambda(lambda_ns)
    mb = baryonic_mass(mass_ns, comp)
    tmp = risco * (q ** 1.352) * -2.269e-3 + (q ** 0.2497) * (1 - 2 * comp) * 4.464e-2 / comp \
        + (1 - mass_ns / mb) * 2.431 + -0.4159
    return mb * np.maximum(tmp, 0), 1.5333330951369120e-2 * q + 0.19066667068621043


def bns_two_component_ejecta(mass_1, mass_2, lambda_1, lambda_2, mtov, zeta):
    """TwoComponentBNS (ejecta_relations.py:170-267) -> (dynamical mej, disk-wind mej [Msun], vej [c])."""
    mej_dyn, vej = bns_ejecta_no_projection(mass_1, mass_2, lambda_1, lambda_2)          # same fits (:201-238)
    q = mass_1 / mass_2
    lambdatilde = (16.0 / 13.0) * (lambda_2 + lambda_1 * (q ** 5) + 12 * lambda_1 * (q ** 4) + 12 * lambda_2 * q) \
        / ((q + 1) ** 5)
    mc = ((mass_1 * mass_2) ** (3. / 5.)) * ((mass_1 + mass_2) ** (-1. / 5.))
    r16 = mc * (lambdatilde / 0.0042) ** (1.0 / 6.0)
    mth = (2.38 - 3.606 * mtov / r16) * mtov
    mdisk = 10 ** (-31.335 * (1 + -0.9760 * np.tanh((1.0474 - (mass_1 + mass_2) / mth) / 0.05957)))
    return mej_dyn, zeta * mdisk, vej


def nsbh_two_component_ejecta(mass_bh, mass_ns, chi_bh, lambda_ns, zeta):
    """TwoComponentNSBH (ejecta_relations.py:290-378)."""
    q = mass_bh / mass_ns
    z1 = 1 + (1 - chi_bh ** 2) ** (1 / 3) * ((1 + chi_bh) ** (1 / 3) + (1 - chi_bh) ** (1 / 3))
    z2 = np.sqrt(3 * chi_bh ** 2 + z1 ** 2)
    risco = 3 + z2 - np.sign(chi_bh) * np.sqrt((3 - z1) * (3 + z1 + 2 * z2))
    comp = compactness_from_lambda(lambda_ns)
    mb = baryonic_mass(mass_ns, comp)
    dyn = mb * np.maximum(4.464e-2 * (q ** 0.2497) * (1 - 2 * comp) / comp + -2.269e-3 * (q ** 1.352) * risco
                          + 2.431 * (1 - mass_ns / mb) + -0.4159, 0)
    rho = (15 * lambda_ns) ** (-1 / 5)
    eta = q / (1 + q) ** 2
    disk = mb * (np.maximum(0.308 * (1 - 2 * rho) / (eta ** (1 / 3)) + -0.124 * risco * rho / eta + 0.283, 0.0)) ** 1.536
    return dyn, zeta * disk, 1.5333330951369120e-2 * q + 0.19066667068621043


def two_component_bns_ejecta_relation(time, redshift, mass_1, mass_2, lambda_1, lambda_2, mtov, zeta, vej_2, kappa_1,
                                      kappa_2, tf_1, tf_2, **kw):
    """kilonova_models.py:1369-1397"""
    mej_1, mej_2, vej_1 = bns_two_component_ejecta(mass_1, mass_2, lambda_1, lambda_2, mtov, zeta)
    return two_component_kilonova_model(time, redshift, mej_1, vej_1, tf_1, kappa_1, mej_2, vej_2, tf_2, kappa_2, **kw)


def two_component_nsbh_ejecta_relation(time, redshift, mass_bh, mass_ns, chi_bh, lambda_ns, zeta, vej_2, kappa_1,
                                       kappa_2, tf_1, tf_2, **kw):
    """kilonova_models.py:1496-1527"""
    mej_1, mej_2, vej_1 = nsbh_two_component_ejecta(mass_bh, mass_ns, chi_bh, lambda_ns, zeta)
    return two_component_kilonova_model(time, redshift, mej_1, vej_1, tf_1, kappa_1, mej_2, vej_2, tf_2, kappa_2, **kw)


def one_component_ejecta_relation(time, redshift, mass_1, mass_2, lambda_1, lambda_2, kappa, **kw):
    """kilonova_models.py:1309-1336"""
    mej, vej = bns_ejecta_no_projection(mass_1, mass_2, lambda_1, lambda_2)
    return one_component_kilonova_model(time, redshift, mej, vej, kappa, **kw)


def one_component_ejecta_relation_projection(time, redshift, mass_1, mass_2, lambda_1, lambda_2, kappa, **kw):
    """kilonova_models.py:1339-1366"""
    mej, vej = bns_ejecta_projection(mass_1, mass_2, lambda_1, lambda_2)
    return one_component_kilonova_model(time, redshift, mej, vej, kappa, **kw)


def one_component_nsbh_ejecta_relation(time, redshift, mass_bh, mass_ns, chi_bh, lambda_ns, kappa, **kw):
    """kilonova_models.py:1470-1493"""
    mej, vej = nsbh_ejecta(mass_bh, mass_ns, chi_bh, lambda_ns)
    return one_component_kilonova_model(time, redshift, mej, vej, kappa, **kw)


def multi_component_kilonova_model(time, redshift, components, *, frequency, dense_resolution, grid_t_max, dl=None):
    """two_component_kilonova_model (kilonova_models.py:1214-1260: grid end 86400*6 s, 500 points) and
    three_component_kilonova_model (:1113-1165: 7e6 s, 300 points), flux_density branch: the components'
    blackbody flux densities are summed.  components: [(mej, vej, kappa, temperature_floor), ...]."""
    time = np.asarray(time, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    time_temp = get_optimal_time_array(1e-2, grid_t_max, dense_resolution, user_times=time * day_to_s / (1. + redshift))
    frequency, t = calc_kcorrected_properties(np.asarray(frequency, dtype=float), redshift, time * day_to_s)
    ff = np.zeros(len(time))
    with np.errstate(all="ignore"):
        for mej, vej, kappa, tfloor in components:
            _, temperature, r_photosphere = one_component_kilonova_internal(time_temp, mej, vej, kappa, tfloor)
            ff += blackbody_to_flux_density(interp1d(time_temp, y=temperature)(t), interp1d(time_temp, y=r_photosphere)(t),
                                            dl, frequency)
        return ff * _ERG_S_CM2_HZ_TO_MJY * (1 + redshift)


def two_component_kilonova_model(time, redshift, mej_1, vej_1, temperature_floor_1, kappa_1, mej_2, vej_2,
                                 temperature_floor_2, kappa_2, *, frequency, dense_resolution=500, dl=None, **_):
    comps = [(mej_1, vej_1, kappa_1, temperature_floor_1), (mej_2, vej_2, kappa_2, temperature_floor_2)]
    return multi_component_kilonova_model(time, redshift, comps, frequency=frequency,
                                          dense_resolution=dense_resolution, grid_t_max=86400 * 6, dl=dl)


def three_component_kilonova_model(time, redshift, mej_1, vej_1, temperature_floor_1, kappa_1, mej_2, vej_2,
                                   temperature_floor_2, kappa_2, mej_3, vej_3, temperature_floor_3, kappa_3, *,
                                   frequency, dense_resolution=300, dl=None, **_):
    comps = [(mej_1, vej_1, kappa_1, temperature_floor_1), (mej_2, vej_2, kappa_2, temperature_floor_2),
             (mej_3, vej_3, kappa_3, temperature_floor_3)]
    return multi_component_kilonova_model(time, redshift, comps, frequency=frequency,
                                          dense_resolution=dense_resolution, grid_t_max=7e6, dl=dl)


# --------------------------------------------------------------------------------------------------
# kilonova: Metzger (2017) shell model (kilonova_models.py:1895-1962, 1965-2068)
# --------------------------------------------------------------------------------------------------
def electron_fraction_from_kappa(kappa):
    """utils.py:1479-1491"""
    kappa_array = np.array([35, 32.2, 22.3, 5.60, 5.36, 3.30, 0.96, 0.5])
    ye_array = np.array([0.10, 0.15, 0.2, 0.25, 0.30, 0.35, 0.4, 0.5])
    return interp1d(kappa_array, y=ye_array, fill_value='extrapolate')(kappa)


def metzger_kilonova_internal(time, mej, vej, beta, kappa, neutron_precursor_switch=True, vmax=0.7):
    """kilonova_models.py:1965-2068.  time: source-frame seconds.  Returns (L_bol, T, R)."""
    with np.errstate(all="ignore"):
        tdays = time / day_to_s
        time_len = len(time)
        mass_len = 200
        av, bv, dv = barnes_kasen_thermalisation(mej, vej)
        e_th = 0.36 * (np.exp(-av * tdays) + np.log1p(2.0 * bv * tdays ** dv) / (2.0 * bv * tdays ** dv))
        electron_fraction = electron_fraction_from_kappa(kappa)
        t0, sig, tau_neutron = 1.3, 0.11, 900
        vel = np.linspace(vej, vmax, mass_len)
        m_array = mej * (vel / vej) ** (-beta)
        v_m = vel * speed_of_light
        time_array = np.tile(time, (mass_len, 1))
        e_th_array = np.tile(e_th, (mass_len, 1))
        edotr = np.zeros((mass_len, time_len))
        time_mask = time > t0
        time_1 = time_array[:, time_mask]
        time_2 = time_array[:, ~time_mask]
        edotr[:, time_mask] = 2.1e10 * e_th_array[:, time_mask] * ((time_1 / (3600. * 24.)) ** (-1.3))
        edotr[:, ~time_mask] = 4.0e18 * (0.5 - (1. / np.pi) * np.arctan((time_2 - t0) / sig)) ** (1.3) * \
            e_th_array[:, ~time_mask]
        energy_v = np.zeros((mass_len, time_len))
        lum_rad = np.zeros((mass_len, time_len))
        td_v = np.zeros((mass_len, time_len))
        tau = np.zeros((mass_len, time_len))
        v_photosphere = np.zeros(time_len)
        r_photosphere = np.zeros(time_len)
        if neutron_precursor_switch:
            neutron_mass = 1e-4 * solar_mass
            xn0 = (2.0 / np.pi) * (1.0 - electron_fraction) * np.arctan(neutron_mass / m_array / solar_mass)
            xr = 1.0 - xn0
            xn0_arr = np.tile(xn0, (time_len, 1)).T
            xr_arr = np.tile(xr, (time_len, 1)).T
            xn_arr = xn0_arr * np.exp(-time_array / tau_neutron)
            edotr = 3.2e14 * xn_arr + edotr
            kappa = 0.4 * (1.0 - xn_arr - xr_arr) + kappa * xr_arr
        dt = np.diff(time)
        dm = np.abs(np.diff(m_array))
        energy_v[:, 0] = 0.5 * m_array * v_m ** 2
        for ii in range(time_len - 1):
            kap = kappa[:-1, ii] if neutron_precursor_switch else kappa
            td_v[:-1, ii] = (kap * m_array[:-1] * solar_mass * 3) / (4 * np.pi * v_m[:-1] * speed_of_light * time[ii] * beta)
            tau[:-1, ii] = (m_array[:-1] * solar_mass * kap / (4 * np.pi * (time[ii] * v_m[:-1]) ** 2))
            lum_rad[:-1, ii] = energy_v[:-1, ii] / (td_v[:-1, ii] + time[ii] * (v_m[:-1] / speed_of_light))
            energy_v[:-1, ii + 1] = (edotr[:-1, ii] - (energy_v[:-1, ii] / time[ii]) - lum_rad[:-1, ii]) * dt[ii] \
                + energy_v[:-1, ii]
            lum_rad[:-1, ii] = lum_rad[:-1, ii] * dm * solar_mass
            tau[mass_len - 1, ii] = tau[mass_len - 2, ii]
            photosphere_index = np.argmin(np.abs(tau[:, ii] - 1))
            v_photosphere[ii] = v_m[photosphere_index]
            r_photosphere[ii] = v_photosphere[ii] * time[ii]
        lbol = np.sum(lum_rad, axis=0)
        temperature = (lbol / (4.0 * np.pi * (r_photosphere) ** (2.0) * sigma_sb)) ** (0.25)
        return lbol, temperature, r_photosphere


def metzger_magnetar_driven_internal(time, mej, vej, beta, kappa, magnetar_luminosity, use_gamma_ray_opacity, *,
                                     thermalisation_efficiency=None, kappa_gamma=None, pair_cascade_switch=True,
                                     ejecta_albedo=0.5, pair_cascade_fraction=0.01, neutron_precursor_switch=True,
                                     vmax=0.7, magnetar_heating="first_layer"):
    """_general_metzger_magnetar_driven_kilonova_model (magnetar_driven_ejecta_models.py:576-757); magnetar_heating
    'first_layer' (default: only the innermost shell is heated by the engine) or 'all_layers' (:698-704).
    time: source-frame seconds.  Returns (L_bol, T, R_photosphere)."""
    if magnetar_heating not in ("first_layer", "all_layers"):
        raise ValueError("magnetar_heating must be 'first_layer' or 'all_layers'")
    with np.errstate(all="ignore"):
        tdays = time / day_to_s
        time_len, mass_len = len(time), 200
        av, bv, dv = barnes_kasen_thermalisation(mej, vej)
        e_th = 0.36 * (np.exp(-av * tdays) + np.log1p(2.0 * bv * tdays ** dv) / (2.0 * bv * tdays ** dv))
        electron_fraction = electron_fraction_from_kappa(kappa)
        t0, sig, tau_neutron = 1.3, 0.11, 900
        m0 = mej * solar_mass
        v0 = vej * speed_of_light
        kinetic_energy = 0.5 * m0 * v0 ** 2
        vel = np.linspace(vej, vmax, mass_len)
        m_array = mej * (vel / vej) ** (-beta)
        m_array = m_array * (mej / np.sum(m_array))                                     # :617-619
        v_m = vel * speed_of_light
        time_array = np.tile(time, (mass_len, 1))
        e_th_array = np.tile(e_th, (mass_len, 1))
        edotr = np.zeros((mass_len, time_len))
        mask = time > t0
        edotr[:, mask] = 2.1e10 * e_th_array[:, mask] * ((time_array[:, mask] / (3600. * 24.)) ** (-1.3))
        edotr[:, ~mask] = 4.0e18 * (0.5 - (1. / np.pi) * np.arctan((time_array[:, ~mask] - t0) / sig)) ** (1.3) * \
            e_th_array[:, ~mask]
        lsd = magnetar_luminosity
        energy_v = np.zeros((mass_len, time_len))
        lum_rad = np.zeros((mass_len, time_len))
        td_v = np.zeros((mass_len, time_len))
        tau = np.zeros((mass_len, time_len))
        r_photosphere = np.zeros(time_len)
        if neutron_precursor_switch:
            xn0 = (2.0 / np.pi) * (1.0 - electron_fraction) * np.arctan(1e-4 * solar_mass / m_array / solar_mass)
            xr = 1.0 - xn0
            xn_arr = np.tile(xn0, (time_len, 1)).T * np.exp(-time_array / tau_neutron)
            xr_arr = np.tile(xr, (time_len, 1)).T
            edotr = 3.2e14 * xn_arr + edotr
            kappa = 0.4 * (1.0 - xn_arr - xr_arr) + kappa * xr_arr
        dt = np.diff(time)
        dm = np.abs(np.diff(m_array))
        energy_v[:, 0] = 0.5 * m_array * v_m ** 2
        for ii in range(time_len - 1):
            kinetic_energy = kinetic_energy + (energy_v[0, ii] / time[ii]) * dt[ii]            # :667
            v0 = (2 * kinetic_energy / m0) ** 0.5
            v_m = v0 * (m_array / mej) ** (-1 / beta)
            v_m[v_m > 3e10] = speed_of_light
            if use_gamma_ray_opacity:
                eff = 1 - np.exp(-(3 * kappa_gamma * mej / (4 * np.pi * vej ** 2)) * time[ii] ** -2)   # :674-676
            else:
                eff = thermalisation_efficiency
            qdot = eff * lsd[ii]
            kap = kappa[:-1, ii] if neutron_precursor_switch else kappa
            td_v[:-1, ii] = (kap * m_array[:-1] * solar_mass * 3) / (4 * np.pi * v_m[:-1] * speed_of_light * time[ii] * beta)
            lum_rad[:-1, ii] = energy_v[:-1, ii] / (td_v[:-1, ii] + time[ii] * (v_m[:-1] / speed_of_light))
            if magnetar_heating == "all_layers":
                energy_v[:-1, ii + 1] = (qdot + edotr[:-1, ii] * dm * solar_mass - (energy_v[:-1, ii] / time[ii])
                                         - lum_rad[:-1, ii]) * dt[ii] + energy_v[:-1, ii]       # :704
            else:
                energy_v[0, ii + 1] = (qdot + edotr[0, ii] * dm[0] * solar_mass - (energy_v[0, ii] / time[ii])
                                       - lum_rad[0, ii]) * dt[ii] + energy_v[0, ii]             # :707
                energy_v[1:-1, ii + 1] = (edotr[1:-1, ii] * dm[1:] * solar_mass - (energy_v[1:-1, ii] / time[ii])
                                          - lum_rad[1:-1, ii]) * dt[ii] + energy_v[1:-1, ii]    # :710
            tau[:-1, ii] = (m_array[:-1] * solar_mass * kap / (4 * np.pi * (time[ii] * v_m[:-1]) ** 2))
            tau[mass_len - 1, ii] = tau[mass_len - 2, ii]
            r_photosphere[ii] = v_m[np.argmin(np.abs(tau[:, ii] - 1))] * time[ii]
        lbol = np.sum(lum_rad, axis=0)
        if pair_cascade_switch:                                                                # :725-728
            tlife = ((0.6 / (1 - ejecta_albedo)) * (pair_cascade_fraction / 0.1) ** 0.5 * (lsd / 1.0e45) ** 0.5
                     * (v0 / (0.3 * speed_of_light)) ** (0.5) * (time / day_to_s) ** (-0.5))
            lbol = lbol / (1.0 + tlife)
        temperature = (lbol / (4.0 * np.pi * (r_photosphere) ** (2.0) * sigma_sb)) ** (0.25)
        return lbol, temperature, r_photosphere


def _metzger_magnetar_driven(time, redshift, lum_fn, mej, vej, beta, kappa_r, use_gamma_ray_opacity, *, frequency,
                             dl=None, resolution=300, **dyn):
    """metzger_magnetar_driven_kilonova_model & co. (:760-905) + _process_flux_density (:240-272) with
    use_relativistic_blackbody=False."""
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    t_grid = get_optimal_time_array(1e-4, 1e7, resolution)
    _, temperature, r_photosphere = metzger_magnetar_driven_internal(t_grid, mej, vej, beta, kappa_r, lum_fn(t_grid),
                                                                     use_gamma_ray_opacity, **dyn)
    frequency, t = calc_kcorrected_properties(np.asarray(frequency, dtype=float), redshift,
                                              np.asarray(time, dtype=float) * day_to_s)
    with np.errstate(all="ignore"):
        fd = blackbody_to_flux_density(interp1d(t_grid, y=temperature)(t), interp1d(t_grid, y=r_photosphere)(t), dl,
                                       frequency)
        return fd * _ERG_S_CM2_HZ_TO_MJY * (1 + redshift)


def general_metzger_magnetar_driven_evolution(time, redshift, mej, vej, beta, kappa_r, logbint, logbext, p0, chi0,
                                              radius, logmoi, kappa_gamma, *, frequency, **kw):
    """general_metzger_magnetar_driven_evolution (:908-956): 500-point grid request (:917)."""
    return _metzger_magnetar_driven(time, redshift, _evolving(logbint, logbext, p0, chi0, radius, logmoi), mej, vej,
                                    beta, kappa_r, True, frequency=frequency, kappa_gamma=kappa_gamma, resolution=500,
                                    **kw)


def metzger_magnetar_driven_kilonova_model(time, redshift, mej, vej, beta, kappa_r, p0, bp, mass_ns, theta_pb,
                                           thermalisation_efficiency, *, frequency, **kw):
    return _metzger_magnetar_driven(time, redshift, lambda t: basic_magnetar(t, p0, bp, mass_ns, theta_pb), mej, vej,
                                    beta, kappa_r, False, frequency=frequency,
                                    thermalisation_efficiency=thermalisation_efficiency, **kw)


def general_metzger_magnetar_driven(time, redshift, mej, vej, beta, kappa_r, l0, tau_sd, nn,
                                    thermalisation_efficiency, *, frequency, **kw):
    return _metzger_magnetar_driven(time, redshift, lambda t: magnetar_only(t, l0, tau_sd, nn), mej, vej, beta,
                                    kappa_r, False, frequency=frequency,
                                    thermalisation_efficiency=thermalisation_efficiency, **kw)


def general_metzger_magnetar_driven_thermalisation(time, redshift, mej, vej, beta, kappa_r, l0, tau_sd, nn,
                                                   kappa_gamma, *, frequency, **kw):
    return _metzger_magnetar_driven(time, redshift, lambda t: magnetar_only(t, l0, tau_sd, nn), mej, vej, beta,
                                    kappa_r, True, frequency=frequency, kappa_gamma=kappa_gamma, **kw)


def metzger_kilonova_model(time, redshift, mej, vej, beta, kappa, *, frequency, dense_resolution=500, dl=None,
                           neutron_precursor_switch=True, vmax=0.7, **_):
    """kilonova_models.py:1895-1935 (flux_density branch).  Returns mJy[E]."""
    time = np.asarray(time, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    t_src_s = time * day_to_s / (1. + redshift)
    time_temp = get_optimal_time_array(1e-2, 7e6, dense_resolution, user_times=t_src_s)
    _, temperature, r_photosphere = metzger_kilonova_internal(time_temp, mej, vej, beta, kappa,
                                                              neutron_precursor_switch, vmax)
    with np.errstate(all="ignore"):
        t = time * day_to_s
        temp_func = interp1d(time_temp, y=temperature)
        rad_func = interp1d(time_temp, y=r_photosphere)
        frequency, t = calc_kcorrected_properties(np.asarray(frequency, dtype=float), redshift, t)
        fd = blackbody_to_flux_density(temp_func(t), rad_func(t), dl, frequency)
        return fd * _ERG_S_CM2_HZ_TO_MJY * (1 + redshift)


# --------------------------------------------------------------------------------------------------
# redback-native structured-jet afterglow (afterglow_models/base_models.py:32-483, 486-640, 885-1010)
# Restated with numpy vectorised over steps / cells; arithmetic order follows the reference.
# --------------------------------------------------------------------------------------------------
AG_MP = 1.6726231e-24
AG_ME = 9.1093897e-28
AG_CC = 2.99792453e10
AG_QE = 4.8032068e-10
AG_C2 = AG_CC * AG_CC
AG_SIGT = (AG_QE * AG_QE / (AG_ME * AG_C2)) ** 2 * (8 * np.pi / 3)
AG_FOURPI = 4 * np.pi
_AG_PXF = np.array([[1.0, 3.0, 0.41], [1.2, 1.4, 0.44], [1.4, 1.1, 0.48], [1.6, 0.86, 0.53], [1.8, 0.725, 0.56],
                    [2.0, 0.637, 0.59], [2.2, 0.579, 0.612], [2.5, 0.520, 0.630], [2.7, 0.487, 0.641],
                    [3.0, 0.451, 0.659], [3.2, 0.434, 0.660], [3.4, 0.420, 0.675]])


def ag_segments(thj, res):
    """get_segments_numba, base_models.py:43-64"""
    latstep = thj / res
    rotstep = 2.0 * np.pi / res
    nlat = int(res)
    nrot = int(2 * np.pi / rotstep)
    phi = np.linspace(rotstep, nrot * rotstep, nrot)
    i = np.arange(nlat)
    omi = ((phi - (phi - rotstep))[None, :] * (np.cos(i * latstep) - np.cos((i + 1) * latstep))[:, None]).ravel()
    thi = np.linspace(latstep - latstep / 2., nlat * latstep - latstep / 2., nlat)
    phii = np.linspace(rotstep - rotstep / 2., nrot * rotstep - rotstep / 2., nrot)
    return omi, thi, phii, rotstep, latstep


def ag_erf(x):
    """erf_approximation_numba (A&S 7.1.26), base_models.py:67-89"""
    a1, a2, a3, a4, a5, p = 0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429, 0.3275911
    sign = np.where(x >= 0, 1.0, -1.0)
    xa = np.abs(x)
    t = 1.0 / (1.0 + p * xa)
    y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * np.exp(-xa * xa)
    return sign * y


def ag_structure(gamma, en, thi, thc, method_id, thj, s=0.01, a=0.5):
    """get_structure_numba, base_models.py:92-148 (1 tophat, 2 Gaussian, 3 power law, 4 alternative power law,
    5 two-component, 6 double Gaussian)"""
    gs = np.full(thi.shape[0], float(gamma))
    ei = np.full(thi.shape[0], float(en))
    with np.errstate(all="ignore"):
        if method_id == 1:
            fac = ag_erf(-(thi - min(thc, thj)) * 1000.0) * 0.5 + 0.5
            return (gs - 1) * fac + 1.0000000000001, ei * fac
        if method_id == 2:
            fac = np.exp(-0.5 * (thi / thc) ** 2)
            return (gs - 1.0) * fac + 1.0000000000001, ei * fac
        if method_id == 3:
            m = thi >= thc
            ei[m] = ei[m] * (thc / thi[m]) ** s
            gs[m] = (gs[m] - 1.0) * (thc / thi[m]) ** a + 1.000000000000001
            return gs, ei
        if method_id == 4:
            fac = (1 + (thi / thc) ** 2) ** 0.5
            return 1.000000000001 + (gs - 1.0) * fac ** (-a), ei * fac ** (-s)
        if method_id == 5:
            m = thi > thc
            gs[m] = a
            ei[m] = ei[m] * s
            return gs, ei
        if method_id == 6:
            ec, ej = np.exp(-0.5 * (thi / thc) ** 2), np.exp(-0.5 * (thi / thj) ** 2)
            eo = ei * ((1.0 - s) * ec + s * ej)
            temp = ((1.0 - s) * ec + s * ej) / ((1.0 - s / a) * ec + (s / a) * ej)
            return np.where(np.isfinite(temp), (gs - 1.0) * temp + 1.0000000000001, a), eo
    raise NotImplementedError(method_id)


def _ag_rk4(ghat, dm, g, m, fac, therm):
    """RK4_step_numba, base_models.py:182-192"""
    ghatm1 = ghat - 1.0
    dm10 = 10.0 ** dm
    g2 = g * g
    return fac * dm10 * (ghat * (g2 - 1.0) - ghatm1 * (g - 1.0 / g)) / \
        (m + dm10 * (therm + (1.0 - therm) * (2.0 * ghat * g - ghatm1 * (1 + 1.0 / g2))))


def ag_gamma(g0, eps, therm, steps, n0, k):
    """get_gamma_numba, base_models.py:195-249.  Returns G, dm (= 10**state_dm), ghat with shape [ring, step]."""
    steps = int(steps)
    nmin = (AG_FOURPI / (3.0 - k)) * n0 * 1e10 ** (3.0 - k)
    nmax = (AG_FOURPI / (3.0 - k)) * n0 * 1e24 ** (3.0 - k)
    dlogm = np.log10(nmin * AG_MP)
    h = np.log10(nmax / nmin) / float(steps)
    fac = -h * np.log(10.0)
    m = eps / (g0 * AG_C2)
    sg = np.zeros((steps, g0.shape[0]))
    sdm = np.zeros((steps, g0.shape[0]))
    sgh = np.zeros((steps, g0.shape[0]))
    g = g0.astype(np.float64).copy()
    dm = np.full(g0.shape[0], dlogm)
    for i in range(steps):
        sg[i] = g
        sdm[i] = dm
        g2m1 = g * g - 1.0
        root = np.sqrt(g2m1)
        theta = root * (root + 1.07 * g2m1) / (3.0 * (1 + root + 1.07 * g2m1))
        z = theta / (0.24 + theta)
        ghat = ((((((1.07136 * z - 2.39332) * z + 2.32513) * z - 0.96583) * z + 0.18203) * z - 1.21937) * z + 5.0) / 3.0
        sgh[i] = ghat
        f1 = _ag_rk4(ghat, dm, g, m, fac, therm)
        f2 = _ag_rk4(ghat, dm + 0.5 * h, g + 0.5 * f1, m, fac, therm)
        f3 = _ag_rk4(ghat, dm + 0.5 * h, g + 0.5 * f2, m, fac, therm)
        f4 = _ag_rk4(ghat, dm + h, g + f3, m, fac, therm)
        g = g + (f1 + 2.0 * (f2 + f3) + f4) / 6.0 + 1e-15
        dm = dm + h
    return sg.T, (10.0 ** sdm).T, sgh.T


def ag_step1(g, dm, p, xp, fx, eb, ee, n, k, thi, ghat, rotstep, latstep, xin, is_expansion, a1, res):
    """calc_afterglow_step1_numba for one ring, base_models.py:252-322"""
    gm1 = g - 1.0
    g2 = g * g
    beta = np.sqrt(1 - 1.0 / g2)
    ne = dm / AG_MP
    cs = np.sqrt(AG_C2 * ghat * (ghat - 1) * gm1 / (1 + ghat * gm1))
    te = np.arcsin(cs / (AG_CC * np.sqrt(g2 - 1.0)))
    fac = 0.5 * latstep
    if is_expansion:
        ex = te / g ** (a1 + 1)
        omg = rotstep * (np.cos(thi - fac) - np.cos(ex / res + thi + fac))
    else:
        ex = np.ones(g.shape[0])
        omg = np.full(g.shape[0], rotstep * (np.cos(thi - fac) - np.cos(thi + fac)))
    exponent = np.ones(g.shape[0])
    exponent[1:] = ((1 - np.cos(latstep + ex[0])) / (1 - np.cos(latstep + ex[1:]))) ** (1.0 / 2.0)
    r = ((3.0 - k) * ne / (AG_FOURPI * n)) ** (1.0 / (3.0 - k))
    r[1:] = np.diff(r) * exponent[1:] ** (1.0 / (3.0 - k))
    r = np.cumsum(r)
    n0 = n * r ** (-k)
    b = np.sqrt(2 * AG_FOURPI * eb * n0 * AG_MP * AG_C2 * ((ghat * g + 1.0) / (ghat - 1.0)) * gm1)
    gmm = np.sqrt(1.5 * AG_FOURPI * AG_QE / (AG_SIGT * b))
    if p > 2:
        gm = ((p - 2) / (p - 1)) * (ee / xin) * gm1 * (AG_MP / AG_ME)
    elif p == 2:
        gm = (1 / np.log(gmm / ((ee / xin) * gm1 * (AG_MP / AG_ME)))) * (ee / xin) * gm1 * (AG_MP / AG_ME)
    else:
        gm = ((2 - p) / (p - 1) * (AG_MP / AG_ME) * (ee / xin) * gm1 * gmm ** (p - 2)) ** (1 / (p - 1))
    nump = 3.0 * xp * gm * gm * AG_QE * b / (AG_FOURPI * AG_ME * AG_CC)
    pp = xin * fx * AG_ME * AG_C2 * AG_SIGT * b / (3.0 * AG_QE)
    kt = gm * AG_ME * AG_C2
    return beta, ne, omg, r, b, gm, nump, pp, kt


def ag_step2(dl, om0, rotstep, obsa, beta, ne, omg, r, b, nump, pp, kt, g, is_expansion):
    """calc_afterglow_step2_numba for one cell, base_models.py:325-370"""
    dl2 = dl * dl
    no = om0 * ne / AG_FOURPI
    cosa = np.cos(obsa)
    om = np.maximum(om0, omg) if is_expansion else omg
    thii = np.arccos(1.0 - om / rotstep)
    rdiff = np.diff(r, prepend=0)
    tobs = np.cumsum(rdiff * (1.0 / beta - cosa) / AG_CC)
    tobso = np.cumsum(rdiff * (1.0 / beta - 1.0) / AG_CC)
    dop = 1.0 / (g * (1.0 - beta * cosa))
    gc = 6.0 * np.pi * AG_ME * AG_CC / (g * AG_SIGT * b * b * tobso)
    nucp = 0.286 * 3.0 * gc * gc * AG_QE * b / (AG_FOURPI * AG_ME * AG_CC)
    num = dop * nump
    nuc = dop * nucp
    fmax = no * pp * dop * dop * dop / (AG_FOURPI * dl2)
    fbb = 2 * om * np.cos(thii) * dop * kt * r * r / (AG_C2 * dl2)
    return fbb, fmax, nuc, num, tobs


def ag_flux(fbb, nuc, num, nu1, fmax, p):
    """get_ag_numba (overlapping if-chain: later conditions overwrite), base_models.py:373-399"""
    with np.errstate(all="ignore"):
        f = np.zeros(num.shape[0])
        c = (nuc < num) & (nu1 < nuc); f[c] = (fmax * (nu1 / nuc) ** (1.0 / 3.0))[c]
        c = (nuc < nu1) & (nuc < num); f[c] = (fmax * (nu1 / nuc) ** (-1.0 / 2.0))[c]
        c = (num < nu1) & (nuc < num); f[c] = (fmax * (num / nuc) ** (-1.0 / 2.0) * (nu1 / num) ** (-p / 2.0))[c]
        c = (num < nuc) & (nu1 < num); f[c] = (fmax * (nu1 / num) ** (1.0 / 3.0))[c]
        c = (num < nu1) & (num < nuc); f[c] = (fmax * (nu1 / num) ** (-(p - 1.0) / 2.0))[c]
        c = (nuc < nu1) & (num < nuc); f[c] = (fmax * (nuc / num) ** (-(p - 1.0) / 2.0) * (nu1 / nuc) ** (-p / 2.0))[c]
        fbb_adj = fbb * nu1 ** 2.0 * np.maximum(1.0, (nu1 / num) ** 0.5)
        return np.minimum(fbb_adj, f)


def calc_ABmag_from_flux_density(fluxdensity_mjy):
    """utils.calc_ABmag_from_flux_density (utils.py:481-488): `(f * mJy).to(ABmag)`.  astropy (absent here) defines
    AB = 10**(48.6 / -2.5) erg s^-1 cm^-2 Hz^-1 (astropy/units/photometric.py) and ABmag = -2.5 log10(f / AB);
    restated from that published definition."""
    with np.errstate(all="ignore"):
        return -2.5 * np.log10(np.asarray(fluxdensity_mjy, dtype=float) * (1e-26 / 10 ** (48.6 / -2.5)))


def ag_default_res(thv, thc, g0):
    """tophat_redback / gaussian_redback resolution rule, base_models.py:932-935"""
    sep = max(thv - thc, 0)
    order = min(int((2 - 10 * sep) * thc * g0), 100)
    return max(10, order)


def redback_afterglow(time, redshift, thv, loge0, thc, logn0, p, logepse, logepsb, g0, xiN, *, frequency,
                      method_id=1, thj=None, res=None, steps=250, k=0, a1=1, expansion=1, dl=None, ss=0.01, aa=0.5,
                      **_):
    """tophat_redback (method_id=1, base_models.py:885-946) / gaussian_redback (method_id=2, :948-1010) through
    RedbackAfterglows.get_lightcurve (:573-640), output_format='flux_density'.  time: observer days.  mJy[E]."""
    time = np.asarray(time, dtype=float) * 86400.0
    frequency = np.asarray(frequency, dtype=float)
    if frequency.ndim == 0:
        frequency = np.ones(len(time)) * float(frequency)
    if k not in (0, 2):
        raise ValueError("k must either be 0 or 2")
    if (p < 1.2) or (p > 3.4):
        raise ValueError("p is out of range, 1.2 < p < 3.4")
    epse, epsb, nism, e0 = 10 ** logepse, 10 ** logepsb, 10 ** logn0, 10 ** loge0
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    if thj is None:
        thj = thc
    if res is None:
        res = ag_default_res(thv, thc, g0)
    n = float(nism) if k == 0 else float(nism * 3e35)
    res = float(res)
    with np.errstate(all="ignore"):
        xp = np.interp(p, _AG_PXF[:, 0], _AG_PXF[:, 1])
        fx = np.interp(p, _AG_PXF[:, 0], _AG_PXF[:, 2])
        nu0 = np.unique(frequency)
        nu = nu0 * (1 + redshift)
        omi, thi, phii, rotstep, latstep = ag_segments(thj, res)
        gs, ei = ag_structure(g0, e0, thi, thc, method_id, thj, ss, aa)
        G, SM, ghat = ag_gamma(gs, ei, 0.0, steps, n, float(k))
        sin_thi, cos_thi = np.sin(thi), np.cos(thi)
        ang = (np.cos(thv) * cos_thi[:, None] + np.cos(0.0) * np.sin(thv) * np.cos(phii)[None, :] * sin_thi[:, None]
               + np.sin(0.0) * np.sin(thv) * np.sin(phii)[None, :] * sin_thi[:, None])
        obsa = np.arccos(np.clip(ang, -1.0, 1.0)).ravel()                      # get_obsangle_numba :151-179
        lc = np.zeros(frequency.size)
        kk = 0
        for i in range(thi.size):
            beta, ne, omg, r, b, gm, nump, pp, kt = ag_step1(G[i], SM[i], p, xp, fx, epsb, epse, n, float(k), thi[i],
                                                             ghat[i], rotstep, latstep, xiN, bool(expansion),
                                                             float(a1), res)
            for _j in range(phii.size):
                fbb, fmax, nuc, num, tobs = ag_step2(dl, omi[kk], rotstep, obsa[kk], beta, ne, omg, r, b, nump, pp,
                                                     kt, G[i], bool(expansion))
                for hh in range(nu.size):
                    sel = frequency == nu0[hh]
                    lc[sel] += np.interp(time[sel] / (1 + redshift), tobs, ag_flux(fbb, nuc, num, nu[hh], fmax, p))
                kk += 1
        return lc * (1 + redshift) / 1e-26


# --------------------------------------------------------------------------------------------------
# magnitude / bandpass branch (sed.py:10-118, 120-195, 911-1009; supernova_models.py:1008-1028 etc.)
#
# PARITY UNPINNED: sncosmo (TimeSeriesSource, Bandpass, integration_grid, ABMagSystem) is a third-party
# dependency that is neither vendored in the reference nor installable here, the reference's tests mock it
# (test/sed_test.py:436-490) and the golden pickle covers flux_density only.  What follows restates the
# published sncosmo 2.10 algorithm on top of the same scipy RectBivariateSpline sncosmo uses.
# --------------------------------------------------------------------------------------------------
HC_ERG_AA = 1.9864458571489284e-08       # sed.py:19
MODEL_BANDFLUX_SPACING = 5.0             # sed.py:20
_C_AA_PER_S = 2.99792458e18


class Bandpass:
    """sncosmo.Bandpass restricted to what the path uses: tabulated transmission, linear interpolation,
    zero outside [minwave, maxwave]."""

    def __init__(self, name, wave, trans):
        self.name = name
        self.wave = np.asarray(wave, dtype=float)
        self.trans = np.asarray(trans, dtype=float)

    def minwave(self):
        return self.wave[0]

    def maxwave(self):
        return self.wave[-1]

    def __call__(self, wave):
        return np.interp(wave, self.wave, self.trans, left=0.0, right=0.0)

    def zpbandflux_ab(self):
        """sncosmo ABMagSystem._refspectrum_bandflux: sum(3631e-23 / h / nu * trans * gradient(nu))."""
        nu = _C_AA_PER_S / self.wave
        tr = self.trans
        if nu[0] > nu[-1]:
            nu, tr = nu[::-1], tr[::-1]
        f = 3631.e-23 / planck / nu
        return np.sum(f * tr * np.gradient(nu))


def integration_grid(low, high, target_spacing):
    """sncosmo.utils.integration_grid"""
    range_diff = high - low
    spacing = range_diff / int(math.ceil(range_diff / target_spacing))
    grid = np.arange(low + spacing / 2., high, spacing)
    return grid, spacing


def flux_density_to_spectrum(flux_density_cgs, redshift, lambda_observer_frame):
    """sed.py:911-934: erg/s/Hz/cm^2 -> (x (1+z)) mJy -> erg/cm^2/s/Angstrom (astropy spectral_density)."""
    mjy = flux_density_cgs * (1 + redshift) * _ERG_S_CM2_HZ_TO_MJY
    return mjy * 1e-26 * _C_AA_PER_S / lambda_observer_frame ** 2


def clean_spectra(values):
    """sed.py:985-992"""
    with np.errstate(all="ignore"):
        valid = np.isfinite(values) & (values != 0)
        mean_value = np.nanmean(np.where(valid, values, np.nan)) if valid.any() else np.nan
    if not np.isfinite(mean_value) or mean_value == 0:
        mean_value = 1.0
    return np.where(valid, values, 1e-30 * mean_value)


def bandmag_from_spectra(time, time_eval, spectra, lambda_array, bands, bandpasses, time_spline_degree=3):
    """get_correct_output_format_from_spectra (sed.py:971-1009, output_format='magnitude') ->
    RedbackTimeSeriesSource.bandmag -> _bandmag_redback -> _bandflux_redback -> _bandflux_single_redback.
    `bands`: array of band names per epoch; `bandpasses`: {name: Bandpass}."""
    from scipy.interpolate import RectBivariateSpline
    spectra = clean_spectra(np.asarray(spectra, dtype=float))
    if len(time_eval) < 4:
        time_spline_degree = max(1, min(time_spline_degree, len(time_eval) - 1))
    spline = RectBivariateSpline(time_eval, lambda_array, spectra, kx=time_spline_degree, ky=3)
    time = np.atleast_1d(np.asarray(time, dtype=float))
    bands = np.broadcast_to(np.asarray(bands), time.shape)
    bandflux = np.zeros(time.shape)
    for name in set(bands.tolist()):
        mask = bands == name
        b = bandpasses[name]
        if b.minwave() < lambda_array[0] or b.maxwave() > lambda_array[-1]:
            raise ValueError(f"bandpass {name!r} outside spectral range")              # sed.py:22-28
        wave, dwave = integration_grid(b.minwave(), b.maxwave(), MODEL_BANDFLUX_SPACING)
        trans = b(wave)
        f = np.abs(spline(time[mask], wave))                                           # sed.py:35-36
        bandflux[mask] = np.sum(wave * trans * f, axis=1) * dwave / HC_ERG_AA          # sed.py:37
    with np.errstate(all="ignore"):
        zp = np.array([bandpasses[n].zpbandflux_ab() for n in bands.tolist()])
        return -2.5 * np.log10(bandflux / zp)                                          # sed.py:111-114


def bandpass_magnitude_to_flux(magnitude, reference_flux):
    """utils.py:707-718 with the per-band reference flux looked up by the caller (tables/filters.csv)."""
    return 10.0 ** (magnitude / (-2.5)) * reference_flux


def sn_magnitude(model, time, redshift, *, bands=None, bandpasses=None, lambda_array=None, dl=None,
                 spectra_transform=None, spectra_only=False, **params):
    """Magnitude branch of arnett / basic_magnetar_powered / slsn (supernova_models.py:1008-1028, 1510-1530,
    1599-1624; csm_interaction :2093-2111).  `model` in {'arnett', 'basic_magnetar_powered', 'slsn', 'type_1a',
    'csm_interaction'}."""
    time_obs = np.asarray(time, dtype=float)
    lam = np.geomspace(100, 60000, 100) if lambda_array is None else np.asarray(lambda_array, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    time_temp = get_optimal_time_array(0.1, 500 if model == "csm_interaction" else 3000, 300)   # :1010, :2095
    time_observer_frame = time_temp * (1. + redshift)
    frequency, t = calc_kcorrected_properties(lambda_to_nu(lam), redshift, time_observer_frame)
    dkw = dict(vej=params["vej"], kappa=params["kappa"], kappa_gamma=params.get("kappa_gamma"))
    if model == "csm_interaction":                                                             # :2099-2100
        lbol = csm_interaction_bolometric(t, params["mej"], params["csm_mass"], params["vej"], params["eta"],
                                          params["rho"], params["kappa"], params["r0"])
    elif model in ("arnett", "type_1a"):
        lbol = arnett_bolometric(t, params["f_nickel"], params["mej"], **dkw)
    else:
        lbol = basic_magnetar_powered_bolometric(t, params["p0"], params["bp"], params["mass_ns"],
                                                 params["theta_pb"], mej=params["mej"], **dkw)
    temp, rad = temperature_floor(t, lbol, params["vej"], params["temperature_floor"])
    with np.errstate(all="ignore"):
        if model in ("slsn", "type_1a"):
            full = cutoff_blackbody_mjy(t, temp, lbol, rad, frequency[:, None], dl,
                                        params.get("cutoff_wavelength", 3000.), params.get("absorption_index", 1.0))
            if model == "type_1a":                                               # :2295-2304
                lk = {k: params[k] for k in ("line_wavelength", "line_width", "line_time", "line_duration",
                                             "line_amplitude") if k in params}
                full = line_sed_cgs(t, lbol, frequency[:, None], full * 1e-26, dl, **lk) * _ERG_S_CM2_HZ_TO_MJY
            full = full.T * (1 + redshift)
            spectra = full * 1e-26 * _C_AA_PER_S / lam ** 2                      # :1613-1616, :2305-2309
        else:
            fd = blackbody_to_flux_density(temp, rad, dl, frequency[:, None]).T
            spectra = flux_density_to_spectrum(fd, redshift, lam)
        if spectra_transform is not None:                                        # extinction_models.py:236-251
            spectra = spectra_transform(spectra, lam)
    if spectra_only:                                                             # output_format='spectra' (:1019-1023)
        return time_observer_frame, lam, spectra
    return bandmag_from_spectra(time_obs, time_observer_frame, spectra, lam, bands, bandpasses)


def general_magnetar_driven_supernova_magnitude(time, redshift, mej, E_sn, kappa, l0, tau_sd, nn, kappa_gamma, *,
                                                bands, bandpasses, temperature_floor, f_nickel=0,
                                                pair_cascade_switch=False, cutoff_wavelength=3000,
                                                absorption_index=1.0, lambda_array=None, dl=None, **_):
    """Magnitude branch of supernova_models.general_magnetar_driven_supernova (:2585-2637): CutoffBlackbody spectra on
    (get_optimal_time_array(1e0, 1e8, 2000) x lambda_array), phases time_temp (1 + z) / day_to_s."""
    time_obs = np.asarray(time, dtype=float)
    lam = np.geomspace(500, 60000, 200) if lambda_array is None else np.asarray(lambda_array, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    t_grid, lbol_g, vej = _magnetar_driven_supernova_dynamics(mej, E_sn, kappa, l0, tau_sd, nn, kappa_gamma, f_nickel,
                                                              pair_cascade_switch)
    time_observer_frame = t_grid * (1. + redshift)
    frequency, _t = calc_kcorrected_properties(lambda_to_nu(lam), redshift, time_observer_frame)
    days = t_grid / day_to_s
    with np.errstate(all="ignore"):
        temp_g, rad_g = globals()["temperature_floor"](days, lbol_g, vej, temperature_floor)
        full = cutoff_blackbody_mjy(days, temp_g, lbol_g, rad_g, frequency[:, None], dl, cutoff_wavelength,
                                    absorption_index)
        full = full.T * (1 + redshift)
        spectra = full * 1e-26 * _C_AA_PER_S / lam ** 2                          # :2628-2629
    return bandmag_from_spectra(time_obs, time_observer_frame / day_to_s, spectra, lam, bands, bandpasses)


def mergernova_magnitude(model, time, redshift, *, bands, bandpasses, lambda_array=None, dl=None, **params):
    """_processing_other_formats (magnetar_driven_ejecta_models.py:198-238) with use_relativistic_blackbody=True for
    `model` in {'basic_mergernova', 'general_mergernova', 'general_mergernova_thermalisation'}."""
    time_obs = np.asarray(time, dtype=float)
    lam = np.geomspace(100, 60000, 100) if lambda_array is None else np.asarray(lambda_array, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    t_grid = get_optimal_time_array(1e-4, 1e8, 500)
    ej = [params[k] for k in ("mej", "beta", "ejecta_radius", "kappa", "n_ism")]
    if model == "basic_mergernova":
        lum = basic_magnetar(t_grid, params["p0"], params["bp"], params["mass_ns"], params["theta_pb"])
        pcs = params.get("pair_cascade_switch", False)
    else:
        lum = magnetar_only(t_grid, params["l0"], params["tau_sd"], params["nn"])
        pcs = params.get("pair_cascade_switch", True)
    gamma_ray = model == "general_mergernova_thermalisation"
    _, temp, rad, dop = ejecta_dynamics_and_interaction(
        t_grid, *ej, lum, pcs, gamma_ray, thermalisation_efficiency=params.get("thermalisation_efficiency"),
        kappa_gamma=params.get("kappa_gamma"))
    time_observer_frame = t_grid * (1. + redshift)
    frequency, _ = calc_kcorrected_properties(lambda_to_nu(lam), redshift, time_observer_frame)
    with np.errstate(all="ignore"):
        nu = frequency[:, None]
        num = 2 * np.pi * planck * nu ** 3 * rad ** 2
        den = dl ** 2 * speed_of_light ** 2 * dop ** 2
        fd = (num / den * (1. / (np.exp((planck * nu) / (boltzmann_constant * temp * dop)) - 1))).T   # :164-173
        spectra = flux_density_to_spectrum(fd, redshift, lam)
    return bandmag_from_spectra(time_obs, time_observer_frame / day_to_s, spectra, lam, bands, bandpasses)


def metzger_magnetar_magnitude(model, time, redshift, *, bands, bandpasses, lambda_array=None, dl=None, **params):
    """_processing_other_formats (magnetar_driven_ejecta_models.py:198-238) with use_relativistic_blackbody=False for
    `model` in {'metzger_magnetar_driven_kilonova_model', 'general_metzger_magnetar_driven',
    'general_metzger_magnetar_driven_thermalisation'}."""
    time_obs = np.asarray(time, dtype=float)
    lam = np.geomspace(100, 60000, 100) if lambda_array is None else np.asarray(lambda_array, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    t_grid = get_optimal_time_array(1e-4, 1e7, 300)
    if model == "metzger_magnetar_driven_kilonova_model":
        lum = basic_magnetar(t_grid, params["p0"], params["bp"], params["mass_ns"], params["theta_pb"])
    else:
        lum = magnetar_only(t_grid, params["l0"], params["tau_sd"], params["nn"])
    gamma_ray = model == "general_metzger_magnetar_driven_thermalisation"
    _, temp, rad = metzger_magnetar_driven_internal(
        t_grid, params["mej"], params["vej"], params["beta"], params["kappa_r"], lum, gamma_ray,
        thermalisation_efficiency=params.get("thermalisation_efficiency"), kappa_gamma=params.get("kappa_gamma"))
    time_observer_frame = t_grid * (1. + redshift)
    frequency, _ = calc_kcorrected_properties(lambda_to_nu(lam), redshift, time_observer_frame)
    with np.errstate(all="ignore"):
        fd = blackbody_to_flux_density(temp, rad, dl, frequency[:, None]).T                      # sed.py:954-963
        spectra = flux_density_to_spectrum(fd, redshift, lam)
    return bandmag_from_spectra(time_obs, time_observer_frame / day_to_s, spectra, lam, bands, bandpasses)


def kn_magnitude(model, time, redshift, *, bands=None, bandpasses=None, lambda_array=None, dl=None,
                 dense_resolution=500, spectra_transform=None, spectra_only=False, **params):
    """Magnitude branch of one_component_kilonova_model / metzger_kilonova_model
    (kilonova_models.py:1579-1603, 1937-1962)."""
    time_obs = np.asarray(time, dtype=float)
    lam = np.geomspace(100, 60000, 200) if lambda_array is None else np.asarray(lambda_array, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    t_src_s = time_obs * day_to_s / (1. + redshift)
    time_temp = get_optimal_time_array(1e-2, 7e6, dense_resolution, user_times=t_src_s)
    if model == "one_component":
        _, temp, rad = one_component_kilonova_internal(time_temp, params["mej"], params["vej"], params["kappa"],
                                                       params.get("temperature_floor", 4000))
    else:
        _, temp, rad = metzger_kilonova_internal(time_temp, params["mej"], params["vej"], params["beta"],
                                                 params["kappa"], params.get("neutron_precursor_switch", True),
                                                 params.get("vmax", 0.7))
    time_observer_frame = time_temp * (1. + redshift)
    frequency, _t = calc_kcorrected_properties(lambda_to_nu(lam), redshift, time_observer_frame)
    with np.errstate(all="ignore"):
        fd = blackbody_to_flux_density(temp, rad, dl, frequency[:, None]).T
        spectra = flux_density_to_spectrum(fd, redshift, lam)
        if spectra_transform is not None:                                        # extinction_models.py:236-251
            spectra = spectra_transform(spectra, lam)
    if spectra_only:                                          # output_format='spectra' (kilonova_models.py:1596-1599)
        return time_observer_frame, lam, spectra
    return bandmag_from_spectra(time_obs, time_observer_frame / day_to_s, spectra, lam, bands, bandpasses)


def multi_component_kn_magnitude(time, redshift, components, *, bands, bandpasses, dense_resolution, grid_t_max,
                                 lambda_array=None, dl=None):
    """Magnitude branch of two_component_kilonova_model (kilonova_models.py:1277-1323: grid end 86400*6 s, 500
    points) and three_component_kilonova_model (:1166-1211: 7e6 s, 300 points): blackbody_to_spectrum of every
    component (sed.py:937-968) summed on the shared grid, then get_correct_output_format_from_spectra.
    components: [(mej, vej, kappa, temperature_floor), ...]."""
    time_obs = np.asarray(time, dtype=float)
    lam = np.geomspace(100, 60000, 200) if lambda_array is None else np.asarray(lambda_array, dtype=float)
    if dl is None:
        dl = cosmo.luminosity_distance_cm(redshift)
    time_temp = get_optimal_time_array(1e-2, grid_t_max, dense_resolution,
                                       user_times=time_obs * day_to_s / (1. + redshift))
    time_observer_frame = time_temp * (1. + redshift)
    frequency, _t = calc_kcorrected_properties(lambda_to_nu(lam), redshift, time_observer_frame)
    full_spec = np.zeros((len(time_temp), len(lam)))
    with np.errstate(all="ignore"):
        for mej, vej, kappa, tfloor in components:
            _, temp, rad = one_component_kilonova_internal(time_temp, mej, vej, kappa, tfloor)
            fd = blackbody_to_flux_density(temp, rad, dl, frequency[:, None]).T
            full_spec += flux_density_to_spectrum(fd, redshift, lam)
    return bandmag_from_spectra(time_obs, time_observer_frame / day_to_s, full_spec, lam, bands, bandpasses)


# --------------------------------------------------------------------------------------------------
# simulate_transients.SimulateTransientWithCadence noise model (simulate_transients.py:457-527)
# --------------------------------------------------------------------------------------------------
def optical_noise_and_cuts(model_mag, ref_flux, limiting_mag, z, noise_type="limiting_mag", snr_threshold=5.0):
    """z: standard-normal draws with the shape of model_mag (the reference calls rng.normal(0, sigma, n_obs),
    which equals sigma * standard_normal for numpy's Generator)."""
    model_mag = np.asarray(model_mag, dtype=float)
    out = {}
    with np.errstate(all="ignore"):
        if noise_type == "limiting_mag":
            model_flux = bandpass_magnitude_to_flux(model_mag, ref_flux)
            flux_error = bandpass_magnitude_to_flux(np.asarray(limiting_mag, dtype=float), ref_flux) / 5.0
            observed_flux = model_flux + z * flux_error
            snr = model_flux / flux_error
            observed_mag = -2.5 * np.log10(observed_flux / ref_flux)                    # utils.py:834-845
            mag_error = np.abs((2.5 / np.log(10)) * (flux_error / observed_flux))      # utils.py:829
            mag_error[np.isnan(observed_flux) | (np.abs(observed_flux) <= 1.0e-20)] = np.nan
            out.update(flux=observed_flux, flux_error=np.broadcast_to(flux_error, model_mag.shape))
        elif noise_type == "gaussian":
            mag_error = np.ones(model_mag.shape) * 0.1
            observed_mag = model_mag + z * 0.1
            snr = 1.0 / mag_error
        elif noise_type == "gaussianmodel":
            mag_error = 0.05 * np.abs(model_mag)
            observed_mag = model_mag + z * mag_error
            snr = np.full(model_mag.shape, 1.0 / 0.05)
        else:
            raise ValueError(noise_type)
    out.update(magnitude=observed_mag, magnitude_error=mag_error, snr=snr, detected=snr >= snr_threshold)
    return out


# ---- counter-based normal variates of the device noise kernel -------------------------------------------------------
# The reference draws its observation noise from numpy's generators (simulate_transients.py:476, 493, 506, 1111); a GPU
# batch cannot replay that stream, so csrc/noise_kernel.cu draws from Philox4x32-10 (Salmon et al. 2011, "Parallel random
# numbers: as easy as 1, 2, 3"; Random123).  This restatement of the published algorithm is pinned to Random123's
# known-answer vectors in tests/test_oracle_golden.py and lets the checker replay the device's variates exactly.
def philox4x32_10(counter, key):
    """counter: 4 arrays of uint32, key: 2 arrays of uint32 -> 4 arrays of uint32."""
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c = [np.asarray(x, dtype=np.uint32) for x in counter]
    k = [np.asarray(x, dtype=np.uint64) for x in key]
    s32 = np.uint64(32)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = m0 * c[0].astype(np.uint64)
        p1 = m1 * c[2].astype(np.uint64)
        k0, k1 = (k[0] & mask).astype(np.uint32), (k[1] & mask).astype(np.uint32)
        c = [(p1 >> s32).astype(np.uint32) ^ c[1] ^ k0, (p1 & mask).astype(np.uint32),
             (p0 >> s32).astype(np.uint32) ^ c[3] ^ k1, (p0 & mask).astype(np.uint32)]
        k = [k[0] + np.uint64(0x9E3779B9), k[1] + np.uint64(0xBB67AE85)]
    return c


def philox_standard_normal(seed, index):
    """Standard-normal variate of global element `index` (array of non-negative ints) for `seed`: two 53-bit uniforms
    in (0, 1) from one Philox block, Box-Muller cosine branch (csrc/noise_kernel.cu::philox_normal)."""
    index = np.asarray(index, dtype=np.uint64)
    seed = int(seed)
    zero = np.zeros(index.shape, dtype=np.uint32)
    r = philox4x32_10([(index & np.uint64(0xFFFFFFFF)).astype(np.uint32), (index >> np.uint64(32)).astype(np.uint32),
                       zero, zero],
                      [np.full(index.shape, seed & 0xFFFFFFFF, dtype=np.uint64),
                       np.full(index.shape, (seed >> 32) & 0xFFFFFFFF, dtype=np.uint64)])
    def uniform(a, b):
        bits = ((a >> np.uint32(5)).astype(np.uint64) << np.uint64(26)) | (b >> np.uint32(6)).astype(np.uint64)
        return (bits.astype(np.float64) + 0.5) / 9007199254740992.0
    u1, u2 = uniform(r[0], r[1]), uniform(r[2], r[3])
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def survey_observation_single(model_magnitude, phase, ref_flux, five_sigma_depth, z, snr_threshold=5.0,
                              add_source_noise=False, source_noise_factor=0.02):
    """SimulateOpticalTransient._make_observation_single (simulate_transients.py:1096-1140) for one transient, given the
    model magnitudes at its pointings (`sncosmo_model.bandmag`, :1105), phase = expMJD - t0, the per-pointing reference
    flux and 5-sigma depth, and standard-normal draws z (the reference calls np.random.normal(loc=flux, scale=err))."""
    model_magnitude = np.asarray(model_magnitude, dtype=float)
    with np.errstate(all="ignore"):
        flux = bandpass_magnitude_to_flux(model_magnitude, ref_flux)                         # :1106
        err = ref_flux * np.power(10.0, -0.4 * np.asarray(five_sigma_depth, dtype=float)) / 5.0   # utils.py:517-541
        if add_source_noise:
            err = np.sqrt(err ** 2 + source_noise_factor ** 2 * flux ** 2)                   # :1108-1109
        observed_flux = flux + z * err                                                       # :1111
        magnitudes = -2.5 * np.log10(observed_flux / ref_flux)                               # utils.py:834-845
        magnitude_errs = np.abs((2.5 / np.log(10)) * (err / flux))                           # utils.py:808-831
        magnitude_errs[np.isnan(flux) | (np.abs(flux) <= 1.0e-20)] = np.nan
        mask = (np.asarray(phase) <= 0.) | np.isnan(magnitudes)                              # :1128
        snr = observed_flux / err                                                            # :1129
        detected = np.ones(model_magnitude.shape)
        detected[mask] = 0
        detected[snr < snr_threshold] = 0                                                    # :1130-1133
    return dict(magnitude=magnitudes, e_magnitude=magnitude_errs, flux=observed_flux, flux_error=err,
                detected=detected.astype(bool), flux_limit=bandpass_magnitude_to_flux(
                    np.asarray(five_sigma_depth, dtype=float), ref_flux))


# --------------------------------------------------------------------------------------------------
# dust extinction (transient_models/extinction_models.py:77-170).  The laws live in the third-party `extinction`
# package (kbarbary/extinction; redback's setup.py / optional_requirements.txt pin no version; not vendored, not
# installable here).  Pinned only to the package's documented example values (README / API docs, as known-answer
# tests in tests/test_extinction.py: ccm89([2000, 4000, 8000] A, 1.0, 3.1) = [2.84252644, 1.4645557, 0.59748901],
# fitzpatrick99(same) = [2.76225609, 1.42325373, 0.55333671]); otherwise PARITY UNPINNED (odonnell94, calzetti00 have
# no vector).  Restated from the papers the package cites -- Cardelli, Clayton & Mathis 1989
# (eqs. 2-5), O'Donnell 1994 (optical coefficients), Calzetti et al. 2000 (eq. 4), Fitzpatrick 1999 (Fitzpatrick &
# Massa 1990 UV curve with the 1999 parameters, cubic spline through nine anchors at longer wavelengths) -- and
# anchored on what test/extinction_test.py:92-141 checks (zero A_V is the identity, extinction dims, host + MW
# dims more, input not mutated) plus the papers' own fixed points (A(V) = A_V at x = 1.82 / micron).
# --------------------------------------------------------------------------------------------------
_CCM_A_OPT = [1.0, 0.17699, -0.50447, -0.02427, 0.72085, 0.01979, -0.77530, 0.32999]          # CCM89 eq. 3a, ascending in y
_CCM_B_OPT = [0.0, 1.41338, 2.28305, 1.07233, -5.38434, -0.62251, 5.30260, -2.09002]          # eq. 3b
_OD94_A_OPT = [1.0, 0.104, -0.609, 0.701, 1.137, -1.718, -0.827, 1.647, -0.505]               # O'Donnell 1994
_OD94_B_OPT = [0.0, 1.952, 2.908, -3.989, -7.985, 11.102, 5.491, -10.805, 3.347]


def _ccm89_like(wave, a_v, r_v, a_opt, b_opt):
    x = 1e4 / np.asarray(wave, dtype=float)
    a, b = np.empty_like(x), np.empty_like(x)
    ir, opt, uv, fuv = x < 1.1, (x >= 1.1) & (x < 3.3), (x >= 3.3) & (x < 8.0), x >= 8.0
    a[ir], b[ir] = 0.574 * x[ir] ** 1.61, -0.527 * x[ir] ** 1.61                                # eq. 2
    y = x[opt] - 1.82
    a[opt] = np.polynomial.polynomial.polyval(y, a_opt)
    b[opt] = np.polynomial.polynomial.polyval(y, b_opt)
    xu = x[uv]
    yf = np.where(xu >= 5.9, xu - 5.9, 0.0)
    a[uv] = 1.752 - 0.316 * xu - 0.104 / ((xu - 4.67) ** 2 + 0.341) - 0.04473 * yf ** 2 - 0.009779 * yf ** 3   # eq. 4
    b[uv] = -3.090 + 1.825 * xu + 1.206 / ((xu - 4.62) ** 2 + 0.263) + 0.2130 * yf ** 2 + 0.1207 * yf ** 3
    y = x[fuv] - 8.0
    a[fuv] = -1.073 - 0.628 * y + 0.137 * y ** 2 - 0.070 * y ** 3                                # eq. 5
    b[fuv] = 13.670 + 4.257 * y - 0.420 * y ** 2 + 0.374 * y ** 3
    return a_v * (a + b / r_v)


def extinction_ccm89(wave, a_v, r_v):
    return _ccm89_like(wave, a_v, r_v, _CCM_A_OPT, _CCM_B_OPT)


def extinction_odonnell94(wave, a_v, r_v):
    return _ccm89_like(wave, a_v, r_v, _OD94_A_OPT, _OD94_B_OPT)


def extinction_calzetti00(wave, a_v, r_v=4.05):
    """Calzetti et al. 2000 eq. 4: k(lambda) over 0.12-0.63 and 0.63-2.2 micron, A = A_V k / R_V with k = k' + R_V.
    The reference calls the package without R_V (:131-133, 'Calzetti law doesn't use R_V'): 4.05, the paper's value."""
    x = 1e4 / np.asarray(wave, dtype=float)
    k = np.where(x > 1.0 / 0.63, 2.659 * (-2.156 + 1.509 * x - 0.198 * x ** 2 + 0.011 * x ** 3),
                 2.659 * (-1.857 + 1.040 * x))
    return a_v * (1.0 + k / r_v)


_F99_ANCHORS_AA = np.array([np.inf, 26500.0, 12200.0, 6000.0, 5470.0, 4670.0, 4110.0, 2700.0, 2600.0])


def _fm90_k(x, r_v):
    """Fitzpatrick & Massa 1990 with Fitzpatrick 1999's x0 = 4.596, gamma = 0.99, c3 = 3.23, c4 = 0.41 and the R_V
    dependent c2 = -0.824 + 4.717 / R_V, c1 = 2.030 - 3.007 c2: E(lambda - V) / E(B - V)."""
    x = np.asarray(x, dtype=float)
    c2 = -0.824 + 4.717 / r_v
    c1 = 2.030 - 3.007 * c2
    drude = x ** 2 / ((x ** 2 - 4.596 ** 2) ** 2 + (x * 0.99) ** 2)
    y = np.where(x >= 5.9, x - 5.9, 0.0)
    return c1 + c2 * x + 3.23 * drude + 0.41 * (0.5392 * y ** 2 + 0.05644 * y ** 3)


def extinction_fitzpatrick99(wave, a_v, r_v=3.1):
    """Fitzpatrick 1999: FM90 below 2700 A, NATURAL cubic spline through the nine anchors of the IDL fm_unred routine at
    longer wavelengths.  The natural end condition reproduces the package's documented example
    fitzpatrick99([2000, 4000, 8000], 1.0, 3.1) = [2.76225609, 1.42325373, 0.55333671] to every printed digit
    (not-a-knot gives 1.42339332, 0.55307718)."""
    from scipy.interpolate import CubicSpline
    x = 1e4 / np.asarray(wave, dtype=float)
    xk = 1e4 / _F99_ANCHORS_AA
    kk = np.array([0.0, 0.26469 * r_v / 3.1, 0.82925 * r_v / 3.1,
                   -0.422809 + 1.00270 * r_v + 2.13572e-04 * r_v ** 2,
                   -5.13540e-02 + 1.00216 * r_v - 7.35778e-05 * r_v ** 2,
                   0.700127 + 1.00184 * r_v - 3.32598e-05 * r_v ** 2,
                   1.19456 + 1.01707 * r_v - 5.46959e-03 * r_v ** 2 + 7.97809e-04 * r_v ** 3 - 4.45636e-05 * r_v ** 4,
                   0.0, 0.0]) - r_v
    kk[7:] = _fm90_k(xk[7:], r_v)
    spline = CubicSpline(xk, kk, bc_type="natural")
    uv = x >= 1e4 / 2700.0
    k = np.where(uv, _fm90_k(x, r_v), spline(np.where(uv, xk[7], x)))
    return a_v / r_v * (k + r_v)


EXTINCTION_LAWS = dict(fitzpatrick99=extinction_fitzpatrick99, calzetti00=extinction_calzetti00,
                       odonnell94=extinction_odonnell94, ccm89=extinction_ccm89)


def perform_extinction(flux_density, angstroms, av_host, rv_host, av_mw=0.0, rv_mw=3.1, host_law="fitzpatrick99",
                       mw_law="fitzpatrick99", *, redshift):
    """extinction_models._perform_extinction (:77-170); `flux_density` [..., len(angstroms)]."""
    for law, who in ((host_law, "host"), (mw_law, "MW")):
        if law not in ("fitzpatrick99", "fm07", "calzetti00", "odonnell94", "ccm89"):              # :112-117
            raise ValueError(f"Unknown {who} extinction law: {law}")
    angstroms = np.atleast_1d(np.asarray(angstroms, dtype=float))
    out = np.array(flux_density, dtype=float, copy=True)                                           # :119
    for a_v, r_v, law, wave in ((av_host, rv_host, host_law, angstroms / (1 + redshift)),          # :122-142
                                (av_mw, rv_mw, mw_law, angstroms)):                                # :145-165
        if not a_v > 0:
            continue
        fn = EXTINCTION_LAWS[law]
        mag = fn(wave, a_v) if law == "calzetti00" else fn(wave, a_v, r_v)
        if a_v < 10:
            mag = np.where(mag > 10, 0.0, mag)                                                     # :137-139
        out = out * 10.0 ** (-0.4 * mag)                                                           # extinction.apply
    return out
