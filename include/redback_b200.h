/*
 * redback_b200 -- C ABI of the B200-native batched light-curve + Gaussian-likelihood engine.
 *
 * This is the drop-in boundary for redback's per-sample Python hot path.  The
 * reference has no native boundary of its own (it is pure Python); each entry point below names
 * the reference Python interface it replaces (paths relative to the redback source tree).
 *
 * Conventions
 *   - plain pointers + sizes, no torch / numpy types;
 *   - `*_f64` entry points take DEVICE pointers and are stream-ordered (`stream` is a cudaStream_t
 *     passed as void*; NULL = default stream); they never synchronise;
 *   - `*_host` entry points take ordinary (pageable) HOST pointers and return when the outputs are complete.  Inside,
 *     samples stream through a per-device two-slot pipeline (pinned staging block -> one H2D copy per chunk -> kernels
 *     -> D2H into pinned memory -> caller's buffer) on two private streams, so copies overlap compute; the pipeline's
 *     streams and buffers are created once per device and reused (no cudaMalloc / cudaDeviceSynchronize per call).
 *     Pointers INSIDE the config (cfg->t0, cfg->f_nickel: [N]; cfg->upper_limits.sigma_level: [E]) are HOST pointers
 *     for `*_host` calls and device pointers for `*_f64` calls.  One host call at a time per device (internally locked);
 *   - `*_host_rows` is the same with the parameter block given as RB_*_NPARAM separate row pointers (a redback
 *     `params_batch` dict of arrays is exactly that): rows[r] = [N] host array, or NULL for the constant fill[r];
 *   - per-sample parameters are struct-of-arrays: params[row * N + sample];
 *   - return 0 on success or a negative rb_status; rb_last_error() gives a message;
 *   - there is no CPU fallback: without a usable CUDA device every compute call fails with
 *     RB_ERR_CUDA.
 */
#ifndef REDBACK_B200_H
#define REDBACK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  RB_OK = 0,
  RB_ERR_INVALID = -1,      /* bad argument (maps to Python ValueError)            */
  RB_ERR_UNSUPPORTED = -2,  /* valid in the reference, not implemented on device   */
  RB_ERR_CUDA = -3          /* CUDA runtime error / no device                      */
} rb_status;

/* ---- luminosity engines: redback/transient_models/supernova_models.py:564-579 (_nickelcobalt_engine),
 *      redback/transient_models/magnetar_models.py:171-185 (basic_magnetar) ---------------------- */
enum { RB_ENGINE_NICKEL = 0, RB_ENGINE_MAGNETAR = 1, RB_ENGINE_CSM = 2,
       RB_ENGINE_MAGNETAR_NICKEL = 3, /* magnetar_nickel (supernova_models.py:1628-1681): both engines summed; the
                                         magnetar rows as for RB_ENGINE_MAGNETAR, f_nickel through cfg.f_nickel */
       RB_ENGINE_MAGNETAR_ONLY = 4,   /* general_magnetar_slsn (:2379-2440): l0 (1 + t/tsd)^((1+nn)/(1-nn));
                                         rows 1-3 = l0 [erg/s], tsd [d], nn */
       RB_ENGINE_EXP_POWERLAW = 5,    /* sn_exponential_powerlaw (:357-381, :501-560; phenomenological_models.py:743-754):
                                         lbol_0 (1 - e^(-t/tp))^a1 (t/tp)^(-a2); rows 1-4 = lbol_0, alpha_1, alpha_2,
                                         tpeak_d */
       RB_ENGINE_TDE_FALLBACK = 6,    /* tde_analytical (tde_models.py:16-27, 682-760): l0 / (max(t, t0) 86400 s)^(5/3);
                                         rows 1-2 = l0, t_0_turn [d] */
       RB_ENGINE_FALLBACK = 7,        /* sn_fallback (supernova_models.py:383-438; phenomenological_models.py:690-702):
                                         rows 1-2 = logl1, tr [d]; never diffused (no_interaction implied) */
       RB_ENGINE_FALLBACK_NICKEL = 8  /* sn_nickel_fallback (:441-498): + nickel-cobalt decay; row 3 = f_nickel */ };
enum { RB_PREC_F64 = 0, RB_PREC_F32 = 1 };

/* ---- output stage: bolometric models return L_bol; flux_density models go through
 *      photosphere.TemperatureFloor (photosphere.py:56-111) and sed.Blackbody (sed.py:475-502) or
 *      sed.CutoffBlackbody (sed.py:301-426) ------------------------------------------------------ */
/* RB_SED_CUTOFF_LINE = CutoffBlackbody modified by sed.Line (sed.py:753-909): redback's `type_1a`
 * (supernova_models.py:2229-2314) with the nickel engine. */
enum { RB_SED_BOLOMETRIC = 0, RB_SED_BLACKBODY = 1, RB_SED_CUTOFF_BLACKBODY = 2, RB_SED_CUTOFF_LINE = 3 };

/* ---- noise model of the fused likelihood reduction ------------------------------------------- */
enum {
  RB_SIGMA_FIXED = 0,      /* GaussianLikelihood(sigma=array|float)   likelihoods.py:146-231        */
  RB_SIGMA_SAMPLED = 1,    /* GaussianLikelihood(sigma=None): `sigma` is a sampled parameter :179-185 */
  RB_SIGMA_QUADRATURE = 2, /* GaussianLikelihoodQuadratureNoise: sqrt(sigma_i^2 + sigma^2) :738-790 */
  RB_SIGMA_FRACTIONAL = 3, /* GaussianLikelihoodWithFractionalNoise: sqrt(sigma_i^2 model^2) :792-851 (no sampled sigma) */
  RB_SIGMA_SYSTEMATIC = 4  /* GaussianLikelihoodWithSystematicNoise: sqrt(sigma_i^2 + model^2 sigma^2) :853-913 */
};

/* Upper limits (GaussianLikelihoodWithUpperLimits, likelihoods.py:234-489): epochs with sigma_level[e] > 0 are
 * non-detections at that many sigma; they contribute log(clip(Phi((y - model)/sigma_meas))) with
 * sigma_meas = y / level (flux-like data) or (2.5/ln 10)/level and the survival function (magnitudes, :392-412)
 * instead of the Gaussian term.  sigma_level is a DEVICE pointer [E] or NULL (all detections). */
typedef struct {
  const double* sigma_level;
  int magnitude_mode;   /* data_mode == 'magnitude' */
} rb_upper_limits;

/* Row indices of the SoA parameter block of the supernova path (rows not used by the chosen
 * engine / sed are ignored and may hold anything). */
enum {
  RB_SN_REDSHIFT = 0,          /* ignored (0) for RB_SED_BOLOMETRIC */
  RB_SN_F_NICKEL = 1,          /* nickel engine */
  RB_SN_P0 = 1,                /* magnetar engine: p0 [ms] */
  RB_SN_BP = 2,                /*   bp [1e14 G]            */
  RB_SN_MASS_NS = 3,           /*   mass_ns [Msun]         */
  RB_SN_THETA_PB = 4,          /*   theta_pb [rad]         */
  RB_SN_CSM_MASS = 1,          /* CSM engine (supernova_models.py:1910-1997): csm_mass [Msun] */
  RB_SN_ETA = 2,               /*   eta: CSM density profile exponent, table range [0, 2]     */
  RB_SN_RHO = 3,               /*   rho: CSM density amplitude [g cm^-3]                      */
  RB_SN_R0 = 4,                /*   r0: inner CSM radius [AU]                                 */
  RB_SN_MEJ = 5,               /* [Msun] */
  RB_SN_VEJ = 6,               /* [km/s] */
  RB_SN_KAPPA = 7,
  RB_SN_KAPPA_GAMMA = 8,
  RB_SN_TEMPERATURE_FLOOR = 9, /* [K]; unused for RB_SED_BOLOMETRIC */
  RB_SN_NPARAM = 10
};

/* Flat LambdaCDM with massive neutrinos, the subset of astropy's FlatLambdaCDM that
 * `cosmology.luminosity_distance(redshift).cgs.value` (supernova_models.py:997) needs. */
typedef struct {
  double hubble_distance_cm;
  double Om0, Ogamma0, Ode0;
  double neff_per_nu;
  double nu_y[3];      /* m_nu / (k_B T_nu0) of the massive species */
  int n_massive;
  int n_massless;
} rb_cosmology;

/* Fill `out` from (H0 [km/s/Mpc], Om0, Tcmb0 [K], Neff, m_nu[3] [eV]). */
int rb_cosmology_flat_lcdm(double H0, double Om0, double Tcmb0, double Neff, const double m_nu[3],
                           rb_cosmology* out);
/* astropy.cosmology.Planck18 (redback's default `cosmo`, supernova_models.py:996). */
int rb_cosmology_planck18(rb_cosmology* out);

typedef struct {
  int engine;               /* RB_ENGINE_*  */
  int sed;                  /* RB_SED_*     */
  int sigma_mode;           /* RB_SIGMA_*   */
  int dense_resolution;     /* kwargs 'dense_resolution', default 1000 (supernova_models.py:964) */
  double cutoff_wavelength; /* Angstrom, default 3000 (supernova_models.py:1582)                 */
  double absorption_index;  /* default 1.0; 4 - index must be a positive integer on device       */
  rb_cosmology cosmology;   /* used only when dl_cm == NULL                                      */
  /* sed.Line parameters (RB_SED_CUTOFF_LINE), kwargs of type_1a (supernova_models.py:2255-2259) */
  double line_wavelength;   /* Angstrom, default 7.5e3 */
  double line_width;        /* Angstrom, default 500   */
  double line_time;         /* days, default 50        */
  double line_duration;     /* days, default 25        */
  double line_amplitude;    /* default 0.3             */
  /* RB_PREC_F32: the diffusion quadrature (the hot loop) runs in FP32 with MUFU exponentials on a
   * cancellation-free form of the exponent; engine tables, photosphere, SED and the likelihood reduction stay
   * FP64.  Target: rtol 1e-5 on model outputs (north_star).  Inputs / outputs remain float64. */
  int precision;            /* RB_PREC_F64 (default) | RB_PREC_F32 */
  rb_upper_limits upper_limits;
  /* RB_ENGINE_CSM only: kwargs nn (12), delta (1), efficiency (0.5) of _csm_engine (supernova_models.py:1934-1936);
     interaction process = CSMDiffusion, method "cumulative" (interaction_processes.py:176-197); dense grid =
     get_optimal_time_array(min(0.1, min t), max t + 100, dense_resolution, user_times = t) (:2026-2030) */
  double csm_nn, csm_delta, csm_efficiency;
  /* phase_models.t0_base_model (redback/transient_models/phase_models.py:19-59): per-sample explosion epoch, DEVICE
     pointer [N] or NULL.  The model is evaluated at t_obs - t0[n]; epochs before t0 return 0 (flux density, flux,
     luminosity) or 1000 (magnitude), and the dense grid ends 100 d after the last epoch with t_obs >= t0
     (RB_ENGINE_NICKEL / RB_ENGINE_MAGNETAR only). */
  const double* t0;
  /* RB_ENGINE_MAGNETAR_NICKEL only: per-sample nickel fraction, DEVICE pointer [N] (row 1 holds p0 there). */
  const double* f_nickel;
  /* interaction_process=None (supernova_models.py:961-968 and the like): the engine luminosity at the epochs goes
     straight to the photosphere / SED; mej, kappa, kappa_gamma are then unused. */
  int no_interaction;
  /* interaction_process=AsphericalDiffusion (interaction_processes.py:66-131): Diffusion times
     1 + 1.4 (2 + u) / (1 + e^u) (area_projection / area_reference - 1), u = t / tau_d / 0.59; nickel / magnetar
     engines, no t0.  area_ratio = area_projection / area_reference (kwargs, per-call scalars). */
  int aspherical;
  double area_ratio;
  /* interaction_process=Viscous (interaction_processes.py:229-273; the operator of the TDE fallback models,
     tde_models.py:1127): L convolved with exp(-(t_e - t) / t_viscous) / t_viscous on 1000 abscissae per epoch instead
     of Diffusion.  t_viscous [days] travels in the RB_SN_KAPPA row of the parameter block (Viscous takes no opacity;
     mej / kappa_gamma are unused).  Every engine except RB_ENGINE_CSM; FP64; no t0. */
  int viscous;
  /* Epochs before the start of the dense grid (t < 0): Diffusion / Viscous leave them out of `uniq_times` and
     `np.searchsorted(uniq_times, time)` (interaction_processes.py:63, :271) hands them the luminosity of the FIRST valid
     epoch, while photosphere and SED see the negative time itself.  The smallest t_obs >= 0 of the call (observer
     frame, days); read when negative_epochs != 0 (set by the caller when some t_obs < 0; Diffusion: nickel, magnetar and
     magnetar_nickel engines without t0 / AsphericalDiffusion; Viscous: every engine). */
  double t_first_valid;
  /* RB_ENGINE_CSM only.  csm_quadrature != 0: CSMDiffusion's legacy per-epoch integral (kwarg csm_diffusion_method =
     "quadrature" | "legacy", interaction_processes.py:199-227) on csm_timesteps abscissae (kwarg
     csm_diffusion_timesteps, default 3000, even, <= 6000) instead of the cumulative recurrence.
     cfg.f_nickel != NULL with RB_ENGINE_CSM selects csm_nickel (supernova_models.py:2120-2165): the nickel-cobalt engine
     diffused through mej + mass_csm_threshold on the same grid is added; kappa_gamma is then read from its row. */
  int csm_quadrature;
  int csm_timesteps;
  int negative_epochs;      /* see t_first_valid */
} rb_sn_config;

/* Defaults: nickel engine, Blackbody, fixed sigma, D = 1000, Planck18. */
int rb_sn_config_default(rb_sn_config* cfg);

/*
 * Fused supernova light curve (+ optional Gaussian log-likelihood) for N parameter samples x E epochs.
 *
 * Replaces, per sample, the reference call chain
 *   GaussianLikelihood.log_likelihood            redback/likelihoods.py:221-231
 *     -> arnett | basic_magnetar_powered | slsn  redback/transient_models/supernova_models.py:971-1007, 1471-1509, 1552-1597
 *        (or the *_bolometric functions :949-969, :1447-1469, :1532-1549)
 *        -> Diffusion.convert_input_luminosity   redback/interaction_processes.py:33-65
 *        -> TemperatureFloor                     redback/photosphere.py:56-111
 *        -> Blackbody | CutoffBlackbody          redback/sed.py:199-222, 301-426
 *
 *   params      [RB_SN_NPARAM][N]  device, SoA
 *   t_obs       [E]  observer-frame days (source-frame for RB_SED_BOLOMETRIC); must be >= 0 and
 *                    max(t_obs) <= t_obs[E-1] + 100  (outside that the reference indexes out of range)
 *   freq_obs    [E]  observer-frame Hz (ignored for RB_SED_BOLOMETRIC; may be NULL then)
 *   dl_cm       [N]  luminosity distance in cm, or NULL to integrate cfg->cosmology on device
 *   y, sigma    [E]  data and (fixed or sigma_i) noise; y == NULL skips the likelihood
 *   sigma_param [N]  sampled `sigma` for RB_SIGMA_SAMPLED / RB_SIGMA_QUADRATURE, else NULL
 *   out_model   [N][E] row-major model output (erg/s bolometric, mJy otherwise) or NULL
 *   out_logl    [N]  np.nan_to_num(log L) or NULL
 */
int rb_sn_eval_f64(const rb_sn_config* cfg, int64_t N, int64_t E, const double* params,
                   const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                   const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                   void* stream);

/* Same computation with HOST buffers: chunked, double-buffered H2D -> kernels -> D2H inside the call (see the
 * conventions at the top).  This is what a non-Python caller binds; redback_b200.likelihoods reaches it through
 * GaussianLikelihood.log_likelihood_batch when the batch is host-resident. */
int rb_sn_eval_f64_host(const rb_sn_config* cfg, int64_t N, int64_t E, const double* params,
                        const double* t_obs, const double* freq_obs, const double* dl_cm,
                        const double* y, const double* sigma, const double* sigma_param,
                        double* out_model, double* out_logl);
int rb_sn_eval_f64_host_rows(const rb_sn_config* cfg, int64_t N, int64_t E, const double* const* rows,
                             const double* fill, const double* t_obs, const double* freq_obs,
                             const double* dl_cm, const double* y, const double* sigma,
                             const double* sigma_param, double* out_model, double* out_logl);

/* ---- kilonova path ----------------------------------------------------------------------------------
 * one_component_kilonova_model   redback/transient_models/kilonova_models.py:1537-1603 (+ :1853-1893)
 * metzger_kilonova_model         redback/transient_models/kilonova_models.py:1895-1962 (+ :1965-2068)
 * both in output_format='flux_density', on the reference's adaptive grid
 * get_optimal_time_array(1e-2, 7e6, dense_resolution, user_times) (redback/utils.py:42-80). */
enum { RB_KN_ONE_COMPONENT = 0, RB_KN_METZGER = 1 };
enum {
  RB_KN_REDSHIFT = 0,
  RB_KN_MEJ = 1,               /* [Msun]                      */
  RB_KN_VEJ = 2,               /* [c]                         */
  RB_KN_KAPPA = 3,
  RB_KN_TEMPERATURE_FLOOR = 4, /* one-component: [K]          */
  RB_KN_BETA = 4,              /* Metzger: velocity power law */
  RB_KN_NPARAM = 5
};

typedef struct {
  int model;                    /* RB_KN_*                                                       */
  int sigma_mode;               /* RB_SIGMA_*                                                    */
  int dense_resolution;         /* kwargs 'dense_resolution', default 500 (kilonova_models.py:1558) */
  int neutron_precursor_switch; /* Metzger only, default 1 (kilonova_models.py:1977)             */
  double vmax;                  /* Metzger only, default 0.7 (kilonova_models.py:1997)           */
  rb_cosmology cosmology;       /* used only when dl_cm == NULL                                  */
  rb_upper_limits upper_limits;
  double grid_t_max;            /* end of the model time grid [s]: 7e6 (kilonova_models.py:1561, default); the
                                   two-component model uses 86400*6 (:1241)                       */
  const double* t0;             /* device [N] or NULL: phase_models.t0_base_model (phase_models.py:19-59) — epochs are
                                   t_obs - t0[n] (ascending t_obs), those before t0 return 0 (1000 mag), and the
                                   sample's model grid follows ITS valid epochs                   */
  int precision;                /* RB_PREC_F64 (default, memset 0) | RB_PREC_F32: one_component_kilonova_model only -- the
                                   smooth per-grid-point factors (thermalisation efficiency and the r-process heating
                                   shape: exp, log1p, atan, three pow) are evaluated in FP32, the exp(+-t^2/tau^2) pair,
                                   the cumulative trapezoid, photosphere, SED and likelihood stay FP64.  Target: rtol
                                   1e-5 on model outputs (north_star).  Inputs / outputs remain float64.             */
} rb_kn_config;

int rb_kn_config_default(rb_kn_config* cfg);

/* Same argument conventions as rb_sn_eval_f64; params is [RB_KN_NPARAM][N]; t_obs in observer-frame days
 * (> 0); epochs whose source-frame time falls outside the model grid [1e-2 s, 7e6 s] (where scipy's
 * interp1d raises ValueError in the reference) come back as NaN. */
int rb_kn_eval_f64(const rb_kn_config* cfg, int64_t N, int64_t E, const double* params, const double* t_obs,
                   const double* freq_obs, const double* dl_cm, const double* y, const double* sigma,
                   const double* sigma_param, double* out_model, double* out_logl, void* stream);
int rb_kn_eval_f64_host(const rb_kn_config* cfg, int64_t N, int64_t E, const double* params,
                        const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                        const double* sigma, const double* sigma_param, double* out_model, double* out_logl);
int rb_kn_eval_f64_host_rows(const rb_kn_config* cfg, int64_t N, int64_t E, const double* const* rows,
                             const double* fill, const double* t_obs, const double* freq_obs,
                             const double* dl_cm, const double* y, const double* sigma,
                             const double* sigma_param, double* out_model, double* out_logl);

/* ---- redback-native afterglow path --------------------------------------------------------------------
 * tophat_redback    redback/transient_models/afterglow_models/base_models.py:885-946
 * gaussian_redback  redback/transient_models/afterglow_models/base_models.py:948-1010
 * through RedbackAfterglows.get_lightcurve (:573-640) and the numba kernels (:43-400),
 * output_format='flux_density'. */
/* RedbackAfterglows._method_to_id (:567-571): 'TH', 'GJ', 'PL', 'PL2', '2C', 'DG' =
 * tophat_redback, gaussian_redback, powerlaw_redback (:1081), alternativepowerlaw_redback (:1150),
 * twocomponent_redback (:1013), doublegaussian_redback (:1218) */
enum { RB_AG_TOPHAT = 1, RB_AG_GAUSSIAN = 2, RB_AG_POWERLAW = 3, RB_AG_POWERLAW_ALT = 4, RB_AG_TWO_COMPONENT = 5,
       RB_AG_DOUBLE_GAUSSIAN = 6 };
enum {
  RB_AG_REDSHIFT = 0, RB_AG_THV = 1, RB_AG_LOGE0 = 2, RB_AG_THC = 3,
  RB_AG_THJ = 4,      /* every jet except the tophat (which uses thj = thc, :937) */
  RB_AG_LOGN0 = 5, RB_AG_P = 6, RB_AG_LOGEPSE = 7, RB_AG_LOGEPSB = 8, RB_AG_G0 = 9, RB_AG_XIN = 10,
  RB_AG_NPARAM = 11
};

typedef struct {
  int jet;            /* RB_AG_*                                                              */
  int sigma_mode;     /* RB_SIGMA_*                                                           */
  int steps;          /* kwargs 'steps', default 250 (:936)                                   */
  int res;            /* kwargs 'res'; 0 = the reference's rule max(10, min(int((2-10*max(thv-thc,0))*thc*g0), 100)) (:932-935) */
  int k;              /* 0 (ISM) or 2 (wind), kwargs 'k' (:918)                               */
  int expansion;      /* kwargs 'expansion', default 1 (:920)                                 */
  double a1;          /* kwargs 'a1', default 1 (:919)                                        */
  rb_cosmology cosmology;
  double ss, aa;      /* kwargs 'ss', 'aa': extra structure parameters of the PL / PL2 / 2C / DG jets */
  rb_upper_limits upper_limits;
  int ab_magnitude;   /* output_format='magnitude': monochromatic AB magnitude of the mJy flux density
                         (utils.calc_ABmag_from_flux_density, utils.py:481-488; base_models.py:944-946) */
  int precision;      /* RB_PREC_F64 (default, memset 0) | RB_PREC_F32: the synchrotron spectrum of a cell step (one log, one
                         exp per frequency) through single-precision hardware transcendentals on FP64-reduced arguments
                         (relative error ~2e-7 per term); dynamics, observer times, interpolation and sums stay FP64.
                         Target: rtol 1e-5 on model outputs (north_star) */
} rb_ag_config;

int rb_ag_config_default(rb_ag_config* cfg);

/* Argument conventions as rb_sn_eval_f64; params is [RB_AG_NPARAM][N]; t_obs in observer-frame days;
 * at most 4 distinct values in freq_obs.  Samples with p outside [1.2, 3.4] (ValueError in the reference,
 * :576-577) come back as NaN. */
int rb_ag_eval_f64(const rb_ag_config* cfg, int64_t N, int64_t E, const double* params, const double* t_obs,
                   const double* freq_obs, const double* dl_cm, const double* y, const double* sigma,
                   const double* sigma_param, double* out_model, double* out_logl, void* stream);
int rb_ag_eval_f64_host(const rb_ag_config* cfg, int64_t N, int64_t E, const double* params,
                        const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                        const double* sigma, const double* sigma_param, double* out_model, double* out_logl);
int rb_ag_eval_f64_host_rows(const rb_ag_config* cfg, int64_t N, int64_t E, const double* const* rows,
                             const double* fill, const double* t_obs, const double* freq_obs,
                             const double* dl_cm, const double* y, const double* sigma,
                             const double* sigma_param, double* out_model, double* out_logl);

/* ---- magnetar-driven relativistic ejecta ("mergernova") path ---------------------------------------------
 * basic_mergernova / general_mergernova / general_mergernova_thermalisation
 * (redback/transient_models/magnetar_driven_ejecta_models.py:275-419): magnetar luminosity on
 * get_optimal_time_array(1e-4, 1e8, 500) [s, source frame] -> _ejecta_dynamics_and_interaction (:14-151, explicit
 * Euler of Lorentz factor, radius, comoving volume and internal energy) -> interp1d of comoving temperature,
 * radius and Doppler factor at the epochs -> _comoving_blackbody_to_flux_density (:153-174) -> mJy (1+z)^2
 * (:262, :272) -> Gaussian log-likelihood. */
enum { RB_MN_BASIC_MAGNETAR = 0,   /* basic_magnetar(p0, bp, mass_ns, theta_pb)  (magnetar_models.py:171-185) */
       RB_MN_MAGNETAR_ONLY = 1,    /* magnetar_only(l0, tau_sd, nn)              (magnetar_models.py:134-144) */
       RB_MN_EVOLVING = 2 };       /* _evolving_gw_and_em_magnetar(bint, bext, p0 [s], chi0, radius, moi)
                                      (magnetar_models.py:187-224): explicit Euler of the spin frequency under EM + GW
                                      torques on the model grid, L = Edot_d; general_mergernova_evolution (:422-474) and
                                      general_metzger_magnetar_driven_evolution (:908-956), both with gamma-ray opacity */
enum {
  RB_MN_REDSHIFT = 0,
  RB_MN_MEJ = 1,            /* [Msun] */
  RB_MN_BETA = 2,           /* initial ejecta velocity [c] */
  RB_MN_EJECTA_RADIUS = 3,  /* [cm] */
  RB_MN_KAPPA = 4,
  RB_MN_N_ISM = 5,          /* [cm^-3] */
  RB_MN_P0 = 6, RB_MN_L0 = 6,
  RB_MN_BP = 7, RB_MN_TAU_SD = 7,
  RB_MN_MASS_NS = 8, RB_MN_NN = 8,
  RB_MN_THETA_PB = 9,
  RB_MN_THERMALISATION_EFFICIENCY = 10, RB_MN_KAPPA_GAMMA = 10,   /* by use_gamma_ray_opacity */
  /* RB_MN_EVOLVING: rows 6-9 = logbint, logbext, p0 [s], chi0; row 10 = kappa_gamma; then */
  RB_MN_LOGBINT = 6, RB_MN_LOGBEXT = 7, RB_MN_EV_P0 = 8, RB_MN_CHI0 = 9,
  RB_MN_NS_RADIUS = 11,     /* [km] */
  RB_MN_LOGMOI = 12,
  RB_MN_NPARAM = 13
};

typedef struct {
  int engine;                  /* RB_MN_BASIC_MAGNETAR | RB_MN_MAGNETAR_ONLY */
  int sigma_mode;              /* RB_SIGMA_* */
  int use_gamma_ray_opacity;   /* general_mergernova_thermalisation (:405-410): row 10 is kappa_gamma */
  int pair_cascade_switch;     /* kwargs 'pair_cascade_switch' (:131-135); default 0 basic, 1 general */
  double ejecta_albedo;        /* kwargs 'ejecta_albedo', default 0.5 */
  double pair_cascade_fraction;/* kwargs 'pair_cascade_fraction', default 0.05 */
  int resolution;              /* points requested from get_optimal_time_array(1e-4, 1e8, .), 500 (:301) */
  rb_cosmology cosmology;
  rb_upper_limits upper_limits;
  /* RB_MN_OUT_FLUX_DENSITY (default): mJy at t_obs [days].  trapped_magnetar (:477-573), t_obs in SECONDS, engine
     RB_MN_MAGNETAR_ONLY, no pair cascade: RB_MN_OUT_TRAPPED_LUMINOSITY = exp(-tau) L_sd + comoving-blackbody
     luminosity at freq_obs [erg/s/Hz -> as the reference returns it]; RB_MN_OUT_TRAPPED_FLUX = that luminosity at
     the k-corrected time / frequency over 4 pi d_L^2 (1+z)^(photon_index - 2). */
  int output;
  double photon_index;
  /* general_magnetar_driven_supernova(_bolometric) (supernova_models.py:2464-2584): the same dynamics with nickel
     heating (use_r_process = 0, f_nickel in row RB_MN_F_NICKEL), gamma-ray opacity, grid
     get_optimal_time_array(1e0, 1e8, 2000); RB_MN_OUT_SUPERNOVA: TemperatureFloor on the grid with the time-dependent
     ejecta velocity (temperature floor in row RB_MN_TEMPERATURE_FLOOR), T / R / L interpolated at the epochs [days],
     then `sed` (RB_SED_BLACKBODY | RB_SED_CUTOFF_BLACKBODY) -> mJy (1+z); RB_MN_OUT_LBOL: L_bol(t) [erg/s] at source-
     frame days (the bolometric variant; redshift row ignored). */
  int use_r_process;           /* default 1 */
  double grid_t_min, grid_t_max;   /* [s], defaults 1e-4, 1e8 */
  int sed;
  double cutoff_wavelength, absorption_index;
  const double* t0;            /* device [N] or NULL: phase_models.t0_base_model — epochs are t_obs - t0[n] (same unit as
                                  t_obs), those before t0 return 0 (1000 mag); RB_MN_OUT_FLUX_DENSITY only */
} rb_mn_config;
enum { RB_MN_OUT_FLUX_DENSITY = 0, RB_MN_OUT_TRAPPED_LUMINOSITY = 1, RB_MN_OUT_TRAPPED_FLUX = 2,
       RB_MN_OUT_SUPERNOVA = 3, RB_MN_OUT_LBOL = 4 };
enum { RB_MN_F_NICKEL = 9, RB_MN_TEMPERATURE_FLOOR = 11 };

int rb_mn_config_default(rb_mn_config* cfg);

/* Same argument conventions as rb_sn_eval_f64; params is [RB_MN_NPARAM][N]; t_obs in observer-frame days.  Epochs
 * whose source-frame time falls outside the grid [1e-4 s, 1e8 s] (scipy's interp1d raises there) come back NaN. */
int rb_mn_eval_f64(const rb_mn_config* cfg, int64_t N, int64_t E, const double* params, const double* t_obs,
                   const double* freq_obs, const double* dl_cm, const double* y, const double* sigma,
                   const double* sigma_param, double* out_model, double* out_logl, void* stream);
int rb_mn_eval_f64_host(const rb_mn_config* cfg, int64_t N, int64_t E, const double* params,
                        const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                        const double* sigma, const double* sigma_param, double* out_model, double* out_logl);
int rb_mn_eval_f64_host_rows(const rb_mn_config* cfg, int64_t N, int64_t E, const double* const* rows,
                             const double* fill, const double* t_obs, const double* freq_obs,
                             const double* dl_cm, const double* y, const double* sigma,
                             const double* sigma_param, double* out_model, double* out_logl);

/* utils.get_optimal_time_array(t_min, t_max, resolution) without user times (utils.py:79-93), HOST function:
 * writes at most `resolution` ascending points to out (host memory) and their count to *n_out. */
int rb_optimal_time_array(double t_min, double t_max, int resolution, double* out, int* n_out);

/* ---- magnetar-driven Metzger shell model ------------------------------------------------------------------
 * metzger_magnetar_driven_kilonova_model / general_metzger_magnetar_driven / ..._thermalisation
 * (magnetar_driven_ejecta_models.py:760-905) -> _general_metzger_magnetar_driven_kilonova_model (:576-757,
 * magnetar_heating='first_layer') on get_optimal_time_array(1e-4, 1e7, 300) -> interp1d of T and R_photosphere ->
 * blackbody (:263-272) -> Gaussian log-likelihood.  The engine ids are RB_MN_BASIC_MAGNETAR / RB_MN_MAGNETAR_ONLY. */
enum {
  RB_MK_REDSHIFT = 0,
  RB_MK_MEJ = 1,       /* [Msun] */
  RB_MK_VEJ = 2,       /* minimum ejecta velocity [c] */
  RB_MK_BETA = 3,      /* velocity power-law slope of the mass profile */
  RB_MK_KAPPA_R = 4,
  RB_MK_P0 = 5, RB_MK_L0 = 5,
  RB_MK_BP = 6, RB_MK_TAU_SD = 6,
  RB_MK_MASS_NS = 7, RB_MK_NN = 7,
  RB_MK_THETA_PB = 8,
  RB_MK_THERMALISATION_EFFICIENCY = 9, RB_MK_KAPPA_GAMMA = 9,
  /* RB_MN_EVOLVING: rows 5-8 = logbint, logbext, p0 [s], chi0; row 9 = kappa_gamma; then */
  RB_MK_LOGBINT = 5, RB_MK_LOGBEXT = 6, RB_MK_EV_P0 = 7, RB_MK_CHI0 = 8,
  RB_MK_NS_RADIUS = 10,     /* [km] */
  RB_MK_LOGMOI = 11,
  RB_MK_NPARAM = 12
};

typedef struct {
  int engine;
  int sigma_mode;
  int use_gamma_ray_opacity;     /* ..._thermalisation: row 9 is kappa_gamma (:674-676) */
  int pair_cascade_switch;       /* default 1 (:590) */
  double ejecta_albedo;          /* 0.5 */
  double pair_cascade_fraction;  /* 0.01 (:592) */
  int neutron_precursor_switch;  /* default 1 (:593) */
  double vmax;                   /* 0.7 (:595) */
  int resolution;                /* get_optimal_time_array(1e-4, 1e7, 300) (:777); 500 for ..._evolution (:917) */
  rb_cosmology cosmology;
  rb_upper_limits upper_limits;
  int heat_all_layers;           /* magnetar_heating: 0 'first_layer' (default, :709-721), 1 'all_layers' (:698-704) */
  const double* t0;              /* device [N] or NULL: phase_models.t0_base_model, as in rb_mn_config */
} rb_mk_config;

int rb_mk_config_default(rb_mk_config* cfg);
int rb_mk_eval_f64(const rb_mk_config* cfg, int64_t N, int64_t E, const double* params, const double* t_obs,
                   const double* freq_obs, const double* dl_cm, const double* y, const double* sigma,
                   const double* sigma_param, double* out_model, double* out_logl, void* stream);
int rb_mk_eval_f64_host(const rb_mk_config* cfg, int64_t N, int64_t E, const double* params,
                        const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                        const double* sigma, const double* sigma_param, double* out_model, double* out_logl);
int rb_mk_eval_f64_host_rows(const rb_mk_config* cfg, int64_t N, int64_t E, const double* const* rows,
                             const double* fill, const double* t_obs, const double* freq_obs,
                             const double* dl_cm, const double* y, const double* sigma,
                             const double* sigma_param, double* out_model, double* out_logl);

/* ---- magnitude / bandpass branch (output_format = 'magnitude' | 'flux') ---------------------------------
 * get_correct_output_format_from_spectra -> RedbackTimeSeriesSource.bandmag -> _bandflux_single_redback
 * (redback/sed.py:971-1009, 120-195, 10-118) and bandpass_magnitude_to_flux (redback/utils.py:707-718).
 * Everything in rb_photometry is a HOST pointer: small per-call tables the library uploads itself.
 * The caller prepares, per band b (what sncosmo does on the host in the reference):
 *   - the integration grid  wave = arange(minwave + d/2, maxwave, d), d = range / ceil(range / 5 A)
 *   - int_wt = wave * transmission(wave), band_scale = d / HC_ERG_AA, band_zp = AB zero-point band flux. */
/* ---- dust extinction (redback/transient_models/extinction_models.py:77-170, _perform_extinction) ---------------
 * Host-galaxy extinction at lambda / (1 + z) with (host_law, av_host, rv_host) and Milky-Way extinction at lambda with
 * (mw_law, av_mw, rv_mw); flux * 10^(-0.4 A); A is zeroed where it exceeds 10 mag while A_V < 10.  The laws are the
 * `extinction` package's (third party, absent here: parity unpinned, see csrc/ext_kernel.cu).  For
 * RB_EXT_FITZPATRICK99 the caller supplies the cubic spline through Fitzpatrick's nine anchor points as 9 breakpoints
 * [1/micron] followed by 8 x 4 piecewise-cubic coefficients (highest power first, in x - breakpoint). */
enum { RB_EXT_CCM89 = 1, RB_EXT_ODONNELL94 = 2, RB_EXT_CALZETTI00 = 3, RB_EXT_FITZPATRICK99 = 4 };
typedef struct {
  int host_law, mw_law;         /* RB_EXT_*                                                          */
  double rv_host, rv_mw;        /* 3.1; calzetti00: 4.05 (the package default, redback passes none)  */
  double av_mw;                 /* 0: no Milky-Way extinction                                        */
  double f99_host[41], f99_mw[41];
} rb_extinction;

/* Flux-density branch of _evaluate_extinction_model (:194-221): model[n][e] *= 10^(-0.4 A), in place, with the observed
 * wavelength c / freq_obs[e].  Device pointers: freq_obs [E], redshift [N], av_host [N], model [N][E]. */
int rb_extinction_apply_f64(const rb_extinction* ext, int64_t N, int64_t E, const double* freq_obs,
                            const double* redshift, const double* av_host, double* model, void* stream);

/* type_1c (supernova_models.py:2317-2352): add the epoch-independent sed.Synchrotron flux density (sed.py:703-751) times
 * (1 + z) to a device-resident (N, E) flux-density matrix, in place.  Device pointers: freq_obs [E], redshift / pp /
 * dl_cm [N], model [N][E]; the reference's source radius is 1e13 cm. */
int rb_synchrotron_add_f64(int64_t N, int64_t E, const double* freq_obs, const double* redshift, const double* pp,
                           const double* dl_cm, double nu_max, double f0, double source_radius, double* model,
                           void* stream);

enum { RB_OUT_MAGNITUDE = 0, RB_OUT_FLUX = 1, RB_OUT_SPECTRA = 2 };
typedef struct {
  int output;                   /* RB_OUT_*                                                         */
  int n_lambda;
  const double* lambda_obs;     /* [n_lambda] kwargs 'lambda_array': observer-frame Angstrom grid   */
  int n_tgrid;                  /* supernova path only: the model time grid                          */
  const double* tgrid;          /* [n_tgrid] get_optimal_time_array(0.1, 3000, 300) in days          */
  int n_bands;
  const int* band_of_epoch;     /* [E]                                                               */
  const int64_t* band_off;      /* [n_bands + 1] offsets into int_wave / int_wt                      */
  const double* int_wave;       /* [band_off[n_bands]]                                               */
  const double* int_wt;
  const double* band_scale;     /* [n_bands]                                                         */
  const double* band_zp;        /* [n_bands]                                                         */
  const double* band_ref_flux;  /* [n_bands] tables/filters.csv reference_flux (RB_OUT_FLUX) or NULL */
  /* magnitude / flux branch of _evaluate_extinction_model (:223-251): the spectra are attenuated per observer-frame
     wavelength before the spline fit.  `extinction` HOST pointer or NULL; `ext_av_host` DEVICE pointer [N] (the
     per-sample host-galaxy A_V), required when `extinction` is given. */
  const rb_extinction* extinction;
  const double* ext_av_host;
  /* output == RB_OUT_SPECTRA: output_format='spectra' (supernova_models.py:1019-1023 and its twins) -- the namedtuple
     (time, lambdas, spectra) per sample, before the clean-up of sed.py:985-992.  DEVICE pointers: spectra_out
     [N][n_t][n_lambda] erg/cm^2/s/Angstrom, spectra_time_out [N][n_t] observer-frame days, spectra_len_out [N] valid
     grid points of sample n (rows beyond are NaN); n_t = n_tgrid (supernova family), dense_resolution (kilonovae) or the
     model's fixed grid (magnetar-driven models).  The band tables, out_model and out_logl are ignored. */
  double* spectra_out;
  double* spectra_time_out;
  int32_t* spectra_len_out;
} rb_photometry;

/* Magnitude-branch twins of rb_sn_eval_f64 / rb_kn_eval_f64: same device arguments minus freq_obs
 * (supernova_models.py:1008-1028, 1510-1530, 1599-1624; kilonova_models.py:1579-1603, 1937-1962).
 * cfg->sed selects Blackbody / CutoffBlackbody for the supernova path; RB_SED_BOLOMETRIC is invalid. */
int rb_sn_mag_eval_f64(const rb_sn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                       const double* params, const double* t_obs, const double* dl_cm, const double* y,
                       const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                       void* stream);
int rb_kn_mag_eval_f64(const rb_kn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                       const double* params, const double* t_obs, const double* dl_cm, const double* y,
                       const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                       void* stream);
/* Magnitude branch of two_component_kilonova_model / three_component_kilonova_model (kilonova_models.py:1166-1211,
 * 1277-1323): the blackbody spectra of n_components one-component models on the shared time grid are summed before
 * the spline / band integration.  params is [n_components][RB_KN_NPARAM][N] (each component carries the redshift
 * row); cfg->model must be RB_KN_ONE_COMPONENT, cfg->grid_t_max the model's grid end. */
int rb_kn_multi_mag_eval_f64(const rb_kn_config* cfg, const rb_photometry* phot, int n_components, int64_t N,
                             int64_t E, const double* params, const double* t_obs, const double* dl_cm,
                             const double* y, const double* sigma, const double* sigma_param, double* out_model,
                             double* out_logl, void* stream);
/* Magnitude branch of the mergernova models (_processing_other_formats, magnetar_driven_ejecta_models.py:198-238,
 * use_relativistic_blackbody = True): spectra on (get_optimal_time_array(1e-4, 1e8, 500) x lambda_array), phases
 * time_temp (1+z) / 86400 d.  phot->tgrid is not used (the library builds the grid itself). */
int rb_mn_mag_eval_f64(const rb_mn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                       const double* params, const double* t_obs, const double* dl_cm, const double* y,
                       const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                       void* stream);
/* The same for the magnetar-driven Metzger shell models (plain blackbody of T and R_photosphere on the model grid). */
int rb_mk_mag_eval_f64(const rb_mk_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                       const double* params, const double* t_obs, const double* dl_cm, const double* y,
                       const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                       void* stream);

/* ---- observation noise + SNR cut for simulated optical light curves ------------------------------------
 * SimulateTransientWithCadence._evaluate_optical_model / _add_optical_noise_and_cuts
 * (redback/simulate_transients.py:412-431, 457-527).  model_mag, z and all outputs are [N][E] (device);
 * ref_flux / limiting_mag are [E]; z holds standard-normal variates (the reference draws rng.normal).
 * obs_flux / flux_err may be NULL (they are NaN for the two Gaussian noise types). */
enum { RB_NOISE_LIMITING_MAG = 0, RB_NOISE_GAUSSIAN = 1, RB_NOISE_GAUSSIANMODEL = 2 };
int rb_optical_noise_f64(int noise_type, int64_t N, int64_t E, const double* model_mag, const double* ref_flux,
                         const double* limiting_mag, const double* z, double snr_threshold, int apply_snr_cut,
                         double* obs_mag, double* mag_err, double* obs_flux, double* flux_err, double* snr,
                         unsigned char* detected, void* stream);

/* The same with the standard-normal variates drawn on device: Philox4x32-10 keyed by `seed`, counter = global element
 * index (sample_base + n) * E + e, two 53-bit uniforms -> Box-Muller.  The draws therefore do not depend on how a
 * population is chunked or sharded over GPUs.  Replaces `np.random.normal` / `rng.normal` of
 * simulate_transients.py:476, 493, 506, 1111 (a different but equally distributed stream; the reference's own stream
 * cannot be reproduced on device).  obs_flux, flux_err, snr and z_out (the variates, for checkers) may be NULL. */
int rb_optical_noise_philox_f64(int noise_type, int64_t N, int64_t E, int64_t sample_base, uint64_t seed,
                                const double* model_mag, const double* ref_flux, const double* limiting_mag,
                                double snr_threshold, int apply_snr_cut, double* obs_mag, double* mag_err,
                                double* obs_flux, double* flux_err, double* snr, unsigned char* detected,
                                double* z_out, void* stream);

/* ---- survey simulation with per-transient pointings --------------------------------------------------------
 * SimulateOpticalTransient._make_observation_single / _make_observations_for_population
 * (redback/simulate_transients.py:1096-1178): every transient is observed at ITS OWN pointings.  The pointing lists
 * are a padded device table [N][E_max] (row n has epoch_count[n] valid entries): `phase` = expMJD - t0 in days
 * (the spline source clamps phases outside the model grid, like sncosmo / FITPACK), `band_index` into the bands of
 * `phot` (whose band_of_epoch must be 0 .. n_bands-1, one entry per band).  Magnitudes as rb_sn_mag_eval_f64;
 * padding entries of out_model are left untouched. */
int rb_sn_mag_eval_ragged_f64(const rb_sn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E_max,
                              const double* params, const double* phase, const int32_t* band_index,
                              const int32_t* epoch_count, const double* dl_cm, double* out_model, void* stream);
/* Noise, magnitude errors and the detection flags of :1100-1134 for that table: five_sigma_depth [N][E_max],
 * band_ref_flux [n_bands] (filters.csv reference_flux), source_noise_factor = 0 switches the source noise off
 * (add_source_noise=False).  Variates: the Philox stream of rb_optical_noise_philox_f64 at element index
 * sample_base + row_start[n] + e, i.e. the pointing's position in the flat (CSR) pointing table, so results do not
 * depend on padding or chunking.  Padding entries: NaN / not detected. */
int rb_survey_noise_philox_f64(int64_t N, int64_t E_max, int64_t sample_base, uint64_t seed,
                               const double* model_mag, const double* phase, const int32_t* band_index,
                               const int32_t* epoch_count, const int64_t* row_start, const double* five_sigma_depth,
                               const double* band_ref_flux, double source_noise_factor, double snr_threshold,
                               double* obs_mag, double* mag_err, double* obs_flux, double* flux_err,
                               unsigned char* detected, double* z_out, void* stream);

/* Luminosity distance [cm] for N redshifts on device (same quadrature the fused kernels use). */
int rb_luminosity_distance_f64(const rb_cosmology* cosmo, int64_t N, const double* redshift,
                               double* out_dl_cm, void* stream);

/*
 * Stand-alone Gaussian log-likelihood reduction over a precomputed model matrix:
 * GaussianLikelihood._gaussian_log_likelihood + np.nan_to_num (likelihoods.py:221-231, 767-790).
 *   model [N][E], y [E], sigma [E], sigma_param [N] or NULL, upper_limits or NULL, out_logl [N]
 */
int rb_gaussian_logl_f64(int sigma_mode, int64_t N, int64_t E, const double* model, const double* y,
                         const double* sigma, const double* sigma_param, const rb_upper_limits* upper_limits,
                         double* out_logl, void* stream);

/* MixtureGaussianLikelihood.log_likelihood (redback/likelihoods.py:542-617) on a device model matrix [N][E]:
 * sum_e logaddexp(log alpha + log N(y - m | sigma_e), log(1 - alpha) + log N(y - m | sigma_out)), alpha [N] and
 * sigma_out [N] per sample, sigma [E].  No nan_to_num (the reference returns the raw sum). */
int rb_mixture_logl_f64(int64_t N, int64_t E, const double* model, const double* y, const double* sigma,
                        const double* alpha, const double* sigma_out, double* out_logl, void* stream);

/* sizeof() of the ABI structs, so that a foreign-language binding can check its own layout before the first call:
 * which = 0 rb_cosmology, 1 rb_upper_limits, 2 rb_sn_config, 3 rb_kn_config, 4 rb_ag_config, 5 rb_mn_config,
 * 6 rb_mk_config, 7 rb_photometry; -1 for anything else. */
int64_t rb_sizeof(int which);

/* FP64 FMA micro-benchmark: runs `iters` dependent-chain DFMAs per thread on a full grid and
 * returns the achieved TFLOP/s (2 flop per FMA) in *out_tflops.  Used as the roofline denominator. */
int rb_measure_fp64_peak(int iters, double* out_tflops, double* out_ms);

/* Number of kernel launches issued by this library in the current process. */
int64_t rb_launch_count(void);

/* Diagnostics of the photometry kernel on the current device: samples processed per clean-up mode since the last reset
 * (out[0] every spectrum entry valid, [1] single pass, no invalid rows, [2] single pass with the row-mask column,
 * [3] two passes; [4..7] reasons single-pass mode was refused).  Synchronises the device.  `out` may be null. */
int rb_phot_mode_counts(uint64_t out[8], int reset);

const char* rb_last_error(void);
const char* rb_version(void);
/* digest of the sources the library was built from (bindings compare it with the tree they ship with) */
const char* rb_source_hash(void);

#ifdef __cplusplus
}
#endif
#endif /* REDBACK_B200_H */
