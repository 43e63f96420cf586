// redback_b200 -- fused supernova light-curve + Gaussian log-likelihood kernels (FP64, sm_100a).
//
// Three launches per call (all on the caller's stream):
//   1. epoch_table_kernel  : per-epoch constants (t, nu, y, 1/sigma, log(2 pi sigma^2)/2), once per launch
//   2. sn_prep_kernel      : one warp per sample: every per-sample scalar of the path (tau_d, trapping
//                            coefficient, engine constants, photosphere/SED constants) and the luminosity
//                            distance by 32-point Gauss-Legendre quadrature -> 16 doubles per sample
//   3. sn_kernel           : persistent thread blocks, grid-stride over samples, one thread per epoch.
//      Per sample the block stages in shared memory
//        * the engine luminosity on the reference's dense grid linspace(0, t[-1]+100, D) as the
//          segments a_k + b_k t of scipy's linear interp1d              (interaction_processes.py:44,56)
//        * the merged quadrature abscissae xm = unique(lsp U 1-lsp) with their trapezoid weights
//                                                                        (interaction_processes.py:49-51,60)
//      and every thread walks the <=100 abscissae of its epoch sequentially in registers
//      (interaction_processes.py:53-61).  Photosphere (photosphere.py:87-111), SED (sed.py:199-222 /
//      301-426) and the Gaussian log-likelihood term (likelihoods.py:229-231) are applied by the same
//      thread and the likelihood is reduced across the block.
//
// The roofline that bounds sn_kernel is the FP64 pipe: 14 FP64 instructions per abscissa;
// HBM traffic is 8*P bytes in and 8 (+8E) bytes out per sample plus the 128-byte scalar record.
#include "rb_common.cuh"
#include "rb_tables.inc"

namespace rb {

constexpr int kNodesHalf = 50;            // int(round(timesteps / 2.0)), interaction_processes.py:34,49
constexpr int kNodes = 2 * kNodesHalf;    // merged abscissae, duplicates kept (zero-width intervals)
constexpr double kExpSkip = -690.0;       // exp() arguments below this contribute < 1e-299: skipped

enum { EP_T = 0, EP_NU, EP_Y, EP_ISIG, EP_LOGT, EP_KIND, kEpochCols };

// per-sample scalar record written by sn_prep_kernel
enum {
  SC_OPZ = 0,    // 1 + z
  SC_TEND,       // dense_times[-1] = t[-1]/(1+z) + 100                       supernova_models.py:965
  SC_TAU2,       // tau_diff^2 [day^2]                                          interaction_processes.py:39
  SC_C1,         // (1024/ln2) / tau_diff^2
  SC_TRAP,       // trap_coeff [day^2]                                          interaction_processes.py:40
  SC_A0,         // log10(tau_diff / dense_times[-1]) - 3                       interaction_processes.py:50
  SC_ENG_A,      // nickel: f_nickel*mej ; magnetar: 2 E_rot / t_p
  SC_ENG_B,      // magnetar: 2 * 86400 / t_p  (1 + SC_ENG_B * t_days = 1 + 2 t / t_p)
  SC_RCV,        // RADIUS_CONSTANT * vej                                       photosphere.py:91
  SC_TFLOOR,     // temperature_floor
  SC_INV_REC,    // 1 / (STEF_CONSTANT * temperature_floor^4)                   photosphere.py:94
  SC_INV_DEN,    // 1 / (dl^2 c^2)                                              sed.py:219
  SC_DL,         // luminosity distance [cm]
  SC_SIGMA,      // sampled sigma (RB_SIGMA_SAMPLED / RB_SIGMA_QUADRATURE)
  SC_INV_TAU2,   // 1 / tau_diff^2
  SC_T0,         // explosion epoch t0 (phase_models.t0_base_model) or 0
  SC_SKIP,       // tau_diff^2 * X: abscissae with te^2 (1 - xm^2) above this are dropped (see sn_prep_kernel)
  kScalars       // = 17
};

struct SnArgs {
  rb_sn_config cfg;
  int64_t N;
  int E;
  const double* params;
  const double* dl_cm;
  const double* sigma_param;
  double* out_model;
  double* out_logl;
  const double* epoch_tab;  // [kEpochCols][E] built by epoch_table_kernel
  double* scalars;          // [N][kScalars]   built by sn_prep_kernel
  int epochs_in_smem;       // copy epoch_tab into shared memory (fits) or read it from global/L2
  int want_logl;
  // grid mode (magnitude branch, supernova_models.py:1008-1028): the "epochs" are the model's own time grid
  // t_temp (source frame: (t_temp*(1+z))/(1+z)); instead of flux densities the kernel writes T, R, L_bol and the
  // observer-frame phase t_temp*(1+z) per grid point for phot_kernel
  int64_t param_stride;     // distance between SoA rows of `params` (= total batch size)
  int grid_mode;
  double* grid_out;         // [N][4][E]
  double* opz_out;          // [N]
  double* dl_out;           // [N]
};

// ---- 1. per-launch epoch table ----------------------------------------------------------------------
// For RB_SIGMA_FIXED, 1/sigma and log(2 pi sigma^2)/2 (likelihoods.py:231) are computed once here instead
// of once per sample x epoch; other modes keep sigma_i itself in the EP_ISIG column.
__global__ void epoch_table_kernel(int E, int sigma_mode, const double* __restrict__ t_obs,
                                   const double* __restrict__ freq_obs, const double* __restrict__ y,
                                   const double* __restrict__ sigma, rb_upper_limits ul,
                                   double* __restrict__ tab) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double kind = 0.0;
  tab[EP_T * E + e] = t_obs[e];
  tab[EP_NU * E + e] = freq_obs ? freq_obs[e] : 0.0;
  tab[EP_Y * E + e] = y ? y[e] : 0.0;
  double isig = 0.0, logt = 0.0;
  if (y && sigma) {
    const double sg = sigma[e];
    if (sigma_mode == RB_SIGMA_FIXED) {
      isig = 1.0 / sg;
      logt = log(2 * kPi * (sg * sg)) / 2;
    } else {
      isig = sg;  // sigma_i, combined with the sampled sigma per sample
    }
  }
  if (y && ul.sigma_level != nullptr && ul.sigma_level[e] > 0.0) {
    // likelihoods.py:392-398: sigma_measurement = y / N (flux-like) or (2.5 / ln 10) / N (magnitudes)
    const double level = ul.sigma_level[e];
    kind = ul.magnitude_mode ? 2.0 : 1.0;
    isig = ul.magnitude_mode ? level / (2.5 / log(10.0)) : level / y[e];
    logt = 0.0;
  }
  tab[EP_ISIG * E + e] = isig;
  tab[EP_LOGT * E + e] = logt;
  tab[EP_KIND * E + e] = kind;
}

// ---- 2. per-sample scalars --------------------------------------------------------------------------
__global__ void sn_prep_kernel(const SnArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int64_t N = a.param_stride;
  const double* P = a.params + n;
  const int sed = a.cfg.sed;
  const double z = (sed == RB_SED_BOLOMETRIC) ? 0.0 : P[RB_SN_REDSHIFT * N];
  const double opz = 1.0 + z;
  double dl = 1.0;
  if (sed != RB_SED_BOLOMETRIC) {
    if (a.dl_cm != nullptr) {
      dl = a.dl_cm[n];
    } else {
      // cosmology.luminosity_distance(redshift) (supernova_models.py:997): 32-point Gauss-Legendre on [0, z]
      double s = RB_GL32_W[lane] * inv_efunc(a.cfg.cosmology, 0.5 * z * (RB_GL32_X[lane] + 1.0));
      s = warp_sum(s);
      dl = opz * a.cfg.cosmology.hubble_distance_cm * (0.5 * z * s);
    }
  }
  if (lane != 0) return;
  const double mej = P[RB_SN_MEJ * N], vej = P[RB_SN_VEJ * N];
  double t_last = a.epoch_tab[EP_T * a.E + a.E - 1];
  double t0 = 0.0;
  if (a.cfg.t0 != nullptr && !a.grid_mode) {
    // phase_models.py:33-48: the base model only sees the epochs with t - t0 >= 0; its time[-1] is the last of them
    t0 = a.cfg.t0[n];
    int e = a.E - 1;
    while (e >= 0 && !(a.epoch_tab[EP_T * a.E + e] - t0 >= 0.0)) --e;
    t_last = (e >= 0) ? a.epoch_tab[EP_T * a.E + e] - t0 : 0.0;
  }
  const double t_end = (a.grid_mode ? (t_last * opz) / opz : t_last / opz) + 100.0;
  const double tau = sqrt(kDiffusionConst * P[RB_SN_KAPPA * N] * mej / vej) / kDay;
  const double tau2 = tau * tau;
  if (a.grid_mode) { a.opz_out[n] = opz; a.dl_out[n] = dl; }
  double* S = a.scalars + n * kScalars;
  S[SC_OPZ] = opz;
  S[SC_TEND] = t_end;
  S[SC_TAU2] = tau2;
  S[SC_C1] = kTabOverLn2 / tau2;
  S[SC_INV_TAU2] = 1.0 / tau2;
  S[SC_TRAP] = (kTrappingConst * P[RB_SN_KAPPA_GAMMA * N] * mej / (vej * vej)) / (kDay * kDay);
  S[SC_A0] = log10(tau / t_end) - 3.0;
  if (a.cfg.engine == RB_ENGINE_NICKEL) {
    S[SC_ENG_A] = P[RB_SN_F_NICKEL * N] * mej;                                 // supernova_models.py:577
    S[SC_ENG_B] = 0.0;
  } else if (a.cfg.engine >= RB_ENGINE_MAGNETAR_ONLY) {
    S[SC_ENG_A] = 0.0;                 // these engines read their parameter rows inside the kernel
    S[SC_ENG_B] = 0.0;
  } else {
    // magnetar_models.py:181-184 with time in seconds (supernova_models.py:1466)
    const double p0 = P[RB_SN_P0 * N], bp = P[RB_SN_BP * N];
    const double m15 = pow(P[RB_SN_MASS_NS * N] / 1.4, 1.5);
    const double sn = sin(P[RB_SN_THETA_PB * N]);
    const double erot = 2.6e52 * m15 * (1.0 / (p0 * p0));
    const double tp = 1.3e5 * (1.0 / (bp * bp)) * (p0 * p0) * m15 * (1.0 / (sn * sn));
    S[SC_ENG_A] = 2.0 * erot / tp;
    S[SC_ENG_B] = 2.0 * kDay / tp;
  }
  // Abscissae whose weight exp(-X) cannot matter.  Rigorous bound for the engines that decrease monotonically: the
  // trapezoid sum is >= L(te) * w_last with w_last = lsp[0] = 10^a0 (the node at xm = 1, exponent 0), the dropped
  // terms sum to <= L(0) * exp(-X) * sum_i xm_i dw_i <= L(0) exp(-X), so with
  //   X = ln(L(0)/L(t_end)) - a0 ln 10 + ln(1e13)
  // they change the result by < 1e-13 relative (gate: 1e-9).  Other engines keep exp(-691) ~ 1e-300.
  double xskip = -(kExpSkip - 1.0);
  {
    double lnr = -1.0;
    const double lnr_ni = log((6.45e43 + 1.45e43) / (6.45e43 * exp(-t_end / 8.8) + 1.45e43 * exp(-t_end / 111.3)));
    const double lnr_mag = 2.0 * log1p(S[SC_ENG_B] * t_end);
    if (a.cfg.engine == RB_ENGINE_NICKEL) lnr = lnr_ni;
    else if (a.cfg.engine == RB_ENGINE_MAGNETAR) lnr = lnr_mag;
    else if (a.cfg.engine == RB_ENGINE_MAGNETAR_NICKEL) lnr = fmax(lnr_ni, lnr_mag);   // (A0+B0)/(A1+B1) <= max ratio
    const double x = lnr - 2.302585092994046 * S[SC_A0] + 29.933606208922594;
    if (lnr >= 0.0 && x < xskip) xskip = x;     // NaN / inf engines (theta_pb = 0) keep the default
  }
  S[SC_SKIP] = xskip * tau2;
  S[SC_RCV] = kRadiusConst * vej;
  double tfloor = 0.0, inv_rec = 0.0;
  if (sed != RB_SED_BOLOMETRIC) {
    tfloor = P[RB_SN_TEMPERATURE_FLOOR * N];
    const double tf2 = tfloor * tfloor;
    inv_rec = 1.0 / (kStefConst * (tf2 * tf2));
  }
  S[SC_TFLOOR] = tfloor;
  S[SC_INV_REC] = inv_rec;
  S[SC_INV_DEN] = 1.0 / ((dl * dl) * (kC * kC));
  S[SC_DL] = dl;
  S[SC_SIGMA] = (a.sigma_param != nullptr && a.want_logl) ? a.sigma_param[n] : 0.0;
  S[SC_T0] = t0;
}

// ---- CutoffBlackbody --------------------------------------------------------------------------------
// regularised incomplete gammas for the normalisation (sed.py:386-421).  Integer orders are closed forms evaluated in
// Horner form with reciprocal constants (no FP64 divisions): Q(s, x) = e^-x sum_{k<s} x^k / k!.
__device__ __forceinline__ double gammaincc_int(int s, double x, double emx) {
  double sum;
  switch (s) {
    case 1: sum = 1.0; break;
    case 2: sum = 1.0 + x; break;
    case 3: sum = fma(x, fma(x, 0.5, 1.0), 1.0); break;
    case 4: sum = fma(x, fma(x, fma(x, 1.0 / 6.0, 0.5), 1.0), 1.0); break;
    default: {
      double term = 1.0;
      sum = 1.0;
      for (int k = 1; k < s; ++k) {
        term *= x / k;
        sum += term;
      }
    }
  }
  return emx * sum;
}

// Q(s, x) for real s > 0 (non-integer absorption_index): series for P below x < s + 1, modified Lentz continued
// fraction for Q above (both converge to ~1e-15; scipy.special.gammaincc is accurate to the same level)
__device__ double gammaincc_real(double s, double x) {
  if (!(x > 0.0)) return 1.0;
  const double lead = exp(s * log(x) - x - lgamma(s));   // x^s e^-x / Gamma(s)
  if (x < s + 1.0) {
    double ap = s, del = 1.0 / s, sum = del;
    for (int n = 0; n < 500; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return 1.0 - sum * lead;
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - s, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 500; ++i) {
    const double an = -(double)i * ((double)i - s);
    b += 2.0;
    d = an * d + b;
    if (fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-16) break;
  }
  return lead * h;
}

// 24 / (k + 4)!, k = 0 .. 19: P(4, x) = e^-x x^4/24 sum_k x^k 24/(k+4)!  (no cancellation; used below x = 1.5, where the
// truncation after 20 terms is < 1e-17 relative)
__device__ const double kP4Series[20] = {
    1.0, 0.2, 0.03333333333333333, 0.004761904761904762, 0.0005952380952380953, 6.613756613756614e-05,
    6.613756613756614e-06, 6.012506012506013e-07, 5.010421677088344e-08, 3.854170520837188e-09,
    2.752978943455134e-10, 1.8353192956367562e-11, 1.1470745597729726e-12, 6.747497410429251e-14,
    3.748609672460695e-15, 1.9729524591898395e-16, 9.864762295949198e-18, 4.697505855213904e-19,
    2.135229934188138e-20, 9.283608409513644e-22};

__device__ __forceinline__ double gammainc4(double x, double emx) {
  // P(4, x); series below 1.5 (no cancellation), 1 - Q(4, x) above
  if (x < 0.25) {
    double sum = kP4Series[11];
#pragma unroll
    for (int k = 10; k >= 0; --k) sum = fma(sum, x, kP4Series[k]);
    const double x2 = x * x;
    return emx * (x2 * x2) * (1.0 / 24.0) * sum;
  }
  return 1.0 - gammaincc_int(4, x, emx);
}

// 1 / i^S, i = 1 .. 10 (row S - 1)
__device__ const double kInvPow[4][10] = {
    {1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10},
    {1.0, 1.0 / 4, 1.0 / 9, 1.0 / 16, 1.0 / 25, 1.0 / 36, 1.0 / 49, 1.0 / 64, 1.0 / 81, 1.0 / 100},
    {1.0, 1.0 / 8, 1.0 / 27, 1.0 / 64, 1.0 / 125, 1.0 / 216, 1.0 / 343, 1.0 / 512, 1.0 / 729, 1.0 / 1000},
    {1.0, 1.0 / 16, 1.0 / 81, 1.0 / 256, 1.0 / 625, 1.0 / 1296, 1.0 / 2401, 1.0 / 4096, 1.0 / 6561, 1.0 / 10000}};

// a^b for the handful of exponents the default configurations use, pow otherwise
__device__ __forceinline__ double pow_small(double a, double b) {
  if (b == 1.0) return a;
  if (b == 0.0) return 1.0;
  if (b == 2.0) return a * a;
  if (b == 3.0) return a * a * a;
  return pow(a, b);
}

// CutoffBlackbody normalisation for one time (sed.py:386-421): norms = L / (FLUX_CONST/A R^2 T) / ((above+below)/T)
// with above = (kT/hc)^4 Gamma(4) sum_n P(4, n x_c)/n^4, below = lam_c^-alpha (kT/hc)^s Gamma(s) sum_n Q(s, n x_c)/n^s,
// s = 4 - alpha.  Integer s uses the closed forms, real s the series / continued fraction.
__device__ __forceinline__ double cutoff_norm(double lbol, double rph, double tph, double lam_c, double alpha) {
  const double sr = 4.0 - alpha;
  const int s = (int)(sr + 0.5);
  const bool s_int = (sr == (double)s);
  const double kt_hc = tph / kXConst;
  const double x_c = 1.0 / (lam_c * kt_hc);
  const double e1 = exp(-x_c);
  double en = 1.0, above = 0.0, below = 0.0;
  // 1 / i^4 and 1 / i^s: compile-time constants after unrolling, a table for the run-time s (a rolled loop was measured
  // neutral for this SED and raised sn_kernel's spills around the call)
#pragma unroll
  for (int i = 1; i <= 10; ++i) {
    en *= e1;  // e^{-i x_c}
    const double di = (double)i;
    const double xi = x_c * di;
    above += gammainc4(xi, en) * (1.0 / (di * di * di * di));
    if (s_int && s <= 4) {
      below += gammaincc_int(s, xi, en) * kInvPow[s - 1][i - 1];
    } else if (s_int) {
      double ns = di;
      for (int q = 1; q < s; ++q) ns *= di;
      below += gammaincc_int(s, xi, en) / ns;
    } else {
      below += gammaincc_real(sr, xi) / pow(di, sr);
    }
  }
  const double kt2 = kt_hc * kt_hc;
  double kts, gam;
  if (s_int) {
    kts = kt_hc;
    gam = 1.0;  // kT_over_hc^s, Gamma(s) = (s-1)!
    for (int q = 1; q < s; ++q) { kts *= kt_hc; gam *= (double)q; }
  } else {
    kts = pow(kt_hc, sr);
    gam = tgamma(sr);
  }
  above = kt2 * kt2 * 6.0 * above;
  below = (1.0 / pow_small(lam_c, alpha)) * kts * gam * below;
  // norms = L / (FLUX_CONST/A R^2 T) / ((above + below) / T): the temperature cancels, one division instead of three
  return lbol / ((kFluxConst / kAngstrom * (rph * rph)) * (above + below));
}

// sed.CutoffBlackbody (sed.py:301-426) + _SED.flux_density (sed.py:239-251) for one (epoch, frequency);
// Returns mJy (before the (1+z) factor of supernova_models.py:1597).
__device__ __noinline__ double cutoff_blackbody_mjy(const rb_sn_config& cfg, double lbol, double rph, double tph,
                                                    double nu, double dl) {
  const double alpha = cfg.absorption_index;
  const double lam_c = cfg.cutoff_wavelength * kAngstrom;
  const double norm = cutoff_norm(lbol, rph, tph, lam_c, alpha);              // :386-421
  const double lam_A = 1.e10 * (kCSI / nu);                                   // utils.py:365-370
  const double lam = lam_A * kAngstrom;
  const double l2 = lam * lam;
  const double l5 = l2 * l2 * lam;
  const double em = expm1(kXConst / (lam * tph));
  // FLUX_CONST R^2 (lam / lam_c)^alpha / lam^5 / expm1(x) * norm / (4 pi dl^2) * lam_A / nu (sed.py:362-384, 239-251):
  // the factors that cannot overflow are gathered into one quotient, the Planck denominator (up to 1e308) divides last
  double num = (kFluxConst * kToMJy) * (rph * rph) * norm * lam_A;
  if (lam < lam_c) num *= pow_small(lam / lam_c, alpha);
  const double den = l5 * (4 * kPi * (dl * dl)) * nu;
  return (num / den) / em;
}

// sed.Line._set_sed (sed.py:859-909) for one (time, frequency): attenuate the base SED by the time-dependent line
// amplitude and add the Gaussian line emission scaled by L_bol.  `base_mjy` is the CutoffBlackbody flux density.
__device__ __forceinline__ double line_sed_mjy(const rb_sn_config& cfg, double base_mjy, double lbol, double te,
                                               double nu, double dl) {
  const double q = (te - cfg.line_time) / cfg.line_duration;
  const double amp = cfg.line_amplitude * exp(-0.5 * (q * q));                // :871-873
  double flux = (base_mjy * 1e-26) * (1 - amp);                               // :880-886 (mJy -> erg/s/cm^2/Hz)
  double amp_scaled = amp * lbol / (4 * kPi * (dl * dl));                     // :893
  amp_scaled /= (cfg.line_width * sqrt(2 * kPi));                             // :894
  const double lam_A = 1.e10 * (kCSI / nu);                                   // utils.py:365-370
  const double u = (lam_A - cfg.line_wavelength) / cfg.line_width;
  const double profile = exp(-0.5 * (u * u));                                 // :897
  const double lam_cm = lam_A * kAngstrom;
  flux += amp_scaled * profile * (lam_cm * lam_cm) / kC;                      // :906
  return flux * kToMJy;
}

// expm1 with the table-driven exp where there is no cancellation (|x| >= 0.5), libm otherwise
__device__ __forceinline__ double expm1_fast(double x, const double* __restrict__ tab) {
  if (fabs(x) >= 0.5 && x > -700.0 && x < 700.0) return exp_tab(x, tab) - 1.0;
  return expm1(x);
}

// ---- 3. the diffusion quadrature ----------------------------------------------------------------------
// Hot-loop exp(x), x in [-700, 0]: k = rint(x * 1024/ln2), r' = x * 1024/ln2 - k in [-0.5, 0.5],
// exp = 2^(k>>10) * tab[k&1023] * (1 + r'(A1 + r' A2)), A_i = (ln2/1024)^i / i!.
// Truncation (ln2/2048)^3/6 = 6.5e-12, argument-scaling error |x| * 2^-52 < 1.6e-13.
constexpr double kA1 = 0.0006769015435155716;
constexpr double kA2 = 2.2909784980688164e-07;

// GUARD: apply the NaN -> 0 replacement of interaction_processes.py:58 (only needed when the dense luminosity
//        contains non-finite values).
// FUSED: evaluate t_i^2 - t_e^2 with one FMA instead of the reference's rounded square + subtraction.  The two
//        differ by the rounding of t_i^2, i.e. by eps/2 * (t_e/tau_d)^2 relative in the integrand -- the
//        reference's own noise floor.  Selected per sample only where that is < 2e-11.
template <bool GUARD, bool FUSED, int UNROLL>
__device__ __forceinline__ double diffusion_sum(const double2* __restrict__ nodes, const double2* __restrict__ seg,
                                                const double* __restrict__ tab, int ms, int Dm1, double te,
                                                double te2, double te_is, double c1) {
  double acc = 0.0;
#pragma unroll UNROLL   // generic instantiations, A/B-measured (tools/bench_variants.py): 2: -1.1 %, 3: +0.1 %, 4: base, 6: +0.6 %,
                        // 8: +0.2 %; the spill-free Blackbody instantiation: 6: base, 10: +0.1 %, 12: +0.9 %, 16: +1.7 %, 25: -3.5 %
  for (int m = ms; m < kNodes; ++m) {
    const double2 nd = nodes[m];                                            // (xm, xm * dw)
    const double t = te * nd.x;                                             // interaction_processes.py:53
    // searchsorted(dense_times, t): v = 2^52 + ceil(t / step) with a round-up FMA
    const double v = __fma_ru(nd.x, te_is, 4503599627370496.0);
    const unsigned k = min((unsigned)__double2loint(v), (unsigned)Dm1);
    const double2 sg = seg[k];                                              // L(t) = a_k + b_k t on (x_{k-1}, x_k]
    const double lum = fma(sg.y, t, sg.x);                                  // scipy interp1d, :56
    const double d = FUSED ? fma(t, t, -te2) : __dsub_rn(__dmul_rn(t, t), te2);   // t_i^2 - t_e^2, :57
    const double vv = fma(d, c1, kExpMagic);
    const int kk = __double2loint(vv);
    const double r = fma(d, c1, -(vv - kExpMagic));
    const double q = fma(r, kA2, kA1);
    const double pp = r * q;
    const double tb = tab[kk & 1023];
    const double res = fma(tb, pp, tb);
    const double ex = __hiloint2double(__double2hiint(res) + ((kk >> 10) << 20), __double2loint(res));
    double g = lum * ex;
    if (GUARD) { if (isnan(g * t)) g = 0.0; }                              // :58 (NaN integrand -> 0)
    acc = fma(g, nd.y, acc);                                                // trapezoid, regrouped
  }
  return acc;
}

// FP32 variant of the quadrature (cfg.precision == RB_PREC_F32).  Differences to the FP64 loop:
//  * the exponent is evaluated as -(te/tau)^2 * (1 - xm^2) with (1 - xm^2) tabulated from FP64 -- algebraically
//    the reference's (t_i^2 - t_e^2)/tau^2 but without its catastrophic cancellation (fatal in FP32);
//  * exp() is MUFU.EX2 (__expf), relative error ~2e-7 + |arg| * 1e-7;
//  * luminosities are scaled by 1 / L(0) so that they fit the FP32 range; slope form of the interpolation.
__device__ __forceinline__ float diffusion_sum_f32(const float4* __restrict__ nodesf, const float2* __restrict__ segf,
                                                   int ms, int Dm1, float te, float te_is, float stepf, float cf) {
  float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 4
  for (int m = ms; m < kNodes; ++m) {
    const float4 nd = nodesf[m];                                   // (xm, xm * dw, 1 - xm^2, -)
    const float t = te * nd.x;
    const float v = __fmaf_ru(nd.x, te_is, 8388608.0f);            // 2^23 + ceil(t / step)
    const unsigned k = min((unsigned)(__float_as_int(v) & 0x7fffff), (unsigned)Dm1);
    const float2 sg = segf[k];                                     // (y_{k-1}, s_k) / L(0)
    const float dx = fmaf(-(v - 8388609.0f), stepf, t);            // t - x_{k-1}
    const float lum = fmaf(sg.y, dx, sg.x);
    const float ex = __expf(cf * nd.z);
    const float g = lum * ex;
    if (m & 1) acc1 = fmaf(g, nd.y, acc1); else acc0 = fmaf(g, nd.y, acc0);
  }
  return acc0 + acc1;
}

// HAS_T0: per-sample explosion epoch (phase_models.t0_base_model); a separate instantiation so that the plain path
// keeps its register allocation.
// ENGINE: 0 = nickel or magnetar chosen at run time (cfg.engine); RB_ENGINE_MAGNETAR_NICKEL = its own instantiation.
// (ptxas' register allocation of the hot loop is sensitive to the surrounding code: adding the third engine to the
// run-time branch cost 4 %, resolving nickel / magnetar at compile time cost the magnetar variant 30 % more spills.)
// ASPH: AsphericalDiffusion's viewing-angle factor on the diffused luminosity (interaction_processes.py:124-125).
// NEG: some epochs lie before the dense grid (cfg.negative_epochs); its own instantiation for the same reason (the select
// in front of the hot loop cost the plain path 2.7 % on the headline and 4.5 % on arnett_bolometric when always compiled in).
// NOCUT: the SED is bolometric or the plain Blackbody: the instantiation carries no CutoffBlackbody / Line code at all, so
// the out-of-line SED's register footprint and code size cannot touch the headline path (the rolled ten-term loop of
// cutoff_norm raised this kernel's spills around the call from 96 to 156 B when both lived in one instantiation).
// BOLO: bolometric output at the epochs (arnett_bolometric & co., BASELINE config 1): no photosphere / SED code at all;
// E = 100: 4.83e9 -> 5.08e9 unit/s against the generic instantiation (gpurun A/B, tools/bench_variants.py).
// BOLO = 1: blocks of at most 256 threads (E <= 256), three of which (six of 128) are resident per SM: the register cap
// follows from that (80 instead of 64: 5.08e9 -> 5.21e9 unit/s at E = 100); BOLO = 2: larger blocks, generic cap.
template <bool HAS_T0, int ENGINE, bool ASPH = false, bool NEG = false, bool NOCUT = false, int BOLO = 0>
__global__ void __launch_bounds__(BOLO == 1 ? 256 : 512, BOLO == 1 ? 3 : 2) sn_kernel(const SnArgs a) {
#ifndef RB_SN_UNROLL_NOCUT
#define RB_SN_UNROLL_NOCUT 16
#endif
  // unroll factor of the quadrature loop, A/B-measured per instantiation (tools/bench_variants.py): the spill-free
  // Blackbody instantiation gains from a deeper unroll, the generic one loses
#ifndef RB_SN_UNROLL_BOLO
#define RB_SN_UNROLL_BOLO 10
#endif
  constexpr int kUnroll = BOLO ? RB_SN_UNROLL_BOLO : (NOCUT ? RB_SN_UNROLL_NOCUT : 6);
  extern __shared__ double2 smem2[];
  const int D = a.cfg.dense_resolution;
  double2* seg = smem2;                                  // [D]      (a_k, b_k)
  double2* nodes = seg + D;                              // [kNodes] (xm, xm * dw)
  double* ytmp = reinterpret_cast<double*>(nodes + kNodes);  // [D]  engine luminosity on the dense grid
  double* lsp = ytmp + D;                                // [kNodesHalf]
  double* xms = lsp + kNodesHalf;                        // [kNodes] merged abscissae
  double* omx2 = xms + kNodes;                           // [kNodes] 1 - xm^2 (skip search)
  double* tab = omx2 + kNodes;                           // [1024]   2^(j/1024)
  double* scb = tab + kExpTabSize;                               // [2][kScalars] per-sample scalars, double-buffered
  double* scratch = scb + 2 * kScalars;                  // [16]     per-warp partial sums
  // FP32 copies only for RB_PREC_F32: the FP64 path does not pay their 10 KB per block (one more resident block per SM
  // at small E); [kEpochCols][E] follows when epochs_in_smem
  float4* nodesf = reinterpret_cast<float4*>(scratch + 16);          // [kNodes]
  float2* segf = reinterpret_cast<float2*>(nodesf + kNodes);         // [D]
  double* ep_s = (a.cfg.precision == RB_PREC_F32) ? reinterpret_cast<double*>(segf + D + (D & 1)) : scratch + 16;

  const int tid = threadIdx.x;
  const int E = a.E;
  const int sed = a.cfg.sed;
  const int sigma_mode = a.cfg.sigma_mode;
  const bool want_logl = a.want_logl != 0;
  const bool fp32 = a.cfg.precision == RB_PREC_F32;
  for (int i = tid; i < kExpTabSize; i += blockDim.x) tab[i] = RB_EXP2_1024[i];
  const double* ep = a.epoch_tab;
  if (a.epochs_in_smem) {
    for (int i = tid; i < kEpochCols * E; i += blockDim.x) ep_s[i] = a.epoch_tab[i];
    ep = ep_s;
  }
  if (tid < kScalars && blockIdx.x < a.N) scb[tid] = a.scalars[(int64_t)blockIdx.x * kScalars + tid];
  __syncthreads();

  int buf = 0;
  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x, buf ^= 1) {
    const double* sc = scb + buf * kScalars;
    // prefetch the next sample's scalar record; it lands in the other buffer before the closing barrier
    double pref = 0.0;
    const int64_t n_next = n + gridDim.x;
    if (tid < kScalars && n_next < a.N) pref = a.scalars[n_next * kScalars + tid];

    const double t_end = sc[SC_TEND];
    const double step = t_end / (double)(D - 1);                              // np.linspace step
    const double inv_step = 1.0 / step;

    // ---- engine on the dense grid --------------------------------------------------------------
    bool bad = false;
    {
      const double ea = sc[SC_ENG_A], eb = sc[SC_ENG_B];
      if (ENGINE == 0 && a.cfg.engine == RB_ENGINE_NICKEL) {
        for (int k = tid; k < D; k += blockDim.x) {                           // supernova_models.py:564-579
          const double x = (k == D - 1) ? t_end : (double)k * step;
          const double y = ea * (6.45e43 * exp_tab(fmax(-x / 8.8, -708.0), tab) +
                                 1.45e43 * exp_tab(fmax(-x / 111.3, -708.0), tab));
          ytmp[k] = y;
          bad |= !isfinite(y);
        }
      } else if (ENGINE == 0) {
        for (int k = tid; k < D; k += blockDim.x) {                           // magnetar_models.py:184
          const double x = (k == D - 1) ? t_end : (double)k * step;
          const double den = fma(x, eb, 1.0);
          const double y = ea / (den * den);
          ytmp[k] = y;
          bad |= !isfinite(y);
        }
      } else if (ENGINE == RB_ENGINE_MAGNETAR_ONLY) {                         // magnetar_models.py:143, :2391
        const double l0 = a.params[1 * a.param_stride + n], tsd = a.params[2 * a.param_stride + n];
        const double nn = a.params[3 * a.param_stride + n];
        const double pw = (1. + nn) / (1. - nn), tau_s = tsd * kDay;
        for (int k = tid; k < D; k += blockDim.x) {
          const double x = (k == D - 1) ? t_end : (double)k * step;
          const double y = l0 * pow(1. + (x * kDay) / tau_s, pw);
          ytmp[k] = y;
          bad |= !isfinite(y);
        }
      } else if (ENGINE == RB_ENGINE_EXP_POWERLAW) {                          // phenomenological_models.py:753
        const double a1 = a.params[1 * a.param_stride + n], al1 = a.params[2 * a.param_stride + n];
        const double al2 = a.params[3 * a.param_stride + n], tp = a.params[4 * a.param_stride + n];
        for (int k = tid; k < D; k += blockDim.x) {
          const double x = (k == D - 1) ? t_end : (double)k * step;
          const double y = a1 * pow(1 - exp(-x / tp), al1) * pow(x / tp, -al2);    // NaN at t = 0, as in the reference
          ytmp[k] = y;
          bad |= !isfinite(y);
        }
      } else if (ENGINE == RB_ENGINE_TDE_FALLBACK) {                          // tde_models.py:23-26
        const double l0 = a.params[1 * a.param_stride + n], tt = a.params[2 * a.param_stride + n];
        for (int k = tid; k < D; k += blockDim.x) {
          const double x = (k == D - 1) ? t_end : (double)k * step;
          const double y = (x - tt > 0.) ? l0 / pow(x * 86400, 5. / 3.) : l0 / pow(tt * 86400, 5. / 3.);
          ytmp[k] = y;
          bad |= !isfinite(y);
        }
      } else {                                                                // magnetar_nickel, :1669-1670
        // f_nickel * mej (:1669); read here rather than through a 17th scalar slot: the record layout feeds the hot
        // loop's register allocation
        const double ec = a.cfg.f_nickel[n] * a.params[RB_SN_MEJ * a.param_stride + n];
        for (int k = tid; k < D; k += blockDim.x) {
          const double x = (k == D - 1) ? t_end : (double)k * step;
          const double den = fma(x, eb, 1.0);
          const double y = ec * (6.45e43 * exp_tab(fmax(-x / 8.8, -708.0), tab) +
                                 1.45e43 * exp_tab(fmax(-x / 111.3, -708.0), tab)) + ea / (den * den);
          ytmp[k] = y;
          bad |= !isfinite(y);
        }
      }
    }
    // ---- quadrature abscissae: lsp = logspace(log10(tau/t_end) - 3, 0, 50) ----------------------
    // (done by the block's last threads, which have one fewer dense-grid iteration than the first ones)
    const int rt = (int)blockDim.x - 1 - tid;
    if (rt < kNodesHalf) {
      const double a0 = sc[SC_A0];
      const double da = (0.0 - a0) / (double)(kNodesHalf - 1);
      const double yv = (rt == kNodesHalf - 1) ? 0.0 : (double)rt * da + a0;
      // 10**y; a last-bit difference to numpy's pow is immaterial here
      lsp[rt] = (yv > -300.0) ? exp_tab(yv * 2.302585092994046, tab) : pow(10.0, yv);
    }
    const bool any_bad = __syncthreads_or(bad);
    // (t_end / tau_d)^2 * eps/2 < 2e-11: the fused t^2 - te^2 is indistinguishable from the reference's
    const bool fused_ok = t_end * t_end < 3.6e5 * sc[SC_TAU2];
    // interp1d segments (interaction_processes.py:44): L(t) = y_{k-1} + s_k (t - x_{k-1}) = a_k + s_k t
    for (int k = tid; k < D; k += blockDim.x) {
      double2 sg = make_double2(0.0, 0.0);
      if (k > 0) {
        const double y0 = ytmp[k - 1];
        const double sk = (ytmp[k] - y0) * inv_step;
        sg = make_double2(fma(-sk, (double)(k - 1) * step, y0), sk);
      }
      seg[k] = sg;
      if (fp32) {
        const double sc0 = 1.0 / ytmp[0];                                     // engines decrease monotonically
        segf[k] = (k > 0) ? make_float2((float)(ytmp[k - 1] * sc0), (float)((ytmp[k] - ytmp[k - 1]) * inv_step * sc0))
                          : make_float2(0.0f, 0.0f);
      }
    }
    // merge lsp (ascending) with 1 - lsp (non-increasing) in np.unique order: rank = own index + number of
    // elements of the other list that sort before it (binary search on the monotone other list)
    if (rt < kNodes) {
      int rank;
      double v;
      if (rt < kNodesHalf) {
        v = lsp[rt];
        int lo = 0, hi = kNodesHalf;  // first j with (1 - lsp[j]) < v; the set is a suffix
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((1.0 - lsp[mid]) < v) hi = mid; else lo = mid + 1;
        }
        rank = rt + (kNodesHalf - lo);
      } else {
        const int j0 = rt - kNodesHalf;
        v = 1.0 - lsp[j0];
        int lo = 0, hi = kNodesHalf;  // number of j with lsp[j] <= v; the set is a prefix
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (lsp[mid] <= v) lo = mid + 1; else hi = mid;
        }
        rank = (kNodesHalf - 1 - j0) + lo;
      }
      xms[rank] = v;
    }
    __syncthreads();
    // trapezoid regrouped by node: sum_i (t_i - t_{i-1})(f_i + f_{i-1}) = te^2 sum_i L_i e_i xm_i dw_i with
    // dw_i = xm_{i+1} - xm_{i-1} (dw_last = xm_last - xm_{last-1}); f_0 = 0 because xm_0 = 0.
    if (rt < kNodes) {
      const double x = xms[rt];
      const double dw = (rt == kNodes - 1) ? (x - xms[rt - 1]) : (xms[rt + 1] - xms[max(rt - 1, 0)]);
      nodes[rt] = make_double2(x, x * dw);
      omx2[rt] = (1.0 - x) * (1.0 + x);
      if (fp32) nodesf[rt] = make_float4((float)x, (float)(x * dw), (float)((1.0 - x) * (1.0 + x)), 0.0f);
    }
    __syncthreads();

    // ---- per-epoch work ------------------------------------------------------------------------
    double logl_part = 0.0;
    for (int e = tid; e < E; e += blockDim.x) {
      const double t_in = HAS_T0 ? ep[EP_T * E + e] - sc[SC_T0] : ep[EP_T * E + e];
      if (HAS_T0 && t_in < 0.0) {                                             // phase_models.py:54-58: before t0
        if (a.out_model != nullptr) a.out_model[n * (int64_t)E + e] = 0.0;
        if (want_logl)
          logl_part += logl_term(sigma_mode, ep[EP_KIND * E + e], ep[EP_Y * E + e], 0.0, ep[EP_ISIG * E + e],
                                 ep[EP_LOGT * E + e], sc[SC_SIGMA]);
        continue;
      }
      const double te_real = a.grid_mode ? (t_in * sc[SC_OPZ]) / sc[SC_OPZ] : t_in / sc[SC_OPZ];   // utils.py:382 (:1012-1014)
      // epochs before the dense grid are left out of uniq_times and np.searchsorted hands them the FIRST valid epoch's
      // luminosity (interaction_processes.py:47, :63); photosphere and SED below see the negative time itself
      const double te = (NEG && te_real < 0.0) ? a.cfg.t_first_valid / sc[SC_OPZ] : te_real;
      const double te2 = te * te;                                             // interaction_processes.py:55
      // first abscissa whose exponent can exceed kExpSkip: te^2 (1 - xm^2) <= -(kExpSkip - 1) tau^2
      int m0 = 0;
      {
        const double lim = sc[SC_SKIP];
        if (te2 > lim) {                 // 1 - xm^2 <= 1: nothing to skip otherwise
          int lo = 0, hi = kNodes;       // first m with te2 * omx2[m] <= lim (omx2 decreases with m)
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (te2 * omx2[mid] > lim) lo = mid + 1; else hi = mid;
          }
          m0 = lo;   // every abscissa kept has exponent >= -691 > -708 (the table exp's domain)
        }
      }
      const int ms = max(m0, 1);  // xm[0] = 1 - lsp[49] = 0 -> integrand 0 there
      // shrunk by 2^-50 so that ceil() never exceeds D-1 and t == x_k lands in (k-1, k]
      const double te_is = te * inv_step * (1.0 - 0x1p-50);
      const double c1 = sc[SC_C1];
      double acc;
      const double l0 = ytmp[0];
      if (fp32 && !any_bad && l0 > 0.0) {
        // FP32 skip: exponents below -87 underflow in FP32
        int mf = ms;
        const double limf = 87.0 * sc[SC_TAU2];
        if (te2 > limf) {
          int lo = 0, hi = kNodes;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (te2 * omx2[mid] > limf) lo = mid + 1; else hi = mid;
          }
          mf = max(lo, 1);
        }
        const float accf = diffusion_sum_f32(nodesf, segf, mf, D - 1, (float)te, (float)te_is, (float)step,
                                             (float)(-te2 * sc[SC_INV_TAU2]));
        acc = (double)accf * l0;
      } else if (any_bad) acc = diffusion_sum<true, false, kUnroll>(nodes, seg, tab, ms, D - 1, te, te2, te_is, c1);
      else if (fused_ok) acc = diffusion_sum<false, true, kUnroll>(nodes, seg, tab, ms, D - 1, te, te2, te_is, c1);
      else acc = diffusion_sum<false, false, kUnroll>(nodes, seg, tab, ms, D - 1, te, te2, te_is, c1);
      double lbol = (0.5 * (te2 * acc)) *
                    (-2.0 * expm1_fast(-sc[SC_TRAP] / te2, tab) * sc[SC_INV_TAU2]);             // :60-61
      if (ASPH) {                                                                               // :124-125
        const double u = te / sqrt(sc[SC_TAU2]) / 0.59;
        lbol *= (1 + 1.4 * (2 + u) / (1 + exp(u)) * (a.cfg.area_ratio - 1));
      }

      double model;
      if (BOLO || sed == RB_SED_BOLOMETRIC) {
        model = lbol;
      } else {
        // ---- photosphere.TemperatureFloor (photosphere.py:87-111) ----
        const double opz = sc[SC_OPZ];
        const double tfloor = sc[SC_TFLOOR];
        const double rr = sc[SC_RCV] * te_real;
        const double r2 = rr * rr;
        const double rec2 = lbol * sc[SC_INV_REC];
        const bool mask = r2 <= rec2;
        const double rph = sqrt(mask ? r2 : rec2);
        const double tph = mask ? sqrt(sqrt(lbol / (kStefConst * r2))) : tfloor;
        if (a.grid_mode) {
          double* go = a.grid_out + (size_t)n * 4 * E;
          go[0 * E + e] = tph;
          go[1 * E + e] = rph;
          go[2 * E + e] = lbol;
          go[3 * E + e] = t_in * opz;                                          // time_observer_frame (:1011)
          continue;
        }
        const double nu = ep[EP_NU * E + e] * opz;                            // utils.py:383
        if (sed == RB_SED_BLACKBODY) {
          // sed.py:199-222
          const double num = 2 * kPi * kH * (nu * nu * nu) * (rph * rph);
          const double em = expm1_fast((kH * nu) / (kKB * tph), tab);
          model = num * sc[SC_INV_DEN] / em * kToMJy * opz;                   // supernova_models.py:1007
        } else if (NOCUT) {
          model = NAN;                                                        // not reachable (sn_kernel_fn)
        } else {
          double mjy = cutoff_blackbody_mjy(a.cfg, lbol, rph, tph, nu, sc[SC_DL]);
          if (sed == RB_SED_CUTOFF_LINE) mjy = line_sed_mjy(a.cfg, mjy, lbol, te_real, nu, sc[SC_DL]);   // sed.py:859-909
          model = mjy * opz;                                                  // supernova_models.py:1597, 2278
        }
      }
      if (a.out_model != nullptr) a.out_model[n * (int64_t)E + e] = model;
      if (want_logl) {
        logl_part += logl_term(sigma_mode, ep[EP_KIND * E + e], ep[EP_Y * E + e], model, ep[EP_ISIG * E + e],
                               ep[EP_LOGT * E + e],
                               sc[SC_SIGMA]);                                 // likelihoods.py:231
      }
    }
    if (want_logl) {
      logl_part = warp_sum(logl_part);
      if ((tid & 31) == 0) scratch[tid >> 5] = logl_part;
    }
    if (tid < kScalars) scb[(buf ^ 1) * kScalars + tid] = pref;
    __syncthreads();  // all epochs done: shared tables may be rebuilt; warp partial sums + next scalars visible
    if (want_logl && tid == 0) {
      double total = 0.0;
      const int nw = (blockDim.x + 31) >> 5;
      for (int w = 0; w < nw; ++w) total += scratch[w];
      a.out_logl[n] = nan_to_num(total);
    }
  }
}

// ---- luminosity distance only ---------------------------------------------------------------------
__global__ void dl_kernel(const rb_cosmology c, int64_t N, const double* __restrict__ z_in,
                          double* __restrict__ out) {
  // one warp per redshift, one 32-point Gauss-Legendre node per lane (rel. error < 1e-14 for z <= 10,
  // 1e-12 at z = 20)
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= N) return;
  const double z = z_in[w];
  double s = RB_GL32_W[lane] * inv_efunc(c, 0.5 * z * (RB_GL32_X[lane] + 1.0));
  s = warp_sum(s);
  if (lane == 0) out[w] = (1.0 + z) * c.hubble_distance_cm * (0.5 * z * s);
}

// ---- stand-alone likelihood reduction -----------------------------------------------------------
__global__ void logl_kernel(int sigma_mode, int64_t N, int E, const double* __restrict__ model,
                            const double* __restrict__ y, const double* __restrict__ sigma,
                            const double* __restrict__ sigma_param, rb_upper_limits ul, double* __restrict__ out) {
  __shared__ double scratch[34];
  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    double part = 0.0;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const double sg = sigma ? sigma[e] : 0.0;
      const double sp = sigma_param ? sigma_param[n] : 0.0;
      if (ul.sigma_level != nullptr && ul.sigma_level[e] > 0.0) {
        const double level = ul.sigma_level[e];
        const double isig = ul.magnitude_mode ? level / (2.5 / log(10.0)) : level / y[e];
        part += logl_term(sigma_mode, ul.magnitude_mode ? 2.0 : 1.0, y[e], model[n * (int64_t)E + e], isig, 0.0, sp);
      } else if (sigma_mode == RB_SIGMA_FIXED) {
        const double rs = (y[e] - model[n * (int64_t)E + e]) / sg;
        part += -(rs * rs) / 2 - log(2 * kPi * (sg * sg)) / 2;
      } else {
        part += logl_term(sigma_mode, 0.0, y[e], model[n * (int64_t)E + e], sg, 0.0, sp);
      }
    }
    const double total = block_sum(part, scratch);
    if (threadIdx.x == 0) out[n] = nan_to_num(total);
  }
}

// ---- FP64 peak micro-benchmark ------------------------------------------------------------------
__global__ void dfma_kernel(int iters, double* out) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 12345.678) out[0] = s;
}

// MixtureGaussianLikelihood._mixture_gaussian_log_likelihood (likelihoods.py:542-577): per datum
// logaddexp(log alpha + log N(r | sigma), log(1 - alpha) + log N(r | sigma_out)); alpha and sigma_out per sample.
__global__ void mixture_logl_kernel(int64_t N, int E, const double* __restrict__ model, const double* __restrict__ y,
                                    const double* __restrict__ sigma, const double* __restrict__ alpha,
                                    const double* __restrict__ sigma_out, double* __restrict__ out) {
  __shared__ double scratch[34];
  const double half_log_2pi = 0.5 * log(2 * kPi);
  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    const double al = alpha[n], so = sigma_out[n];
    const double la = log(al), lb = log(1 - al), lso = log(so);
    double part = 0.0;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const double res = y[e] - model[n * (int64_t)E + e];
      const double sg = sigma[e];
      const double zi = res / sg, zo = res / so;
      const double ti = la + (-half_log_2pi - log(sg) - 0.5 * (zi * zi));
      const double to = lb + (-half_log_2pi - lso - 0.5 * (zo * zo));
      // np.logaddexp: max + log1p(exp(-|d|)); NaN propagates, equal infinities return themselves
      const double hi = fmax(ti, to), d = -fabs(ti - to);
      part += (ti == to) ? ti + 0.6931471805599453 : ((isnan(ti) || isnan(to)) ? NAN : hi + log1p(exp(d)));
    }
    const double total = block_sum(part, scratch);
    if (threadIdx.x == 0) out[n] = total;
    __syncthreads();
  }
}


}  // namespace rb
