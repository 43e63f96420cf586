// Magnetar-driven relativistic ejecta: _ejecta_dynamics_and_interaction
// (magnetar_driven_ejecta_models.py:14-151) + _process_flux_density (:240-272) + Gaussian likelihood.
//
// The reference integrates four coupled quantities (Lorentz factor, radius, comoving volume, internal energy) with
// explicit Euler over the 498 points of get_optimal_time_array(1e-4, 1e8, 500): a strictly sequential recurrence per
// parameter sample.  One THREAD per sample keeps the whole state in registers and marches the grid once; the epochs
// (sorted by time once per launch) are consumed on the fly as the march passes them -- interp1d of (T', R, D)
// between the previous and the current grid point, relativistic blackbody, likelihood term -- so no per-sample
// time series ever goes to memory.
// (included by capi.cu after sn_kernel.cu and kn_kernel.cu)

namespace rb {

constexpr double kProtonMass = 1.67262192369e-24;      // constants.proton_mass
constexpr double kRadiationConst = kSigmaSB * 4 / kC;  // constants.py:14

constexpr double kGravConst = 6.674299999999999e-08;    // constants.graviational_constant (astropy G, cgs)

// _evolving_gw_and_em_magnetar (magnetar_models.py:187-211): coefficients of the spin-down ODE and of Edot_d
struct EvolvingMagnetar {
  double c_em, c_gw, c_lum, omega;     // d omega/dt = -c_em omega^3 - c_gw omega^5;  L = c_lum omega^4
  __device__ __forceinline__ void init(double logbint, double logbext, double p0, double chi0, double radius_km,
                                       double logmoi) {
    const double bint = pow(10.0, logbint), bext = pow(10.0, logbext), radius = radius_km * kKm, moi = pow(10.0, logmoi);
    const double rb = bint / bext, b16 = bext / 1e16;
    const double eps = -3e-4 * (rb * rb) * (b16 * b16);                                          // :192
    const double sn = sin(chi0), sn2 = sn * sn;
    const double r6 = (radius * radius * radius) * (radius * radius * radius);
    const double c3 = kC * kC * kC, c5 = c3 * kC * kC;
    c_em = (bext * bext * r6 / (moi * c3)) * (1. + sn);                                          // :202
    c_gw = (2.0 * kGravConst * moi * (eps * eps) / (5.0 * c5)) * sn2 * (1.0 + 15.0 * sn2);       // :203
    c_lum = (bext * bext * r6 / (4 * c3)) * (1.0 + sn2);                                         // :205
    omega = 2.0 * kPi / p0;                                                                      // :193
  }
  __device__ __forceinline__ double luminosity() const { const double o2 = omega * omega; return c_lum * (o2 * o2); }
  __device__ __forceinline__ void step(double dt) {                                              // :200-203
    const double o2 = omega * omega, o3 = o2 * omega;
    omega = omega + dt * (-c_em * o3 - c_gw * (o3 * o2));
  }
};

// sorted epoch table columns
enum { MS_T = 0, MS_NU, MS_Y, MS_ISIG, MS_LOGT, MS_KIND, MS_ORIG, kMnEpochCols };

struct MnArgs {
  rb_mn_config cfg;
  int64_t N;
  int E;
  int G;                     // grid points
  const double* params;      // [RB_MN_NPARAM][param_stride]
  int64_t param_stride;
  const double* dl_cm;       // [N] (always given: the launcher integrates the cosmology first)
  const double* sigma_param;
  const double* tgrid;       // [G] source-frame seconds
  const double* sorted_ep;   // [kMnEpochCols][E], ascending MS_T (observer-frame days)
  const double* t0;          // [N] (offset to this chunk) or nullptr: per-sample merger epoch (t0_base_model)
  double* out_model;
  double* out_logl;
  int want_logl;
  int grid_mode;             // magnitude branch: write (T' D, R / D, L, phase [d]) per grid point for phot_kernel
  double* grid_out;          // [N][4][G]
  double* opz_out;
  double* dl_out;
};

// stable sort of the epoch table by time (rank = #earlier + #equal with a smaller index); one block, O(E^2/threads)
__global__ void mn_sort_epochs_kernel(int E, const double* __restrict__ ep, double* __restrict__ out) {
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const double te = ep[EP_T * E + e];
    int rank = 0;
    for (int f = 0; f < E; ++f) {
      const double tf = ep[EP_T * E + f];
      rank += (tf < te || (tf == te && f < e)) ? 1 : 0;
    }
    out[MS_T * E + rank] = te;
    out[MS_NU * E + rank] = ep[EP_NU * E + e];
    out[MS_Y * E + rank] = ep[EP_Y * E + e];
    out[MS_ISIG * E + rank] = ep[EP_ISIG * E + e];
    out[MS_LOGT * E + rank] = ep[EP_LOGT * E + e];
    out[MS_KIND * E + rank] = ep[EP_KIND * E + e];
    out[MS_ORIG * E + rank] = (double)e;
  }
}

// SN: the general_magnetar_driven_supernova outputs (RB_MN_OUT_SUPERNOVA / RB_MN_OUT_LBOL) as a separate instantiation, so
// that the mergernova outputs keep their register budget (96 vs 168 registers).
template <bool SN>
__global__ void __launch_bounds__(128, SN ? 3 : 5) mn_kernel(const MnArgs a) {
  extern __shared__ double mn_tg[];     // [G] time grid
  const int G = a.G, E = a.E;
  for (int i = threadIdx.x; i < G; i += blockDim.x) mn_tg[i] = a.tgrid[i];
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.N) return;
  const int64_t S = a.param_stride;
  const double* P = a.params + n;
  const double z = P[RB_MN_REDSHIFT * S];
  const double opz = 1.0 + z;
  const double dl = a.dl_cm[n];
  if (a.grid_mode) { a.opz_out[n] = opz; a.dl_out[n] = dl; }
  const double m = P[RB_MN_MEJ * S] * kMsun;                                   // :41
  const double kappa = P[RB_MN_KAPPA * S], n_ism = P[RB_MN_N_ISM * S];
  const double beta0 = P[RB_MN_BETA * S];
  double rad = P[RB_MN_EJECTA_RADIUS * S];
  double e_int = 0.5 * (beta0 * beta0) * m * (kC * kC);                        // :51
  double vol = (4.0 / 3.0) * kPi * (rad * rad * rad);                          // :52
  double gam = 1.0 / sqrt(1.0 - beta0 * beta0);                                // :53
  double d_gam = 0.0, d_rad = 0.0, d_vol = 0.0, d_e = 0.0;
  // magnetar engine constants
  double eng_a, eng_b, eng_c = 0.0;
  if (a.cfg.engine == RB_MN_BASIC_MAGNETAR) {                                  // magnetar_models.py:182-184
    const double p0 = P[RB_MN_P0 * S], bp = P[RB_MN_BP * S];
    const double m15 = pow(P[RB_MN_MASS_NS * S] / 1.4, 1.5);
    const double sn = sin(P[RB_MN_THETA_PB * S]);
    const double erot = 2.6e52 * m15 * (1.0 / (p0 * p0));
    const double tp = 1.3e5 * (1.0 / (bp * bp)) * (p0 * p0) * m15 * (1.0 / (sn * sn));
    eng_a = 2.0 * erot / tp;
    eng_b = tp;
  } else if (a.cfg.engine == RB_MN_MAGNETAR_ONLY) {                            // magnetar_models.py:143
    eng_a = P[RB_MN_L0 * S];
    eng_b = P[RB_MN_TAU_SD * S];
    const double nn = P[RB_MN_NN * S];
    eng_c = (1.0 + nn) / (1.0 - nn);
  } else {
    eng_a = eng_b = 0.0;
  }
  EvolvingMagnetar ev;
  ev.c_em = ev.c_gw = ev.c_lum = ev.omega = 0.0;
  if (a.cfg.engine == RB_MN_EVOLVING)
    ev.init(P[RB_MN_LOGBINT * S], P[RB_MN_LOGBEXT * S], P[RB_MN_EV_P0 * S], P[RB_MN_CHI0 * S], P[RB_MN_NS_RADIUS * S],
            P[RB_MN_LOGMOI * S]);
  const double row10 = P[RB_MN_THERMALISATION_EFFICIENCY * S];
  const double gprefac = 3.0 * row10 * m / (4.0 * kPi);                        // :103 (kappa_gamma variant), / vej^2 below
  const double lrad_coef = 4e49 * (m / 2e33 * 100.0);                          // :80
  const double nickel_mass = a.cfg.use_r_process ? 0.0 : P[RB_MN_F_NICKEL * S] * m / kMsun;   // :83
  const double pc_coef = (0.6 / (1.0 - a.cfg.ejecta_albedo)) * sqrt(a.cfg.pair_cascade_fraction / 0.1);   // :132
  const double sp = (a.sigma_param != nullptr && a.want_logl) ? a.sigma_param[n] : 0.0;
  const double inv_den = 1.0 / ((dl * dl) * (kC * kC));
  const double* ep = a.sorted_ep;
  const double t0g = mn_tg[0], t1g = mn_tg[G - 1];
  const int out_mode = a.cfg.output;
  const bool trapped = out_mode == RB_MN_OUT_TRAPPED_LUMINOSITY || out_mode == RB_MN_OUT_TRAPPED_FLUX;   // seconds (:477-573)
  const bool sn_mode = SN && out_mode == RB_MN_OUT_SUPERNOVA;       // interpolation axis in days (supernova_models.py:2573)
  const bool lbol_mode = SN && out_mode == RB_MN_OUT_LBOL;          // seconds, no redshift (:2484-2487)
  const double kcorr = pow(opz, a.cfg.photon_index - 2.0);          // :544
  double tfloor = 0.0, inv_rec = 0.0;
  rb_sn_config sed_cfg;                                             // only the cut-off parameters are read
  sed_cfg.cutoff_wavelength = a.cfg.cutoff_wavelength;
  sed_cfg.absorption_index = a.cfg.absorption_index;
  if (sn_mode) {
    tfloor = P[RB_MN_TEMPERATURE_FLOOR * S];
    const double tf2 = tfloor * tfloor;
    inv_rec = 1.0 / (kStefConst * (tf2 * tf2));
  }
  // epoch coordinate on the interpolation axis of the mode
  const double t_zero = a.t0 != nullptr ? a.t0[n] : 0.0;             // phase_models.py:33-45
  auto epoch_x = [&](int e) -> double {
    const double t_obs = ep[MS_T * E + e] - t_zero;
    if (trapped || sn_mode) return t_obs / opz;                     // :537 / supernova_models.py:2561
    if (lbol_mode) return t_obs * kDay;                             // :2484
    return (t_obs * kDay) / opz;                                    // :252-254
  };

  int ei = 0;
  double logl = 0.0;
  // epochs before the grid: interp1d raises in the reference -> NaN; epochs before t0 (t0_base_model): 0
  while (ei < E) {
    const double ts = epoch_x(ei);
    if (ts >= (sn_mode ? t0g / kDay : t0g)) break;
    if (!a.grid_mode) {
      const double fill = ts < 0.0 ? 0.0 : NAN;                      // ts < 0 only with a per-sample t0
      if (a.out_model != nullptr) a.out_model[n * (int64_t)E + (int)ep[MS_ORIG * E + ei]] = fill;
      if (ts < 0.0) {
        if (a.want_logl)
          logl += logl_term(a.cfg.sigma_mode, ep[MS_KIND * E + ei], ep[MS_Y * E + ei], 0.0, ep[MS_ISIG * E + ei],
                            ep[MS_LOGT * E + ei], sp);
      } else {
        logl = NAN;
      }
    }
    ++ei;
  }
  double pt = 0.0, ptemp = 0.0, prad = 0.0, pdop = 0.0, ptau = 0.0, plb = 0.0;
  for (int i = 0; i < G; ++i) {
    const double t = mn_tg[i];
    const double beta = sqrt(1.0 - 1.0 / (gam * gam));                         // :65
    const double dop = 1.0 / (gam * (1.0 - beta));                             // :66
    if (i > 0) {
      const double dt = t - mn_tg[i - 1];
      gam = gam + d_gam * dt;                                                  // :69-72
      rad = rad + d_rad * dt;
      vol = vol + d_vol * dt;
      e_int = e_int + d_e * dt;
    }
    const double swept = (4.0 / 3.0) * kPi * (rad * rad * rad) * n_ism * kProtonMass;            // :74
    const double pressure = e_int / (3.0 * vol);
    const double t_com = dop * t;
    const double dvdt_com = 4.0 * kPi * (rad * rad) * beta * kC;
    // :78 -- product and difference rounded separately as in the reference: at late times the difference is ~1e-9,
    // a fused multiply-subtract (no rounding of the ~0.5 product) would move it by up to 1.5e-8 relative
    const double denom = __dsub_rn(0.5, __dmul_rn(1.0 / 3.141592654, atan_cr((t_com - 1.3) / 0.11)));
    double l_rad;
    if (!SN || a.cfg.use_r_process) {
      l_rad = lrad_coef * pow(denom, 1.3);                                                        // :80
    } else {                                                                                      // :82-84
      l_rad = nickel_mass * (6.45e43 * exp(-t_com / (8.8 * 86400)) + 1.45e43 * exp(-t_com / (111.3 * 86400)));
    }
    const double tau = kappa * (m / vol) * (rad / gam);                                           // :85
    double l_emit, temp;
    if (tau <= 1.0) {
      l_emit = (e_int * kC) / (rad / gam);
      temp = sqrt(sqrt(e_int / (kRadiationConst * vol)));
    } else {
      l_emit = (e_int * kC) / (tau * rad / gam);
      temp = sqrt(sqrt(e_int / (kRadiationConst * vol * tau)));
    }
    double mag_lum;
    if (a.cfg.engine == RB_MN_BASIC_MAGNETAR) {
      const double den = 1.0 + 2.0 * t / eng_b;
      mag_lum = eng_a / (den * den);
    } else if (a.cfg.engine == RB_MN_MAGNETAR_ONLY) {
      mag_lum = eng_a * pow(1.0 + t / eng_b, eng_c);
    } else {
      mag_lum = ev.luminosity();                       // Edot_d at grid point i; omega advances to i + 1 below
      if (i + 1 < G) ev.step(mn_tg[i + 1] - t);
    }
    double eff = row10;
    if (a.cfg.use_gamma_ray_opacity) {
      const double vej = kC * sqrt(1.0 - 1.0 / (gam * gam));                                      // utils.py:1524
      eff = 1.0 - exp(-(gprefac / (vej * vej)) * (1.0 / (t * t)));                                // :103-104
    }
    const double dop2 = dop * dop;
    d_rad = (beta * kC) / (1.0 - beta);                                                           // :108
    const double d_swept = 4.0 * kPi * (rad * rad) * n_ism * kProtonMass * d_rad;
    const double dedt = eff * mag_lum + dop2 * l_rad - dop2 * l_emit;                             // :110-111
    const double de_com = eff * (1.0 / dop2) * mag_lum + l_rad - l_emit - pressure * dvdt_com;    // :112-113
    d_vol = dvdt_com * dop;
    d_e = de_com * dop;
    d_gam = (dedt - gam * dop * de_com - (gam * gam - 1.0) * (kC * kC) * d_swept) /
            (m * (kC * kC) + e_int + 2.0 * gam * swept * (kC * kC));                              // :116-118
    double lbol = l_emit * dop2;                                                                  // :93
    if (a.cfg.pair_cascade_switch) {                                                              // :131-135
      const double v0 = sqrt((1.0 / gam) * (1.0 / gam) + 1.0) * kC;
      const double tlife = pc_coef * sqrt(mag_lum / 1.0e45) * sqrt(v0 / (0.3 * kC)) * (1.0 / sqrt(t / 86400.0));
      lbol = lbol / (1.0 + tlife);
      temp = sqrt(sqrt(lbol / (4.0 * kPi * (rad * rad) * kSigmaSB)));
    }
    if (a.grid_mode) {
      double* go = a.grid_out + (size_t)n * 4 * G;
      if (SN && sn_mode) {               // supernova_models.py:2598-2600, 2622-2626: photosphere on the grid
        const double vej_km = kC * sqrt(1.0 - 1.0 / (gam * gam)) / kKm;       // utils.py:1524
        double t_ph, r_ph;
        temperature_floor_dev(lbol, t / kDay, kRadiusConst * vej_km, tfloor, inv_rec, t_ph, r_ph);
        go[0 * G + i] = t_ph;
        go[1 * G + i] = r_ph;
      } else {
        go[0 * G + i] = temp * dop;      // exp(h nu / (k T' D)): an ordinary blackbody of temperature T' D ...
        go[1 * G + i] = rad / dop;       // ... and radius R / D (:164-172)
      }
      go[2 * G + i] = lbol;
      go[3 * G + i] = (t * opz) / kDay;  // time_observer_frame / day_to_s (:236)
    } else {
      // quantities interpolated by the mode: (T', R, D, tau) of the ejecta, or (T_ph, R_ph, L) of the photosphere
      double q_temp = temp, q_rad = rad;
      const double xg = sn_mode ? t / kDay : t;                               // supernova_models.py:2573
      if (sn_mode) {                                                          // :2570-2572: TemperatureFloor on the grid
        const double vej_km = kC * sqrt(1.0 - 1.0 / (gam * gam)) / kKm;       // utils.py:1524
        temperature_floor_dev(lbol, xg, kRadiusConst * vej_km, tfloor, inv_rec, q_temp, q_rad);
      }
      if (i > 0) {
        // epochs in (x_{i-1}, x_i]: interp1d (:256-261) + the mode's emission law
        while (ei < E) {
          const double ts = epoch_x(ei);
          if (!(ts <= xg)) break;
          const double w = ts - pt, inv = 1.0 / (xg - pt);
          const double te = (q_temp - ptemp) * inv * w + ptemp;
          const double re = (q_rad - prad) * inv * w + prad;
          const double nu = ep[MS_NU * E + ei] * opz;
          double model;
          if (SN && lbol_mode) {
            model = (lbol - plb) * inv * w + plb;                             // :2481-2485
          } else if (SN && sn_mode) {
            const double lb = (lbol - plb) * inv * w + plb;
            if (a.cfg.sed == RB_SED_BLACKBODY) {
              const double num = 2 * kPi * kH * (nu * nu * nu) * (re * re);
              model = num * inv_den / expm1((kH * nu) / (kKB * te)) * kToMJy * opz;
            } else {
              model = cutoff_blackbody_mjy(sed_cfg, lb, re, te, nu, dl) * opz;     // :2577-2581
            }
          } else if (!SN) {
            const double de = (dop - pdop) * inv * w + pdop;
            const double frac = 1.0 / (exp((kH * nu) / (kKB * te * de)) - 1.0);   // the reference's exp(x) - 1 (:172)
            if (!trapped) {
              const double num = 2 * kPi * kH * (nu * nu * nu) * (re * re);
              model = num * (inv_den / (de * de)) * frac * opz * kToMJy * opz;                    // :262, :272
            } else {
              // _comoving_blackbody_to_luminosity (:177-196) + the leaking spin-down luminosity (:509-511)
              const double optical_depth = (tau - ptau) * inv * w + ptau;
              const double num = 8 * (kPi * kPi) * kH * ((nu * nu) * (nu * nu)) * (re * re);
              const double l_ej = num / ((kC * kC) * (de * de)) * frac;
              const double lsd = eng_a * pow(1.0 + ts / eng_b, eng_c);
              model = exp(-optical_depth) * lsd + l_ej;
              if (out_mode == RB_MN_OUT_TRAPPED_FLUX) model = model / (4 * kPi * (dl * dl) * kcorr);   // :545
            }
          } else {
            model = NAN;
          }
          if (a.out_model != nullptr) a.out_model[n * (int64_t)E + (int)ep[MS_ORIG * E + ei]] = model;
          if (a.want_logl)
            logl += logl_term(a.cfg.sigma_mode, ep[MS_KIND * E + ei], ep[MS_Y * E + ei], model, ep[MS_ISIG * E + ei],
                              ep[MS_LOGT * E + ei], sp);
          ++ei;
        }
      }
      pt = xg; ptemp = q_temp; prad = q_rad; pdop = dop; ptau = tau; plb = lbol;
      continue;
    }
    pt = t; ptemp = temp; prad = rad; pdop = dop; ptau = tau;
  }
  if (!a.grid_mode) {
    for (; ei < E; ++ei) {               // beyond the grid (t_src > 1e8 s or NaN)
      if (a.out_model != nullptr) a.out_model[n * (int64_t)E + (int)ep[MS_ORIG * E + ei]] = NAN;
      logl = NAN;
    }
    if (a.want_logl) a.out_logl[n] = nan_to_num(logl);
  }
  (void)t1g;
}

}  // namespace rb
