// Magnetar-driven Metzger shell model: _general_metzger_magnetar_driven_kilonova_model
// (magnetar_driven_ejecta_models.py:576-757, magnetar_heating = 'first_layer' or 'all_layers') behind
// metzger_magnetar_driven_kilonova_model / general_metzger_magnetar_driven(_thermalisation) (:760-905), flux density
// through _process_flux_density (:240-272, plain blackbody) and the Gaussian likelihood.
//
// One block per parameter sample.  The 200 mass shells are threads; every shell integrates its own explicit-Euler
// energy equation over the 298-point grid get_optimal_time_array(1e-4, 1e7, 300).  Unlike the plain Metzger model the
// shells are coupled: the bulk velocity follows the kinetic energy, which is fed by the thermal energy of shell 0.
// Instead of a block barrier per time step each thread carries a private replica of shell 0's state (identical
// arithmetic in every thread, hence bit-identical), so the march needs no synchronisation; per step only the shell sum
// of the radiated luminosity and the first-minimum argmin |tau - 1| are reduced with warp shuffles.
// (included by capi.cu after kn_kernel.cu, whose helpers -- atan_cr, bk_interp, Ye table -- it uses)

namespace rb {

constexpr int kMkThreads = 256;

struct MkArgs {
  rb_mk_config cfg;
  int64_t N;
  int E;
  const double* t0;          // [N] (offset to this chunk) or nullptr: per-sample merger epoch (t0_base_model)
  int G;
  const double* params;      // [RB_MK_NPARAM][param_stride]
  int64_t param_stride;
  const double* dl_cm;       // [N]
  const double* sigma_param;
  const double* tgrid;       // [G] source-frame seconds
  const double* epoch_tab;   // [kEpochCols][E]
  double* out_model;
  double* out_logl;
  int want_logl;
  int grid_mode;             // magnitude branch: write (T, R_photosphere, L, phase [d]) per grid point for phot_kernel
  double* grid_out;          // [N][4][G]
  double* opz_out;
  double* dl_out;
};

__global__ void __launch_bounds__(kMkThreads) mk_kernel(const MkArgs a) {
  extern __shared__ double mk_smem[];
  const int G = a.G, E = a.E;
  double* tg = mk_smem;                    // [G]
  double* w1 = tg + G;                     // [G] r-process heating rate per unit mass (:629-630)
  double* w2 = w1 + G;                     // [G] exp(-t / tau_neutron)
  double* lsd = w2 + G;                    // [G] magnetar luminosity
  double* v0g = lsd + G;                   // [G] bulk velocity per step
  double* Tg = v0g + G;                    // [G]
  double* Rg = Tg + G;                     // [G]
  double* lpart = Rg + G;                  // [kShellWarps][G]
  double* tpart = lpart + kShellWarps * G; // [kShellWarps][G]
  double* vfac = tpart + kShellWarps * G;  // [kShells] (m / mej)^(-1/beta)
  double* marr = vfac + kShells;           // [kShells] normalised shell masses [Msun]
  double* red = marr + kShells;            // [40]
  int* ipart = reinterpret_cast<int*>(red + 40);   // [kShellWarps][G]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < G; i += blockDim.x) tg[i] = a.tgrid[i];
  const bool npre = a.cfg.neutron_precursor_switch != 0;
  const double pc_coef = (0.6 / (1.0 - a.cfg.ejecta_albedo)) * sqrt(a.cfg.pair_cascade_fraction / 0.1);

  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();
    const int64_t S = a.param_stride;
    const double* P = a.params + n;
    const double z = P[RB_MK_REDSHIFT * S], opz = 1.0 + z;
    const double mej = P[RB_MK_MEJ * S], vej = P[RB_MK_VEJ * S], beta = P[RB_MK_BETA * S], kappa = P[RB_MK_KAPPA_R * S];
    const double row9 = P[RB_MK_THERMALISATION_EFFICIENCY * S];
    const double dl = a.dl_cm[n];
    const double m0 = mej * kMsun;
    // magnetar engine
    double eng_a, eng_b, eng_c = 0.0;
    if (a.cfg.engine == RB_MN_BASIC_MAGNETAR) {                                 // magnetar_models.py:182-184
      const double p0 = P[RB_MK_P0 * S], bp = P[RB_MK_BP * S];
      const double m15 = pow(P[RB_MK_MASS_NS * S] / 1.4, 1.5);
      const double sn = sin(P[RB_MK_THETA_PB * S]);
      const double erot = 2.6e52 * m15 * (1.0 / (p0 * p0));
      const double tp = 1.3e5 * (1.0 / (bp * bp)) * (p0 * p0) * m15 * (1.0 / (sn * sn));
      eng_a = 2.0 * erot / tp;
      eng_b = tp;
    } else if (a.cfg.engine == RB_MN_MAGNETAR_ONLY) {
      eng_a = P[RB_MK_L0 * S];
      eng_b = P[RB_MK_TAU_SD * S];
      const double nn = P[RB_MK_NN * S];
      eng_c = (1.0 + nn) / (1.0 - nn);
    } else {
      eng_a = eng_b = 0.0;
      if (tid == 0) {                                   // sequential spin-down ODE over the grid (magnetar_models.py:200-205)
        EvolvingMagnetar ev;
        ev.init(P[RB_MK_LOGBINT * S], P[RB_MK_LOGBEXT * S], P[RB_MK_EV_P0 * S], P[RB_MK_CHI0 * S],
                P[RB_MK_NS_RADIUS * S], P[RB_MK_LOGMOI * S]);
        for (int k = 0; k < G; ++k) {
          lsd[k] = ev.luminosity();
          if (k + 1 < G) ev.step(tg[k + 1] - tg[k]);
        }
      }
    }
    // ---- grid quantities (:596-631) -------------------------------------------------------------------------------
    {
      const double av = bk_interp(kBkA, mej, vej), bv = bk_interp(kBkB, mej, vej), dv = bk_interp(kBkD, mej, vej);
      for (int k = tid; k < G; k += blockDim.x) {
        const double t = tg[k];
        const double tdays = t / kDay;
        const double xx = 2.0 * bv * pow(tdays, dv);
        const double e_th = 0.36 * (exp(-av * tdays) + log1p(xx) / xx);
        if (t > 1.3) w1[k] = 2.1e10 * e_th * pow(t / (3600. * 24.), -1.3);
        else w1[k] = 4.0e18 * pow(0.5 - (1. / kPi) * atan_cr((t - 1.3) / 0.11), 1.3) * e_th;
        w2[k] = exp(-t / 900.0);
        Tg[k] = 1.0 / t;               // shared by all shell threads during the march (Tg is filled after it)
        if (a.cfg.engine == RB_MN_BASIC_MAGNETAR) {
          const double den = 1.0 + 2.0 * t / eng_b;
          lsd[k] = eng_a / (den * den);
        } else if (a.cfg.engine == RB_MN_MAGNETAR_ONLY) {
          lsd[k] = eng_a * pow(1.0 + t / eng_b, eng_c);
        }
      }
    }
    // ---- shell masses, normalised to mej (:614-619) --------------------------------------------------------------------
    const double vstep = (a.cfg.vmax - vej) / (double)(kShells - 1);           // np.linspace(vmin, vmax, 200)
    double raw = 0.0;
    if (tid < kShells) {
      const double vel = (tid == kShells - 1) ? a.cfg.vmax : (double)tid * vstep + vej;
      raw = mej * pow(vel / vej, -beta);
    }
    const double total = block_sum(raw, red);
    if (tid < kShells) {
      const double m = raw * (mej / total);
      marr[tid] = m;
      vfac[tid] = pow(m / mej, -1 / beta);                                      // :671
    }
    __syncthreads();
    // Ye from kappa (utils.electron_fraction_from_kappa)
    double ye;
    {
      int i = 1;
      while (i < 7 && kYeKappa[i] < kappa) ++i;
      const double slope = (kYeValue[i] - kYeValue[i - 1]) / (kYeKappa[i] - kYeKappa[i - 1]);
      ye = slope * (kappa - kYeKappa[i - 1]) + kYeValue[i - 1];
    }
    // ---- the march: one thread per shell + a private replica of shell 0 ---------------------------------------------
    if (tid < kShellWarps * 32) {
      const int s = min(tid, kShells - 1);
      const bool ode = tid < kShells - 1;                                       // the [:-1] slices
      const double ms = marr[s], ms0 = marr[0];
      const double dm = fabs(marr[min(s + 1, kShells - 1)] - ms), dm0 = fabs(marr[1] - ms0);   // :658
      const double vf = vfac[s], vf0 = vfac[0];
      double xn0 = 0.0, xr = 1.0, xn00 = 0.0, xr0 = 1.0;
      if (npre) {                                                               // :646-648
        xn0 = (2.0 / kPi) * (1.0 - ye) * atan(1e-4 * kMsun / ms / kMsun);
        xr = 1.0 - xn0;
        xn00 = (2.0 / kPi) * (1.0 - ye) * atan(1e-4 * kMsun / ms0 / kMsun);
        xr0 = 1.0 - xn00;
      }
      const double vel = (s == kShells - 1) ? a.cfg.vmax : (double)s * vstep + vej;
      double en = 0.5 * ms * ((vel * kC) * (vel * kC));                        // :660 (initial v_m = vel c)
      double en0 = 0.5 * ms0 * ((vej * kC) * (vej * kC));
      double ke = 0.5 * m0 * ((vej * kC) * (vej * kC));                        // :610
      const double gpref = 3 * row9 * mej / (4 * kPi * (vej * vej));           // :675 (kappa_gamma variant)
      // Divisions per step: 1 / t is tabulated per grid point, 1 / v_m = (1 / v0) (1 / vf) needs one division for both
      // shells of the thread; with the two L = E / (t_d + t v / c) that leaves three instead of eight (:667-712)
      const double ivf = 1.0 / vf, ivf0 = 1.0 / vf0, inv_c = 1.0 / kC;
      const double c_td = (ms * kMsun * 3) / (4 * kPi * kC * beta), c_td0 = (ms0 * kMsun * 3) / (4 * kPi * kC * beta);
      const double c_tau = (ms * kMsun) / (4 * kPi);
      for (int ii = 0; ii < G - 1; ++ii) {
        const double t = tg[ii], dt = tg[ii + 1] - t, it = Tg[ii];
        ke = ke + (en0 * it) * dt;                                             // :667
        const double v0 = sqrt(2 * ke / m0);
        const double iv0 = 1.0 / v0;
        if (tid == 0) v0g[ii] = v0;
        double vm = v0 * vf, vm0 = v0 * vf0;                                   // :671-672
        double ivm = iv0 * ivf, ivm0 = iv0 * ivf0;
        if (vm > 3e10) { vm = kC; ivm = inv_c; }
        if (vm0 > 3e10) { vm0 = kC; ivm0 = inv_c; }
        double eff = row9;
        if (a.cfg.use_gamma_ray_opacity) eff = 1 - exp(-gpref * (1.0 / (t * t)));   // :676 (1 - exp, as the reference)
        const double qdot = eff * lsd[ii];
        double kap = kappa, kap0 = kappa, edot = w1[ii], edot0 = w1[ii];
        if (npre) {
          const double xn = xn0 * w2[ii], xnb = xn00 * w2[ii];
          edot = 3.2e14 * xn + edot;                                           // :651-652
          edot0 = 3.2e14 * xnb + edot0;
          kap = 0.4 * (1.0 - xn - xr) + kappa * xr;                            // :653-655
          kap0 = 0.4 * (1.0 - xnb - xr0) + kappa * xr0;
        }
        // shell 0 replica (:697-707)
        {
          const double td0 = (kap0 * c_td0) * (ivm0 * it);
          const double lum0 = en0 / (td0 + t * (vm0 * inv_c));
          en0 = (qdot + edot0 * dm0 * kMsun - (en0 * it) - lum0) * dt + en0;
        }
        // this thread's shell
        const double td = (kap * c_td) * (ivm * it);
        const double lum = en / (td + t * (vm * inv_c));
        if (tid == 0) en = en0;                                                // same arithmetic; keep them identical
        else if (a.cfg.heat_all_layers) en = (qdot + edot * dm * kMsun - (en * it) - lum) * dt + en;   // :704
        else en = (edot * dm * kMsun - (en * it) - lum) * dt + en;             // :710
        const double itv = it * ivm;
        const double tau = (kap * c_tau) * (itv * itv);                         // :712
        double lo = ode ? lum : 0.0;
        double dist = ode ? fabs(tau - 1.0) : INFINITY;                         // shell 199 copies 198: never first
        if (isnan(dist)) dist = -INFINITY;                                      // np.argmin returns the first NaN
        int idx = s;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lo += __shfl_xor_sync(0xffffffffu, lo, o);
        warp_argmin_first(dist, idx);
        if (lane == 0) {
          lpart[warp * G + ii] = lo;
          tpart[warp * G + ii] = dist;
          ipart[warp * G + ii] = idx;
        }
      }
    }
    __syncthreads();
    // ---- bolometric luminosity, photosphere, temperature (:717-729) -----------------------------------------------------
    {
      const double v0_last = v0g[G - 2];
      for (int k = tid; k < G; k += blockDim.x) {
        double lb = 0.0, rad = 0.0;
        if (k < G - 1) {
          double best = tpart[k];
          int bi = ipart[k];
          lb = lpart[k];
          for (int w = 1; w < kShellWarps; ++w) {
            lb += lpart[w * G + k];
            const double d = tpart[w * G + k];
            if (d < best) { best = d; bi = ipart[w * G + k]; }
          }
          double vm = v0g[k] * vfac[bi];
          if (vm > 3e10) vm = kC;
          rad = vm * tg[k];                                                     // :715-716
        }
        if (a.cfg.pair_cascade_switch) {                                        // :725-728 (v0 = its final value)
          const double tlife = pc_coef * sqrt(lsd[k] / 1.0e45) * sqrt(v0_last / (0.3 * kC)) * (1.0 / sqrt(tg[k] / kDay));
          lb = lb / (1.0 + tlife);
        }
        Tg[k] = sqrt(sqrt(lb / (4.0 * kPi * (rad * rad) * kSigmaSB)));         // :729
        Rg[k] = rad;
      }
    }
    __syncthreads();
    if (a.grid_mode) {                       // _processing_other_formats (:198-238), use_relativistic_blackbody = False
      double* go = a.grid_out + (size_t)n * 4 * G;
      for (int k = tid; k < G; k += blockDim.x) {
        go[0 * G + k] = Tg[k];
        go[1 * G + k] = Rg[k];
        go[2 * G + k] = 0.0;
        go[3 * G + k] = (tg[k] * opz) / kDay;                                   // time_observer_frame / day_to_s (:236)
      }
      if (tid == 0) { a.opz_out[n] = opz; a.dl_out[n] = dl; }
      continue;
    }
    // ---- epochs: interp1d + blackbody (:263-272) + likelihood ------------------------------------------------------------
    const double inv_den = 1.0 / ((dl * dl) * (kC * kC));
    const double sp = (a.sigma_param != nullptr && a.want_logl) ? a.sigma_param[n] : 0.0;
    double logl_part = 0.0;
    for (int e = tid; e < E; e += blockDim.x) {
      const double dt_obs = a.epoch_tab[EP_T * E + e] - (a.t0 != nullptr ? a.t0[n] : 0.0);   // t0_base_model
      const double t = (dt_obs * kDay) / opz;
      const double nu = a.epoch_tab[EP_NU * E + e] * opz;
      double model_v;
      if (dt_obs < 0.0) {
        model_v = 0.0;                                                         // phase_models.py:52-58
      } else if (!(t >= tg[0] && t <= tg[G - 1])) {
        model_v = NAN;                       // interp1d raises ValueError in the reference
      } else {
        int lo = 0, hi = G;                  // searchsorted(tg, t, 'left')
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (tg[mid] < t) lo = mid + 1; else hi = mid;
        }
        const int i = min(max(lo, 1), G - 1);
        const double x0 = tg[i - 1], dx = tg[i] - x0;
        const double temp = ((Tg[i] - Tg[i - 1]) / dx) * (t - x0) + Tg[i - 1];
        const double rad = ((Rg[i] - Rg[i - 1]) / dx) * (t - x0) + Rg[i - 1];
        const double num = 2 * kPi * kH * (nu * nu * nu) * (rad * rad);         // sed.py:216
        const double frac = 1.0 / expm1((kH * nu) / (kKB * temp));
        model_v = num * inv_den * frac * kToMJy * opz;                          // :272
      }
      if (a.out_model != nullptr) a.out_model[n * (int64_t)E + e] = model_v;
      if (a.want_logl)
        logl_part += logl_term(a.cfg.sigma_mode, a.epoch_tab[EP_KIND * E + e], a.epoch_tab[EP_Y * E + e], model_v,
                               a.epoch_tab[EP_ISIG * E + e], a.epoch_tab[EP_LOGT * E + e], sp);
    }
    if (a.want_logl) {
      const double total = block_sum(logl_part, red);
      if (tid == 0) a.out_logl[n] = nan_to_num(total);
    }
  }
}

}  // namespace rb
