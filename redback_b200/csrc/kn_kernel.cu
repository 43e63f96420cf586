// redback_b200 -- fused kilonova light-curve + Gaussian log-likelihood kernels (FP64, sm_100a).
//
//   one_component_kilonova_model / _one_component_kilonova_model   kilonova_models.py:1537-1603, 1853-1893
//   metzger_kilonova_model / _metzger_kilonova_model               kilonova_models.py:1895-1962, 1965-2068
//   get_optimal_time_array (user_times branch)                      utils.py:42-80
//   interpolated_barnes_and_kasen_thermalisation_efficiency         utils.py:1161-1187
//   electron_fraction_from_kappa                                    utils.py:1479-1491
//
// Launches per call: epoch_table_kernel (shared with the SN path), kn_prep_kernel (one warp per sample:
// d_L by 32-point Gauss-Legendre + per-sample scalars), kn_kernel (persistent blocks, one sample at a
// time).  Inside kn_kernel:
//   phase A  one thread per point of the reference's adaptive time grid (<= dense_resolution points):
//            thermalisation, engine, and either
//              * one-component: integrand -> block-wide warp-shuffle scan = scipy cumulative_trapezoid
//              * Metzger: one thread per mass shell integrates its explicit-Euler energy equation over the
//                grid in registers; per step the shells' luminosities and the tau = 1 shell are reduced
//                with warp shuffles
//            -> temperature and photospheric radius on the grid, in shared memory
//   phase B  one thread per epoch: scipy linear interp1d of (T, R) at the source-frame epoch, blackbody
//            flux density (sed.py:199-222), Gaussian log-likelihood term, block reduction.
#include "rb_common.cuh"

namespace rb {

enum { KN_ONE_COMPONENT = RB_KN_ONE_COMPONENT, KN_METZGER = RB_KN_METZGER };
constexpr int kShells = 200;          // mass_len, kilonova_models.py:1983
constexpr int kShellWarps = 7;        // ceil(200 / 32)

enum {
  KS_OPZ = 0, KS_DL, KS_INV_DEN, KS_AV, KS_BV, KS_DV, KS_M0, KS_V0, KS_TDIFF, KS_TFLOOR, KS_SIGMA,
  KS_LOG_EMIN, KS_LOG_EMAX, KS_EMIN, KS_EMAX, KS_MEJ, KS_VEJ, KS_BETA, KS_KAPPA, KS_YE,
  KS_NB, KS_NU, KS_NA,  // segment sizes (stored as doubles)
  KS_PAD, kKnScalars    // = 24
};

struct KnArgs {
  rb_kn_config cfg;
  int64_t N;
  int E;
  const double* params;
  const double* dl_cm;
  const double* sigma_param;
  double* out_model;
  double* out_logl;
  const double* epoch_tab;
  double* scalars;
  double tmin_obs, tmax_obs;  // min / max of the observer-frame epochs [days]
  const double* t0;           // [N] (already offset to this chunk) or nullptr: per-sample explosion epoch
  const double* t_obs;        // [n_obs] ascending observer-frame epochs, needed with t0
  int n_obs;
  int want_logl;
  // grid mode (magnitude branch, kilonova_models.py:1579-1603): write T, R and the observer-frame phase
  // (time_temp*(1+z))/86400 on the model grid for phot_kernel instead of evaluating the epochs
  int64_t param_stride;  // distance between SoA rows of `params` (= total batch size)
  int grid_mode;
  double* grid_out;   // [N][4][dense_resolution]
  int* grid_len;      // [N]
  double* opz_out;    // [N]
  double* dl_out;     // [N]
};

// ---- correctly rounded atan for large arguments ------------------------------------------------------
// 0.5 - arctan(x)/pi (kilonova_models.py:1878, 2016) cancels catastrophically for x >> 1: a one-ulp
// difference in atan(x) becomes a relative error 2.2e-16 * pi * x of the engine luminosity.  numpy's
// arctan is correctly rounded for 99.9 % of arguments, so the device evaluates atan(x) = pi/2 - atan(1/x)
// in double-double and rounds once (error < 1e-19 before rounding).  For x < 16 the cancellation is
// harmless and CUDA's atan (1 ulp) is used.
__device__ __forceinline__ double atan_cr(double x) {
  if (!(x >= 16.0) || isinf(x)) return atan(x);
  const double yh = 1.0 / x;
  const double yl = fma(-yh, x, 1.0) / x;
  const double y2 = yh * yh;
  double p = fma(y2, 1.0 / 17.0, -1.0 / 15.0);
  p = fma(y2, p, 1.0 / 13.0);
  p = fma(y2, p, -1.0 / 11.0);
  p = fma(y2, p, 1.0 / 9.0);
  p = fma(y2, p, -1.0 / 7.0);
  p = fma(y2, p, 1.0 / 5.0);
  p = fma(y2, p, -1.0 / 3.0);
  const double corr = fma(yh * y2, p, yl * (1.0 - y2));  // atan(y) = yh + corr
  const double pio2_hi = 1.5707963267948966, pio2_lo = 6.123233995736766e-17;
  const double s1 = pio2_hi - yh;
  const double e1 = (pio2_hi - s1) - yh;
  return s1 + ((e1 + pio2_lo) - corr);
}

// Lookup tables live in global memory (two dynamically indexed function-local const arrays were observed to
// alias each other in the sm_100a build of this file, nvcc 12.9).
__device__ const double kBkMass[5] = {1.e-3, 5.e-3, 1.e-2, 5.e-2, 1.e-1};
__device__ const double kBkVel[4] = {0.1, 0.2, 0.3, 0.4};
// electron_fraction_from_kappa (utils.py:1487-1489), sorted by kappa as interp1d does
__device__ const double kYeKappa[8] = {0.5, 0.96, 3.30, 5.36, 5.60, 22.3, 32.2, 35};
__device__ const double kYeValue[8] = {0.5, 0.4, 0.35, 0.30, 0.25, 0.2, 0.15, 0.10};

// bilinear RegularGridInterpolator with linear extrapolation (scipy _rgi_cython.evaluate_linear_2d)
__device__ __forceinline__ double bk_interp(const double (*tab)[4], double mej, double vej) {
  int i0 = 0, i1 = 0;
  for (int i = 1; i <= 3; ++i) if (mej >= kBkMass[i]) i0 = i;   // x[i] <= v < x[i+1], clipped to [0, n-2]
  for (int i = 1; i <= 2; ++i) if (vej >= kBkVel[i]) i1 = i;
  const double y0 = (mej - kBkMass[i0]) / (kBkMass[i0 + 1] - kBkMass[i0]);
  const double y1 = (vej - kBkVel[i1]) / (kBkVel[i1 + 1] - kBkVel[i1]);
  return tab[i0][i1] * (1 - y0) * (1 - y1) + tab[i0][i1 + 1] * (1 - y0) * y1 + tab[i0 + 1][i1] * y0 * (1 - y1) +
         tab[i0 + 1][i1 + 1] * y0 * y1;
}

__device__ const double kBkA[5][4] = {{2.01, 4.52, 8.16, 16.3}, {0.81, 1.9, 3.2, 5.0}, {0.56, 1.31, 2.19, 3.0},
                                      {.27, .55, .95, 2.0}, {0.20, 0.39, 0.65, 0.9}};
__device__ const double kBkB[5][4] = {{0.28, 0.62, 1.19, 2.4}, {0.19, 0.28, 0.45, 0.65}, {0.17, 0.21, 0.31, 0.45},
                                      {0.10, 0.13, 0.15, 0.17}, {0.06, 0.11, 0.12, 0.12}};
__device__ const double kBkD[5][4] = {{1.12, 1.39, 1.52, 1.65}, {0.86, 1.21, 1.39, 1.5}, {0.74, 1.13, 1.32, 1.4},
                                      {0.6, 0.9, 1.13, 1.25}, {0.63, 0.79, 1.04, 1.5}};

// ---- per-sample scalars ------------------------------------------------------------------------------
__global__ void kn_prep_kernel(const KnArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int64_t N = a.param_stride;
  const double* P = a.params + n;
  const double z = P[RB_KN_REDSHIFT * N];
  const double opz = 1.0 + z;
  double dl;
  if (a.dl_cm != nullptr) {
    dl = a.dl_cm[n];
  } else {
    double s = RB_GL32_W[lane] * inv_efunc(a.cfg.cosmology, 0.5 * z * (RB_GL32_X[lane] + 1.0));
    s = warp_sum(s);
    dl = opz * a.cfg.cosmology.hubble_distance_cm * (0.5 * z * s);
  }
  if (lane != 0) return;
  const double mej = P[RB_KN_MEJ * N], vej = P[RB_KN_VEJ * N], kappa = P[RB_KN_KAPPA * N];
  double* S = a.scalars + n * kKnScalars;
  S[KS_OPZ] = opz;
  S[KS_DL] = dl;
  S[KS_INV_DEN] = 1.0 / ((dl * dl) * (kC * kC));
  S[KS_AV] = bk_interp(kBkA, mej, vej);
  S[KS_BV] = bk_interp(kBkB, mej, vej);
  S[KS_DV] = bk_interp(kBkD, mej, vej);
  const double m0 = mej * kMsun, v0 = vej * kC;
  S[KS_M0] = m0;
  S[KS_V0] = v0;
  S[KS_TDIFF] = sqrt(2.0 * kappa * m0 / (13.7 * v0 * kC));                      // kilonova_models.py:1877
  S[KS_TFLOOR] = (a.cfg.model == KN_ONE_COMPONENT) ? P[RB_KN_TEMPERATURE_FLOOR * N] : 0.0;
  S[KS_SIGMA] = (a.sigma_param != nullptr && a.want_logl) ? a.sigma_param[n] : 0.0;
  // get_optimal_time_array(1e-2, 7e6, R, user_times = t_obs * 86400 / (1+z))  (utils.py:50-80)
  const double t_min = 1e-2, t_max = a.cfg.grid_t_max;
  double first = a.tmin_obs, last = a.tmax_obs, t0 = 0.0;
  if (a.t0 != nullptr) {                                                         // phase_models.py:33-45
    t0 = a.t0[n];
    int lo = 0, hi = a.n_obs;              // first epoch with t_obs - t0 >= 0
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (a.t_obs[mid] - t0 >= 0.0) hi = mid; else lo = mid + 1;
    }
    if (lo < a.n_obs) { first = a.t_obs[lo] - t0; last = a.t_obs[a.n_obs - 1] - t0; }
    else first = last = 1.0;               // no epoch after t0: every output is the placeholder
  }
  const double user_min = first * kDay / opz, user_max = last * kDay / opz;
  const double eval_min = fmax(t_min, user_min * 0.5);
  const double eval_max = fmin(t_max, user_max * 2.0);
  const int R = a.cfg.dense_resolution;
  int nb = (int)(R * 0.15), nu = (int)(R * 0.70);
  int na = R - nb - nu;
  if (!(eval_min > t_min)) { nu += nb; nb = 1; }
  if (!(eval_max < t_max)) { nu += na; na = 1; }
  S[KS_EMIN] = eval_min;
  S[KS_EMAX] = eval_max;
  S[KS_LOG_EMIN] = log10(eval_min);
  S[KS_LOG_EMAX] = log10(eval_max);
  S[KS_NB] = (double)nb;
  S[KS_NU] = (double)nu;
  S[KS_NA] = (double)na;
  S[KS_MEJ] = mej;
  S[KS_VEJ] = vej;
  S[KS_KAPPA] = kappa;
  double beta = 0.0, ye = 0.0;
  if (a.cfg.model == KN_METZGER) {
    beta = P[RB_KN_BETA * N];
    // electron_fraction_from_kappa: interp1d on the kappa-sorted Tanaka+19 table, linear extrapolation
    int i = 1;  // searchsorted(kx, kappa) clipped to [1, 7]
    while (i < 7 && kYeKappa[i] < kappa) ++i;
    const double slope = (kYeValue[i] - kYeValue[i - 1]) / (kYeKappa[i] - kYeKappa[i - 1]);
    ye = slope * (kappa - kYeKappa[i - 1]) + kYeValue[i - 1];
  }
  S[KS_BETA] = beta;
  S[KS_YE] = ye;
  S[KS_PAD] = t0;
}

// np.geomspace(a, b, n)[i] with log10 endpoints la, lb precomputed (numpy function_base.geomspace)
__device__ __forceinline__ double geom_point(double a, double b, double la, double lb, int n, int i,
                                             const double* __restrict__ tab) {
  if (i == 0) return a;
  if (i == n - 1) return b;
  const double step = (lb - la) / (double)(n - 1);
  const double y = (double)i * step + la;
  return exp_tab(y * 2.302585092994046, tab);  // 10**y (|y| <= 7: inside exp_tab's range)
}

// inclusive block scan of v over consecutive chunks of blockDim.x; `carry` (shared) holds the running total
__device__ __forceinline__ double block_scan_inclusive(double v, double* wsum, double* carry) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += u;
  }
  if (lane == 31) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double w = (lane < nw) ? wsum[lane] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += u;
    }
    if (lane < nw) wsum[lane] = w;  // inclusive warp totals
  }
  __syncthreads();
  const double base = *carry + (warp > 0 ? wsum[warp - 1] : 0.0);
  const double out = base + v;
  __syncthreads();
  if (threadIdx.x == blockDim.x - 1) *carry = out;
  __syncthreads();
  return out;
}

// MODEL resolved at compile time: each instantiation carries only its own model's code (instruction cache, registers)
template <int MODEL>
__global__ void __launch_bounds__(512, 1) kn_kernel(const KnArgs a) {
  extern __shared__ double ksm[];
  const int R = a.cfg.dense_resolution;
  const int E = a.E;
  double* tg = ksm;                 // [R] time grid (source-frame seconds)
  double* Tg = tg + R;              // [R] temperature on the grid
  double* Rg = Tg + R;              // [R] photospheric radius on the grid
  double* w1 = Rg + R;              // [R] work: integrand | edot_base
  double* w2 = w1 + R;              // [R] work: L | exp(-t/900)
  double* tab = w2 + R;             // [1024]
  double* sc = tab + kExpTabSize;           // [kKnScalars]
  double* wsum = sc + kKnScalars;   // [34] scan / reduction scratch (+ carry at [33])
  double* ep = wsum + 34;           // [kEpochCols][E]
  double* lpart = ep + kEpochCols * E;                 // Metzger: [kShellWarps][R] partial luminosities
  double* tpart = lpart + kShellWarps * R;             // Metzger: [kShellWarps][R] min |tau - 1|
  int* ipart = reinterpret_cast<int*>(tpart + kShellWarps * R);  // Metzger: [kShellWarps][R] argmin shell

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int model = MODEL;
  const int sigma_mode = a.cfg.sigma_mode;
  const bool want_logl = a.want_logl != 0;
  const bool fp32 = a.cfg.precision == RB_PREC_F32;
  const double grid_t_max = a.cfg.grid_t_max, log_grid_t_max = log10(a.cfg.grid_t_max);
  for (int i = tid; i < kExpTabSize; i += blockDim.x) tab[i] = RB_EXP2_1024[i];
  for (int i = tid; i < kEpochCols * E; i += blockDim.x) ep[i] = a.epoch_tab[i];
  __syncthreads();

  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    if (tid < kKnScalars) sc[tid] = a.scalars[n * kKnScalars + tid];
    __syncthreads();
    const int nb = (int)sc[KS_NB], nu = (int)sc[KS_NU], na = (int)sc[KS_NA];
    const int G = nb + nu + na - 2;   // np.unique drops the two junction duplicates
    const double av = sc[KS_AV], bv = sc[KS_BV], dv = sc[KS_DV];
    const double tdiff = sc[KS_TDIFF], m0 = sc[KS_M0], v0 = sc[KS_V0];

    // ---- phase A.1: time grid + per-point physics ------------------------------------------------
    for (int k = tid; k < G; k += blockDim.x) {
      double t;
      if (k < nb) t = geom_point(1e-2, sc[KS_EMIN], -2.0, sc[KS_LOG_EMIN], nb, k, tab);
      else if (k < nb + nu - 1) t = geom_point(sc[KS_EMIN], sc[KS_EMAX], sc[KS_LOG_EMIN], sc[KS_LOG_EMAX], nu, k - nb + 1, tab);
      else t = geom_point(sc[KS_EMAX], grid_t_max, sc[KS_LOG_EMAX], log_grid_t_max, na, k - nb - nu + 2, tab);
      if (nb == 1 && k == 0) t = 1e-2;
      tg[k] = t;
      const double tdays = t / kDay;
      if (model == KN_ONE_COMPONENT && fp32) {
        // RB_PREC_F32: the smooth O(1) factors in single precision (six of the grid point's nine transcendentals); the
        // shape 0.5 - atan(u) / pi is taken as atan(1 / u) / pi for u > 1, where the subtraction would cancel in FP32
        // (down to 5e-9 at the end of the grid).  Everything that needs the FP64 range stays FP64.
        const float tdf = (float)tdays;
        const float xxf = 2.0f * (float)bv * powf(tdf, (float)dv);
        const float ethf = 0.36f * (expf(-(float)av * tdf) + log1pf(xxf) / xxf);
        const float u = (float)((t - 1.3) * (1.0 / 0.11));
        const float shape = (u > 1.0f) ? atanf(1.0f / u) * 0.31830988618379067f
                                       : 0.5f - atanf(u) * 0.31830988618379067f;
        const double smooth = (double)(powf(shape, 1.3f) * ethf);
        w1[k] = (4.0e18 * m0) * smooth * (t / tdiff) * exp((t * t) / (tdiff * tdiff));
        continue;
      }
      const double xx = 2.0 * bv * pow(tdays, dv);
      const double e_th = 0.36 * (exp(-av * tdays) + log1p(xx) / xx);          // :1867
      if (model == KN_ONE_COMPONENT) {
        const double lum_in = 4.0e18 * m0 * pow(0.5 - atan_cr((t - 1.3) / 0.11) / kPi, 1.3);   // :1878
        w1[k] = lum_in * e_th * (t / tdiff) * exp((t * t) / (tdiff * tdiff));  // :1879
      } else {
        double eb;
        if (t > 1.3) eb = 2.1e10 * e_th * pow(t / (3600. * 24.), -1.3);        // :2013
        else eb = 4.0e18 * pow(0.5 - (1. / kPi) * atan_cr((t - 1.3) / 0.11), 1.3) * e_th;      // :2014
        w1[k] = eb;
        w2[k] = exp(-t / 900.0);                                               // :2032
        Tg[k] = 1.0 / t;             // shared by the 200 shell threads of the march (Tg is filled after it)
      }
    }
    __syncthreads();

    if (model == KN_ONE_COMPONENT) {
      // ---- phase A.2: cumulative_trapezoid (block scan), diffusion, photosphere ------------------
      if (tid == 0) wsum[33] = 0.0;
      __syncthreads();
      for (int base = 0; base < G; base += blockDim.x) {
        const int k = base + tid;
        double inc = 0.0;
        if (k >= 1 && k < G) inc = (tg[k] - tg[k - 1]) * (w1[k] + w1[k - 1]) / 2.0;
        const double cum = block_scan_inclusive(inc, wsum, wsum + 33);
        if (k < G) w2[k] = cum;
      }
      __syncthreads();
      const double tfloor = sc[KS_TFLOOR];
      const double tf2 = tfloor * tfloor;
      for (int k = tid; k < G; k += blockDim.x) {
        const double t = tg[k];
        double lb = (k == 0) ? w2[1] : w2[k];                                  // :1881-1882
        lb = lb * exp(-(t * t) / (tdiff * tdiff)) / tdiff;                     // :1883
        double temp = sqrt(sqrt(lb / (4.0 * kPi * kSigmaSB * (v0 * v0) * (t * t))));   // :1885
        double rad = sqrt(lb / (4.0 * kPi * kSigmaSB * (tf2 * tf2)));          // :1886
        if (temp <= tfloor) temp = tfloor; else rad = v0 * t;                  // :1889-1892
        Tg[k] = temp;
        Rg[k] = rad;
      }
    } else {
      // ---- phase A.2 (Metzger): one thread per mass shell, explicit Euler over the grid ----------
      if (tid < kShellWarps * 32) {
        const int s = tid;
        const bool active = s < kShells;
        const double mej = sc[KS_MEJ], vej = sc[KS_VEJ], beta = sc[KS_BETA], kappa = sc[KS_KAPPA];
        const double vstep = (a.cfg.vmax - vej) / (double)(kShells - 1);       // np.linspace(vmin, vmax, 200)
        const int sc_ = min(s, kShells - 1);
        const double vel = (sc_ == kShells - 1) ? a.cfg.vmax : (double)sc_ * vstep + vej;
        const double vel1 = (sc_ + 1 >= kShells - 1) ? a.cfg.vmax : (double)(sc_ + 1) * vstep + vej;
        const double ms = mej * pow(vel / vej, -beta);                         // :1999
        const double ms1 = mej * pow(vel1 / vej, -beta);
        const double dm = fabs(ms1 - ms);                                      // :2040
        const double vm = vel * kC;
        double xn0 = 0.0, xr = 1.0;
        const bool npre = a.cfg.neutron_precursor_switch != 0;
        if (npre) {
          xn0 = (2.0 / kPi) * (1.0 - sc[KS_YE]) * atan(1e-4 * kMsun / ms / kMsun);   // :2027-2028
          xr = 1.0 - xn0;
        }
        double en = 0.5 * ms * (vm * vm);                                      // :2043
        const bool ode = active && s < kShells - 1;                            // the [:-1] slices
        // per-shell constants of the step: the four divisions of :2050-2059 collapse into one (1 / t is shared by
        // all shells and tabulated per grid point)
        const double c_td = (ms * kMsun * 3) / (4 * kPi * vm * kC * beta);     // t_d = kappa c_td / t        (:2050)
        const double c_tau = (ms * kMsun) / (4 * kPi * (vm * vm));             // tau = kappa c_tau / t^2     (:2052)
        const double vmc = vm / kC;
        for (int ii = 0; ii < G - 1; ++ii) {
          const double t = tg[ii], it = Tg[ii];
          double kap = kappa, edot = w1[ii];
          if (npre) {
            const double xn = xn0 * w2[ii];
            edot = 3.2e14 * xn + edot;                                         // :2033-2034
            kap = 0.4 * (1.0 - xn - xr) + kappa * xr;                          // :2035-2037
          }
          const double td = (kap * c_td) * it;
          const double tau = (kap * c_tau) * (it * it);
          const double lum = en / (td + t * vmc);                               // :2058
          const double dt = tg[ii + 1] - t;
          en = (edot - (en * it) - lum) * dt + en;                              // :2059
          double lo = ode ? lum * dm * kMsun : 0.0;                             // :2060
          // shell 199 copies tau of shell 198 (:2062): it can never win the first-minimum argmin
          double dist = ode ? fabs(tau - 1.0) : INFINITY;
          if (isnan(dist)) dist = -INFINITY;  // np.argmin returns the first NaN
          int idx = s;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) lo += __shfl_xor_sync(0xffffffffu, lo, o);
          warp_argmin_first(dist, idx);
          if (lane == 0) {
            lpart[warp * R + ii] = lo;
            tpart[warp * R + ii] = dist;
            ipart[warp * R + ii] = idx;
          }
        }
      }
      __syncthreads();
      const double vstep = (a.cfg.vmax - sc[KS_VEJ]) / (double)(kShells - 1);
      for (int k = tid; k < G; k += blockDim.x) {
        double lb = 0.0, rad = 0.0;
        if (k < G - 1) {
          double best = tpart[k];
          int bi = ipart[k];
          lb = lpart[k];
          for (int w = 1; w < kShellWarps; ++w) {
            lb += lpart[w * R + k];
            const double d = tpart[w * R + k];
            if (d < best) { best = d; bi = ipart[w * R + k]; }
          }
          const double vel = (bi == kShells - 1) ? a.cfg.vmax : (double)bi * vstep + sc[KS_VEJ];
          rad = (vel * kC) * tg[k];                                             // :2064-2065
        }
        Tg[k] = sqrt(sqrt(lb / (4.0 * kPi * (rad * rad) * kSigmaSB)));         // :2069
        Rg[k] = rad;
      }
    }
    __syncthreads();

    const double opz = sc[KS_OPZ];
    if (a.grid_mode) {
      double* go = a.grid_out + (size_t)n * 4 * R;
      for (int k = tid; k < G; k += blockDim.x) {
        go[0 * R + k] = Tg[k];
        go[1 * R + k] = Rg[k];
        go[2 * R + k] = 0.0;
        go[3 * R + k] = (tg[k] * opz) / kDay;                                  // :1580, :1600
      }
      if (tid == 0) { a.grid_len[n] = G; a.opz_out[n] = opz; a.dl_out[n] = sc[KS_DL]; }
      __syncthreads();
      continue;
    }
    // ---- phase B: epochs -----------------------------------------------------------------------
    double logl_part = 0.0;
    for (int e = tid; e < E; e += blockDim.x) {
      const double dt_obs = ep[EP_T * E + e] - sc[KS_PAD];                     // t0_base_model; t0 = 0 otherwise
      const double t = (dt_obs * kDay) / opz;                                  // :1563, utils.py:382
      const double nu = ep[EP_NU * E + e] * opz;
      double model_v;
      if (dt_obs < 0.0) {
        model_v = 0.0;                                                         // phase_models.py:52-58
      } else if (!(t >= tg[0] && t <= tg[G - 1])) {
        model_v = NAN;  // interp1d would raise ValueError (bounds_error): flagged as NaN on device
      } else {
        int lo = 0, hi = G;  // searchsorted(tg, t, 'left')
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (tg[mid] < t) lo = mid + 1; else hi = mid;
        }
        const int i = min(max(lo, 1), G - 1);
        const double x0 = tg[i - 1], dx = tg[i] - x0;
        const double temp = ((Tg[i] - Tg[i - 1]) / dx) * (t - x0) + Tg[i - 1];  // scipy _call_linear
        const double rad = ((Rg[i] - Rg[i - 1]) / dx) * (t - x0) + Rg[i - 1];
        const double num = 2 * kPi * kH * (nu * nu * nu) * (rad * rad);         // sed.py:216
        const double frac = 1.0 / expm1((kH * nu) / (kKB * temp));
        model_v = num * sc[KS_INV_DEN] * frac * kToMJy * opz;                   // :1577
      }
      if (a.out_model != nullptr) a.out_model[n * (int64_t)E + e] = model_v;
      if (want_logl) {
        logl_part += logl_term(sigma_mode, ep[EP_KIND * E + e], ep[EP_Y * E + e], model_v, ep[EP_ISIG * E + e],
                               ep[EP_LOGT * E + e],
                               sc[KS_SIGMA]);
      }
    }
    if (want_logl) {
      logl_part = warp_sum(logl_part);
      if (lane == 0) wsum[warp] = logl_part;
    }
    __syncthreads();
    if (want_logl && tid == 0) {
      double total = 0.0;
      const int nw = (blockDim.x + 31) >> 5;
      for (int w = 0; w < nw; ++w) total += wsum[w];
      a.out_logl[n] = nan_to_num(total);
    }
    __syncthreads();
  }
}

}  // namespace rb
