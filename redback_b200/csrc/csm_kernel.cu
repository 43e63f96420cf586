// CSM-interaction path: _csm_engine (supernova_models.py:1910-1997) on the adaptive dense grid
// of get_optimal_time_array (utils.py:42-145), CSMDiffusion's cumulative recurrence (interaction_processes.py:176-197),
// TemperatureFloor + SED epilogue and the Gaussian likelihood, one block per parameter sample.
//
// The recurrence  new[i+1] = new[i] e^{-dt/t0} + L_i (1 - e^{-dt/t0}) + s_i (dt - t0 (1 - e^{-dt/t0}))  over the merged
// knots unique(dense ∪ epochs) is a chain of affine maps; it is evaluated as a block-wide scan of (a_i, b_i) pairs
// (per-thread chunks + warp-shuffle scan), not by the reference's sequential Python loop.
// Two further options of the same path:
//  * csm_diffusion_method = "quadrature" | "legacy" (interaction_processes.py:199-227): per unique epoch the convolution
//    with exp((t - t_e) / t0) / t0 on `csm_diffusion_timesteps` abscissae between the grid start and the epoch, the
//    luminosity by np.interp on the adaptive grid; one thread per unique epoch, the interval index walks forward;
//  * csm_nickel (supernova_models.py:2120-2165): the nickel-cobalt engine on the SAME adaptive grid through
//    Diffusion (:33-65) with mej_eff = mej + mass_csm_threshold / Msun, added to the CSM luminosity.
// (included by capi.cu after sn_kernel.cu and kn_kernel.cu, whose device helpers it uses)

namespace rb {

#include "rb_csm_table.inc"

constexpr double kAu = 14959787070000.0;   // constants.au_cgs

enum CsmScalar {
  CS_OPZ = 0, CS_T0, CS_TFS, CS_TRS, CS_CFS, CS_CRS, CS_ALPHA, CS_EFF,
  CS_RCV, CS_TFLOOR, CS_INV_REC, CS_INV_DEN, CS_DL, CS_SIGMA,
  CS_TMIN, CS_EMIN, CS_EMAX, CS_TMAX, CS_LTMIN, CS_LEMIN, CS_LEMAX, CS_LTMAX, CS_NB, CS_NU, CS_NA, CS_MTH,
  kCsmScalars
};

struct CsmArgs {
  SnArgs s;
  const double* utime;   // [E] sorted unique observer-frame epochs (first *n_unique valid)
  const int* uidx;       // [E] epoch -> index into utime
  const int* n_unique;   // device scalar
  double* cscal;         // [N][kCsmScalars]
};

// sorted unique epochs (np.unique inside CSMDiffusion, interaction_processes.py:181-182) -- one block, O(E^2 / threads)
__global__ void unique_times_kernel(int E, const double* __restrict__ t, double* __restrict__ utime,
                                    int* __restrict__ uidx, int* __restrict__ n_unique) {
  extern __shared__ int first_occ[];
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  int mine = 0;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const double te = t[e];
    int first = 1;
    for (int f = 0; f < e; ++f)
      if (t[f] == te) { first = 0; break; }
    first_occ[e] = first;
    mine += first;
  }
  atomicAdd(&total, mine);
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const double te = t[e];
    int cnt = 0;
    for (int f = 0; f < E; ++f) cnt += (first_occ[f] && t[f] < te) ? 1 : 0;
    uidx[e] = cnt;
    if (first_occ[e]) utime[cnt] = te;
  }
  if (threadIdx.x == 0) *n_unique = total;
}

// utils.get_csm_properties (utils.py:288-315): bilinear interpolation on (n, s); NaN outside the table (the reference
// raises "out of bounds")
__device__ inline void csm_properties(double nn, double eta, double& aa, double& bf, double& br) {
  aa = bf = br = NAN;
  if (!(nn >= RB_CSM_N[0] && nn <= RB_CSM_N[29] && eta >= RB_CSM_S[0] && eta <= RB_CSM_S[9])) return;
  int i = 0, j = 0;
  while (i < 28 && RB_CSM_N[i + 1] < nn) ++i;      // searchsorted(grid, x) - 1, clipped to [0, size - 2]
  while (j < 8 && RB_CSM_S[j + 1] < eta) ++j;
  const double y0 = (nn - RB_CSM_N[i]) / (RB_CSM_N[i + 1] - RB_CSM_N[i]);
  const double y1 = (eta - RB_CSM_S[j]) / (RB_CSM_S[j + 1] - RB_CSM_S[j]);
  const double w00 = (1.0 - y0) * (1.0 - y1), w01 = (1.0 - y0) * y1, w10 = y0 * (1.0 - y1), w11 = y0 * y1;
  const int q = i * 10 + j;
  aa = RB_CSM_AA[q] * w00 + RB_CSM_AA[q + 1] * w01 + RB_CSM_AA[q + 10] * w10 + RB_CSM_AA[q + 11] * w11;
  bf = RB_CSM_BF[q] * w00 + RB_CSM_BF[q + 1] * w01 + RB_CSM_BF[q + 10] * w10 + RB_CSM_BF[q + 11] * w11;
  br = RB_CSM_BR[q] * w00 + RB_CSM_BR[q + 1] * w01 + RB_CSM_BR[q + 10] * w10 + RB_CSM_BR[q + 11] * w11;
}

// ---- per-sample scalars: shock constants, diffusion time scale, dense-grid segments, d_L ------------------------
__global__ void csm_prep_kernel(const CsmArgs c) {
  const SnArgs& a = c.s;
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
      double s = RB_GL32_W[lane] * inv_efunc(a.cfg.cosmology, 0.5 * z * (RB_GL32_X[lane] + 1.0));
      s = warp_sum(s);
      dl = opz * a.cfg.cosmology.hubble_distance_cm * (0.5 * z * s);
    }
  }
  if (lane != 0) return;
  if (a.grid_mode) { a.opz_out[n] = opz; a.dl_out[n] = dl; }
  double* S = c.cscal + n * kCsmScalars;
  const double nn = a.cfg.csm_nn, delta = a.cfg.csm_delta;
  const double eta = P[RB_SN_ETA * N], rho = P[RB_SN_RHO * N], kappa = P[RB_SN_KAPPA * N];
  const double vkm = P[RB_SN_VEJ * N];
  const double mej = P[RB_SN_MEJ * N] * kMsun;                                  // supernova_models.py:1927-1932
  const double mcsm = P[RB_SN_CSM_MASS * N] * kMsun;
  const double r0 = P[RB_SN_R0 * N] * kAu;
  const double v = vkm * kKm;
  const double esn = 3. * (v * v) * mej / 10.;
  double aa, bf, br;
  csm_properties(nn, eta, aa, bf, br);
  const double qq = rho * pow(r0, eta);                                                                   // :1942
  const double rcsm = pow((3.0 - eta) / (4.0 * kPi * qq) * mcsm + pow(r0, 3.0 - eta), 1.0 / (3.0 - eta));  // :1944
  const double rph = fabs(pow(-2.0 * (1.0 - eta) / (3.0 * kappa * qq) + pow(rcsm, 1.0 - eta), 1.0 / (1.0 - eta)));
  const double mth = fabs(4.0 * kPi * qq / (3.0 - eta) * (pow(rph, 3.0 - eta) - pow(r0, 3.0 - eta)));      // :1949
  const double gn = 1.0 / (4.0 * kPi * (nn - delta)) * pow(2.0 * (5.0 - delta) * (nn - 5.0) * esn, (nn - 3.) / 2.0) /
                    pow((3.0 - delta) * (nn - 3.0) * mej, (nn - 5.0) / 2.0);                                // :1953
  const double pw = (nn - eta) / ((nn - 3.0) * (3.0 - eta));
  const double tfs = pow(fabs((3.0 - eta) * pow(qq, (3.0 - nn) / (nn - eta)) * pow(aa * gn, (eta - 3.0) / (nn - eta)) /
                              (4.0 * kPi * pow(bf, 3.0 - eta))), pw) * pow(mth, pw);                       // :1958
  const double trs = pow(v / (br * pow(aa * gn / qq, 1.0 / (nn - eta))) *
                         pow(1.0 - (3.0 - nn) * mej / (4.0 * kPi * pow(v, 3.0 - nn) * gn), 1.0 / (3.0 - nn)),
                         (nn - eta) / (eta - 3.0));                                                        // :1965
  const double nme = nn - eta;
  S[CS_OPZ] = opz;
  S[CS_T0] = kappa * mth / (4. * (kPi * kPi * kPi) / 9. * kC * rph) / kDay;      // interaction_processes.py:171-173
  S[CS_TFS] = tfs;
  S[CS_TRS] = trs;
  S[CS_CFS] = 2.0 * kPi / (nme * nme * nme) * pow(gn, (5.0 - eta) / nme) * pow(qq, (nn - 5.0) / nme) *
              ((nn - 3.0) * (nn - 3.0)) * (nn - 5.0) * pow(bf, 5.0 - eta) * pow(aa, (5.0 - eta) / nme);   // :1975
  const double r3 = (3.0 - eta) / nme;
  S[CS_CRS] = 2.0 * kPi * pow(aa * gn / qq, (5.0 - nn) / nme) * pow(br, 5.0 - nn) * gn * (r3 * r3 * r3);  // :1979
  S[CS_ALPHA] = (2.0 * nn + 6.0 * eta - nn * eta - 15.) / nme;
  S[CS_EFF] = isnan(aa) ? NAN : a.cfg.csm_efficiency;   // eta outside the table: the reference raises, NaN here
  S[CS_RCV] = kRadiusConst * vkm;
  double tfloor = 0.0, inv_rec = 0.0;
  if (sed != RB_SED_BOLOMETRIC) {
    tfloor = P[RB_SN_TEMPERATURE_FLOOR * N];
    const double tf2 = tfloor * tfloor;
    inv_rec = 1.0 / (kStefConst * (tf2 * tf2));
  }
  S[CS_TFLOOR] = tfloor;
  S[CS_INV_REC] = inv_rec;
  S[CS_INV_DEN] = 1.0 / ((dl * dl) * (kC * kC));
  S[CS_DL] = dl;
  S[CS_SIGMA] = (a.sigma_param != nullptr && a.want_logl) ? a.sigma_param[n] : 0.0;
  // dense grid of csm_interaction_bolometric (:2026-2030): get_optimal_time_array(min(0.1, min t), max t + 100, D, t)
  const int Eu = *c.n_unique;
  const double u0 = c.utime[0], u1 = c.utime[Eu - 1];
  const double umin = a.grid_mode ? (u0 * opz) / opz : u0 / opz;
  const double umax = a.grid_mode ? (u1 * opz) / opz : u1 / opz;
  // dense_time_min: min(0.1, min t) for csm_interaction (:2028), min(0.01, min t) for csm_nickel (:2143)
  const double tmin = fmin(a.cfg.f_nickel != nullptr ? 0.01 : 0.1, umin), tmax = umax + 100.0;
  const double emin = fmax(tmin, umin * 0.5), emax = fmin(tmax, umax * 2.0);       // utils.py:57-58
  const int D = a.cfg.dense_resolution;
  int nb = (int)((double)D * 0.15), nu = (int)((double)D * 0.70);
  int na = D - nb - nu;
  if (!(emin > tmin)) { nu += nb; nb = 1; }                                         // utils.py:64-68
  if (!(emax < tmax)) { nu += na; na = 1; }                                         // utils.py:72-77
  S[CS_TMIN] = tmin;
  S[CS_EMIN] = emin;
  S[CS_EMAX] = emax;
  S[CS_TMAX] = tmax;
  S[CS_LTMIN] = log10(tmin);
  S[CS_LEMIN] = log10(emin);
  S[CS_LEMAX] = log10(emax);
  S[CS_LTMAX] = log10(tmax);
  S[CS_NB] = (double)nb;
  S[CS_NU] = (double)nu;
  S[CS_NA] = (double)na;
  S[CS_MTH] = mth;                                                                  // mass_csm_threshold [g], :1949
}

// photosphere.TemperatureFloor (photosphere.py:87-111) for one epoch
__device__ __forceinline__ void temperature_floor_dev(double lbol, double te, double rcv, double tfloor,
                                                      double inv_rec, double& tph, double& rph) {
  const double rr = rcv * te;
  const double r2 = rr * rr;
  const double rec2 = lbol * inv_rec;
  const bool mask = r2 <= rec2;
  rph = sqrt(mask ? r2 : rec2);
  tph = mask ? sqrt(sqrt(lbol / (kStefConst * r2))) : tfloor;
}

// SED (sed.py:199-222, 301-426, 859-909) for one epoch: observer-frame flux density in mJy (already times 1 + z)
__device__ __forceinline__ double sed_mjy(const rb_sn_config& cfg, double lbol, double te, double tph, double rph,
                                          double nu_obs, double opz, double inv_den, double dl,
                                          const double* __restrict__ tab) {
  const double nu = nu_obs * opz;                                               // utils.py:383
  if (cfg.sed == RB_SED_BLACKBODY) {
    const double num = 2 * kPi * kH * (nu * nu * nu) * (rph * rph);
    const double em = expm1_fast((kH * nu) / (kKB * tph), tab);
    return num * inv_den / em * kToMJy * opz;
  }
  double mjy = cutoff_blackbody_mjy(cfg, lbol, rph, tph, nu, dl);
  if (cfg.sed == RB_SED_CUTOFF_LINE) mjy = line_sed_mjy(cfg, mjy, lbol, te, nu, dl);
  return mjy * opz;
}

constexpr int kCsmThreads = 512;

__global__ void __launch_bounds__(kCsmThreads, 2) csm_kernel(const CsmArgs c) {
  extern __shared__ double csm_smem[];
  const SnArgs& a = c.s;
  const int D = a.cfg.dense_resolution, E = a.E;
  const int Eu = *c.n_unique;
  const int Mmax = D + Eu;
  double* tab = csm_smem;              // [1024] 2^(j/1024)
  double* tg = tab + kExpTabSize;      // [D]    dense grid (days, source frame)
  double* ld = tg + D;                 // [D]    engine luminosity on it
  double* mt = ld + D;                 // [Mmax] merged knots
  double* ml = mt + Mmax;              // [Mmax] luminosity at the knots; after the scan: diffused luminosity
  double* ca = ml + Mmax;              // [Mmax] affine map of interval i: new' = ca new + cb
  double* cb = ca + Mmax;              // [Mmax]
  // the four knot arrays also hold the quadrature method's timesteps / 2 + timesteps abscissae (host: csm_smem_bytes)
  const int M4 = max(4 * Mmax, a.cfg.csm_quadrature ? 3 * ((a.cfg.csm_timesteps + 1) / 2) + 4 : 0);
  double* us = mt + M4;                // [Eu]   unique source-frame epochs
  double* ldn = us + Eu;               // [D]    nickel-cobalt luminosity on the dense grid (csm_nickel)
  double* lq = ldn + D;                // [Eu]   CSM luminosity at the unique epochs, quadrature method
  double* lni = lq + Eu;               // [Eu]   diffused nickel luminosity at the unique epochs (csm_nickel)
  double* nk = lni + Eu;               // [160]  Diffusion's 50 + 100 abscissae (csm_nickel)
  double* sc = nk + 160;               // [32]   per-sample scalars
  double* wsa = sc + 32;               // [32]   warp aggregates
  double* wsb = wsa + 32;              // [32]
  int* upos = reinterpret_cast<int*>(wsb + 32);   // [Eu] position of each unique epoch among the knots

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int sed = a.cfg.sed;
  const bool want_logl = a.want_logl != 0;
  const bool quad = a.cfg.csm_quadrature != 0;
  const bool nickel = a.cfg.f_nickel != nullptr;
  const int nq_half = (a.cfg.csm_timesteps + 1) / 2;     // int(round(timesteps / 2.0)) for the even defaults, :209
  const int NQ = 2 * nq_half;
  double* qlsp = mt;                   // quadrature method: [nq_half] and [NQ] abscissae in the (then unused) knot arrays
  double* qx = mt + nq_half;
  for (int i = tid; i < kExpTabSize; i += blockDim.x) tab[i] = RB_EXP2_1024[i];

  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();   // previous sample finished with every shared array
    if (tid < kCsmScalars) sc[tid] = c.cscal[n * kCsmScalars + tid];
    __syncthreads();
    const int nb = (int)sc[CS_NB], nu = (int)sc[CS_NU], na = (int)sc[CS_NA];
    const int G = nb + nu + na - 2;          // np.unique drops the two shared segment end points
    const double opz = sc[CS_OPZ];
    // ---- dense grid + engine (supernova_models.py:1972-1985) ----------------------------------------------------
    {
      const double alpha = sc[CS_ALPHA], tfs = sc[CS_TFS], trs = sc[CS_TRS], cfs = sc[CS_CFS], crs = sc[CS_CRS];
      for (int k = tid; k < G; k += blockDim.x) {
        double t;
        if (k < nb) t = geom_point(sc[CS_TMIN], sc[CS_EMIN], sc[CS_LTMIN], sc[CS_LEMIN], nb, k, tab);
        else if (k < nb + nu - 1) t = geom_point(sc[CS_EMIN], sc[CS_EMAX], sc[CS_LEMIN], sc[CS_LEMAX], nu, k - nb + 1, tab);
        else t = geom_point(sc[CS_EMAX], sc[CS_TMAX], sc[CS_LEMAX], sc[CS_LTMAX], na, k - nb - nu + 2, tab);
        if (nb == 1 && k == 0) t = sc[CS_TMIN];
        if (na == 1 && k == G - 1) t = sc[CS_TMAX];
        tg[k] = t;
        const double ts = t * kDay;
        // (ts + ti)^alpha, |alpha ln x| < 40: 1e-15 relative through the table exp
        const double pw = exp_tab(alpha * log(ts + 1.0), tab);
        const double lfs = (tfs - ts > 0) ? cfs * pw : 0.0;
        const double lrs = (trs - ts > 0) ? crs * pw : 0.0;
        ld[k] = sc[CS_EFF] * (lfs + lrs);
        if (nickel)                                                              // supernova_models.py:564-579 at :2152
          ldn[k] = a.cfg.f_nickel[n] * a.params[RB_SN_MEJ * a.param_stride + n] *
                   (6.45e43 * exp(-t / 8.8) + 1.45e43 * exp(-t / 111.3));
      }
      for (int j = tid; j < Eu; j += blockDim.x) {
        const double u = c.utime[j];
        us[j] = a.grid_mode ? (u * opz) / opz : u / opz;                        // utils.py:382
      }
    }
    __syncthreads();
    if (quad) {
      // ---- csm_diffusion_method = "quadrature" (interaction_processes.py:199-227) ------------------------------------
      const double t0 = sc[CS_T0], t_end = tg[G - 1], tb = fmax(0.0, tg[0]);                       // :203-206
      {
        const double a0 = fmin(log10(t0 / t_end) - 3.0, 0.0);                                      // :210
        const double da = (0.0 - a0) / (double)(nq_half - 1);
        for (int j = tid; j < nq_half; j += blockDim.x)
          qlsp[j] = pow(10.0, (j == nq_half - 1) ? 0.0 : (double)j * da + a0);                     // :211
      }
      __syncthreads();
      // xm = unique(lsp U 1 - lsp) (+ 0 and 1, which are already members): merge by rank, duplicates kept as
      // zero-width intervals
      for (int r = tid; r < NQ; r += blockDim.x) {
        int rank;
        double v;
        if (r < nq_half) {
          v = qlsp[r];
          int lo = 0, hi = nq_half;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((1.0 - qlsp[mid]) < v) hi = mid; else lo = mid + 1;
          }
          rank = r + (nq_half - lo);
        } else {
          const int j0 = r - nq_half;
          v = 1.0 - qlsp[j0];
          int lo = 0, hi = nq_half;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (qlsp[mid] <= v) lo = mid + 1; else hi = mid;
          }
          rank = (nq_half - 1 - j0) + lo;
        }
        qx[rank] = v;
      }
      __syncthreads();
      const double inv_t0 = 1.0 / t0;
      for (int j = tid; j < Eu; j += blockDim.x) {
        const double te = us[j], span = te - tb;
        int idx = 0, idx_slope = -1;
        double slope = 0.0, acc = 0.0;
        for (int m = 0; m < NQ; ++m) {
          const double x = qx[m];
          const double t = tb + span * x;                                                         // :217
          while (idx < G - 2 && tg[idx + 1] <= t) ++idx;
          if (idx != idx_slope) { slope = (ld[idx + 1] - ld[idx]) / (tg[idx + 1] - tg[idx]); idx_slope = idx; }
          const double lum = slope * (t - tg[idx]) + ld[idx];                                     // np.interp, :220
          const double arg = (t - te) * inv_t0;
          double g = lum * ((arg > -700.0) ? exp_tab(arg, tab) : exp(arg));                       // :221
          if (isnan(g)) g = 0.0;                                                                  // :222
          acc = fma(g, qx[min(m + 1, NQ - 1)] - qx[max(m - 1, 0)], acc);
        }
        lq[j] = (0.5 * span * acc) * inv_t0;                                                      // :224-225
      }
    } else {
    // ---- knots = unique(dense ∪ epochs) (interaction_processes.py:181-183): rank = own index + number of elements
    // of the other list sorting before it; an epoch equal to a dense point follows it as a zero-length interval -----
    for (int k = tid; k < G; k += blockDim.x) {
      const double x = tg[k];
      int lo = 0, hi = Eu;                 // number of epochs < x
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (us[mid] < x) lo = mid + 1; else hi = mid;
      }
      mt[k + lo] = x;
      ml[k + lo] = ld[k];
    }
    for (int j = tid; j < Eu; j += blockDim.x) {
      const double x = us[j];
      int lo = 0, hi = G;                  // number of dense points <= x
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tg[mid] <= x) lo = mid + 1; else hi = mid;
      }
      const int kk = min(max(lo - 1, 0), G - 2);
      const double slope = (ld[kk + 1] - ld[kk]) / (tg[kk + 1] - tg[kk]);       // np.interp
      mt[j + lo] = x;
      ml[j + lo] = slope * (x - tg[kk]) + ld[kk];
      upos[j] = j + lo;
    }
    __syncthreads();
    const int M = G + Eu;
    // ---- affine maps of the intervals (interaction_processes.py:185-195) ---------------------------------------------
    {
      const double t0 = sc[CS_T0];
      for (int i = tid; i < M - 1; i += blockDim.x) {
        const double dt = mt[i + 1] - mt[i];
        double av = 1.0, bv = 0.0;
        if (dt != 0.0) {
          const double slope = (ml[i + 1] - ml[i]) / dt;
          const double sdt = dt / t0;
          av = (sdt < 700.0) ? exp_tab(-sdt, tab) : exp(-sdt);
          const double om = -expm1_fast(-sdt, tab);
          bv = ml[i] * om + slope * (dt - t0 * om);
        }
        ca[i] = av;
        cb[i] = bv;
      }
    }
    __syncthreads();
    // ---- block scan of the maps: per-thread chunk, warp-shuffle scan of the chunk aggregates ------------------------
    {
      const int chunk = (M - 1 + (int)blockDim.x - 1) / (int)blockDim.x;
      const int i0 = min(tid * chunk, M - 1), i1 = min(i0 + chunk, M - 1);
      double A = 1.0, B = 0.0;
      for (int i = i0; i < i1; ++i) {
        B = B * ca[i] + cb[i];
        A *= ca[i];
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double pa = __shfl_up_sync(0xffffffffu, A, o), pb = __shfl_up_sync(0xffffffffu, B, o);
        if (lane >= o) { B = A * pb + B; A = A * pa; }
      }
      if (lane == 31) { wsa[warp] = A; wsb[warp] = B; }
      double ea = __shfl_up_sync(0xffffffffu, A, 1), eb = __shfl_up_sync(0xffffffffu, B, 1);   // lane-exclusive
      if (lane == 0) { ea = 1.0; eb = 0.0; }
      __syncthreads();
      if (warp == 0) {
        double WA = (lane < nw) ? wsa[lane] : 1.0, WB = (lane < nw) ? wsb[lane] : 0.0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double pa = __shfl_up_sync(0xffffffffu, WA, o), pb = __shfl_up_sync(0xffffffffu, WB, o);
          if (lane >= o) { WB = WA * pb + WB; WA = WA * pa; }
        }
        if (lane < nw) { wsa[lane] = WA; wsb[lane] = WB; }
      }
      __syncthreads();
      const double wb = (warp > 0) ? wsb[warp - 1] : 0.0;     // state entering this warp (new_lums[0] = 0)
      double v = ea * wb + eb;                                 // state entering this thread's chunk
      if (tid == 0) ml[0] = 0.0;
      for (int i = i0; i < i1; ++i) {
        v = v * ca[i] + cb[i];
        ml[i + 1] = v;
      }
    }
    }
    if (nickel) {
      // ---- csm_nickel: Diffusion (interaction_processes.py:33-65) of the nickel-cobalt engine on the adaptive grid,
      // ejecta mass mej + mass_csm_threshold (supernova_models.py:2148-2157) ------------------------------------------------
      const int64_t S = a.param_stride;
      const double mej_eff = a.params[RB_SN_MEJ * S + n] + sc[CS_MTH] / kMsun;
      const double vkm = a.params[RB_SN_VEJ * S + n];
      const double tau = sqrt(kDiffusionConst * a.params[RB_SN_KAPPA * S + n] * mej_eff / vkm) / kDay;            // :39
      const double trap = (kTrappingConst * a.params[RB_SN_KAPPA_GAMMA * S + n] * mej_eff / (vkm * vkm)) / (kDay * kDay);
      const double t_end = tg[G - 1], tb = fmax(0.0, tg[0]);
      double* nlsp = nk;               // [50]
      double* nx = nk + kNodesHalf;    // [100]
      if (tid < kNodesHalf) {
        const double a0 = log10(tau / t_end) - 3.0;                                               // :50
        const double da = (0.0 - a0) / (double)(kNodesHalf - 1);
        nlsp[tid] = pow(10.0, (tid == kNodesHalf - 1) ? 0.0 : (double)tid * da + a0);
      }
      __syncthreads();
      if (tid < kNodes) {
        int rank;
        double v;
        if (tid < kNodesHalf) {
          v = nlsp[tid];
          int lo = 0, hi = kNodesHalf;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((1.0 - nlsp[mid]) < v) hi = mid; else lo = mid + 1;
          }
          rank = tid + (kNodesHalf - lo);
        } else {
          const int j0 = tid - kNodesHalf;
          v = 1.0 - nlsp[j0];
          int lo = 0, hi = kNodesHalf;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (nlsp[mid] <= v) lo = mid + 1; else hi = mid;
          }
          rank = (kNodesHalf - 1 - j0) + lo;
        }
        nx[rank] = v;
      }
      __syncthreads();
      const double inv_tau2 = 1.0 / (tau * tau);
      for (int j = tid; j < Eu; j += blockDim.x) {
        const double te = us[j], span = te - tb;
        const double te_n = fmin(fmax(tb + span * nx[kNodes - 1], tb), t_end);                     // int_times[:, -1], :53-55
        const double te2 = te_n * te_n;
        int idx = 0, idx_slope = -1;
        double slope = 0.0, acc = 0.0;
        for (int m = 0; m < kNodes; ++m) {
          const double t = fmin(fmax(tb + span * nx[m], tb), t_end);                              // :53
          while (idx < G - 2 && tg[idx + 1] < t) ++idx;                                           // interp1d: x_lo < t <= x_hi
          if (idx != idx_slope) { slope = (ldn[idx + 1] - ldn[idx]) / (tg[idx + 1] - tg[idx]); idx_slope = idx; }
          const double lum = slope * (t - tg[idx]) + ldn[idx];                                    // :56
          const double arg = (__dmul_rn(t, t) - te2) * inv_tau2;                                  // :57
          double g = lum * t * ((arg > -700.0) ? exp_tab(arg, tab) : exp(arg));
          if (isnan(g)) g = 0.0;                                                                  // :58
          const double tp = fmin(fmax(tb + span * nx[min(m + 1, kNodes - 1)], tb), t_end);
          const double tm = fmin(fmax(tb + span * nx[max(m - 1, 0)], tb), t_end);
          acc = fma(g, tp - tm, acc);
        }
        lni[j] = (0.5 * acc) * (-2.0 * expm1(-trap / te2) * inv_tau2);                            // :60-61
      }
    }
    __syncthreads();
    // ---- per-epoch epilogue -----------------------------------------------------------------------------------------
    double logl_part = 0.0;
    for (int e = tid; e < E; e += blockDim.x) {
      const int j = c.uidx[e];
      double lbol = quad ? lq[j] : ml[upos[j]];                // np.interp(time, knots, new_lums) at a knot
      if (nickel) lbol += lni[j];                              // supernova_models.py:2165
      const double te = us[j];
      double model;
      if (sed == RB_SED_BOLOMETRIC) {
        model = lbol;
      } else {
        double tph, rph;
        temperature_floor_dev(lbol, te, sc[CS_RCV], sc[CS_TFLOOR], sc[CS_INV_REC], tph, rph);
        if (a.grid_mode) {
          double* go = a.grid_out + (size_t)n * 4 * E;
          go[0 * E + e] = tph;
          go[1 * E + e] = rph;
          go[2 * E + e] = lbol;
          go[3 * E + e] = a.epoch_tab[EP_T * E + e] * opz;                       // time_observer_frame (:2096)
          continue;
        }
        model = sed_mjy(a.cfg, lbol, te, tph, rph, a.epoch_tab[EP_NU * E + e], opz, sc[CS_INV_DEN], sc[CS_DL], tab);
      }
      if (a.out_model != nullptr) a.out_model[n * (int64_t)E + e] = model;
      if (want_logl)
        logl_part += logl_term(a.cfg.sigma_mode, a.epoch_tab[EP_KIND * E + e], a.epoch_tab[EP_Y * E + e], model,
                               a.epoch_tab[EP_ISIG * E + e], a.epoch_tab[EP_LOGT * E + e], sc[CS_SIGMA]);
    }
    if (want_logl) {
      logl_part = warp_sum(logl_part);
      __syncthreads();
      if (lane == 0) wsa[warp] = logl_part;
      __syncthreads();
      if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < nw; ++w) total += wsa[w];
        a.out_logl[n] = nan_to_num(total);
      }
    }
  }
}

}  // namespace rb
