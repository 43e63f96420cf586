// redback_b200 -- redback-native structured-jet afterglow (tophat / Gaussian) + Gaussian log-likelihood.
//
// Reference: redback/transient_models/afterglow_models/base_models.py
//   constants :32-39 (the module's own, not astropy's)        get_segments_numba :43-64
//   erf_approximation_numba :67-89   get_structure_numba :92-148   get_obsangle_numba :151-179
//   RK4_step_numba :182-192          get_gamma_numba :195-249
//   calc_afterglow_step1_numba :252-322   calc_afterglow_step2_numba :325-370   get_ag_numba :373-399
//   RedbackAfterglows.get_lightcurve / calc_afterglow_numba / calc_lightcurve :573-640
//   tophat_redback :885-946   gaussian_redback :948-1010
//
// The jet is discretised into `res` rings x N_rot cells; work per sample is data dependent (res in
// [10, 100]).  Samples are processed in chunks; per chunk:
//   ag_prep_kernel   one warp per sample: d_L (32-point Gauss-Legendre), derived scalars, res -> ring count
//   ag_scan_kernel   exclusive scan of ring counts -> ring offsets (one block)
//   ag_ring_kernel   one THREAD per (sample, ring): RK4 of Gamma(log m) over `steps` steps in registers and
//                    the ring's synchrotron parameters (step1) streamed to a ring-major scratch buffer
//   ag_cell_kernel   one BLOCK per sample, one WARP per cell: each lane owns a contiguous run of steps; the
//                    observer-time cumsums are warp-shuffle scans; per-step spectra go to shared memory and
//                    the lanes then interpolate them at the epochs (np.interp) into per-warp accumulators
//                    (deterministic summation order: cells within a warp in sequence, then warps in order).
#include "rb_common.cuh"

namespace rb {

// base_models.py:32-39
constexpr double AG_MP = 1.6726231e-24;
constexpr double AG_ME = 9.1093897e-28;
constexpr double AG_CC = 2.99792453e10;
constexpr double AG_QE = 4.8032068e-10;
constexpr double AG_C2 = AG_CC * AG_CC;
constexpr double AG_FOURPI = 4 * kPi;
// SIGT = (QE*QE/(ME*C2))**2 * (8*pi/3): the square is a Python float pow; value computed on the host
// with the same operations (see ag_sigt()) and passed in AgConst.

constexpr int kAgMaxNu = 16;     // unique frequencies per call (the per-warp spectra buffer grows with them)
constexpr int kAgWarps = 5;      // warps per block in ag_cell_kernel
constexpr int kAgCellsPerUnit = 100;  // cells handled by one block of ag_cell_kernel: res = 10 -> 10 rings per unit,
                                      // res = 100 -> one ring per unit (equal work per block whatever the resolution)
__host__ __device__ inline int ag_rings_per_unit(int nrot) { return nrot >= kAgCellsPerUnit ? 1 : max(1, kAgCellsPerUnit / max(nrot, 1)); }
// per ring and step, everything that does not depend on the cell's azimuth (written by ag_ring_kernel, read by every
// cell of the ring straight from L1 / L2, lane = step):
//   RA_RD    R[s] - R[s-1]            (np.diff(R, prepend=0), :347-350)
//   RA_IBETA 1 / beta                 RA_BETA, RA_G: the Doppler factor needs them separately (:343)
//   RA_OMG   ring solid angle (:305-306), RA_OMCOS = omega * cos(acos(1 - omega / rotstep)) (:351-352)
//   RA_FMR   Ne * P'  (F_max without the cell's solid angle, Doppler factor and distance, :365-366)
//   RA_FBR   2 kT R^2 / c^2           (black-body limit without solid angle, Doppler factor and distance, :367)
//   RA_NUMP, RA_NUCP comoving nu_m, nu_c and their logarithms RA_LNUMP, RA_LNUCP.  nu_c (:363-364) depends on the ring
//   only: the on-axis observer time it needs is accumulated in the ring kernel's own sequential loop (= np.cumsum order)
enum { RA_RD = 0, RA_IBETA, RA_BETA, RA_G, RA_OMG, RA_OMCOS, RA_FMR, RA_FBR, RA_NUMP, RA_NUCP, RA_LNUMP, RA_LNUCP,
       kRingArrays };

enum {
  AS_OPZ = 0, AS_DL, AS_THV, AS_N, AS_EB, AS_EE, AS_P, AS_XP, AS_FX, AS_XIN, AS_G0, AS_EK, AS_THC, AS_THJ,
  AS_LATSTEP, AS_ROTSTEP, AS_RES, AS_NROT, AS_SIGMA, AS_PAD, kAgScalars  // = 20
};

struct AgArgs {
  rb_ag_config cfg;
  int64_t N;            // samples in this chunk
  int64_t n_base;       // index of the chunk's first sample in the full batch
  int64_t N_total;
  int E;
  const double* params; // [RB_AG_NPARAM][N_total]
  const double* dl_cm;
  const double* sigma_param;
  double* out_model;
  double* out_logl;
  const double* epoch_tab;   // [kEpochCols][E]
  // epochs in time order, `epl` consecutive ones per lane, stored transposed (entry k of lane l at k * 32 + l):
  const double* ep_t_T;      // [32 * epl] observer time [days]
  const int* ep_nu_T;        // [32 * epl] index into nu_unique, -1 = padding
  const int* ep_orig_T;      // [32 * epl] index of the epoch in the caller's order, -1 = padding
  int epl;
  double t_epoch_min, t_epoch_max;   // smallest / largest (non-NaN) observer time of the call [days]
  double nu_unique[kAgMaxNu];
  int n_nu;
  double* scalars;           // [N][kAgScalars]
  int* ring_count;           // [N]
  int64_t* ring_off;         // [N + 1]
  double* ring_scratch;      // [rings][kRingArrays][steps]
  int* unit_count;           // [N]   work units (groups of rings, about kAgCellsPerUnit cells) per sample
  int64_t* unit_off;         // [N + 1]
  double* partial;           // [units][E] per-unit light-curve partial sums
  double sigt;
  int want_logl;
};

__device__ const double kPxfP[12] = {1.0, 1.2, 1.4, 1.6, 1.8, 2.0, 2.2, 2.5, 2.7, 3.0, 3.2, 3.4};
__device__ const double kPxfX[12] = {3.0, 1.4, 1.1, 0.86, 0.725, 0.637, 0.579, 0.520, 0.487, 0.451, 0.434, 0.420};
__device__ const double kPxfF[12] = {0.41, 0.44, 0.48, 0.53, 0.56, 0.59, 0.612, 0.630, 0.641, 0.659, 0.660, 0.675};

// np.interp(p, xp, fp) for p inside the table (base_models.py:584-585)
__device__ __forceinline__ double pxf_interp(const double* fp, double p) {
  if (p <= kPxfP[0]) return fp[0];
  if (p >= kPxfP[11]) return fp[11];
  int j = 0;
  while (j < 10 && kPxfP[j + 1] <= p) ++j;
  const double slope = (fp[j + 1] - fp[j]) / (kPxfP[j + 1] - kPxfP[j]);
  return slope * (p - kPxfP[j]) + fp[j];
}

// ---- per-sample scalars -----------------------------------------------------------------------------------
__global__ void ag_prep_kernel(const AgArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const int64_t NT = a.N_total;
  const double* P = a.params + a.n_base + n;
  const double z = P[RB_AG_REDSHIFT * NT];
  const double opz = 1.0 + z;
  double dl;
  if (a.dl_cm != nullptr) {
    dl = a.dl_cm[a.n_base + n];
  } else {
    double s = RB_GL32_W[lane] * inv_efunc(a.cfg.cosmology, 0.5 * z * (RB_GL32_X[lane] + 1.0));
    s = warp_sum(s);
    dl = opz * a.cfg.cosmology.hubble_distance_cm * (0.5 * z * s);
  }
  if (lane != 0) return;
  const double thv = P[RB_AG_THV * NT], thc = P[RB_AG_THC * NT], g0 = P[RB_AG_G0 * NT];
  const double thj = (a.cfg.jet != RB_AG_TOPHAT) ? P[RB_AG_THJ * NT] : thc;       // tophat: thj = thc (:937)
  const double p = P[RB_AG_P * NT];
  double nism = pow(10.0, P[RB_AG_LOGN0 * NT]);
  if (a.cfg.k == 2) nism = nism * 3e35;                                            // :540-541
  // resolution rule (:932-935)
  int res = a.cfg.res;
  if (res <= 0) {
    const double sep = fmax(thv - thc, 0.0);
    const double ord = (2 - 10 * sep) * thc * g0;
    int order = (ord >= 2147483647.0) ? 2147483647 : (ord <= -2147483648.0 ? -2147483647 - 1 : (int)ord);
    order = min(order, 100);
    res = max(10, order);
  }
  const double resd = (double)res;
  const double latstep = thj / resd;                                               // :46-49
  const double rotstep = 2.0 * kPi / resd;
  const int nrot = (int)(2 * kPi / rotstep);
  double* S = a.scalars + n * kAgScalars;
  S[AS_OPZ] = opz;
  S[AS_DL] = dl;
  S[AS_THV] = thv;
  S[AS_N] = nism;
  S[AS_EB] = pow(10.0, P[RB_AG_LOGEPSB * NT]);
  S[AS_EE] = pow(10.0, P[RB_AG_LOGEPSE * NT]);
  S[AS_P] = p;
  S[AS_XP] = pxf_interp(kPxfX, p);
  S[AS_FX] = pxf_interp(kPxfF, p);
  S[AS_XIN] = P[RB_AG_XIN * NT];
  S[AS_G0] = g0;
  S[AS_EK] = pow(10.0, P[RB_AG_LOGE0 * NT]);
  S[AS_THC] = thc;
  S[AS_THJ] = thj;
  S[AS_LATSTEP] = latstep;
  S[AS_ROTSTEP] = rotstep;
  S[AS_RES] = resd;
  S[AS_NROT] = (double)nrot;
  S[AS_SIGMA] = (a.sigma_param != nullptr && a.want_logl) ? a.sigma_param[a.n_base + n] : 0.0;
  S[AS_PAD] = 0.0;
  // p outside (1.2, 3.4) raises ValueError in the reference (:576-577): flagged by zero rings -> NaN output
  const int rings = (p < 1.2 || p > 3.4 || !(p == p)) ? 0 : res;
  a.ring_count[n] = rings;
  const int rpu = ag_rings_per_unit(nrot);
  a.unit_count[n] = (rings + rpu - 1) / rpu;
}

// exclusive scan of ring_count[0..N) -> ring_off[0..N]; single block
__global__ void ag_scan_kernel(int64_t N, const int* __restrict__ cnt, int64_t* __restrict__ off) {
  __shared__ int64_t wsum[33];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int64_t base = 0; base < N; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    int64_t v = (i < N) ? cnt[i] : 0;
    const int64_t own = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int64_t w = (lane < nw) ? wsum[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      if (lane < nw) wsum[lane] = w;
    }
    __syncthreads();
    const int64_t incl = carry + (warp > 0 ? wsum[warp - 1] : 0) + v;
    if (i < N) off[i] = incl - own;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[N] = carry;
}

// x ** (1 / (3 - k)) for the two media the reference allows (k = 0: cube root, k = 2: identity); numpy's pow with the
// rounded exponent 1/3 differs from the cube root by ln(x) * 1.9e-17 relative
__device__ __forceinline__ double pow_inv3k(double x, int k, double inv3k) {
  return (k == 0) ? cbrt(x) : ((k == 2) ? x : pow(x, inv3k));
}

// RK4_step_numba (:182-192)
__device__ __forceinline__ double ag_rk4(double ghat, double dm, double g, double m, double fac, double therm) {
  const double ghatm1 = ghat - 1.0;
  const double dm10 = exp10(dm);   // 10**dm
  const double g2 = g * g;
  const double ig = 1.0 / g;
  return fac * dm10 * (ghat * (g2 - 1.0) - ghatm1 * (g - ig)) /
         (m + dm10 * (therm + (1.0 - therm) * (2.0 * ghat * g - ghatm1 * (1 + 1.0 / g2))));
}

// ring r -> (sample n, latitude index i, ring centre thi)
struct AgRing {
  int64_t n;
  int i;
  double thi;
};
__device__ __forceinline__ AgRing ag_ring_of(const AgArgs& a, int64_t r) {
  int64_t lo = 0, hi = a.N;      // last n with ring_off[n] <= r
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (a.ring_off[mid] <= r) lo = mid; else hi = mid;
  }
  AgRing g;
  g.n = lo;
  g.i = (int)(r - a.ring_off[lo]);
  const double* S = a.scalars + lo * kAgScalars;
  const double latstep = S[AS_LATSTEP];
  const int nlat = (int)S[AS_RES];
  // thi = linspace(latstep - latstep/2, nlat*latstep - latstep/2, nlat)[i]  (:61)
  const double th_lo = latstep - latstep / 2., th_hi = (double)nlat * latstep - latstep / 2.;
  if (nlat == 1) g.thi = th_lo;
  else if (g.i == nlat - 1) g.thi = th_hi;
  else g.thi = (double)g.i * ((th_hi - th_lo) / (double)(nlat - 1)) + th_lo;
  return g;
}

// mass grid of get_gamma_numba (:195-249): log10 of the first swept-up mass, the step in log10 m and -h ln 10
__device__ __forceinline__ void ag_mass_grid(double n_amb, double kk, int steps, double& dlogm, double& h, double& fac) {
  const double Nmin = (AG_FOURPI / (3.0 - kk)) * n_amb * pow(1e10, 3.0 - kk);
  const double Nmax = (AG_FOURPI / (3.0 - kk)) * n_amb * pow(1e24, 3.0 - kk);
  dlogm = log10(Nmin * AG_MP);
  h = log10(Nmax / Nmin) / (double)steps;
  fac = -h * log(10.0);
}

// adiabatic index fit (:200-206)
__device__ __forceinline__ double ag_ghat(double G) {
  const double g2m1 = G * G - 1.0;
  const double root = sqrt(g2m1);
  const double theta = root * (root + 1.07 * g2m1) / (3.0 * (1 + root + 1.07 * g2m1));
  const double zz = theta / (0.24 + theta);
  return ((((((1.07136 * zz - 2.39332) * zz + 2.32513) * zz - 0.96583) * zz + 0.18203) * zz - 1.21937) * zz + 5.0) / 3.0;
}

// ---- one THREAD per ring: jet structure and the sequential RK4 march of Gamma(log m) (get_gamma_numba) -----------
// Writes the Lorentz factor and log10 m BEFORE each update (what step1 sees) into the ring's RA_G / RA_RD slots; the
// latter is overwritten by ag_ring_kernel.
__global__ void __launch_bounds__(128) ag_gamma_kernel(const AgArgs a, int64_t total_rings) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= total_rings) return;
  const AgRing ring = ag_ring_of(a, r);
  const double* S = a.scalars + ring.n * kAgScalars;
  const int steps = a.cfg.steps;
  const double kk = (double)a.cfg.k;
  const double thi = ring.thi;
  // jet structure (:92-112)
  const double g0 = S[AS_G0], ek = S[AS_EK], thc = S[AS_THC], thj = S[AS_THJ];
  double fac_s;
  if (a.cfg.jet == RB_AG_TOPHAT) {
    const double x = -(thi - fmin(thc, thj)) * 1000.0;
    const double sign = (x >= 0) ? 1.0 : -1.0;
    const double xa = fabs(x);
    const double t = 1.0 / (1.0 + 0.3275911 * xa);
    const double yv = 1.0 - (((((1.061405429 * t + -1.453152027) * t) + 1.421413741) * t + -0.284496736) * t +
                             0.254829592) * t * exp(-xa * xa);
    fac_s = (sign * yv) * 0.5 + 0.5;
  } else {
    const double q = thi / thc;
    fac_s = exp(-0.5 * (q * q));
  }
  double Gs = (g0 - 1) * fac_s + 1.0000000000001;
  double Ei = ek * fac_s;
  const double ss = a.cfg.ss, aa = a.cfg.aa;
  if (a.cfg.jet == RB_AG_POWERLAW) {                         // :114-120
    Gs = g0;
    Ei = ek;
    if (thi >= thc) {
      Ei = ek * pow(thc / thi, ss);
      Gs = (g0 - 1.0) * pow(thc / thi, aa) + 1.000000000000001;
    }
  } else if (a.cfg.jet == RB_AG_POWERLAW_ALT) {              // :122-125
    const double q = thi / thc;
    const double fac = sqrt(1 + q * q);
    Gs = 1.000000000001 + (g0 - 1.0) * pow(fac, -aa);
    Ei = ek * pow(fac, -ss);
  } else if (a.cfg.jet == RB_AG_TWO_COMPONENT) {             // :127-131
    Gs = g0;
    Ei = ek;
    if (thi > thc) { Gs = aa; Ei = ek * ss; }
  } else if (a.cfg.jet == RB_AG_DOUBLE_GAUSSIAN) {           // :133-146
    const double qc = thi / thc, qj = thi / thj;
    const double ec = exp(-0.5 * (qc * qc)), ej = exp(-0.5 * (qj * qj));
    Ei = ek * ((1.0 - ss) * ec + ss * ej);
    const double temp = ((1.0 - ss) * ec + ss * ej) / ((1.0 - ss / aa) * ec + (ss / aa) * ej);
    Gs = isfinite(temp) ? (g0 - 1.0) * temp + 1.0000000000001 : aa;
  }
  double dlogm, h, fac;
  ag_mass_grid(S[AS_N], kk, steps, dlogm, h, fac);
  const double M = Ei / (Gs * AG_C2);
  double* out = a.ring_scratch + (size_t)r * kRingArrays * steps;
  double G = Gs, dm = dlogm;
  for (int s = 0; s < steps; ++s) {
    out[RA_G * steps + s] = G;
    out[RA_RD * steps + s] = dm;
    const double ghat = ag_ghat(G);
    // ---- RK4 (:238-246) ----
    const double F1 = ag_rk4(ghat, dm, G, M, fac, 0.0);
    const double F2 = ag_rk4(ghat, dm + 0.5 * h, G + 0.5 * F1, M, fac, 0.0);
    const double F3 = ag_rk4(ghat, dm + 0.5 * h, G + 0.5 * F2, M, fac, 0.0);
    const double F4 = ag_rk4(ghat, dm + h, G + F3, M, fac, 0.0);
    G = G + (F1 + 2.0 * (F2 + F3) + F4) / 6.0 + 1e-15;
    dm = dm + h;
  }
}

// inclusive warp scan of v with the running total `carry` of the previous 32-step chunks
__device__ __forceinline__ double ag_scan(double v, double& carry, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += up;
  }
  v += carry;
  carry = __shfl_sync(0xffffffffu, v, 31);
  return v;
}

// ---- one WARP per ring, lane = step: calc_afterglow_step1 (:252-322) and the azimuth-independent part of step2 --------
// Everything except the two running sums (radius R = cumsum of the increments, on-axis observer time) is local to a
// step once Gamma is known; the sums are warp scans per chunk of 32 steps with carried totals (np.cumsum).
__global__ void __launch_bounds__(128) ag_ring_kernel(const AgArgs a, int64_t total_rings) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= total_rings) return;
  const AgRing ring = ag_ring_of(a, r);
  const double* S = a.scalars + ring.n * kAgScalars;
  const int steps = a.cfg.steps;
  const double kk = (double)a.cfg.k;
  const double thi = ring.thi;
  const double latstep = S[AS_LATSTEP], rotstep = S[AS_ROTSTEP], resd = S[AS_RES];
  const double n_amb = S[AS_N];
  const double p = S[AS_P], xp = S[AS_XP], Fx = S[AS_FX], EB = S[AS_EB], Ee = S[AS_EE], xiN = S[AS_XIN];
  const bool expansion = a.cfg.expansion != 0;
  const double a1 = a.cfg.a1;
  const double hfac = 0.5 * latstep;
  const double cos_lo = cos(thi - hfac);
  const double omg_noexp = rotstep * (cos_lo - cos(thi + hfac));
  const double inv3k = 1.0 / (3.0 - kk);
  const double sigt = a.sigt;
  double* out = a.ring_scratch + (size_t)r * kRingArrays * steps;

  double ex0 = 1.0, r_orig_last = 0.0, r_cum_last = 0.0, carry_r = 0.0, carry_t = 0.0;
  for (int sb = 0; sb < steps; sb += 32) {
    const int s = sb + lane;
    const bool live = s < steps;
    const int sl = live ? s : steps - 1;
    const double G = out[RA_G * steps + sl], dm = out[RA_RD * steps + sl];
    const double ghat = ag_ghat(G);
    const double Gm1 = G - 1.0, G2 = G * G;
    const double beta = sqrt(1 - 1.0 / G2);
    const double Ne = exp10(dm) / AG_MP;
    const double cs = sqrt(AG_C2 * ghat * (ghat - 1) * Gm1 / (1 + ghat * Gm1));
    const double te = asin(cs / (AG_CC * sqrt(G2 - 1.0)));
    double ex = 1.0, omg = omg_noexp;
    if (expansion) {
      ex = te / pow(G, a1 + 1);
      omg = rotstep * (cos_lo - cos(ex / resd + thi + hfac));
    }
    if (sb == 0) ex0 = __shfl_sync(0xffffffffu, ex, 0);
    const double r_orig = pow_inv3k((3.0 - kk) * Ne / (AG_FOURPI * n_amb), a.cfg.k, inv3k);
    double r_prev = __shfl_up_sync(0xffffffffu, r_orig, 1);
    if (lane == 0) r_prev = r_orig_last;
    r_orig_last = __shfl_sync(0xffffffffu, r_orig, 31);
    double r_inc = r_orig;
    if (s > 0) {
      const double expo = sqrt((1 - cos(latstep + ex0)) / (1 - cos(latstep + ex)));
      r_inc = (r_orig - r_prev) * pow_inv3k(expo, a.cfg.k, inv3k);
    }
    const double r_cum = ag_scan(live ? r_inc : 0.0, carry_r, lane);
    double r_before = __shfl_up_sync(0xffffffffu, r_cum, 1);
    if (lane == 0) r_before = r_cum_last;
    r_cum_last = __shfl_sync(0xffffffffu, r_cum, 31);
    const double r_diff = (s == 0) ? r_cum : r_cum - r_before;                       // np.diff(R, prepend=0), :347-350
    const double ibeta = 1.0 / beta;
    const double tobso = ag_scan(live ? r_diff * (ibeta - 1.0) / AG_CC : 0.0, carry_t, lane);   // :353, :356
    if (!live) continue;
    const double n0 = (a.cfg.k == 0) ? n_amb : n_amb * pow(r_cum, -kk);             // r**-0.0 == 1
    const double B = sqrt(2 * AG_FOURPI * EB * n0 * AG_MP * AG_C2 * ((ghat * G + 1.0) / (ghat - 1.0)) * Gm1);
    double gm;
    if (p > 2) {
      gm = ((p - 2) / (p - 1)) * (Ee / xiN) * Gm1 * (AG_MP / AG_ME);
    } else {
      const double gmm = sqrt(1.5 * AG_FOURPI * AG_QE / (sigt * B));
      if (p == 2) gm = (1 / log(gmm / ((Ee / xiN) * Gm1 * (AG_MP / AG_ME)))) * (Ee / xiN) * Gm1 * (AG_MP / AG_ME);
      else gm = pow((2 - p) / (p - 1) * (AG_MP / AG_ME) * (Ee / xiN) * Gm1 * pow(gmm, p - 2), 1 / (p - 1));
    }
    const double gc = 6.0 * kPi * AG_ME * AG_CC / (G * sigt * B * B * tobso);        // :360
    const double nucp = 0.286 * 3.0 * gc * gc * AG_QE * B / (AG_FOURPI * AG_ME * AG_CC);          // :361
    const double nump = 3.0 * xp * gm * gm * AG_QE * B / (AG_FOURPI * AG_ME * AG_CC);
    const double pp = xiN * Fx * AG_ME * AG_C2 * sigt * B / (3.0 * AG_QE);
    const double kt = gm * AG_ME * AG_C2;
    out[RA_RD * steps + s] = r_diff;
    out[RA_IBETA * steps + s] = ibeta;
    out[RA_BETA * steps + s] = beta;
    out[RA_OMG * steps + s] = omg;
    out[RA_OMCOS * steps + s] = omg * (1.0 - omg / rotstep);       // cos(arccos(x)) = x to an ulp (x in [-1, 1], a factor)
    out[RA_FMR * steps + s] = Ne * pp;
    out[RA_FBR * steps + s] = 2.0 * kt * r_cum * r_cum / AG_C2;
    out[RA_NUMP * steps + s] = nump;
    out[RA_NUCP * steps + s] = nucp;
    out[RA_LNUMP * steps + s] = log(nump);
    out[RA_LNUCP * steps + s] = log(nucp);
  }
}

// get_ag_numba (:373-399): the overlapping if-chain reduces to the last matching condition
__device__ __forceinline__ double ag_flux(double fbb, double nuc, double num, double nu1, double fmax, double p) {
  double f = 0.0;
  if (nuc < num) {
    if (num < nu1) f = fmax * (1.0 / sqrt(num / nuc)) * pow(nu1 / num, -p / 2.0);
    else if (nuc < nu1) f = fmax * (1.0 / sqrt(nu1 / nuc));
    else if (nu1 < nuc) f = fmax * cbrt(nu1 / nuc);
  } else if (num < nuc) {
    if (nuc < nu1) f = fmax * pow(nuc / num, -(p - 1.0) / 2.0) * pow(nu1 / nuc, -p / 2.0);
    else if (num < nu1) f = fmax * pow(nu1 / num, -(p - 1.0) / 2.0);
    else if (nu1 < num) f = fmax * cbrt(nu1 / num);
  }
  const double adj = fbb * (nu1 * nu1) * fmax_nan(1.0, sqrt(nu1 / num));
  return (f < adj) ? f : adj;  // min(FBB_adj, Fluxt): first argument unless the second is smaller
}

// The same with the power laws evaluated as exp(exponent * (log a - log b)): the logarithms of nu_c and nu_m are shared
// by all frequencies of a step and log(nu) is a per-sample constant, so a step costs two log and one exp per frequency
// instead of two to four pow.  Comparisons use the original values; |exponent * log| < 1e3 keeps the result within
// 1e-13 relative of the pow form.  Callers route non-positive / non-finite nu_c, nu_m to ag_flux.
__device__ __forceinline__ double ag_flux_log(double fbb, double nuc, double num, double nu1, double lnc, double lnm,
                                              double ln1, double fmax, double p, double nu1_sq, double sqrt_nu1_over_num) {
  double f = 0.0;
  if (nuc < num) {
    if (num < nu1) f = fmax * exp(-0.5 * (lnm - lnc) - 0.5 * p * (ln1 - lnm));
    else if (nuc < nu1) f = fmax * exp(-0.5 * (ln1 - lnc));
    else if (nu1 < nuc) f = fmax * exp((ln1 - lnc) * (1.0 / 3.0));
  } else if (num < nuc) {
    if (nuc < nu1) f = fmax * exp(-0.5 * (p - 1.0) * (lnc - lnm) - 0.5 * p * (ln1 - lnc));
    else if (num < nu1) f = fmax * exp(-0.5 * (p - 1.0) * (ln1 - lnm));
    else if (nu1 < num) f = fmax * exp((ln1 - lnm) * (1.0 / 3.0));
  }
  const double adj = fbb * nu1_sq * fmax_nan(1.0, sqrt_nu1_over_num);
  return (f < adj) ? f : adj;
}

// RB_PREC_F32: exp(x) through the single-precision hardware exponential on an FP64-reduced argument: x = k ln2 + r with
// |r| <= 0.35 in FP64, e^r in FP32 (relative error ~2e-7), 2^k into the exponent field.
__device__ __forceinline__ double exp_sp(double x) {
  if (!(x > -708.0)) return (x == x) ? 0.0 : x;
  if (x > 708.0) return INFINITY;
  const double v = fma(x, 1.4426950408889634, 6755399441055744.0);
  const int k = __double2loint(v);
  const double r = fma(v - 6755399441055744.0, -0.6931471805599453, x);
  const double res = (double)expf((float)r);
  return __hiloint2double(__double2hiint(res) + (k << 20), __double2loint(res));
}

// ag_flux_log with exp_sp (RB_PREC_F32)
__device__ __forceinline__ double ag_flux_log_sp(double fbb, double nuc, double num, double nu1, double lnc, double lnm,
                                                 double ln1, double fmax, double p, double nu1_sq,
                                                 double sqrt_nu1_over_num) {
  double f = 0.0;
  if (nuc < num) {
    if (num < nu1) f = fmax * exp_sp(-0.5 * (lnm - lnc) - 0.5 * p * (ln1 - lnm));
    else if (nuc < nu1) f = fmax * exp_sp(-0.5 * (ln1 - lnc));
    else if (nu1 < nuc) f = fmax * exp_sp((ln1 - lnc) * (1.0 / 3.0));
  } else if (num < nuc) {
    if (nuc < nu1) f = fmax * exp_sp(-0.5 * (p - 1.0) * (lnc - lnm) - 0.5 * p * (ln1 - lnc));
    else if (num < nu1) f = fmax * exp_sp(-0.5 * (p - 1.0) * (ln1 - lnm));
    else if (nu1 < num) f = fmax * exp_sp((ln1 - lnm) * (1.0 / 3.0));
  }
  const double adj = fbb * nu1_sq * fmax_nan(1.0, sqrt_nu1_over_num);
  return (f < adj) ? f : adj;
}

// ---- one block per work unit (about kAgCellsPerUnit cells: a few rings of one sample), one warp per cell -------------------------
// Warps are independent (no block barrier inside the ring loop): the ring arrays are read straight from the scratch
// buffer (coalesced, lane = step; the ~24 KB of a ring stay in L1 while its cells are processed), only the per-warp
// observer-time / spectrum arrays and light-curve accumulators live in shared memory.
// Epochs are handed to the lanes in TIME ORDER, a contiguous run of `epl` epochs per lane (tables built by the host,
// transposed so that entry k of lane l sits at k * 32 + l): within a run the np.interp interval index only moves
// forward, so after one binary search per lane the search is a short linear walk.
// SP: RB_PREC_F32 (its own instantiation: the FP64 path carries none of it)
#ifndef RB_AG_CELL_MINB
#define RB_AG_CELL_MINB 4   // resident blocks per SM the register cap is derived from (A/B switch)
#endif
template <bool SP>
__global__ void __launch_bounds__(kAgWarps * 32, RB_AG_CELL_MINB) ag_cell_kernel(const AgArgs a, int64_t total_units) {
  extern __shared__ double asm_[];
  const int steps = a.cfg.steps, E = a.E, n_nu = a.n_nu, epl = a.epl, EP = 32 * a.epl;
  constexpr bool fp32 = SP;
  double* wbuf = asm_;                                           // per warp: tobs[steps] + flux[n_nu][steps]
  double* acc = wbuf + (size_t)kAgWarps * (1 + n_nu) * steps;    // [kAgWarps][EP]  (position = k * 32 + lane)
  double* sc = acc + (size_t)kAgWarps * EP;                      // [kAgScalars]
  double* nutab = sc + kAgScalars;                               // [4][kAgMaxNu]: nu (1+z), its log, sqrt, square

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t u = blockIdx.x;
  if (u >= total_units) return;
  // unit -> sample: last n with unit_off[n] <= u
  int64_t lo_n = 0, hi_n = a.N;
  while (hi_n - lo_n > 1) {
    const int64_t mid = (lo_n + hi_n) >> 1;
    if (a.unit_off[mid] <= u) lo_n = mid; else hi_n = mid;
  }
  const int64_t n = lo_n;
  const int rpu = ag_rings_per_unit((int)a.scalars[n * kAgScalars + AS_NROT]);
  const int ring_begin = (int)(u - a.unit_off[n]) * rpu;
  const int ring_end = min(ring_begin + rpu, a.ring_count[n]);

  double* my_tobs = wbuf + (size_t)warp * (1 + n_nu) * steps;
  double* my_flux = my_tobs + steps;
  double* my_acc = acc + (size_t)warp * EP;
  if (tid < kAgScalars) sc[tid] = a.scalars[n * kAgScalars + tid];
  if (tid < kAgMaxNu) {
    const double nu1 = (tid < n_nu) ? a.nu_unique[tid] * a.scalars[n * kAgScalars + AS_OPZ] : 1.0;   // nu = nu0*(1+z) :587
    nutab[tid] = nu1;
    nutab[kAgMaxNu + tid] = log(nu1);
    nutab[2 * kAgMaxNu + tid] = sqrt(nu1);
    nutab[3 * kAgMaxNu + tid] = nu1 * nu1;
  }
  for (int e = tid; e < kAgWarps * EP; e += blockDim.x) acc[e] = 0.0;
  __syncthreads();
  {
    const int nrot = (int)sc[AS_NROT];
    const double opz = sc[AS_OPZ], dl = sc[AS_DL], rotstep = sc[AS_ROTSTEP], latstep = sc[AS_LATSTEP];
    const double thv = sc[AS_THV], p = sc[AS_P];
    const double inv_dl2 = 1.0 / (dl * dl);
    const double inv_cc = 1.0 / AG_CC;
    const bool expansion = a.cfg.expansion != 0;
    const double sin_tho = sin(thv), cos_tho = cos(thv);
    const int nlat = (int)sc[AS_RES];
    const double th_lo = latstep - latstep / 2., th_hi = (double)nlat * latstep - latstep / 2.;
    const double ph_lo = rotstep - rotstep / 2., ph_hi = (double)nrot * rotstep - rotstep / 2.;
    const double day_over_opz = kDay / opz;
    const double x_epoch_min = a.t_epoch_min * day_over_opz, x_epoch_max = a.t_epoch_max * day_over_opz;

    for (int i = ring_begin; i < ring_end; ++i) {
      const double* __restrict__ ring = a.ring_scratch + (size_t)(a.ring_off[n] + i) * kRingArrays * steps;
      double thi;
      if (nlat == 1) thi = th_lo;
      else if (i == nlat - 1) thi = th_hi;
      else thi = (double)i * ((th_hi - th_lo) / (double)(nlat - 1)) + th_lo;
      const double sin_thi = sin(thi), cos_thi = cos(thi);
      // Omi row factor (:57): (cos(i*latstep) - cos((i+1)*latstep))
      const double dcos = cos((double)i * latstep) - cos((double)(i + 1) * latstep);

      // geometry of this warp's cells (j = warp, warp + 5, ...), 32 at a time with lane = cell: the trigonometric
      // functions of get_obsangle_numba run once per cell instead of once per lane of the cell's warp
      double cosa_l = 0.0, om0_l = 0.0;
      for (int j = warp, q = 0; j < nrot; j += kAgWarps, ++q) {
        if ((q & 31) == 0) {
          const int jl = j + kAgWarps * lane;
          // phi = linspace(rotstep, nrot*rotstep, nrot)[j]; phii = linspace(rotstep/2, ...)[j]   (:53,62)
          double phi, phii;
          if (nrot == 1) { phi = rotstep; phii = ph_lo; }
          else {
            const double phi_hi = (double)nrot * rotstep;
            phi = (jl == nrot - 1) ? phi_hi : (double)jl * ((phi_hi - rotstep) / (double)(nrot - 1)) + rotstep;
            phii = (jl == nrot - 1) ? ph_hi : (double)jl * ((ph_hi - ph_lo) / (double)(nrot - 1)) + ph_lo;
          }
          om0_l = (phi - (phi - rotstep)) * dcos;
          // get_obsangle_numba (:151-179) with phi = 0
          const double f1 = 0.0 * sin_tho * sin(phii) * sin_thi;
          const double f2 = 1.0 * sin_tho * cos(phii) * sin_thi;
          double ang = cos_tho * cos_thi + f2 + f1;
          ang = (ang > 1.0) ? 1.0 : ((ang < -1.0) ? -1.0 : ang);
          cosa_l = cos(acos(ang));
        }
        const double Om0 = __shfl_sync(0xffffffffu, om0_l, q & 31);
        const double cosa = __shfl_sync(0xffffffffu, cosa_l, q & 31);
        // the cell's own solid angle (expansion: used until the ring's exceeds it)
        const double omcos0 = Om0 * (1.0 - Om0 / rotstep);
        const double c_fmax = Om0 / AG_FOURPI / AG_FOURPI * inv_dl2;       // NO * P' * dop^3 / (4 pi dl^2), :365-366
        // ---- step2 (:325-370).  Lane l owns steps l, l + 32, ...
        // Pass 1: observer times of all steps (cumsum = warp scan per chunk of 32 steps, running total carried between
        // chunks), and how many of them lie at or below the first / above the last epoch.
        double carry = 0.0;
        int n_le = 0, n_gt = 0;
        for (int sb = 0; sb < steps; sb += 32) {
          const int s = sb + lane;
          const bool live = s < steps;
          const int sc_ = live ? s : steps - 1;
          double tobs = live ? __ldg(ring + RA_RD * steps + sc_) * (__ldg(ring + RA_IBETA * steps + sc_) - cosa) * inv_cc
                             : 0.0;                                                     // :353-356
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, tobs, o);
            if (lane >= o) tobs += up;
          }
          tobs += carry;
          carry = __shfl_sync(0xffffffffu, tobs, 31);
          if (live) my_tobs[s] = tobs;
          n_le += __popc(__ballot_sync(0xffffffffu, live && tobs <= x_epoch_min));
          n_gt += __popc(__ballot_sync(0xffffffffu, live && tobs > x_epoch_max));
        }
        // np.interp only reads the spectra at the steps bracketing an epoch -- from the last step at or below the first
        // epoch to the first step above the last one -- and at the two ends (epochs outside the observer-time range).
        // For typical cells that is a fifth of the 250 steps (the step grid spans 14 decades in radius, the epochs 3.5 in
        // time): the spectra of the others are never evaluated.  A NaN observer time counts on neither side, which
        // keeps the whole range.
        const int s_lo = max(n_le - 1, 0), s_hi = min(steps - n_gt, steps - 1);
        __syncwarp();
        // Pass 2: spectra of the needed steps
        for (int sb = 0; sb < steps; sb += 32) {
          const int s = sb + lane;
          const bool need = s < steps && ((s >= s_lo && s <= s_hi) || s == 0 || s == steps - 1);
          if (!__any_sync(0xffffffffu, need)) continue;
          if (!need) continue;
          const double beta = __ldg(ring + RA_BETA * steps + s), G = __ldg(ring + RA_G * steps + s);
          const double nump = __ldg(ring + RA_NUMP * steps + s), nucp = __ldg(ring + RA_NUCP * steps + s);
          const double fmr = __ldg(ring + RA_FMR * steps + s), fbr = __ldg(ring + RA_FBR * steps + s);
          const double lnump = __ldg(ring + RA_LNUMP * steps + s), lnucp = __ldg(ring + RA_LNUCP * steps + s);
          double omcos = __ldg(ring + RA_OMCOS * steps + s);
          if (expansion) {                                                              // :345-352
            const double omg = __ldg(ring + RA_OMG * steps + s);
            const double Om = fmax_nan(Om0, omg);
            if (!(Om == omg)) omcos = omcos0;
          }
          const double dop = 1.0 / (G * (1.0 - beta * cosa));                           // :343
          const double num = dop * nump;
          const double nuc = dop * nucp;
          const double Fmax = c_fmax * fmr * (dop * dop * dop);
          const double FBB = omcos * dop * fbr * inv_dl2;                               // :367
          const bool regular = num > 0.0 && num < INFINITY && nuc > 0.0 && nuc < INFINITY;
          if (regular) {
            // log(dop * nu') = log(dop) + log(nu'): one logarithm per cell step, the ring's are tabulated
            const double ld = fp32 ? (double)logf((float)dop) : log(dop);
            const double lnm = ld + lnump, lnc = ld + lnucp;
            const double rs = rsqrt(num);
            if (fp32) {
              for (int h = 0; h < n_nu; ++h)
                my_flux[h * steps + s] = ag_flux_log_sp(FBB, nuc, num, nutab[h], lnc, lnm, nutab[kAgMaxNu + h], Fmax, p,
                                                        nutab[3 * kAgMaxNu + h], nutab[2 * kAgMaxNu + h] * rs);
            } else {
              for (int h = 0; h < n_nu; ++h)
                my_flux[h * steps + s] = ag_flux_log(FBB, nuc, num, nutab[h], lnc, lnm, nutab[kAgMaxNu + h], Fmax, p,
                                                     nutab[3 * kAgMaxNu + h], nutab[2 * kAgMaxNu + h] * rs);
            }
          } else {
            for (int h = 0; h < n_nu; ++h) my_flux[h * steps + s] = ag_flux(FBB, nuc, num, nutab[h], Fmax, p);
          }
        }
        __syncwarp();
        // ---- calc_lightcurve (:632-640): np.interp(t/(1+z), tobs, Flux[h]) summed over cells ----
        const double x_first = my_tobs[0], x_last = my_tobs[steps - 1];
        int lo = -1;                     // largest j <= steps - 2 with tobs[j] <= x, carried along the lane's time-ordered run
        int lo_slope = -1;               // interval whose 1 / (tobs[lo+1] - tobs[lo]) is in `inv_dt`
        double inv_dt = 0.0;
        // epochs observed at the same time in several bands sit next to each other in the run: the interval search and
        // the abscissa part of the interpolation are done once per distinct time (mode 0 interior, 1 / 2 clamped to the
        // first / last step, 3 NaN)
        double x_prev = NAN, dx = 0.0;
        int mode = 3;
        double t_next = a.ep_t_T[lane];
        int h_next = a.ep_nu_T[lane];
        for (int k = 0; k < epl; ++k) {
          const int pos = k * 32 + lane;
          const int hh = h_next;
          const double x = t_next * day_over_opz;                        // :923  time * day_to_s / (1 + z)
          if (k + 1 < epl) { t_next = a.ep_t_T[pos + 32]; h_next = a.ep_nu_T[pos + 32]; }
          if (hh < 0) continue;          // padding
          const double* fp = my_flux + hh * steps;
          if (!(x == x_prev)) {
            x_prev = x;
            if (!(x == x)) mode = 3;
            else if (x <= x_first) mode = 1;        // numpy: x < xp[0] -> left; x == xp[0] -> fp[0]
            else if (x >= x_last) mode = 2;
            else {
              mode = 0;
              int walk = 0;
              if (lo >= 0)
                while (walk < 4 && lo < steps - 2 && my_tobs[lo + 1] <= x) { ++lo; ++walk; }
              if (lo < 0 || walk == 4) {
                int hi = steps - 1;
                lo = max(lo, 0);
                while (hi - lo > 1) {
                  const int mid = (lo + hi) >> 1;
                  if (my_tobs[mid] <= x) lo = mid; else hi = mid;
                }
              }
              const double t0 = my_tobs[lo];
              if (lo != lo_slope) { inv_dt = 1.0 / (my_tobs[lo + 1] - t0); lo_slope = lo; }
              dx = x - t0;
            }
          }
          double v;
          if (mode == 0) {
            const double f0 = fp[lo];
            v = fma((fp[lo + 1] - f0) * inv_dt, dx, f0);                 // slope * (x - xp[lo]) + fp[lo]
          } else if (mode == 1) v = fp[0];
          else if (mode == 2) v = fp[steps - 1];
          else v = x;
          my_acc[pos] += v;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // per-unit partial light curve (deterministic: warps summed in order)
    for (int q = tid; q < EP; q += blockDim.x) {
      const int e = a.ep_orig_T[q];
      if (e < 0) continue;
      double lc = 0.0;
      for (int w = 0; w < kAgWarps; ++w) lc += acc[w * EP + q];
      a.partial[u * (int64_t)E + e] = lc;
    }
  }
}

// ---- one block per sample: sum the sample's unit partials in order, (1+z)/1e-26 -> mJy, likelihood ------------
__global__ void __launch_bounds__(128) ag_finish_kernel(const AgArgs a) {
  __shared__ double wsum[4];
  const int E = a.E, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    const int64_t u0 = a.unit_off[n], u1 = a.unit_off[n + 1];
    const double opz = a.scalars[n * kAgScalars + AS_OPZ];
    const double sp = a.scalars[n * kAgScalars + AS_SIGMA];
    double logl_part = 0.0;
    for (int e = tid; e < E; e += blockDim.x) {
      double lc = 0.0;
      for (int64_t u = u0; u < u1; ++u) lc += a.partial[u * (int64_t)E + e];
      double model_v = (u1 > u0) ? lc * opz / 1e-26 : NAN;                         // :640, :942
      // astropy: (f * mJy).to(ABmag) = -2.5 log10(f * 1e-26 / AB), AB = 10^(48.6 / -2.5) erg s^-1 cm^-2 Hz^-1
      if (a.cfg.ab_magnitude) model_v = -2.5 * log10(model_v * 2.754228703338175e-07);
      if (a.out_model != nullptr) a.out_model[(a.n_base + n) * (int64_t)E + e] = model_v;
      if (a.want_logl) {
        logl_part += logl_term(a.cfg.sigma_mode, a.epoch_tab[EP_KIND * E + e], a.epoch_tab[EP_Y * E + e], model_v,
                               a.epoch_tab[EP_ISIG * E + e],
                               a.epoch_tab[EP_LOGT * E + e], sp);
      }
    }
    if (a.want_logl) {
      logl_part = warp_sum(logl_part);
      __syncthreads();
      if (lane == 0) wsum[warp] = logl_part;
      __syncthreads();
      if (tid == 0) a.out_logl[a.n_base + n] = nan_to_num(wsum[0] + wsum[1] + wsum[2] + wsum[3]);
    }
  }
}

}  // namespace rb
