// redback_b200 -- shared device helpers (constants, exp, reductions, cosmology quadrature).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "redback_b200.h"

namespace rb {

// ---- constants: redback/constants.py:1-23 (astropy CODATA-2018 cgs), SURVEY.md §8 a20 ----------
constexpr double kC = 29979245800.0;
constexpr double kCSI = 299792458.0;
constexpr double kH = 6.62607015e-27;
constexpr double kKB = 1.380649e-16;
constexpr double kSigmaSB = 5.6703744191844314e-05;
constexpr double kMsun = 1.988409870698051e33;
constexpr double kKm = 100000.0;
constexpr double kDay = 86400.0;
constexpr double kAngstrom = 1e-08;
constexpr double kPi = 3.141592653589793;

// interaction_processes.py:36-37
constexpr double kDiffusionConst = 2.0 * kMsun / (13.7 * kC * kKm);
constexpr double kTrappingConst = 3.0 * kMsun / (4 * kPi * (kKm * kKm));
// photosphere.py:58-59
constexpr double kRadiusConst = kKm * kDay;
constexpr double kStefConst = 4 * kPi * kSigmaSB;
// sed.py:303-304
constexpr double kXConst = kH * kC / kKB;
constexpr double kFluxConst = 4 * kPi * 2 * kPi * kH * (kC * kC) * kAngstrom;
// erg s^-1 cm^-2 Hz^-1 -> mJy
constexpr double kToMJy = 1e26;

// ---- table-driven exp -------------------------------------------------------------------------
// exp(x) = 2^m * 2^(j/1024) * exp(r), k = rint(x*1024/ln2) = 1024 m + j, |r| <= ln2/2048 = 3.4e-4.
// `tab` points at a shared-memory copy of RB_EXP2_1024 (8 KB).  Valid for -708 < x < 709 (no denormal /
// overflow handling: callers clamp).
//   exp_tab  : degree-3 expm1 polynomial (truncation 5.5e-16) + two-word ln2/1024: ~1 ulp, used for the
//              engine luminosities and anywhere full accuracy matters.
//   The diffusion hot loop inlines a cheaper variant (degree 2, error 6.5e-12; see sn_kernel.cu).
constexpr int kExpTabSize = 1024;
constexpr double kExpMagic = 6755399441055744.0;  // 1.5 * 2^52
constexpr double kTabOverLn2 = 1477.3197218702985;   // 1024 / ln 2
constexpr double kLn2OverTabHi = 0.0006769015435155716;
constexpr double kLn2OverTabLo = 2.2658131339052534e-20;

__device__ __forceinline__ double exp_tab(double x, const double* __restrict__ tab) {
  const double v = fma(x, kTabOverLn2, kExpMagic);
  const int k = __double2loint(v);
  const double kd = v - kExpMagic;
  double r = fma(kd, -kLn2OverTabHi, x);
  r = fma(kd, -kLn2OverTabLo, r);
  double q = fma(r, 0.16666666666666666, 0.5);
  q = fma(r, q, 1.0);
  const double p = r * q;
  const double t = tab[k & 1023];
  const double res = fma(t, p, t);
  return __hiloint2double(__double2hiint(res) + ((k >> 10) << 20), __double2loint(res));
}

// ---- reductions -------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; result valid in every thread.  `scratch` has >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double s = (lane < nw) ? scratch[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) scratch[32] = s;
  }
  __syncthreads();
  return scratch[32];
}

// np.argmin over the warp's lanes for values that are >= 0, +inf, or -inf (the caller's stand-in for NaN, which np.argmin
// returns first): the smallest value and, among equals, the smallest index, left in every lane.  Non-negative doubles
// order like their bit patterns, so three integer warp reductions (REDUX.MIN on the high word, the low word among the
// lanes that tie on it, the index among those that tie on both) replace five shuffle stages of (value, index) pairs.
__device__ __forceinline__ void warp_argmin_first(double& dist, int& idx) {
  const unsigned long long key = (dist < 0.0) ? 0ull : (unsigned long long)__double_as_longlong(dist) + 1ull;
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned m_hi = __reduce_min_sync(0xffffffffu, hi);
  const bool c1 = hi == m_hi;
  const unsigned m_lo = __reduce_min_sync(0xffffffffu, c1 ? lo : 0xffffffffu);
  const bool c2 = c1 && lo == m_lo;
  idx = (int)__reduce_min_sync(0xffffffffu, c2 ? (unsigned)idx : 0xffffffffu);
  const unsigned long long kmin = ((unsigned long long)m_hi << 32) | m_lo;
  dist = (kmin == 0ull) ? -INFINITY : __longlong_as_double((long long)(kmin - 1ull));
}

// np.maximum / Python max(a, b) on floats: NaN-propagating maximum
__device__ __forceinline__ double fmax_nan(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  return (b > a) ? b : a;
}

// np.nan_to_num for a float64 scalar (likelihoods.py:227)
__device__ __forceinline__ double nan_to_num(double v) {
  if (isnan(v)) return 0.0;
  if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
  return v;
}

// ---- Gaussian log-likelihood term of one datum (likelihoods.py:229-231) for every noise model -----------------
//   RB_SIGMA_FIXED       sigma given: `isig` = 1/sigma, `logt` = log(2 pi sigma^2)/2 precomputed per epoch
//   RB_SIGMA_SAMPLED     sigma = sampled parameter `sp`                                   (likelihoods.py:179-185)
//   RB_SIGMA_QUADRATURE  sqrt(sigma_i^2 + sp^2)                                           (likelihoods.py:767)
//   RB_SIGMA_FRACTIONAL  sqrt(sigma_i^2 * model^2)                                        (likelihoods.py:826)
//   RB_SIGMA_SYSTEMATIC  sqrt(sigma_i^2 + model^2 * sp^2)                                 (likelihoods.py:888)
// For the modes other than FIXED `isig` carries sigma_i itself.
// kind != 0: the datum is an upper limit (GaussianLikelihoodWithUpperLimits._upper_limit_log_likelihood,
// likelihoods.py:377-419): `isig` = 1/sigma_measurement; kind 1 = "true value < limit" (flux-like), 2 = magnitudes.
__device__ __forceinline__ double logl_term(int mode, double kind, double y, double model, double isig, double logt,
                                            double sp) {
  const double res = y - model;
  if (kind != 0.0) {
    double cdf = 0.5 * (1.0 + erf((res * isig) / 1.4142135623730951));
    if (kind == 2.0) cdf = 1.0 - cdf;
    cdf = fmin(fmax(cdf, 1e-30), 1.0 - 1e-30);   // np.clip(cdf, eps, 1 - eps), NaN propagates
    return log(cdf);
  }
  if (mode == RB_SIGMA_FIXED) {
    const double rs = res * isig;
    return -(rs * rs) / 2 - logt;
  }
  double sg = sp;
  if (mode == RB_SIGMA_QUADRATURE) sg = sqrt(isig * isig + sp * sp);
  else if (mode == RB_SIGMA_FRACTIONAL) sg = sqrt((isig * isig) * (model * model));
  else if (mode == RB_SIGMA_SYSTEMATIC) sg = sqrt(isig * isig + (model * model) * (sp * sp));
  const double rs = res / sg;
  return -(rs * rs) / 2 - log(2 * kPi * (sg * sg)) / 2;
}

// ---- cosmology --------------------------------------------------------------------------------
// 1/E(z) of flat LambdaCDM with massive neutrinos (astropy FLRW.nu_relative_density / inv_efunc).
__device__ __forceinline__ double inv_efunc(const rb_cosmology& c, double z) {
  const double zp1 = 1.0 + z;
  double rel = (double)c.n_massless;
  for (int i = 0; i < c.n_massive; ++i) {
    double cur = c.nu_y[i] / zp1;
    rel += pow(1.0 + pow(0.3173 * cur, 1.83), 0.54644808743);
  }
  const double nu = 0.22710731766 * c.neff_per_nu * rel;
  const double Or = c.Ogamma0 * (1.0 + nu);
  return rsqrt(zp1 * zp1 * zp1 * (Or * zp1 + c.Om0) + c.Ode0);
}

}  // namespace rb
