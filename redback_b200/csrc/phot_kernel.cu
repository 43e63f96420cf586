// redback_b200 -- spectral grid -> bicubic spline -> bandpass synthetic photometry -> magnitudes (+ log L).
//
// Reference path (output_format = 'magnitude' | 'flux'):
//   blackbody_to_spectrum / flux_density_to_spectrum             redback/sed.py:911-968
//   get_correct_output_format_from_spectra                       redback/sed.py:971-1009
//   RedbackTimeSeriesSource (sncosmo.TimeSeriesSource: scipy RectBivariateSpline(phase, wave, flux, kx=3, ky=3),
//     i.e. FITPACK regrid with s = 0: the interpolating not-a-knot bicubic tensor spline)   redback/sed.py:120-195
//   _bandflux_single_redback / _bandflux_redback / _bandmag_redback                        redback/sed.py:10-118
//   bandpass_magnitude_to_flux                                   redback/utils.py:707-718
//
// The spline value at (epoch e, wavelength w) is  b_t(e)^T A_t^-1 S A_lambda^-T b_lambda(w)  with S the N_t x N_lambda
// spectra, A_t / A_lambda the B-spline collocation matrices and b the four non-zero B-splines.  The two solves commute,
// so the N_t x N_lambda coefficient matrix is never formed: one persistent block per sample streams S through shared
// memory in BLOCKS OF W WAVELENGTH COLUMNS,
//   1. spectra of the block (all N_t rows, W columns; non-finite / zero entries replaced, sed.py:985-992);
//   2. time-axis solve of the W columns in shared memory: every column is cut into P row segments (P a power of two
//      <= 32, W * P = threads).  ONE thread owns a (column, segment) through the whole solve: forward recurrence from a
//      zero state, the true incoming boundary from a P-lane warp scan of the segments' affine maps (their 2 x 2 matrices
//      depend on the LU only and are tabulated per sample), backward recurrence, reverse scan, and the final fix-up with
//      the segment's homogeneous solutions -- no block barrier inside the solve;
//   3. contraction with the E epochs' time B-splines (E x W values) and, because the blocks come in increasing
//      wavelength order, the forward elimination of the WAVELENGTH-axis solve on those E rows (state carried per epoch);
//      the eliminated values go to a per-block scratch of E x (N_lambda - first band column) doubles (L2-resident);
// then per epoch: wavelength-axis back substitution (thread per epoch) and the band integral.  The reference sums
// |f| over the band's 5-Angstrom grid (sed.py:36-37); where all of the band's spline coefficients are non-negative the
// spline is non-negative (B-splines are), |f| = f, and the integral is the dot product of the coefficients with
// per-band weights contracted on the host; otherwise the points are evaluated one by one, split over a warp's lanes.
// Columns more than `blue_margin` below the bluest band are not evaluated: the inverse of the wavelength-axis collocation
// matrix decays like rho^|i-j| (rho = 2 - sqrt(3) = 0.268 on a uniform grid, <= 0.268 q^2 on a grid whose neighbouring
// wavelengths differ by at most a factor q; prefactor < 2.5) and spectra grow towards the blue by at most lambda^-4 =
// q^4 per column, so the dropped columns change a band coefficient by < 2.5 x^M / (1 - x) relative to the spectrum at
// the band's blue edge, x = 0.268 q^6.  The host picks the smallest M with that bound below 1e-11 (29 columns on the
// default 100-point grid, 24 on the kilonova models' 200 points; a fixed 40 until late in round 2 cost 12-14 %), below
// 1e-15 with dust extinction (its 10-mag cap can leave far-UV columns 1e4 times less attenuated than the band).  The mean needed by the clean-up is taken in a statistics pre-pass over ALL
// columns (terms e^-50 below their row's largest are counted, not summed), skipped when a bound shows that every entry
// is finite and non-zero; the per-sample banded LU of the time axis (kilonova grids) runs on one thread underneath it.
// SINGLE-PASS MODE (plain blackbody / cutoff spectra): when a bound shows that every entry of the KEPT columns is valid
// in every valid row, the only replaced entries are whole rows (NaN / zero / infinite T or R: e.g. the kilonova grid
// points after exp(t^2 / tau^2) has overflowed).  The replacement value r then enters as  S = S0 + r m 1^T  (m = row
// mask), and because both solves are linear and B-splines sum to one, every coefficient of epoch e changes by the same
// r U_e with U_e = b_t(e)^T A_t^-1 m: one extra column through the time-axis solve.  The spectra are then evaluated ONCE
// (rows of m as zeros), their sum is accumulated on the way, the pre-pass shrinks to the columns below the kept ones
// (mostly counted, not evaluated), and r U_e is added in the epoch phase.  Without invalid rows nothing is replaced in
// the kept columns and the pre-pass is skipped altogether.
#include "rb_common.cuh"

namespace rb {

constexpr int kPhotMaxSpan = 96;     // max B-spline coefficients (along wavelength) overlapped by one band
constexpr int kEpCols = 12;          // per-epoch record in the block's scratch: h0..h3, first row, y1, y2, band flux, U_e,
                                     // Ga, Gb, s0 (the backward fix-up folded into the contraction, see 2b)
constexpr int kScanSteps = 5;        // log2(32)

enum { PH_SED_BLACKBODY = 1, PH_SED_CUTOFF = 2, PH_SED_CUTOFF_LINE = 3 };
enum { PH_T = 0, PH_R, PH_L, PH_X, kPhotGridCols };   // per-sample model grid: temperature, radius, L_bol, phase

// diagnostics (rb_phot_mode_counts): samples per clean-up mode -- 0 every entry valid, 1 single pass without invalid rows,
// 2 single pass with the row-mask column, 3 two passes; 4..7 why single-pass mode was refused (negative factors, Planck
// exponent bound of the kept columns, a zero factor, magnitude bounds)
__device__ unsigned long long g_phot_modes[8];

struct PhotArgs {
  int64_t N;
  int64_t n_base;
  int E;
  int n_t_max;            // allocated grid length per sample
  int n_lambda;
  int sed;                // PH_SED_*
  int output;             // RB_OUT_MAGNITUDE | RB_OUT_FLUX
  int sigma_mode;
  int want_logl;
  int common_time_lu;     // 1: time-axis LU supplied (grid identical up to scaling for all samples)
  int n_comp;             // blackbody components summed into one spectrum (multi-component kilonovae), else 1
  int W;                  // wavelength columns per block
  int P;                  // row segments per column (W * P <= blockDim.x)
  int nt_pad;             // odd leading dimension of the column buffer
  int j_store_lo;         // first wavelength column kept for the back substitution (= min band_jlo)
  int blue_margin;        // columns evaluated below j_store_lo (see the header comment)
  int j_store_hi;         // last wavelength coefficient used by any band (= max band_jlo + band_span - 1)
  double q_red;           // largest ratio of neighbouring wavelengths above j_store_hi (red-side cut, section 0c)
  int span_max;           // max band_span
  int e_pad;              // leading dimension of the per-block epoch arrays
  int64_t scratch_stride; // doubles of scratch per block
  int64_t grid_comp_stride;   // doubles between the model grids of consecutive components
  double cutoff_wavelength, absorption_index;
  double line_wavelength, line_width, line_time, line_duration, line_amplitude;   // sed.Line (type_1a)
  const double* grid;     // [N][kPhotGridCols][n_t_max]   (written by the model kernels in grid mode)
  const int* grid_len;    // [N] or nullptr (= n_t_max)
  const double* opz;      // [N] 1 + z
  const double* dl;       // [N]
  const double* sigma_param_s;  // [N] sampled sigma (already sliced to the chunk) or nullptr
  const double* t0_s;           // [N] explosion epochs of t0_base_model (sliced to the chunk) or nullptr
  const double* lambda_obs;     // [n_lambda]
  const double* lu_lambda;      // [n_lambda][5]: l2, l1, d, u1, u2
  const double* lu_time;        // [n_t_max][5] when common_time_lu
  const double* epoch_tab;      // [kEpochCols][E] (EP_T = observer-frame days)
  const int* band_of_epoch;     // [E]
  const int* band_off;          // [n_bands + 1] into the integration-grid arrays
  const int* band_jlo;          // [n_bands] first wavelength coefficient used by the band
  const int* band_span;         // [n_bands] number of coefficients used
  const double* band_scale;     // [n_bands] dwave / HC_ERG_AA
  const double* band_zp;        // [n_bands]
  const double* band_ref;       // [n_bands] reference flux ('flux' output) or nullptr
  const int* epoch_order;       // [E] epochs grouped by band
  // ragged mode (survey simulation: every transient has its own pointings), all nullptr otherwise:
  const double* t_ragged;       // [N_total][E] observer-frame phase of sample n's pointings (days)
  const int* band_ragged;       // [N_total][E] band index per pointing
  const int* epoch_count;       // [N_total] pointings of sample n (<= E); the rest of the row is padding
  const double* int_rec;        // [total][6] per integration point: the 4 non-zero wavelength B-splines, wave * trans,
                                //            index of the first of them relative to the band's first coefficient
  const double* band_weight;    // [n_bands][span_max] sum over the band's points of wave * trans * B_j(wave)
  double* scratch;              // [gridDim][scratch_stride]
  // extinction of the spectra per observer-frame wavelength (extinction_models.py:223-251), see ext_kernel.cu
  const rb_extinction* ext;     // device copy of the laws, or nullptr
  const double* ext_av_host_s;  // [N] host-galaxy A_V (sliced to the chunk)
  double* out_model;
  double* out_logl;
};

// shared-memory doubles: row factors, time LU + homogeneous solutions, column factors, wavelength LU, scan matrices,
// reduction scratch, exp table, column buffer (also the per-epoch coefficient buffer of the last phase)
__host__ __device__ inline size_t phot_fixed_doubles(int NT, int NL, int n_comp, bool line, bool tables_global) {
  return (size_t)(2 * n_comp + (line ? 2 : 0)) * NT + (tables_global ? 0 : 9 * (size_t)NT) + 4 * (size_t)NL +
         5 * (size_t)NL + 2 * kScanSteps * 32 * 4 + 128 + 256 + 6 * (size_t)((NT + 31) / 32) + kExpTabSize + 4;
}

// FITPACK fpbspl for k = 3: the four non-zero cubic B-splines at x, t[l] <= x < t[l+1] (0-based l)
__device__ __forceinline__ void fpbspl3(const double* __restrict__ t, int l, double x, double h[4]) {
  double hh[3];
  h[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= 3; ++j) {
#pragma unroll
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
#pragma unroll
    for (int i = 1; i <= j; ++i) {
      const int li = l + i, lj = li - j;
      const double f = hh[i - 1] / (t[li] - t[lj]);
      h[i - 1] = h[i - 1] + f * (t[li] - x);
      h[i] = f * (x - t[lj]);
    }
  }
}

// knot vector of the interpolating (not-a-knot) cubic spline through x[0..n): t[0..3] = x0, t[4..n-1] = x[2..n-3],
// t[n..n+3] = x[n-1]   (FITPACK fpregr/fpgrre with s = 0)
__device__ __forceinline__ double nak_knot(const double* __restrict__ x, int n, int i) {
  if (i < 4) return x[0];
  if (i >= n) return x[n - 1];
  return x[i - 2];
}

// Banded LU rows (l2, l1, d, u1, u2) are kept as (l2, l1, 1/d, u1/d, u2/d): a back-substitution step is then
// x = b/d - (u1/d) x[i+1] - (u2/d) x[i+2] with a single FMA on the serial dependency chain.
__device__ __forceinline__ void load_lu_row(double* __restrict__ dst, const double* __restrict__ src) {
  const double rd = 1.0 / src[2];
  dst[0] = src[0];
  dst[1] = src[1];
  dst[2] = rd;
  dst[3] = src[3] * rd;
  dst[4] = src[4] * rd;
}

// 1 / d to ~1 ulp without the IEEE division's slow path: hardware seed + two Newton steps.  d = 0 / inf / NaN give a
// non-finite or zero result, which the caller classifies as an invalid spectrum sample like the reference's inf / 0.
__device__ __forceinline__ double rcp_fast(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// exp(x) for -700 <= x <= 700 with the 1024-entry table and a degree-2 polynomial (truncation 6.5e-12 relative,
// argument reduction with the high word of ln2/1024 only: |x| 1477 * 2.3e-20 < 3e-14): six FP64 operations.
__device__ __forceinline__ double exp_lean(double x, const double* __restrict__ tab) {
  const double v = fma(x, kTabOverLn2, kExpMagic);
  const int k = __double2loint(v);
  const double r = fma(v - kExpMagic, -kLn2OverTabHi, x);
  const double p = r * fma(r, 0.5, 1.0);
  const double t = tab[k & 1023];
  const double res = fma(t, p, t);
  return __hiloint2double(__double2hiint(res) + ((k >> 10) << 20), __double2loint(res));
}

// expm1(x) for 0 < x <= 700 from the same table: e^x - 1 = T (1 + p) - 1 with T = 2^m 2^(j/1024) and p = e^r - 1 by a
// degree-4 polynomial (|r| <= 3.4e-4: truncation 4e-20), evaluated as fma(T, p, T - 1): T - 1 is exact for T in [1, 2),
// so there is no cancellation for small x (relative error < 3e-16 for x >= 1e-12).
__device__ __forceinline__ double expm1_tab(double x, const double* __restrict__ tab) {
  const double v = fma(x, kTabOverLn2, kExpMagic);
  const int k = __double2loint(v);
  const double r = fma(v - kExpMagic, -kLn2OverTabHi, x);
  double q = fma(r, 1.0 / 24.0, 1.0 / 6.0);
  q = fma(r, q, 0.5);
  q = fma(r, q, 1.0);
  const double p = r * q;
  const double t = tab[k & 1023];
  const double Ts = __hiloint2double(__double2hiint(t) + ((k >> 10) << 20), __double2loint(t));
  return fma(Ts, p, Ts - 1.0);
}

// General 1 / expm1(x) of the Planck function (any x, per lane).  The reference's expm1 overflows to inf (-> value 0 ->
// replaced) above 709.78, reproduced here; between 700 and there the result is subnormal and libm's exp is used.
__device__ __noinline__ double planck_inv_general(double x, const double* __restrict__ tab) {
  if (x > 40.0) {
    if (x > 709.782712893384) return 0.0;
    if (x > 700.0) return exp(-x);
    return exp_lean(-x, tab);
  }
  if (x >= 1e-12) return rcp_fast(expm1_tab(x, tab));
  return rcp_fast(expm1(x));
}

// Columns to keep beyond the reddest band's last coefficient (phot_kernel, section 0c); out of line: pow / log / ceil
// once per sample must not weigh on the kernel's register allocation.
__device__ __noinline__ int red_margin(double xe, double q, int NL) {
  const double xr = 0.268 * (q * q) * pow(q, fmax(xe - 4.0, 0.0));
  if (!(xr < 0.8)) return NL;
  return max(8, (int)ceil(log(1e-11 * (1.0 - xr) / 2.5) / log(xr)));
}

#ifdef RB_PHOT_PROFILE
// developer aid (tools/phot_phases.sh): block 0 prints the cycles it spent in each phase, summed over all its samples
#define PHOT_PROF_DECL() long long pt_[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long pa_ = 0; long long pn_ = 0;
#define PHOT_PROF_BEGIN() pa_ = clock64(); ++pn_;
#define PHOT_PROF_ACC(i) { __syncthreads(); const long long c_ = clock64(); pt_[i] += c_ - pa_; pa_ = c_; }
#define PHOT_PROF_END()
#define PHOT_PROF_REPORT()                                                                                            \
  if (blockIdx.x == 0 && tid == 0)                                                                                    \
    printf("phot phases [cycles per sample, %lld samples] setup %lld stats+lu %lld spectra %lld solve %lld contract "  \
           "%lld backsub %lld integrate %lld finish %lld\n", pn_, pt_[0] / pn_, pt_[1] / pn_, pt_[2] / pn_,            \
           pt_[3] / pn_, pt_[4] / pn_, pt_[5] / pn_, pt_[6] / pn_, pt_[7] / pn_);
#else
#define PHOT_PROF_DECL()
#define PHOT_PROF_BEGIN()
#define PHOT_PROF_ACC(i)
#define PHOT_PROF_END()
#define PHOT_PROF_REPORT()
#endif

// two adjacent doubles: one 16-byte load where the base is known to be 16-byte aligned (shared-memory tables)
template <bool SCALAR>
__device__ __forceinline__ double2 ld_pair(const double* __restrict__ p) {
  if (SCALAR) return make_double2(p[0], p[1]);
  return *reinterpret_cast<const double2*>(p);
}

// GT: the time-axis LU and its homogeneous solutions live in global memory (grids too long for shared memory)
// RED: the red-side column cut (section 0c) is compiled in; the host picks it where the grid extends well beyond the
// reddest band (the kilonova models' 200-point default), the other instantiation is the same code without it (the cut's
// few lines cost the 100-point Arnett case 2 % through the register allocation of the hot loops).
template <bool GT, bool RED = false>
__global__ void __launch_bounds__(256, 2) phot_kernel(const PhotArgs a) {
  extern __shared__ __align__(16) double psm[];
  const int NT = a.n_t_max, NL = a.n_lambda, E = a.E, NC = a.n_comp;
  const int T = (int)blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const bool line = a.sed == PH_SED_CUTOFF_LINE;
  const int W = a.W, P = a.P, NTP = a.nt_pad;
  double* gs = a.scratch + (size_t)blockIdx.x * a.scratch_stride;
  double* rowA = psm;                          // [NC][NT]  1 / (k_B T)  or 1 / T
  double* rowB = rowA + NC * NT;               // [NC][NT]  R^2          or R^2 * cutoff normalisation
  double* rowC = rowB + NC * NT;               // [NT]  1 - line amplitude           (line SED only)
  double* rowD = rowC + (line ? NT : 0);       // [NT]  scaled line amplitude        (line SED only)
  double* after_rows = rowD + (line ? NT : 0);
  // time-axis banded LU, scaled rows (l2, l1, 1/d, u1/d, u2/d), stored as THREE arrays so that a sweep of the segmented
  // solve fetches its two coefficients with one 16-byte load: [NT] pairs (l2, l1), [NT] pairs (u1/d, u2/d), [NT] 1/d;
  // the homogeneous solutions of the segmented solve likewise as [NT] pairs (forward 0, 1) and [NT] pairs (backward 0, 1).
  // Both bases are 16-byte aligned in shared memory; with the tables in global memory (GT) the loads stay scalar.
  double* lut = GT ? gs : after_rows + ((after_rows - psm) & 1);    // (no pointer <-> integer round trip: it would
                                                                     // turn every access into a generic load)
  double* hom = lut + 5 * NT + (NT & 1);
#define LUT0(i) lut[2 * (i)]
#define LUT1(i) lut[2 * (i) + 1]
#define LUT2(i) lut[4 * NT + (i)]
#define LUT3(i) lut[2 * NT + 2 * (i)]
#define LUT4(i) lut[2 * NT + 2 * (i) + 1]
  // The LU is first order away from the two not-a-knot rows (l2 != 0 only in row G-2, u2 != 0 only in row 1), and
  // set_segments keeps row G-2 off the start of its segment: one boundary value per direction crosses a segment border,
  // so there is ONE homogeneous solution per direction and the sweeps read l1 / u1 from dense copies (8-byte loads at the
  // conflict-free odd stride L; the 16-byte pair loads cost twice the shared-memory wavefronts, and the solve is bound by
  // that pipe: ~34 wavefronts per row and column before, 22 now).
#define HOMF(i) hom[(i)]
#define HOMB(i) hom[NT + (i)]
#define L1D(i) hom[2 * NT + (i)]
#define U1D(i) hom[3 * NT + (i)]
  double* colA = GT ? after_rows : hom + 4 * NT;   // [NL]  h nu         or hc / (k_B lambda)
  double* colB = colA + NL;                    // [NL]  everything else that depends on wavelength only
  double* colC = colB + NL;                    // [NL]  line profile in output units
  double* colD = colC + NL;                    // [NL]  2 pi h nu^3 (the reference's numerator, for its underflow point)
  double* lul = colD + NL;                     // [NL][5]  wavelength-axis banded LU, scaled rows
  double* mtab = lul + 5 * NL;                 // [2][kScanSteps][32][4] scan matrices (forward, backward)
  double* red = mtab + 2 * kScanSteps * 32 * 4;   // [64]
  double* red2 = red + 64;                     // [64]
  double* xfix = red2 + 64;                    // [256] per (column, segment) thread: the backward boundary value X1
  double* grp = xfix + 256;                    // [6][ngroups] bounds of rowA / rowB over each group of 32 rows
  double* tab = grp + 6 * ((NT + 31) / 32);    // [1024]
  double* buf = tab + kExpTabSize;             // [W][NTP] column block; later [epochs][span_max] coefficients
  double* ept = gs + (GT ? 9 * (size_t)NT + 2 : 0);   // [kEpCols][e_pad]
  double* ys = ept + (size_t)kEpCols * a.e_pad;       // [NL - j_store_lo][e_pad]
  const int EP = a.e_pad;
  const int lgP = 31 - __clz(P);
  const int j_first = max(0, a.j_store_lo - a.blue_margin);

  for (int i = tid; i < kExpTabSize; i += T) tab[i] = RB_EXP2_1024[i];
  for (int j = tid; j < NL; j += T) load_lu_row(lul + 5 * j, a.lu_lambda + 5 * j);
  int L = 0, Pe = 0;                                  // segment length (odd) and number of non-empty segments
  auto set_segments = [&](int G) {
    L = (G + P - 1) / P;
    if (L < 3) L = 3;
    L |= 1;                                           // odd: conflict-free shared-memory strides
    while (G % L == 2) L += 2;                        // row G-2 (the only one with l2 != 0) never starts a segment
    Pe = (G + L - 1) / L;
  };
  // Tables of the segmented solve for the current LU and segmentation.
  //  hom: forward  y[i] = -l2 y[i-2] - l1 y[i-1]   from (y[m-1], y[m-2]) = (1, 0) at the segment start m,
  //       backward x[i] = -u2' x[i+2] - u1' x[i+1] from (x[m'+1], x[m'+2]) = (1, 0) at the segment end m'
  //       (the responses to the second boundary value vanish, see HOMF above).
  //  mtab: segment s maps its incoming boundary state to its outgoing one, out = p_s + H_s in, with H_s made of hom at
  //       the segment's last (forward) / first (backward) two rows; step d of the log-step scan needs the product of
  //       the d matrices ending (forward) / starting (backward) at s -- built here by doubling.
  auto build_tables = [&](int G) {
    for (int t2 = tid; t2 < 2 * Pe; t2 += T) {
      const int s = t2 >> 1, which = t2 & 1;
      const int i_lo = s * L, i_hi = min(i_lo + L, G);
      double p1 = 1.0, p2 = 0.0;
      if (which == 0) {
        for (int i = i_lo; i < i_hi; ++i) {
          const double y = -LUT0(i) * p2 - LUT1(i) * p1;
          HOMF(i) = y;
          p2 = p1;
          p1 = y;
        }
      } else {
        for (int i = i_hi - 1; i >= i_lo; --i) {
          const double x = -LUT4(i) * p2 - LUT3(i) * p1;
          HOMB(i) = x;
          p2 = p1;
          p1 = x;
        }
      }
    }
    for (int i = tid; i < G; i += T) {
      L1D(i) = LUT1(i);
      U1D(i) = LUT3(i);
    }
    __syncthreads();
    if (tid < 64) {
      const int dir = tid >> 5, s = tid & 31;
      double* m0 = mtab + (size_t)dir * kScanSteps * 128 + 4 * s;
      double h00 = 0.0, h01 = 0.0, h10 = 0.0, h11 = 0.0;
      if (s < Pe) {
        const int i_lo = s * L, i_hi = min(i_lo + L, G);
        if (dir == 0) {          // (y[i_hi-1], y[i_hi-2]) from (y[i_lo-1], y[i_lo-2]); all but the last segment have >= 3 rows
          h00 = HOMF(i_hi - 1);
          if (i_hi - 2 >= i_lo) { h10 = HOMF(i_hi - 2); } else { h10 = 1.0; }
        } else {                 // (x[i_lo], x[i_lo+1]) from (x[i_hi], x[i_hi+1])
          h00 = HOMB(i_lo);
          if (i_lo + 1 < i_hi) { h10 = HOMB(i_lo + 1); } else { h10 = 1.0; }
        }
      }
      m0[0] = h00; m0[1] = h01; m0[2] = h10; m0[3] = h11;
    }
    __syncthreads();
    for (int st = 1; st < kScanSteps; ++st) {
      const int d = 1 << (st - 1);
      if (tid < 64) {
        const int dir = tid >> 5, s = tid & 31;
        const double* prev = mtab + (size_t)dir * kScanSteps * 128 + (size_t)(st - 1) * 128;
        double* cur = mtab + (size_t)dir * kScanSteps * 128 + (size_t)st * 128 + 4 * s;
        const int o = dir == 0 ? s - d : s + d;
        const double* A = prev + 4 * s;
        if (o >= 0 && o < 32) {
          const double* B = prev + 4 * o;
          cur[0] = A[0] * B[0] + A[1] * B[2];
          cur[1] = A[0] * B[1] + A[1] * B[3];
          cur[2] = A[2] * B[0] + A[3] * B[2];
          cur[3] = A[2] * B[1] + A[3] * B[3];
        } else {
          cur[0] = A[0]; cur[1] = A[1]; cur[2] = A[2]; cur[3] = A[3];
        }
      }
      __syncthreads();
    }
  };
  if (a.common_time_lu) {
    for (int i = tid; i < NT; i += T) {
      double row[5];
      load_lu_row(row, a.lu_time + 5 * i);
      LUT0(i) = row[0]; LUT1(i) = row[1]; LUT2(i) = row[2]; LUT3(i) = row[3]; LUT4(i) = row[4];
    }
    set_segments(NT);
    __syncthreads();
    build_tables(NT);
  }
  __syncthreads();

  PHOT_PROF_DECL();
  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    PHOT_PROF_BEGIN();
    const int G = a.grid_len ? a.grid_len[n] : NT;
    const double* gT = a.grid + (size_t)n * kPhotGridCols * NT + PH_T * NT;
    const double* gR = a.grid + (size_t)n * kPhotGridCols * NT + PH_R * NT;
    const double* gL = a.grid + (size_t)n * kPhotGridCols * NT + PH_L * NT;
    const double* gX = a.grid + (size_t)n * kPhotGridCols * NT + PH_X * NT;
    const double opz = a.opz[n], dl = a.dl[n];
    const double lam_c = a.cutoff_wavelength * kAngstrom;
    const bool ragged = a.t_ragged != nullptr;
    const int En = ragged ? a.epoch_count[a.n_base + n] : E;                       // epochs of this sample
    const double* t_row = ragged ? a.t_ragged + (size_t)(a.n_base + n) * E : a.epoch_tab + (size_t)EP_T * E;
    const int* band_row = ragged ? a.band_ragged + (size_t)(a.n_base + n) * E : a.band_of_epoch;
    double* xk = buf;                                 // [G] data sites; dead once the set-up is done
    if (!a.common_time_lu) set_segments(G);
    // ---- 0. set-up.  Every spectrum sample factorises around the Planck denominator:
    //   v[k][j] = colB[j] * rowB[k] / expm1(colA[j] * rowA[k])   (+ the sed.Line terms for type Ia),
    // so divisions, powers and unit conversions are done once per row / column.
    double rmin = INFINITY, rmax = 0.0, amax = 0.0, amin = INFINITY;   // bounds for the all-valid test
    // the same over the VALID rows only (T and R^2 finite and positive), and the rows that are neither valid nor plainly
    // dead (a negative factor gives finite non-zero entries of the wrong sign: those keep the two-pass path)
    double v_rmin = INFINITY, v_rmax = 0.0, v_amax = 0.0, v_amin = INFINITY, n_dead = 0.0, n_odd = 0.0;
    auto row_stats = [&](double ra, double rb_) {
      const bool ok = ra > 0.0 && ra < INFINITY && rb_ > 0.0 && rb_ < INFINITY;
      if (ok) {
        v_rmin = fmin(v_rmin, rb_); v_rmax = fmax(v_rmax, rb_); v_amax = fmax(v_amax, ra); v_amin = fmin(v_amin, ra);
      } else {
        n_dead += 1.0;
        if (ra < 0.0 || rb_ < 0.0) n_odd += 1.0;
      }
    };
    for (int k = tid; k < G; k += T) {
      xk[k] = gX[k];
      const double Tk = gT[k], R = gR[k];
      if (a.sed == PH_SED_BLACKBODY) {
        for (int c = 0; c < NC; ++c) {                          // kilonova_models.py:1150-1175, 1278-1300
          const double Tc = gT[c * a.grid_comp_stride + k], Rc = gR[c * a.grid_comp_stride + k];
          const double ra = 1.0 / (kKB * Tc), rb_ = Rc * Rc;
          rowA[c * NT + k] = ra;
          rowB[c * NT + k] = rb_;
          rmin = fmin(rmin, (rb_ > 0.0 && isfinite(rb_)) ? rb_ : 0.0);
          rmax = fmax(rmax, isfinite(rb_) ? rb_ : INFINITY);
          amax = fmax(amax, (ra > 0.0 && isfinite(ra)) ? ra : INFINITY);
          amin = fmin(amin, (ra > 0.0) ? ra : 0.0);
          row_stats(ra, rb_);
        }
      } else {
        const double ra = 1.0 / Tk;
        const double rb_ = (R * R) * cutoff_norm(gL[k], R, Tk, lam_c, a.absorption_index);
        rowA[k] = ra;
        rowB[k] = rb_;
        rmin = fmin(rmin, (rb_ > 0.0 && isfinite(rb_)) ? rb_ : 0.0);
        rmax = fmax(rmax, isfinite(rb_) ? rb_ : INFINITY);
        amax = fmax(amax, (ra > 0.0 && isfinite(ra)) ? ra : INFINITY);
        amin = fmin(amin, (ra > 0.0) ? ra : 0.0);
        row_stats(ra, rb_);
        if (line) {
          // sed.Line on the (N_lambda, N_t) grid (sed.py:859-909; supernova_models.py:2295-2303); time is source frame
          const double te = gX[k] / opz;
          const double q = (te - a.line_time) / a.line_duration;
          const double amp = a.line_amplitude * exp(-0.5 * (q * q));
          double amp_scaled = amp * gL[k] / (4 * kPi * (dl * dl));
          amp_scaled /= (a.line_width * sqrt(2 * kPi));
          rowC[k] = 1 - amp;
          rowD[k] = amp_scaled;
          if (!(amp_scaled >= 0.0 && isfinite(amp_scaled) && amp < 1.0)) rmax = INFINITY;   // force the pre-pass
        }
      }
    }
    double cmin = INFINITY, cmax = 0.0, camax = 0.0;
    double k_cmin = INFINITY, k_cmax = 0.0;            // over the kept columns j >= j_first
    for (int j = tid; j < NL; j += T) {
      const double lam_obs = a.lambda_obs[j];
      const double nu = (kCSI / (lam_obs * 1.e-10)) * opz;                       // utils.py:357-362, 383
      const double to_flam = 1e-26 * 2.99792458e18 / (lam_obs * lam_obs);        // mJy -> erg/cm^2/s/Angstrom
      double ca, cb;
      if (a.sed == PH_SED_BLACKBODY) {
        ca = kH * nu;
        cb = 2 * kPi * kH * (nu * nu * nu) / ((dl * dl) * (kC * kC)) * opz * kToMJy * to_flam;   // sed.py:216, 929
        colD[j] = 2 * kPi * kH * (nu * nu * nu);                                  // sed.py:217 num without R^2
      } else {
        const double lam_A = 1.e10 * (kCSI / nu);
        const double lam = lam_A * kAngstrom;
        const double l2 = lam * lam, l5 = l2 * l2 * lam;
        const double absorb = (lam < lam_c) ? pow(lam / lam_c, a.absorption_index) : 1.0;
        ca = kXConst / lam;
        double c = kFluxConst * (absorb / l5) / (4 * kPi * (dl * dl)) * lam_A / nu * kToMJy;        // sed.py:362-384
        if (line) {
          const double u = (lam_A - a.line_wavelength) / a.line_width;
          colC[j] = exp(-0.5 * (u * u)) * (lam * lam) / kC * kToMJy * opz * to_flam;
          c = c * 1e-26 * kToMJy;
        }
        cb = c * opz * to_flam;                                                   // supernova_models.py:1611-1612, 2304-2305
        colD[j] = 0.0;
      }
      if (a.ext != nullptr) {
        // spectra * 10^(-0.4 A(lambda_obs)): a factor per wavelength column (the line term of type Ia carries it too)
        const double att = ext_factor(*a.ext, lam_obs, opz, a.ext_av_host_s[n]);
        cb *= att;
        if (line) colC[j] *= att;
      }
      colA[j] = ca;
      colB[j] = cb;
      cmin = fmin(cmin, (cb > 0.0 && isfinite(cb)) ? cb : 0.0);
      cmax = fmax(cmax, isfinite(cb) ? cb : INFINITY);
      camax = fmax(camax, (ca > 0.0 && isfinite(ca)) ? ca : INFINITY);
      if (j >= j_first) {
        k_cmin = fmin(k_cmin, (cb > 0.0 && isfinite(cb)) ? cb : 0.0);
        k_cmax = fmax(k_cmax, isfinite(cb) ? cb : INFINITY);
      }
    }
    // block-wide bounds: red[0..7] min r, [8..15] max r, [16..23] max rowA, [24..31] min c, [32..39] max c, [40..47] max colA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      rmin = fmin(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
      rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
      amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      amin = fmin(amin, __shfl_xor_sync(0xffffffffu, amin, o));
      cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
      cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
      camax = fmax(camax, __shfl_xor_sync(0xffffffffu, camax, o));
      v_rmin = fmin(v_rmin, __shfl_xor_sync(0xffffffffu, v_rmin, o));
      v_rmax = fmax(v_rmax, __shfl_xor_sync(0xffffffffu, v_rmax, o));
      v_amax = fmax(v_amax, __shfl_xor_sync(0xffffffffu, v_amax, o));
      v_amin = fmin(v_amin, __shfl_xor_sync(0xffffffffu, v_amin, o));
      n_dead += __shfl_xor_sync(0xffffffffu, n_dead, o);
      n_odd += __shfl_xor_sync(0xffffffffu, n_odd, o);
      k_cmin = fmin(k_cmin, __shfl_xor_sync(0xffffffffu, k_cmin, o));
      k_cmax = fmax(k_cmax, __shfl_xor_sync(0xffffffffu, k_cmax, o));
    }
    if (lane == 0) {
      red[warp] = rmin; red[8 + warp] = rmax; red[16 + warp] = amax;
      red[24 + warp] = cmin; red[32 + warp] = cmax; red[40 + warp] = camax; red[48 + warp] = amin;
      red2[warp] = v_rmin; red2[8 + warp] = v_rmax; red2[16 + warp] = v_amax; red2[24 + warp] = v_amin;
      red2[32 + warp] = n_dead; red2[40 + warp] = n_odd; red2[48 + warp] = k_cmin; red2[56 + warp] = k_cmax;
    }
    __syncthreads();
    // bounds of the row factors over each group of 32 consecutive rows (what one warp handles per column): with the
    // column's factors they decide the evaluation regime of the whole group by a few scalar multiplies
    const int ng = (G + 31) >> 5;
    double* g_ralo = grp;                 // min rowA over the group's VALID rows (finite, > 0, finite R^2 > 0); +inf if none
    double* g_rahi = grp + ng;            // max rowA                               } over the whole group,
    double* g_rblo = grp + 2 * ng;        // log min rowB                           } meaningful when g_clean
    double* g_rbhi = grp + 3 * ng;        // log max rowB
    double* g_clean = grp + 4 * ng;       // 1: every row of the group is valid
    for (int g = warp; g < ng; g += nwarps) {
      const int k = g * 32 + lane;
      const bool in = k < G;
      const double ra = in ? rowA[k] : 0.0, rb = in ? rowB[k] : 0.0;
      const bool ok = in && ra > 0.0 && ra < INFINITY && rb > 0.0 && rb < INFINITY;
      double lo_v = ok ? ra : INFINITY, hi = ok ? ra : 0.0, blo = ok ? rb : INFINITY, bhi = ok ? rb : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo_v = fmin(lo_v, __shfl_xor_sync(0xffffffffu, lo_v, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        blo = fmin(blo, __shfl_xor_sync(0xffffffffu, blo, o));
        bhi = fmax(bhi, __shfl_xor_sync(0xffffffffu, bhi, o));
      }
      const bool clean = __all_sync(0xffffffffu, ok || !in);
      if (lane == 0) {
        g_ralo[g] = lo_v; g_rahi[g] = hi; g_clean[g] = clean ? 1.0 : 0.0;
        g_rblo[g] = clean ? log(blo) : -INFINITY;          // logarithms: the validity bounds below are sums
        g_rbhi[g] = clean ? log(bhi) : INFINITY;
      }
    }
    // ---- 0b. per-epoch time basis (knot interval + the four non-zero B-splines) and the collocation rows ----------
    for (int e = tid; e < En; e += T) {
      double x = t_row[e];
      double h[4] = {0.0, 0.0, 0.0, 0.0};
      double first = -1.0;
      bool live = true;
      if (a.t0_s != nullptr) {                                                    // phase_models.py:33-58
        x -= a.t0_s[n];
        if (x < 0.0) live = false;
      }
      if (live) {
        if (x < xk[0]) x = xk[0];                                                 // fpbisp clamps to [tb, te]
        if (x > xk[G - 1]) x = xk[G - 1];
        // knot interval: largest l in [3, G-1] with t[l] <= x  (t[l] = x[l-2] for 4 <= l <= G-1)
        int l = 3;
        int lo = 4, hi = G - 1;
        while (lo <= hi) {
          const int mid = (lo + hi) >> 1;
          if (xk[mid - 2] <= x) { l = mid; lo = mid + 1; } else hi = mid - 1;
        }
        double tl[8];
        for (int q = 0; q < 8; ++q) tl[q] = nak_knot(xk, G, l - 3 + q);
        fpbspl3(tl, 3, x, h);
        first = (double)(l - 3);
      }
      ept[0 * EP + e] = h[0]; ept[1 * EP + e] = h[1]; ept[2 * EP + e] = h[2]; ept[3 * EP + e] = h[3];
      ept[4 * EP + e] = first;
      ept[5 * EP + e] = 0.0;   // y[j-1] of the wavelength-axis forward elimination
      ept[6 * EP + e] = 0.0;   // y[j-2]
      ept[8 * EP + e] = 0.0;   // U_e: the row mask's time-axis solution contracted with this epoch's B-splines
    }
    if (!a.common_time_lu) {
      for (int i = tid; i < G; i += T) {
        // row i of B_j(x_i): interval l with t[l] <= x_i < t[l+1]; non-zero columns l-3..l
        double row[5] = {0, 0, 0, 0, 0};   // columns i-2..i+2
        if (i == 0) row[2] = 1.0;
        else if (i == G - 1) row[2] = 1.0;
        else {
          int l;
          if (i == 1) l = 3;
          else if (i >= G - 2) l = G - 1;
          else l = i + 2;
          double tl[8];
          for (int q = 0; q < 8; ++q) tl[q] = nak_knot(xk, G, l - 3 + q);
          double h[4];
          fpbspl3(tl, 3, xk[i], h);
          for (int q = 0; q < 4; ++q) {
            const int col = l - 3 + q, off = col - i + 2;
            if (off >= 0 && off < 5) row[off] = h[q];
          }
        }
        LUT0(i) = row[0]; LUT1(i) = row[1]; LUT2(i) = row[2]; LUT3(i) = row[3]; LUT4(i) = row[4];
      }
    }
    bool all_valid;
    double x_hi;
    {
      double r_lo = INFINITY, r_hi = 0.0, a_hi = 0.0, a_lo = INFINITY, c_lo = INFINITY, c_hi = 0.0, ca_hi = 0.0;
      for (int w = 0; w < nwarps; ++w) {
        r_lo = fmin(r_lo, red[w]); r_hi = fmax(r_hi, red[8 + w]); a_hi = fmax(a_hi, red[16 + w]);
        c_lo = fmin(c_lo, red[24 + w]); c_hi = fmax(c_hi, red[32 + w]); ca_hi = fmax(ca_hi, red[40 + w]);
        a_lo = fmin(a_lo, red[48 + w]);
      }
      x_hi = a_hi * ca_hi;
      const double x_lo = a_lo * colA[NL - 1];
      // every product colB rowB / expm1(x) is then finite and non-zero, in the reference's own evaluation order too (its
      // erg/s/Hz/cm^2 value is >= 3e-15 of ours): value >= cb R^2 e^-x > 1e-290 and <= cb R^2 / x < 1e250
      all_valid = (x_hi < 600.0) && (x_lo > 1e-12) && (r_lo > 0.0) && (c_lo > 0.0) &&
                  (log(r_lo) + log(c_lo) - x_hi > -667.0) &&
                  (log(r_hi) + log(c_hi) + log((double)NC) - log(fmin(x_lo, 1.0)) < 575.0);
    }
    // single-pass mode: the same bounds over the valid rows and the kept columns only
    bool single_pass = false;
    double dead_rows = 0.0;
    if (NC == 1 && !line && !all_valid) {
      double r_lo = INFINITY, r_hi = 0.0, a_hi = 0.0, a_lo = INFINITY, c_lo = INFINITY, c_hi = 0.0, odd = 0.0;
      for (int w = 0; w < nwarps; ++w) {
        r_lo = fmin(r_lo, red2[w]); r_hi = fmax(r_hi, red2[8 + w]); a_hi = fmax(a_hi, red2[16 + w]);
        a_lo = fmin(a_lo, red2[24 + w]); dead_rows += red2[32 + w]; odd += red2[40 + w];
        c_lo = fmin(c_lo, red2[48 + w]); c_hi = fmax(c_hi, red2[56 + w]);
      }
      const double xk_hi = a_hi * colA[j_first], xk_lo = a_lo * colA[NL - 1];
      single_pass = odd == 0.0 && dead_rows < (double)G && (xk_hi < 600.0) && (xk_lo > 1e-12) && (r_lo > 0.0) &&
                    (c_lo > 0.0) && (log(r_lo) + log(c_lo) - xk_hi > -667.0) &&
                    (log(r_hi) + log(c_hi) - log(fmin(xk_lo, 1.0)) < 575.0);
      if (tid == 0 && !single_pass) {
        const int why = (odd != 0.0) ? 4 : (!(xk_hi < 600.0) || !(xk_lo > 1e-12) || !(dead_rows < (double)G)) ? 5
                        : (!(r_lo > 0.0) || !(c_lo > 0.0)) ? 6 : 7;
        atomicAdd(&g_phot_modes[why], 1ull);
      }
    }
    const bool need_mask = single_pass && dead_rows > 0.0;
    // ---- 0c. red-side column cut (plain blackbody / cutoff spectra without extinction).  Towards the red a valid entry
    // grows by at most lambda^(x - 4) with x = h nu / k T <= x_e at the reddest band's last coefficient for the coolest
    // valid row (d ln B_lambda / d ln lambda = -5 + x e^x / (e^x - 1) <= x - 4), i.e. by g = q^max(x_e - 4, 0) per column,
    // and the collocation inverse decays like 0.268 q^2 per column: columns more than M beyond that coefficient change a
    // band coefficient by < 2.5 xr^M / (1 - xr), xr = 0.268 q^2 g.  M from the same 1e-11 target as the blue side; no
    // cut for samples with xr >= 0.8 (very cool rows) or with rows of the wrong sign.
    int j_last = NL - 1;
    if (RED && NC == 1 && !line && a.ext == nullptr && a.q_red > 0.0) {
      double ah = 0.0, odd = 0.0;
      if (all_valid) {
        for (int w = 0; w < nwarps; ++w) ah = fmax(ah, red[16 + w]);
      } else {
        for (int w = 0; w < nwarps; ++w) { ah = fmax(ah, red2[16 + w]); odd += red2[40 + w]; }
      }
      if (odd == 0.0 && ah > 0.0) j_last = min(NL - 1, a.j_store_hi + red_margin(ah * colA[a.j_store_hi], a.q_red, NL));
    }
    if (tid == 0) atomicAdd(&g_phot_modes[all_valid ? 0 : (single_pass ? (need_mask ? 2 : 1) : 3)], 1ull);
    __syncthreads();                                  // xk (in buf) is dead from here on
    PHOT_PROF_ACC(0);

    const double inv_den = 1.0 / ((dl * dl) * (kC * kC));                           // sed.py:218
    const double ca_min = colA[NL - 1];
    // General evaluation of one spectrum sample (column j, row k): any regime, any SED, any number of components.
    auto spec = [&](double ca, double cb, double cd, int k, int j) -> double {
      const double x = ca * rowA[k];
      double v = cb * rowB[k] * planck_inv_general(x, tab);
      if (x > 650.0 && x <= 709.782712893384 && a.sed == PH_SED_BLACKBODY) {
        // the reference evaluates (2 pi h nu^3 R^2 / (dl^2 c^2)) * (1 / expm1(x)) in erg/s/Hz/cm^2 (sed.py:217-221),
        // which underflows to 0 (-> replaced) ~1e7 earlier than the product in our output units
        const double vc = ((cd * rowB[k]) * inv_den) * (1.0 / expm1(x));
        if (vc == 0.0) v = 0.0;
      }
      for (int c = 1; c < NC; ++c) {
        const double xc = ca * rowA[c * NT + k];
        double vc2 = cb * rowB[c * NT + k] * planck_inv_general(xc, tab);
        if (xc > 650.0 && xc <= 709.782712893384 && ((cd * rowB[c * NT + k]) * inv_den) * (1.0 / expm1(xc)) == 0.0) vc2 = 0.0;
        v += vc2;
      }
      if (line) v = v * rowC[k] + rowD[k] * colC[j];
      return v;
    };
    // Regime of (column, group of 32 rows) from the group's bounds -- all scalar, uniform over the warp:
    //   0 dead   : every row overflows the Planck denominator or is itself invalid -> every sample is replaced
    //   1 wien   : 40 < x <= 650, all samples valid: value = cb R^2 e^-x (the two differ by e^-x < 4e-18)
    //   2 planck : x <= 40, all samples valid: value = cb R^2 / expm1(x)
    //   3 general: anything else, and every multi-component / line spectrum
    const bool plain = NC == 1 && !line;
    // every sample of a regime-1/2 group is finite and non-zero, here AND in the reference's own evaluation order (its
    // erg/s/Hz/cm^2 value is >= 3e-15 of ours): value >= cb R^2 e^-x > 1e-290 and <= cb R^2 / x < 1e250, as sums of logs
    auto regime = [&](double ca, double lcb, int g) -> int {
      if (!plain) return 3;
      if (ca * g_ralo[g] > 709.782712893384) return 0;
      const double xlo = ca * g_ralo[g], xhi = ca * g_rahi[g];
      if (g_clean[g] == 0.0 || !(lcb + g_rblo[g] - xhi > -667.0) || !(lcb + g_rbhi[g] < 540.0)) return 3;
      if (xlo > 40.0 && xhi <= 650.0) return 1;
      if (xhi <= 40.0 && xlo > 1e-12) return 2;
      return 3;
    };

    // The regimes of all row groups of one column at once: lane = group, the answers travel as ballot masks, so the
    // loops over groups below test bits instead of re-deriving the regime (which costs more than a fast evaluation).
    // m[0] dead, m[1] wien, m[2] planck, m[3] wien groups whose terms are all e^-50 below their row's largest (`dca` =
    // colA - colA of the reddest column); bit g of word g / 32.
    auto regime_masks = [&](double ca, double lcb, double dca, unsigned long long (&m)[4]) {
      m[0] = m[1] = m[2] = m[3] = 0ull;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        if (w == 1 && ng <= 32) break;
        const int g = w * 32 + lane;
        const int r = g < ng ? regime(ca, lcb, g) : 15;
        const bool tail = r == 1 && g_ralo[min(g, ng - 1)] * dca > 50.0;
        m[0] |= (unsigned long long)__ballot_sync(0xffffffffu, r == 0) << (32 * w);
        m[1] |= (unsigned long long)__ballot_sync(0xffffffffu, r == 1) << (32 * w);
        m[2] |= (unsigned long long)__ballot_sync(0xffffffffu, r == 2) << (32 * w);
        m[3] |= (unsigned long long)__ballot_sync(0xffffffffu, tail) << (32 * w);
      }
    };
    auto quad_bits = [&](unsigned long long mw, int g0) -> unsigned { return (unsigned)(mw >> g0) & 15u; };

    // ---- 1. statistics pre-pass for the clean-up value (sed.py:985-992), with the serial LU underneath ----------
    // Banded LU of the collocation matrix without pivoting (totally positive), in place, in the scaled layout of
    // load_lu_row.  The matrix is tridiagonal except for the two not-a-knot rows (1 and G-2), so for rows 3 .. G-3 the
    // pivots obey d_i = b_i - a_i c_{i-1} / d_{i-1}; with d_i = P_i / P_{i-1} that is the LINEAR recurrence
    // P_i = b_i P_{i-1} - (a_i c_{i-1}) P_{i-2}: one FMA per row on the serial chain instead of multiply, FMA and a
    // reciprocal (~200 cycles per row at FP64 latencies).  One thread walks the chain (P in `hom`, which is rebuilt
    // afterwards) while the other warps run the statistics pre-pass; the reciprocals and row scalings are then done by
    // all threads in parallel, and the last two rows by one thread again.  |P| shrinks like 0.6^i: no underflow for
    // any grid that fits.
    const bool lu_thread = !a.common_time_lu && warp == nwarps - 1;
    const bool lu_fast = !a.common_time_lu && G >= 10;
    auto lu_row = [&](int i) {
      const double a0 = LUT0(i), a1 = LUT1(i), a2 = LUT2(i), a3 = LUT3(i), a4 = LUT4(i);
      double rd1 = 0.0, s1_1 = 0.0, s2_1 = 0.0, rd2 = 0.0, s1_2 = 0.0, s2_2 = 0.0;   // 1/d, u1/d, u2/d of rows i-1, i-2
      if (i >= 1) { rd1 = LUT2(i - 1); s1_1 = LUT3(i - 1); s2_1 = LUT4(i - 1); }
      if (i >= 2) { rd2 = LUT2(i - 2); s1_2 = LUT3(i - 2); s2_2 = LUT4(i - 2); }
      const double am1 = a1 - a0 * s1_2;
      const double d = (a2 - a0 * s2_2) - am1 * s1_1;
      const double u1 = a3 - am1 * s2_1;
      const double rd = rcp_fast(d);
      LUT0(i) = a0 * rd2;
      LUT1(i) = am1 * rd1;
      LUT2(i) = rd;
      LUT3(i) = u1 * rd;
      LUT4(i) = a4 * rd;
    };
    if (lu_thread && lane == 0) {
      if (!lu_fast) {
        for (int i = 0; i < G; ++i) lu_row(i);
      } else {
        lu_row(0);
        lu_row(1);
        lu_row(2);
        double pm2 = 1.0, pm1 = LUT2(3) - LUT1(3) * LUT3(2);      // P_2 = 1, P_3 = d_3 = b_3 - a_3 (u1/d)_2
        hom[2] = pm2;
        hom[3] = pm1;
#pragma unroll 4
        for (int i = 4; i <= G - 3; ++i) {
          const double ac = LUT1(i) * LUT3(i - 1);                // a_i c_{i-1} (raw rows)
          const double pi = fma(LUT2(i), pm1, -(ac * pm2));
          hom[i] = pi;
          pm2 = pm1;
          pm1 = pi;
        }
      }
    }
    double repl = 0.0;
    double pre_sum = 0.0, pre_cnt = 0.0;             // single-pass mode: sum / count of the columns below the kept ones
    if (!all_valid && !(single_pass && !need_mask)) {
      double vsum = 0.0, vcnt = 0.0;
      const int wstat = a.common_time_lu ? nwarps : nwarps - 1;     // the LU warp does not take part
      // single-pass mode: the columns outside the kept range [j_first, j_last]; otherwise all of them
      const int j_stat_end = single_pass ? j_first + (NL - 1 - j_last) : NL;
      if (warp < wstat) {
        for (int jj = warp; jj < j_stat_end; jj += wstat) {
          const int j = (single_pass && jj >= j_first) ? j_last + 1 + (jj - j_first) : jj;
          const double ca = colA[j], cb = colB[j], cd = colD[j];
          const double lcb = (cb > 0.0 && cb < INFINITY) ? log(cb) : NAN;
          const double dca = ca - ca_min;
          unsigned long long m[4];
          regime_masks(ca, lcb, dca, m);
          // A term more than e^-50 below its row's largest one (the reddest column's, or the Planck peak) cannot change
          // the sum at 1e-16: such groups (all samples valid) are counted, not evaluated.
          if (lane == 0) {
            const int nt = __popcll(m[3]);
            const bool last_tail = (m[3] >> (ng - 1)) & 1ull;
            vcnt += (double)(32 * nt - (last_tail ? 32 * ng - G : 0));
          }
          // four groups of 32 rows per trip: when all of them are in a fast regime the four Planck evaluations of a lane
          // are independent chains that overlap (the kernel runs at 16 warps per SM: latency is hidden by ILP, not TLP)
          for (int g0 = 0; g0 < ng; g0 += 4) {
            const unsigned skip = quad_bits(m[0], g0) | quad_bits(m[3], g0);       // dead or counted
            const unsigned qw = quad_bits(m[1], g0) & ~skip, qp = quad_bits(m[2], g0);
            const unsigned in4 = ng - g0 >= 4 ? 15u : (1u << (ng - g0)) - 1u;
            if (skip == in4) continue;
            if ((qw | qp) == 15u) {
              double x[4], c[4], pv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int k = min((g0 + u) * 32 + lane, G - 1);
                x[u] = ca * rowA[k];
                c[u] = cb * rowB[k];
              }
              if (qw == 15u) {
#pragma unroll
                for (int u = 0; u < 4; ++u) pv[u] = exp_lean(-x[u], tab);
              } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) pv[u] = expm1_tab(x[u], tab);
#pragma unroll
                for (int u = 0; u < 4; ++u) pv[u] = rcp_fast(pv[u]);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if ((g0 + u) * 32 + lane < G) { vsum += c[u] * pv[u]; vcnt += 1.0; }
              continue;
            }
            for (int u = 0; u < 4; ++u) {
              if (((skip | ~in4) >> u) & 1u) continue;
              const int k = min((g0 + u) * 32 + lane, G - 1);
              const bool live = (g0 + u) * 32 + lane < G;
              if (((qw | qp) >> u) & 1u) {
                const double x = ca * rowA[k];
                const double pv = ((qw >> u) & 1u) ? exp_lean(-x, tab) : rcp_fast(expm1_tab(x, tab));
                if (live) { vsum += (cb * rowB[k]) * pv; vcnt += 1.0; }
              } else {
                const double v = spec(ca, cb, cd, k, j);
                if (live && isfinite(v) && v != 0.0) { vsum += v; vcnt += 1.0; }
              }
            }
          }
        }
      }
      vsum = warp_sum(vsum);
      vcnt = warp_sum(vcnt);
      if (lane == 0) { red[48 + warp] = vsum; red[56 + warp] = vcnt; }
      __syncthreads();
      double s = 0.0, c = 0.0;
      for (int w = 0; w < nwarps; ++w) { s += red[48 + w]; c += red[56 + w]; }
      if (single_pass) {
        pre_sum = s;
        pre_cnt = c;
      } else {
        double mean = (c > 0) ? s / c : NAN;
        if (!isfinite(mean) || mean == 0.0) mean = 1.0;
        repl = 1e-30 * mean;
      }
    } else {
      __syncthreads();
    }
    if (lu_fast) {
      for (int i = 3 + tid; i <= G - 3; i += T) {
        const double rd = hom[i - 1] / hom[i];                    // 1 / d_i
        const double rd_prev = (i == 3) ? LUT2(2) : hom[i - 2] / hom[i - 1];
        const double a1 = LUT1(i), c1 = LUT3(i), c2 = LUT4(i);
        LUT0(i) = 0.0;
        LUT1(i) = a1 * rd_prev;
        LUT2(i) = rd;
        LUT3(i) = c1 * rd;
        LUT4(i) = c2 * rd;
      }
      __syncthreads();
      if (tid == 0) {
        lu_row(G - 2);
        lu_row(G - 1);
      }
      __syncthreads();
    }
    if (!a.common_time_lu) build_tables(G);
    // per epoch: the weights of the two boundary values that the rows l .. l+3 of its B-splines respond to (2b)
    for (int e = tid; e < En; e += T) {
      const int l = (int)ept[4 * EP + e];
      if (l < 0) continue;
      const int s0 = l / L;
      const int cnt = min(4, (s0 + 1) * L - l);
      double ga = 0.0, gb = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double g = ept[q * EP + e] * HOMB(l + q);
        if (q < cnt) ga += g; else gb += g;
      }
      ept[9 * EP + e] = ga;
      ept[10 * EP + e] = gb;
      ept[11 * EP + e] = (double)s0;
    }
    __syncthreads();
    PHOT_PROF_ACC(1);

    // ---- 2. wavelength-column blocks ---------------------------------------------------------------------------
    const int seg = tid & (P - 1), jl = tid >> lgP;
    const int i_lo = seg * L, i_hi = min(i_lo + L, G);
    const double* mF = mtab + 4 * seg;
    const double* mB = mtab + kScanSteps * 128 + 4 * seg;
    const double fill = single_pass ? 0.0 : repl;     // what a replaced entry holds during the solves
    double ksum = 0.0;                                // single-pass mode: sum of the kept columns' valid entries
    for (int it = need_mask ? -1 : 0;; ++it) {
      const bool mask_it = it < 0;                    // the row mask as an extra column (single-pass mode)
      const int j0 = j_first + max(it, 0) * W;
      if (!mask_it && j0 > j_last) break;
      const int wc = mask_it ? 1 : min(W, j_last + 1 - j0);
      // 2a. spectra of the block, cleaned
      if (mask_it) {
        for (int k = tid; k < G; k += T) {
          const double ra = rowA[k], rb = rowB[k];
          buf[k] = (ra > 0.0 && ra < INFINITY && rb > 0.0 && rb < INFINITY) ? 0.0 : 1.0;
        }
      }
      for (int q = warp; q < (mask_it ? 0 : wc); q += nwarps) {
        double* col = buf + q * NTP;
        const double ca = colA[j0 + q], cb = colB[j0 + q], cd = colD[j0 + q];
        const double lcb = (cb > 0.0 && cb < INFINITY) ? log(cb) : NAN;
        unsigned long long m[4];
        regime_masks(ca, lcb, 0.0, m);
        for (int g0 = 0; g0 < ng; g0 += 4) {
          const unsigned qd = quad_bits(m[0], g0), qw = quad_bits(m[1], g0), qp = quad_bits(m[2], g0);
          const unsigned in4 = ng - g0 >= 4 ? 15u : (1u << (ng - g0)) - 1u;
          if ((qw | qp) == 15u) {
            double x[4], c[4], pv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int k = min((g0 + u) * 32 + lane, G - 1);
              x[u] = ca * rowA[k];
              c[u] = cb * rowB[k];
            }
            if (qw == 15u) {
#pragma unroll
              for (int u = 0; u < 4; ++u) pv[u] = exp_lean(-x[u], tab);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u) pv[u] = expm1_tab(x[u], tab);
#pragma unroll
              for (int u = 0; u < 4; ++u) pv[u] = rcp_fast(pv[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int k = (g0 + u) * 32 + lane;
              if (k < G) { const double v = c[u] * pv[u]; col[k] = v; ksum += v; }
            }
            continue;
          }
          for (int u = 0; u < 4; ++u) {
            if (!((in4 >> u) & 1u)) continue;
            const int k = min((g0 + u) * 32 + lane, G - 1);
            const bool live = (g0 + u) * 32 + lane < G;
            double v;
            bool valid = true;
            if ((qd >> u) & 1u) {
              valid = false;
            } else if ((qw >> u) & 1u) {
              v = (cb * rowB[k]) * exp_lean(-(ca * rowA[k]), tab);
            } else if ((qp >> u) & 1u) {
              v = (cb * rowB[k]) * rcp_fast(expm1_tab(ca * rowA[k], tab));
            } else {
              v = spec(ca, cb, cd, k, j0 + q);
              valid = isfinite(v) && v != 0.0;
            }
            if (!valid) v = fill;
            if (live) { col[k] = v; if (valid) ksum += v; }
          }
        }
      }
      __syncthreads();
      PHOT_PROF_ACC(2);
      // 2b. time axis: the thread's (column, segment) through forward recurrence, boundary scan, backward recurrence,
      // reverse scan and fix-up; the P lanes of a column exchange boundary states by shuffles only
      {
        const bool active = jl < wc && i_lo < G;
        double* col = buf + jl * NTP;
        double p1 = 0.0, p2 = 0.0;
        if (active) {
#pragma unroll 4
          for (int i = i_lo; i < i_hi; ++i) {
            const double y = col[i] - L1D(i) * p1;
            col[i] = y;
            p2 = p1;
            p1 = y;
          }
          const int g2 = G - 2;                       // the not-a-knot row with l2 != 0; nothing after it depends on it
          if (g2 >= i_lo && g2 < i_hi && g2 - 2 >= i_lo) {     // (l1 of row G-1 is zero); y[G-4] outside: in HOMF
            const double y = col[g2] - LUT0(g2) * col[g2 - 2];
            col[g2] = y;
            if (g2 == i_hi - 1) p1 = y; else p2 = y;
          }
          if (i_hi - i_lo < 2) p2 = 0.0;              // one-row segment: y[i_hi-2] belongs to the previous one (H row = [1 0])
        }
        // inclusive scan of the segments' affine maps: afterwards (p1, p2) is the TRUE state at the segment's end
        for (int st = 0; st < lgP; ++st) {
          const int d = 1 << st;
          const double q1 = __shfl_up_sync(0xffffffffu, p1, d, P), q2 = __shfl_up_sync(0xffffffffu, p2, d, P);
          if (seg >= d) {
            const double* M = mF + st * 128;
            const double n1 = p1 + M[0] * q1 + M[1] * q2;
            const double n2 = p2 + M[2] * q1 + M[3] * q2;
            p1 = n1;
            p2 = n2;
          }
        }
        double Y1 = __shfl_up_sync(0xffffffffu, p1, 1, P), Y2 = __shfl_up_sync(0xffffffffu, p2, 1, P);
        if (seg == 0) { Y1 = 0.0; Y2 = 0.0; }
        p1 = 0.0;
        p2 = 0.0;
        if (active) {
          // forward fix-up, division by the pivot and the backward recurrence in one sweep (one read and one write of the
          // column less than as separate passes); the fix-up and the scaling are off the recurrence's dependency chain
#pragma unroll 4
          for (int i = i_hi - 1; i >= i_lo; --i) {
            const double yv = (col[i] + HOMF(i) * Y1) * LUT2(i);
            const double x = yv - U1D(i) * p1;
            col[i] = x;
            p2 = p1;
            p1 = x;
          }
          if (i_lo == 0 && 3 < i_hi) {                // row 1, the not-a-knot row with u2 != 0 (u1 of row 0 is zero);
            const double x = col[1] - LUT4(1) * col[3];   // x[3] outside the segment: in HOMB
            col[1] = x;
            if (i_hi - i_lo >= 2) p2 = x;
          }
          if (i_hi - i_lo < 2) p2 = 0.0;
        }
        for (int st = 0; st < lgP; ++st) {
          const int d = 1 << st;
          const double q1 = __shfl_down_sync(0xffffffffu, p1, d, P), q2 = __shfl_down_sync(0xffffffffu, p2, d, P);
          if (seg + d < P) {
            const double* M = mB + st * 128;
            const double n1 = p1 + M[0] * q1 + M[1] * q2;
            const double n2 = p2 + M[2] * q1 + M[3] * q2;
            p1 = n1;
            p2 = n2;
          }
        }
        double X1 = __shfl_down_sync(0xffffffffu, p1, 1, P), X2 = __shfl_down_sync(0xffffffffu, p2, 1, P);
        if (seg == P - 1) { X1 = 0.0; X2 = 0.0; }
        // The backward fix-up x[i] += HOMB(i) X1 is NOT swept over the column (a read and a write of every row): the
        // only consumer is the contraction below, r_e = sum_q h_q x[l+q], so it takes the local solution and adds
        // Ga_e X1[seg(l)] + Gb_e X1[seg(l)+1] with Ga / Gb = sum of h_q HOMB(l+q) over the rows of either segment.
        xfix[tid] = active ? X1 : 0.0;
      }
      __syncthreads();
      PHOT_PROF_ACC(3);
      // 2c. contraction with the epochs' time B-splines and forward elimination of the wavelength-axis solve on the E
      // resulting rows
      for (int e = tid; e < En; e += T) {
        const int l = (int)ept[4 * EP + e];
        if (l >= 0 && mask_it) {
          const double* c = buf + l;
          const int s0 = (int)ept[11 * EP + e];
          ept[8 * EP + e] = ept[0 * EP + e] * c[0] + ept[1 * EP + e] * c[1] + ept[2 * EP + e] * c[2] + ept[3 * EP + e] * c[3] +
                            ept[9 * EP + e] * xfix[s0] + ept[10 * EP + e] * xfix[min(s0 + 1, P - 1)];
        } else if (l >= 0) {
          const double h0 = ept[0 * EP + e], h1 = ept[1 * EP + e], h2 = ept[2 * EP + e], h3 = ept[3 * EP + e];
          const double ga = ept[9 * EP + e], gb = ept[10 * EP + e];
          const int s0 = (int)ept[11 * EP + e], s1 = min(s0 + 1, P - 1);
          double y1 = ept[5 * EP + e], y2 = ept[6 * EP + e];
          const double* c = buf + l;
#pragma unroll 2
          for (int q = 0; q < wc; ++q, c += NTP) {
            const double r = h0 * c[0] + h1 * c[1] + h2 * c[2] + h3 * c[3] + ga * xfix[q * P + s0] + gb * xfix[q * P + s1];
            const int j = j0 + q;
            const double y = (r - lul[5 * j + 0] * y2) - lul[5 * j + 1] * y1;
            if (j >= a.j_store_lo) ys[(size_t)(j - a.j_store_lo) * EP + e] = y;
            y2 = y1;
            y1 = y;
          }
          ept[5 * EP + e] = y1;
          ept[6 * EP + e] = y2;
        }
      }
      __syncthreads();
      PHOT_PROF_ACC(4);
    }
    if (RED && j_last < NL - 1) {
      // the columns beyond j_last count as zero spectra: the wavelength-axis forward elimination just runs on
      for (int e = tid; e < En; e += T) {
        if ((int)ept[4 * EP + e] < 0) continue;
        double y1 = ept[5 * EP + e], y2 = ept[6 * EP + e];
        for (int j = j_last + 1; j < NL; ++j) {
          const double y = -(lul[5 * j + 0] * y2) - lul[5 * j + 1] * y1;
          if (j >= a.j_store_lo) ys[(size_t)(j - a.j_store_lo) * EP + e] = y;
          y2 = y1;
          y1 = y;
        }
      }
      __syncthreads();
    }

    // ---- 3. epochs: wavelength-axis back substitution, band integrals, magnitudes, likelihood ------------------
    // chunks of epochs (grouped by band through epoch_order) sized by the coefficient buffer:
    //   a. thread per epoch: back substitution from the last column down to the band's first coefficient; the band's
    //      coefficients go to the shared buffer; if none of them is negative the band integral is their dot product
    //      with the band's contracted weights, done on the way;
    //   b. warp per epoch, only where a coefficient is negative: the integration points split between the lanes, |f|;
    //   c. thread per epoch: magnitude / flux, likelihood term, coalesced store.
    if (need_mask) {
      // replacement value (sed.py:985-992) from the sums gathered on the way: kept columns + the columns below them;
      // every entry of a valid row is valid in the kept columns (single_pass bound), so their count is a product
      ksum = warp_sum(ksum);
      if (lane == 0) red[warp] = ksum;
      __syncthreads();
      double sacc = pre_sum;
      for (int w = 0; w < nwarps; ++w) sacc += red[w];
      const double cacc = pre_cnt + ((double)G - dead_rows) * (double)(j_last + 1 - j_first);
      double mean = (cacc > 0) ? sacc / cacc : NAN;
      if (!isfinite(mean) || mean == 0.0) mean = 1.0;
      repl = 1e-30 * mean;
      __syncthreads();
    }
    double logl_part = 0.0;
    const int SM = a.span_max;
    const int chunk = max(1, min((W * NTP) / SM, 4096));
    const double sp = a.sigma_param_s ? a.sigma_param_s[n] : 0.0;
    for (int e0 = 0; e0 < En; e0 += chunk) {
      const int ne = min(chunk, En - e0);
      for (int i = tid; i < ne; i += T) {
        const int e = ragged ? e0 + i : a.epoch_order[e0 + i];
        const int b = band_row[e];
        const int jlo = a.band_jlo[b], span = a.band_span[b];
        double* cb = buf + (size_t)i * SM;
        if (ept[4 * EP + e] < 0.0) continue;                    // before t0: nothing to integrate
        double x1 = 0.0, x2 = 0.0, lin = 0.0, cneg = 0.0;
        const double* yp = ys + e;
        const double* wgt = a.band_weight + (size_t)b * SM;
        const double corr = need_mask ? repl * ept[8 * EP + e] : 0.0;   // single-pass mode: r U_e on every coefficient
        for (int j = NL - 1; j >= jlo; --j) {
          const double x = (yp[(size_t)(j - a.j_store_lo) * EP] * lul[5 * j + 2] - lul[5 * j + 4] * x2) - lul[5 * j + 3] * x1;
          if (j < jlo + span) {
            const double xc = x + corr;
            cb[j - jlo] = xc;
            lin = fma(xc, __ldg(wgt + (j - jlo)), lin);
            cneg = fmin(cneg, xc);                              // NaN-safe: fmin drops a NaN operand, handled below
            if (!(xc == xc)) cneg = -1.0;
          }
          x2 = x1;
          x1 = x;
        }
        ept[7 * EP + e] = lin;
        ept[5 * EP + e] = cneg;                                 // < 0: the point-by-point path is needed
      }
      __syncthreads();
      PHOT_PROF_ACC(5);
      for (int i = warp; i < ne; i += nwarps) {
        const int e = ragged ? e0 + i : a.epoch_order[e0 + i];
        if (ept[4 * EP + e] < 0.0 || !(ept[5 * EP + e] < 0.0)) continue;
        const int b = band_row[e];
        const int w0 = a.band_off[b], w1 = a.band_off[b + 1];
        const double* dw = buf + (size_t)i * SM;
        double acc = 0.0;
        // four integration points per lane and trip: the loads of a trip overlap instead of forming one chain per point
        constexpr int kPts = 4;
        for (int wb = w0 + lane; wb < w1; wb += 32 * kPts) {
          double2 r0[kPts], r1[kPts], r2[kPts];
#pragma unroll
          for (int u = 0; u < kPts; ++u) {
            const int w = wb + 32 * u;
            const double2* rec = reinterpret_cast<const double2*>(a.int_rec) + 3 * (size_t)min(w, w1 - 1);
            r0[u] = __ldg(rec);
            r1[u] = __ldg(rec + 1);
            r2[u] = __ldg(rec + 2);
            if (w >= w1) r2[u].x = 0.0;
          }
#pragma unroll
          for (int u = 0; u < kPts; ++u) {
            const double* dq = dw + (int)r2[u].y;
            const double f = r0[u].x * dq[0] + r0[u].y * dq[1] + r1[u].x * dq[2] + r1[u].y * dq[3];
            acc += r2[u].x * fabs(f);                                             // sed.py:36-37
          }
        }
        acc = warp_sum(acc);
        if (lane == 0) ept[7 * EP + e] = acc;
      }
      __syncthreads();
      PHOT_PROF_ACC(6);
      for (int i = tid; i < ne; i += T) {
        const int e = ragged ? e0 + i : a.epoch_order[e0 + i];
        double model_v;
        if (ept[4 * EP + e] < 0.0) {
          model_v = (a.output == RB_OUT_FLUX) ? 0.0 : 1000.0;
        } else {
          const int b = band_row[e];
          const double bandflux = ept[7 * EP + e] * a.band_scale[b];
          model_v = -2.5 * log10(bandflux / a.band_zp[b]);                        // sed.py:114
          if (a.output == RB_OUT_FLUX) model_v = pow(10.0, model_v / (-2.5)) * a.band_ref[b];   // utils.py:716-718
        }
        if (a.out_model != nullptr) a.out_model[(a.n_base + n) * (int64_t)E + e] = model_v;
        if (a.want_logl)
          logl_part += logl_term(a.sigma_mode, a.epoch_tab[EP_KIND * E + e], a.epoch_tab[EP_Y * E + e], model_v,
                                 a.epoch_tab[EP_ISIG * E + e], a.epoch_tab[EP_LOGT * E + e], sp);
      }
      __syncthreads();
    }
    if (a.want_logl) {
      logl_part = warp_sum(logl_part);
      if (lane == 0) red[warp] = logl_part;
      __syncthreads();
      if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < nwarps; ++w) total += red[w];
        a.out_logl[a.n_base + n] = nan_to_num(total);
      }
    }
    __syncthreads();
    PHOT_PROF_ACC(7);
    PHOT_PROF_END();
  }
  PHOT_PROF_REPORT();
}

}  // namespace rb
