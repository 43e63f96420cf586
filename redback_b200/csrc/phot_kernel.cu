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
// One persistent block per sample works on a private N_t x N_lambda scratch tile in global memory (L2):
//   1. spectra S[k][j] from the model's (T_k, R_k[, L_k]) on its time grid         (all threads, elementwise)
//   2. clean: non-finite / zero -> 1e-30 * mean(valid)                              (block reduction)
//   3. spline coefficients C = A_t^-1 S A_lambda^-T: banded (pentadiagonal) LU of the B-spline collocation
//      matrices, no pivoting (totally positive); time axis: one thread per wavelength column; wavelength
//      axis: one thread per time row, staged through shared-memory tiles
//   4. one thread per epoch: FITPACK fpbspl time basis, contraction over time, then the band's 5-Angstrom
//      integration grid with host-precomputed wavelength bases; |f| is summed as the reference does
//   5. magnitude / flux, Gaussian log-likelihood term, block reduction.
#include "rb_common.cuh"

namespace rb {

constexpr int kPhotThreads = 256;
constexpr int kPhotMaxSpan = 96;     // max B-spline coefficients (along wavelength) overlapped by one band
constexpr int kPhotTileRows = 32;    // rows per shared-memory tile in the wavelength-axis solve
constexpr int kEpRec = 10;           // doubles per epoch in the hand-over array of the epoch phase
constexpr int kPhotBatch = 8;        // rows of the scratch tile loaded ahead of the time-axis recurrence

// shared-memory doubles of the tile region: a tile of rows for the wavelength-axis solve, or the 2 (per component) + 2
// per-epoch and 3
// per-wavelength factor arrays of the spectra phase, or the per-warp coefficient buffers and per-epoch hand-over arrays
// of the epoch phase (at least 128 epochs per chunk)
__host__ __device__ inline int phot_tile_elems(int NT, int NL, int n_comp) {
  const int t = kPhotTileRows * (NL + 1), f = (2 * n_comp + 2) * NT + 3 * NL, e = (kPhotThreads / 32) * kPhotMaxSpan + kEpRec * 128;
  return max(max(t, f), e);
}

enum { PH_SED_BLACKBODY = 1, PH_SED_CUTOFF = 2, PH_SED_CUTOFF_LINE = 3 };
enum { PH_T = 0, PH_R, PH_L, PH_X, kPhotGridCols };   // per-sample model grid: temperature, radius, L_bol, phase

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
  const double* int_rec;        // [total][6] per integration point: the 4 non-zero wavelength B-splines, wave * trans,
                                //            index of the first of them relative to the band's first coefficient
  double* scratch;              // [gridDim][n_t_max][n_lambda]
  double* out_model;
  double* out_logl;
};

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

// Banded LU rows (l2, l1, d, u1, u2) are kept in shared memory as (l2, l1, 1/d, u1/d, u2/d): a back-substitution step
// is then x = b/d - (u1/d) x[i+1] - (u2/d) x[i+2] with a single FMA on the serial dependency chain (b/d and the x[i+2]
// term do not depend on the previous step), instead of two FMAs and a division.
__device__ __forceinline__ void load_lu_row(double* __restrict__ dst, const double* __restrict__ src) {
  const double rd = 1.0 / src[2];
  dst[0] = src[0];
  dst[1] = src[1];
  dst[2] = rd;
  dst[3] = src[3] * rd;
  dst[4] = src[4] * rd;
}

#ifdef RB_PHOT_PROFILE
// developer aid (tools/phot_phases.sh): block 0 prints the cycles it spent in each phase of its first samples
#define PHOT_PROF_BEGIN() long long pt_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; int pn_ = 0; long long pc_ = clock64(), pa_ = pc_;
#define PHOT_PROF_ACC(i) { __syncthreads(); const long long c_ = clock64(); pt_[i] += c_ - pa_; pa_ = c_; }
#define PHOT_PROF(i) { __syncthreads(); const long long c_ = clock64(); pt_[pn_++] = c_ - pc_; pc_ = c_; }
#define PHOT_PROF_END()                                                                                             \
  if (blockIdx.x == 0 && tid == 0 && n < 3 * (int64_t)gridDim.x)                                                      \
    printf("phot phases [cycles] setup %lld spectra %lld cleanup+lu %lld forward %lld back+lambda %lld (back %lld "     \
           "lambda %lld store %lld) epochs %lld (basis %lld integrate %lld)\n", pt_[0], pt_[1], pt_[2], pt_[3], pt_[4], pt_[8], \
           pt_[9], pt_[10], pt_[5], pt_[6], pt_[7]);
#else
#define PHOT_PROF_BEGIN()
#define PHOT_PROF(i)
#define PHOT_PROF_ACC(i)
#define PHOT_PROF_END()
#endif

__global__ void __launch_bounds__(kPhotThreads, 2) phot_kernel(const PhotArgs a) {
  extern __shared__ double psm[];
  const int NT = a.n_t_max, NL = a.n_lambda, E = a.E, NC = a.n_comp;
  double* xk = psm;                       // [NT]      time-axis data sites (observer-frame phase, days)
  double* lut = xk + NT;                  // [NT][5]   time-axis banded LU
  double* lul = lut + 5 * NT;             // [NL][5]   wavelength-axis banded LU
  double* tile = lul + 5 * NL;            // [kPhotTileRows][NL + 1]; holds the row / column factors while the spectra are made
  double* red = tile + phot_tile_elems(NT, NL, NC);    // [40]
  double* bst = red + 40;                 // [NL][2]   time-axis back-substitution state per wavelength column
  double* hom = bst + 2 * NL;             // [NL][4]   homogeneous solutions of the partitioned wavelength-axis solve
  double* bnd = hom + 4 * NL;             // [nwarps][kPhotTileRows][2] boundary states of that solve
  double* tab = bnd + 2 * kPhotTileRows * (kPhotThreads / 32);   // [1024]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kExpTabSize; i += blockDim.x) tab[i] = RB_EXP2_1024[i];
  for (int j = tid; j < NL; j += blockDim.x) load_lu_row(lul + 5 * j, a.lu_lambda + 5 * j);
  double* S = a.scratch + (size_t)blockIdx.x * NT * NL;
  __syncthreads();
  // segments of the partitioned wavelength-axis solve (at least 2 columns each) and their homogeneous solutions:
  // forward  y[j] = -l2 y[j-2] - l1 y[j-1] from (y[m-1], y[m-2]) = (1, 0) and (0, 1) at the segment start m,
  // backward x[j] = -u2' x[j+2] - u1' x[j+1] from (x[m'+1], x[m'+2]) = (1, 0) and (0, 1) at the segment end m'
  const int seg_len = max((NL + (int)(blockDim.x >> 5) - 1) / (int)(blockDim.x >> 5), 2);
  const int nseg = (NL + seg_len - 1) / seg_len;
  if (tid < 4 * nseg) {
    const int k = tid >> 2, which = tid & 3;
    const int j_lo = k * seg_len, j_hi = min(j_lo + seg_len, NL);
    double p1 = (which & 1) ? 0.0 : 1.0, p2 = (which & 1) ? 1.0 : 0.0;
    if (which < 2) {
      for (int j = j_lo; j < j_hi; ++j) {
        const double y = -lul[5 * j + 0] * p2 - lul[5 * j + 1] * p1;
        hom[4 * j + which] = y;
        p2 = p1;
        p1 = y;
      }
    } else {
      for (int j = j_hi - 1; j >= j_lo; --j) {
        const double x = -lul[5 * j + 4] * p2 - lul[5 * j + 3] * p1;
        hom[4 * j + which] = x;
        p2 = p1;
        p1 = x;
      }
    }
  }
  __syncthreads();

  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    PHOT_PROF_BEGIN();
    const int G = a.grid_len ? a.grid_len[n] : NT;
    const double* gT = a.grid + (size_t)n * kPhotGridCols * NT + PH_T * NT;
    const double* gR = a.grid + (size_t)n * kPhotGridCols * NT + PH_R * NT;
    const double* gL = a.grid + (size_t)n * kPhotGridCols * NT + PH_L * NT;
    const double* gX = a.grid + (size_t)n * kPhotGridCols * NT + PH_X * NT;
    const double opz = a.opz[n], dl = a.dl[n];
    const double lam_c = a.cutoff_wavelength * kAngstrom;
    // Every spectrum sample factorises into a per-epoch and a per-wavelength part around the Planck denominator:
    //   v[k][j] = colB[j] * rowB[k] / expm1(colA[j] * rowA[k])   (+ the sed.Line terms for type Ia),
    // so the divisions, powers and unit conversions are done once per row / column (they live in the tile region,
    // which is free until the solves) and a sample costs one expm1, one division and three multiplies.
    double* rowA = tile;             // [NC][NT]  1 / (k_B T)  or 1 / T
    double* rowB = rowA + NC * NT;   // [NC][NT]  R^2          or R^2 * cutoff normalisation
    double* rowC = rowB + NC * NT;   // [NT]  1 - line amplitude
    double* rowD = rowC + NT;        // [NT]  scaled line amplitude
    double* colA = rowD + NT;        // [NL]  h nu         or hc / (k_B lambda)
    double* colB = colA + NL;        // [NL]  everything else that depends on wavelength only
    double* colC = colB + NL;        // [NL]  line profile in output units
    const bool line = a.sed == PH_SED_CUTOFF_LINE;
    for (int k = tid; k < G; k += blockDim.x) {
      xk[k] = gX[k];
      const double T = gT[k], R = gR[k];
      if (a.sed == PH_SED_BLACKBODY) {
        rowA[k] = 1.0 / (kKB * T);
        rowB[k] = R * R;
        for (int c = 1; c < NC; ++c) {                         // kilonova_models.py:1150-1175, 1278-1300
          const double Tc = gT[c * a.grid_comp_stride + k], Rc = gR[c * a.grid_comp_stride + k];
          rowA[c * NT + k] = 1.0 / (kKB * Tc);
          rowB[c * NT + k] = Rc * Rc;
        }
      } else {
        rowA[k] = 1.0 / T;
        rowB[k] = (R * R) * cutoff_norm(gL[k], R, T, lam_c, a.absorption_index);
        if (line) {
          // sed.Line on the (N_lambda, N_t) grid (sed.py:859-909; supernova_models.py:2295-2303); time is source frame
          const double te = gX[k] / opz;
          const double q = (te - a.line_time) / a.line_duration;
          const double amp = a.line_amplitude * exp(-0.5 * (q * q));
          double amp_scaled = amp * gL[k] / (4 * kPi * (dl * dl));
          amp_scaled /= (a.line_width * sqrt(2 * kPi));
          rowC[k] = 1 - amp;
          rowD[k] = amp_scaled;
        }
      }
    }
    for (int j = tid; j < NL; j += blockDim.x) {
      const double lam_obs = a.lambda_obs[j];
      const double nu = (kCSI / (lam_obs * 1.e-10)) * opz;                       // utils.py:357-362, 383
      const double to_flam = 1e-26 * 2.99792458e18 / (lam_obs * lam_obs);        // mJy -> erg/cm^2/s/Angstrom
      if (a.sed == PH_SED_BLACKBODY) {
        colA[j] = kH * nu;
        colB[j] = 2 * kPi * kH * (nu * nu * nu) / ((dl * dl) * (kC * kC)) * opz * kToMJy * to_flam;   // sed.py:216, 929
      } else {
        const double lam_A = 1.e10 * (kCSI / nu);
        const double lam = lam_A * kAngstrom;
        const double l2 = lam * lam, l5 = l2 * l2 * lam;
        const double absorb = (lam < lam_c) ? pow(lam / lam_c, a.absorption_index) : 1.0;
        colA[j] = kXConst / lam;
        double c = kFluxConst * (absorb / l5) / (4 * kPi * (dl * dl)) * lam_A / nu * kToMJy;        // sed.py:362-384
        if (line) {
          const double u = (lam_A - a.line_wavelength) / a.line_width;
          colC[j] = exp(-0.5 * (u * u)) * (lam * lam) / kC * kToMJy * opz * to_flam;
          c = c * 1e-26 * kToMJy;
        }
        colB[j] = c * opz * to_flam;                                              // supernova_models.py:1611-1612, 2304-2305
      }
    }
    if (a.common_time_lu)
      for (int i = tid; i < G; i += blockDim.x) load_lu_row(lut + 5 * i, a.lu_time + 5 * i);
    __syncthreads();

    PHOT_PROF(0);
    // ---- 1. spectra + statistics for the clean-up -------------------------------------------------------
    double vsum = 0.0, vcnt = 0.0;
    {
      int k = tid / NL, j = tid - k * NL;
      const int dk = (int)blockDim.x / NL, dj = (int)blockDim.x - dk * NL;
      for (; k < G; k += dk) {
        double v = colB[j] * rowB[k] / expm1_fast(colA[j] * rowA[k], tab);
        for (int c = 1; c < NC; ++c) v += colB[j] * rowB[c * NT + k] / expm1_fast(colA[j] * rowA[c * NT + k], tab);
        if (line) v = v * rowC[k] + rowD[k] * colC[j];
        S[k * NL + j] = v;
        if (isfinite(v) && v != 0.0) { vsum += v; vcnt += 1.0; }
        j += dj;
        if (j >= NL) { j -= NL; ++k; }
      }
    }
    __syncthreads();   // the tile region is reused by the solves below
    PHOT_PROF(1);
    // ---- 2. clean-up value (sed.py:985-992) ---------------------------------------------------------------
    vsum = warp_sum(vsum);
    vcnt = warp_sum(vcnt);
    if (lane == 0) { red[warp] = vsum; red[8 + warp] = vcnt; }
    __syncthreads();
    if (tid == 0) {
      double s = 0.0, c = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += red[w]; c += red[8 + w]; }
      double mean = (c > 0) ? s / c : NAN;
      if (!isfinite(mean) || mean == 0.0) mean = 1.0;
      red[32] = 1e-30 * mean;
    }
    // ---- 3a. time-axis collocation + LU when the grid is per-sample ----------------------------------------
    if (!a.common_time_lu) {
      for (int i = tid; i < G; i += blockDim.x) {
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
        for (int q = 0; q < 5; ++q) lut[5 * i + q] = row[q];
      }
      __syncthreads();
      if (tid == 0) {
        // banded LU without pivoting, in place, stored in the scaled layout of load_lu_row.  The two previous rows
        // stay in registers and the next row is loaded ahead, so the serial chain per row is multiply, FMA, reciprocal.
        double rd1 = 0.0, u1_1 = 0.0, u2_1 = 0.0, rd2 = 0.0, u1_2 = 0.0, u2_2 = 0.0;
        double n0 = lut[0], n1 = lut[1], n2 = lut[2], n3 = lut[3], n4 = lut[4];
        for (int i = 0; i < G; ++i) {
          const double a0 = n0, a1 = n1, u2 = n4;
          double d = n2, u1 = n3;
          if (i + 1 < G) { n0 = lut[5 * i + 5]; n1 = lut[5 * i + 6]; n2 = lut[5 * i + 7]; n3 = lut[5 * i + 8]; n4 = lut[5 * i + 9]; }
          const double l2 = a0 * rd2;
          const double am1 = a1 - l2 * u1_2;
          d -= l2 * u2_2;
          const double l1 = am1 * rd1;
          d -= l1 * u1_1;
          u1 -= l1 * u2_1;
          const double rd = 1.0 / d;
          lut[5 * i + 0] = l2;
          lut[5 * i + 1] = l1;
          lut[5 * i + 2] = rd;
          lut[5 * i + 3] = u1 * rd;
          lut[5 * i + 4] = u2 * rd;
          rd2 = rd1; u1_2 = u1_1; u2_2 = u2_1;
          rd1 = rd; u1_1 = u1; u2_1 = u2;
        }
      }
    }
    __syncthreads();
    const double repl = red[32];

    PHOT_PROF(2);
    // ---- 3b. time-axis forward elimination: one thread per wavelength column (coalesced across columns).  The scratch tile
    // lives in global memory / L2: loads are issued kPhotBatch rows ahead of the recurrence to hide latency.
    for (int j = tid; j < NL; j += blockDim.x) {
      double y1 = 0.0, y2 = 0.0;   // y[i-1], y[i-2]
      double nb[kPhotBatch];       // the batch after the one being eliminated: its loads overlap the recurrence
#pragma unroll
      for (int u = 0; u < kPhotBatch; ++u) nb[u] = (u < G) ? S[u * NL + j] : 0.0;
      for (int i0 = 0; i0 < G; i0 += kPhotBatch) {
        double b[kPhotBatch];
#pragma unroll
        for (int u = 0; u < kPhotBatch; ++u) {
          b[u] = nb[u];
          nb[u] = (i0 + kPhotBatch + u < G) ? S[(i0 + kPhotBatch + u) * NL + j] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < kPhotBatch; ++u) {
          const int i = i0 + u;
          if (i < G) {
            double v = b[u];
            if (!(isfinite(v) && v != 0.0)) v = repl;
            const double y = (v - lut[5 * i + 0] * y2) - lut[5 * i + 1] * y1;   // y[i-2] term first: one FMA on the chain
            S[i * NL + j] = y;
            y2 = y1;
            y1 = y;
          }
        }
      }
    }
    for (int j = tid; j < NL; j += blockDim.x) bst[2 * j] = bst[2 * j + 1] = 0.0;   // x[i+1], x[i+2] per column
    __syncthreads();
    PHOT_PROF(3);
    // ---- 3c. time-axis back substitution fused with the wavelength-axis solve: the rows come out last-to-first, so
    // each tile of kPhotTileRows finished rows goes straight into shared memory, is solved along wavelength there and
    // is written to the scratch once (one read and one write pass fewer than solving the two axes separately)
    PHOT_PROF_ACC(11);
    // with one column per thread the first batch of the next tile is loaded before the wavelength-axis solve of the
    // current one, so no tile starts with an exposed round trip to the scratch
    const bool one_col = NL <= (int)blockDim.x;
    double nbx[kPhotBatch];
#pragma unroll
    for (int u = 0; u < kPhotBatch; ++u)
      nbx[u] = (one_col && tid < NL && G - 1 - u >= max(G - kPhotTileRows, 0)) ? S[(G - 1 - u) * NL + tid] : 0.0;
    for (int r1 = G; r1 > 0; r1 -= kPhotTileRows) {
      const int r0 = max(r1 - kPhotTileRows, 0), rows = r1 - r0;
      for (int j = tid; j < NL; j += blockDim.x) {
        double x1 = bst[2 * j], x2 = bst[2 * j + 1];
        double nb[kPhotBatch];
#pragma unroll
        for (int u = 0; u < kPhotBatch; ++u)
          nb[u] = one_col ? nbx[u] : ((r1 - 1 - u >= r0) ? S[(r1 - 1 - u) * NL + j] : 0.0);
        for (int i0 = r1 - 1; i0 >= r0; i0 -= kPhotBatch) {
          double b[kPhotBatch];
#pragma unroll
          for (int u = 0; u < kPhotBatch; ++u) {
            b[u] = nb[u];
            nb[u] = (i0 - kPhotBatch - u >= r0) ? S[(i0 - kPhotBatch - u) * NL + j] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < kPhotBatch; ++u) {
            const int i = i0 - u;
            if (i >= r0) {
              const double x = (b[u] * lut[5 * i + 2] - lut[5 * i + 4] * x2) - lut[5 * i + 3] * x1;
              tile[(i - r0) * (NL + 1) + j] = x;
              x2 = x1;
              x1 = x;
            }
          }
        }
        bst[2 * j] = x1;
        bst[2 * j + 1] = x2;
        if (one_col) {
          const int q0 = max(r0 - kPhotTileRows, 0);
#pragma unroll
          for (int u = 0; u < kPhotBatch; ++u) nbx[u] = (r0 - 1 - u >= q0) ? S[(r0 - 1 - u) * NL + j] : 0.0;
        }
      }
      __syncthreads();
      PHOT_PROF_ACC(8);
      // Wavelength-axis solve of the tile, partitioned: each row is cut into one segment per warp (lane = row), every
      // segment runs the recurrence from a zero state, and the true solution is that particular solution plus the
      // segment's two homogeneous solutions (hom, set up once per block) scaled by the true state at the segment
      // boundary.  The boundary states follow from a short serial pass over the segments; the serial chain per tile
      // drops from 2 NL steps to 2 NL / nwarps.
      const int seg = warp, row_i = lane;
      const int j_lo = min(seg * seg_len, NL), j_hi = min(j_lo + seg_len, NL);
      double* row = tile + row_i * (NL + 1);
      if (row_i < rows) {                                      // forward elimination, particular solution
        double y1 = 0.0, y2 = 0.0;
        for (int j0 = j_lo; j0 < j_hi; j0 += kPhotBatch) {
          double b[kPhotBatch], c0[kPhotBatch], c1[kPhotBatch];
#pragma unroll
          for (int u = 0; u < kPhotBatch; ++u) {
            const int j = min(j0 + u, NL - 1);
            b[u] = row[j];
            c0[u] = lul[5 * j + 0];
            c1[u] = lul[5 * j + 1];
          }
#pragma unroll
          for (int u = 0; u < kPhotBatch; ++u) {
            if (j0 + u < j_hi) {
              const double y = (b[u] - c0[u] * y2) - c1[u] * y1;
              row[j0 + u] = y;
              y2 = y1;
              y1 = y;
            }
          }
        }
      }
      __syncthreads();
      if (warp == 0 && row_i < rows) {                         // true (y[j-1], y[j-2]) at the start of every segment
        double Y1 = 0.0, Y2 = 0.0;
        for (int k = 1; k < nseg; ++k) {
          const int m = k * seg_len;                           // first column of segment k; segment k-1 is [m - seg_len, m)
          const double e1 = row[m - 1] + hom[4 * (m - 1) + 0] * Y1 + hom[4 * (m - 1) + 1] * Y2;
          const double e2 = row[m - 2] + hom[4 * (m - 2) + 0] * Y1 + hom[4 * (m - 2) + 1] * Y2;
          Y1 = e1;
          Y2 = e2;
          bnd[(k * kPhotTileRows + row_i) * 2 + 0] = Y1;
          bnd[(k * kPhotTileRows + row_i) * 2 + 1] = Y2;
        }
      }
      __syncthreads();
      if (row_i < rows && j_lo < j_hi) {                       // back substitution, particular solution
        const double Y1 = seg > 0 ? bnd[(seg * kPhotTileRows + row_i) * 2 + 0] : 0.0;
        const double Y2 = seg > 0 ? bnd[(seg * kPhotTileRows + row_i) * 2 + 1] : 0.0;
        double x1 = 0.0, x2 = 0.0;
        for (int j0 = j_hi - 1; j0 >= j_lo; j0 -= kPhotBatch) {
          double b[kPhotBatch], c3[kPhotBatch], c4[kPhotBatch];
#pragma unroll
          for (int u = 0; u < kPhotBatch; ++u) {
            const int j = max(j0 - u, 0);
            b[u] = (row[j] + hom[4 * j + 0] * Y1 + hom[4 * j + 1] * Y2) * lul[5 * j + 2];   // forward fix-up folded in
            c3[u] = lul[5 * j + 3];
            c4[u] = lul[5 * j + 4];
          }
#pragma unroll
          for (int u = 0; u < kPhotBatch; ++u) {
            if (j0 - u >= j_lo) {
              const double x = (b[u] - c4[u] * x2) - c3[u] * x1;
              row[j0 - u] = x;
              x2 = x1;
              x1 = x;
            }
          }
        }
      }
      __syncthreads();
      if (warp == 0 && row_i < rows) {                         // true (x[j+1], x[j+2]) at the end of every segment
        double X1 = 0.0, X2 = 0.0;
        for (int k = nseg - 2; k >= 0; --k) {
          const int m = (k + 1) * seg_len;                     // first column of segment k + 1
          const double s1 = row[m] + hom[4 * m + 2] * X1 + hom[4 * m + 3] * X2;
          const double s2 = (m + 1 < NL) ? row[m + 1] + hom[4 * (m + 1) + 2] * X1 + hom[4 * (m + 1) + 3] * X2 : 0.0;
          X1 = s1;
          X2 = s2;
          bnd[(k * kPhotTileRows + row_i) * 2 + 0] = X1;
          bnd[(k * kPhotTileRows + row_i) * 2 + 1] = X2;
        }
      }
      __syncthreads();
      PHOT_PROF_ACC(9);
      for (int j = tid; j < NL; j += blockDim.x) {                // backward fix-up folded into the store
        const int k = j / seg_len;
        const double hA = (k < nseg - 1) ? hom[4 * j + 2] : 0.0, hB = (k < nseg - 1) ? hom[4 * j + 3] : 0.0;
        const double* bk = bnd + (min(k, nseg - 2 >= 0 ? nseg - 2 : 0) * kPhotTileRows) * 2;
#pragma unroll 4
        for (int r = 0; r < rows; ++r)
          S[(r0 + r) * NL + j] = tile[r * (NL + 1) + j] + hA * bk[2 * r] + hB * bk[2 * r + 1];
      }
      __syncthreads();
      PHOT_PROF_ACC(10);
    }

    PHOT_PROF(5);
    // ---- 4./5. epochs: bispev at (t_e, band grid), band flux, magnitude, likelihood -------------------------
    // Three sub-phases per chunk of epochs, each with the thread mapping that suits it (the tile region is free again
    // and holds the per-epoch hand-over arrays):
    //   a. one thread per epoch: knot interval and the four non-zero B-splines along time (divisions, a search);
    //   b. one warp per epoch: gather the band's wavelength coefficients from the scratch with coalesced loads into a
    //      per-warp buffer and split the band-pass integration points between the lanes;
    //   c. one thread per epoch: magnitude, likelihood term, coalesced store.
    double logl_part = 0.0;
    const int nwarps = (int)(blockDim.x >> 5);
    double* dw = tile + warp * kPhotMaxSpan;
    double* eh = tile + nwarps * kPhotMaxSpan;                 // [chunk][kEpRec]: h0..h3, first row (< 0: before t0), band
                                                               // flux, then the band's jlo, span, first and last point
    const int chunk = (phot_tile_elems(NT, NL, NC) - nwarps * kPhotMaxSpan) / kEpRec;
    const double sp = a.sigma_param_s ? a.sigma_param_s[n] : 0.0;
    PHOT_PROF_ACC(11);
    for (int e0 = 0; e0 < E; e0 += chunk) {
      const int ne = min(chunk, E - e0);
      for (int i = tid; i < ne; i += blockDim.x) {
        const int e = a.epoch_order[e0 + i];
        double x = a.epoch_tab[EP_T * E + e];
        double* o = eh + kEpRec * i;
        const int b = a.band_of_epoch[e];                      // looked up here so that the warps of sub-phase b do not
        o[6] = (double)a.band_jlo[b];                          // start each epoch with a chain of dependent global loads
        o[7] = (double)a.band_span[b];
        o[8] = (double)a.band_off[b];
        o[9] = (double)a.band_off[b + 1];
        if (a.t0_s != nullptr) {                                                  // phase_models.py:33-58
          x -= a.t0_s[n];
          if (x < 0.0) { o[0] = o[1] = o[2] = o[3] = 0.0; o[4] = -1.0; continue; }
        }
        if (x < xk[0]) x = xk[0];                                                 // fpbisp clamps to [tb, te]
        if (x > xk[G - 1]) x = xk[G - 1];
        // knot interval: largest l in [3, G-1] with t[l] <= x  (t[l] = x[l-2] for 4 <= l <= G-1)
        int l = 3;
        {
          int lo = 4, hi = G - 1;   // find count of knots t[4..G-1] = xk[2..G-3] that are <= x
          while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            if (xk[mid - 2] <= x) { l = mid; lo = mid + 1; } else hi = mid - 1;
          }
        }
        double tl[8], h[4];
        for (int q = 0; q < 8; ++q) tl[q] = nak_knot(xk, G, l - 3 + q);
        fpbspl3(tl, 3, x, h);
        o[0] = h[0]; o[1] = h[1]; o[2] = h[2]; o[3] = h[3];
        o[4] = (double)(l - 3);
      }
      __syncthreads();
      PHOT_PROF_ACC(6);
      // The scratch is larger than L2 across the grid, so every gather may be a DRAM round trip: all its loads are
      // issued together, and those of the warp's next epoch before the current one is integrated.
      constexpr int kGather = kPhotMaxSpan / 32;
      double g0[kGather], g1[kGather], g2[kGather], g3[kGather];
      auto gather = [&](int i) {
        if (i >= ne) return;
        const double* o = eh + kEpRec * i;
        const int span = (int)o[7];
        const double* Sx = S + (size_t)max((int)o[4], 0) * NL + (int)o[6];
#pragma unroll
        for (int u = 0; u < kGather; ++u) {
          const int q = min(lane + 32 * u, span - 1);
          g0[u] = Sx[q];
          g1[u] = Sx[NL + q];
          g2[u] = Sx[2 * NL + q];
          g3[u] = Sx[3 * NL + q];
        }
      };
      gather(warp);
      for (int i = warp; i < ne; i += nwarps) {
        const double* o = eh + kEpRec * i;
        const int span = (int)o[7];
        const double h0 = o[0], h1 = o[1], h2 = o[2], h3 = o[3];
#pragma unroll
        for (int u = 0; u < kGather; ++u) {
          const int q = lane + 32 * u;
          if (q < span) dw[q] = h0 * g0[u] + h1 * g1[u] + h2 * g2[u] + h3 * g3[u];
        }
        gather(i + nwarps);
        if (o[4] < 0.0) continue;                              // before t0: no spectrum to integrate
        __syncwarp();
        double acc = 0.0;
        // four integration points per lane and trip, indices first, then weights, then coefficients: the loads of a
        // trip overlap instead of forming one dependent chain per point
        const int w0 = (int)o[8], w1 = (int)o[9];
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
        __syncwarp();
        if (lane == 0) eh[kEpRec * i + 5] = acc;
      }
      __syncthreads();
      PHOT_PROF_ACC(7);
      for (int i = tid; i < ne; i += blockDim.x) {
        const int e = a.epoch_order[e0 + i];
        double model_v;
        if (eh[kEpRec * i + 4] < 0.0) {
          model_v = (a.output == RB_OUT_FLUX) ? 0.0 : 1000.0;
        } else {
          const int b = a.band_of_epoch[e];
          const double bandflux = eh[kEpRec * i + 5] * a.band_scale[b];
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
      __syncthreads();
      if (lane == 0) red[warp] = logl_part;
      __syncthreads();
      if (tid == 0) {
        double total = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += red[w];
        a.out_logl[a.n_base + n] = nan_to_num(total);
      }
    }
    __syncthreads();
    PHOT_PROF(6);
    PHOT_PROF_END();
  }
}

}  // namespace rb
