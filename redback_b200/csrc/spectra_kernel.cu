// output_format='spectra' (supernova_models.py:1019-1023 and its twins in every model family; the namedtuple
// (time, lambdas, spectra) that extinction_models.py:223-235 and user code consume): the observer-frame spectra
// erg / cm^2 / s / Angstrom on the model's own time grid x lambda_array, BEFORE the clean-up of
// get_correct_output_format_from_spectra (sed.py:985-992).  Fed by the same grid-mode model kernels as phot_kernel;
// unlike phot_kernel this kernel materialises the N_t x N_lambda tile (it is the requested output).
// Block per sample (grid-stride); row / column factors exactly as phot_kernel's set-up, libm expm1 per entry.
// (included by capi.cu after phot_kernel.cu)

namespace rb {

struct SpectraArgs {
  int64_t N;              // samples in this chunk
  int n_t_max, n_lambda, sed, n_comp;
  int64_t grid_comp_stride;
  double cutoff_wavelength, absorption_index;
  double line_wavelength, line_width, line_time, line_duration, line_amplitude;
  const double* grid;     // [N][kPhotGridCols][n_t_max]
  const int* grid_len;    // [N] or nullptr
  const double* opz;      // [N]
  const double* dl;       // [N]
  const double* lambda_obs;
  const rb_extinction* ext;
  const double* ext_av_host_s;
  double* spectra_out;    // [N][n_t_max][n_lambda]   (rows >= grid_len[n] are NaN)
  double* time_out;       // [N][n_t_max]             observer-frame days (NaN beyond grid_len[n])
  int* len_out;           // [N]
};

__global__ void __launch_bounds__(256) spectra_kernel(const SpectraArgs a) {
  extern __shared__ __align__(16) double ssm[];
  const int NT = a.n_t_max, NL = a.n_lambda, NC = a.n_comp;
  const int T = (int)blockDim.x, tid = threadIdx.x;
  const bool line = a.sed == PH_SED_CUTOFF_LINE;
  double* rowA = ssm;                    // [NC][NT]
  double* rowB = rowA + NC * NT;         // [NC][NT]
  double* rowC = rowB + NC * NT;         // [NT] (line)
  double* rowD = rowC + (line ? NT : 0); // [NT] (line)
  double* colA = rowD + (line ? NT : 0); // [NL]
  double* colB = colA + NL;
  double* colC = colB + NL;
  double* colD = colC + NL;
  const double lam_c = a.cutoff_wavelength * kAngstrom;
  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    const int G = a.grid_len ? a.grid_len[n] : NT;
    const double opz = a.opz[n], dl = a.dl[n];
    const double* g0 = a.grid + (size_t)n * kPhotGridCols * NT;
    const double *gT = g0 + PH_T * NT, *gR = g0 + PH_R * NT, *gL = g0 + PH_L * NT, *gX = g0 + PH_X * NT;
    __syncthreads();
    for (int k = tid; k < G; k += T) {
      if (a.sed == PH_SED_BLACKBODY) {
        for (int c = 0; c < NC; ++c) {
          const double Tc = gT[c * a.grid_comp_stride + k], Rc = gR[c * a.grid_comp_stride + k];
          rowA[c * NT + k] = 1.0 / (kKB * Tc);
          rowB[c * NT + k] = Rc * Rc;
        }
      } else {
        const double Tk = gT[k], R = gR[k];
        rowA[k] = 1.0 / Tk;
        rowB[k] = (R * R) * cutoff_norm(gL[k], R, Tk, lam_c, a.absorption_index);
        if (line) {                                                             // sed.py:859-909
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
    for (int j = tid; j < NL; j += T) {
      const double lam_obs = a.lambda_obs[j];
      const double nu = (kCSI / (lam_obs * 1.e-10)) * opz;                       // utils.py:357-362, 383
      const double to_flam = 1e-26 * 2.99792458e18 / (lam_obs * lam_obs);        // mJy -> erg/cm^2/s/Angstrom
      double ca, cb;
      colC[j] = 0.0;
      if (a.sed == PH_SED_BLACKBODY) {
        ca = kH * nu;
        cb = 2 * kPi * kH * (nu * nu * nu) / ((dl * dl) * (kC * kC)) * opz * kToMJy * to_flam;   // sed.py:216, 929
        colD[j] = 2 * kPi * kH * (nu * nu * nu);
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
        cb = c * opz * to_flam;
        colD[j] = 0.0;
      }
      if (a.ext != nullptr) {
        const double att = ext_factor(*a.ext, lam_obs, opz, a.ext_av_host_s[n]);
        cb *= att;
        colC[j] *= att;
      }
      colA[j] = ca;
      colB[j] = cb;
    }
    __syncthreads();
    const double inv_den = 1.0 / ((dl * dl) * (kC * kC));                        // sed.py:218
    double* out = a.spectra_out + (size_t)n * NT * NL;
    for (int idx = tid; idx < NT * NL; idx += T) {
      const int k = idx / NL, j = idx - k * NL;
      double v = NAN;
      if (k < G) {
        v = 0.0;
        for (int c = 0; c < NC; ++c) {
          const double x = colA[j] * rowA[c * NT + k];
          double vc = colB[j] * rowB[c * NT + k] * (1.0 / expm1(x));
          // the reference's erg/s/Hz/cm^2 product underflows to 0 earlier than ours in output units (sed.py:217-221)
          if (a.sed == PH_SED_BLACKBODY && x > 650.0 && ((colD[j] * rowB[c * NT + k]) * inv_den) * (1.0 / expm1(x)) == 0.0)
            vc = 0.0;
          v += vc;
        }
        if (line) v = v * rowC[k] + rowD[k] * colC[j];
      }
      out[idx] = v;
    }
    for (int k = tid; k < NT; k += T) a.time_out[(size_t)n * NT + k] = (k < G) ? gX[k] : NAN;
    if (tid == 0) a.len_out[n] = G;
  }
}

}  // namespace rb
