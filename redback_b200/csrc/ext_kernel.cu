// Dust extinction (redback/transient_models/extinction_models.py:77-170, `_perform_extinction`): host-galaxy extinction
// at the rest wavelength lambda / (1 + z) and Milky-Way extinction at the observed wavelength, flux * 10^(-0.4 A).
// The laws are those of the third-party `extinction` package (kbarbary/extinction, not vendored in the reference and
// not installable here; restated from the published papers and anchored on the package's documented example values,
// tests/test_extinction.py; odonnell94 / calzetti00 have no such vector: PARITY UNPINNED):
//   ccm89          Cardelli, Clayton & Mathis 1989, A(lambda) = A_V (a(x) + b(x) / R_V)
//   odonnell94     the same with O'Donnell's 1994 optical coefficients
//   calzetti00     Calzetti et al. 2000, A = A_V (1 + k(x) / R_V), R_V = 4.05 (redback never passes another value)
//   fitzpatrick99  Fitzpatrick 1999: Fitzpatrick & Massa 1990 in the UV (lambda < 2700 A), cubic spline through nine
//                  anchor points at longer wavelengths; the NATURAL spline is solved on the host
//                  (extinction_models.f99_spline_table) and arrives as piecewise-cubic coefficients
// x is in inverse microns.  `_perform_extinction` zeroes the extinction where it exceeds 10 mag while A_V < 10.
// (included by capi.cu before phot_kernel.cu, which applies the same factor per wavelength column)

namespace rb {

__device__ __forceinline__ void ccm89_ab(double x, bool odonnell, double& a, double& b) {
  if (x < 1.1) {                       // infrared
    const double y = pow(x, 1.61);
    a = 0.574 * y;
    b = -0.527 * y;
  } else if (x < 3.3) {                // optical / near infrared
    const double y = x - 1.82;
    if (odonnell) {
      a = ((((((( -0.505 * y + 1.647) * y - 0.827) * y - 1.718) * y + 1.137) * y + 0.701) * y - 0.609) * y + 0.104) * y + 1.0;
      b = (((((((3.347 * y - 10.805) * y + 5.491) * y + 11.102) * y - 7.985) * y - 3.989) * y + 2.908) * y + 1.952) * y;
    } else {
      a = ((((((0.329990 * y - 0.77530) * y + 0.01979) * y + 0.72085) * y - 0.02427) * y - 0.50447) * y + 0.17699) * y + 1.0;
      b = ((((((-2.09002 * y + 5.30260) * y - 0.62251) * y - 5.38434) * y + 1.07233) * y + 2.28305) * y + 1.41338) * y;
    }
  } else if (x < 8.0) {                // ultraviolet
    double y = x - 4.67;
    a = 1.752 - 0.316 * x - (0.104 / (y * y + 0.341));
    y = x - 4.62;
    b = -3.090 + 1.825 * x + (1.206 / (y * y + 0.263));
    if (x > 5.9) {
      y = x - 5.9;
      const double y2 = y * y, y3 = y2 * y;
      a += -0.04473 * y2 - 0.009779 * y3;
      b += 0.2130 * y2 + 0.1207 * y3;
    }
  } else {                             // far ultraviolet
    const double y = x - 8.0, y2 = y * y, y3 = y2 * y;
    a = -0.070 * y3 + 0.137 * y2 - 0.628 * y - 1.073;
    b = 0.374 * y3 - 0.420 * y2 + 4.257 * y + 13.670;
  }
}

// Fitzpatrick & Massa 1990 ultraviolet curve with Fitzpatrick's 1999 parameters: k(x) = E(lambda - V) / E(B - V)
__host__ __device__ inline double f99_uv_k(double x, double rv) {
  const double x0 = 4.596, gamma = 0.99, c3 = 3.23, c4 = 0.41;
  const double c2 = -0.824 + 4.717 / rv, c1 = 2.030 - 3.007 * c2;
  const double x2 = x * x, d0 = x2 - x0 * x0;
  const double drude = x2 / (d0 * d0 + x2 * gamma * gamma);
  double k = c1 + c2 * x + c3 * drude;
  if (x >= 5.9) {
    const double y = x - 5.9, y2 = y * y;
    k += c4 * (0.5392 * y2 + 0.05644 * y2 * y);
  }
  return k;
}

// A(lambda) in magnitudes.  `tab`: fitzpatrick99 only -- 9 breakpoints then 8 x 4 coefficients (highest power first) of
// the spline pieces in (x - breakpoint).
__device__ __forceinline__ double ext_mag(int law, double wave_angstrom, double av, double rv, const double* tab) {
  const double x = 1e4 / wave_angstrom;
  double a_lam;
  if (law == RB_EXT_CCM89 || law == RB_EXT_ODONNELL94) {
    double a, b;
    ccm89_ab(x, law == RB_EXT_ODONNELL94, a, b);
    a_lam = av * (a + b / rv);
  } else if (law == RB_EXT_CALZETTI00) {
    const double k = (x > 1.5873015873015872) ? 2.659 * (((0.011 * x - 0.198) * x + 1.509) * x - 2.156)
                                               : 2.659 * (1.040 * x - 1.857);
    a_lam = av * (1.0 + k / rv);
  } else {
    double k;
    if (x >= 1e4 / 2700.0) {
      k = f99_uv_k(x, rv);
    } else {
      int i = 0;
      while (i < 7 && x >= tab[i + 1]) ++i;
      const double dx = x - tab[i];
      const double* c = tab + 9 + 4 * i;
      k = ((c[0] * dx + c[1]) * dx + c[2]) * dx + c[3];
    }
    a_lam = av / rv * (k + rv);
  }
  if (av < 10.0 && a_lam > 10.0) a_lam = 0.0;                                   // extinction_models.py:137-139, :160-162
  return a_lam;
}

// total attenuation factor 10^(-0.4 (A_host + A_MW)) at observed wavelength `wave` [Angstrom] for a source at 1 + z = opz
__device__ __forceinline__ double ext_factor_inline(const rb_extinction& e, double wave, double opz, double av_host) {
  double mag = 0.0;
  if (av_host > 0.0) mag += ext_mag(e.host_law, wave / opz, av_host, e.rv_host, e.f99_host);   // :127-142
  if (e.av_mw > 0.0) mag += ext_mag(e.mw_law, wave, e.av_mw, e.rv_mw, e.f99_mw);               // :145-165
  return pow(10.0, -0.4 * mag);                                                                // extinction.apply
}

// out-of-line for phot_kernel (per wavelength column, off its hot loops; `e` in global memory)
__device__ __noinline__ double ext_factor(const rb_extinction& e, double wave, double opz, double av_host) {
  return ext_factor_inline(e, wave, opz, av_host);
}

// flux-density branch (extinction_models.py:194-221): model[n][e] *= factor(lambda_e, z_n, av_host_n), in place
__global__ void ext_apply_kernel(const rb_extinction e, int64_t N, int E, const double* __restrict__ freq_obs,
                                 const double* __restrict__ redshift, const double* __restrict__ av_host,
                                 double* __restrict__ model) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * (int64_t)E) return;
  const int64_t n = i / E;
  const int ep = (int)(i - n * E);
  const double wave = 1.e10 * (kCSI / freq_obs[ep]);                            // utils.nu_to_lambda, :199
  model[i] *= ext_factor_inline(e, wave, 1.0 + redshift[n], av_host[n]);
}

// sed.Synchrotron (sed.py:703-751) + _SED.flux_density (sed.py:239-251) as type_1c adds it to the arnett flux density
// (supernova_models.py:2343-2352): model[n][e] += mJy(nu_e (1 + z_n), pp_n, dl_n) (1 + z_n), in place.  Epoch-independent.
__global__ void synchrotron_add_kernel(int64_t N, int E, const double* __restrict__ freq_obs,
                                       const double* __restrict__ redshift, const double* __restrict__ pp,
                                       const double* __restrict__ dl, double nu_max, double f0, double source_radius,
                                       double* __restrict__ model) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * (int64_t)E) return;
  const int64_t n = i / E;
  const int ep = (int)(i - n * E);
  const double opz = 1.0 + redshift[n];
  const double nu = freq_obs[ep] * opz;                                        // utils.py:357-384
  const double r2 = source_radius * source_radius;
  double sed;
  if (nu < nu_max) sed = f0 * r2 * pow(nu / nu_max, 2.5) * kAngstrom / kC * (nu * nu);                         // :742-744
  else sed = (f0 * r2 * pow(nu_max, 2.5)) * pow(nu / nu_max, -(pp[n] - 1.) / 2.) * kAngstrom / kC * (nu * nu);  // :745-747
  const double d = dl[n];
  const double mjy = sed / (4 * kPi * (d * d)) * (1.e10 * (kCSI / nu)) / nu * kToMJy;
  model[i] += mjy * opz;                                                       // :2351-2352
}

}  // namespace rb
