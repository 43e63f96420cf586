// redback_b200 -- observation noise + SNR cut for simulated optical light curves, batched over transients.
// Reference: SimulateTransientWithCadence._evaluate_optical_model / _add_optical_noise_and_cuts
// (redback/simulate_transients.py:412-431, 457-527) with redback/utils.py:707-718 (magnitude -> flux),
// :834-845 (flux -> magnitude), :808-831 (magnitude error).  One thread per (transient, observation);
// the standard-normal variates are supplied by the caller (the reference draws them from numpy's Generator).
#include "rb_common.cuh"

namespace rb {

// Philox4x32-10 (Salmon et al. 2011, Random123): counter-based, so the variate of (transient, observation) depends only
// on the seed and the GLOBAL element index -- results do not change with chunking or with the number of ranks.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// standard normal from one Philox block: two 53-bit uniforms in (0, 1), Box-Muller (cosine branch)
__device__ __forceinline__ double philox_normal(uint64_t seed, uint64_t index) {
  uint32_t r[4];
  philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u1 = ((double)(((uint64_t)(r[0] >> 5) << 26) | (uint64_t)(r[1] >> 6)) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)(((uint64_t)(r[2] >> 5) << 26) | (uint64_t)(r[3] >> 6)) + 0.5) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// z == nullptr: the variates are drawn here (Philox, element index = (sample_base + n) * E + e); z_out (optional)
// receives them so that a checker can replay the noise model on the host.
__global__ void optical_noise_kernel(int noise_type, int64_t total, int E, const double* __restrict__ model_mag,
                                     const double* __restrict__ ref_flux, const double* __restrict__ limiting_mag,
                                     const double* __restrict__ z, uint64_t seed, int64_t sample_base,
                                     double snr_threshold, int apply_cut,
                                     double* __restrict__ obs_mag, double* __restrict__ mag_err,
                                     double* __restrict__ obs_flux, double* __restrict__ flux_err,
                                     double* __restrict__ snr_out, unsigned char* __restrict__ detected,
                                     double* __restrict__ z_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx % E);
  const double mm = model_mag[idx];
  const double zz = z ? z[idx] : philox_normal(seed, (uint64_t)(sample_base * E + idx));
  if (z_out) z_out[idx] = zz;
  double om, me, of = NAN, fe = NAN, snr;
  if (noise_type == RB_NOISE_LIMITING_MAG) {
    const double rf = ref_flux[e];
    const double model_flux = pow(10.0, mm / (-2.5)) * rf;                 // utils.py:716-718
    const double limiting_flux = pow(10.0, limiting_mag[e] / (-2.5)) * rf;
    fe = limiting_flux / 5.0;                                              // simulate_transients.py:472
    of = model_flux + zz * fe;                                             // :476  rng.normal(0, flux_error)
    snr = model_flux / fe;                                                 // :479
    om = -2.5 * log10(of / rf);                                            // utils.py:843-844
    me = fabs((2.5 / log(10.0)) * (fe / of));                              // utils.py:829
    if (isnan(of) || fabs(of) <= 1.0e-20) me = NAN;                        // utils.py:827-830
  } else if (noise_type == RB_NOISE_GAUSSIAN) {
    me = 0.1;                                                              // :492-494
    om = mm + zz * 0.1;
    snr = 1.0 / me;
  } else {  // RB_NOISE_GAUSSIANMODEL
    me = 0.05 * fabs(mm);                                                  // :505-507
    om = mm + zz * me;
    snr = 1.0 / 0.05;
  }
  obs_mag[idx] = om;
  mag_err[idx] = me;
  if (obs_flux) obs_flux[idx] = of;
  if (flux_err) flux_err[idx] = fe;
  if (snr_out) snr_out[idx] = snr;
  detected[idx] = apply_cut ? (unsigned char)(snr >= snr_threshold) : (unsigned char)1;   // :516-523
}

// Survey simulation with per-transient pointings: SimulateOpticalTransient._make_observation_single
// (simulate_transients.py:1096-1140) for a padded [N][E] table; row n has epoch_count[n] valid entries.
//   flux = bandpass_magnitude_to_flux(model_mag)                 utils.py:707-718
//   err  = ref_flux 10^(-0.4 depth) / 5  [+ source noise in quadrature]   utils.py:517-541, :1108-1109
//   observed_flux ~ N(flux, err)                                 :1111
//   magnitude = -2.5 log10(observed_flux / ref_flux)             utils.py:834-845
//   magnitude error from the MODEL flux                          utils.py:808-831, :1113
//   detected = not (phase <= 0 or isnan(magnitude)) and observed_flux / err >= snr_threshold   :1128-1134
__global__ void survey_noise_kernel(int64_t total, int E, const double* __restrict__ model_mag,
                                    const double* __restrict__ phase, const int* __restrict__ band_index,
                                    const int* __restrict__ epoch_count, const int64_t* __restrict__ row_start,
                                    const double* __restrict__ depth,
                                    const double* __restrict__ band_ref_flux, uint64_t seed, int64_t sample_base,
                                    double source_noise_factor, double snr_threshold, double* __restrict__ obs_mag,
                                    double* __restrict__ mag_err, double* __restrict__ obs_flux,
                                    double* __restrict__ flux_err, unsigned char* __restrict__ detected,
                                    double* __restrict__ z_out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t n = idx / E;
  const int e = (int)(idx - n * E);
  if (e >= epoch_count[n]) {                               // padding
    obs_mag[idx] = NAN; mag_err[idx] = NAN; obs_flux[idx] = NAN; flux_err[idx] = NAN; detected[idx] = 0;
    if (z_out) z_out[idx] = 0.0;
    return;
  }
  // variate of the pointing's position in the FLAT pointing table: independent of padding and chunking
  const double zz = philox_normal(seed, (uint64_t)(sample_base + row_start[n] + e));
  if (z_out) z_out[idx] = zz;
  const double rf = band_ref_flux[band_index[idx]];
  const double flux = pow(10.0, model_mag[idx] / (-2.5)) * rf;
  double err = (rf * pow(10.0, -0.4 * depth[idx])) / 5.0;
  if (source_noise_factor > 0.0) err = sqrt(err * err + (source_noise_factor * source_noise_factor) * (flux * flux));
  const double of = flux + zz * err;
  const double om = -2.5 * log10(of / rf);
  double me = fabs((2.5 / log(10.0)) * (err / flux));
  if (isnan(flux) || fabs(flux) <= 1.0e-20) me = NAN;
  const bool masked = (phase[idx] <= 0.0) || isnan(om);
  const bool low_snr = (of / err) < snr_threshold;
  obs_mag[idx] = om;
  mag_err[idx] = me;
  obs_flux[idx] = of;
  flux_err[idx] = err;
  detected[idx] = (unsigned char)(!masked && !low_snr);
}

}  // namespace rb
