// redback_b200 -- observation noise + SNR cut for simulated optical light curves, batched over transients.
// Reference: SimulateTransientWithCadence._evaluate_optical_model / _add_optical_noise_and_cuts
// (redback/simulate_transients.py:412-431, 457-527) with redback/utils.py:707-718 (magnitude -> flux),
// :834-845 (flux -> magnitude), :808-831 (magnitude error).  One thread per (transient, observation);
// the standard-normal variates are supplied by the caller (the reference draws them from numpy's Generator).
#include "rb_common.cuh"

namespace rb {

__global__ void optical_noise_kernel(int noise_type, int64_t total, int E, const double* __restrict__ model_mag,
                                     const double* __restrict__ ref_flux, const double* __restrict__ limiting_mag,
                                     const double* __restrict__ z, double snr_threshold, int apply_cut,
                                     double* __restrict__ obs_mag, double* __restrict__ mag_err,
                                     double* __restrict__ obs_flux, double* __restrict__ flux_err,
                                     double* __restrict__ snr_out, unsigned char* __restrict__ detected) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int e = (int)(idx % E);
  const double mm = model_mag[idx];
  const double zz = z[idx];
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
  snr_out[idx] = snr;
  detected[idx] = apply_cut ? (unsigned char)(snr >= snr_threshold) : (unsigned char)1;   // :516-523
}

}  // namespace rb
