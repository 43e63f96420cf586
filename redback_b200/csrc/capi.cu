// redback_b200 -- C ABI (include/redback_b200.h): argument checking, launches, host-buffer variants.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "sn_kernel.cu"
#include "kn_kernel.cu"
#include "csm_kernel.cu"
#include "mn_kernel.cu"
#include "direct_kernel.cu"
#include "viscous_kernel.cu"
#include "mk_kernel.cu"
#include "ag_kernel.cu"
#include "ext_kernel.cu"
#include "phot_kernel.cu"
#include "spectra_kernel.cu"
#include "noise_kernel.cu"
#include <algorithm>
#include <vector>

namespace {

thread_local std::string g_err;
std::atomic<int64_t>* launch_counter() {
  static std::atomic<int64_t> c{0};
  return &c;
}

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(RB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define RB_CUDA(call)                                        \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);    \
  } while (0)

// One-time set-up PER DEVICE: cudaFuncSetAttribute and the memory-pool threshold act on the current device only, so a
// process that evaluates on cuda:1 after cuda:0 (the public `device=` kwarg, or threads on different devices) must run
// them again there.  A failed set-up is retried on the next call.
struct DeviceOnce {
  std::mutex mu;
  bool done[64] = {};
  template <typename F> cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return f();
    std::lock_guard<std::mutex> g(mu);
    if (done[dev]) return cudaSuccess;
    e = f();
    if (e == cudaSuccess) done[dev] = true;
    return e;
  }
};

// Keep the stream-ordered workspace pool resident across synchronisations: by default the driver trims it to zero
// at every sync, and every call would pay for hundreds of MB of fresh allocations (measured: 2x on the photometry path).
void keep_workspace_pool() {
  static DeviceOnce once;
  once.run([] {
    int dev = 0;
    cudaMemPool_t pool;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetDefaultMemPool(&pool, dev);
    if (e == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return e;
  });
}

// stream-ordered device allocation released (cudaFreeAsync) when it goes out of scope, on every return path
struct StreamAlloc {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  StreamAlloc() = default;
  StreamAlloc(const StreamAlloc&) = delete;
  StreamAlloc& operator=(const StreamAlloc&) = delete;
  ~StreamAlloc() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t bytes, cudaStream_t s) {
    keep_workspace_pool();
    st = s;
    return cudaMallocAsync(&p, bytes ? bytes : 8, s);
  }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

size_t sn_smem_bytes(int D, int64_t E, bool epochs_in_smem, bool fp32 = false) {
  size_t n = (size_t)(2 * D + 2 * rb::kNodes + D + rb::kNodesHalf + rb::kNodes + rb::kNodes + rb::kExpTabSize + 2 * rb::kScalars + 16);  // incl. omx2
  if (epochs_in_smem) n += (size_t)rb::kEpochCols * (size_t)E;
  if (fp32) n += (size_t)(2 * rb::kNodes + (D + 1) / 2 * 2);   // FP32 copies: float4 nodes, float2 segments
  return sizeof(double) * n;
}

size_t csm_smem_bytes(int D, int64_t E, int quad_timesteps = 0) {
  size_t M4 = 4 * ((size_t)D + (size_t)E);
  // quadrature method: the knot arrays hold timesteps / 2 + timesteps abscissae instead
  if ((size_t)quad_timesteps * 3 / 2 + 4 > M4) M4 = (size_t)quad_timesteps * 3 / 2 + 4;
  return sizeof(double) * ((size_t)rb::kExpTabSize + 3 * (size_t)D + M4 + 3 * (size_t)E + 160 + 96) +
         sizeof(int) * (size_t)E;
}

// CSM engine (RB_ENGINE_CSM): unique epochs -> per-sample scalars -> one block per sample.  `t_dev` are the E
// observer-frame epochs (or the model's own time grid in grid mode) that a.epoch_tab was built from.
int launch_csm(rb::SnArgs a, const double* t_dev, cudaStream_t st, int sms) {
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t attr_err = cudaFuncSetAttribute(rb::csm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    return attr_err;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(csm_kernel)");
  const int E = a.E;
  const size_t smem = csm_smem_bytes(a.cfg.dense_resolution, E, a.cfg.csm_quadrature ? a.cfg.csm_timesteps : 0);
  if (smem > 220 * 1024)
    return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval (CSM): dense_resolution + number of epochs exceed shared memory");
  StreamAlloc b_ut, b_ui, b_sc;
  RB_CUDA(b_ut.alloc((size_t)E * sizeof(double), st));
  RB_CUDA(b_ui.alloc(((size_t)E + 1) * sizeof(int), st));
  RB_CUDA(b_sc.alloc((size_t)a.N * rb::kCsmScalars * sizeof(double), st));
  rb::CsmArgs c{};
  c.s = a;
  c.utime = b_ut.as<double>();
  c.uidx = b_ui.as<int>();
  c.n_unique = b_ui.as<int>() + E;
  c.cscal = b_sc.as<double>();
  rb::unique_times_kernel<<<1, 1024, (size_t)E * sizeof(int), st>>>(E, t_dev, b_ut.as<double>(), b_ui.as<int>(),
                                                                   b_ui.as<int>() + E);
  rb::csm_prep_kernel<<<(unsigned)((a.N + 7) / 8), 256, 0, st>>>(c);
  int occ = 0;
  RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rb::csm_kernel, rb::kCsmThreads, smem));
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)sms * occ * 4;
  if (grid > a.N) grid = a.N;
  rb::csm_kernel<<<(unsigned)grid, rb::kCsmThreads, smem, st>>>(c);
  launch_counter()->fetch_add(3);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

typedef void (*SnKernelFn)(const rb::SnArgs);
SnKernelFn sn_kernel_fn(int engine, bool has_t0, bool aspherical = false, bool negative = false, bool nocut = false,
                        int bolo = 0) {
  // bolometric output: 1 = blocks of at most 256 threads (higher register cap), 2 = larger blocks
  if (bolo && !aspherical && !negative && !has_t0 && (engine == RB_ENGINE_NICKEL || engine == RB_ENGINE_MAGNETAR))
    return bolo == 1 ? rb::sn_kernel<false, 0, false, false, false, 1> : rb::sn_kernel<false, 0, false, false, false, 2>;
  // the headline path (nickel / magnetar engine, bolometric or Blackbody SED, no t0): its own instantiation without
  // the CutoffBlackbody / Line code
  if (nocut && !aspherical && !negative && !has_t0 && (engine == RB_ENGINE_NICKEL || engine == RB_ENGINE_MAGNETAR))
    return rb::sn_kernel<false, 0, false, false, true>;
  if (aspherical) return rb::sn_kernel<false, 0, true>;
  if (negative)      // epochs before the dense grid: nickel / magnetar / magnetar_nickel without t0 (checked in sn_check)
    return engine == RB_ENGINE_MAGNETAR_NICKEL ? rb::sn_kernel<false, RB_ENGINE_MAGNETAR_NICKEL, false, true>
                                               : rb::sn_kernel<false, 0, false, true>;
  switch (engine) {
    case RB_ENGINE_MAGNETAR_NICKEL:
      return has_t0 ? rb::sn_kernel<true, RB_ENGINE_MAGNETAR_NICKEL> : rb::sn_kernel<false, RB_ENGINE_MAGNETAR_NICKEL>;
    case RB_ENGINE_MAGNETAR_ONLY:
      return has_t0 ? rb::sn_kernel<true, RB_ENGINE_MAGNETAR_ONLY> : rb::sn_kernel<false, RB_ENGINE_MAGNETAR_ONLY>;
    case RB_ENGINE_EXP_POWERLAW:
      return has_t0 ? rb::sn_kernel<true, RB_ENGINE_EXP_POWERLAW> : rb::sn_kernel<false, RB_ENGINE_EXP_POWERLAW>;
    case RB_ENGINE_TDE_FALLBACK:
      return has_t0 ? rb::sn_kernel<true, RB_ENGINE_TDE_FALLBACK> : rb::sn_kernel<false, RB_ENGINE_TDE_FALLBACK>;
    default:
      return has_t0 ? rb::sn_kernel<true, 0> : rb::sn_kernel<false, 0>;
  }
}

cudaError_t sn_kernel_attributes() {
  cudaError_t err = cudaSuccess;
  const int engines[5] = {RB_ENGINE_NICKEL, RB_ENGINE_MAGNETAR_NICKEL, RB_ENGINE_MAGNETAR_ONLY, RB_ENGINE_EXP_POWERLAW,
                          RB_ENGINE_TDE_FALLBACK};
  for (int e = 0; e < 5 && err == cudaSuccess; ++e)
    for (int t = 0; t < 2 && err == cudaSuccess; ++t)
      err = cudaFuncSetAttribute(sn_kernel_fn(engines[e], t != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 220 * 1024);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(sn_kernel_fn(0, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(sn_kernel_fn(RB_ENGINE_NICKEL, false, false, false, true),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(sn_kernel_fn(RB_ENGINE_NICKEL, false, false, false, false, 1),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(sn_kernel_fn(RB_ENGINE_NICKEL, false, false, false, false, 2),
                               cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int e = 0; e < 2 && err == cudaSuccess; ++e)
    err = cudaFuncSetAttribute(sn_kernel_fn(engines[e], false, false, true), cudaFuncAttributeMaxDynamicSharedMemorySize,
                               220 * 1024);
  return err;
}

int sn_block_threads(int64_t E) {
  int b = (int)((E + 31) / 32) * 32;
  if (b < 128) b = 128;
  if (b > 512) b = 512;
  return b;
}

}  // namespace

extern "C" {

const char* rb_last_error(void) { return g_err.c_str(); }
const char* rb_version(void) { return "redback_b200 0.2 (sm_100a)"; }
#ifndef RB_SOURCE_HASH
#define RB_SOURCE_HASH "unknown"
#endif
const char* rb_source_hash(void) { return "RBHASH:" RB_SOURCE_HASH; }
int64_t rb_launch_count(void) { return launch_counter()->load(); }

int rb_phot_mode_counts(uint64_t out[8], int reset) {
  unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (out) {
    RB_CUDA(cudaDeviceSynchronize());
    RB_CUDA(cudaMemcpyFromSymbol(h, rb::g_phot_modes, sizeof(h)));
    for (int i = 0; i < 8; ++i) out[i] = h[i];
  }
  if (reset) {
    const unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    RB_CUDA(cudaMemcpyToSymbol(rb::g_phot_modes, z, sizeof(z)));
  }
  return RB_OK;
}

int rb_cosmology_flat_lcdm(double H0, double Om0, double Tcmb0, double Neff, const double m_nu[3],
                           rb_cosmology* out) {
  if (!out || !(H0 > 0)) return fail(RB_ERR_INVALID, "rb_cosmology_flat_lcdm: bad arguments");
  const double mpc_cm = 3.0856775814913674e24, G_si = 6.6743e-11, kb_ev_k = 8.617333262145179e-05;
  const double H0_s = H0 * 1e5 / mpc_cm;
  const double rho_crit = 3.0 * H0_s * H0_s / (8.0 * M_PI * G_si) * 1e-3;
  const double a_B_c2 = 4.0 * (rb::kSigmaSB * 1e-3) / (rb::kCSI * rb::kCSI * rb::kCSI) * 1e-3;
  std::memset(out, 0, sizeof(*out));
  out->Om0 = Om0;
  out->Ogamma0 = a_B_c2 * std::pow(Tcmb0, 4) / rho_crit;
  const double Tnu0 = 0.7137658555036082 * Tcmb0;
  const int nnu = (int)std::floor(Neff);
  out->neff_per_nu = nnu > 0 ? Neff / nnu : 0.0;
  int nm = 0;
  for (int i = 0; i < 3 && i < nnu; ++i)
    if (m_nu && m_nu[i] > 0) out->nu_y[nm++] = m_nu[i] / (kb_ev_k * Tnu0);
  out->n_massive = nm;
  out->n_massless = nnu - nm;
  double rel = out->n_massless;
  for (int i = 0; i < nm; ++i) rel += std::pow(1.0 + std::pow(0.3173 * out->nu_y[i], 1.83), 0.54644808743);
  const double Onu0 = out->Ogamma0 * 0.22710731766 * out->neff_per_nu * rel;
  out->Ode0 = 1.0 - Om0 - out->Ogamma0 - Onu0;
  out->hubble_distance_cm = rb::kCSI / 1e3 / H0 * mpc_cm;
  return RB_OK;
}

int rb_cosmology_planck18(rb_cosmology* out) {
  const double m_nu[3] = {0.06, 0.0, 0.0};
  return rb_cosmology_flat_lcdm(67.66, 0.30966, 2.7255, 3.046, m_nu, out);
}

int rb_sn_config_default(rb_sn_config* cfg) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_sn_config_default: null config");
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->engine = RB_ENGINE_NICKEL;
  cfg->sed = RB_SED_BLACKBODY;
  cfg->sigma_mode = RB_SIGMA_FIXED;
  cfg->dense_resolution = 1000;
  cfg->cutoff_wavelength = 3000.0;
  cfg->absorption_index = 1.0;
  cfg->line_wavelength = 7.5e3;
  cfg->line_width = 500.0;
  cfg->line_time = 50.0;
  cfg->line_duration = 25.0;
  cfg->line_amplitude = 0.3;
  cfg->csm_nn = 12.0;
  cfg->csm_delta = 1.0;
  cfg->csm_efficiency = 0.5;
  cfg->csm_timesteps = 3000;
  return rb_cosmology_planck18(&cfg->cosmology);
}

// sn_viscous_kernel: one block per sample (grid-stride), dense grid + 1000 abscissae in shared memory
static cudaError_t launch_viscous(const rb::SnArgs& a, int64_t n, cudaStream_t st, int sms) {
  const size_t smem = sizeof(double) * rb::viscous_smem_doubles(a.cfg.dense_resolution);
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    return cudaFuncSetAttribute(rb::sn_viscous_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  });
  if (attr_err != cudaSuccess) return attr_err;
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  rb::sn_viscous_kernel<<<(unsigned)std::min<int64_t>(n, (int64_t)sms * 16), 256, smem, st>>>(a);
  return cudaGetLastError();
}

static int sn_check(const rb_sn_config* cfg, int64_t N, int64_t E, const void* params, const void* t_obs,
                    const void* freq_obs, const void* y, const void* sigma, const void* sigma_param,
                    const void* out_model, const void* out_logl) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_sn_eval: null config");
  if (N < 0 || E <= 0 || E > (1 << 24)) return fail(RB_ERR_INVALID, "rb_sn_eval: need N >= 0 and 0 < E <= 2^24");
  if (cfg->engine < RB_ENGINE_NICKEL || cfg->engine > RB_ENGINE_FALLBACK_NICKEL)
    return fail(RB_ERR_INVALID, "rb_sn_eval: unknown engine");
  if (cfg->aspherical && (cfg->engine > RB_ENGINE_MAGNETAR || cfg->t0 != nullptr || cfg->no_interaction))
    return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval: AsphericalDiffusion is available for the nickel / magnetar engines "
                                    "without t0");
  if (cfg->engine == RB_ENGINE_MAGNETAR_NICKEL && !cfg->f_nickel)
    return fail(RB_ERR_INVALID, "rb_sn_eval: cfg.f_nickel is required for RB_ENGINE_MAGNETAR_NICKEL");
  if (cfg->negative_epochs && !cfg->viscous && !cfg->no_interaction && cfg->engine < RB_ENGINE_FALLBACK &&
      (cfg->t0 != nullptr || cfg->aspherical ||
       !(cfg->engine == RB_ENGINE_NICKEL || cfg->engine == RB_ENGINE_MAGNETAR || cfg->engine == RB_ENGINE_MAGNETAR_NICKEL)))
    return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval: epochs before the dense grid are available for the nickel / magnetar "
                                    "engines with Diffusion (no t0, not aspherical) and for Viscous");
  if (cfg->viscous && (cfg->engine == RB_ENGINE_CSM || cfg->t0 != nullptr || cfg->no_interaction || cfg->aspherical ||
                       cfg->precision != RB_PREC_F64))
    return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval: Viscous is FP64, without t0, for every engine except the CSM shock");
  if (cfg->engine == RB_ENGINE_CSM) {
    // utils.py:304-310: RegularGridInterpolator raises outside the tabulated range of n
    if (!(cfg->csm_nn >= 6.0 && cfg->csm_nn <= 14.0))
      return fail(RB_ERR_INVALID, "One of the requested xi is out of bounds in dimension 0 (nn must be in [6, 14])");
    if (cfg->precision != RB_PREC_F64) return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval: the CSM path is FP64 only");
    if (cfg->dense_resolution < 20) return fail(RB_ERR_INVALID, "rb_sn_eval: dense_resolution must be >= 20 (CSM)");
    if (cfg->t0 != nullptr) return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval: t0 is not available for the CSM engine");
    if (cfg->csm_quadrature && (cfg->csm_timesteps < 4 || cfg->csm_timesteps > 6000 || (cfg->csm_timesteps & 1)))
      return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval: csm_diffusion_timesteps must be even and in [4, 6000] on device");
  }
  if (cfg->sed < RB_SED_BOLOMETRIC || cfg->sed > RB_SED_CUTOFF_LINE)
    return fail(RB_ERR_INVALID, "rb_sn_eval: unknown sed");
  if (cfg->sigma_mode < RB_SIGMA_FIXED || cfg->sigma_mode > RB_SIGMA_SYSTEMATIC)
    return fail(RB_ERR_INVALID, "rb_sn_eval: unknown sigma_mode");
  if (cfg->precision != RB_PREC_F64 && cfg->precision != RB_PREC_F32)
    return fail(RB_ERR_INVALID, "rb_sn_eval: unknown precision");
  if (cfg->precision == RB_PREC_F32 && (cfg->no_interaction || cfg->engine >= RB_ENGINE_FALLBACK))
    return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval: FP32 applies to the diffusion quadrature only");
  if (cfg->precision == RB_PREC_F32 && cfg->engine >= RB_ENGINE_EXP_POWERLAW)
    return fail(RB_ERR_UNSUPPORTED, "rb_sn_eval: the FP32 quadrature needs a monotonically decreasing engine");
  if (cfg->dense_resolution < 2 || cfg->dense_resolution > 8000)
    return fail(RB_ERR_INVALID, "rb_sn_eval: dense_resolution must be in [2, 8000]");
  if (!params || !t_obs) return fail(RB_ERR_INVALID, "rb_sn_eval: params and t_obs are required");
  if (cfg->sed != RB_SED_BOLOMETRIC && !freq_obs)
    return fail(RB_ERR_INVALID, "rb_sn_eval: freq_obs is required for flux_density output");
  if (cfg->sed == RB_SED_CUTOFF_BLACKBODY || cfg->sed == RB_SED_CUTOFF_LINE) {
    // sed.py:396-400: absorption_index >= 4 raises ValueError in the reference
    if (cfg->absorption_index >= 4.0)
      return fail(RB_ERR_INVALID, "absorption_index >= 4 is not supported: the integral below the cutoff diverges");
    if (!(cfg->cutoff_wavelength > 0)) return fail(RB_ERR_INVALID, "cutoff_wavelength must be positive");
  }
  if (y) {
    if (!out_logl) return fail(RB_ERR_INVALID, "rb_sn_eval: y given but out_logl is null");
    if (cfg->sigma_mode != RB_SIGMA_SAMPLED && !sigma)
      return fail(RB_ERR_INVALID, "rb_sn_eval: sigma is required for this sigma_mode");
    if (cfg->sigma_mode != RB_SIGMA_FIXED && cfg->sigma_mode != RB_SIGMA_FRACTIONAL && !sigma_param)
      return fail(RB_ERR_INVALID, "rb_sn_eval: sigma_param is required for this sigma_mode");
  } else if (out_logl) {
    return fail(RB_ERR_INVALID, "rb_sn_eval: out_logl requested without data y");
  }
  if (!out_model && !out_logl) return fail(RB_ERR_INVALID, "rb_sn_eval: nothing to compute");
  return RB_OK;
}

int rb_sn_eval_f64(const rb_sn_config* cfg, int64_t N, int64_t E, const double* params,
                   const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                   const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                   void* stream) {
  int rc = sn_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  RB_CUDA(cudaGetDevice(&dev));
  RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t attr_err = sn_kernel_attributes();
    return attr_err;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(sn_kernel)");

  // workspace: per-launch epoch table [kEpochCols][E] + per-sample scalar records [N][kScalars]
  const size_t ws_doubles = (size_t)rb::kEpochCols * (size_t)E + (size_t)rb::kScalars * (size_t)N;
  StreamAlloc ws_buf;
  RB_CUDA(ws_buf.alloc(ws_doubles * sizeof(double), st));
  double* ws = ws_buf.as<double>();
  double* epoch_tab = ws;
  rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, cfg->sigma_mode, t_obs, freq_obs, y,
                                                                    y ? sigma : nullptr, cfg->upper_limits, epoch_tab);
  rb::SnArgs a{};
  a.cfg = *cfg;
  a.N = N;
  a.E = (int)E;
  a.params = params;
  a.dl_cm = dl_cm;
  a.sigma_param = sigma_param;
  a.out_model = out_model;
  a.out_logl = y ? out_logl : nullptr;
  a.want_logl = y ? 1 : 0;
  a.epoch_tab = epoch_tab;
  a.scalars = ws + (size_t)rb::kEpochCols * (size_t)E;
  a.param_stride = N;
  if (cfg->engine == RB_ENGINE_CSM) {
    launch_counter()->fetch_add(1);
    return launch_csm(a, t_obs, st, sms);
  }
  if (cfg->viscous) {
    rb::sn_prep_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(a);
    RB_CUDA(launch_viscous(a, N, st, sms));
    launch_counter()->fetch_add(3);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
  }
  if (cfg->no_interaction || cfg->engine >= RB_ENGINE_FALLBACK) {
    rb::sn_prep_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(a);
    rb::sn_direct_kernel<<<(unsigned)std::min<int64_t>(N, (int64_t)sms * 32), 256, 0, st>>>(a);
    launch_counter()->fetch_add(3);
    RB_CUDA(cudaGetLastError());
    return RB_OK;
  }
  rb::sn_prep_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(a);
  const int threads = sn_block_threads(E);
  const bool fp32 = cfg->precision == RB_PREC_F32;
  a.epochs_in_smem = (sn_smem_bytes(cfg->dense_resolution, E, true, true) <= 64 * 1024) ? 1 : 0;
  size_t smem = sn_smem_bytes(cfg->dense_resolution, E, a.epochs_in_smem != 0, fp32);
  if (smem > 220 * 1024) return fail(RB_ERR_INVALID, "rb_sn_eval: dense_resolution too large for shared memory");
  int occ = 0;
  const bool nocut = cfg->sed == RB_SED_BLACKBODY;   // (bolometric, E = 100: measured 1.4 % slower in that instantiation)
  const SnKernelFn kern = sn_kernel_fn(cfg->engine, cfg->t0 != nullptr, cfg->aspherical != 0, cfg->negative_epochs != 0,
                                       nocut, (cfg->sed == RB_SED_BOLOMETRIC && !a.grid_mode) ? (threads <= 256 ? 1 : 2) : 0);
  RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (a.epochs_in_smem) {
    // the epoch table stays in global memory (L1) where that buys another resident block: E = 100 -> 6 blocks of 128
    // threads per SM instead of 5, arnett_bolometric 4.58e9 -> 4.82e9 unit/s; E = 300 is limited by registers either way
    const size_t smem_g = sn_smem_bytes(cfg->dense_resolution, E, false, fp32);
    int occ_g = 0;
    RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_g, kern, threads, smem_g));
    if (occ_g > occ) { a.epochs_in_smem = 0; smem = smem_g; occ = occ_g; }
  }
  if (occ < 1) occ = 1;
  // persistent blocks, grid-stride over samples; 4 waves' worth of blocks for load balance at mid-size N
  int64_t grid = (int64_t)sms * occ * 4;
  if (grid > N) grid = N;
  kern<<<(unsigned)grid, threads, smem, st>>>(a);
  launch_counter()->fetch_add(3);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

int rb_luminosity_distance_f64(const rb_cosmology* cosmo, int64_t N, const double* redshift,
                               double* out_dl_cm, void* stream) {
  if (!cosmo || !redshift || !out_dl_cm || N < 0) return fail(RB_ERR_INVALID, "rb_luminosity_distance: bad arguments");
  if (N == 0) return RB_OK;
  const int threads = 256, wpb = threads / 32;
  rb::dl_kernel<<<(unsigned)((N + wpb - 1) / wpb), threads, 0, (cudaStream_t)stream>>>(*cosmo, N, redshift, out_dl_cm);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

int rb_gaussian_logl_f64(int sigma_mode, int64_t N, int64_t E, const double* model, const double* y,
                         const double* sigma, const double* sigma_param, const rb_upper_limits* upper_limits,
                         double* out_logl, void* stream) {
  if (!model || !y || !out_logl || N < 0 || E <= 0) return fail(RB_ERR_INVALID, "rb_gaussian_logl: bad arguments");
  if (sigma_mode < RB_SIGMA_FIXED || sigma_mode > RB_SIGMA_SYSTEMATIC)
    return fail(RB_ERR_INVALID, "rb_gaussian_logl: unknown sigma_mode");
  if (sigma_mode != RB_SIGMA_SAMPLED && !sigma) return fail(RB_ERR_INVALID, "rb_gaussian_logl: sigma required");
  if (sigma_mode != RB_SIGMA_FIXED && sigma_mode != RB_SIGMA_FRACTIONAL && !sigma_param)
    return fail(RB_ERR_INVALID, "rb_gaussian_logl: sigma_param required");
  if (N == 0) return RB_OK;
  const int threads = E >= 256 ? 256 : (E >= 128 ? 128 : 64);
  const int64_t grid = N < 1048576 ? N : 1048576;
  rb::logl_kernel<<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(sigma_mode, N, (int)E, model, y,
                                                                         sigma ? sigma : y, sigma_param,
                                                                         upper_limits ? *upper_limits : rb_upper_limits{nullptr, 0}, out_logl);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

int rb_mixture_logl_f64(int64_t N, int64_t E, const double* model, const double* y, const double* sigma,
                        const double* alpha, const double* sigma_out, double* out_logl, void* stream) {
  if (!model || !y || !sigma || !alpha || !sigma_out || !out_logl || N < 0 || E <= 0)
    return fail(RB_ERR_INVALID, "rb_mixture_logl: bad arguments");
  if (N == 0) return RB_OK;
  const int threads = E >= 256 ? 256 : (E >= 128 ? 128 : 64);
  const int64_t grid = N < 1048576 ? N : 1048576;
  rb::mixture_logl_kernel<<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(N, (int)E, model, y, sigma, alpha,
                                                                                 sigma_out, out_logl);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

}  // extern "C"

#include "host_pipe.inc"

extern "C" {

int rb_sn_eval_f64_host(const rb_sn_config* cfg, int64_t N, int64_t E, const double* params,
                        const double* t_obs, const double* freq_obs, const double* dl_cm,
                        const double* y, const double* sigma, const double* sigma_param,
                        double* out_model, double* out_logl) {
  int rc = sn_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_sn_eval_f64, cfg, rows_block(RB_SN_NPARAM, params), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 524288);
}

int rb_sn_eval_f64_host_rows(const rb_sn_config* cfg, int64_t N, int64_t E, const double* const* rows,
                             const double* fill, const double* t_obs, const double* freq_obs,
                             const double* dl_cm, const double* y, const double* sigma,
                             const double* sigma_param, double* out_model, double* out_logl) {
  int rc = sn_check(cfg, N, E, rows, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_sn_eval_f64, cfg, rows_list(RB_SN_NPARAM, rows, fill), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 524288);
}

int rb_extinction_apply_f64(const rb_extinction* ext, int64_t N, int64_t E, const double* freq_obs,
                            const double* redshift, const double* av_host, double* model, void* stream) {
  if (!ext || !freq_obs || !redshift || !av_host || !model || N < 0 || E <= 0)
    return fail(RB_ERR_INVALID, "rb_extinction_apply: bad arguments");
  for (int law : {ext->host_law, ext->mw_law})
    if (law < RB_EXT_CCM89 || law > RB_EXT_FITZPATRICK99) return fail(RB_ERR_INVALID, "rb_extinction_apply: unknown law");
  if (N == 0) return RB_OK;
  const int64_t total = N * E;
  rb::ext_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*ext, N, (int)E, freq_obs,
                                                                                        redshift, av_host, model);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

int rb_synchrotron_add_f64(int64_t N, int64_t E, const double* freq_obs, const double* redshift, const double* pp,
                           const double* dl_cm, double nu_max, double f0, double source_radius, double* model,
                           void* stream) {
  if (!freq_obs || !redshift || !pp || !dl_cm || !model || N < 0 || E <= 0)
    return fail(RB_ERR_INVALID, "rb_synchrotron_add: bad arguments");
  if (N == 0) return RB_OK;
  const int64_t total = N * E;
  rb::synchrotron_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      N, (int)E, freq_obs, redshift, pp, dl_cm, nu_max, f0, source_radius, model);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

int64_t rb_sizeof(int which) {
  switch (which) {
    case 0: return (int64_t)sizeof(rb_cosmology);
    case 1: return (int64_t)sizeof(rb_upper_limits);
    case 2: return (int64_t)sizeof(rb_sn_config);
    case 3: return (int64_t)sizeof(rb_kn_config);
    case 4: return (int64_t)sizeof(rb_ag_config);
    case 5: return (int64_t)sizeof(rb_mn_config);
    case 6: return (int64_t)sizeof(rb_mk_config);
    case 7: return (int64_t)sizeof(rb_photometry);
    case 8: return (int64_t)sizeof(rb_extinction);
    default: return -1;
  }
}

int rb_measure_fp64_peak(int iters, double* out_tflops, double* out_ms) {
  if (iters <= 0 || !out_tflops) return fail(RB_ERR_INVALID, "rb_measure_fp64_peak: bad arguments");
  int dev = 0, sms = 0;
  RB_CUDA(cudaGetDevice(&dev));
  RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* d = nullptr;
  RB_CUDA(cudaMalloc(&d, sizeof(double)));
  cudaEvent_t e0, e1;
  RB_CUDA(cudaEventCreate(&e0));
  RB_CUDA(cudaEventCreate(&e1));
  const int threads = 512, blocks = sms * 4;
  rb::dfma_kernel<<<blocks, threads>>>(iters / 8 + 1, d);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    rb::dfma_kernel<<<blocks, threads>>>(iters, d);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  launch_counter()->fetch_add(6);
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "dfma_kernel");
  const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
  *out_tflops = flops / (best * 1e-3) / 1e12;
  if (out_ms) *out_ms = best;
  return RB_OK;
}

}  // extern "C"

// ---- kilonova path ------------------------------------------------------------------------------------
namespace {
size_t kn_smem_bytes(const rb_kn_config* cfg, int64_t E) {
  const size_t R = (size_t)cfg->dense_resolution;
  size_t bytes = sizeof(double) * (5 * R + rb::kExpTabSize + rb::kKnScalars + 34 + (size_t)rb::kEpochCols * (size_t)E);
  if (cfg->model == RB_KN_METZGER) bytes += sizeof(double) * 2 * rb::kShellWarps * R + sizeof(int) * rb::kShellWarps * R;
  return bytes;
}

__global__ void minmax_kernel(int E, const double* __restrict__ t, double* __restrict__ out) {
  // single block: min / max of the epochs (tiny E)
  __shared__ double smin[32], smax[32];
  double mn = INFINITY, mx = -INFINITY;
  for (int e = threadIdx.x; e < E; e += blockDim.x) { mn = fmin(mn, t[e]); mx = fmax(mx, t[e]); }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fmin(mn, smin[w]); mx = fmax(mx, smax[w]); }
    out[0] = mn;
    out[1] = mx;
  }
}
}  // namespace

extern "C" int rb_kn_config_default(rb_kn_config* cfg) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_kn_config_default: null config");
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->model = RB_KN_ONE_COMPONENT;
  cfg->sigma_mode = RB_SIGMA_FIXED;
  cfg->dense_resolution = 500;
  cfg->neutron_precursor_switch = 1;
  cfg->vmax = 0.7;
  cfg->grid_t_max = 7e6;
  return rb_cosmology_planck18(&cfg->cosmology);
}

static int kn_check(const rb_kn_config* cfg, int64_t N, int64_t E, const void* params, const void* t_obs,
                    const void* freq_obs, const void* y, const void* sigma, const void* sigma_param,
                    const void* out_model, const void* out_logl) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_kn_eval: null config");
  if (N < 0 || E <= 0) return fail(RB_ERR_INVALID, "rb_kn_eval: need N >= 0 and E > 0");
  if (cfg->model != RB_KN_ONE_COMPONENT && cfg->model != RB_KN_METZGER)
    return fail(RB_ERR_INVALID, "rb_kn_eval: unknown model");
  if (cfg->sigma_mode < RB_SIGMA_FIXED || cfg->sigma_mode > RB_SIGMA_SYSTEMATIC)
    return fail(RB_ERR_INVALID, "rb_kn_eval: unknown sigma_mode");
  if (cfg->dense_resolution < 20 || cfg->dense_resolution > 2000)
    return fail(RB_ERR_INVALID, "rb_kn_eval: dense_resolution must be in [20, 2000]");
  if (!params || !t_obs || !freq_obs) return fail(RB_ERR_INVALID, "rb_kn_eval: params, t_obs and freq_obs are required");
  if (cfg->model == RB_KN_METZGER && !(cfg->vmax > 0)) return fail(RB_ERR_INVALID, "rb_kn_eval: vmax must be positive");
  if (!(cfg->grid_t_max > 1.0 && cfg->grid_t_max <= 1e9))
    return fail(RB_ERR_INVALID, "rb_kn_eval: grid_t_max must be in (1, 1e9] seconds");
  if (y) {
    if (!out_logl) return fail(RB_ERR_INVALID, "rb_kn_eval: y given but out_logl is null");
    if (cfg->sigma_mode != RB_SIGMA_SAMPLED && !sigma) return fail(RB_ERR_INVALID, "rb_kn_eval: sigma is required");
    if (cfg->sigma_mode != RB_SIGMA_FIXED && cfg->sigma_mode != RB_SIGMA_FRACTIONAL && !sigma_param) return fail(RB_ERR_INVALID, "rb_kn_eval: sigma_param is required");
  } else if (out_logl) {
    return fail(RB_ERR_INVALID, "rb_kn_eval: out_logl requested without data y");
  }
  if (!out_model && !out_logl) return fail(RB_ERR_INVALID, "rb_kn_eval: nothing to compute");
  if (kn_smem_bytes(cfg, E) > 220 * 1024)
    return fail(RB_ERR_UNSUPPORTED, "rb_kn_eval: dense_resolution / number of epochs exceed shared memory");
  return RB_OK;
}

extern "C" int rb_kn_eval_f64(const rb_kn_config* cfg, int64_t N, int64_t E, const double* params,
                              const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                              const double* sigma, const double* sigma_param, double* out_model,
                              double* out_logl, void* stream) {
  int rc = kn_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  RB_CUDA(cudaGetDevice(&dev));
  RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t attr_err = cudaFuncSetAttribute(rb::kn_kernel<RB_KN_ONE_COMPONENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(rb::kn_kernel<RB_KN_METZGER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    return attr_err;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(kn_kernel)");
  const size_t ws_doubles = (size_t)rb::kEpochCols * (size_t)E + (size_t)rb::kKnScalars * (size_t)N + 2;
  StreamAlloc ws_buf;
  RB_CUDA(ws_buf.alloc(ws_doubles * sizeof(double), st));
  double* ws = ws_buf.as<double>();
  double* epoch_tab = ws;
  double* scal = ws + (size_t)rb::kEpochCols * (size_t)E;
  double* mm = scal + (size_t)rb::kKnScalars * (size_t)N;
  rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, cfg->sigma_mode, t_obs, freq_obs, y,
                                                                    y ? sigma : nullptr, cfg->upper_limits, epoch_tab);
  minmax_kernel<<<1, 256, 0, st>>>((int)E, t_obs, mm);
  double h_mm[2];
  cudaError_t ce = cudaMemcpyAsync(h_mm, mm, 2 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);  // two doubles; needed to size the time grid
  if (ce != cudaSuccess) return cuda_fail(ce, "epoch min/max");
  if (!cfg->t0 && !(h_mm[0] > 0.0)) return fail(RB_ERR_INVALID, "rb_kn_eval: epochs must be strictly positive (days)");
  rb::KnArgs a{};
  a.cfg = *cfg;
  a.N = N;
  a.E = (int)E;
  a.params = params;
  a.dl_cm = dl_cm;
  a.sigma_param = sigma_param;
  a.out_model = out_model;
  a.out_logl = y ? out_logl : nullptr;
  a.want_logl = y ? 1 : 0;
  a.epoch_tab = epoch_tab;
  a.scalars = scal;
  a.param_stride = N;
  a.tmin_obs = h_mm[0];
  a.tmax_obs = h_mm[1];
  a.t0 = cfg->t0;
  a.t_obs = t_obs;
  a.n_obs = (int)E;
  rb::kn_prep_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(a);
  const size_t smem = kn_smem_bytes(cfg, E);
  // measured (tools/bench_paths.py): one-component 512 threads 45.4 ms, 256: 35.1, 128: 29.8; Metzger (needs 224 shell
  // threads) 512: 404 ms, 256: 241 -- smaller blocks, more of them per SM hide the per-sample barriers
  const int threads = (cfg->model == RB_KN_METZGER) ? 256 : 128;
  int occ = 0;
  void (*kn_kern)(const rb::KnArgs) = (cfg->model == RB_KN_METZGER) ? rb::kn_kernel<RB_KN_METZGER> : rb::kn_kernel<RB_KN_ONE_COMPONENT>;
  RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kn_kern, threads, smem));
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)sms * occ * 4;
  if (grid > N) grid = N;
  kn_kern<<<(unsigned)grid, threads, smem, st>>>(a);
  launch_counter()->fetch_add(4);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

extern "C" int rb_kn_eval_f64_host(const rb_kn_config* cfg, int64_t N, int64_t E, const double* params,
                                   const double* t_obs, const double* freq_obs, const double* dl_cm,
                                   const double* y, const double* sigma, const double* sigma_param,
                                   double* out_model, double* out_logl) {
  int rc = kn_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_kn_eval_f64, cfg, rows_block(RB_KN_NPARAM, params), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 262144);
}

extern "C" int rb_kn_eval_f64_host_rows(const rb_kn_config* cfg, int64_t N, int64_t E, const double* const* rows,
                                        const double* fill, const double* t_obs, const double* freq_obs,
                                        const double* dl_cm, const double* y, const double* sigma,
                                        const double* sigma_param, double* out_model, double* out_logl) {
  int rc = kn_check(cfg, N, E, rows, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_kn_eval_f64, cfg, rows_list(RB_KN_NPARAM, rows, fill), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 262144);
}

// ---- magnetar-driven ejecta (mergernova) path ---------------------------------------------------------------
namespace {
// numpy.geomspace(a, b, n) for 0 < a < b (numpy/_core/function_base.py): 10 ** linspace(log10 a, log10 b, n), end
// points set exactly
void host_geomspace(double a, double b, int n, std::vector<double>& out) {
  const double la = std::log10(a), lb = std::log10(b);
  const double step = n > 1 ? (lb - la) / (double)(n - 1) : 0.0;
  for (int i = 0; i < n; ++i) {
    double v = std::pow(10.0, (double)i * step + la);
    if (i == 0) v = a;
    if (i == n - 1 && n > 1) v = b;
    out.push_back(v);
  }
}

int mn_check(const rb_mn_config* cfg, int64_t N, int64_t E, const void* params, const void* t_obs,
             const void* freq_obs, const void* y, const void* sigma, const void* sigma_param, const void* out_model,
             const void* out_logl) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_mn_eval: null config");
  if (N < 0 || E <= 0) return fail(RB_ERR_INVALID, "rb_mn_eval: need N >= 0 and E > 0");
  if (E > 65536) return fail(RB_ERR_UNSUPPORTED, "rb_mn_eval: at most 65536 epochs (the per-launch epoch sort is O(E^2))");
  if (cfg->engine < RB_MN_BASIC_MAGNETAR || cfg->engine > RB_MN_EVOLVING)
    return fail(RB_ERR_INVALID, "rb_mn_eval: unknown engine");
  if (cfg->sigma_mode < RB_SIGMA_FIXED || cfg->sigma_mode > RB_SIGMA_SYSTEMATIC)
    return fail(RB_ERR_INVALID, "rb_mn_eval: unknown sigma_mode");
  if (cfg->resolution < 8 || cfg->resolution > 4096)
    return fail(RB_ERR_INVALID, "rb_mn_eval: resolution must be in [8, 4096]");
  if (cfg->output < RB_MN_OUT_FLUX_DENSITY || cfg->output > RB_MN_OUT_LBOL)
    return fail(RB_ERR_INVALID, "rb_mn_eval: unknown output");
  if (!(cfg->grid_t_min > 0.0 && cfg->grid_t_max > cfg->grid_t_min))
    return fail(RB_ERR_INVALID, "rb_mn_eval: need 0 < grid_t_min < grid_t_max");
  if (cfg->t0 && cfg->output != RB_MN_OUT_FLUX_DENSITY)
    return fail(RB_ERR_UNSUPPORTED, "rb_mn_eval: cfg->t0 is available for RB_MN_OUT_FLUX_DENSITY only");
  if (cfg->output == RB_MN_OUT_SUPERNOVA && cfg->sed != RB_SED_BLACKBODY && cfg->sed != RB_SED_CUTOFF_BLACKBODY)
    return fail(RB_ERR_INVALID, "rb_mn_eval: sed must be RB_SED_BLACKBODY or RB_SED_CUTOFF_BLACKBODY");
  if (cfg->output == RB_MN_OUT_SUPERNOVA && cfg->sed == RB_SED_CUTOFF_BLACKBODY && cfg->absorption_index >= 4.0)
    return fail(RB_ERR_INVALID, "absorption_index >= 4 is not supported: the integral below the cutoff diverges");
  if ((cfg->output == RB_MN_OUT_TRAPPED_LUMINOSITY || cfg->output == RB_MN_OUT_TRAPPED_FLUX) &&
      (cfg->engine != RB_MN_MAGNETAR_ONLY || cfg->pair_cascade_switch || cfg->use_gamma_ray_opacity))
    return fail(RB_ERR_INVALID, "rb_mn_eval: the trapped-magnetar outputs need the magnetar_only engine without pair "
                                "cascade / gamma-ray opacity");
  if (!params || !t_obs || (!freq_obs && cfg->output != RB_MN_OUT_LBOL))
    return fail(RB_ERR_INVALID, "rb_mn_eval: params, t_obs and freq_obs are required");
  if (y) {
    if (!out_logl) return fail(RB_ERR_INVALID, "rb_mn_eval: y given but out_logl is null");
    if (cfg->sigma_mode != RB_SIGMA_SAMPLED && !sigma) return fail(RB_ERR_INVALID, "rb_mn_eval: sigma is required");
    if (cfg->sigma_mode != RB_SIGMA_FIXED && cfg->sigma_mode != RB_SIGMA_FRACTIONAL && !sigma_param)
      return fail(RB_ERR_INVALID, "rb_mn_eval: sigma_param is required");
  } else if (out_logl) {
    return fail(RB_ERR_INVALID, "rb_mn_eval: out_logl requested without data y");
  }
  if (!out_model && !out_logl) return fail(RB_ERR_INVALID, "rb_mn_eval: nothing to compute");
  return RB_OK;
}
}  // namespace

extern "C" int rb_optimal_time_array(double t_min, double t_max, int resolution, double* out, int* n_out) {
  if (!out || !n_out || resolution < 2 || !(t_min > 0.0) || !(t_max > t_min))
    return fail(RB_ERR_INVALID, "rb_optimal_time_array: need 0 < t_min < t_max, resolution >= 2");
  std::vector<double> v;
  const double log_range = std::log10(t_max) - std::log10(t_min);          // utils.py:80
  if (log_range > 6) {
    const int n_early = (int)(resolution * 0.25), n_mid = (int)(resolution * 0.50);
    const int n_late = resolution - n_early - n_mid;
    const double trans1 = t_min * std::pow(t_max / t_min, 0.25), trans2 = t_min * std::pow(t_max / t_min, 0.75);
    host_geomspace(t_min, trans1, n_early, v);
    host_geomspace(trans1, trans2, n_mid, v);
    host_geomspace(trans2, t_max, n_late, v);
    std::sort(v.begin(), v.end());                                         // np.unique
    v.erase(std::unique(v.begin(), v.end()), v.end());
  } else {
    host_geomspace(t_min, t_max, resolution, v);
  }
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
  *n_out = (int)v.size();
  return RB_OK;
}

extern "C" int rb_mn_config_default(rb_mn_config* cfg) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_mn_config_default: null config");
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->engine = RB_MN_BASIC_MAGNETAR;
  cfg->sigma_mode = RB_SIGMA_FIXED;
  cfg->ejecta_albedo = 0.5;
  cfg->pair_cascade_fraction = 0.05;
  cfg->resolution = 500;
  cfg->output = RB_MN_OUT_FLUX_DENSITY;
  cfg->photon_index = 2.0;
  cfg->use_r_process = 1;
  cfg->grid_t_min = 1e-4;
  cfg->grid_t_max = 1e8;
  cfg->sed = RB_SED_BLACKBODY;
  cfg->cutoff_wavelength = 3000.0;
  cfg->absorption_index = 1.0;
  return rb_cosmology_planck18(&cfg->cosmology);
}

extern "C" int rb_mn_eval_f64(const rb_mn_config* cfg, int64_t N, int64_t E, const double* params,
                              const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                              const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                              void* stream) {
  int rc = mn_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<double> tg((size_t)cfg->resolution);
  int G = 0;
  rc = rb_optimal_time_array(cfg->grid_t_min, cfg->grid_t_max, cfg->resolution, tg.data(), &G);   // :301, supernova_models.py:2476
  if (rc != RB_OK) return rc;
  StreamAlloc b_tg, b_ep, b_sorted, b_dl;
  RB_CUDA(b_tg.alloc((size_t)G * sizeof(double), st));
  RB_CUDA(cudaMemcpyAsync(b_tg.p, tg.data(), (size_t)G * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaStreamSynchronize(st));   // tg is a stack-owned staging buffer
  RB_CUDA(b_ep.alloc((size_t)rb::kEpochCols * (size_t)E * sizeof(double), st));
  RB_CUDA(b_sorted.alloc((size_t)rb::kMnEpochCols * (size_t)E * sizeof(double), st));
  rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, cfg->sigma_mode, t_obs, freq_obs, y,
                                                                    y ? sigma : nullptr, cfg->upper_limits,
                                                                    b_ep.as<double>());
  rb::mn_sort_epochs_kernel<<<1, 1024, 0, st>>>((int)E, b_ep.as<double>(), b_sorted.as<double>());
  launch_counter()->fetch_add(2);
  const double* dl = dl_cm;
  if (!dl) {
    RB_CUDA(b_dl.alloc((size_t)N * sizeof(double), st));
    rb::dl_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(cfg->cosmology, N, params + (size_t)RB_MN_REDSHIFT * N,
                                                          b_dl.as<double>());   // z = 0 -> d_L = 0 (unused: RB_MN_OUT_LBOL)
    launch_counter()->fetch_add(1);
    dl = b_dl.as<double>();
  }
  rb::MnArgs a{};
  a.cfg = *cfg;
  a.N = N;
  a.E = (int)E;
  a.G = G;
  a.params = params;
  a.param_stride = N;
  a.dl_cm = dl;
  a.sigma_param = sigma_param;
  a.tgrid = b_tg.as<double>();
  a.sorted_ep = b_sorted.as<double>();
  a.t0 = cfg->t0;
  a.out_model = out_model;
  a.out_logl = y ? out_logl : nullptr;
  a.want_logl = y ? 1 : 0;
  if (cfg->output >= RB_MN_OUT_SUPERNOVA || !cfg->use_r_process)
    rb::mn_kernel<true><<<(unsigned)((N + 127) / 128), 128, (size_t)G * sizeof(double), st>>>(a);
  else
    rb::mn_kernel<false><<<(unsigned)((N + 127) / 128), 128, (size_t)G * sizeof(double), st>>>(a);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

extern "C" int rb_mn_eval_f64_host(const rb_mn_config* cfg, int64_t N, int64_t E, const double* params,
                                   const double* t_obs, const double* freq_obs, const double* dl_cm,
                                   const double* y, const double* sigma, const double* sigma_param,
                                   double* out_model, double* out_logl) {
  int rc = mn_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_mn_eval_f64, cfg, rows_block(RB_MN_NPARAM, params), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 262144);
}

extern "C" int rb_mn_eval_f64_host_rows(const rb_mn_config* cfg, int64_t N, int64_t E, const double* const* rows,
                                        const double* fill, const double* t_obs, const double* freq_obs,
                                        const double* dl_cm, const double* y, const double* sigma,
                                        const double* sigma_param, double* out_model, double* out_logl) {
  int rc = mn_check(cfg, N, E, rows, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_mn_eval_f64, cfg, rows_list(RB_MN_NPARAM, rows, fill), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 262144);
}

// ---- magnetar-driven Metzger shell model ----------------------------------------------------------------------
namespace {
size_t mk_smem_bytes(int G) {
  return sizeof(double) * ((size_t)7 * G + 2 * (size_t)rb::kShellWarps * G + 2 * rb::kShells + 40) +
         sizeof(int) * (size_t)rb::kShellWarps * G;
}

int mk_check(const rb_mk_config* cfg, int64_t N, int64_t E, const void* params, const void* t_obs,
             const void* freq_obs, const void* y, const void* sigma, const void* sigma_param, const void* out_model,
             const void* out_logl) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_mk_eval: null config");
  if (N < 0 || E <= 0) return fail(RB_ERR_INVALID, "rb_mk_eval: need N >= 0 and E > 0");
  if (cfg->engine < RB_MN_BASIC_MAGNETAR || cfg->engine > RB_MN_EVOLVING)
    return fail(RB_ERR_INVALID, "rb_mk_eval: unknown engine");
  if (cfg->sigma_mode < RB_SIGMA_FIXED || cfg->sigma_mode > RB_SIGMA_SYSTEMATIC)
    return fail(RB_ERR_INVALID, "rb_mk_eval: unknown sigma_mode");
  if (cfg->resolution < 8 || cfg->resolution > 1024)
    return fail(RB_ERR_INVALID, "rb_mk_eval: resolution must be in [8, 1024]");
  if (!(cfg->vmax > 0)) return fail(RB_ERR_INVALID, "rb_mk_eval: vmax must be positive");
  if (!params || !t_obs || !freq_obs) return fail(RB_ERR_INVALID, "rb_mk_eval: params, t_obs and freq_obs are required");
  if (y) {
    if (!out_logl) return fail(RB_ERR_INVALID, "rb_mk_eval: y given but out_logl is null");
    if (cfg->sigma_mode != RB_SIGMA_SAMPLED && !sigma) return fail(RB_ERR_INVALID, "rb_mk_eval: sigma is required");
    if (cfg->sigma_mode != RB_SIGMA_FIXED && cfg->sigma_mode != RB_SIGMA_FRACTIONAL && !sigma_param)
      return fail(RB_ERR_INVALID, "rb_mk_eval: sigma_param is required");
  } else if (out_logl) {
    return fail(RB_ERR_INVALID, "rb_mk_eval: out_logl requested without data y");
  }
  if (!out_model && !out_logl) return fail(RB_ERR_INVALID, "rb_mk_eval: nothing to compute");
  return RB_OK;
}
}  // namespace

extern "C" int rb_mk_config_default(rb_mk_config* cfg) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_mk_config_default: null config");
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->engine = RB_MN_BASIC_MAGNETAR;
  cfg->sigma_mode = RB_SIGMA_FIXED;
  cfg->pair_cascade_switch = 1;
  cfg->ejecta_albedo = 0.5;
  cfg->pair_cascade_fraction = 0.01;
  cfg->neutron_precursor_switch = 1;
  cfg->vmax = 0.7;
  cfg->resolution = 300;
  return rb_cosmology_planck18(&cfg->cosmology);
}

extern "C" int rb_mk_eval_f64(const rb_mk_config* cfg, int64_t N, int64_t E, const double* params,
                              const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                              const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                              void* stream) {
  int rc = mk_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  RB_CUDA(cudaGetDevice(&dev));
  RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t attr_err = cudaFuncSetAttribute(rb::mk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    return attr_err;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(mk_kernel)");
  std::vector<double> tg((size_t)cfg->resolution);
  int G = 0;
  rc = rb_optimal_time_array(1e-4, 1e7, cfg->resolution, tg.data(), &G);   // magnetar_driven_ejecta_models.py:777
  if (rc != RB_OK) return rc;
  const size_t smem = mk_smem_bytes(G);
  if (smem > 220 * 1024) return fail(RB_ERR_UNSUPPORTED, "rb_mk_eval: resolution exceeds shared memory");
  StreamAlloc b_tg, b_ep, b_dl;
  RB_CUDA(b_tg.alloc((size_t)G * sizeof(double), st));
  RB_CUDA(cudaMemcpyAsync(b_tg.p, tg.data(), (size_t)G * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaStreamSynchronize(st));
  RB_CUDA(b_ep.alloc((size_t)rb::kEpochCols * (size_t)E * sizeof(double), st));
  rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, cfg->sigma_mode, t_obs, freq_obs, y,
                                                                    y ? sigma : nullptr, cfg->upper_limits,
                                                                    b_ep.as<double>());
  launch_counter()->fetch_add(1);
  const double* dl = dl_cm;
  if (!dl) {
    RB_CUDA(b_dl.alloc((size_t)N * sizeof(double), st));
    rb::dl_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(cfg->cosmology, N, params + (size_t)RB_MK_REDSHIFT * N,
                                                          b_dl.as<double>());
    launch_counter()->fetch_add(1);
    dl = b_dl.as<double>();
  }
  rb::MkArgs a{};
  a.cfg = *cfg;
  a.N = N;
  a.t0 = cfg->t0;
  a.E = (int)E;
  a.G = G;
  a.params = params;
  a.param_stride = N;
  a.dl_cm = dl;
  a.sigma_param = sigma_param;
  a.tgrid = b_tg.as<double>();
  a.epoch_tab = b_ep.as<double>();
  a.out_model = out_model;
  a.out_logl = y ? out_logl : nullptr;
  a.want_logl = y ? 1 : 0;
  int occ = 0;
  RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rb::mk_kernel, rb::kMkThreads, smem));
  if (occ < 1) occ = 1;
  rb::mk_kernel<<<(unsigned)std::min<int64_t>(N, (int64_t)sms * occ * 4), rb::kMkThreads, smem, st>>>(a);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

extern "C" int rb_mk_eval_f64_host(const rb_mk_config* cfg, int64_t N, int64_t E, const double* params,
                                   const double* t_obs, const double* freq_obs, const double* dl_cm,
                                   const double* y, const double* sigma, const double* sigma_param,
                                   double* out_model, double* out_logl) {
  int rc = mk_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_mk_eval_f64, cfg, rows_block(RB_MK_NPARAM, params), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 65536);
}

extern "C" int rb_mk_eval_f64_host_rows(const rb_mk_config* cfg, int64_t N, int64_t E, const double* const* rows,
                                        const double* fill, const double* t_obs, const double* freq_obs,
                                        const double* dl_cm, const double* y, const double* sigma,
                                        const double* sigma_param, double* out_model, double* out_logl) {
  int rc = mk_check(cfg, N, E, rows, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_mk_eval_f64, cfg, rows_list(RB_MK_NPARAM, rows, fill), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 65536);
}

// ---- afterglow path ---------------------------------------------------------------------------------------
extern "C" int rb_ag_config_default(rb_ag_config* cfg) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_ag_config_default: null config");
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->jet = RB_AG_TOPHAT;
  cfg->sigma_mode = RB_SIGMA_FIXED;
  cfg->steps = 250;
  cfg->res = 0;
  cfg->k = 0;
  cfg->expansion = 1;
  cfg->a1 = 1.0;
  cfg->ss = 0.01;
  cfg->aa = 0.5;
  return rb_cosmology_planck18(&cfg->cosmology);
}

static int ag_check(const rb_ag_config* cfg, int64_t N, int64_t E, const void* params, const void* t_obs,
                    const void* freq_obs, const void* y, const void* sigma, const void* sigma_param,
                    const void* out_model, const void* out_logl) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_ag_eval: null config");
  if (N < 0 || E <= 0) return fail(RB_ERR_INVALID, "rb_ag_eval: need N >= 0 and E > 0");
  if (cfg->jet < RB_AG_TOPHAT || cfg->jet > RB_AG_DOUBLE_GAUSSIAN) return fail(RB_ERR_INVALID, "rb_ag_eval: unknown jet");
  if (cfg->k != 0 && cfg->k != 2) return fail(RB_ERR_INVALID, "k must either be 0 or 2");  // base_models.py:574-575
  if (cfg->sigma_mode < RB_SIGMA_FIXED || cfg->sigma_mode > RB_SIGMA_SYSTEMATIC)
    return fail(RB_ERR_INVALID, "rb_ag_eval: unknown sigma_mode");
  if (cfg->steps < 2 || cfg->steps > 1000) return fail(RB_ERR_INVALID, "rb_ag_eval: steps must be in [2, 1000]");
  if (cfg->res < 0 || cfg->res > 400) return fail(RB_ERR_INVALID, "rb_ag_eval: res must be in [1, 400] (0 = default rule)");
  if (!params || !t_obs || !freq_obs) return fail(RB_ERR_INVALID, "rb_ag_eval: params, t_obs and freq_obs are required");
  if (y) {
    if (!out_logl) return fail(RB_ERR_INVALID, "rb_ag_eval: y given but out_logl is null");
    if (cfg->sigma_mode != RB_SIGMA_SAMPLED && !sigma) return fail(RB_ERR_INVALID, "rb_ag_eval: sigma is required");
    if (cfg->sigma_mode != RB_SIGMA_FIXED && cfg->sigma_mode != RB_SIGMA_FRACTIONAL && !sigma_param) return fail(RB_ERR_INVALID, "rb_ag_eval: sigma_param is required");
  } else if (out_logl) {
    return fail(RB_ERR_INVALID, "rb_ag_eval: out_logl requested without data y");
  }
  if (!out_model && !out_logl) return fail(RB_ERR_INVALID, "rb_ag_eval: nothing to compute");
  return RB_OK;
}

extern "C" int rb_ag_eval_f64(const rb_ag_config* cfg, int64_t N, int64_t E, const double* params,
                              const double* t_obs, const double* freq_obs, const double* dl_cm, const double* y,
                              const double* sigma, const double* sigma_param, double* out_model,
                              double* out_logl, void* stream) {
  int rc = ag_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  RB_CUDA(cudaGetDevice(&dev));
  RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t attr_err = cudaFuncSetAttribute(rb::ag_cell_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(rb::ag_cell_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    return attr_err;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(ag_cell_kernel)");

  // unique frequencies (np.unique(self.freq), base_models.py:586) and the time order of the epochs: 2 E doubles to the
  // host once per call
  std::vector<double> h_freq((size_t)E), h_time((size_t)E);
  RB_CUDA(cudaMemcpyAsync(h_freq.data(), freq_obs, (size_t)E * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaMemcpyAsync(h_time.data(), t_obs, (size_t)E * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  std::vector<double> uniq(h_freq);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  if ((int)uniq.size() > rb::kAgMaxNu)
    return fail(RB_ERR_UNSUPPORTED, "rb_ag_eval: more than 16 distinct frequencies in one call");
  // ag_cell_kernel's epoch tables: epochs in ascending time (NaN last), a run of `epl` consecutive ones per lane,
  // stored transposed (entry k of lane l at k * 32 + l), padded with -1
  const int epl = (int)((E + 31) / 32);
  const size_t EP = (size_t)32 * epl;
  std::vector<int64_t> order((size_t)E);
  for (int64_t e = 0; e < E; ++e) order[(size_t)e] = e;
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
    const double tx = h_time[(size_t)x], ty = h_time[(size_t)y];
    if (tx != tx) return false;
    if (ty != ty) return true;
    return tx < ty;
  });
  double t_epoch_min = INFINITY, t_epoch_max = -INFINITY;
  for (int64_t e = 0; e < E; ++e) {
    const double te = h_time[(size_t)e];
    if (te == te) { t_epoch_min = std::min(t_epoch_min, te); t_epoch_max = std::max(t_epoch_max, te); }
  }
  std::vector<double> h_tT(EP, 0.0);
  std::vector<int> h_nuT(EP, -1), h_origT(EP, -1);
  for (int64_t r = 0; r < E; ++r) {
    const int64_t e = order[(size_t)r];
    const size_t pos = (size_t)(r % epl) * 32 + (size_t)(r / epl);
    h_tT[pos] = h_time[(size_t)e];
    h_nuT[pos] = (int)(std::lower_bound(uniq.begin(), uniq.end(), h_freq[(size_t)e]) - uniq.begin());
    h_origT[pos] = (int)e;
  }

  const int steps = cfg->steps, n_nu = (int)uniq.size();
  const size_t cell_smem = sizeof(double) * ((size_t)rb::kAgWarps * (1 + n_nu) * steps + (size_t)rb::kAgWarps * EP +
                                             rb::kAgScalars + 4 * rb::kAgMaxNu);
  if (cell_smem > 220 * 1024) return fail(RB_ERR_UNSUPPORTED, "rb_ag_eval: steps / epochs exceed shared memory");

  const int64_t chunk_max = 8192;
  StreamAlloc b_epoch, b_tT, b_nuT, b_origT, b_scal, b_cnt, b_off, b_ucnt, b_uoff;
  RB_CUDA(b_epoch.alloc((size_t)rb::kEpochCols * (size_t)E * sizeof(double), st));
  RB_CUDA(b_tT.alloc(EP * sizeof(double), st));
  RB_CUDA(b_nuT.alloc(EP * sizeof(int), st));
  RB_CUDA(b_origT.alloc(EP * sizeof(int), st));
  double* d_epoch = b_epoch.as<double>();
  RB_CUDA(cudaMemcpyAsync(b_tT.as<double>(), h_tT.data(), EP * sizeof(double), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(b_nuT.as<int>(), h_nuT.data(), EP * sizeof(int), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaMemcpyAsync(b_origT.as<int>(), h_origT.data(), EP * sizeof(int), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaStreamSynchronize(st));  // the tables are pageable host memory
  rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, cfg->sigma_mode, t_obs, freq_obs, y,
                                                                    y ? sigma : nullptr, cfg->upper_limits, d_epoch);
  launch_counter()->fetch_add(1);

  const double q = rb::AG_QE * rb::AG_QE / (rb::AG_ME * rb::AG_C2);
  const double sigt = (q * q) * (8 * M_PI / 3);   // base_models.py:37
  RB_CUDA(b_scal.alloc((size_t)chunk_max * rb::kAgScalars * sizeof(double), st));
  RB_CUDA(b_cnt.alloc((size_t)chunk_max * sizeof(int), st));
  RB_CUDA(b_off.alloc((size_t)(chunk_max + 1) * sizeof(int64_t), st));
  RB_CUDA(b_ucnt.alloc((size_t)chunk_max * sizeof(int), st));
  RB_CUDA(b_uoff.alloc((size_t)(chunk_max + 1) * sizeof(int64_t), st));
  int* d_cnt = b_cnt.as<int>();
  int64_t* d_off = b_off.as<int64_t>();
  int* d_ucnt = b_ucnt.as<int>();
  int64_t* d_uoff = b_uoff.as<int64_t>();

  const size_t ring_bytes = (size_t)rb::kRingArrays * steps * sizeof(double);
  const size_t scratch_cap = (size_t)6 << 30;   // bound the ring scratch at 6 GiB per chunk
  int64_t n0 = 0;
  int64_t chunk = chunk_max;
  while (n0 < N) {
    const int64_t nc = std::min<int64_t>(chunk, N - n0);
    rb::AgArgs a{};
    a.cfg = *cfg;
    a.N = nc;
    a.n_base = n0;
    a.N_total = N;
    a.E = (int)E;
    a.params = params;
    a.dl_cm = dl_cm;
    a.sigma_param = sigma_param;
    a.out_model = out_model;
    a.out_logl = y ? out_logl : nullptr;
    a.want_logl = y ? 1 : 0;
    a.epoch_tab = d_epoch;
    a.ep_t_T = b_tT.as<double>();
    a.ep_nu_T = b_nuT.as<int>();
    a.ep_orig_T = b_origT.as<int>();
    a.epl = epl;
    a.t_epoch_min = t_epoch_min;
    a.t_epoch_max = t_epoch_max;
    for (int h = 0; h < rb::kAgMaxNu; ++h) a.nu_unique[h] = h < n_nu ? uniq[(size_t)h] : 0.0;
    a.n_nu = n_nu;
    a.scalars = b_scal.as<double>();
    a.ring_count = d_cnt;
    a.ring_off = d_off;
    a.unit_count = d_ucnt;
    a.unit_off = d_uoff;
    a.sigt = sigt;
    rb::ag_prep_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, st>>>(a);
    rb::ag_scan_kernel<<<1, 1024, 0, st>>>(nc, d_cnt, d_off);
    rb::ag_scan_kernel<<<1, 1024, 0, st>>>(nc, d_ucnt, d_uoff);
    int64_t total = 0, total_units = 0;
    RB_CUDA(cudaMemcpyAsync(&total, d_off + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaMemcpyAsync(&total_units, d_uoff + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    RB_CUDA(cudaStreamSynchronize(st));
    launch_counter()->fetch_add(3);
    if ((size_t)total * ring_bytes > scratch_cap && nc > 64) {
      chunk = std::max<int64_t>(64, nc / 2);   // too many rings (large res): retry with a smaller chunk
      continue;
    }
    StreamAlloc b_ring, b_part;
    RB_CUDA(b_ring.alloc(std::max<size_t>((size_t)total, 1) * ring_bytes, st));
    RB_CUDA(b_part.alloc(std::max<size_t>((size_t)total_units, 1) * (size_t)E * sizeof(double), st));
    a.ring_scratch = b_ring.as<double>();
    a.partial = b_part.as<double>();
    if (total > 0) {
      rb::ag_gamma_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, total);
      rb::ag_ring_kernel<<<(unsigned)((total + 3) / 4), 128, 0, st>>>(a, total);
      if (cfg->precision == RB_PREC_F32)
        rb::ag_cell_kernel<true><<<(unsigned)total_units, rb::kAgWarps * 32, cell_smem, st>>>(a, total_units);
      else
        rb::ag_cell_kernel<false><<<(unsigned)total_units, rb::kAgWarps * 32, cell_smem, st>>>(a, total_units);
      launch_counter()->fetch_add(3);
    }
    rb::ag_finish_kernel<<<(unsigned)std::min<int64_t>(nc, (int64_t)sms * 16), 128, 0, st>>>(a);
    launch_counter()->fetch_add(1);
    RB_CUDA(cudaGetLastError());
    n0 += nc;
  }
  return RB_OK;
}

extern "C" int rb_ag_eval_f64_host(const rb_ag_config* cfg, int64_t N, int64_t E, const double* params,
                                   const double* t_obs, const double* freq_obs, const double* dl_cm,
                                   const double* y, const double* sigma, const double* sigma_param,
                                   double* out_model, double* out_logl) {
  int rc = ag_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_ag_eval_f64, cfg, rows_block(RB_AG_NPARAM, params), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 32768);
}

extern "C" int rb_ag_eval_f64_host_rows(const rb_ag_config* cfg, int64_t N, int64_t E, const double* const* rows,
                                        const double* fill, const double* t_obs, const double* freq_obs,
                                        const double* dl_cm, const double* y, const double* sigma,
                                        const double* sigma_param, double* out_model, double* out_logl) {
  int rc = ag_check(cfg, N, E, rows, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  return eval_host_std(rb_ag_eval_f64, cfg, rows_list(RB_AG_NPARAM, rows, fill), N, E, t_obs, freq_obs, dl_cm, y, sigma,
                       sigma_param, out_model, out_logl, 32768);
}

// ---- magnitude / bandpass branch ---------------------------------------------------------------------------
namespace {

// FITPACK fpbspl (k = 3) on the host; same arithmetic as the device version
void host_fpbspl3(const double* t, int l, double x, double h[4]) {
  double hh[3];
  h[0] = 1.0;
  for (int j = 1; j <= 3; ++j) {
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
    for (int i = 1; i <= j; ++i) {
      const int li = l + i, lj = li - j;
      const double f = hh[i - 1] / (t[li] - t[lj]);
      h[i - 1] = h[i - 1] + f * (t[li] - x);
      h[i] = f * (x - t[lj]);
    }
  }
}

std::vector<double> nak_knots(const double* x, int n) {
  std::vector<double> t((size_t)n + 4);
  for (int i = 0; i < n + 4; ++i) t[(size_t)i] = (i < 4) ? x[0] : (i >= n ? x[n - 1] : x[i - 2]);
  return t;
}

// banded LU (l2, l1, d, u1, u2 per row) of the not-a-knot cubic B-spline collocation matrix on sites x[0..n)
std::vector<double> collocation_lu(const double* x, int n) {
  std::vector<double> t = nak_knots(x, n);
  std::vector<double> lu((size_t)5 * n, 0.0);
  for (int i = 0; i < n; ++i) {
    double* row = &lu[(size_t)5 * i];
    if (i == 0 || i == n - 1) { row[2] = 1.0; continue; }
    const int l = (i == 1) ? 3 : ((i >= n - 2) ? n - 1 : i + 2);
    double h[4];
    host_fpbspl3(t.data(), l, x[i], h);
    for (int q = 0; q < 4; ++q) {
      const int off = (l - 3 + q) - i + 2;
      if (off >= 0 && off < 5) row[off] = h[q];
    }
  }
  for (int i = 0; i < n; ++i) {
    double* r = &lu[(size_t)5 * i];
    double l2 = 0.0, l1 = 0.0, am1 = r[1], d = r[2], u1 = r[3];
    if (i >= 2) {
      const double* r2 = &lu[(size_t)5 * (i - 2)];
      l2 = r[0] / r2[2];
      am1 -= l2 * r2[3];
      d -= l2 * r2[4];
    }
    if (i >= 1) {
      const double* r1 = &lu[(size_t)5 * (i - 1)];
      l1 = am1 / r1[2];
      d -= l1 * r1[3];
      u1 -= l1 * r1[4];
    }
    r[0] = l2; r[1] = l1; r[2] = d; r[3] = u1;
  }
  return lu;
}

struct PhotDevice {
  // device copies of the per-call tables; freed stream-ordered by release()
  std::vector<void*> allocs;
  cudaStream_t st;
  template <typename T>
  cudaError_t upload(const std::vector<T>& h, const T** out) {
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, std::max<size_t>(h.size(), 1) * sizeof(T), st);
    if (e != cudaSuccess) return e;
    allocs.push_back(p);
    if (!h.empty()) e = cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    *out = (const T*)p;
    return e;
  }
  void release() {
    for (void* p : allocs) cudaFreeAsync(p, st);
    allocs.clear();
  }
};

int phot_check(const rb_photometry* ph, int64_t E, bool need_tgrid) {
  if (!ph) return fail(RB_ERR_INVALID, "rb_*_mag_eval: null photometry");
  if (ph->output != RB_OUT_MAGNITUDE && ph->output != RB_OUT_FLUX && ph->output != RB_OUT_SPECTRA)
    return fail(RB_ERR_INVALID, "Output format must be 'flux', 'magnitude', 'sncosmo_source', or 'flux_density'");
  if (ph->n_lambda < 8 || ph->n_lambda > 1024 || !ph->lambda_obs)
    return fail(RB_ERR_INVALID, "rb_*_mag_eval: lambda_array needs 8..1024 points");
  if (need_tgrid && (ph->n_tgrid < 8 || ph->n_tgrid > 2000 || !ph->tgrid))
    return fail(RB_ERR_INVALID, "rb_sn_mag_eval: time grid needs 8..2000 points");
  if (ph->output == RB_OUT_SPECTRA) {
    if (!ph->spectra_out || !ph->spectra_time_out || !ph->spectra_len_out)
      return fail(RB_ERR_INVALID, "rb_*_mag_eval: spectra output needs spectra_out, spectra_time_out and spectra_len_out");
    for (int i = 1; i < ph->n_lambda; ++i)
      if (!(ph->lambda_obs[i] > ph->lambda_obs[i - 1]))
        return fail(RB_ERR_INVALID, "rb_*_mag_eval: lambda_array must be strictly increasing");
    return RB_OK;
  }
  if (ph->n_bands < 1 || !ph->band_of_epoch || !ph->band_off || !ph->int_wave || !ph->int_wt || !ph->band_scale ||
      !ph->band_zp)
    return fail(RB_ERR_INVALID, "rb_*_mag_eval: incomplete band tables");
  if (ph->output == RB_OUT_FLUX && !ph->band_ref_flux)
    return fail(RB_ERR_INVALID, "rb_*_mag_eval: band_ref_flux is required for flux output");
  for (int64_t e = 0; e < E; ++e)
    if (ph->band_of_epoch[e] < 0 || ph->band_of_epoch[e] >= ph->n_bands)
      return fail(RB_ERR_INVALID, "rb_*_mag_eval: band_of_epoch out of range");
  for (int i = 1; i < ph->n_lambda; ++i)
    if (!(ph->lambda_obs[i] > ph->lambda_obs[i - 1]))
      return fail(RB_ERR_INVALID, "rb_*_mag_eval: lambda_array must be strictly increasing");
  return RB_OK;
}

// Build and upload everything phot_kernel needs that does not depend on the sample.
int phot_prepare(const rb_photometry* ph, int64_t E, PhotDevice& dev, rb::PhotArgs& pa) {
  const int NL = ph->n_lambda;
  std::vector<double> lam(ph->lambda_obs, ph->lambda_obs + NL);
  std::vector<double> lu_l = collocation_lu(lam.data(), NL);
  std::vector<double> tk = nak_knots(lam.data(), NL);
  const int64_t total = ph->band_off[ph->n_bands];
  std::vector<int> ly((size_t)total), boff((size_t)ph->n_bands + 1), jlo((size_t)ph->n_bands), span((size_t)ph->n_bands);
  std::vector<double> rec((size_t)total * 6);   // per integration point: 4 B-spline values, wave * trans, first coefficient
  for (int b = 0; b <= ph->n_bands; ++b) boff[(size_t)b] = (int)ph->band_off[b];
  for (int b = 0; b < ph->n_bands; ++b) {
    int lo = NL, hi = -1;
    if (ph->band_off[b + 1] <= ph->band_off[b]) return fail(RB_ERR_INVALID, "rb_*_mag_eval: empty band");
    for (int64_t w = ph->band_off[b]; w < ph->band_off[b + 1]; ++w) {
      double x = ph->int_wave[w];
      if (x < lam[0] || x > lam[(size_t)NL - 1])   // sed.py:22-28
        return fail(RB_ERR_INVALID, "bandpass outside spectral range of lambda_array");
      int l = 3;
      for (int lo2 = 4, hi2 = NL - 1; lo2 <= hi2;) {
        const int mid = (lo2 + hi2) >> 1;
        if (lam[(size_t)mid - 2] <= x) { l = mid; lo2 = mid + 1; } else hi2 = mid - 1;
      }
      double h[4];
      host_fpbspl3(tk.data(), l, x, h);
      ly[(size_t)w] = l - 3;
      for (int q = 0; q < 4; ++q) rec[(size_t)w * 6 + q] = h[q];
      rec[(size_t)w * 6 + 4] = ph->int_wt[w];
      lo = std::min(lo, l - 3);
      hi = std::max(hi, l);
    }
    jlo[(size_t)b] = lo;
    span[(size_t)b] = hi - lo + 1;
    for (int64_t w = ph->band_off[b]; w < ph->band_off[b + 1]; ++w) rec[(size_t)w * 6 + 5] = (double)(ly[(size_t)w] - lo);
    if (span[(size_t)b] > rb::kPhotMaxSpan)
      return fail(RB_ERR_UNSUPPORTED, "rb_*_mag_eval: a band overlaps too many wavelength-grid points");
  }
  // per-band weights of the spline coefficients for the all-positive case: sum_w wave trans B_j(wave)
  const int span_max = *std::max_element(span.begin(), span.end());
  std::vector<double> bw((size_t)ph->n_bands * (size_t)span_max, 0.0);
  for (int b = 0; b < ph->n_bands; ++b)
    for (int64_t w = ph->band_off[b]; w < ph->band_off[b + 1]; ++w) {
      const int rel = ly[(size_t)w] - jlo[(size_t)b];
      for (int q = 0; q < 4; ++q) bw[(size_t)b * span_max + rel + q] += rec[(size_t)w * 6 + 4] * rec[(size_t)w * 6 + q];
    }
  std::vector<int> boe(ph->band_of_epoch, ph->band_of_epoch + E);
  // epochs grouped by band: the warps of a block then integrate over the same band at the same time and its table of
  // integration points stays in L1
  std::vector<int> order((size_t)E);
  for (int64_t e = 0; e < E; ++e) order[(size_t)e] = (int)e;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return boe[(size_t)x] < boe[(size_t)y]; });
  std::vector<double> scale(ph->band_scale, ph->band_scale + ph->n_bands), zp(ph->band_zp, ph->band_zp + ph->n_bands);
  std::vector<double> ref;
  if (ph->band_ref_flux) ref.assign(ph->band_ref_flux, ph->band_ref_flux + ph->n_bands);
  cudaError_t e = dev.upload(lam, &pa.lambda_obs);
  if (e == cudaSuccess) e = dev.upload(lu_l, &pa.lu_lambda);
  if (e == cudaSuccess) e = dev.upload(boe, &pa.band_of_epoch);
  if (e == cudaSuccess) e = dev.upload(boff, &pa.band_off);
  if (e == cudaSuccess) e = dev.upload(jlo, &pa.band_jlo);
  if (e == cudaSuccess) e = dev.upload(span, &pa.band_span);
  if (e == cudaSuccess) e = dev.upload(scale, &pa.band_scale);
  if (e == cudaSuccess) e = dev.upload(zp, &pa.band_zp);
  if (e == cudaSuccess) e = dev.upload(ref, &pa.band_ref);
  if (e == cudaSuccess) e = dev.upload(order, &pa.epoch_order);
  if (e == cudaSuccess) e = dev.upload(rec, &pa.int_rec);
  if (e == cudaSuccess) e = dev.upload(bw, &pa.band_weight);
  if (e == cudaSuccess) e = cudaStreamSynchronize(dev.st);   // host vectors go out of scope
  if (e != cudaSuccess) return cuda_fail(e, "photometry table upload");
  pa.n_lambda = NL;
  pa.output = ph->output;
  pa.j_store_lo = *std::min_element(jlo.begin(), jlo.end());
  pa.span_max = span_max;
  // columns kept below the bluest band (phot_kernel.cu header): smallest M with 2.5 x^M / (1 - x) < 1e-11, x = 0.268 q^6
  {
    double q = 1.0;
    for (int j = 1; j <= pa.j_store_lo && j < NL; ++j) q = std::max(q, lam[(size_t)j] / lam[(size_t)j - 1]);
    const double x = 0.268 * std::pow(q, 6.0);
    // with dust extinction the 10-mag cap of _perform_extinction (extinction_models.py:137-139) can leave far-UV columns
    // up to 1e4 times LESS attenuated than the band: that factor comes off the target
    const double target = ph->extinction != nullptr ? 1e-15 : 1e-11;
    int margin = NL;                                   // very coarse grids: keep every column
    if (x < 0.8) margin = (int)std::ceil(std::log(target * (1.0 - x) / 2.5) / std::log(x));
    pa.blue_margin = std::max(8, margin);
    // red side (phot_kernel.cu, 0c): last coefficient any band uses and the grid's neighbour ratio beyond it
    int jhi = 0;
    for (int b = 0; b < ph->n_bands; ++b) jhi = std::max(jhi, jlo[(size_t)b] + span[(size_t)b] - 1);
    pa.j_store_hi = std::min(jhi, NL - 1);
    double qr = 1.0;
    for (int j = pa.j_store_hi + 1; j < NL; ++j) qr = std::max(qr, lam[(size_t)j] / lam[(size_t)j - 1]);
    pa.q_red = qr;
  }
  return RB_OK;
}

// Launch shape of phot_kernel: W wavelength columns per block and P row segments per column (W * P <= threads), chosen
// so that two blocks fit on an SM when at least 8 columns then fit, else one block per SM with as many columns as fit.
struct PhotShape {
  int W = 0, P = 0, nt_pad = 0, blocks_per_sm = 0;
  bool tables_global = false;
  size_t smem = 0;
};

// Launch shape of phot_kernel: P row segments per column (a power of two <= 32, the lanes that exchange boundary
// states by shuffles) and W = 256 / P wavelength columns per block -- the widest block that fits two blocks per SM,
// else one block per SM, else with the time-axis tables in global memory (very long grids).
bool phot_shape(int NT, int NL, int n_comp, bool line, PhotShape* out) {
  const int threads = 256;
  const int ntp = NT | 1;
  const size_t col = sizeof(double) * (size_t)ntp;
  const size_t budget2 = 115000, budget1 = 227 * 1024;
  int w_env = 0, bps_env = 0;
  if (const char* e = std::getenv("RB_PHOT_W")) w_env = std::atoi(e);                 // kernel experiments
  if (const char* e = std::getenv("RB_PHOT_BLOCKS_PER_SM")) bps_env = std::atoi(e);
  for (int pass = 0; pass < 3; ++pass) {
    const bool tg = pass == 2;
    const size_t budget = (pass == 0 && bps_env != 1) ? budget2 : budget1;
    if (pass == 0 && bps_env == 1) continue;
    const size_t fixed = sizeof(double) * rb::phot_fixed_doubles(NT, NL, n_comp, line, tg);
    if (fixed >= budget) continue;
    int w = 0;
    for (int P = 1; P <= 32; P <<= 1) {                       // widest first
      const int cand = threads / P;
      if (w_env > 0 && cand != w_env) continue;
      if (fixed + col * (size_t)cand <= budget) { w = cand; break; }
    }
    if (w == 0) continue;
    out->W = w;
    out->P = threads / w;
    out->nt_pad = ntp;
    out->blocks_per_sm = pass == 0 ? 2 : 1;
    out->tables_global = tg;
    out->smem = fixed + col * (size_t)w;
    return true;
  }
  return false;
}

// ---- shared driver of the magnitude / bandpass branch -------------------------------------------------------------
// The model kernels run in grid mode and leave (T, R, L, phase) per grid point in a chunk-sized scratch; phot_kernel
// turns them into spectra, spline coefficients, band integrals, magnitudes and the likelihood.  What differs between
// the model families is only how a chunk's grids are filled (`fill`) and a few SED / time-axis options.
struct PhotJob {
  const char* who;
  const rb_photometry* phot;
  int64_t N, E;
  int NT;                          // grid points per sample (allocated)
  const double* lu_grid;           // host [NT]: common time axis up to a per-sample scale -> one collocation LU; or
                                   // nullptr: every sample has its own grid (phot_kernel factorises per sample)
  bool with_len;                   // per-sample grid lengths
  int n_comp = 1;                  // blackbody components summed per sample; component c's grids start at
                                   // d_grid + c * (chunk capacity) * 4 * NT
  int sed = rb::PH_SED_BLACKBODY;
  double cutoff_wavelength = 3000.0, absorption_index = 1.0;
  double line_wavelength = 0, line_width = 0, line_time = 0, line_duration = 0, line_amplitude = 0;
  int sigma_mode;
  rb_upper_limits upper_limits;
  const double *t_obs, *y, *sigma, *sigma_param, *t0;
  double *out_model, *out_logl;
  // ragged mode (device pointers, see rb_sn_mag_eval_ragged_f64): per-sample pointings instead of the shared t_obs
  const double* t_ragged = nullptr;
  const int* band_ragged = nullptr;
  const int* epoch_count = nullptr;
};

// output_format='spectra': the same grid-mode fill, then spectra_kernel writes the N_t x N_lambda tile per sample
extern "C++" template <typename Fill>
int run_spectra(const PhotJob& j, cudaStream_t st, int sms, Fill fill) {
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    return cudaFuncSetAttribute(rb::spectra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(spectra_kernel)");
  const int NT = j.NT, NL = j.phot->n_lambda;
  const int n_comp = j.n_comp > 0 ? j.n_comp : 1;
  const bool line = j.sed == rb::PH_SED_CUTOFF_LINE;
  const size_t smem = ((size_t)(2 * n_comp + (line ? 2 : 0)) * NT + 4 * (size_t)NL) * sizeof(double);
  if (smem > 227 * 1024) return fail(RB_ERR_UNSUPPORTED, std::string(j.who) + ": grids exceed shared memory");
  const int64_t N = j.N, chunk_max = 32768, nc_max = std::min<int64_t>(chunk_max, N);
  double *d_lam = nullptr, *d_grid = nullptr, *d_opz = nullptr, *d_dl = nullptr;
  int* d_len = nullptr;
  rb_extinction* d_ext = nullptr;
  cudaError_t ce = cudaMallocAsync(&d_lam, (size_t)NL * sizeof(double), st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_lam, j.phot->lambda_obs, (size_t)NL * sizeof(double), cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaMallocAsync(&d_grid, (size_t)n_comp * nc_max * 4 * NT * sizeof(double), st);
  if (ce == cudaSuccess) ce = cudaMallocAsync(&d_opz, (size_t)nc_max * sizeof(double), st);
  if (ce == cudaSuccess) ce = cudaMallocAsync(&d_dl, (size_t)nc_max * sizeof(double), st);
  if (ce == cudaSuccess && j.with_len) ce = cudaMallocAsync(&d_len, (size_t)nc_max * sizeof(int), st);
  if (ce == cudaSuccess && j.phot->extinction != nullptr) {
    ce = cudaMallocAsync(&d_ext, sizeof(rb_extinction), st);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(d_ext, j.phot->extinction, sizeof(rb_extinction), cudaMemcpyHostToDevice, st);
  }
  int status = ce == cudaSuccess ? RB_OK : cuda_fail(ce, "spectra workspace");
  for (int64_t n0 = 0; n0 < N && status == RB_OK; n0 += chunk_max) {
    const int64_t nc = std::min<int64_t>(chunk_max, N - n0);
    status = fill(n0, nc, d_grid, d_opz, d_dl, d_len);
    if (status != RB_OK) break;
    rb::SpectraArgs sa{};
    sa.N = nc;
    sa.n_t_max = NT; sa.n_lambda = NL; sa.sed = j.sed; sa.n_comp = n_comp;
    sa.grid_comp_stride = (int64_t)nc_max * 4 * NT;
    sa.cutoff_wavelength = j.cutoff_wavelength; sa.absorption_index = j.absorption_index;
    sa.line_wavelength = j.line_wavelength; sa.line_width = j.line_width; sa.line_time = j.line_time;
    sa.line_duration = j.line_duration; sa.line_amplitude = j.line_amplitude;
    sa.grid = d_grid; sa.grid_len = d_len; sa.opz = d_opz; sa.dl = d_dl; sa.lambda_obs = d_lam;
    sa.ext = d_ext;
    sa.ext_av_host_s = d_ext ? j.phot->ext_av_host + n0 : nullptr;
    sa.spectra_out = j.phot->spectra_out + (size_t)n0 * NT * NL;
    sa.time_out = j.phot->spectra_time_out + (size_t)n0 * NT;
    sa.len_out = j.phot->spectra_len_out + n0;
    rb::spectra_kernel<<<(unsigned)std::min<int64_t>((int64_t)sms * 4, nc), 256, smem, st>>>(sa);
    launch_counter()->fetch_add(1);
    ce = cudaGetLastError();
    if (ce != cudaSuccess) status = cuda_fail(ce, "spectra kernel");
  }
  if (ce == cudaSuccess && status == RB_OK) ce = cudaStreamSynchronize(st);   // lambda_obs / extinction are host memory
  if (d_ext) cudaFreeAsync(d_ext, st);
  if (d_len) cudaFreeAsync(d_len, st);
  if (d_dl) cudaFreeAsync(d_dl, st);
  if (d_opz) cudaFreeAsync(d_opz, st);
  if (d_grid) cudaFreeAsync(d_grid, st);
  if (d_lam) cudaFreeAsync(d_lam, st);
  return status;
}

// fill(n0, nc, d_grid, d_opz, d_dl, d_len) -> RB_OK or an error code (already recorded with fail())
extern "C++" template <typename Fill>
int run_photometry(const PhotJob& j, cudaStream_t st, int sms, Fill fill) {
  keep_workspace_pool();
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t attr_err = cudaFuncSetAttribute(rb::phot_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(rb::phot_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(rb::phot_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return attr_err;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(phot_kernel)");
  const int NT = j.NT, NL = j.phot->n_lambda;
  const int64_t N = j.N, E = j.E;
  const int n_comp = j.n_comp > 0 ? j.n_comp : 1;
  if (j.phot->extinction != nullptr) {
    if (!j.phot->ext_av_host) return fail(RB_ERR_INVALID, std::string(j.who) + ": extinction needs ext_av_host");
    for (int law : {j.phot->extinction->host_law, j.phot->extinction->mw_law})
      if (law < RB_EXT_CCM89 || law > RB_EXT_FITZPATRICK99)
        return fail(RB_ERR_INVALID, std::string(j.who) + ": unknown extinction law");
  }
  if (j.phot->output == RB_OUT_SPECTRA) return run_spectra(j, st, sms, fill);
  PhotShape shape;
  if (!phot_shape(NT, NL, n_comp, j.sed == rb::PH_SED_CUTOFF_LINE, &shape))
    return fail(RB_ERR_UNSUPPORTED, std::string(j.who) + ": grids exceed shared memory");
  const size_t psmem = shape.smem;
  PhotDevice pd;
  pd.st = st;
  rb::PhotArgs pa{};
  const bool ragged = j.t_ragged != nullptr;
  int rc = phot_prepare(j.phot, ragged ? j.phot->n_bands : E, pd, pa);   // ragged: band_of_epoch is one entry per band
  if (rc != RB_OK) { pd.release(); return rc; }
  cudaError_t ce = cudaSuccess;
  if (j.lu_grid != nullptr) {
    std::vector<double> lu_t = collocation_lu(j.lu_grid, NT);
    ce = pd.upload(lu_t, &pa.lu_time);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { pd.release(); return cuda_fail(ce, "time grid upload"); }
  }
  double* d_ep_obs = nullptr;
  ce = cudaMallocAsync(&d_ep_obs, (size_t)rb::kEpochCols * E * sizeof(double), st);
  if (ce != cudaSuccess) { pd.release(); return cuda_fail(ce, "epoch table"); }
  if (!ragged) {
    rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, j.sigma_mode, j.t_obs, nullptr, j.y,
                                                                      j.y ? j.sigma : nullptr, j.upper_limits, d_ep_obs);
    launch_counter()->fetch_add(1);
  }

  const int64_t chunk_max = 32768;
  int occ_ph = 1;
  if (shape.tables_global)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ph, rb::phot_kernel<true>, 256, psmem);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ph, rb::phot_kernel<false>, 256, psmem);
  if (occ_ph < 1) occ_ph = 1;
  if (occ_ph > shape.blocks_per_sm) occ_ph = shape.blocks_per_sm;
  const int64_t ph_grid_max = (int64_t)sms * occ_ph;
  const int64_t nc_max = std::min<int64_t>(chunk_max, N);
  // per-block scratch: the epoch records + the forward-eliminated rows of the wavelength-axis solve (L2-resident)
  const int e_pad = (int)((E + 31) / 32 * 32);
  const int64_t scratch_stride = (int64_t)(rb::kEpCols + NL - pa.j_store_lo) * e_pad +
                                 (shape.tables_global ? 9 * (int64_t)NT + 2 : 0);
  double *d_grid = nullptr, *d_opz = nullptr, *d_dl = nullptr, *d_scratch = nullptr;
  int* d_len = nullptr;
  ce = cudaMallocAsync(&d_grid, (size_t)n_comp * nc_max * 4 * NT * sizeof(double), st);
  if (ce == cudaSuccess) ce = cudaMallocAsync(&d_opz, (size_t)nc_max * sizeof(double), st);
  if (ce == cudaSuccess) ce = cudaMallocAsync(&d_dl, (size_t)nc_max * sizeof(double), st);
  if (ce == cudaSuccess && j.with_len) ce = cudaMallocAsync(&d_len, (size_t)nc_max * sizeof(int), st);
  if (ce == cudaSuccess)
    ce = cudaMallocAsync(&d_scratch, (size_t)std::min<int64_t>(ph_grid_max, nc_max) * scratch_stride * sizeof(double), st);
  rb_extinction* d_ext = nullptr;
  if (ce == cudaSuccess && j.phot->extinction != nullptr) {
    ce = cudaMallocAsync(&d_ext, sizeof(rb_extinction), st);
    if (ce == cudaSuccess)   // pageable source: the copy is staged before the call returns
      ce = cudaMemcpyAsync(d_ext, j.phot->extinction, sizeof(rb_extinction), cudaMemcpyHostToDevice, st);
  }
  int status = RB_OK;
  if (ce != cudaSuccess) status = cuda_fail(ce, "magnitude workspace");
  // red-side cut (phot_kernel.cu, 0c): worth its instantiation only if at least 8 columns can go for the hottest samples
  bool red_cut = false;
  {
    const double xr0 = 0.268 * pa.q_red * pa.q_red;
    if (xr0 < 0.8) {
      const int m0 = std::max(8, (int)std::ceil(std::log(1e-11 * (1.0 - xr0) / 2.5) / std::log(xr0)));
      red_cut = NL - 1 - pa.j_store_hi >= m0 + 8;
    }
  }
  for (int64_t n0 = 0; n0 < N && status == RB_OK; n0 += chunk_max) {
    const int64_t nc = std::min<int64_t>(chunk_max, N - n0);
    status = fill(n0, nc, d_grid, d_opz, d_dl, d_len);
    if (status != RB_OK) break;
    pa.N = nc;
    pa.n_base = n0;
    pa.E = (int)E;
    pa.n_t_max = NT;
    pa.sed = j.sed;
    pa.cutoff_wavelength = j.cutoff_wavelength;
    pa.absorption_index = j.absorption_index;
    pa.line_wavelength = j.line_wavelength;
    pa.line_width = j.line_width;
    pa.line_time = j.line_time;
    pa.line_duration = j.line_duration;
    pa.line_amplitude = j.line_amplitude;
    pa.sigma_mode = j.sigma_mode;
    pa.want_logl = j.y ? 1 : 0;
    pa.common_time_lu = j.lu_grid != nullptr ? 1 : 0;
    pa.n_comp = n_comp;
    pa.grid_comp_stride = (int64_t)nc_max * 4 * NT;
    pa.grid = d_grid;
    pa.grid_len = d_len;
    pa.opz = d_opz;
    pa.dl = d_dl;
    pa.sigma_param_s = j.sigma_param ? j.sigma_param + n0 : nullptr;
    pa.t0_s = j.t0 ? j.t0 + n0 : nullptr;
    pa.ext = d_ext;
    pa.ext_av_host_s = d_ext ? j.phot->ext_av_host + n0 : nullptr;
    pa.epoch_tab = d_ep_obs;
    pa.t_ragged = j.t_ragged;
    pa.band_ragged = j.band_ragged;
    pa.epoch_count = j.epoch_count;
    pa.scratch = d_scratch;
    pa.scratch_stride = scratch_stride;
    pa.W = shape.W;
    pa.P = shape.P;
    pa.nt_pad = shape.nt_pad;
    pa.e_pad = e_pad;
    pa.out_model = j.out_model;
    pa.out_logl = j.y ? j.out_logl : nullptr;
    if (shape.tables_global)
      rb::phot_kernel<true><<<(unsigned)std::min<int64_t>(ph_grid_max, nc), 256, psmem, st>>>(pa);
    else if (red_cut)   // the red-side cut where the grid extends well beyond the reddest band
      rb::phot_kernel<false, true><<<(unsigned)std::min<int64_t>(ph_grid_max, nc), 256, psmem, st>>>(pa);
    else
      rb::phot_kernel<false><<<(unsigned)std::min<int64_t>(ph_grid_max, nc), 256, psmem, st>>>(pa);
    launch_counter()->fetch_add(1);
    ce = cudaGetLastError();
    if (ce != cudaSuccess) status = cuda_fail(ce, "magnitude kernels");
  }
  if (d_ext) cudaFreeAsync(d_ext, st);
  cudaFreeAsync(d_scratch, st);
  if (d_len) cudaFreeAsync(d_len, st);
  cudaFreeAsync(d_dl, st);
  cudaFreeAsync(d_opz, st);
  cudaFreeAsync(d_grid, st);
  cudaFreeAsync(d_ep_obs, st);
  pd.release();
  return status;
}

// upload a small host array for the lifetime of one call
struct CallBuffer {
  double* p = nullptr;
  cudaStream_t st = nullptr;
  ~CallBuffer() { if (p) cudaFreeAsync(p, st); }
  cudaError_t alloc(size_t n, cudaStream_t s) {
    keep_workspace_pool();
    st = s;
    return cudaMallocAsync(&p, (n ? n : 1) * sizeof(double), s);
  }
  cudaError_t upload(const std::vector<double>& h, cudaStream_t s) {
    keep_workspace_pool();
    st = s;
    cudaError_t e = cudaMallocAsync(&p, h.size() * sizeof(double), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);     // h may be a temporary
    return e;
  }
};

int device_and_sms(int* sms) {
  int dev = 0;
  RB_CUDA(cudaGetDevice(&dev));
  RB_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  return RB_OK;
}
}  // namespace

static int sn_mag_eval(const rb_sn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                       const double* params, const double* t_obs, const int* band_ragged, const int* epoch_count,
                       const double* dl_cm, const double* y, const double* sigma, const double* sigma_param,
                       double* out_model, double* out_logl, void* stream) {
  static const double dummy = 0.0;
  const bool ragged = band_ragged != nullptr;
  int rc = sn_check(cfg, N, E, params, t_obs, cfg && cfg->sed != RB_SED_BOLOMETRIC ? &dummy : nullptr, y, sigma,
                    sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (cfg->sed == RB_SED_BOLOMETRIC) return fail(RB_ERR_INVALID, "rb_sn_mag_eval: bolometric models have no magnitude output");
  if (ragged && (!epoch_count || cfg->t0 != nullptr))
    return fail(RB_ERR_INVALID, "rb_sn_mag_eval_ragged: epoch_count is required and cfg->t0 must be NULL (pass phases)");
  rc = phot_check(phot, ragged ? phot->n_bands : E, true);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0;
  rc = device_and_sms(&sms);
  if (rc != RB_OK) return rc;
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] { return sn_kernel_attributes(); });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(sn_kernel)");
  const int NT = phot->n_tgrid;
  std::vector<double> tg(phot->tgrid, phot->tgrid + NT);
  CallBuffer b_tgrid, b_ep_grid, b_scal;
  RB_CUDA(b_tgrid.upload(tg, st));
  // epoch table of the model grid (the "epochs" of the model kernel in grid mode)
  RB_CUDA(b_ep_grid.alloc((size_t)rb::kEpochCols * NT, st));
  rb::epoch_table_kernel<<<(unsigned)((NT + 255) / 256), 256, 0, st>>>(NT, RB_SIGMA_FIXED, b_tgrid.p, nullptr, nullptr,
                                                                     nullptr, rb_upper_limits{nullptr, 0}, b_ep_grid.p);
  launch_counter()->fetch_add(1);
  const int64_t nc_max = std::min<int64_t>(32768, N);
  RB_CUDA(b_scal.alloc((size_t)nc_max * rb::kScalars, st));
  const int sn_threads = sn_block_threads(NT);
  const int ep_in_smem = (sn_smem_bytes(cfg->dense_resolution, NT, true, true) <= 64 * 1024) ? 1 : 0;
  const size_t sn_smem = sn_smem_bytes(cfg->dense_resolution, NT, ep_in_smem != 0);
  const SnKernelFn sn_kern = sn_kernel_fn(cfg->engine, false, cfg->aspherical != 0);
  int occ_sn = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sn, sn_kern, sn_threads, sn_smem);
  if (occ_sn < 1) occ_sn = 1;

  PhotJob j{};
  j.who = "rb_sn_mag_eval";
  j.phot = phot; j.N = N; j.E = E; j.NT = NT;
  j.lu_grid = tg.data();
  j.with_len = false;
  j.sed = (cfg->sed == RB_SED_CUTOFF_BLACKBODY) ? rb::PH_SED_CUTOFF
          : (cfg->sed == RB_SED_CUTOFF_LINE ? rb::PH_SED_CUTOFF_LINE : rb::PH_SED_BLACKBODY);
  j.cutoff_wavelength = cfg->cutoff_wavelength;
  j.absorption_index = cfg->absorption_index;
  j.line_wavelength = cfg->line_wavelength; j.line_width = cfg->line_width; j.line_time = cfg->line_time;
  j.line_duration = cfg->line_duration; j.line_amplitude = cfg->line_amplitude;
  j.sigma_mode = cfg->sigma_mode; j.upper_limits = cfg->upper_limits;
  j.t_obs = t_obs; j.y = y; j.sigma = sigma; j.sigma_param = sigma_param;
  j.t0 = (cfg->t0 && cfg->engine != RB_ENGINE_CSM) ? cfg->t0 : nullptr;
  j.out_model = out_model; j.out_logl = out_logl;
  if (ragged) { j.t_ragged = t_obs; j.band_ragged = band_ragged; j.epoch_count = epoch_count; }
  return run_photometry(j, st, sms, [&](int64_t n0, int64_t nc, double* d_grid, double* d_opz, double* d_dl, int*) {
    rb::SnArgs a{};
    a.cfg = *cfg;
    a.N = nc;
    a.E = NT;
    a.params = params + n0;          // SoA rows are N_total apart: handled through the stride below
    if (cfg->f_nickel) a.cfg.f_nickel = cfg->f_nickel + n0;
    a.dl_cm = dl_cm ? dl_cm + n0 : nullptr;
    a.epoch_tab = b_ep_grid.p;
    a.scalars = b_scal.p;
    a.grid_mode = 1;
    a.grid_out = d_grid;
    a.opz_out = d_opz;
    a.dl_out = d_dl;
    a.param_stride = N;
    a.epochs_in_smem = ep_in_smem;
    if (cfg->engine == RB_ENGINE_CSM) return launch_csm(a, b_tgrid.p, st, sms);
    rb::sn_prep_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, st>>>(a);
    if (cfg->viscous) {
      if (launch_viscous(a, nc, st, sms) != cudaSuccess) return fail(RB_ERR_CUDA, "sn_viscous_kernel launch");
    } else if (cfg->no_interaction || cfg->engine >= RB_ENGINE_FALLBACK)
      rb::sn_direct_kernel<<<(unsigned)std::min<int64_t>(nc, (int64_t)sms * 32), 256, 0, st>>>(a);
    else
      sn_kern<<<(unsigned)std::min<int64_t>(nc, (int64_t)sms * occ_sn * 4), sn_threads, sn_smem, st>>>(a);
    launch_counter()->fetch_add(2);
    return (int)RB_OK;
  });
}

extern "C" int rb_sn_mag_eval_f64(const rb_sn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                                  const double* params, const double* t_obs, const double* dl_cm, const double* y,
                                  const double* sigma, const double* sigma_param, double* out_model,
                                  double* out_logl, void* stream) {
  return sn_mag_eval(cfg, phot, N, E, params, t_obs, nullptr, nullptr, dl_cm, y, sigma, sigma_param, out_model, out_logl,
                     stream);
}

extern "C" int rb_sn_mag_eval_ragged_f64(const rb_sn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E_max,
                                         const double* params, const double* phase, const int32_t* band_index,
                                         const int32_t* epoch_count, const double* dl_cm, double* out_model,
                                         void* stream) {
  if (!band_index) return fail(RB_ERR_INVALID, "rb_sn_mag_eval_ragged: band_index is required");
  return sn_mag_eval(cfg, phot, N, E_max, params, phase, band_index, epoch_count, dl_cm, nullptr, nullptr, nullptr,
                     out_model, nullptr, stream);
}

static int kn_mag_eval(const rb_kn_config* cfg, const rb_photometry* phot, int n_comp, int64_t N, int64_t E,
                       const double* params, const double* t_obs, const double* dl_cm, const double* y,
                       const double* sigma, const double* sigma_param, double* out_model, double* out_logl,
                       void* stream) {
  static const double dummy = 0.0;
  int rc = kn_check(cfg, N, E, params, t_obs, &dummy, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (n_comp < 1 || n_comp > 4) return fail(RB_ERR_INVALID, "rb_kn_multi_mag_eval: 1..4 components");
  if (n_comp > 1 && cfg->model != RB_KN_ONE_COMPONENT)
    return fail(RB_ERR_UNSUPPORTED, "rb_kn_multi_mag_eval: components are one-component kilonova models");
  rc = phot_check(phot, E, false);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0;
  rc = device_and_sms(&sms);
  if (rc != RB_OK) return rc;
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t attr_err = cudaFuncSetAttribute(rb::kn_kernel<RB_KN_ONE_COMPONENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(rb::kn_kernel<RB_KN_METZGER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    return attr_err;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(kn_kernel)");
  // epoch range for the adaptive grid (get_optimal_time_array(..., user_times)): one tiny D2H read
  CallBuffer b_mm, b_ep1, b_scal;
  RB_CUDA(b_mm.alloc(2, st));
  minmax_kernel<<<1, 256, 0, st>>>((int)E, t_obs, b_mm.p);
  double h_mm[2];
  RB_CUDA(cudaMemcpyAsync(h_mm, b_mm.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  launch_counter()->fetch_add(1);
  if (!cfg->t0 && !(h_mm[0] > 0.0)) return fail(RB_ERR_INVALID, "rb_kn_mag_eval: epochs must be strictly positive (days)");
  const int NT = cfg->dense_resolution;
  const int64_t nc_max = std::min<int64_t>(32768, N);   // = run_photometry's chunk capacity
  RB_CUDA(b_ep1.alloc((size_t)rb::kEpochCols, st));
  RB_CUDA(b_scal.alloc((size_t)nc_max * rb::kKnScalars, st));
  rb_kn_config kcfg = *cfg;
  const size_t ksmem = kn_smem_bytes(&kcfg, 1);
  const int kn_threads = (cfg->model == RB_KN_METZGER) ? 256 : 512;   // grid mode: 512 measured faster than 128 (457 vs 537 ms)
  int occ_kn = 1;
  void (*kn_kern)(const rb::KnArgs) = (cfg->model == RB_KN_METZGER) ? rb::kn_kernel<RB_KN_METZGER> : rb::kn_kernel<RB_KN_ONE_COMPONENT>;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_kn, kn_kern, kn_threads, ksmem);
  if (occ_kn < 1) occ_kn = 1;

  PhotJob j{};
  j.who = "rb_kn_mag_eval";
  j.phot = phot; j.N = N; j.E = E; j.NT = NT;
  j.lu_grid = nullptr;             // the grid depends on the sample's redshift
  j.with_len = true;
  j.n_comp = n_comp;
  j.sigma_mode = cfg->sigma_mode; j.upper_limits = cfg->upper_limits;
  j.t_obs = t_obs; j.y = y; j.sigma = sigma; j.sigma_param = sigma_param; j.t0 = cfg->t0;
  j.out_model = out_model; j.out_logl = out_logl;
  return run_photometry(j, st, sms, [&](int64_t n0, int64_t nc, double* d_grid, double* d_opz, double* d_dl, int* d_len) {
    for (int c = 0; c < n_comp; ++c) {   // the components share redshift, grid and epochs; only (T, R) differ
      rb::KnArgs a{};
      a.cfg = *cfg;
      a.N = nc;
      a.E = 1;                       // the kernel only stages the epoch table; unused in grid mode
      a.params = params + (size_t)c * RB_KN_NPARAM * N + n0;
      a.param_stride = N;
      a.dl_cm = dl_cm ? dl_cm + n0 : nullptr;
      a.epoch_tab = b_ep1.p;
      a.scalars = b_scal.p;
      a.tmin_obs = h_mm[0];
      a.tmax_obs = h_mm[1];
      a.t0 = cfg->t0 ? cfg->t0 + n0 : nullptr;
      a.t_obs = t_obs;
      a.n_obs = (int)E;
      a.grid_mode = 1;
      a.grid_out = d_grid + (size_t)c * nc_max * 4 * NT;
      a.grid_len = d_len;
      a.opz_out = d_opz;
      a.dl_out = d_dl;
      rb::kn_prep_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, st>>>(a);
      kn_kern<<<(unsigned)std::min<int64_t>(nc, (int64_t)sms * occ_kn * 4), kn_threads, ksmem, st>>>(a);
      launch_counter()->fetch_add(2);
    }
    return (int)RB_OK;
  });
}

extern "C" int rb_kn_mag_eval_f64(const rb_kn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                                  const double* params, const double* t_obs, const double* dl_cm, const double* y,
                                  const double* sigma, const double* sigma_param, double* out_model,
                                  double* out_logl, void* stream) {
  return kn_mag_eval(cfg, phot, 1, N, E, params, t_obs, dl_cm, y, sigma, sigma_param, out_model, out_logl, stream);
}

extern "C" int rb_kn_multi_mag_eval_f64(const rb_kn_config* cfg, const rb_photometry* phot, int n_components, int64_t N,
                                        int64_t E, const double* params, const double* t_obs, const double* dl_cm,
                                        const double* y, const double* sigma, const double* sigma_param,
                                        double* out_model, double* out_logl, void* stream) {
  return kn_mag_eval(cfg, phot, n_components, N, E, params, t_obs, dl_cm, y, sigma, sigma_param, out_model, out_logl,
                     stream);
}

extern "C" int rb_mn_mag_eval_f64(const rb_mn_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                                  const double* params, const double* t_obs, const double* dl_cm, const double* y,
                                  const double* sigma, const double* sigma_param, double* out_model,
                                  double* out_logl, void* stream) {
  static const double dummy = 0.0;
  int rc = mn_check(cfg, N, E, params, t_obs, &dummy, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  const bool supernova = cfg->output == RB_MN_OUT_SUPERNOVA && !cfg->use_r_process;   // supernova_models.py:2585-2637
  if (!supernova && (cfg->output != RB_MN_OUT_FLUX_DENSITY || !cfg->use_r_process))
    return fail(RB_ERR_UNSUPPORTED, "rb_mn_mag_eval: magnitudes are available for the mergernova flux-density set-up "
                                    "and for general_magnetar_driven_supernova");
  rc = phot_check(phot, E, false);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0;
  rc = device_and_sms(&sms);
  if (rc != RB_OK) return rc;
  std::vector<double> tg((size_t)cfg->resolution);
  int G = 0;
  rc = rb_optimal_time_array(cfg->grid_t_min, cfg->grid_t_max, cfg->resolution, tg.data(), &G);
  if (rc != RB_OK) return rc;
  tg.resize((size_t)G);
  std::vector<double> tdays((size_t)G);
  for (int i = 0; i < G; ++i) tdays[(size_t)i] = tg[(size_t)i] / 86400.0;     // time_eval (:236), up to the (1+z) scale
  CallBuffer b_tgrid, b_dlc;
  RB_CUDA(b_tgrid.upload(tg, st));
  if (!dl_cm) RB_CUDA(b_dlc.alloc((size_t)std::min<int64_t>(32768, N), st));

  PhotJob j{};
  j.who = "rb_mn_mag_eval";
  j.phot = phot; j.N = N; j.E = E; j.NT = G;
  j.lu_grid = tdays.data();
  j.with_len = false;
  if (supernova && cfg->sed == RB_SED_CUTOFF_BLACKBODY) {
    j.sed = rb::PH_SED_CUTOFF;
    j.cutoff_wavelength = cfg->cutoff_wavelength;
    j.absorption_index = cfg->absorption_index;
  }
  j.sigma_mode = cfg->sigma_mode; j.upper_limits = cfg->upper_limits;
  j.t_obs = t_obs; j.y = y; j.sigma = sigma; j.sigma_param = sigma_param; j.t0 = cfg->t0;
  j.out_model = out_model; j.out_logl = out_logl;
  return run_photometry(j, st, sms, [&](int64_t n0, int64_t nc, double* d_grid, double* d_opz, double* d_dl, int*) {
    const double* dl = dl_cm ? dl_cm + n0 : b_dlc.p;
    if (!dl_cm) {
      rb::dl_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, st>>>(cfg->cosmology, nc,
                                                            params + (size_t)RB_MN_REDSHIFT * N + n0, b_dlc.p);
      launch_counter()->fetch_add(1);
    }
    rb::MnArgs a{};
    a.cfg = *cfg;
    a.N = nc;
    a.E = 0;                       // no epochs are consumed in grid mode
    a.G = G;
    a.params = params + n0;
    a.param_stride = N;
    a.dl_cm = dl;
    a.tgrid = b_tgrid.p;
    a.grid_mode = 1;
    a.grid_out = d_grid;
    a.opz_out = d_opz;
    a.dl_out = d_dl;
    if (supernova)
      rb::mn_kernel<true><<<(unsigned)((nc + 127) / 128), 128, (size_t)G * sizeof(double), st>>>(a);
    else
      rb::mn_kernel<false><<<(unsigned)((nc + 127) / 128), 128, (size_t)G * sizeof(double), st>>>(a);
    launch_counter()->fetch_add(1);
    return (int)RB_OK;
  });
}

extern "C" int rb_mk_mag_eval_f64(const rb_mk_config* cfg, const rb_photometry* phot, int64_t N, int64_t E,
                                  const double* params, const double* t_obs, const double* dl_cm, const double* y,
                                  const double* sigma, const double* sigma_param, double* out_model,
                                  double* out_logl, void* stream) {
  static const double dummy = 0.0;
  int rc = mk_check(cfg, N, E, params, t_obs, &dummy, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  rc = phot_check(phot, E, false);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0;
  rc = device_and_sms(&sms);
  if (rc != RB_OK) return rc;
  static DeviceOnce once;
  const cudaError_t attr_err = once.run([] {
    cudaError_t attr_err = cudaFuncSetAttribute(rb::mk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    return attr_err;
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(mk_kernel)");
  std::vector<double> tg((size_t)cfg->resolution);
  int G = 0;
  rc = rb_optimal_time_array(1e-4, 1e7, cfg->resolution, tg.data(), &G);
  if (rc != RB_OK) return rc;
  tg.resize((size_t)G);
  std::vector<double> tdays((size_t)G);
  for (int i = 0; i < G; ++i) tdays[(size_t)i] = tg[(size_t)i] / 86400.0;
  const size_t mk_smem = mk_smem_bytes(G);
  if (mk_smem > 220 * 1024) return fail(RB_ERR_UNSUPPORTED, "rb_mk_mag_eval: resolution exceeds shared memory");
  CallBuffer b_tgrid, b_dlc;
  RB_CUDA(b_tgrid.upload(tg, st));
  if (!dl_cm) RB_CUDA(b_dlc.alloc((size_t)std::min<int64_t>(32768, N), st));

  PhotJob j{};
  j.who = "rb_mk_mag_eval";
  j.phot = phot; j.N = N; j.E = E; j.NT = G;
  j.lu_grid = tdays.data();
  j.with_len = false;
  j.sigma_mode = cfg->sigma_mode; j.upper_limits = cfg->upper_limits;
  j.t_obs = t_obs; j.y = y; j.sigma = sigma; j.sigma_param = sigma_param; j.t0 = cfg->t0;
  j.out_model = out_model; j.out_logl = out_logl;
  return run_photometry(j, st, sms, [&](int64_t n0, int64_t nc, double* d_grid, double* d_opz, double* d_dl, int*) {
    const double* dl = dl_cm ? dl_cm + n0 : b_dlc.p;
    if (!dl_cm) {
      rb::dl_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, st>>>(cfg->cosmology, nc,
                                                            params + (size_t)RB_MK_REDSHIFT * N + n0, b_dlc.p);
      launch_counter()->fetch_add(1);
    }
    rb::MkArgs a{};
    a.cfg = *cfg;
    a.N = nc;
    a.t0 = cfg->t0 ? cfg->t0 + n0 : nullptr;
    a.E = 0;
    a.G = G;
    a.params = params + n0;
    a.param_stride = N;
    a.dl_cm = dl;
    a.tgrid = b_tgrid.p;
    a.grid_mode = 1;
    a.grid_out = d_grid;
    a.opz_out = d_opz;
    a.dl_out = d_dl;
    rb::mk_kernel<<<(unsigned)std::min<int64_t>(nc, (int64_t)sms * 8), rb::kMkThreads, mk_smem, st>>>(a);
    launch_counter()->fetch_add(1);
    return (int)RB_OK;
  });
}

namespace {
int optical_noise(int noise_type, int64_t N, int64_t E, const double* model_mag, const double* ref_flux,
                  const double* limiting_mag, const double* z, uint64_t seed, int64_t sample_base, double snr_threshold,
                  int apply_snr_cut, double* obs_mag, double* mag_err, double* obs_flux, double* flux_err, double* snr,
                  unsigned char* detected, double* z_out, void* stream) {
  if (noise_type < RB_NOISE_LIMITING_MAG || noise_type > RB_NOISE_GAUSSIANMODEL)
    return fail(RB_ERR_INVALID, "rb_optical_noise: unknown noise_type");
  if (N < 0 || E <= 0 || sample_base < 0 || !model_mag || !obs_mag || !mag_err || !detected)
    return fail(RB_ERR_INVALID, "rb_optical_noise: bad arguments");
  if (noise_type == RB_NOISE_LIMITING_MAG && (!ref_flux || !limiting_mag))
    return fail(RB_ERR_INVALID, "rb_optical_noise: ref_flux and limiting_mag are required for limiting_mag noise");
  if (N == 0) return RB_OK;
  const int64_t total = N * E;
  rb::optical_noise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      noise_type, total, (int)E, model_mag, ref_flux, limiting_mag, z, seed, sample_base, snr_threshold, apply_snr_cut,
      obs_mag, mag_err, obs_flux, flux_err, snr, detected, z_out);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}
}  // namespace

extern "C" int rb_optical_noise_f64(int noise_type, int64_t N, int64_t E, const double* model_mag,
                                    const double* ref_flux, const double* limiting_mag, const double* z,
                                    double snr_threshold, int apply_snr_cut, double* obs_mag, double* mag_err,
                                    double* obs_flux, double* flux_err, double* snr, unsigned char* detected,
                                    void* stream) {
  if (!z || !snr) return fail(RB_ERR_INVALID, "rb_optical_noise: z and snr are required");
  return optical_noise(noise_type, N, E, model_mag, ref_flux, limiting_mag, z, 0, 0, snr_threshold, apply_snr_cut,
                       obs_mag, mag_err, obs_flux, flux_err, snr, detected, nullptr, stream);
}

extern "C" int rb_optical_noise_philox_f64(int noise_type, int64_t N, int64_t E, int64_t sample_base, uint64_t seed,
                                           const double* model_mag, const double* ref_flux,
                                           const double* limiting_mag, double snr_threshold, int apply_snr_cut,
                                           double* obs_mag, double* mag_err, double* obs_flux, double* flux_err,
                                           double* snr, unsigned char* detected, double* z_out, void* stream) {
  return optical_noise(noise_type, N, E, model_mag, ref_flux, limiting_mag, nullptr, seed, sample_base, snr_threshold,
                       apply_snr_cut, obs_mag, mag_err, obs_flux, flux_err, snr, detected, z_out, stream);
}

extern "C" int rb_survey_noise_philox_f64(int64_t N, int64_t E, int64_t sample_base, uint64_t seed,
                                          const double* model_mag, const double* phase, const int32_t* band_index,
                                          const int32_t* epoch_count, const int64_t* row_start,
                                          const double* five_sigma_depth,
                                          const double* band_ref_flux, double source_noise_factor, double snr_threshold,
                                          double* obs_mag, double* mag_err, double* obs_flux, double* flux_err,
                                          unsigned char* detected, double* z_out, void* stream) {
  if (N < 0 || E <= 0 || sample_base < 0 || !model_mag || !phase || !band_index || !epoch_count || !row_start || !five_sigma_depth ||
      !band_ref_flux || !obs_mag || !mag_err || !obs_flux || !flux_err || !detected)
    return fail(RB_ERR_INVALID, "rb_survey_noise: bad arguments");
  if (N == 0) return RB_OK;
  const int64_t total = N * E;
  rb::survey_noise_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      total, (int)E, model_mag, phase, band_index, epoch_count, row_start, five_sigma_depth, band_ref_flux, seed, sample_base,
      source_noise_factor, snr_threshold, obs_mag, mag_err, obs_flux, flux_err, detected, z_out);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}
