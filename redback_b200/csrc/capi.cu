// redback_b200 -- C ABI (include/redback_b200.h): argument checking, launches, host-buffer variants.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "sn_kernel.cu"
#include "kn_kernel.cu"
#include "ag_kernel.cu"
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

size_t sn_smem_bytes(int D, int64_t E, bool epochs_in_smem) {
  size_t n = (size_t)(2 * D + 2 * rb::kNodes + D + rb::kNodesHalf + rb::kNodes + 256 + 2 * rb::kScalars + 16);
  if (epochs_in_smem) n += (size_t)rb::kEpochCols * (size_t)E;
  return sizeof(double) * n;
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
const char* rb_version(void) { return "redback_b200 0.1 (sm_100a)"; }
int64_t rb_launch_count(void) { return launch_counter()->load(); }

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
  return rb_cosmology_planck18(&cfg->cosmology);
}

static int sn_check(const rb_sn_config* cfg, int64_t N, int64_t E, const void* params, const void* t_obs,
                    const void* freq_obs, const void* y, const void* sigma, const void* sigma_param,
                    const void* out_model, const void* out_logl) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_sn_eval: null config");
  if (N < 0 || E <= 0 || E > (1 << 24)) return fail(RB_ERR_INVALID, "rb_sn_eval: need N >= 0 and 0 < E <= 2^24");
  if (cfg->engine != RB_ENGINE_NICKEL && cfg->engine != RB_ENGINE_MAGNETAR)
    return fail(RB_ERR_INVALID, "rb_sn_eval: unknown engine");
  if (cfg->sed < RB_SED_BOLOMETRIC || cfg->sed > RB_SED_CUTOFF_BLACKBODY)
    return fail(RB_ERR_INVALID, "rb_sn_eval: unknown sed");
  if (cfg->sigma_mode < RB_SIGMA_FIXED || cfg->sigma_mode > RB_SIGMA_QUADRATURE)
    return fail(RB_ERR_INVALID, "rb_sn_eval: unknown sigma_mode");
  if (cfg->dense_resolution < 2 || cfg->dense_resolution > 8000)
    return fail(RB_ERR_INVALID, "rb_sn_eval: dense_resolution must be in [2, 8000]");
  if (!params || !t_obs) return fail(RB_ERR_INVALID, "rb_sn_eval: params and t_obs are required");
  if (cfg->sed != RB_SED_BOLOMETRIC && !freq_obs)
    return fail(RB_ERR_INVALID, "rb_sn_eval: freq_obs is required for flux_density output");
  if (cfg->sed == RB_SED_CUTOFF_BLACKBODY) {
    // sed.py:396-400: absorption_index >= 4 raises ValueError in the reference
    if (cfg->absorption_index >= 4.0)
      return fail(RB_ERR_INVALID, "absorption_index >= 4 is not supported: the integral below the cutoff diverges");
    const double s = 4.0 - cfg->absorption_index;
    if (s != std::floor(s) || s < 1.0)
      return fail(RB_ERR_UNSUPPORTED, "CutoffBlackbody on device needs integer 4 - absorption_index in {1,2,3,4}");
    if (!(cfg->cutoff_wavelength > 0)) return fail(RB_ERR_INVALID, "cutoff_wavelength must be positive");
  }
  if (y) {
    if (!out_logl) return fail(RB_ERR_INVALID, "rb_sn_eval: y given but out_logl is null");
    if (cfg->sigma_mode != RB_SIGMA_SAMPLED && !sigma)
      return fail(RB_ERR_INVALID, "rb_sn_eval: sigma is required for this sigma_mode");
    if (cfg->sigma_mode != RB_SIGMA_FIXED && !sigma_param)
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
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [dev] {
    attr_err = cudaFuncSetAttribute(rb::sn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    // keep the stream-ordered workspace pool resident across synchronisations (default trims it to 0)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(sn_kernel)");

  // workspace: per-launch epoch table [kEpochCols][E] + per-sample scalar records [N][kScalars]
  const size_t ws_doubles = (size_t)rb::kEpochCols * (size_t)E + (size_t)rb::kScalars * (size_t)N;
  double* ws = nullptr;
  RB_CUDA(cudaMallocAsync(&ws, ws_doubles * sizeof(double), st));
  double* epoch_tab = ws;
  rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, cfg->sigma_mode, t_obs, freq_obs, y,
                                                                    y ? sigma : nullptr, epoch_tab);
  rb::SnArgs a;
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
  rb::sn_prep_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(a);
  const int threads = sn_block_threads(E);
  a.epochs_in_smem = (sn_smem_bytes(cfg->dense_resolution, E, true) <= 64 * 1024) ? 1 : 0;
  const size_t smem = sn_smem_bytes(cfg->dense_resolution, E, a.epochs_in_smem != 0);
  if (smem > 220 * 1024) {
    cudaFreeAsync(ws, st);
    return fail(RB_ERR_INVALID, "rb_sn_eval: dense_resolution too large for shared memory");
  }
  int occ = 0;
  cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rb::sn_kernel, threads, smem);
  if (oe != cudaSuccess) {
    cudaFreeAsync(ws, st);
    return cuda_fail(oe, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  }
  if (occ < 1) occ = 1;
  // persistent blocks, grid-stride over samples; 4 waves' worth of blocks for load balance at mid-size N
  int64_t grid = (int64_t)sms * occ * 4;
  if (grid > N) grid = N;
  rb::sn_kernel<<<(unsigned)grid, threads, smem, st>>>(a);
  launch_counter()->fetch_add(3);
  cudaError_t le = cudaGetLastError();
  cudaFreeAsync(ws, st);
  if (le != cudaSuccess) return cuda_fail(le, "sn_kernel launch");
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
                         const double* sigma, const double* sigma_param, double* out_logl, void* stream) {
  if (!model || !y || !out_logl || N < 0 || E <= 0) return fail(RB_ERR_INVALID, "rb_gaussian_logl: bad arguments");
  if (sigma_mode < RB_SIGMA_FIXED || sigma_mode > RB_SIGMA_QUADRATURE)
    return fail(RB_ERR_INVALID, "rb_gaussian_logl: unknown sigma_mode");
  if (sigma_mode != RB_SIGMA_SAMPLED && !sigma) return fail(RB_ERR_INVALID, "rb_gaussian_logl: sigma required");
  if (sigma_mode != RB_SIGMA_FIXED && !sigma_param) return fail(RB_ERR_INVALID, "rb_gaussian_logl: sigma_param required");
  if (N == 0) return RB_OK;
  const int threads = E >= 256 ? 256 : (E >= 128 ? 128 : 64);
  const int64_t grid = N < 1048576 ? N : 1048576;
  rb::logl_kernel<<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(sigma_mode, N, (int)E, model, y,
                                                                         sigma ? sigma : y, sigma_param, out_logl);
  launch_counter()->fetch_add(1);
  RB_CUDA(cudaGetLastError());
  return RB_OK;
}

namespace {
struct DevBuf {
  double* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n * sizeof(double) : sizeof(double)); }
};
}  // namespace

int rb_sn_eval_f64_host(const rb_sn_config* cfg, int64_t N, int64_t E, const double* params,
                        const double* t_obs, const double* freq_obs, const double* dl_cm,
                        const double* y, const double* sigma, const double* sigma_param,
                        double* out_model, double* out_logl) {
  int rc = sn_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  DevBuf d_par, d_t, d_f, d_dl, d_y, d_s, d_sp, d_m, d_l;
  const size_t nN = (size_t)N, nE = (size_t)E;
  auto up = [&](DevBuf& b, const double* h, size_t n) -> cudaError_t {
    if (!h) return cudaSuccess;
    cudaError_t e = b.alloc(n);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(b.p, h, n * sizeof(double), cudaMemcpyHostToDevice);
  };
  RB_CUDA(up(d_par, params, (size_t)RB_SN_NPARAM * nN));
  RB_CUDA(up(d_t, t_obs, nE));
  RB_CUDA(up(d_f, freq_obs, nE));
  RB_CUDA(up(d_dl, dl_cm, nN));
  RB_CUDA(up(d_y, y, nE));
  RB_CUDA(up(d_s, sigma, nE));
  RB_CUDA(up(d_sp, sigma_param, nN));
  if (out_model) RB_CUDA(d_m.alloc(nN * nE));
  if (out_logl) RB_CUDA(d_l.alloc(nN));
  rc = rb_sn_eval_f64(cfg, N, E, d_par.p, d_t.p, d_f.p, d_dl.p, d_y.p, d_s.p, d_sp.p, d_m.p, d_l.p, nullptr);
  if (rc != RB_OK) return rc;
  if (out_model) RB_CUDA(cudaMemcpy(out_model, d_m.p, nN * nE * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_logl) RB_CUDA(cudaMemcpy(out_logl, d_l.p, nN * sizeof(double), cudaMemcpyDeviceToHost));
  RB_CUDA(cudaDeviceSynchronize());
  return RB_OK;
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
  size_t bytes = sizeof(double) * (5 * R + 256 + rb::kKnScalars + 34 + (size_t)rb::kEpochCols * (size_t)E);
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
  return rb_cosmology_planck18(&cfg->cosmology);
}

static int kn_check(const rb_kn_config* cfg, int64_t N, int64_t E, const void* params, const void* t_obs,
                    const void* freq_obs, const void* y, const void* sigma, const void* sigma_param,
                    const void* out_model, const void* out_logl) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_kn_eval: null config");
  if (N < 0 || E <= 0) return fail(RB_ERR_INVALID, "rb_kn_eval: need N >= 0 and E > 0");
  if (cfg->model != RB_KN_ONE_COMPONENT && cfg->model != RB_KN_METZGER)
    return fail(RB_ERR_INVALID, "rb_kn_eval: unknown model");
  if (cfg->sigma_mode < RB_SIGMA_FIXED || cfg->sigma_mode > RB_SIGMA_QUADRATURE)
    return fail(RB_ERR_INVALID, "rb_kn_eval: unknown sigma_mode");
  if (cfg->dense_resolution < 20 || cfg->dense_resolution > 2000)
    return fail(RB_ERR_INVALID, "rb_kn_eval: dense_resolution must be in [20, 2000]");
  if (!params || !t_obs || !freq_obs) return fail(RB_ERR_INVALID, "rb_kn_eval: params, t_obs and freq_obs are required");
  if (cfg->model == RB_KN_METZGER && !(cfg->vmax > 0)) return fail(RB_ERR_INVALID, "rb_kn_eval: vmax must be positive");
  if (y) {
    if (!out_logl) return fail(RB_ERR_INVALID, "rb_kn_eval: y given but out_logl is null");
    if (cfg->sigma_mode != RB_SIGMA_SAMPLED && !sigma) return fail(RB_ERR_INVALID, "rb_kn_eval: sigma is required");
    if (cfg->sigma_mode != RB_SIGMA_FIXED && !sigma_param) return fail(RB_ERR_INVALID, "rb_kn_eval: sigma_param is required");
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
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [dev] {
    attr_err = cudaFuncSetAttribute(rb::kn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(kn_kernel)");
  const size_t ws_doubles = (size_t)rb::kEpochCols * (size_t)E + (size_t)rb::kKnScalars * (size_t)N + 2;
  double* ws = nullptr;
  RB_CUDA(cudaMallocAsync(&ws, ws_doubles * sizeof(double), st));
  double* epoch_tab = ws;
  double* scal = ws + (size_t)rb::kEpochCols * (size_t)E;
  double* mm = scal + (size_t)rb::kKnScalars * (size_t)N;
  rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, cfg->sigma_mode, t_obs, freq_obs, y,
                                                                    y ? sigma : nullptr, epoch_tab);
  minmax_kernel<<<1, 256, 0, st>>>((int)E, t_obs, mm);
  double h_mm[2];
  cudaError_t ce = cudaMemcpyAsync(h_mm, mm, 2 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);  // two doubles; needed to size the time grid
  if (ce != cudaSuccess) {
    cudaFreeAsync(ws, st);
    return cuda_fail(ce, "epoch min/max");
  }
  if (!(h_mm[0] > 0.0)) {
    cudaFreeAsync(ws, st);
    return fail(RB_ERR_INVALID, "rb_kn_eval: epochs must be strictly positive (days)");
  }
  rb::KnArgs a;
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
  a.tmin_obs = h_mm[0];
  a.tmax_obs = h_mm[1];
  rb::kn_prep_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(a);
  const size_t smem = kn_smem_bytes(cfg, E);
  const int threads = 512;
  int occ = 0;
  cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rb::kn_kernel, threads, smem);
  if (oe != cudaSuccess) {
    cudaFreeAsync(ws, st);
    return cuda_fail(oe, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  }
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)sms * occ * 4;
  if (grid > N) grid = N;
  rb::kn_kernel<<<(unsigned)grid, threads, smem, st>>>(a);
  launch_counter()->fetch_add(4);
  cudaError_t le = cudaGetLastError();
  cudaFreeAsync(ws, st);
  if (le != cudaSuccess) return cuda_fail(le, "kn_kernel launch");
  return RB_OK;
}

extern "C" int rb_kn_eval_f64_host(const rb_kn_config* cfg, int64_t N, int64_t E, const double* params,
                                   const double* t_obs, const double* freq_obs, const double* dl_cm,
                                   const double* y, const double* sigma, const double* sigma_param,
                                   double* out_model, double* out_logl) {
  int rc = kn_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  DevBuf d_par, d_t, d_f, d_dl, d_y, d_s, d_sp, d_m, d_l;
  const size_t nN = (size_t)N, nE = (size_t)E;
  auto up = [&](DevBuf& b, const double* h, size_t n) -> cudaError_t {
    if (!h) return cudaSuccess;
    cudaError_t e = b.alloc(n);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(b.p, h, n * sizeof(double), cudaMemcpyHostToDevice);
  };
  RB_CUDA(up(d_par, params, (size_t)RB_KN_NPARAM * nN));
  RB_CUDA(up(d_t, t_obs, nE));
  RB_CUDA(up(d_f, freq_obs, nE));
  RB_CUDA(up(d_dl, dl_cm, nN));
  RB_CUDA(up(d_y, y, nE));
  RB_CUDA(up(d_s, sigma, nE));
  RB_CUDA(up(d_sp, sigma_param, nN));
  if (out_model) RB_CUDA(d_m.alloc(nN * nE));
  if (out_logl) RB_CUDA(d_l.alloc(nN));
  rc = rb_kn_eval_f64(cfg, N, E, d_par.p, d_t.p, d_f.p, d_dl.p, d_y.p, d_s.p, d_sp.p, d_m.p, d_l.p, nullptr);
  if (rc != RB_OK) return rc;
  if (out_model) RB_CUDA(cudaMemcpy(out_model, d_m.p, nN * nE * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_logl) RB_CUDA(cudaMemcpy(out_logl, d_l.p, nN * sizeof(double), cudaMemcpyDeviceToHost));
  RB_CUDA(cudaDeviceSynchronize());
  return RB_OK;
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
  return rb_cosmology_planck18(&cfg->cosmology);
}

static int ag_check(const rb_ag_config* cfg, int64_t N, int64_t E, const void* params, const void* t_obs,
                    const void* freq_obs, const void* y, const void* sigma, const void* sigma_param,
                    const void* out_model, const void* out_logl) {
  if (!cfg) return fail(RB_ERR_INVALID, "rb_ag_eval: null config");
  if (N < 0 || E <= 0) return fail(RB_ERR_INVALID, "rb_ag_eval: need N >= 0 and E > 0");
  if (cfg->jet != RB_AG_TOPHAT && cfg->jet != RB_AG_GAUSSIAN)
    return fail(RB_ERR_UNSUPPORTED, "rb_ag_eval: only the tophat ('TH') and Gaussian ('GJ') jets are on device");
  if (cfg->k != 0 && cfg->k != 2) return fail(RB_ERR_INVALID, "k must either be 0 or 2");  // base_models.py:574-575
  if (cfg->sigma_mode < RB_SIGMA_FIXED || cfg->sigma_mode > RB_SIGMA_QUADRATURE)
    return fail(RB_ERR_INVALID, "rb_ag_eval: unknown sigma_mode");
  if (cfg->steps < 2 || cfg->steps > 1000) return fail(RB_ERR_INVALID, "rb_ag_eval: steps must be in [2, 1000]");
  if (cfg->res < 0 || cfg->res > 400) return fail(RB_ERR_INVALID, "rb_ag_eval: res must be in [1, 400] (0 = default rule)");
  if (!params || !t_obs || !freq_obs) return fail(RB_ERR_INVALID, "rb_ag_eval: params, t_obs and freq_obs are required");
  if (y) {
    if (!out_logl) return fail(RB_ERR_INVALID, "rb_ag_eval: y given but out_logl is null");
    if (cfg->sigma_mode != RB_SIGMA_SAMPLED && !sigma) return fail(RB_ERR_INVALID, "rb_ag_eval: sigma is required");
    if (cfg->sigma_mode != RB_SIGMA_FIXED && !sigma_param) return fail(RB_ERR_INVALID, "rb_ag_eval: sigma_param is required");
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
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [dev] {
    attr_err = cudaFuncSetAttribute(rb::ag_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  });
  if (attr_err != cudaSuccess) return cuda_fail(attr_err, "cudaFuncSetAttribute(ag_cell_kernel)");

  // unique frequencies (np.unique(self.freq), base_models.py:586): E doubles to the host once per call
  std::vector<double> h_freq((size_t)E);
  RB_CUDA(cudaMemcpyAsync(h_freq.data(), freq_obs, (size_t)E * sizeof(double), cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  std::vector<double> uniq(h_freq);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  if ((int)uniq.size() > rb::kAgMaxNu)
    return fail(RB_ERR_UNSUPPORTED, "rb_ag_eval: more than 4 distinct frequencies in one call");
  std::vector<int> h_idx((size_t)E);
  for (int64_t e = 0; e < E; ++e)
    h_idx[(size_t)e] = (int)(std::lower_bound(uniq.begin(), uniq.end(), h_freq[(size_t)e]) - uniq.begin());

  const int steps = cfg->steps, n_nu = (int)uniq.size();
  const size_t cell_smem = sizeof(double) * ((size_t)rb::kRingArrays * steps + (size_t)E +
                                             (size_t)rb::kAgWarps * (1 + n_nu) * steps + (size_t)rb::kAgWarps * E +
                                             rb::kAgScalars) + sizeof(int) * (size_t)E;
  if (cell_smem > 220 * 1024) return fail(RB_ERR_UNSUPPORTED, "rb_ag_eval: steps / epochs exceed shared memory");

  const int64_t chunk_max = 8192;
  const size_t fixed_doubles = (size_t)rb::kEpochCols * (size_t)E;
  double* d_epoch = nullptr;
  int* d_idx = nullptr;
  RB_CUDA(cudaMallocAsync(&d_epoch, fixed_doubles * sizeof(double), st));
  RB_CUDA(cudaMallocAsync(&d_idx, (size_t)E * sizeof(int), st));
  RB_CUDA(cudaMemcpyAsync(d_idx, h_idx.data(), (size_t)E * sizeof(int), cudaMemcpyHostToDevice, st));
  RB_CUDA(cudaStreamSynchronize(st));  // h_idx is pageable host memory
  rb::epoch_table_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>((int)E, cfg->sigma_mode, t_obs, freq_obs, y,
                                                                    y ? sigma : nullptr, d_epoch);
  launch_counter()->fetch_add(1);

  const double q = rb::AG_QE * rb::AG_QE / (rb::AG_ME * rb::AG_C2);
  const double sigt = (q * q) * (8 * M_PI / 3);   // base_models.py:37
  double* d_scal = nullptr;
  int* d_cnt = nullptr;
  int64_t* d_off = nullptr;
  int* d_ucnt = nullptr;
  int64_t* d_uoff = nullptr;
  RB_CUDA(cudaMallocAsync(&d_scal, (size_t)chunk_max * rb::kAgScalars * sizeof(double), st));
  RB_CUDA(cudaMallocAsync(&d_cnt, (size_t)chunk_max * sizeof(int), st));
  RB_CUDA(cudaMallocAsync(&d_off, (size_t)(chunk_max + 1) * sizeof(int64_t), st));
  RB_CUDA(cudaMallocAsync(&d_ucnt, (size_t)chunk_max * sizeof(int), st));
  RB_CUDA(cudaMallocAsync(&d_uoff, (size_t)(chunk_max + 1) * sizeof(int64_t), st));

  int status = RB_OK;
  const size_t ring_bytes = (size_t)rb::kRingArrays * steps * sizeof(double);
  const size_t scratch_cap = (size_t)6 << 30;   // bound the ring scratch at 6 GiB per chunk
  int64_t n0 = 0;
  int64_t chunk = chunk_max;
  while (n0 < N && status == RB_OK) {
    const int64_t nc = std::min<int64_t>(chunk, N - n0);
    rb::AgArgs a;
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
    a.nu_index = d_idx;
    for (int h = 0; h < rb::kAgMaxNu; ++h) a.nu_unique[h] = h < n_nu ? uniq[(size_t)h] : 0.0;
    a.n_nu = n_nu;
    a.scalars = d_scal;
    a.ring_count = d_cnt;
    a.ring_off = d_off;
    a.ring_scratch = nullptr;
    a.unit_count = d_ucnt;
    a.unit_off = d_uoff;
    a.partial = nullptr;
    a.sigt = sigt;
    rb::ag_prep_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, st>>>(a);
    rb::ag_scan_kernel<<<1, 1024, 0, st>>>(nc, d_cnt, d_off);
    rb::ag_scan_kernel<<<1, 1024, 0, st>>>(nc, d_ucnt, d_uoff);
    int64_t total = 0, total_units = 0;
    cudaError_t ce = cudaMemcpyAsync(&total, d_off + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(&total_units, d_uoff + nc, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    launch_counter()->fetch_add(3);
    if (ce != cudaSuccess) { status = cuda_fail(ce, "ring count"); break; }
    if ((size_t)total * ring_bytes > scratch_cap && nc > 64) {
      chunk = std::max<int64_t>(64, nc / 2);   // too many rings (large res): retry with a smaller chunk
      continue;
    }
    double* d_ring = nullptr;
    ce = cudaMallocAsync(&d_ring, std::max<size_t>((size_t)total, 1) * ring_bytes, st);
    if (ce != cudaSuccess) { status = cuda_fail(ce, "ring scratch allocation"); break; }
    a.ring_scratch = d_ring;
    double* d_part = nullptr;
    ce = cudaMallocAsync(&d_part, std::max<size_t>((size_t)total_units, 1) * (size_t)E * sizeof(double), st);
    if (ce != cudaSuccess) { cudaFreeAsync(d_ring, st); status = cuda_fail(ce, "partial sums allocation"); break; }
    a.partial = d_part;
    if (total > 0) {
      rb::ag_ring_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(a, total);
      rb::ag_cell_kernel<<<(unsigned)total_units, rb::kAgWarps * 32, cell_smem, st>>>(a, total_units);
      launch_counter()->fetch_add(2);
    }
    rb::ag_finish_kernel<<<(unsigned)std::min<int64_t>(nc, (int64_t)sms * 16), 128, 0, st>>>(a);
    launch_counter()->fetch_add(1);
    ce = cudaGetLastError();
    cudaFreeAsync(d_part, st);
    cudaFreeAsync(d_ring, st);
    if (ce != cudaSuccess) { status = cuda_fail(ce, "afterglow kernels"); break; }
    n0 += nc;
  }
  cudaFreeAsync(d_uoff, st);
  cudaFreeAsync(d_ucnt, st);
  cudaFreeAsync(d_off, st);
  cudaFreeAsync(d_cnt, st);
  cudaFreeAsync(d_scal, st);
  cudaFreeAsync(d_idx, st);
  cudaFreeAsync(d_epoch, st);
  return status;
}

extern "C" int rb_ag_eval_f64_host(const rb_ag_config* cfg, int64_t N, int64_t E, const double* params,
                                   const double* t_obs, const double* freq_obs, const double* dl_cm,
                                   const double* y, const double* sigma, const double* sigma_param,
                                   double* out_model, double* out_logl) {
  int rc = ag_check(cfg, N, E, params, t_obs, freq_obs, y, sigma, sigma_param, out_model, out_logl);
  if (rc != RB_OK) return rc;
  if (N == 0) return RB_OK;
  DevBuf d_par, d_t, d_f, d_dl, d_y, d_s, d_sp, d_m, d_l;
  const size_t nN = (size_t)N, nE = (size_t)E;
  auto up = [&](DevBuf& b, const double* h, size_t n) -> cudaError_t {
    if (!h) return cudaSuccess;
    cudaError_t e = b.alloc(n);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(b.p, h, n * sizeof(double), cudaMemcpyHostToDevice);
  };
  RB_CUDA(up(d_par, params, (size_t)RB_AG_NPARAM * nN));
  RB_CUDA(up(d_t, t_obs, nE));
  RB_CUDA(up(d_f, freq_obs, nE));
  RB_CUDA(up(d_dl, dl_cm, nN));
  RB_CUDA(up(d_y, y, nE));
  RB_CUDA(up(d_s, sigma, nE));
  RB_CUDA(up(d_sp, sigma_param, nN));
  if (out_model) RB_CUDA(d_m.alloc(nN * nE));
  if (out_logl) RB_CUDA(d_l.alloc(nN));
  rc = rb_ag_eval_f64(cfg, N, E, d_par.p, d_t.p, d_f.p, d_dl.p, d_y.p, d_s.p, d_sp.p, d_m.p, d_l.p, nullptr);
  if (rc != RB_OK) return rc;
  if (out_model) RB_CUDA(cudaMemcpy(out_model, d_m.p, nN * nE * sizeof(double), cudaMemcpyDeviceToHost));
  if (out_logl) RB_CUDA(cudaMemcpy(out_logl, d_l.p, nN * sizeof(double), cudaMemcpyDeviceToHost));
  RB_CUDA(cudaDeviceSynchronize());
  return RB_OK;
}
