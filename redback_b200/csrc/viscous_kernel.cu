// interaction_process = Viscous (redback/interaction_processes.py:229-273): the engine luminosity convolved with
// exp(-(t_e - t) / t_viscous) / t_viscous,
//   L_out(t_e) = trapezoid( L(t_i) exp((t_i - t_e) / t_visc), t_i ) / t_visc,   t_i = t_e * xm,
//   xm = unique(lsp U 1 - lsp), lsp = logspace(log10(t_visc / dense_times[-1]) - 3, 0, 500)      (:252-262)
// on the dense grid linspace(0, time[-1] + 100, dense_resolution) of the calling model, followed by the same
// TemperatureFloor / SED / likelihood epilogue as every supernova path.  One block per sample (grid-stride), one thread
// per epoch walking the 1000 abscissae; the trapezoid is regrouped by node (sum_i f_i (xm_{i+1} - xm_{i-1}) t_e / 2,
// duplicate abscissae are zero-width intervals).  t_viscous is read from the RB_SN_KAPPA row of the parameter block
// (Viscous takes no opacity).  (included by capi.cu after direct_kernel.cu)

namespace rb {

constexpr int kViscHalf = 500;             // int(round(timesteps / 2.0)), timesteps = 1000 (:249, :257)
constexpr int kViscNodes = 2 * kViscHalf;

__host__ __device__ inline size_t viscous_smem_doubles(int D) {
  return 2 * (size_t)D + 2 * (size_t)kViscNodes + (size_t)D + kViscHalf + kViscNodes + kExpTabSize + kScalars + 34;
}

__global__ void __launch_bounds__(256) sn_viscous_kernel(const SnArgs a) {
  extern __shared__ double2 vsm2[];
  const int D = a.cfg.dense_resolution;
  double2* seg = vsm2;                                        // [D]  (a_k, b_k): L(t) = a_k + b_k t on (x_{k-1}, x_k]
  double2* nodes = seg + D;                                   // [kViscNodes] (xm, xm_{i+1} - xm_{i-1})
  double* ytmp = reinterpret_cast<double*>(nodes + kViscNodes);   // [D] engine luminosity on the dense grid
  double* lsp = ytmp + D;                                     // [kViscHalf]
  double* xms = lsp + kViscHalf;                              // [kViscNodes] merged abscissae
  double* tab = xms + kViscNodes;                             // [1024]
  double* sc = tab + kExpTabSize;                             // [kScalars]
  double* red = sc + kScalars;                                // [34]
  const int E = a.E, tid = threadIdx.x, T = blockDim.x;
  for (int i = tid; i < kExpTabSize; i += T) tab[i] = RB_EXP2_1024[i];
  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();
    if (tid < kScalars) sc[tid] = a.scalars[n * kScalars + tid];
    __syncthreads();
    const double opz = sc[SC_OPZ], t_end = sc[SC_TEND];
    const double tvisc = a.params[RB_SN_KAPPA * a.param_stride + n];
    const double step = t_end / (double)(D - 1), inv_step = 1.0 / step;
    for (int k = tid; k < D; k += T) ytmp[k] = direct_engine(a, n, sc, (k == D - 1) ? t_end : (double)k * step);
    {
      const double a0 = log10(tvisc / t_end) - 3.0;           // :258
      const double da = (0.0 - a0) / (double)(kViscHalf - 1);
      for (int j = tid; j < kViscHalf; j += T) lsp[j] = pow(10.0, (j == kViscHalf - 1) ? 0.0 : (double)j * da + a0);
    }
    __syncthreads();
    for (int k = tid; k < D; k += T) {                        // scipy interp1d (linear), :246
      double2 sg = make_double2(ytmp[0], 0.0);                // t = 0 (the abscissa xm = 0) lands here: L(0)
      if (k > 0) {
        const double y0 = ytmp[k - 1];
        const double sk = (ytmp[k] - y0) * inv_step;
        sg = make_double2(fma(-sk, (double)(k - 1) * step, y0), sk);
      }
      seg[k] = sg;
    }
    // merge lsp (ascending) with 1 - lsp (non-increasing) in sorted order: rank = own index + number of elements of the
    // other list that sort before it
    for (int r = tid; r < kViscNodes; r += T) {
      int rank;
      double v;
      if (r < kViscHalf) {
        v = lsp[r];
        int lo = 0, hi = kViscHalf;      // first j with (1 - lsp[j]) < v; the set is a suffix
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((1.0 - lsp[mid]) < v) hi = mid; else lo = mid + 1;
        }
        rank = r + (kViscHalf - lo);
      } else {
        const int j0 = r - kViscHalf;
        v = 1.0 - lsp[j0];
        int lo = 0, hi = kViscHalf;      // number of j with lsp[j] <= v; the set is a prefix
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (lsp[mid] <= v) lo = mid + 1; else hi = mid;
        }
        rank = (kViscHalf - 1 - j0) + lo;
      }
      xms[rank] = fmin(fmax(v, 0.0), 1.0);                    // np.clip(tb + (t_e - tb) xm, tb, dense_times[-1]), :260
    }
    __syncthreads();
    for (int r = tid; r < kViscNodes; r += T) {
      const double x = xms[r];
      const double dw = xms[min(r + 1, kViscNodes - 1)] - xms[max(r - 1, 0)];
      nodes[r] = make_double2(x, dw);
    }
    __syncthreads();

    const double inv_tv = 1.0 / tvisc;
    double part = 0.0;
    for (int e = tid; e < E; e += T) {
      const double t_in = a.epoch_tab[EP_T * E + e];
      const double te = a.grid_mode ? (t_in * opz) / opz : t_in / opz;          // utils.py:382
      // epochs before the grid take the first valid epoch's luminosity (np.searchsorted of an excluded time, :271)
      const double td = (te >= 0.0 || !a.cfg.negative_epochs) ? te : a.cfg.t_first_valid / opz;
      // first abscissa whose weight exp((xm - 1) t_e / t_visc) is above exp(-691)
      int m0 = 0;
      {
        const double lim = 691.0 * tvisc;
        if (td > lim) {
          int lo = 0, hi = kViscNodes;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((1.0 - xms[mid]) * td > lim) lo = mid + 1; else hi = mid;
          }
          m0 = lo;
        }
      }
      const double te_is = td * inv_step * (1.0 - 0x1p-50);
      double acc = 0.0;
#pragma unroll 4
      for (int m = m0; m < kViscNodes; ++m) {
        const double2 nd = nodes[m];
        const double t = td * nd.x;
        const double v = __fma_ru(nd.x, te_is, 4503599627370496.0);             // 2^52 + ceil(t / step)
        const unsigned k = min((unsigned)__double2loint(v), (unsigned)(D - 1));
        const double2 sg = seg[k];
        const double lum = fma(sg.y, t, sg.x);
        double g = lum * exp_tab(fmax((t - td) * inv_tv, -700.0), tab);        // :264
        if (isnan(g)) g = 0.0;                                                  // :265
        acc = fma(g, nd.y, acc);
      }
      const double lbol = (0.5 * td * acc) * inv_tv;                            // :267

      double model;
      if (a.cfg.sed == RB_SED_BOLOMETRIC) {
        model = lbol;
      } else {
        double tph, rph;
        temperature_floor_dev(lbol, te, sc[SC_RCV], sc[SC_TFLOOR], sc[SC_INV_REC], tph, rph);
        if (a.grid_mode) {
          double* go = a.grid_out + (size_t)n * 4 * E;
          go[0 * E + e] = tph;
          go[1 * E + e] = rph;
          go[2 * E + e] = lbol;
          go[3 * E + e] = t_in * opz;
          continue;
        }
        model = sed_mjy(a.cfg, lbol, te, tph, rph, a.epoch_tab[EP_NU * E + e], opz, sc[SC_INV_DEN], sc[SC_DL], tab);
      }
      if (a.out_model != nullptr) a.out_model[n * (int64_t)E + e] = model;
      if (a.want_logl)
        part += logl_term(a.cfg.sigma_mode, a.epoch_tab[EP_KIND * E + e], a.epoch_tab[EP_Y * E + e], model,
                          a.epoch_tab[EP_ISIG * E + e], a.epoch_tab[EP_LOGT * E + e], sc[SC_SIGMA]);
    }
    if (a.want_logl) {
      const double total = block_sum(part, red);
      if (tid == 0) a.out_logl[n] = nan_to_num(total);
    }
  }
}

}  // namespace rb
