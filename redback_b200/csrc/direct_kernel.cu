// Supernova-family models WITHOUT an interaction process: the engine luminosity goes straight to TemperatureFloor and
// the SED (sn_fallback / sn_nickel_fallback, supernova_models.py:383-498; and `interaction_process=None` of the
// diffusion models, e.g. :961-968).  Element-wise per (sample, epoch): one block per sample, one thread per epoch,
// the per-sample scalars of sn_prep_kernel, the same epilogue helpers as the CSM path.
// (included by capi.cu after csm_kernel.cu)

namespace rb {

// engine luminosity at source-frame time t [days]
__device__ __forceinline__ double direct_engine(const SnArgs& a, int64_t n, const double* __restrict__ sc, double t) {
  const int64_t S = a.param_stride;
  const double* P = a.params + n;
  switch (a.cfg.engine) {
    case RB_ENGINE_NICKEL:                                                      // supernova_models.py:564-579
      return sc[SC_ENG_A] * (6.45e43 * exp(-t / 8.8) + 1.45e43 * exp(-t / 111.3));
    case RB_ENGINE_MAGNETAR: {                                                  // magnetar_models.py:184
      const double den = fma(t, sc[SC_ENG_B], 1.0);
      return sc[SC_ENG_A] / (den * den);
    }
    case RB_ENGINE_MAGNETAR_NICKEL: {
      const double den = fma(t, sc[SC_ENG_B], 1.0);
      return sc[SC_ENG_A] / (den * den) +
             a.cfg.f_nickel[n] * P[RB_SN_MEJ * S] * (6.45e43 * exp(-t / 8.8) + 1.45e43 * exp(-t / 111.3));   // :1662-1664
    }
    case RB_ENGINE_MAGNETAR_ONLY: {                                             // magnetar_models.py:143
      const double nn = P[3 * S];
      return P[1 * S] * pow(1. + (t * kDay) / (P[2 * S] * kDay), (1. + nn) / (1. - nn));
    }
    case RB_ENGINE_EXP_POWERLAW: {                                              // phenomenological_models.py:753
      const double tp = P[4 * S];
      return P[1 * S] * pow(1 - exp(-t / tp), P[2 * S]) * pow(t / tp, -P[3 * S]);
    }
    case RB_ENGINE_TDE_FALLBACK: {                                              // tde_models.py:23-26
      const double tt = P[2 * S];
      return (t - tt > 0.) ? P[1 * S] / pow(t * 86400, 5. / 3.) : P[1 * S] / pow(tt * 86400, 5. / 3.);
    }
    default: {                                                                  // phenomenological_models.py:697-702
      const double l1 = pow(10.0, P[1 * S]);
      const double ts = t * 86400, trs = P[2 * S] * 86400;
      double lbol = (ts < trs) ? l1 * pow(trs, -5. / 3.) : l1 * pow(ts, -5. / 3.);
      if (a.cfg.engine == RB_ENGINE_FALLBACK_NICKEL)                            // supernova_models.py:467
        lbol = lbol + P[3 * S] * P[RB_SN_MEJ * S] * (6.45e43 * exp(-t / 8.8) + 1.45e43 * exp(-t / 111.3));
      return lbol;
    }
  }
}

__global__ void __launch_bounds__(256) sn_direct_kernel(const SnArgs a) {
  __shared__ double sc[kScalars];
  __shared__ double red[34];
  __shared__ double tab[kExpTabSize];
  const int E = a.E, tid = threadIdx.x;
  for (int i = tid; i < kExpTabSize; i += blockDim.x) tab[i] = RB_EXP2_1024[i];
  for (int64_t n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();
    if (tid < kScalars) sc[tid] = a.scalars[n * kScalars + tid];
    __syncthreads();
    const double opz = sc[SC_OPZ];
    double part = 0.0;
    for (int e = tid; e < E; e += blockDim.x) {
      const double t_in = a.epoch_tab[EP_T * E + e] - sc[SC_T0];
      double model;
      if (a.cfg.t0 != nullptr && !a.grid_mode && t_in < 0.0) {
        model = 0.0;                                                            // phase_models.py:54-58
      } else {
        const double te = a.grid_mode ? (t_in * opz) / opz : t_in / opz;        // utils.py:382
        const double lbol = direct_engine(a, n, sc, te);
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
