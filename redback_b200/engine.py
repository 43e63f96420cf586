"""Device-resident evaluation paths: epochs / frequencies / data are uploaded once, parameter batches
stream through the fused kernels.  torch is used for device memory and streams only."""
import ctypes as C

import numpy as np
import torch

from . import _capi

_F64 = torch.float64


def _device(device=None):
    if not torch.cuda.is_available():
        raise _capi.RedbackB200Error(
            "redback_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def as_device_f64(x, device):
    """1-D float64 contiguous device tensor from a python scalar / sequence / ndarray / tensor."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=_F64).contiguous()
    arr = np.array(x, dtype=np.float64, order="C", copy=True)   # own, writable buffer (inputs may be read-only views)
    return torch.from_numpy(arr).to(device)


def _precision(dtype):
    name = str(dtype).replace("torch.", "").replace("numpy.", "")
    if name in ("float64", "f64", "double", "<class 'float'>"):
        return _capi.PREC_F64
    if name in ("float32", "f32", "single"):
        return _capi.PREC_F32
    raise ValueError(f"dtype must be float64 or float32, got {dtype!r}")


def batch_size(params):
    n = None
    for k, v in params.items():
        if isinstance(v, torch.Tensor):
            m = v.numel() if v.dim() > 0 else None
        elif np.ndim(v) > 0:
            m = int(np.size(v))
        else:
            m = None
        if m is not None:
            if n is not None and m != n:
                raise ValueError(f"params_batch entries have different lengths ({n} vs {m} for '{k}')")
            n = m
    return n


class _FusedPath:
    """Common host side of a fused model+likelihood path: epochs / frequencies / data resident on device,
    SoA packing of parameter batches, stream-ordered launch through the C ABI."""

    ROWS = {}
    NPARAM = 0
    NEEDS_FREQUENCY = True

    def _eval_fn(self):
        raise NotImplementedError

    def _check_time(self, t):
        pass

    def _launch(self, N, params_dev, dl, y, sigma, sigma_param, model, logl, stream):
        return self._eval_fn()(C.byref(self.cfg), N, self.E, _ptr(params_dev), _ptr(self.time), _ptr(self.frequency),
                               _ptr(dl), _ptr(y), _ptr(sigma), _ptr(sigma_param), _ptr(model), _ptr(logl),
                               C.c_void_p(stream))

    def _launch_mag(self, fn, N, params_dev, dl, y, sigma, sigma_param, model, logl, stream):
        return fn(C.byref(self.cfg), C.addressof(self.phot), N, self.E, _ptr(params_dev), _ptr(self.time), _ptr(dl),
                  _ptr(y), _ptr(sigma), _ptr(sigma_param), _ptr(model), _ptr(logl), C.c_void_p(stream))

    def _setup(self, time, frequency, y, sigma, device):
        self.device = _device(device)
        t = np.atleast_1d(np.asarray(time.detach().cpu().numpy() if isinstance(time, torch.Tensor) else time,
                                     dtype=np.float64))
        if t.ndim != 1 or t.size == 0:
            raise ValueError("time must be a non-empty 1-D array")
        self._check_time(t)
        self.E = int(t.size)
        self._h_time = np.ascontiguousarray(t)               # host copies: the *_host entry points take host pointers
        self._h_freq = self._h_y = self._h_sigma = self._h_ul = None
        self.time = torch.from_numpy(self._h_time).to(self.device)
        self.frequency = None
        if self.NEEDS_FREQUENCY:
            if frequency is None:
                raise KeyError("frequency")  # the reference does kwargs['frequency']
            f = np.asarray(frequency.detach().cpu().numpy() if isinstance(frequency, torch.Tensor) else frequency,
                           dtype=np.float64)
            if f.ndim == 0:
                f = np.full(self.E, float(f))
            if f.shape != (self.E,):
                raise ValueError("frequency must be a single number or the same length as time")
            self._h_freq = np.ascontiguousarray(f)
            self.frequency = torch.from_numpy(self._h_freq).to(self.device)
        self.y = self.sigma = None
        if y is not None:
            self.y = as_device_f64(y, self.device)
            self._h_y = self.y.cpu().numpy()
            if self.y.shape != (self.E,):
                raise ValueError("y must have the same length as time")
            if sigma is not None:
                s = np.asarray(sigma, dtype=np.float64) if not isinstance(sigma, torch.Tensor) else sigma
                if np.ndim(s) == 0:
                    s = np.full(self.E, float(s))
                self.sigma = as_device_f64(s, self.device)
                if self.sigma.shape != (self.E,):
                    raise ValueError('Sigma must be either float or array-like x.')
                self._h_sigma = self.sigma.cpu().numpy()

    def set_extinction(self, ext, av_host):
        """Attenuate the spectra of a magnitude / flux path per observer-frame wavelength before the spline fit
        (extinction_models.py:223-251): `ext` a _capi.Extinction, `av_host` a contiguous float64 device tensor [N] that
        must match the next evaluate(); None clears it."""
        if not hasattr(self, "phot"):
            raise NotImplementedError("extinction of the spectra needs a magnitude / flux path")
        self._ext_keep = (ext, av_host)
        self.phot.extinction = C.addressof(ext) if ext is not None else None
        self.phot.ext_av_host = av_host.data_ptr() if av_host is not None else None

    SPECTRA_TIME_SCALE = 1.0      # observer-frame days of the model grid -> the unit of the reference's `time` field

    def _spectra_grid_points(self):
        raise NotImplementedError("output_format='spectra' needs a magnitude / flux path")

    def evaluate_spectra(self, params_dev, dl=None, aux=None):
        """output_format='spectra' (supernova_models.py:1019-1023 and its twins): (time [N, n_t], spectra [N, n_t, n_lambda]
        erg/cm^2/s/Angstrom, grid_points [N]) on the model's own time grid, before any clean-up; rows beyond
        grid_points[n] are NaN.  The path must have been built with output='spectra'."""
        if getattr(self, "phot", None) is None or self.phot.output != 2:
            raise ValueError("the path was not built with output='spectra'")
        N, n_t, n_l = params_dev.shape[1], self._spectra_grid_points(), self.phot.n_lambda
        spectra = torch.empty((N, n_t, n_l), dtype=_F64, device=self.device)
        tt = torch.empty((N, n_t), dtype=_F64, device=self.device)
        count = torch.empty((N,), dtype=torch.int32, device=self.device)
        self.phot.spectra_out, self.phot.spectra_time_out = spectra.data_ptr(), tt.data_ptr()
        self.phot.spectra_len_out = count.data_ptr()
        try:
            self.evaluate(params_dev, dl=dl, aux=aux)
        finally:
            self.phot.spectra_out = self.phot.spectra_time_out = self.phot.spectra_len_out = None
        if self.SPECTRA_TIME_SCALE != 1.0:
            tt *= self.SPECTRA_TIME_SCALE
        return tt, spectra, count

    def set_upper_limits(self, sigma_level, magnitude_mode=False):
        """Mark epochs with sigma_level[e] > 0 as upper limits at that many sigma
        (GaussianLikelihoodWithUpperLimits, likelihoods.py:234-489); None clears them."""
        if sigma_level is None:
            self._ul_level = self._h_ul = None
            self.cfg.upper_limits.sigma_level = None
            self.cfg.upper_limits.magnitude_mode = 0
            return
        self._ul_level = as_device_f64(sigma_level, self.device)   # kept alive by the path
        if self._ul_level.shape != (self.E,):
            raise ValueError("upper-limit sigma levels must have the same length as time")
        self._h_ul = self._ul_level.cpu().numpy()
        self.cfg.upper_limits.sigma_level = self._ul_level.data_ptr()
        self.cfg.upper_limits.magnitude_mode = int(bool(magnitude_mode))

    def pack(self, params, N):
        """SoA [NPARAM, N] device tensor from {name: scalar | array[N] | tensor[N]}.  Host values (scalars and
        numpy arrays) are written row by row into ONE pinned block that goes up in a single asynchronous copy;
        device tensors are copied row-wise on the device."""
        out = torch.empty((self.NPARAM, N), dtype=_F64, device=self.device)
        stage = torch.zeros((self.NPARAM, N), dtype=_F64, pin_memory=True)
        view = stage.numpy()
        dev_rows = []
        for name, val in params.items():
            if name not in self.ROWS:
                continue
            row = self.ROWS[name]
            if isinstance(val, torch.Tensor):
                if val.is_cuda:
                    dev_rows.append((row, val))
                else:
                    view[row] = val.detach().to(_F64).reshape(-1).numpy()
            else:
                view[row] = np.asarray(val, dtype=np.float64).reshape(-1) if np.ndim(val) else float(val)
        out.copy_(stage, non_blocking=True)
        for row, val in dev_rows:
            v = val.to(device=self.device, dtype=_F64).reshape(-1)
            out[row] = v.expand(N) if v.numel() == 1 else v
        # `stage` may be freed by the caching host allocator only after the copy has run (stream-ordered reuse)
        return out

    # ---- host-resident batches: the C ABI's *_host_rows entry (chunked, double-buffered pipeline in the library) ----
    def _host_rows_fn(self):
        return None

    def supports_host_rows(self):
        return self._host_rows_fn() is not None

    def evaluate_host(self, params, N, dl=None, sigma_param=None, t0=None, aux=None, want_model=False, want_logl=True):
        """{name: scalar | ndarray[N]} -> (model ndarray[N, E] | None, logl ndarray[N] | None) through
        rb_*_eval_f64_host_rows: the arrays are read where they lie (no packing copy on the Python side)."""
        fn = self._host_rows_fn()
        if fn is None:
            raise NotImplementedError("this path has no host-buffer entry point")
        if want_logl and self._h_y is None:
            raise ValueError("log-likelihood requested but no data were given")
        keep = []

        def host(v, n):
            arr = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)))
            keep.append(arr)
            return arr.ctypes.data_as(C.c_void_p)

        rows = (C.c_void_p * self.NPARAM)()
        fill = (C.c_double * self.NPARAM)()
        for name, val in params.items():
            if name not in self.ROWS:
                continue
            row = self.ROWS[name]
            if np.ndim(val) == 0:
                rows[row], fill[row] = None, float(val)
            else:
                arr = np.ascontiguousarray(np.asarray(val, dtype=np.float64).reshape(-1))
                if arr.size != N:
                    raise ValueError(f"parameter '{name}' has {arr.size} values, expected {N}")
                keep.append(arr)
                rows[row] = arr.ctypes.data
        cfg = type(self.cfg).from_buffer_copy(self.cfg)
        cfg.upper_limits.sigma_level = self._h_ul.ctypes.data if self._h_ul is not None else None
        if t0 is not None:
            if not hasattr(cfg, "t0"):
                raise NotImplementedError("a per-sample t0 is not available on this path")
            cfg.t0 = host(t0, N)
        aux_name = getattr(self, "AUX_PARAM", None)
        if aux_name is not None:
            if aux is None:
                raise ValueError(f"this path needs the per-sample parameter '{aux_name}'")
            setattr(cfg, aux_name, host(aux, N))
        model = np.empty((N, self.E), dtype=np.float64) if want_model else None
        logl = np.empty((N,), dtype=np.float64) if want_logl else None
        if N == 0:
            return model, logl
        ptr = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        with torch.cuda.device(self.device):
            rc = fn(C.byref(cfg), N, self.E, rows, fill, ptr(self._h_time), ptr(self._h_freq),
                    host(dl, N) if dl is not None else None, ptr(self._h_y) if want_logl else None,
                    ptr(self._h_sigma) if want_logl else None,
                    host(sigma_param, N) if (sigma_param is not None and want_logl) else None, ptr(model), ptr(logl))
        _capi.check(rc)
        return model, logl

    def evaluate(self, params_dev, dl=None, sigma_param=None, want_model=True, want_logl=False, out_logl=None,
                 t0=None, aux=None):
        """Launch the fused kernel on the current stream.  Returns (model[N,E] | None, logl[N] | None).
        `t0`: optional device tensor [N] of explosion epochs (phase_models.t0_base_model; supernova paths only).
        `aux`: device tensor [N] of the path's extra per-sample parameter `AUX_PARAM` (f_nickel of magnetar_nickel)."""
        if params_dev.dtype != _F64 or params_dev.dim() != 2 or params_dev.shape[0] != self.NPARAM:
            raise ValueError(f"params_dev must be float64 [{self.NPARAM}, N]")
        params_dev = params_dev.contiguous()
        N = params_dev.shape[1]
        model = torch.empty((N, self.E), dtype=_F64, device=self.device) if want_model else None
        logl = None
        if want_logl:
            logl = out_logl if out_logl is not None else torch.empty((N,), dtype=_F64, device=self.device)
            if logl.shape != (N,) or logl.dtype != _F64 or not logl.is_contiguous():
                raise ValueError("out_logl must be a contiguous float64 [N] device tensor")
        if want_logl and self.y is None:
            raise ValueError("log-likelihood requested but no data were given")
        if N == 0:      # empty batch: nothing to launch
            return model, logl
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if t0 is not None:
            if not hasattr(self.cfg, "t0"):
                raise NotImplementedError("a per-sample t0 is only available on the supernova paths")
            if t0.dtype != _F64 or t0.shape != (N,) or not t0.is_contiguous():
                raise ValueError("t0 must be a contiguous float64 [N] device tensor")
            self.cfg.t0 = t0.data_ptr()
        aux_name = getattr(self, "AUX_PARAM", None)
        if aux_name is not None:
            if aux is None or aux.dtype != _F64 or aux.shape != (N,) or not aux.is_contiguous():
                raise ValueError(f"this path needs aux = contiguous float64 [N] device tensor of '{aux_name}'")
            setattr(self.cfg, aux_name, aux.data_ptr())
        try:
            with torch.cuda.device(self.device):
                rc = self._launch(N, params_dev, dl, self.y if want_logl else None,
                                  self.sigma if want_logl else None, sigma_param if want_logl else None, model, logl,
                                  stream)
        finally:
            if t0 is not None:
                self.cfg.t0 = None
            if aux_name is not None:
                setattr(self.cfg, aux_name, None)
        _capi.check(rc)
        return model, logl


    def log_likelihood_host(self, params_host, out_host=None, chunk=1 << 21):
        """End-to-end entry for host-resident batches: `params_host` is a (pinned) CPU float64 tensor
        [RB_SN_NPARAM, N]; chunks are copied H2D, evaluated and the log-likelihoods copied back D2H on
        two alternating streams so that copies overlap compute.  Returns a pinned CPU tensor [N]."""
        if params_host.dtype != _F64 or params_host.dim() != 2 or params_host.shape[0] != self.NPARAM:
            raise ValueError(f"params_host must be float64 [{self.NPARAM}, N]")
        N = params_host.shape[1]
        if out_host is None:
            out_host = torch.empty((N,), dtype=_F64, pin_memory=True)
        chunk = max(1, min(int(chunk), N))
        if getattr(self, "_slots", None) is None or self._slots[0][0].shape[1] != chunk:
            self._slots = [(torch.empty((self.NPARAM, chunk), dtype=_F64, device=self.device),
                            torch.empty((chunk,), dtype=_F64, device=self.device),
                            torch.cuda.Stream(self.device)) for _ in range(2)]
        cur = torch.cuda.current_stream(self.device)
        for _, _, st in self._slots:
            st.wait_stream(cur)
        for i, a in enumerate(range(0, N, chunk)):
            b = min(a + chunk, N)
            dbuf, lbuf, st = self._slots[i % 2]
            with torch.cuda.stream(st):
                if b - a == chunk:
                    view = dbuf
                    for r in range(self.NPARAM):
                        dbuf[r].copy_(params_host[r, a:b], non_blocking=True)
                else:
                    view = torch.empty((self.NPARAM, b - a), dtype=_F64, device=self.device)
                    for r in range(self.NPARAM):
                        view[r].copy_(params_host[r, a:b], non_blocking=True)
                _, logl = self.evaluate(view, want_model=False, want_logl=True,
                                        out_logl=lbuf[: b - a] if b - a == chunk else None)
                out_host[a:b].copy_(logl, non_blocking=True)
        for _, _, st in self._slots:
            cur.wait_stream(st)
        torch.cuda.current_stream(self.device).synchronize()
        return out_host


class SupernovaPath(_FusedPath):
    """Fused Diffusion -> TemperatureFloor -> Blackbody|CutoffBlackbody -> Gaussian log L.

    Mirrors the reference call chain GaussianLikelihood.log_likelihood -> arnett / slsn / ... ->
    Diffusion (redback/interaction_processes.py:33-65) -> TemperatureFloor (redback/photosphere.py:56-111)
    -> Blackbody / CutoffBlackbody (redback/sed.py:199-222, 301-426), for N samples at once.
    """

    ROWS = _capi.SN_ROWS
    NPARAM = _capi.SN_NPARAM

    def __init__(self, engine, sed, time, frequency=None, y=None, sigma=None,
                 sigma_mode=_capi.SIGMA_FIXED, dense_resolution=1000, cutoff_wavelength=3000.0,
                 absorption_index=1.0, cosmology=None, device=None, line=None, dtype="float64", csm=None,
                 no_interaction=False, area_ratio=None, viscous=False):
        self.NEEDS_FREQUENCY = sed != _capi.SED_BOLOMETRIC
        self.cfg = _capi.sn_config(engine, sed, sigma_mode, dense_resolution, cutoff_wavelength,
                                   absorption_index, cosmology, line, _precision(dtype), csm, no_interaction,
                                   area_ratio, viscous)
        self._csm_nickel = bool(csm and csm.get("nickel"))
        self._engine_rows(engine)
        if viscous:      # interaction_processes.Viscous takes t_viscous instead of the opacities: it rides in kappa's row
            self.ROWS = dict(self.ROWS, t_viscous=self.ROWS["kappa"])
        self._setup(time, frequency, y, sigma, device)

    def _eval_fn(self):
        return _capi.lib().rb_sn_eval_f64

    def _host_rows_fn(self):
        return _capi.lib().rb_sn_eval_f64_host_rows

    def _engine_rows(self, engine):
        if engine == _capi.ENGINE_MAGNETAR_NICKEL:
            # f_nickel and p0 share row 1 of the parameter block: f_nickel travels through cfg.f_nickel instead
            self.ROWS = {k: v for k, v in _capi.SN_ROWS.items() if k != "f_nickel"}
            self.AUX_PARAM = "f_nickel"
        elif engine == _capi.ENGINE_FALLBACK_NICKEL:
            self.ROWS = dict(_capi.SN_ROWS, f_nickel=3)      # rows 1-3 = logl1, tr, f_nickel
        elif engine == _capi.ENGINE_CSM and getattr(self, "_csm_nickel", False):
            # csm_nickel: f_nickel shares row 1 with csm_mass, so it travels through cfg.f_nickel
            self.ROWS = {k: v for k, v in _capi.SN_ROWS.items() if k != "f_nickel"}
            self.AUX_PARAM = "f_nickel"

    def _check_time(self, t):
        if self.cfg.engine == _capi.ENGINE_CSM:
            # the dense grid is np.geomspace(min(0.1, min(time)), ...) (supernova_models.py:2026-2030)
            if np.any(~(t > 0)):
                raise ValueError("Geometric sequence cannot include zero: csm_interaction needs time > 0")
            return
        if self.cfg.no_interaction or self.cfg.engine >= _capi.ENGINE_FALLBACK:
            return                                   # the engine is evaluated at the epochs themselves, any sign
        if np.any(t < 0):
            # interaction_processes.py:47, :63: epochs before the dense grid get the first valid epoch's luminosity
            valid = t[t >= 0]
            if valid.size == 0:
                raise IndexError("index 0 is out of bounds for axis 0 with size 0")      # uniq_lums is empty
            plain = self.cfg.engine in (_capi.ENGINE_NICKEL, _capi.ENGINE_MAGNETAR, _capi.ENGINE_MAGNETAR_NICKEL) and \
                not self.cfg.aspherical
            if not (plain or self.cfg.viscous):
                raise NotImplementedError("negative times are fused for the nickel / magnetar engines with Diffusion and "
                                          "for Viscous only")
            self.cfg.t_first_valid = float(valid.min())
            self.cfg.negative_epochs = 1
        if np.max(t) > t[-1] + 100:
            # interaction_processes.py:63 would index past uniq_lums here
            raise IndexError("max(time) exceeds time[-1] + 100: outside the reference's dense grid")


class KilonovaPath(_FusedPath):
    """Fused one-component / Metzger kilonova -> interp1d -> blackbody -> Gaussian log L
    (redback/transient_models/kilonova_models.py:1537-1603, 1853-2068)."""

    ROWS = _capi.KN_ROWS
    NPARAM = _capi.KN_NPARAM

    def __init__(self, model, time, frequency=None, y=None, sigma=None, sigma_mode=_capi.SIGMA_FIXED,
                 dense_resolution=500, neutron_precursor_switch=True, vmax=0.7, cosmology=None, device=None,
                 grid_t_max=7e6, dtype="float64"):
        precision = _precision(dtype)
        if precision != _capi.PREC_F64 and model != _capi.KN_ONE_COMPONENT:
            raise NotImplementedError("dtype='float32' is available for one_component_kilonova_model only")
        self.cfg = _capi.kn_config(model, sigma_mode, dense_resolution, neutron_precursor_switch, vmax, cosmology,
                                   grid_t_max, precision)
        self._setup(time, frequency, y, sigma, device)

    def _eval_fn(self):
        return _capi.lib().rb_kn_eval_f64

    def _host_rows_fn(self):
        return _capi.lib().rb_kn_eval_f64_host_rows

    def _check_time(self, t):
        if np.any(t <= 0):
            raise ValueError("A value in x_new is below the interpolation range.")  # scipy interp1d's error


class SupernovaMagPath(SupernovaPath):
    """Magnitude / bandpass-flux branch of arnett / basic_magnetar_powered / slsn
    (supernova_models.py:1008-1028, 1510-1530, 1599-1624): model on get_optimal_time_array(0.1, 3000, 300) x
    lambda_array, bicubic spline, band integrals -- all on device (csrc/phot_kernel.cu)."""

    def __init__(self, engine, sed, time, bands, output="magnitude", lambda_array=None, y=None, sigma=None,
                 sigma_mode=_capi.SIGMA_FIXED, dense_resolution=1000, cutoff_wavelength=3000.0, absorption_index=1.0,
                 cosmology=None, device=None, line=None, csm=None, no_interaction=False, area_ratio=None,
                 viscous=False):
        from . import bands as _bands
        if sed == _capi.SED_BOLOMETRIC:
            raise ValueError("bolometric models have no magnitude output")
        self.NEEDS_FREQUENCY = False
        self.cfg = _capi.sn_config(engine, sed, sigma_mode, dense_resolution, cutoff_wavelength, absorption_index,
                                   cosmology, line, 0, csm, no_interaction, area_ratio, viscous)
        self._csm_nickel = bool(csm and csm.get("nickel"))
        self._engine_rows(engine)
        if viscous:
            self.ROWS = dict(self.ROWS, t_viscous=self.ROWS["kappa"])
        self._setup(time, None, y, sigma, device)
        lam = np.geomspace(100, 60000, 100) if lambda_array is None else np.asarray(lambda_array, dtype=np.float64)
        # get_optimal_time_array(0.1, 3000, 300) (utils.py:108); csm_interaction: (0.1, 500, 300) (:2095)
        tgrid = np.geomspace(0.1, 500 if engine == _capi.ENGINE_CSM and not self._csm_nickel else 3000, 300)   # csm_nickel: :2213
        self.phot, self._keep = _bands.build_photometry(bands, self.E, lam, output, tgrid=tgrid)

    def _launch(self, *args):
        return self._launch_mag(_capi.lib().rb_sn_mag_eval_f64, *args)

    def _spectra_grid_points(self):
        return int(self.phot.n_tgrid)

    def evaluate_ragged(self, params_dev, phase, band_index, epoch_count, dl=None):
        """Magnitudes at per-sample pointings (rb_sn_mag_eval_ragged_f64): `phase` [N, E_max] float64 (days since
        explosion, observer frame), `band_index` [N, E_max] int32 into this path's sorted band list, `epoch_count` [N]
        int32; the path must have been built with bands = its sorted unique band names, one epoch per band.
        Returns [N, E_max] (padding entries NaN)."""
        N, E = phase.shape
        if params_dev.shape != (self.NPARAM, N) or band_index.shape != (N, E) or epoch_count.shape != (N,):
            raise ValueError("params [NPARAM, N], phase / band_index [N, E_max] and epoch_count [N] must agree")
        if phase.dtype != _F64 or band_index.dtype != torch.int32 or epoch_count.dtype != torch.int32:
            raise ValueError("phase must be float64, band_index and epoch_count int32")
        out = torch.full((N, E), float("nan"), dtype=_F64, device=self.device)
        if N == 0 or E == 0:
            return out
        with torch.cuda.device(self.device):
            rc = _capi.lib().rb_sn_mag_eval_ragged_f64(
                C.byref(self.cfg), C.addressof(self.phot), N, E, _ptr(params_dev.contiguous()), _ptr(phase.contiguous()),
                _ptr(band_index.contiguous()), _ptr(epoch_count.contiguous()), _ptr(dl), _ptr(out),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _capi.check(rc)
        return out

    def _host_rows_fn(self):
        return None     # the magnitude branch has device entry points only


class KilonovaMagPath(KilonovaPath):
    """Magnitude / bandpass-flux branch of one_component_kilonova_model / metzger_kilonova_model
    (kilonova_models.py:1579-1603, 1937-1962)."""

    NEEDS_FREQUENCY = False

    def __init__(self, model, time, bands, output="magnitude", lambda_array=None, y=None, sigma=None,
                 sigma_mode=_capi.SIGMA_FIXED, dense_resolution=500, neutron_precursor_switch=True, vmax=0.7,
                 cosmology=None, device=None):
        from . import bands as _bands
        self.cfg = _capi.kn_config(model, sigma_mode, dense_resolution, neutron_precursor_switch, vmax, cosmology)
        self._setup(time, None, y, sigma, device)
        lam = np.geomspace(100, 60000, 200) if lambda_array is None else np.asarray(lambda_array, dtype=np.float64)
        self.phot, self._keep = _bands.build_photometry(bands, self.E, lam, output)

    def _launch(self, *args):
        return self._launch_mag(_capi.lib().rb_kn_mag_eval_f64, *args)

    SPECTRA_TIME_SCALE = 86400.0        # time_observer_frame in seconds (kilonova_models.py:1594-1600)

    def _spectra_grid_points(self):
        return int(self.cfg.dense_resolution)

    def _host_rows_fn(self):
        return None     # the magnitude branch has device entry points only


class KilonovaMultiMagPath(KilonovaMagPath):
    """Magnitude / bandpass-flux branch of two_component_kilonova_model / three_component_kilonova_model
    (kilonova_models.py:1166-1211, 1277-1323): the components' blackbody spectra are summed on the shared grid.
    Parameter rows: the one-component rows once per component, named `<row>_<component>` (every `redshift_<c>` row
    holds the same redshift)."""

    def __init__(self, n_components, time, bands, grid_t_max, **kwargs):
        super().__init__(_capi.KN_ONE_COMPONENT, time, bands, **kwargs)
        self.cfg.grid_t_max = float(grid_t_max)
        self.n_components = int(n_components)
        self.NPARAM = self.n_components * _capi.KN_NPARAM
        self.ROWS = {f"{k}_{c + 1}": c * _capi.KN_NPARAM + r for c in range(self.n_components)
                     for k, r in _capi.KN_ROWS.items() if k != "beta"}

    def _launch(self, N, params_dev, dl, y, sigma, sigma_param, model, logl, stream):
        return _capi.lib().rb_kn_multi_mag_eval_f64(
            C.byref(self.cfg), C.addressof(self.phot), self.n_components, N, self.E, _ptr(params_dev), _ptr(self.time),
            _ptr(dl), _ptr(y), _ptr(sigma), _ptr(sigma_param), _ptr(model), _ptr(logl), C.c_void_p(stream))


class AfterglowPath(_FusedPath):
    """redback-native tophat / Gaussian afterglow -> Gaussian log L
    (redback/transient_models/afterglow_models/base_models.py:43-640, 885-1010)."""

    ROWS = _capi.AG_ROWS
    NPARAM = _capi.AG_NPARAM

    def __init__(self, jet, time, frequency=None, y=None, sigma=None, sigma_mode=_capi.SIGMA_FIXED, steps=250,
                 res=None, k=0, expansion=1, a1=1, cosmology=None, device=None, ss=0.01, aa=0.5,
                 ab_magnitude=False, dtype="float64"):
        if k not in (0, 2):
            raise ValueError("k must either be 0 or 2")
        self.cfg = _capi.ag_config(jet, sigma_mode, steps, 0 if res is None else int(res), k, expansion, a1, cosmology,
                                   ss, aa, ab_magnitude, _precision(dtype))
        self._setup(time, frequency, y, sigma, device)

    def _eval_fn(self):
        return _capi.lib().rb_ag_eval_f64

    def _host_rows_fn(self):
        return _capi.lib().rb_ag_eval_f64_host_rows


class MergernovaPath(_FusedPath):
    """Magnetar-driven relativistic ejecta -> relativistic blackbody -> Gaussian log L
    (redback/transient_models/magnetar_driven_ejecta_models.py:14-151, 240-419)."""

    ROWS = _capi.MN_ROWS
    NPARAM = _capi.MN_NPARAM

    def __init__(self, engine, time, frequency=None, y=None, sigma=None, sigma_mode=_capi.SIGMA_FIXED,
                 use_gamma_ray_opacity=False, pair_cascade_switch=False, ejecta_albedo=0.5,
                 pair_cascade_fraction=0.05, cosmology=None, device=None, output=0, photon_index=2.0,
                 supernova=None):
        self.cfg = _capi.mn_config(engine, sigma_mode, use_gamma_ray_opacity, pair_cascade_switch, ejecta_albedo,
                                   pair_cascade_fraction, cosmology, output, photon_index, supernova)
        if engine == _capi.MN_EVOLVING:
            self.ROWS = _capi.MN_ROWS_EVOLVING
        if supernova is not None:
            self.ROWS = _capi.MN_ROWS_SUPERNOVA
        self.NEEDS_FREQUENCY = output != _capi.MN_OUT_LBOL
        self._setup(time, frequency, y, sigma, device)

    def _eval_fn(self):
        return _capi.lib().rb_mn_eval_f64

    def _host_rows_fn(self):
        return _capi.lib().rb_mn_eval_f64_host_rows

    def _check_time(self, t):
        if np.any(~(t > 0)):
            # scipy's interp1d raises below the model grid (starts at 1e-4 s, source frame)
            raise ValueError("A value in x_new is below the interpolation range: mergernova models need time > 0")


class MetzgerMagnetarPath(_FusedPath):
    """Magnetar-driven Metzger shell model -> blackbody -> Gaussian log L
    (redback/transient_models/magnetar_driven_ejecta_models.py:576-905)."""

    ROWS = _capi.MK_ROWS
    NPARAM = _capi.MK_NPARAM

    def __init__(self, engine, time, frequency=None, y=None, sigma=None, sigma_mode=_capi.SIGMA_FIXED,
                 use_gamma_ray_opacity=False, pair_cascade_switch=True, ejecta_albedo=0.5, pair_cascade_fraction=0.01,
                 neutron_precursor_switch=True, vmax=0.7, cosmology=None, device=None, resolution=300,
                 magnetar_heating="first_layer"):
        self.cfg = _capi.mk_config(engine, sigma_mode, use_gamma_ray_opacity, pair_cascade_switch, ejecta_albedo,
                                   pair_cascade_fraction, neutron_precursor_switch, vmax, cosmology, resolution,
                                   magnetar_heating)
        if engine == _capi.MN_EVOLVING:
            self.ROWS = _capi.MK_ROWS_EVOLVING
        self._setup(time, frequency, y, sigma, device)

    def _eval_fn(self):
        return _capi.lib().rb_mk_eval_f64

    def _host_rows_fn(self):
        return _capi.lib().rb_mk_eval_f64_host_rows

    def _check_time(self, t):
        if np.any(~(t > 0)):
            raise ValueError("A value in x_new is below the interpolation range: these models need time > 0")


class MetzgerMagnetarMagPath(MetzgerMagnetarPath):
    """Magnitude / bandpass-flux branch of the magnetar-driven Metzger shell models (_processing_other_formats,
    magnetar_driven_ejecta_models.py:198-238, plain blackbody)."""

    NEEDS_FREQUENCY = False

    def __init__(self, engine, time, bands, output="magnitude", lambda_array=None, y=None, sigma=None,
                 sigma_mode=_capi.SIGMA_FIXED, use_gamma_ray_opacity=False, pair_cascade_switch=True, ejecta_albedo=0.5,
                 pair_cascade_fraction=0.01, neutron_precursor_switch=True, vmax=0.7, cosmology=None, device=None,
                 resolution=300, magnetar_heating="first_layer"):
        from . import bands as _bands
        self.cfg = _capi.mk_config(engine, sigma_mode, use_gamma_ray_opacity, pair_cascade_switch, ejecta_albedo,
                                   pair_cascade_fraction, neutron_precursor_switch, vmax, cosmology, resolution,
                                   magnetar_heating)
        if engine == _capi.MN_EVOLVING:
            self.ROWS = _capi.MK_ROWS_EVOLVING
        self._setup(time, None, y, sigma, device)
        lam = np.geomspace(100, 60000, 100) if lambda_array is None else np.asarray(lambda_array, dtype=np.float64)
        self.phot, self._keep = _bands.build_photometry(bands, self.E, lam, output)

    def _launch(self, *args):
        return self._launch_mag(_capi.lib().rb_mk_mag_eval_f64, *args)

    SPECTRA_TIME_SCALE = 86400.0        # time_observer_frame in seconds (magnetar_driven_ejecta_models.py:212, :232)

    def _spectra_grid_points(self):
        return _optimal_grid_points(1e-4, 1e7, int(self.cfg.resolution))

    def _host_rows_fn(self):
        return None     # the magnitude branch has device entry points only


class MergernovaMagPath(MergernovaPath):
    """Magnitude / bandpass-flux branch of the mergernova models (_processing_other_formats,
    magnetar_driven_ejecta_models.py:198-238)."""

    NEEDS_FREQUENCY = False

    def __init__(self, engine, time, bands, output="magnitude", lambda_array=None, y=None, sigma=None,
                 sigma_mode=_capi.SIGMA_FIXED, use_gamma_ray_opacity=False, pair_cascade_switch=False,
                 ejecta_albedo=0.5, pair_cascade_fraction=0.05, cosmology=None, device=None, supernova=None):
        from . import bands as _bands
        if supernova is None:
            self.cfg = _capi.mn_config(engine, sigma_mode, use_gamma_ray_opacity, pair_cascade_switch, ejecta_albedo,
                                       pair_cascade_fraction, cosmology)
        else:   # general_magnetar_driven_supernova (supernova_models.py:2585-2637): lambda default :2587
            self.cfg = _capi.mn_config(engine, sigma_mode, use_gamma_ray_opacity, pair_cascade_switch, ejecta_albedo,
                                       pair_cascade_fraction, cosmology, _capi.MN_OUT_SUPERNOVA, 2.0, supernova)
            self.ROWS = _capi.MN_ROWS_SUPERNOVA
        if engine == _capi.MN_EVOLVING:
            self.ROWS = _capi.MN_ROWS_EVOLVING
        self._setup(time, None, y, sigma, device)
        default_lam = np.geomspace(100, 60000, 100) if supernova is None else np.geomspace(500, 60000, 200)
        lam = default_lam if lambda_array is None else np.asarray(lambda_array, dtype=np.float64)
        self.phot, self._keep = _bands.build_photometry(bands, self.E, lam, output)

    def _launch(self, *args):
        return self._launch_mag(_capi.lib().rb_mn_mag_eval_f64, *args)

    SPECTRA_TIME_SCALE = 86400.0        # time_observer_frame in seconds (magnetar_driven_ejecta_models.py:212, :232)

    def _spectra_grid_points(self):
        return _optimal_grid_points(self.cfg.grid_t_min, self.cfg.grid_t_max, int(self.cfg.resolution))

    def _host_rows_fn(self):
        return None     # the magnitude branch has device entry points only


def _optimal_grid_points(t_min, t_max, resolution):
    """Length of get_optimal_time_array(t_min, t_max, resolution) without user times (host logic of the library)."""
    buf = (C.c_double * int(resolution))()
    count = C.c_int(0)
    _capi.check(_capi.lib().rb_optimal_time_array(float(t_min), float(t_max), int(resolution), buf, C.byref(count)))
    return int(count.value)


def luminosity_distance(redshift, cosmology=None, device=None):
    """Planck18 (default) luminosity distance in cm for an array of redshifts, on device."""
    dev = _device(device)
    z = as_device_f64(np.atleast_1d(redshift) if not isinstance(redshift, torch.Tensor) else redshift, dev)
    out = torch.empty_like(z)
    cosmo = cosmology
    if cosmo is None:
        cosmo = _capi.Cosmology()
        _capi.check(_capi.lib().rb_cosmology_planck18(C.byref(cosmo)))
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().rb_luminosity_distance_f64(
            C.byref(cosmo), z.numel(), _ptr(z), _ptr(out), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def gaussian_logl(model, y, sigma, sigma_param=None, sigma_mode=_capi.SIGMA_FIXED, ul_level=None,
                  ul_magnitude=False):
    """Stand-alone reduction over a device model matrix [N, E] (likelihoods.py:221-231, 767-790); epochs with
    ul_level[e] > 0 are upper limits (likelihoods.py:377-419)."""
    dev = model.device
    N, E = model.shape
    y = as_device_f64(y, dev)
    if sigma is not None and not isinstance(sigma, torch.Tensor) and np.ndim(sigma) == 0:
        sigma = np.full(E, float(sigma))
    sigma = as_device_f64(sigma, dev) if sigma is not None else None
    sp = as_device_f64(sigma_param, dev) if sigma_param is not None else None
    out = torch.empty((N,), dtype=_F64, device=dev)
    ul = None
    if ul_level is not None:
        level = as_device_f64(ul_level, dev)
        if level.shape != (E,):
            raise ValueError("upper-limit sigma levels must have the same length as y")
        ul = _capi.UpperLimits(level.data_ptr(), int(bool(ul_magnitude)))
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().rb_gaussian_logl_f64(
            sigma_mode, N, E, _ptr(model.contiguous()), _ptr(y), _ptr(sigma), _ptr(sp),
            C.byref(ul) if ul is not None else None, _ptr(out),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out
