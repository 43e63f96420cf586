"""Batched drop-ins for redback.likelihoods.GaussianLikelihood / GaussianLikelihoodQuadratureNoise.

Same constructor, `.parameters`, `.log_likelihood(parameters=None)`, `.noise_log_likelihood()` as
redback/likelihoods.py:22-231 and :738-790, plus `log_likelihood_batch(params_batch) -> ndarray[N]`.
bilby is not a dependency: the two things the reference takes from `bilby.Likelihood` (the `parameters`
dict inferred from the model signature and `log_likelihood_ratio`) are provided here, so instances can be
handed to bilby samplers (which only call those methods) when bilby is installed.
"""
import inspect

import numpy as np
import torch

from . import _capi, _fused
from . import afterglow_models as _ag
from . import kilonova_models as _kn
from . import magnetar_driven_ejecta_models as _mn
from . import phase_models as _ph
from . import supernova_models as _sn
from . import tde_models as _tde
from .engine import batch_size, gaussian_logl

FUSED = {}
FUSED.update(_sn.FUSED)
FUSED.update(_kn.FUSED)
FUSED.update(_ag.FUSED)
FUSED.update(_mn.FUSED)
FUSED.update(_ph.FUSED)
FUSED.update(_tde.FUSED)


def infer_parameters_from_function(func):
    """bilby.core.utils.introspection.infer_parameters_from_function: all positional-or-keyword argument
    names after the first one; *args / **kwargs are ignored (likelihoods.py:50)."""
    sig = inspect.signature(func)
    names = [n for n, p in sig.parameters.items()
             if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    return [n for n in names[1:] if n != "params_batch"]


try:    # the reference's base class (likelihoods.py:22): subclass it whenever bilby is importable, so isinstance checks
    # in bilby's samplers / result classes hold; without bilby the two inherited members are provided below
    import bilby as _bilby
    _LikelihoodBase = _bilby.Likelihood
except ImportError:
    _bilby = None
    _LikelihoodBase = object


class _RedbackLikelihood(_LikelihoodBase):
    """redback/likelihoods.py:22-143"""

    def __init__(self, x, y, function, kwargs=None, priors=None, fiducial_parameters=None):
        self.x = x
        self.y = y
        self.function = function
        self.kwargs = kwargs
        self.priors = priors
        self.fiducial_parameters = fiducial_parameters
        parameters = dict.fromkeys(infer_parameters_from_function(function))
        if _bilby is not None:
            super().__init__(parameters=parameters)
        else:
            self.parameters = parameters
            self._marginalized_parameters = []

    def _update_parameters(self, parameters=None):
        if parameters is not None:
            self.parameters.update(parameters)

    @property
    def kwargs(self):
        return self._kwargs

    @kwargs.setter
    def kwargs(self, kwargs):
        self._kwargs = dict() if kwargs is None else kwargs

    @property
    def n(self):
        return len(self.x)

    def log_likelihood_ratio(self):
        return self.log_likelihood() - self.noise_log_likelihood()


class GaussianLikelihood(_RedbackLikelihood):
    """redback/likelihoods.py:146-231 with a vectorised `log_likelihood_batch`."""

    _sigma_mode_sampled = _capi.SIGMA_SAMPLED

    def __init__(self, x, y, sigma, function, kwargs=None, priors=None, fiducial_parameters=None):
        self._noise_log_likelihood = None
        self._path = None
        super().__init__(x=x, y=y, function=function, kwargs=kwargs, priors=priors,
                         fiducial_parameters=fiducial_parameters)
        self.sigma = sigma
        if self.sigma is None:
            self.parameters['sigma'] = None

    @property
    def sigma(self):
        return self.parameters.get('sigma', self._sigma)

    @sigma.setter
    def sigma(self, sigma):
        if sigma is None:
            self._sigma = sigma
        elif isinstance(sigma, float) or isinstance(sigma, int):
            self._sigma = sigma
        elif len(sigma) == self.n:
            self._sigma = sigma
        elif sigma.shape == ((2, len(self.x))):
            self._sigma = sigma
        else:
            raise ValueError('Sigma must be either float or array-like x.')

    @property
    def model_output(self):
        return self.function(self.x, **self.parameters, **self.kwargs)

    @property
    def residual(self):
        return self.y - self.model_output

    # ---- single-sample API (what bilby samplers call) -------------------------------------------
    def noise_log_likelihood(self):
        if self._noise_log_likelihood is None:
            self._noise_log_likelihood = self._gaussian_log_likelihood(res=self.y, sigma=self._noise_sigma())
        return self._noise_log_likelihood

    def _noise_sigma(self):
        return self.sigma

    def log_likelihood(self, parameters=None):
        self._update_parameters(parameters)
        if self.function in FUSED:
            batch = {k: np.array([v], dtype=np.float64) for k, v in self.parameters.items() if v is not None}
            return float(self.log_likelihood_batch(batch)[0])
        return float(np.nan_to_num(self._gaussian_log_likelihood(res=self.residual, sigma=self._full_sigma())))

    def _full_sigma(self):
        return self.sigma

    def _gaussian_log_likelihood(self, res, sigma):
        """likelihoods.py:229-231 -- evaluated by the device reduction kernel (one 'sample')."""
        res = np.asarray(res, dtype=np.float64)
        model = torch.zeros((1, res.size), dtype=torch.float64, device=self._dev())
        sig = np.broadcast_to(np.asarray(sigma, dtype=np.float64), res.shape)
        out = gaussian_logl(model, res, sig, sigma_mode=_capi.SIGMA_FIXED)
        # the kernel applies nan_to_num; the reference does not inside this helper but every caller does
        return float(out[0].item())

    def _dev(self):
        return torch.device(self.kwargs.get("device") or f"cuda:{torch.cuda.current_device()}") \
            if torch.cuda.is_available() else _raise_no_cuda()

    # ---- batched API ----------------------------------------------------------------------------
    def _sigma_mode(self, batch):
        if self._sigma is None:
            return _capi.SIGMA_SAMPLED
        return _capi.SIGMA_FIXED

    def _split_sigma_param(self, batch, N):
        """Return the per-sample `sigma` parameter array (or None) for the chosen noise mode."""
        mode = self._sigma_mode(batch)
        if mode in (_capi.SIGMA_FIXED, _capi.SIGMA_FRACTIONAL):
            return mode, None
        sp = batch.get('sigma', self.parameters.get('sigma'))
        if sp is None:
            raise KeyError('sigma')
        return mode, sp

    def log_likelihood_batch(self, params_batch, device_output=False):
        """np.nan_to_num(log L) for every sample of `params_batch` ({name: array[N]}); parameters not in the
        batch are taken from `self.parameters` / `self.kwargs`.  Shape (N,)."""
        batch = dict(params_batch)
        merged = {k: v for k, v in self.parameters.items() if v is not None}
        merged.update(batch)
        N = batch_size(merged)
        if N is None:
            raise ValueError("params_batch must contain at least one array")
        mode, sp = self._split_sigma_param(merged, N)
        if self.function in FUSED:
            logl = self._fused(merged, N, mode, sp, host=not device_output)
            if isinstance(logl, np.ndarray):
                return logl
        else:
            kw = dict(self.kwargs)
            kw["device_output"] = True
            fn_params = {k: v for k, v in merged.items() if k != 'sigma' or 'sigma' in
                         inspect.signature(self.function).parameters}
            try:
                model = self.function(self.x, params_batch=fn_params, **kw)
            except TypeError as exc:
                raise NotImplementedError(
                    "log_likelihood_batch needs a model function with a params_batch entry point") from exc
            if not isinstance(model, torch.Tensor):
                model = torch.as_tensor(np.asarray(model, dtype=np.float64), device=self._dev())
            sp_dev = None
            if sp is not None:
                sp_dev = sp if isinstance(sp, torch.Tensor) else np.broadcast_to(np.asarray(sp, dtype=np.float64), (N,))
            level, ul_mag = self._upper_limit_levels()
            logl = gaussian_logl(model, self.y, self._sigma, sp_dev, sigma_mode=mode, ul_level=level,
                                 ul_magnitude=ul_mag)
        return logl if device_output else logl.cpu().numpy()

    def _upper_limit_levels(self):
        """(per-epoch sigma level with 0 = detection, magnitude mode) or (None, False): no upper limits."""
        return None, False

    def _fused(self, merged, N, mode, sp, host=False):
        spec = FUSED[self.function]
        spec = spec.for_kwargs(self.kwargs)      # wrappers / kwargs-selected variants (base_model, interaction_process)
        kw = dict(self.kwargs)
        if self.function is _kn.one_component_kilonova_model:
            kw.setdefault("temperature_floor", 4000)
        # the path holds device copies of x / y / sigma: key it on their content so that editing the likelihood's
        # attributes (assignment or in place) is honoured on the next call, as in the reference (likelihoods.py:206-231)
        key = (spec.cache_key(kw), mode, _fused.content_key(self.x), _fused.content_key(self.y),
               _fused.content_key(self._sigma), _fused.content_key(self._upper_limit_levels()[0]))
        if self._path is None or self._path[0] != key:
            self._path = (key, spec.build(self.x, kw, y=self.y, sigma=self._sigma, sigma_mode=mode))
            level, ul_mag = self._upper_limit_levels()
            if level is not None:
                self._path[1].set_upper_limits(level, ul_mag)
        path = self._path[1]
        params, _ = _fused.collect(merged, kw, None, spec.needed, spec.defaults)
        params = spec.transform(params)
        if host and path.supports_host_rows() and not any(
                isinstance(v, torch.Tensor) for v in list(params.values()) + [sp, kw.get("dl")]):
            # host-resident batch: the library's chunked, double-buffered H2D -> kernels -> D2H pipeline reads the
            # caller's arrays where they lie (rb_*_eval_f64_host_rows)
            _, logl = path.evaluate_host(params, N, dl=kw.get("dl"), sigma_param=sp, t0=params.get("t0"),
                                         aux=params.get(getattr(path, "AUX_PARAM", None) or ""), want_model=False,
                                         want_logl=True)
            return logl
        dev = path.pack(params, N)
        sp_dev = None
        if sp is not None:
            if isinstance(sp, torch.Tensor):
                sp_dev = sp.to(device=path.device, dtype=torch.float64).reshape(-1).expand(N).contiguous()
            else:
                sp_dev = torch.from_numpy(np.array(np.broadcast_to(np.asarray(sp, dtype=np.float64), (N,)))).to(path.device)
        _, logl = path.evaluate(dev, dl=_fused.dl_from_kwargs(kw, path, N), sigma_param=sp_dev, want_model=False,
                                want_logl=True, t0=_fused.t0_tensor(params, path, N),
                                aux=_fused.aux_tensor(params, path, N))
        return logl


def _raise_no_cuda():
    raise _capi.RedbackB200Error("redback_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


class GaussianLikelihoodWithUpperLimits(GaussianLikelihood):
    """redback/likelihoods.py:234-489: non-detections contribute log Phi((limit - model)/sigma_meas) (flux-like
    data) or the survival function (magnitudes); detections keep the Gaussian term.  The per-epoch split is done
    inside the fused kernels (an extra epoch-table column), so a batch costs the same as without upper limits."""

    def __init__(self, x, y, sigma, function, kwargs=None, priors=None, fiducial_parameters=None, detections=None,
                 upper_limit_sigma=3.0, data_mode='flux'):
        super().__init__(x=x, y=y, sigma=sigma, function=function, kwargs=kwargs, priors=priors,
                         fiducial_parameters=fiducial_parameters)
        self.detections = detections
        self.upper_limit_sigma = upper_limit_sigma
        self.data_mode = data_mode
        if not np.all(self._detections):
            nan_ul = np.isnan(np.asarray(self.y, dtype=np.float64)[~self._detections])
            if np.any(nan_ul):
                raise ValueError(
                    f"{int(np.sum(nan_ul))} upper limit(s) have NaN y-values. Upper limits require a finite "
                    f"value (e.g. the limiting magnitude or flux) to compute likelihood. "
                    f"Replace NaN values with the upper limit value, or remove those data points.")

    @property
    def detections(self):
        return self._detections

    @detections.setter
    def detections(self, detections):
        if detections is None:
            self._detections = np.ones(len(self.x), dtype=bool)
        elif len(detections) == len(self.x):
            self._detections = np.array(detections, dtype=bool)
        else:
            raise ValueError('detections must have the same length as x.')
        self._path = None

    @property
    def upper_limits(self):
        return ~self._detections

    @property
    def upper_limit_sigma(self):
        return self._upper_limit_sigma

    @upper_limit_sigma.setter
    def upper_limit_sigma(self, upper_limit_sigma):
        if isinstance(upper_limit_sigma, (float, int)):
            self._upper_limit_sigma = float(upper_limit_sigma)
        elif isinstance(upper_limit_sigma, np.ndarray):
            if len(upper_limit_sigma) == len(self.x) or \
                    (hasattr(self, '_detections') and len(upper_limit_sigma) == np.sum(~self._detections)):
                self._upper_limit_sigma = upper_limit_sigma
            else:
                raise ValueError('upper_limit_sigma array must have length equal to x or to number of upper limits.')
        else:
            raise ValueError('upper_limit_sigma must be a float or array.')
        self._path = None

    @property
    def data_mode(self):
        return self._data_mode

    @data_mode.setter
    def data_mode(self, data_mode):
        if data_mode.lower() not in ['flux', 'flux_density', 'luminosity', 'magnitude', 'mag']:
            raise ValueError("data_mode must be 'flux', 'flux_density', 'luminosity', or 'magnitude' (or 'mag')")
        self._data_mode = 'magnitude' if data_mode.lower() == 'mag' else data_mode.lower()
        self._path = None

    def get_upper_limit_sigma_values(self):
        if not np.any(self.upper_limits):
            return np.array([])
        if isinstance(self.upper_limit_sigma, (float, int)):
            return np.full(np.sum(self.upper_limits), self.upper_limit_sigma)
        if len(self.upper_limit_sigma) == len(self.x):
            return self.upper_limit_sigma[self.upper_limits]
        return self.upper_limit_sigma

    def _upper_limit_levels(self):
        if not np.any(self.upper_limits):
            return None, False
        level = np.zeros(len(self.x), dtype=np.float64)
        level[self.upper_limits] = self.get_upper_limit_sigma_values()
        if np.any(~(level[self.upper_limits] > 0)):
            raise ValueError('upper_limit_sigma must be positive for every upper limit.')
        return level, self.data_mode == 'magnitude'

    def _device_logl(self, model_row):
        """detections + upper limits for one model vector through the device reduction (likelihoods.py:421-468)."""
        model = torch.as_tensor(np.asarray(model_row, dtype=np.float64).reshape(1, -1), device=self._dev())
        level, ul_mag = self._upper_limit_levels()
        sig = self.sigma
        if sig is None:
            raise KeyError('sigma')
        sig = np.array(np.broadcast_to(np.asarray(sig, dtype=np.float64), (len(self.x),)))
        sig[self.upper_limits] = 1.0    # never read for upper limits; keeps NaN placeholders out of the table
        out = gaussian_logl(model, self.y, sig, sigma_mode=_capi.SIGMA_FIXED, ul_level=level, ul_magnitude=ul_mag)
        return float(out[0].item())

    def noise_log_likelihood(self):
        if self._noise_log_likelihood is None:
            fill = 60.0 if self.data_mode == 'magnitude' else 0.0
            # model = 0 at detections makes the residual y itself (:432-435); 0 / 60 mag at the limits (:440-445)
            noise_model = np.where(self.detections, 0.0, fill)
            self._noise_log_likelihood = self._device_logl(noise_model)
        return self._noise_log_likelihood

    def log_likelihood(self, parameters=None):
        self._update_parameters(parameters)
        if self.function in FUSED:
            batch = {k: np.array([v], dtype=np.float64) for k, v in self.parameters.items() if v is not None}
            return float(self.log_likelihood_batch(batch)[0])
        return self._device_logl(self.model_output)

    def summary(self):
        n_upper_limits = np.sum(self.upper_limits)
        return {'total_data_points': len(self.x), 'detections': np.sum(self.detections),
                'upper_limits': n_upper_limits, 'data_mode': self.data_mode,
                'upper_limit_sigma_levels': self.get_upper_limit_sigma_values() if n_upper_limits > 0 else None}


class MixtureGaussianLikelihood(GaussianLikelihood):
    """redback/likelihoods.py:492-665: every datum is alpha * N(r | sigma) + (1 - alpha) * N(r | sigma_out); `alpha`
    and `sigma_out` are parameters (sampled or fixed).  The batch path evaluates the model matrix on device in
    chunks and reduces it with the mixture kernel (the two extra per-sample parameters are not part of the fused
    kernels' single noise-parameter slot)."""

    _CHUNK_BYTES = 1 << 28

    def __init__(self, x, y, sigma, function, kwargs=None, priors=None, fiducial_parameters=None):
        super().__init__(x=x, y=y, sigma=sigma, function=function, kwargs=kwargs, priors=priors,
                         fiducial_parameters=fiducial_parameters)
        if 'alpha' not in self.parameters:
            self.parameters['alpha'] = 0.9
        if 'sigma_out' not in self.parameters:
            self.parameters['sigma_out'] = sigma * 10 if sigma is not None and isinstance(sigma, (int, float)) \
                else 10.0
        self._noise_log_likelihood = None

    def _mixture(self, model, alpha, sigma_out):
        """model: device tensor [N, E]; alpha, sigma_out: [N] arrays -> device tensor [N]."""
        import ctypes as C
        if self._sigma is None:
            raise NotImplementedError("MixtureGaussianLikelihood with a sampled inlier sigma is not available on device")
        dev = model.device
        N, E = model.shape
        from .engine import as_device_f64, _ptr
        y = as_device_f64(self.y, dev)
        sig = as_device_f64(np.broadcast_to(np.asarray(self._sigma, dtype=np.float64), (E,)), dev)
        al = as_device_f64(np.broadcast_to(np.asarray(alpha, dtype=np.float64), (N,)), dev)
        so = as_device_f64(np.broadcast_to(np.asarray(sigma_out, dtype=np.float64), (N,)), dev)
        out = torch.empty((N,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().rb_mixture_logl_f64(
                N, E, _ptr(model.contiguous()), _ptr(y), _ptr(sig), _ptr(al), _ptr(so), _ptr(out),
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return out

    def log_likelihood(self, parameters=None):
        self._update_parameters(parameters)
        fn_params = {k: v for k, v in self.parameters.items() if k not in ('alpha', 'sigma_out', 'sigma')
                     and v is not None}
        model = np.asarray(self.function(self.x, **fn_params, **self.kwargs), dtype=np.float64)
        dev_model = torch.as_tensor(model.reshape(1, -1), device=self._dev())
        return float(self._mixture(dev_model, self.parameters['alpha'], self.parameters['sigma_out'])[0].item())

    def log_likelihood_batch(self, params_batch, device_output=False):
        merged = {k: v for k, v in self.parameters.items() if v is not None}
        merged.update(params_batch)
        N = batch_size(merged)
        if N is None:
            raise ValueError("params_batch must contain at least one array")
        alpha = np.broadcast_to(np.asarray(merged.pop('alpha'), dtype=np.float64), (N,))
        sigma_out = np.broadcast_to(np.asarray(merged.pop('sigma_out'), dtype=np.float64), (N,))
        merged.pop('sigma', None)
        kw = dict(self.kwargs)
        kw["device_output"] = True
        step = max(1, self._CHUNK_BYTES // (8 * len(self.x)))
        parts = []
        for n0 in range(0, N, step):
            sl = slice(n0, min(N, n0 + step))
            chunk = {k: (v[sl] if np.ndim(v) > 0 and len(v) == N else v) for k, v in merged.items()}
            chunk = {k: (np.full(sl.stop - sl.start, float(v)) if np.ndim(v) == 0 else v) for k, v in chunk.items()}
            try:
                model = self.function(self.x, params_batch=chunk, **kw)
            except TypeError as exc:
                raise NotImplementedError(
                    "log_likelihood_batch needs a model function with a params_batch entry point") from exc
            if not isinstance(model, torch.Tensor):
                model = torch.as_tensor(np.asarray(model, dtype=np.float64), device=self._dev())
            parts.append(self._mixture(model, alpha[sl], sigma_out[sl]))
        logl = torch.cat(parts)
        return logl if device_output else logl.cpu().numpy()

    def _density(self, r, sigma):
        dev = self._dev()
        r = torch.as_tensor(np.asarray(r, dtype=np.float64), device=dev)
        sigma = torch.as_tensor(np.asarray(sigma, dtype=np.float64), device=dev)
        return (1 / (float(np.sqrt(2 * np.pi)) * sigma)) * torch.exp(-0.5 * (r / sigma) ** 2)

    def p_in(self, r):
        """likelihoods.py:579-593"""
        return self._density(r, self.sigma).cpu().numpy()

    def p_out(self, r):
        """likelihoods.py:595-609"""
        return self._density(r, self.parameters.get('sigma_out')).cpu().numpy()

    def calculate_outlier_posteriors(self, model_prediction):
        """likelihoods.py:637-665 (a per-fit diagnostic on one model vector; elementwise on device)."""
        r = np.asarray(self.y, dtype=np.float64) - np.asarray(model_prediction, dtype=np.float64)
        alpha = float(self.parameters.get('alpha'))
        pin = self._density(r, self.sigma)
        pout = self._density(r, self.parameters.get('sigma_out'))
        numerator = (1 - alpha) * pout
        denominator = alpha * pin + (1 - alpha) * pout
        return torch.where(denominator > 0, numerator / denominator, torch.zeros_like(denominator)).cpu().numpy()


class GaussianLikelihoodQuadratureNoise(GaussianLikelihood):
    """redback/likelihoods.py:738-790: sigma_full = sqrt(sigma_i^2 + sigma^2), `sigma` sampled."""

    def __init__(self, x, y, sigma_i, function, kwargs=None, priors=None, fiducial_parameters=None):
        self.sigma_i = sigma_i
        super().__init__(x=x, y=y, sigma=sigma_i, function=function, kwargs=kwargs, priors=priors,
                         fiducial_parameters=fiducial_parameters)

    @property
    def full_sigma(self):
        return np.sqrt(self.sigma_i ** 2. + self.sigma ** 2.)

    def _noise_sigma(self):
        return self.full_sigma

    def _full_sigma(self):
        return self.full_sigma

    def _sigma_mode(self, batch):
        return _capi.SIGMA_QUADRATURE


class GaussianLikelihoodWithFractionalNoise(GaussianLikelihood):
    """redback/likelihoods.py:792-851: sigma = sqrt(sigma_i^2 * model^2) (noise proportional to the model)."""

    def __init__(self, x, y, sigma_i, function, kwargs=None, priors=None, fiducial_parameters=None):
        self.sigma_i = sigma_i
        super().__init__(x=x, y=y, sigma=sigma_i, function=function, kwargs=kwargs, priors=priors,
                         fiducial_parameters=fiducial_parameters)

    @property
    def full_sigma(self):
        model_y = self.model_output
        return np.sqrt(self.sigma_i ** 2. * model_y ** 2)

    def _noise_sigma(self):
        return self.sigma_i

    def _full_sigma(self):
        return self.full_sigma

    def _sigma_mode(self, batch):
        return _capi.SIGMA_FRACTIONAL


class GaussianLikelihoodWithSystematicNoise(GaussianLikelihood):
    """redback/likelihoods.py:853-913: sigma = sqrt(sigma_i^2 + model^2 * sigma^2), `sigma` sampled."""

    def __init__(self, x, y, sigma_i, function, kwargs=None, priors=None, fiducial_parameters=None):
        self.sigma_i = sigma_i
        super().__init__(x=x, y=y, sigma=sigma_i, function=function, kwargs=kwargs, priors=priors,
                         fiducial_parameters=fiducial_parameters)

    @property
    def full_sigma(self):
        model_y = self.model_output
        return np.sqrt(self.sigma_i ** 2. + model_y ** 2 * self.sigma ** 2.)

    def _noise_sigma(self):
        return self.sigma_i

    def _full_sigma(self):
        return self.full_sigma

    def _sigma_mode(self, batch):
        return _capi.SIGMA_SYSTEMATIC
