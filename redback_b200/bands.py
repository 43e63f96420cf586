"""Host-side bandpass tables for the magnitude branch.

In the reference this is sncosmo's job (`get_bandpass`, `Bandpass.__call__`, `utils.integration_grid`,
`get_magsystem('ab').zpbandflux`; call sites redback/sed.py:20-37, 48-49, 101-113).  sncosmo is an optional
import here: if it is installed, unknown band names are fetched from its registry; otherwise bandpasses must be
registered with `register_bandpass(name, wave, trans)`.  Only small per-call tables are produced on the host; the
spline fit, the band integrals and the magnitudes run on the device (csrc/phot_kernel.cu).
"""
import ctypes as C
import math

import numpy as np

HC_ERG_AA = 1.9864458571489284e-08   # redback/sed.py:19
MODEL_BANDFLUX_SPACING = 5.0         # redback/sed.py:20
_C_AA_PER_S = 2.99792458e18
_PLANCK = 6.62607015e-27

# redback/tables/filters.csv:54-59, 65-67, column reference_flux (used by output_format='flux', utils.py:707-718)
REFERENCE_FLUX = dict(lsstu=3.822e-06, lsstg=6.074e-06, lsstr=3.433e-06, lssti=2.269e-06, lsstz=1.445e-06,
                      lssty=1.005e-06, ztfg=5.785e-06, ztfr=3.796e-06, ztfi=2.340e-06)


class Bandpass:
    """Tabulated transmission; linear interpolation, zero outside [minwave, maxwave] (sncosmo.Bandpass)."""

    def __init__(self, name, wave, trans, reference_flux=None):
        self.name = name
        self.wave = np.ascontiguousarray(wave, dtype=np.float64)
        self.trans = np.ascontiguousarray(trans, dtype=np.float64)
        if self.wave.ndim != 1 or self.wave.shape != self.trans.shape or self.wave.size < 2:
            raise ValueError("wave and trans must be 1-D arrays of the same length")
        if np.any(np.diff(self.wave) <= 0):
            raise ValueError("bandpass wavelengths must be strictly increasing")
        self.reference_flux = REFERENCE_FLUX.get(name) if reference_flux is None else float(reference_flux)

    def minwave(self):
        return float(self.wave[0])

    def maxwave(self):
        return float(self.wave[-1])

    def __call__(self, wave):
        return np.interp(wave, self.wave, self.trans, left=0.0, right=0.0)

    def integration_grid(self):
        """sncosmo.utils.integration_grid(minwave, maxwave, 5 A)"""
        low, high = self.minwave(), self.maxwave()
        range_diff = high - low
        spacing = range_diff / int(math.ceil(range_diff / MODEL_BANDFLUX_SPACING))
        return np.arange(low + spacing / 2., high, spacing), spacing

    def zpbandflux_ab(self):
        """sncosmo ABMagSystem.zpbandflux: sum(3631e-23 / h / nu * trans * gradient(nu))"""
        nu = _C_AA_PER_S / self.wave
        tr = self.trans
        if nu[0] > nu[-1]:
            nu, tr = nu[::-1], tr[::-1]
        return float(np.sum(3631.e-23 / _PLANCK / nu * tr * np.gradient(nu)))


_REGISTRY = {}


def register_bandpass(name, wave, trans, reference_flux=None):
    _REGISTRY[name] = Bandpass(name, wave, trans, reference_flux)
    return _REGISTRY[name]


def get_bandpass(name):
    if name in _REGISTRY:
        return _REGISTRY[name]
    try:
        import sncosmo  # optional: tables only
    except ImportError:
        raise KeyError(f"Band {name} is not registered (sncosmo is not installed: use "
                       f"redback_b200.bands.register_bandpass(name, wave, trans))") from None
    b = sncosmo.get_bandpass(name)
    return register_bandpass(name, b.wave, b.trans)


def synthetic_bandpass(name, lo, hi, edge=25.0, peak=0.8, step=10.0):
    """Smoothed top-hat between `lo` and `hi` Angstrom -- a documented stand-in when the survey's real
    transmission curves are not available (no network / no sncosmo data)."""
    w = np.arange(lo - 8 * edge, hi + 8 * edge + step / 2, step)
    tr = 0.5 * peak * (np.tanh((w - lo) / edge) - np.tanh((w - hi) / edge))
    tr[0] = tr[-1] = 0.0
    return register_bandpass(name, w, tr)


# published LSST band edges (nm): u 320-400, g 400-552, r 552-691, i 691-818, z 818-922, y 922-1060
SYNTHETIC_LSST_EDGES = dict(lsstu=(3200, 4000), lsstg=(4000, 5520), lsstr=(5520, 6910), lssti=(6910, 8180),
                            lsstz=(8180, 9220), lssty=(9220, 10600))


def register_synthetic_lsst():
    return {n: synthetic_bandpass(n, lo, hi) for n, (lo, hi) in SYNTHETIC_LSST_EDGES.items()}


class Photometry(C.Structure):
    """rb_photometry (include/redback_b200.h); all pointers are host pointers."""
    _fields_ = [("output", C.c_int), ("n_lambda", C.c_int), ("lambda_obs", C.c_void_p), ("n_tgrid", C.c_int),
                ("tgrid", C.c_void_p), ("n_bands", C.c_int), ("band_of_epoch", C.c_void_p), ("band_off", C.c_void_p),
                ("int_wave", C.c_void_p), ("int_wt", C.c_void_p), ("band_scale", C.c_void_p), ("band_zp", C.c_void_p),
                ("band_ref_flux", C.c_void_p), ("extinction", C.c_void_p), ("ext_av_host", C.c_void_p),
                ("spectra_out", C.c_void_p), ("spectra_time_out", C.c_void_p), ("spectra_len_out", C.c_void_p)]


def build_photometry(bands, n_epochs, lambda_array, output, tgrid=None):
    """Return (Photometry struct, keepalive list).  `bands`: band name or array of names (one per epoch); ignored for
    output='spectra' (no band tables: the spectra themselves are the output, engine.evaluate_spectra)."""
    if output == "spectra":
        lam = np.ascontiguousarray(lambda_array, dtype=np.float64)
        ph = Photometry()
        ph.output = 2
        ph.n_lambda, ph.lambda_obs = lam.size, lam.ctypes.data
        keep = [lam]
        ph.n_tgrid, ph.tgrid = 0, None
        if tgrid is not None:
            tg = np.ascontiguousarray(tgrid, dtype=np.float64)
            keep.append(tg)
            ph.n_tgrid, ph.tgrid = tg.size, tg.ctypes.data
        return ph, keep
    names = np.broadcast_to(np.asarray(bands), (n_epochs,))
    uniq = sorted(set(names.tolist()))
    bps = [get_bandpass(n) for n in uniq]
    lam = np.ascontiguousarray(lambda_array, dtype=np.float64)
    index = {n: i for i, n in enumerate(uniq)}
    boe = np.ascontiguousarray([index[n] for n in names.tolist()], dtype=np.int32)
    waves, wts, scale, zp, off = [], [], [], [], [0]
    for b in bps:
        if b.minwave() < lam[0] or b.maxwave() > lam[-1]:   # sed.py:22-28
            raise ValueError('bandpass {0!r:s} [{1:.6g}, .., {2:.6g}] outside spectral range [{3:.6g}, .., {4:.6g}]'
                             .format(b.name, b.minwave(), b.maxwave(), lam[0], lam[-1]))
        w, dw = b.integration_grid()
        waves.append(w)
        wts.append(w * b(w))
        scale.append(dw / HC_ERG_AA)
        zp.append(b.zpbandflux_ab())
        off.append(off[-1] + w.size)
    int_wave = np.ascontiguousarray(np.concatenate(waves))
    int_wt = np.ascontiguousarray(np.concatenate(wts))
    off = np.ascontiguousarray(off, dtype=np.int64)
    scale = np.ascontiguousarray(scale, dtype=np.float64)
    zp = np.ascontiguousarray(zp, dtype=np.float64)
    keep = [lam, boe, int_wave, int_wt, off, scale, zp]
    ph = Photometry()
    ph.output = 1 if output == "flux" else 0
    ph.n_lambda, ph.lambda_obs = lam.size, lam.ctypes.data
    ph.n_tgrid, ph.tgrid = 0, None
    if tgrid is not None:
        tg = np.ascontiguousarray(tgrid, dtype=np.float64)
        keep.append(tg)
        ph.n_tgrid, ph.tgrid = tg.size, tg.ctypes.data
    ph.n_bands = len(bps)
    ph.band_of_epoch, ph.band_off = boe.ctypes.data, off.ctypes.data
    ph.int_wave, ph.int_wt = int_wave.ctypes.data, int_wt.ctypes.data
    ph.band_scale, ph.band_zp = scale.ctypes.data, zp.ctypes.data
    ph.band_ref_flux = None
    ph.extinction = ph.ext_av_host = None
    if output == "flux":
        missing = [b.name for b in bps if b.reference_flux is None]
        if missing:
            raise KeyError(f"Band {missing[0]} is not defined in filters.csv!")   # utils.py:869
        ref = np.ascontiguousarray([b.reference_flux for b in bps], dtype=np.float64)
        keep.append(ref)
        ph.band_ref_flux = ref.ctypes.data
    return ph, keep
