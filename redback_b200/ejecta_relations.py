"""Ejecta relations: gravitational-wave source parameters -> kilonova ejecta mass [Msun] and velocity [c]
(the fits used by redback/ejecta_relations.py: Dietrich & Ujevic 2017 for binary neutron stars, Kawaguchi et al.
2016 for neutron-star black-hole binaries).  Host-side and vectorised over the sample axis: a handful of flops per
sample that feed `params_batch` of the fused kilonova kernels.  Each class exposes `ejecta_mass` and
`ejecta_velocity` like the reference's, evaluated once at construction."""
import numpy as np

# quasi-universal relations (ejecta_relations.py:435-444, 473-482)
_C_OF_LOG_LAMBDA = (0.371, -0.0391, 0.001056)
_MB_COEF, _MB_POWER = 0.8857853174243745, 1.2082383572002926
# (scale, offset, compactness coefficient) of the symmetric mass-ratio fits  s * [q (1 + c C1) + (1 + c C2) / q] + o
_V_TOTAL = (-0.3090, 0.657, -1.879)        # ejecta_relations.py:40-43
_V_RHO = (-0.219479, 0.444836, -2.67385)   # :492-494
_V_Z = (-0.315585, 0.63808, -1.00757)      # :514-516


def _arr(x):
    return np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64)


def calc_compactness_from_lambda(lambda_1):
    u = np.log(_arr(lambda_1))
    k0, k1, k2 = _C_OF_LOG_LAMBDA
    return k0 + k1 * u + k2 * u ** 2


def calc_baryonic_mass(mass, compactness):
    return _arr(mass) * (1 + _MB_COEF * _arr(compactness) ** _MB_POWER)


def _symmetric_fit(coef, ma, mb, ca, cb):
    scale, offset, c = coef
    return ((ma / mb) * (1.0 + c * ca) + (mb / ma) * (1.0 + c * cb)) * scale + offset


def calc_vrho(mass_1, mass_2, lambda_1, lambda_2):
    """Mean ejecta velocity in the orbital plane."""
    return _symmetric_fit(_V_RHO, _arr(mass_1), _arr(mass_2), calc_compactness_from_lambda(lambda_1),
                          calc_compactness_from_lambda(lambda_2))


def calc_vz(mass_1, mass_2, lambda_1, lambda_2):
    """Mean ejecta velocity orthogonal to the orbital plane."""
    return _symmetric_fit(_V_Z, _arr(mass_1), _arr(mass_2), calc_compactness_from_lambda(lambda_1),
                          calc_compactness_from_lambda(lambda_2))


class _BinaryNeutronStar:
    def __init__(self, mass_1, mass_2, lambda_1, lambda_2):
        self.mass_1, self.mass_2 = _arr(mass_1), _arr(mass_2)
        self.lambda_1, self.lambda_2 = _arr(lambda_1), _arr(lambda_2)
        self.c1 = calc_compactness_from_lambda(self.lambda_1)
        self.c2 = calc_compactness_from_lambda(self.lambda_2)
        self.vrho = _symmetric_fit(_V_RHO, self.mass_1, self.mass_2, self.c1, self.c2)
        self.vz = _symmetric_fit(_V_Z, self.mass_1, self.mass_2, self.c1, self.c2)
        self.ejecta_velocity = self._velocity()
        self.ejecta_mass = self._mass()


class OneComponentBNSNoProjection(_BinaryNeutronStar):
    """One component, no projection (ejecta_relations.py:5-65)."""

    def _velocity(self):
        scale, offset, c = _V_TOTAL      # the reference adds the two terms and the offset left to right (:43)
        m1, m2 = self.mass_1, self.mass_2
        return scale * (m1 / m2) * (1 + c * self.c1) + scale * (m2 / m1) * (1 + c * self.c2) + offset

    def _mass(self):
        m1, m2, c1, c2 = self.mass_1, self.mass_2, self.c1, self.c2
        tidal = m1 * (1 - 2 * c1) / c1 + m2 * (1 - 2 * c2) / c2
        ratio = m1 * (m2 / m1) ** -2.905 + m2 * (m1 / m2) ** -2.905
        return 10 ** (-0.0719 * tidal + 0.2116 * ratio + -2.42)                 # :56-64, log10 of the mass


class OneComponentBNSProjection(_BinaryNeutronStar):
    """One component with the orbital-plane / orthogonal velocity projection (ejecta_relations.py:88-147)."""

    def _velocity(self):
        return (self.vrho ** 2.0 + self.vz ** 2.0) ** 0.5

    def _mass(self):
        m1, m2, c1, c2 = self.mass_1, self.mass_2, self.c1, self.c2
        b1, b2 = calc_baryonic_mass(m1, c1), calc_baryonic_mass(m2, c2)
        tidal = (b1 * ((m2 / m1) ** (1.0 / 3.0)) * (1.0 - 2.0 * c1) / c1) + (b2 * ((m1 / m2) ** (1.0 / 3.0)) * (1.0 - 2.0 * c2) / c2)
        ratio = b1 * ((m2 / m1) ** -2.5484) + b2 * ((m1 / m2) ** -2.5484)
        binding = b1 * (1.0 - m1 / b1) + b2 * (1.0 - m2 / b2)
        milli_msun = tidal * -1.35695 + ratio * 6.11252 + binding * -49.43355 + 16.1144               # :130-146
        return np.maximum(milli_msun, 0) / 1000.0


class TwoComponentBNS(OneComponentBNSNoProjection):
    """Dynamical + disk-wind ejecta of a binary neutron star merger (ejecta_relations.py:170-267, Coughlin et al. 2019):
    the dynamical component uses the one-component fits, the disk mass follows from the threshold mass for prompt
    collapse; `zeta` is the fraction of the disk that is unbound."""

    def __init__(self, mass_1, mass_2, lambda_1, lambda_2, mtov, zeta):
        super().__init__(mass_1, mass_2, lambda_1, lambda_2)
        self.mtov, self.zeta = _arr(mtov), _arr(zeta)
        m1, m2, l1, l2 = self.mass_1, self.mass_2, self.lambda_1, self.lambda_2
        q = m1 / m2
        lambda_tilde = (16.0 / 13.0) * (l2 + l1 * (q ** 5) + 12 * l1 * (q ** 4) + 12 * l2 * q) / ((q + 1) ** 5)
        chirp = ((m1 * m2) ** (3. / 5.)) * ((m1 + m2) ** (-1. / 5.))
        radius_16 = chirp * (lambda_tilde / 0.0042) ** (1.0 / 6.0)
        threshold = (2.38 - 3.606 * self.mtov / radius_16) * self.mtov
        log10_disk = -31.335 * (1 + -0.9760 * np.tanh((1.0474 - (m1 + m2) / threshold) / 0.05957))   # :255-262
        self.dynamical_mej = self.ejecta_mass
        self.disk_wind_mej = self.zeta * 10 ** log10_disk


class TwoComponentNSBH:
    """Dynamical (Kawaguchi et al. 2016) + disk-wind (Foucart et al. 2018) ejecta of a neutron-star black-hole merger
    (ejecta_relations.py:290-378)."""

    def __init__(self, mass_bh, mass_ns, chi_bh, lambda_ns, zeta):
        self.mass_bh, self.mass_ns, self.chi_bh = _arr(mass_bh), _arr(mass_ns), _arr(chi_bh)
        self.lambda_ns, self.zeta = _arr(lambda_ns), _arr(zeta)
        q = self.mass_ratio = self.mass_bh / self.mass_ns
        chi = self.chi_bh
        z1 = 1 + (1 - chi ** 2) ** (1 / 3) * ((1 + chi) ** (1 / 3) + (1 - chi) ** (1 / 3))
        z2 = np.sqrt(3 * chi ** 2 + z1 ** 2)
        self.risco = 3 + z2 - np.sign(chi) * np.sqrt((3 - z1) * (3 + z1 + 2 * z2))                  # :318-328
        self.ejecta_velocity = 1.5333330951369120e-2 * q + 0.19066667068621043
        compactness = calc_compactness_from_lambda(self.lambda_ns)
        baryonic = calc_baryonic_mass(self.mass_ns, compactness)
        unbound = 4.464e-2 * (q ** 0.2497) * (1 - 2 * compactness) / compactness + -2.269e-3 * (q ** 1.352) * self.risco \
            + 2.431 * (1 - self.mass_ns / baryonic) + -0.4159                                         # :349-356
        self.dynamical_mej = baryonic * np.maximum(unbound, 0)
        rho = (15 * self.lambda_ns) ** (-1 / 5)
        eta = q / (1 + q) ** 2
        remnant = 0.308 * (1 - 2 * rho) / (eta ** (1 / 3)) + -0.124 * self.risco * rho / eta + 0.283  # :369-376
        self.disk_wind_mej = self.zeta * (baryonic * (np.maximum(remnant, 0.0)) ** 1.536)


class OneComponentNSBH:
    """Neutron-star black-hole merger, one component (ejecta_relations.py:380-432)."""

    def __init__(self, mass_bh, mass_ns, chi_bh, lambda_ns):
        self.mass_bh, self.mass_ns, self.chi_bh = _arr(mass_bh), _arr(mass_ns), _arr(chi_bh)
        self.lambda_ns = _arr(lambda_ns)
        self.mass_ratio = self.mass_bh / self.mass_ns
        self.risco = self._isco_radius(self.chi_bh)
        self.ejecta_velocity = 1.5333330951369120e-2 * self.mass_ratio + 0.19066667068621043
        compactness = calc_compactness_from_lambda(self.lambda_ns)
        baryonic = calc_baryonic_mass(self.mass_ns, compactness)
        q = self.mass_ratio
        bracket = self.risco * (q ** 1.352) * -2.269e-3 + (q ** 0.2497) * (1 - 2 * compactness) * 4.464e-2 / compactness \
            + ((1 - self.mass_ns / baryonic) * 2.431 + -0.4159)
        self.ejecta_mass = baryonic * np.maximum(bracket, 0)

    @staticmethod
    def _isco_radius(chi):
        """Bardeen, Press & Teukolsky radius of the innermost stable circular orbit in units of G M / c^2."""
        z1 = 1 + ((1 - chi * chi) ** (1 / 3.0)) * (((1 + chi) ** (1 / 3.0)) + (1 - chi) ** (1 / 3.0))
        z2 = (3 * chi * chi + z1 * z1) ** (1 / 2.0)
        return 3 + z2 - np.sign(chi) * ((3 - z1) * (3 + z1 + 2 * z2)) ** (1 / 2.0)
