import sys, numpy as np
sys.path.insert(0, '.')
import redback_b200 as rb
from oracle import restated as r
t = np.array([0.1, 1.0, 5.0]); nu = np.full(3, 1e14)
kw = dict(mej=0.01, vej=0.1, beta=3.0, kappa=1.0)
for npre in (True, False):
    out = rb.kilonova_models.metzger_kilonova_model(t, 0.01, frequency=nu, output_format='flux_density', neutron_precursor_switch=npre, **kw)
    ref = r.metzger_kilonova_model(t, 0.01, frequency=nu, neutron_precursor_switch=npre, **kw)
    print(npre, out, ref, out/ref)
tt = np.geomspace(0.01, 40, 12)
out = rb.kilonova_models.metzger_kilonova_model(tt, 0.01, frequency=1e14, output_format='flux_density', **kw)
ref = r.metzger_kilonova_model(tt, 0.01, frequency=np.full(12,1e14), **kw)
print(out/ref)
