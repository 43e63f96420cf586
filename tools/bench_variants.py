"""A/B timing of alternate builds of the library (same ABI) on the supernova workloads.

    python tools/bench_variants.py lib_a.so lib_b.so ...

Each library is timed in its own subprocess (REDBACK_B200_LIB) on the headline workload (slsn + Blackbody, E = 300),
the same with the CutoffBlackbody SED and on arnett_bolometric (E = 100); prints one line per library."""
import json
import os
import subprocess
import sys

CHILD = r'''
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(sys.argv[1])), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(sys.argv[1])), "..", "tests"))
import helpers as h
from redback_b200 import _capi
from redback_b200.engine import SupernovaPath
def timed(path, params, steps=5, warmup=3):
    for _ in range(warmup):
        path.evaluate(params, want_model=False, want_logl=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        path.evaluate(params, want_model=False, want_logl=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
rng = np.random.default_rng(0)
out = {}
t2, nu2 = h.config2_epochs(np.random.default_rng(0))
p = SupernovaPath(_capi.ENGINE_MAGNETAR, _capi.SED_BLACKBODY, t2, frequency=nu2, y=np.full(300, 1e-3), sigma=1e-4)
n = 2_000_000
out["slsn_bb_E300"] = n * 300 / timed(p, p.pack(h.slsn_prior(rng, n), n)) / 1e6
p = SupernovaPath(_capi.ENGINE_MAGNETAR, _capi.SED_CUTOFF_BLACKBODY, t2, frequency=nu2, y=np.full(300, 1e-3), sigma=1e-4)
out["slsn_cutoff_E300"] = n * 300 / timed(p, p.pack(h.slsn_prior(rng, n), n)) / 1e6
t1 = h.config1_time()
p = SupernovaPath(_capi.ENGINE_NICKEL, _capi.SED_BOLOMETRIC, t1, y=np.full(100, 1e42), sigma=1e41)
out["arnett_bol_E100"] = n * 100 / timed(p, p.pack(h.arnett_bolometric_prior(rng, n), n)) / 1e6
print(json.dumps(out))
'''

for lib in sys.argv[1:]:
    env = dict(os.environ, REDBACK_B200_LIB=os.path.abspath(lib))
    res = subprocess.run([sys.executable, "-c", CHILD, __file__], env=env, capture_output=True, text=True)
    line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-400:]
    print(os.path.basename(lib), line, "(1e9 sample*epoch/s)")
