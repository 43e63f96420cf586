"""A/B timing of alternate builds of the library (same ABI) on the two magnitude workloads of the bench (config 3:
kilonova LSST magnitudes, config 5: Arnett population magnitudes), a quarter of the bench sample counts.

    python tools/phot_variants.py lib_a.so lib_b.so ...
"""
import json, os, subprocess, sys
CHILD = r'''
import json, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import redback_b200 as rb, bench_paths as bp
dev = torch.device("cuda:0")
out = {}
for w in bp.workloads(rb, dev, 0.25):
    if w["key"].split("_")[0] not in ("config3", "config5"): continue
    y, sigma = bp.synthetic_data(w)
    path = w["build"](y, sigma)
    n = w["n"]; p = w["prior"](np.random.default_rng(3), n); params = path.pack(p, n)
    fn = lambda: path.evaluate(params, want_model=True, want_logl=False)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    out[w["key"].split("_")[0]] = n * path.E / (e0.elapsed_time(e1) / 3) / 1e5
print(json.dumps(out))
'''
for lib in sys.argv[1:]:
    env = dict(os.environ, REDBACK_B200_LIB=os.path.abspath(lib))
    res = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-300:], "(1e8 unit/s)")
