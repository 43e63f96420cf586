"""Developer aid: does the ORDER of the samples inside a rank's shard change the throughput of the afterglow path?
(balanced_partition hands every rank its samples sorted by descending cost.)  python tools/ag_order_probe.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import redback_b200 as rb
import bench_paths as bp
from redback_b200 import distributed as rd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
dev = torch.device("cuda:0")
w = [x for x in bp.workloads(rb, dev, 1.0) if x["key"].startswith("config4")][0]
y, sigma = bp.synthetic_data(w)
path = w["build"](y, sigma)
gp = w["prior"](np.random.default_rng(2024), n * 2)
cost, res = rd.afterglow_cost(gp["thv"], gp["thc"], gp["g0"])
part = rd.balanced_partition(cost, 2)[0]


def rate(index, label):
    p = {k: v[index] for k, v in gp.items()}
    params = path.pack(p, len(index))
    logl = torch.empty(len(index), dtype=torch.float64, device=dev)
    for _ in range(2):
        path.evaluate(params, want_model=False, want_logl=True, out_logl=logl)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        path.evaluate(params, want_model=False, want_logl=True, out_logl=logl)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    print(f"{label:34s} {len(index)} samples, cost {cost[index].sum():.4e}, {ms:8.2f} ms, "
          f"{len(index) * path.E / (ms * 1e-3):.4e} unit/s", flush=True)


rng = np.random.default_rng(0)
rate(np.arange(n), "contiguous first half")
rate(part, "balanced, descending cost")
rate(rng.permutation(part), "balanced, shuffled")
rate(part[::-1], "balanced, ascending cost")
