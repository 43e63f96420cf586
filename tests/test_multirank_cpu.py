"""CPU (gloo, world_size 2): the host-side sharding logic used by bench.py for N > 1 GPUs -- contiguous sample
shards per rank, no data-path collective except the all-gather of per-sample log L."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bench.shard_bounds(n_total, rank, world)
    # stand-in for the device kernel: a deterministic per-sample value computed from the global sample index
    rows = bench.prior_rows_numpy(np.random.default_rng(1000 + rank), hi - lo)
    local = torch.from_numpy(rows.sum(axis=0) + np.arange(lo, hi))
    sizes = [bench.shard_bounds(n_total, r, world) for r in range(world)]
    gathered = [torch.empty(b - a, dtype=torch.float64) for a, b in sizes]
    dist.all_gather(gathered, local)
    full = torch.cat(gathered)
    tm = torch.tensor([float(rank + 1)])
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)   # max-over-ranks timing reduction
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    assert tm.item() == world
    dist.destroy_process_group()


def test_two_rank_sharding_and_allgather(tmp_path):
    world, n_total = 2, 1000   # equal shards: bench.py uses a fixed number of samples per GPU (weak scaling)
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "r0.npy")
    b = np.load(tmp_path / "r1.npy")
    assert a.shape == (n_total,) and np.array_equal(a, b)
    sys.path.insert(0, ROOT)
    import bench
    bounds = [bench.shard_bounds(n_total, r, world) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n_total and bounds[0][1] == bounds[1][0]
    assert abs((bounds[0][1] - bounds[0][0]) - (bounds[1][1] - bounds[1][0])) <= 1
    odd = [bench.shard_bounds(1001, r, 3) for r in range(3)]
    assert odd[0][0] == 0 and odd[-1][1] == 1001 and all(odd[i][1] == odd[i + 1][0] for i in range(2))
