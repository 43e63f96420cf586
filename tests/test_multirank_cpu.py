"""CPU (gloo, world_size 2): the product's multi-GPU layer, redback_b200.distributed -- sample shards per rank, the
chunked evaluate + all-gather of per-sample log L with an injected evaluation callable standing in for
the device kernels, and the cost-balanced partition used for the afterglow path."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _value(idx):
    return np.sin(idx * 0.37) * 1e3 - idx


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    from redback_b200 import distributed as rd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # --- contiguous shards, uneven sizes, 3 chunks ---------------------------------------------------------
    full = dict(a=np.arange(n_total, dtype=np.float64), b=2.5)
    mine = rd.shard_params(full, rank, world)
    lo, hi = rd.shard_bounds(n_total, rank, world)
    assert mine["b"] == 2.5 and np.array_equal(mine["a"], np.arange(lo, hi))
    calls = []

    def evaluate(a, b, out):
        calls.append((a, b))
        out.copy_(torch.from_numpy(_value(mine["a"][a:b])))

    ev = rd.ShardedEvaluator(hi - lo, evaluate, chunks=3)
    got = ev.concatenated().numpy().copy()
    again = ev.concatenated().numpy()            # buffers are reused across calls
    assert np.array_equal(got, again)
    assert sum(b - a for a, b in calls[: len(calls) // 2]) == hi - lo
    # --- cost-balanced partition (afterglow): same collective, results restored to the caller's order -------
    rng = np.random.default_rng(5)
    thv, thc, g0 = np.arccos(rng.uniform(0, 1, n_total)), rng.uniform(0.01, 0.1, n_total), rng.uniform(100, 2000, n_total)
    cost, res = rd.afterglow_cost(thv, thc, g0)
    parts = rd.balanced_partition(cost, world)
    idx = parts[rank]

    def evaluate2(a, b, out):
        out.copy_(torch.from_numpy(_value(idx[a:b].astype(np.float64))))

    ev2 = rd.ShardedEvaluator(len(idx), evaluate2, chunks=2)
    restored = rd.unpermute(ev2.log_likelihood(), parts, n_total).numpy()
    # --- rows of E values per sample (light curves on every rank) ---------------------
    width = 7

    def evaluate3(a, b, out):
        out.copy_(torch.from_numpy(_value(mine["a"][a:b])[:, None] + np.arange(width)[None, :] * 0.5))

    ev3 = rd.ShardedEvaluator(hi - lo, evaluate3, chunks=3, width=width)
    curves = ev3.concatenated().numpy()
    assert curves.shape == (n_total, width)
    assert np.array_equal(curves, _value(np.arange(n_total, dtype=np.float64))[:, None] + np.arange(width)[None, :] * 0.5)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.stack([got, restored]))
    np.save(os.path.join(out_dir, f"c{rank}.npy"), np.array([cost[p].sum() for p in parts]))
    dist.destroy_process_group()


def test_two_rank_sharded_evaluator(tmp_path):
    world, n_total = 2, 1001
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "r0.npy")
    b = np.load(tmp_path / "r1.npy")
    expect = _value(np.arange(n_total, dtype=np.float64))
    assert np.array_equal(a, b)
    assert np.array_equal(a[0], expect) and np.array_equal(a[1], expect)
    per_rank_cost = np.load(tmp_path / "c0.npy")
    assert per_rank_cost.max() / per_rank_cost.min() < 1.02      # res spans 10..100 -> 100x cost spread per sample


def test_partitions():
    from redback_b200 import distributed as rd
    bounds = [rd.shard_bounds(1001, r, 3) for r in range(3)]
    assert bounds[0][0] == 0 and bounds[-1][1] == 1001 and all(bounds[i][1] == bounds[i + 1][0] for i in range(2))
    assert max(b - a for a, b in bounds) - min(b - a for a, b in bounds) <= 1
    rng = np.random.default_rng(0)
    cost = np.where(rng.uniform(size=10_000) < 0.006, 100.0, 1.0) * rng.lognormal(0, 0.3, 10_000)   # afterglow-like
    for world in (2, 4, 8):
        parts = rd.balanced_partition(cost, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(cost.size))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
        sums = np.array([cost[p].sum() for p in parts])
        contiguous = np.array([cost[slice(*rd.shard_bounds(cost.size, r, world))].sum() for r in range(world)])
        assert sums.max() / sums.min() < 1.05 and sums.max() / sums.min() <= contiguous.max() / contiguous.min()
    one = rd.ShardedEvaluator(5, lambda a, b, out: out.copy_(torch.arange(a, b, dtype=torch.float64)))
    assert one.concatenated().tolist() == [0, 1, 2, 3, 4]
