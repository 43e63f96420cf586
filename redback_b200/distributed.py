"""Multi-GPU evaluation of a parameter batch: one process per GPU (torch.distributed), samples sharded across
ranks, and the path's only exchange -- an all-gather of the per-sample log-likelihoods.

Every sample is independent, so there is no data-path collective inside the model evaluation.  Two partitions:

* `shard_bounds`: contiguous blocks of ceil(N / world) samples (the default: uniform cost per sample);
* `balanced_partition`: for paths whose cost per sample varies by orders of magnitude -- the redback afterglow picks
  its jet resolution from (thv, thc, g0) (afterglow_models/base_models.py:932-935), a 100x spread in work -- samples
  are sorted by cost and dealt to the ranks boustrophedon-wise, so every rank gets the same number of samples and
  (nearly) the same total cost; `unpermute` restores the caller's order after the gather.

`ShardedEvaluator.log_likelihood` evaluates the local shard chunk by chunk and issues the all-gather of chunk i on a
side stream while the kernels of chunk i+1 run, so the collective is hidden behind compute.  The evaluation callable
is injected, so the same code runs under gloo on CPU in the tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total, rank, world):
    """Contiguous shard [lo, hi) of `n_total` samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_params(params_batch, rank, world, index=None):
    """This rank's slice of every array in `params_batch` ({name: array[N] | scalar}): the contiguous shard, or the
    rows `index` (from `balanced_partition`)."""
    n = None
    for v in params_batch.values():
        if np.ndim(v) > 0:
            n = len(v)
            break
    if n is None:
        return dict(params_batch)
    if index is None:
        lo, hi = shard_bounds(n, rank, world)
        index = slice(lo, hi)
    return {k: (v[index] if np.ndim(v) > 0 else v) for k, v in params_batch.items()}


def balanced_partition(cost, world, lpt_rounds=2048):
    """Deal samples to `world` ranks so that the per-rank cost sums are nearly equal and the counts differ by at most
    one.  Returns a list of index arrays (rank r evaluates samples parts[r], in that order: ascending cost).  Samples are sorted by
    descending cost; the heavy head (`lpt_rounds` rounds of `world` samples) is dealt round by round with the largest
    sample of the round going to the currently lightest rank (longest-processing-time greedy with equal counts), the
    light tail in snake order (0..w-1, w-1..0, ...)."""
    cost = np.asarray(cost, dtype=np.float64)
    order = np.argsort(-cost, kind="stable")
    n = order.size
    owner = np.empty(n, dtype=np.int64)
    sums = np.zeros(world)
    head = min(n, world * int(lpt_rounds))
    for a in range(0, head, world):
        idx = order[a:a + world]                      # descending cost
        ranks = np.argsort(sums, kind="stable")[: idx.size]   # ascending load
        owner[a:a + idx.size] = ranks
        sums[ranks] += cost[idx]
    if head < n:
        # tail: snake, started from the lightest rank so that the head's residual imbalance is not amplified
        rot = np.argsort(sums, kind="stable")
        pos = np.arange(n - head)
        rnd, k = np.divmod(pos, world)
        owner[head:] = rot[np.where(rnd % 2 == 0, k, world - 1 - k)]
    # ascending cost inside a rank: measured on the afterglow path (tools/ag_order_probe.py, 4e5 samples), the same
    # shard runs 7 % slower with its heavy samples first and 2.6 % faster with them last than in random order
    return [order[owner == r][::-1].copy() for r in range(world)]


def unpermute(gathered_parts, parts, n_total):
    """Inverse of `balanced_partition`: gathered_parts[r] are rank r's results for samples parts[r]."""
    first = gathered_parts[0]
    if isinstance(first, torch.Tensor):
        out = torch.empty((n_total,) + tuple(first.shape[1:]), dtype=first.dtype, device=first.device)
        for g, idx in zip(gathered_parts, parts):
            out[torch.as_tensor(idx, device=first.device)] = g[: len(idx)]
        return out
    out = np.empty((n_total,) + tuple(np.shape(first)[1:]), dtype=np.asarray(first).dtype)
    for g, idx in zip(gathered_parts, parts):
        out[idx] = np.asarray(g)[: len(idx)]
    return out


def afterglow_cost(thv, thc, g0, n_rot=None):
    """Relative cost of one redback-afterglow sample: flop-eq ~ 5.4e4 C + 8.8e4 R with R = res from the reference's
    rule (base_models.py:932-935) and C = R * N_rot cells."""
    thv, thc, g0 = (np.asarray(x, dtype=np.float64) for x in (thv, thc, g0))
    res = np.clip(((2 - 10 * np.maximum(thv - thc, 0)) * thc * g0).astype(np.int64), 10, 100)
    rot = res if n_rot is None else n_rot      # rotstep = 2 pi / res -> N_rot = res
    return 5.4e4 * res * rot + 8.8e4 * res, res


class ShardedEvaluator:
    """All-gathered per-sample outputs of a sharded evaluation.

    evaluate(lo, hi, out) must write the results of local samples [lo, hi) into `out` (a view of length hi - lo) and
    be stream-ordered on the current CUDA stream (any redback_b200 path's `evaluate(..., out_logl=out)` is).
    `width`: None for one value per sample (log L); E for a row of E values per sample (light curves / magnitudes on
    every rank) -- `out` is then a [hi - lo, E] view.
    """

    def __init__(self, n_local, evaluate, device=None, group=None, chunks=4, dtype=torch.float64, width=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_local = int(n_local)
        self.evaluate = evaluate
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        self.chunks = max(1, min(int(chunks), self.n_local)) if self.n_local else 1
        # every rank must take part in every collective with the same sizes: agree on the padded shard length
        n_max = torch.tensor([self.n_local], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(n_max, op=dist.ReduceOp.MAX, group=group)
        self.n_pad = int(n_max.item())
        self.chunk = -(-self.n_pad // self.chunks) if self.n_pad else 0
        tail = () if width is None else (int(width),)
        self.tail = tail
        self.local = torch.zeros((self.chunk * self.chunks,) + tail, dtype=dtype, device=self.device)
        self.stage = [torch.empty((self.world * self.chunk,) + tail, dtype=dtype, device=self.device) for _ in range(2)]
        self.gathered = torch.empty((self.world, self.chunk * self.chunks) + tail, dtype=dtype, device=self.device)
        self.comm = torch.cuda.Stream(self.device) if self.cuda else None
        if self.world > 1:
            sizes = torch.zeros(self.world, dtype=torch.int64, device=self.device)
            sizes[self.rank] = self.n_local
            dist.all_reduce(sizes, group=group)
            self.sizes = [int(s) for s in sizes.tolist()]
        else:
            self.sizes = [self.n_local]

    def log_likelihood(self):
        """Evaluate the local shard and return the list of per-rank result tensors (views, rank r has sizes[r] rows).
        With world == 1 no collective is issued."""
        if self.world == 1:
            if self.n_local:
                self.evaluate(0, self.n_local, self.local[: self.n_local])
            return [self.local[: self.n_local]]
        cur = torch.cuda.current_stream(self.device) if self.cuda else None
        works = []
        for c in range(self.chunks):
            lo = c * self.chunk
            hi = min(lo + self.chunk, self.n_local)
            if hi > lo:
                self.evaluate(lo, hi, self.local[lo:hi])
            src = self.local[lo: lo + self.chunk]
            stage = self.stage[c % 2]
            if self.cuda:
                # gather chunk c on the side stream while the compute stream goes on with chunk c+1
                self.comm.wait_stream(cur)
                with torch.cuda.stream(self.comm):
                    dist.all_gather_into_tensor(stage, src, group=self.group)
                    self.gathered[:, lo: lo + self.chunk].copy_(stage.view((self.world, self.chunk) + self.tail))
            else:
                parts = [torch.empty_like(src) for _ in range(self.world)]
                works.append(dist.all_gather(parts, src, group=self.group))
                self.gathered[:, lo: lo + self.chunk].copy_(torch.stack(parts))
        if self.cuda:
            cur.wait_stream(self.comm)
        return [self.gathered[r, : self.sizes[r]] for r in range(self.world)]

    def concatenated(self):
        """log_likelihood() as one tensor in rank order (the contiguous-shard layout of the full batch)."""
        return torch.cat(self.log_likelihood())
