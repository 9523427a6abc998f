"""Data-parallel plumbing for the SGC-LL path: graphs are independent, so a
step shards them over the ranks with no data-path collective; the only exchange is ONE all-reduce of
a flat fp32 gradient buffer (NCCL over NVLink on the GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_graphs(n_nodes, world_size, rank, balance="work"):
    """Indices of the graphs rank `rank` owns.

    balance="count": contiguous equal-count blocks (what the weak-scaling benchmark uses).
    balance="work" : longest-processing-time bin packing on n_g^2 (the per-graph cost of the n x n
                     kernels) with equal graph counts per rank, for ragged batches (Sydney shape).
    """
    n = np.asarray(n_nodes, dtype=np.int64)
    B = n.size
    if balance == "count":
        per = (B + world_size - 1) // world_size
        return np.arange(rank * per, min(B, (rank + 1) * per), dtype=np.int64)
    if balance != "work":
        raise ValueError(balance)
    cap = (B + world_size - 1) // world_size
    order = np.argsort(-(n * n), kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    count = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(B, dtype=np.int64)
    for g in order:
        open_ranks = np.nonzero(count < cap)[0]
        r = open_ranks[np.argmin(load[open_ranks])]
        owner[g] = r
        load[r] += n[g] * n[g]
        count[r] += 1
    return np.sort(np.nonzero(owner == rank)[0])


class FlatGradBuffer(object):
    """All parameter gradients in one contiguous buffer; every `.grad` is a view into it, so the
    gradient exchange of a step is a single all-reduce (latency-bound: ~0.5 M floats)."""

    def __init__(self, params):
        self.params = list(params)
        total = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(total, device=first.device, dtype=first.dtype)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, world_size=None):
        if dist.is_available() and dist.is_initialized() and (world_size or dist.get_world_size()) > 1:
            dist.all_reduce(self.flat)

    def numel(self):
        return int(self.flat.numel())
