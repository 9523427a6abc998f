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

    def __init__(self, params, direct=()):
        """direct: parameters whose producers write the gradient straight into the view (the SGC-LL
        backward overwrites its outputs, include/agcn_sgcll.h), so autograd has nothing to add."""
        self.params = list(params)
        total = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(total, device=first.device, dtype=first.dtype)
        off = 0
        direct = set(id(p) for p in direct)
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            if id(p) in direct:
                p._agcn_grad_out = p.grad
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, world_size=None):
        if dist.is_available() and dist.is_initialized() and (world_size or dist.get_world_size()) > 1:
            dist.all_reduce(self.flat)

    def numel(self):
        return int(self.flat.numel())


class FlatParamBuffer(object):
    """All parameters in one contiguous buffer (every parameter keeps its identity and becomes a view
    into it), paired with a FlatGradBuffer: the optimizer then updates ONE tensor per step instead of
    one per parameter."""

    def __init__(self, params, grads):
        self.params = list(params)
        first = self.params[0]
        total = sum(p.numel() for p in self.params)
        assert total == grads.numel()
        self.flat = torch.empty(total, device=first.device, dtype=first.dtype)
        off = 0
        with torch.no_grad():
            for p in self.params:
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view
                off += p.numel()
        self.leaf = self.flat.requires_grad_(True)
        self.leaf.grad = grads.flat

    def chunks(self, grads, chunk=8192):
        """The flat buffer as a list of leaf tensors (storage aliases) with `.grad` aliases into the flat
        gradient: a multi-tensor optimizer then spreads one update over many thread blocks instead of
        walking a single 0.5 M-element tensor with a handful of them."""
        out = []
        for a in range(0, self.flat.numel(), chunk):
            t = self.flat.data[a:a + chunk].requires_grad_(True)
            t.grad = grads.flat[a:a + chunk]
            out.append(t)
        return out
