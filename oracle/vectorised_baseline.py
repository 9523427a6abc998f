"""`cpu_vectorised`: the "good CPU implementation" baseline of BASELINE.md section 4 item 2.  TEST / BENCH
INFRASTRUCTURE ONLY (imported by tests/ and by bench.py's CPU legs, never by the product).

The same SimpleAGCN training step as oracle/network_oracle.py (4 x SGC_LL + DenseMol + GraphGatherMol + logits + loss,
forward + backward by torch autograd), but batched: graphs are bucketed by size, zero-padded inside a bucket and run
through torch.bmm / broadcasting with row masks, on all host threads.  It computes what the GPU arm computes for the
same semantics: in `reference_literal` mode res_L == I (graphconv.py:195-200) and the similarity matrix is
only an optional output, so it is not evaluated (the GPU arm makes it lazy too); in `paper` mode the
learned metric, the Gaussian kernel and the normalised residual Laplacian are evaluated and differentiated.

tests/test_oracle_golden.py checks this module against the per-graph oracle (loss and gradients).
"""
from __future__ import annotations

import numpy as np
import torch


def make_buckets(n_nodes, limits=(16, 32, 48, 64, 96, 132, 192, 256, 384, 512, 768, 1024, 2048, 4096)):
    """Indices of the graphs per padded size (smallest limit >= n)."""
    n_nodes = np.asarray(n_nodes)
    out = {}
    for g, n in enumerate(n_nodes):
        lim = next((l for l in limits if n <= l), int(n))
        out.setdefault(lim, []).append(g)
    return {k: np.asarray(v) for k, v in sorted(out.items())}


def _leaky(x, alpha):
    return torch.relu(x) - alpha * torch.relu(-x)


def sgc_ll_bucket(X, L, mask, p, K, laplacian, metric_grad):
    """X [b, N, F] (rows >= n_g zero), L [b, N, N] (zero outside the real block), mask [b, N] bool -> activated
    output [b, N, Fo] with zero padding rows.  graphconv.py:145-251 for a whole bucket."""
    b, N, F = X.shape
    m2 = (mask[:, :, None] & mask[:, None, :])
    eye = torch.eye(N, dtype=X.dtype).expand(b, N, N) * m2
    numel = mask.sum(1).to(X.dtype) ** 2                                   # elements of the real n x n block
    if laplacian == "reference_literal":
        res_L = eye
    else:
        Xm, Mm = (X, p["M_L"]) if metric_grad == "full" else (X.detach(), p["M_L"].detach())
        xw = Xm @ Mm
        d2 = ((xw[:, :, None, :] - xw[:, None, :, :]) ** 2).sum(-1) if N <= 256 else \
            (torch.cdist(xw, xw, compute_mode="donot_use_mm_for_euclid_dist") ** 2)
        pos = d2 > 0
        dist = torch.where(pos, torch.sqrt(torch.where(pos, d2, torch.ones_like(d2))), torch.zeros_like(d2))
        W = torch.exp(-dist) * m2 * (1 - torch.eye(N, dtype=X.dtype))
        d = W.sum(1)
        ok = d > 2.0 ** -80
        dis = torch.where(ok, 1.0 / torch.sqrt(torch.where(ok, d, torch.ones_like(d))), torch.zeros_like(d))
        res_L = eye - dis[:, :, None] * W * dis[:, None, :]
        if metric_grad != "full":
            res_L = res_L.detach()
    ss = (res_L * res_L).sum((1, 2))
    inv = torch.rsqrt(torch.clamp(ss, min=1e-300))
    scale = torch.minimum(inv * numel, torch.ones_like(inv))                # tf.clip_by_average_norm(., 1)
    L_all = _leaky(res_L * scale[:, None, None], p["alpha"]) + L          # graphconv.py:212-216
    T = [X]
    if K > 1:
        T.append(torch.bmm(L_all, X))
    for _ in range(2, K):
        T.append(2 * torch.bmm(L_all, T[-1]) - T[-2])
    xc = torch.stack(T, 0).permute(1, 2, 3, 0).reshape(b, N, F * K)
    y = xc @ p["weight"] + p["bias"]
    return torch.relu(y) * mask[:, :, None]                                # pad AFTER the bias (graphconv.py:249-251)


def simple_agcn_loss(buckets, layer_params, head_params, K, global_batch, laplacian="reference_literal",
                     metric_grad="reference", loss_kind="sigmoid_ce"):
    """buckets: list of dicts X [b,N,F], L [b,N,N], mask [b,N], targets, weights.  Returns the scalar loss."""
    total = 0.0
    for bk in buckets:
        x = bk["X"]
        for p in layer_params:
            x = sgc_ll_bucket(x, bk["L"], bk["mask"], p, K, laplacian, metric_grad)
        d = (x @ head_params["dense_W"] + head_params["dense_b"]) * bk["mask"][:, :, None]
        mol = torch.tanh(d.sum(1))
        logits = mol @ head_params["head_W"] + head_params["head_b"]
        if loss_kind == "sigmoid_ce":
            t = bk["targets"]
            costs = torch.clamp(logits, min=0) - logits * t + torch.log1p(torch.exp(-logits.abs()))
            total = total + (costs * bk["weights"]).sum()
        else:
            ce = torch.logsumexp(logits, 1) - (logits * bk["targets"]).sum(1)
            total = total + (ce * bk["weights"]).sum()
    return total / global_batch


def prepare_buckets(X, L, n_nodes, targets, weights, dtype=torch.float32):
    """Padded wire-layout batch (numpy) -> size buckets of torch tensors."""
    out = []
    for lim, idx in make_buckets(n_nodes).items():
        lim = min(lim, X.shape[1])
        n = torch.as_tensor(np.asarray(n_nodes)[idx].astype(np.int64))
        out.append({"X": torch.as_tensor(X[idx, :lim]).to(dtype), "L": torch.as_tensor(L[idx, :lim, :lim]).to(dtype),
                    "mask": torch.arange(lim)[None, :] < n[:, None],
                    "targets": torch.as_tensor(targets[idx]).to(dtype), "weights": torch.as_tensor(weights[idx]).to(dtype)})
    return out


def make_params(dims, K, Fm, Nt, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)

    def glorot(r, c):
        lim = float(np.sqrt(6.0 / (r + c)))
        return ((torch.rand(r, c, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype).requires_grad_(True)

    layers = [{"weight": glorot(dims[i] * K, dims[i + 1]), "bias": torch.zeros(dims[i + 1], dtype=dtype, requires_grad=True),
               "M_L": glorot(dims[i], dims[i]), "alpha": torch.ones(1, dtype=dtype, requires_grad=True)}
              for i in range(len(dims) - 1)]
    head = {"dense_W": glorot(dims[-1], Fm), "dense_b": torch.zeros(Fm, dtype=dtype, requires_grad=True),
            "head_W": (torch.randn(Fm, Nt, generator=g, dtype=torch.float64) * 0.01).to(dtype).requires_grad_(True),
            "head_b": torch.zeros(Nt, dtype=dtype, requires_grad=True)}
    return layers, head
