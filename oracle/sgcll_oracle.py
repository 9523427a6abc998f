"""CPU oracle for the SGC-LL hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference's SGC-LL layer
(uta-smile/Adaptive-Graph-Convolutional-Network).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product (``agcn_b200``) never imports it.

Parity pin status
-----------------
* ``metric_block_literal`` (the body of the reference's ``func`` closure,
  models/layers/graphconv.py:163-209 and graphconv_reslap.py:136-182) is PINNED:
  ``tests/golden/make_golden.py`` executes the reference's own source for that
  closure and for ``models/graph_structure.py`` in the build container and
  commits the outputs; ``tests/test_oracle_golden.py`` checks this file against
  them.
* Everything that the reference delegates to TensorFlow (``tf.clip_by_norm``,
  ``tf.clip_by_average_norm``, ``tf.nn.relu``, ``tf.matmul``, ``tf.pad``,
  ``tf.slice``, autodiff, ``py_func`` = stop-gradient) is restated from TF's
  documented semantics because TensorFlow (~0.12, Python 2) cannot be installed
  here: that part is "parity unpinned" -- the reference ships no golden vectors
  or tests for this path.

Two named semantics:
* ``laplacian="reference_literal"``: what the reference code literally computes.
  Inside ``func`` ``D * W * D`` is an elementwise product of ndarrays and W has
  a zero diagonal, so ``res_L == I`` exactly (graphconv.py:195-200).
* ``laplacian="paper"``: the intended ``I - D^-1/2 W D^-1/2``.
* ``metric_grad="reference"``: ``tf.py_func`` has no gradient, so nothing flows
  to ``M_L`` nor to ``x`` through the metric (graphconv.py:211).
* ``metric_grad="full"``: the differentiable metric of the paper.
"""
from __future__ import annotations

import numpy as np
import torch

EPS_F32 = float(np.spacing(np.array(0, np.float32)))  # 1.4e-45, graphconv.py:196
DEGREE_FLOOR = 2.0 ** -80                               # see residual_laplacian


# --------------------------------------------------------------------------
# metric block, literal restatement (numpy, fp32, interpreted double loop)
# --------------------------------------------------------------------------
def metric_block_literal(x: np.ndarray, M: np.ndarray, flavour: str = "SGC_LL"):
    """graphconv.py:163-209 (flavour "SGC_LL") / graphconv_reslap.py:136-182
    (flavour "SGC_LL_Reslap").  Returns (L, W) float32, exactly as the
    reference's py_func does -- including the elementwise ``D * W * D``.
    """
    x = np.asarray(x, np.float32)
    M = np.asarray(M, np.float32)
    x_w = np.dot(x, M)                                   # :164
    n = x_w.shape[0]
    rows = [[] for _ in range(n)]                        # :168
    for i in range(n):                                   # :169-178
        for j in range(n):
            if j == i:
                rows[i].append(0.0)
                continue
            dist = np.linalg.norm(x_w[i] - x_w[j])
            rows[i].append(1 * np.exp(-1 * dist))
    W = np.asarray(rows).astype(np.float32)              # :180
    d = W.sum(axis=0)                                    # :195
    d += np.spacing(np.array(0, W.dtype))                # :196
    if flavour == "SGC_LL":
        d = 1 / np.sqrt(d)                               # :197
        D = np.diag(d.squeeze())                         # :198
    else:
        d = np.power(d.squeeze(), -0.5).flatten()        # reslap :170
        D = np.diag(d)                                   # reslap :171
    I = np.identity(d.size, dtype=W.dtype)               # :199
    L = I - D * W * D                                    # :200  (elementwise!)
    return L.astype(np.float32), W.astype(np.float32)    # :209


def similarity_numpy(x: np.ndarray, M: np.ndarray) -> np.ndarray:
    """Vectorised fp32 restatement of graphconv.py:164-180 (W only), direct
    difference norm like ``np.linalg.norm(u - v)``."""
    x_w = np.dot(np.asarray(x, np.float32), np.asarray(M, np.float32))
    diff = x_w[:, None, :] - x_w[None, :, :]
    dist = np.sqrt(np.einsum("ijk,ijk->ij", diff, diff).astype(np.float32))
    W = np.exp(-dist).astype(np.float32)
    np.fill_diagonal(W, 0.0)
    return W


# --------------------------------------------------------------------------
# TensorFlow op semantics restated
# --------------------------------------------------------------------------
def clip_by_norm(t: torch.Tensor, clip_norm: float = 1.0) -> torch.Tensor:
    """tf.clip_by_norm: t * clip_norm * min(rsqrt(sum t^2), 1/clip_norm)."""
    ss = (t * t).sum()
    inv = torch.where(ss > 0, torch.rsqrt(torch.where(ss > 0, ss, torch.ones_like(ss))),
                      torch.full_like(ss, float("inf")))
    scale = clip_norm * torch.minimum(inv, torch.full_like(ss, 1.0 / clip_norm))
    return t * scale


def clip_by_average_norm(t: torch.Tensor, clip_norm: float = 1.0) -> torch.Tensor:
    """tf.clip_by_average_norm: t * clip_norm * min(numel * rsqrt(sum t^2), 1/clip_norm)."""
    ss = (t * t).sum()
    inv = torch.where(ss > 0, torch.rsqrt(torch.where(ss > 0, ss, torch.ones_like(ss))),
                      torch.full_like(ss, float("inf")))
    scale = clip_norm * torch.minimum(inv * t.numel(), torch.full_like(ss, 1.0 / clip_norm))
    return t * scale


def leaky(x: torch.Tensor, alpha) -> torch.Tensor:
    """models/operators/model_operatos.py:534-558 with a Variable alpha:
    relu(x) - alpha * relu(-x).  relu'(0) = 0 on both branches (TF ReluGrad)."""
    return torch.relu(x) - alpha * torch.relu(-x)


# --------------------------------------------------------------------------
# the metric block in torch (any dtype), both semantics
# --------------------------------------------------------------------------
def similarity(x_w: torch.Tensor) -> torch.Tensor:
    """W_ij = exp(-||xw_i - xw_j||_2), W_ii = 0 (graphconv.py:168-180).  The
    sqrt has sub-gradient 0 at exact duplicates."""
    diff = x_w[:, None, :] - x_w[None, :, :]
    d2 = (diff * diff).sum(-1)
    pos = d2 > 0
    dist = torch.where(pos, torch.sqrt(torch.where(pos, d2, torch.ones_like(d2))), torch.zeros_like(d2))
    W = torch.exp(-dist)
    eye = torch.eye(x_w.shape[0], dtype=torch.bool)
    return torch.where(eye, torch.zeros_like(W), W)


def residual_laplacian(W: torch.Tensor, laplacian: str) -> torch.Tensor:
    n = W.shape[0]
    I = torch.eye(n, dtype=W.dtype)
    if laplacian == "reference_literal":
        # graphconv.py:198-200: diag(d) * W * diag(d) elementwise with zero-diagonal W == 0
        return I
    if laplacian != "paper":
        raise ValueError(laplacian)
    d = W.sum(dim=0)                                       # :195 (column sums)
    # :196-197.  eps = 1.4e-45 is a denormal (0 under FTZ): the guard is written
    # out: d <= DEGREE_FLOOR = 2^-80 -> d^-1/2 := 0, i.e. a node farther
    # than ~55 from every other node in the learned metric is isolated and its row
    # of L is the identity row.  (The reference evaluates 1/sqrt(d + 1.4e-45) ~ 1e22
    # on sums of float32 denormals there; the floor keeps every term and its
    # derivative finite in fp32.)
    pos = d > DEGREE_FLOOR
    dis = torch.where(pos, 1.0 / torch.sqrt(torch.where(pos, d, torch.ones_like(d))), torch.zeros_like(d))
    return I - (dis[:, None] * W) * dis[None, :]


# --------------------------------------------------------------------------
# one graph through one SGC-LL layer
# --------------------------------------------------------------------------
def sgc_ll_graph(x, L_int, params, K, variant="SGC_LL", laplacian="reference_literal",
                 metric_grad="reference", L_prev=None, compute_similarity=True):
    """x [n,F], L_int [n,n], params dict(weight[F*K,Fo], bias[Fo], M_L[F,F],
    alpha[1] (, beta[1])) -> (y [n,Fo] pre-activation, res_L, res_W, L_all).

    graphconv.py:145-251 (variant "SGC_LL"); graphconv_reslap.py:119-229
    (variant "SGC_LL_Reslap", L_prev = the previous save_lap layer's L_all or None).
    """
    n, F = x.shape
    M_L, alpha = params["M_L"], params["alpha"]
    xm, Mm = (x, M_L) if metric_grad == "full" else (x.detach(), M_L.detach())
    if compute_similarity or laplacian != "reference_literal":
        x_w = xm @ Mm                                       # :164
        res_W = similarity(x_w)
    else:
        # callers that only need y / gradients in literal mode: res_L == I whatever W is (:198-200), so the
        # O(n^2 F) similarity (returned, never used downstream) is skipped; res_W comes back as None
        res_W = None
    res_L = residual_laplacian(res_W if res_W is not None else torch.zeros(n, n, dtype=x.dtype), laplacian)
    if metric_grad != "full":                               # py_func: no gradient (:211)
        res_W, res_L = (None if res_W is None else res_W.detach()), res_L.detach()
    if variant == "SGC_LL":
        res_L = leaky(clip_by_average_norm(res_L, 1.0), alpha)     # :212-213
        L_all = res_L + L_int                                      # :216
    elif variant == "SGC_LL_Reslap":
        res_L = leaky(clip_by_norm(res_L, 1.0), alpha)             # reslap :185-186
        if L_prev is not None:
            L_all = res_L + L_int + L_prev * params["beta"]        # reslap :190
        else:
            L_all = res_L + L_int                                  # reslap :193
        L_all = leaky(clip_by_norm(L_all, 1.0), alpha)             # reslap :194-195
    else:
        raise ValueError(variant)
    # Chebyshev recurrence (:221-236); intended stacking [K, n, F] for any K
    T = [x]
    if K > 1:
        T.append(L_all @ x)
    for _ in range(2, K):
        T.append(2 * (L_all @ T[-1]) - T[-2])
    xc = torch.stack(T, 0).permute(1, 2, 0).reshape(n, F * K)     # :238-244, col = f*K + k
    y = xc @ params["weight"] + params["bias"]                    # :245-247
    return y, res_L, res_W, L_all


def sgc_ll_batch(X, L, n_nodes, params, K, variant="SGC_LL", laplacian="reference_literal",
                 metric_grad="reference", L_prev=None, activation="relu", compute_similarity=True):
    """Whole padded batch.  X [B,Nmax,F], L [B,Nmax,Nmax], n_nodes [B];
    L_prev: list of B [n,n] tensors or None.  Returns (Y [B,Nmax,Fo] activated,
    zero rows >= n (graphconv.py:249-251 pads AFTER the bias), lists res_L,
    res_W, L_all of unpadded [n,n])."""
    B, Nmax, _ = X.shape
    Fo = params["weight"].shape[1]
    Ys, RL, RW, LA = [], [], [], []
    for g in range(B):
        n = int(n_nodes[g])
        y, rl, rw, la = sgc_ll_graph(X[g, :n], L[g, :n, :n], params, K, variant, laplacian, metric_grad,
                                     None if L_prev is None else L_prev[g], compute_similarity)
        if activation == "relu":
            y = torch.relu(y)
        elif activation not in (None, "linear"):
            y = getattr(torch, activation)(y)
        Ys.append(torch.cat([y, torch.zeros(Nmax - n, Fo, dtype=y.dtype)], 0))
        RL.append(rl), RW.append(rw), LA.append(la)
    return torch.stack(Ys, 0), RL, RW, LA


def make_params(F, Fo, K, variant="SGC_LL", seed=0, dtype=torch.float32, perturb=True):
    """Glorot-uniform weight / M_L, zero bias, alpha = beta = 1 (graphconv.py:66-83,
    graphconv_reslap.py:21-43).  ``perturb`` moves bias/alpha/beta off their
    initial values so that tests exercise them."""
    g = torch.Generator().manual_seed(seed)

    def glorot(r, c):
        lim = float(np.sqrt(6.0 / (r + c)))
        return (torch.rand(r, c, generator=g, dtype=torch.float64) * 2 - 1) * lim

    p = {"weight": glorot(F * K, Fo), "bias": torch.zeros(Fo, dtype=torch.float64),
         "M_L": glorot(F, F), "alpha": torch.ones(1, dtype=torch.float64)}
    if variant == "SGC_LL_Reslap":
        p["beta"] = torch.ones(1, dtype=torch.float64)
    if perturb:
        p["bias"] = torch.randn(Fo, generator=g, dtype=torch.float64) * 0.1
        p["alpha"] = torch.tensor([0.7], dtype=torch.float64)
        if "beta" in p:
            p["beta"] = torch.tensor([0.6], dtype=torch.float64)
    return {k: v.to(dtype) for k, v in p.items()}


# --------------------------------------------------------------------------
# GraphPoolMol restated (models/layers/graphpool.py:91-105)
# --------------------------------------------------------------------------
def graph_pool_literal(x: np.ndarray, L: np.ndarray) -> np.ndarray:
    """Line-by-line restatement of the ``func`` closure of graphpool.py:91-105: atom i takes the
    feature-wise maximum over the atoms its Laplacian row marks (non-zero entries); a row without
    a non-zero entry keeps its own features."""
    acc_x = []
    for i, l in enumerate(list(L)):
        idx = np.nonzero(l)
        self_neighbor_atoms = x[idx]
        if len(self_neighbor_atoms) != 0:
            pooled_x = np.amax(self_neighbor_atoms, axis=0)
        else:
            pooled_x = x[i]
        acc_x.append(pooled_x)
    return np.vstack(acc_x)


def graph_pool(x: torch.Tensor, L: torch.Tensor) -> torch.Tensor:
    """Vectorised form of the same (differentiable through the arg-max, which the reference is not)."""
    mask = L != 0
    big = x.unsqueeze(0).expand(L.shape[0], -1, -1).masked_fill(~mask.unsqueeze(-1), float("-inf"))
    pooled = big.max(dim=1).values
    empty = ~mask.any(dim=1)
    return torch.where(empty.unsqueeze(-1), x, pooled)


# --------------------------------------------------------------------------
# host-side graph preprocessing restated (models/graph_structure.py:75-130)
# --------------------------------------------------------------------------
def compute_laplacian_dense(adj_lists) -> np.ndarray:
    """Dense float64 restatement of Graph.compute_laplacian:
    A (undirected, from_dict_of_lists :79-83) -> A+I -> D~^-1/2 (A+I) D~^-1/2
    (:110-125) -> L = I - D^-1/2 A^ D^-1/2 with D from the column sums of A^
    (:87-104).  The adjacency is cast to float32 (:127); the rest runs in
    float64 because ``sp.eye`` promotes."""
    n = len(adj_lists)
    A = np.zeros((n, n), np.float32)
    for i, nbrs in enumerate(adj_lists):
        for j in nbrs:
            A[i, j] = 1.0
            A[j, i] = 1.0
    A = A.astype(np.float64) + np.eye(n)
    rowsum = A.sum(1)
    with np.errstate(divide="ignore"):
        dinv = np.power(rowsum, -0.5)
    dinv[np.isinf(dinv)] = 0.0
    An = (A * dinv[None, :]).T * dinv[None, :]
    d = An.sum(axis=0) + np.spacing(np.array(0, An.dtype))
    d = 1 / np.sqrt(d)
    return np.eye(n) - (d[:, None] * An) * d[None, :]


# --------------------------------------------------------------------------
# synthetic inputs of the BASELINE shapes
# --------------------------------------------------------------------------
def tox21_like_features(rng: np.random.Generator, n: int) -> np.ndarray:
    """75-d one-hot-ish atom features (blocks 44,11,5,7 one-hot; charge, radicals;
    one-hot 5; aromatic flag), mirroring utils/feature/graph_features.py:156-180.
    Yields exactly duplicated rows on purpose."""
    x = np.zeros((n, 75), np.float32)
    col = 0
    for width, conc in ((44, 4), (11, 4), (5, 3), (7, 3)):
        idx = np.minimum(rng.geometric(1.0 / conc, n) - 1, width - 1)
        x[np.arange(n), col + idx] = 1.0
        col += width
    x[:, col] = rng.choice([-1.0, 0.0, 1.0], n, p=[0.03, 0.94, 0.03]); col += 1
    col += 1  # radical electrons: 0
    x[np.arange(n), col + np.minimum(rng.geometric(0.5, n) - 1, 4)] = 1.0; col += 5
    x[:, col] = rng.random(n) < 0.4
    return x


def molecule_like_adjacency(rng: np.random.Generator, n: int):
    """Random spanning tree plus a few ring closures, degree <= 4."""
    deg = np.zeros(n, np.int32)
    adj = [[] for _ in range(n)]

    def link(a, b):
        if a != b and b not in adj[a] and deg[a] < 4 and deg[b] < 4:
            adj[a].append(b), adj[b].append(a)
            deg[a] += 1; deg[b] += 1
            return True
        return False

    for v in range(1, n):
        for _ in range(16):
            if link(v, int(rng.integers(max(0, v - 6), v))):
                break
        else:
            cands = [u for u in range(v) if deg[u] < 4]
            link(v, cands[-1])
    for _ in range(max(1, n // 6)):
        a = int(rng.integers(0, n)); b = int(min(n - 1, a + rng.integers(3, 7)))
        link(a, b)
    return adj


def tox21_like_sizes(rng: np.random.Generator, B: int, Nmax: int = 132) -> np.ndarray:
    n = np.clip(np.round(rng.lognormal(np.log(17.0), 0.55, B)), 4, Nmax).astype(np.int32)
    n[0] = Nmax                                            # one forced maximum-size molecule
    return n


def synthetic_molecule_batch(B: int, Nmax: int = 132, seed: int = 1234):
    """C1/C2 of SURVEY 8(d): returns X [B,Nmax,75] f32, L [B,Nmax,Nmax] f32 (padded
    like models/tf_modules/graph_topology.py:84-98) and n_nodes [B] int32."""
    rng = np.random.default_rng(seed)
    n_nodes = tox21_like_sizes(rng, B, Nmax)
    X = np.zeros((B, Nmax, 75), np.float32)
    L = np.zeros((B, Nmax, Nmax), np.float32)
    for g, n in enumerate(n_nodes):
        X[g, :n] = tox21_like_features(rng, n)
        L[g, :n, :n] = compute_laplacian_dense(molecule_like_adjacency(rng, n)).astype(np.float32)
    return X, L, n_nodes


def rel_err(a, b) -> float:
    """max |a-b| / max(|b|, tiny): the relative error the 1e-4 parity budget is
    stated in."""
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))
