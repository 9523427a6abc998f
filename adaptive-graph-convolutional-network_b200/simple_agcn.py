"""SimpleAGCN-shaped training step used to MEASURE the SGC-LL hot path (bench.py).

The reference assembles the same stack in models/networks/basic_AGCN.py:35-47: four SGC_LL layers
(l_n_filters = [64, 128, 128, 64], utils/hyper_parameters.py:14), DenseMol, GraphGatherMol(tanh),
then MultitaskGraphClassifier's per-task 2-class logits with weighted sigmoid cross-entropy divided
by the batch size (models/tf_modules/multitask_classifier.py:41-44,187-209) and Adam
(:233-237).  Only the SGC_LL layers are this repository's product; the dense / gather / head / Adam
pieces are plain torch ops here (SURVEY.md section 8f lists them as the next rows to fuse).

Data parallel: every rank holds a replica and its own shard of graphs; one all-reduce of a flat
gradient buffer per step.  The loss is normalised by the GLOBAL batch size.
"""
import numpy as np
import torch
import torch.distributed as dist

from .batch import GraphBatch, PackedLaplacians, PackedNodes
from .data_parallel import FlatGradBuffer, FlatParamBuffer
from .layers import SGC_LL
from .layers import graphconv as _gc


class SimpleAGCNStep(object):
    def __init__(self, n_feat=75, filters=(64, 128, 128, 64), final_feature_n=256, n_tasks=12, K=3, batch_size=256,
                 learning_rate=2e-3, device="cuda", world_size=1, laplacian="reference_literal",
                 metric_grad="reference", seed=123, fused_head=True):
        self.device = torch.device(device)
        self.world_size = world_size
        self.global_batch = batch_size * world_size
        self.n_tasks = n_tasks
        # DenseMol + GraphGatherMol + heads + loss as ONE library call (agcn_head_loss_grad) instead of ~45 torch ops
        self.fused_head = fused_head and filters[-1] % 4 == 0 and 32 <= filters[-1] <= 128 and \
            final_feature_n % 4 == 0 and 32 <= final_feature_n <= 256 and 2 * n_tasks >= 32
        torch.manual_seed(seed)  # identical replicas on every rank
        _gc.DEFAULT_DEVICE[0] = str(self.device)
        dims = [n_feat] + list(filters)
        self.layers = [SGC_LL(dims[i + 1], dims[i], batch_size, K=K, activation='relu', laplacian=laplacian,
                              metric_grad=metric_grad) for i in range(len(filters))]
        for l in self.layers:
            l.build()
        lim = float(np.sqrt(6.0 / (filters[-1] + final_feature_n)))
        self.dense_W = ((torch.rand(filters[-1], final_feature_n) * 2 - 1) * lim).to(self.device).requires_grad_(True)
        self.dense_b = torch.zeros(final_feature_n, device=self.device, requires_grad=True)
        # n_tasks independent [n_feature, 2] logits heads == one [n_feature, 2 * n_tasks] matrix
        self.head_W = torch.nn.init.trunc_normal_(torch.empty(final_feature_n, 2 * n_tasks), std=0.01, a=-0.02,
                                                  b=0.02).to(self.device).requires_grad_(True)
        self.head_b = torch.zeros(2 * n_tasks, device=self.device, requires_grad=True)
        self.params = [v for l in self.layers for v in l.vars.values()] + [self.dense_W, self.dense_b, self.head_W,
                                                                           self.head_b]
        # one flat gradient buffer; every .grad is a view into it -> a single all-reduce per step
        # the SGC-LL backward writes its parameter gradients straight into the views (no accumulation kernels)
        direct = [v for l in self.layers for v in l.vars.values()]
        if self.fused_head:
            direct += [self.dense_W, self.dense_b, self.head_W, self.head_b]
        self.grads = FlatGradBuffer(self.params, direct=direct)
        self.flat_grad = self.grads.flat
        # ... and one flat parameter tensor: Adam is a single-tensor update
        self.flat_params = FlatParamBuffer(self.params, self.grads)
        self.opt = torch.optim.Adam(self.flat_params.chunks(self.grads), lr=learning_rate, betas=(0.9, 0.999), eps=1e-7,
                                    fused=True, capturable=True)

    def _n_nodes_f(self, batch):
        t = getattr(batch, "_n_nodes_float", None)
        if t is None:
            t = batch.n_nodes_device().float()
            batch._n_nodes_float = t
        return t

    def n_parameters(self):
        return int(self.flat_grad.numel())

    def forward_loss(self, X, Lint, batch, onehot, weights):
        """X [R, F] packed, Lint packed, onehot [B, 2*T] float, weights [B, 2*T] float -> scalar loss."""
        x = {'node_features': PackedNodes(X, batch), 'original_laplacian': PackedLaplacians(Lint, batch),
             'data_slice': None, 'lap_slice': None, '_batch': batch}
        for layer in self.layers:
            out, _, _ = layer(x)
            x = dict(x, node_features=out)
        if self.fused_head:
            from .functional import head_loss
            return head_loss(out.data, self.dense_W, self.dense_b, self.head_W, self.head_b, onehot, weights, batch,
                             1.0 / self.global_batch, unit_grad=True)
        # DenseMol applies no activation (dense_layer.py:42-50), so the per-graph row sum of GraphGatherMol
        # (graphgather.py:68-77) commutes with it: sum_i (h_i W + b) = (sum_i h_i) W + n_g b.  Gathering first
        # shrinks the dense GEMM from R = sum n_g rows to B rows.
        hsum = torch.zeros(batch.batch_size, out.data.shape[1], device=out.data.device).index_add_(
            0, batch.graph_ids(), out.data)
        mol = torch.addmm(self._n_nodes_f(batch)[:, None] * self.dense_b[None, :], hsum, self.dense_W)
        mol = torch.tanh(mol)                                                  # GraphGatherMol
        logits = torch.addmm(self.head_b, mol, self.head_W)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, onehot, weight=weights, reduction='sum')
        return loss / self.global_batch

    def step(self, X, Lint, batch, onehot, weights):
        """One training step: forward, backward, gradient all-reduce, Adam.  Returns the loss tensor."""
        self.grads.zero()
        loss = self.forward_loss(X, Lint, batch, onehot, weights)
        loss.backward()
        if self.world_size > 1:
            self.grads.all_reduce()
        self.opt.step()
        return loss


def synthetic_labels(B, n_tasks, seed, device):
    """Bernoulli(0.1) labels as one-hot float targets [B, 2T], weights 1."""
    rng = np.random.default_rng(seed)
    y = (rng.random((B, n_tasks)) < 0.1)
    onehot = np.zeros((B, n_tasks, 2), np.float32)
    onehot[..., 0] = ~y
    onehot[..., 1] = y
    w = np.ones((B, 2 * n_tasks), np.float32)
    return torch.from_numpy(onehot.reshape(B, 2 * n_tasks)).to(device), torch.from_numpy(w).to(device)
