"""CPU oracle of the callers around the SGC-LL layer.  TEST INFRASTRUCTURE ONLY (same rules as
sgcll_oracle.py: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product).

Restates, with torch ops on the CPU (any dtype, autograd for the gradients):
  * DenseMol                   models/layers/dense_layer.py:33-50       per graph X W + b, NO activation applied
  * GraphGatherMol             models/layers/graphgather.py:40-78       row sum over the real atoms, then tanh
  * multitask_logits           models/operators/model_operatos.py:792-864   n_tasks independent [n_feature, 2] heads
  * weighted sigmoid CE loss   models/tf_modules/multitask_classifier.py:41-44,187-209   sum / batch_size
  * the SimpleAGCN stack       models/networks/basic_AGCN.py:35-47      4 x SGC_LL(relu) + DenseMol + GraphGatherMol
  * tf.train.AdamOptimizer     models/tf_modules/multitask_classifier.py:233-237 (TF's update rule, restated)
  * BlockEnd / DenseBlockEnd / MLP   models/layers/blockend.py:67-86, densenet_block.py:98-131, MLP.py:69-83

Parity pin status: these are TensorFlow graph ops (matmul, reduce_sum, tanh, sigmoid_cross_entropy_with_logits,
AdamOptimizer) restated from TF's documented semantics -- "parity unpinned" like the TF part of sgcll_oracle.py
(TensorFlow 0.12 / Python 2 cannot run here and the reference ships no golden vectors).
"""
from __future__ import annotations

import numpy as np
import torch

from . import sgcll_oracle as O


def sigmoid_cross_entropy_with_logits(x, t):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x, 0) - x t + log(1 + exp(-|x|))."""
    return torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))


def head_loss(H_list, dense_W, dense_b, head_W, head_b, targets, weights, scale):
    """H_list: B tensors [n_g, Fh] (the real rows of the last SGC-LL layer's output).
    head_W [Fm, 2 T] is the n_tasks independent [Fm, 2] heads side by side (column 2 t + c = class c of task t),
    targets / weights [B, 2 T] (one-hot labels; the reference's per-task weight repeated for both classes).
    Returns the scalar loss  scale * sum_b,t,c w * sigmoid_ce(logit, target)."""
    mols = []
    for h in H_list:
        d = h @ dense_W + dense_b                        # dense_layer.py:42-49 (bias per atom, no activation)
        mols.append(d.sum(0))                            # graphgather.py:68-74
    mol = torch.tanh(torch.stack(mols, 0))               # graphgather.py:53 activation="tanh"
    logits = mol @ head_W + head_b                       # model_operatos.py:792-864
    costs = sigmoid_cross_entropy_with_logits(logits, targets) * weights     # multitask_classifier.py:41-44
    return costs.sum() * scale                           # :203-208 (scale = 1 / batch_size)


def simple_agcn_features(X, L, n_nodes, layer_params, K, laplacian="reference_literal", metric_grad="reference",
                         relu_masks=None):
    """The SGC_LL stack of basic_AGCN.py:35-45 over a padded batch -> list of B [n_g, Fo] outputs of the last layer.

    relu_masks (optional): per layer a bool tensor [R, Fo_l] over the packed rows (graph after graph) saying which
    outputs the implementation under test found positive.  The gradient of a ReLU network is discontinuous where a
    pre-activation crosses zero: a value within rounding error of 0 can land on either side in two correct
    implementations, and that single flip moves a weight gradient by ~1/sqrt(R) of its size.  With the masks given,
    relu(y) is evaluated as y * mask, so both sides differentiate the SAME piecewise-linear branch (the forward
    values differ by the rounding error of those near-zero entries only)."""
    H, row = [], 0
    skip = laplacian == "reference_literal"
    for g in range(X.shape[0]):
        n = int(n_nodes[g])
        x, Lg = X[g, :n], L[g, :n, :n]
        for l, p in enumerate(layer_params):
            y, _, _, _ = O.sgc_ll_graph(x, Lg, p, K, "SGC_LL", laplacian, metric_grad, compute_similarity=not skip)
            if relu_masks is None:
                x = torch.relu(y)                        # graphconv.py:118-123
            else:
                x = y * relu_masks[l][row:row + n].to(y.dtype)
        H.append(x)
        row += n
    return H


def simple_agcn_loss(X, L, n_nodes, layer_params, head_params, targets, weights, global_batch, K,
                     laplacian="reference_literal", metric_grad="reference", relu_masks=None):
    """basic_AGCN.py:35-47 over a padded batch: X [B,Nmax,F], L [B,Nmax,Nmax] tensors, n_nodes [B].
    layer_params: list of dicts (weight, bias, M_L, alpha); head_params: dict dense_W, dense_b, head_W, head_b."""
    H = simple_agcn_features(X, L, n_nodes, layer_params, K, laplacian, metric_grad, relu_masks)
    return head_loss(H, head_params["dense_W"], head_params["dense_b"], head_params["head_W"], head_params["head_b"],
                     targets, weights, 1.0 / global_batch)


def adam_tf(param, grad, m, v, t, lr=2e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer's update (t = 1 for the first step):
        lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t);  m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2
        param -= lr_t m / (sqrt(v) + eps)
    Returns (param, m, v) as new tensors."""
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = beta1 * m + (1.0 - beta1) * grad
    v = beta2 * v + (1.0 - beta2) * grad * grad
    return param - lr_t * m / (torch.sqrt(v) + eps), m, v


# ---- residual / dense-connected block ends and the MLP (per graph, real rows only; rows >= n_g stay 0)
def block_end(x, x_res, weight, activation=torch.relu):
    """blockend.py:67-86: act(x_res W + x)."""
    return activation(x_res @ weight + x)


def dense_block_end(x, inblock, w_inblock, outblock, w_outblock, beta_in, beta_out, activation=torch.relu):
    """densenet_block.py:98-131: act(x + sum_l beta1 (a_l W_l) + sum_b beta2 (o_b W_b))."""
    for a, w in zip(inblock, w_inblock):
        x = x + (a @ w) * beta_in
    for o, w in zip(outblock, w_outblock):
        x = x + (o @ w) * beta_out
    return activation(x)


def mlp(x, weights, biases):
    """MLP.py:69-83: x = x W_i + b_i for every layer (no activation in between), relu AFTER the zero padding."""
    for w, b in zip(weights, biases):
        x = x @ w + b
    return torch.relu(x)
