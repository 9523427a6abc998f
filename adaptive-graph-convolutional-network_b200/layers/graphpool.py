"""GraphPoolMol -- neighbourhood max pooling over the Laplacian's sparsity pattern, B200 build.

Mirrors models/layers/graphpool.py of the reference: ``GraphPoolMol(batch_size, **kwargs)`` and the
``call(dict)`` contract of the graph layers (keys ``node_features``, ``original_laplacian``, ``data_slice``,
``lap_slice``), returning the pooled node features (graphpool.py:38-53).  The reference pools inside
``tf.py_func`` (graphpool.py:91-105: a Python loop over atoms, no gradient); here one kernel
(``agcn_graph_pool``, include/agcn_sgcll.h) pools the whole packed batch.

``pool_grad`` in {"reference", "argmax"}: "reference" (default) stops the gradient like the reference's
py_func; "argmax" routes it to the node that supplied each maximum.
"""
import torch

from .. import _lib
from ..batch import PackedNodes, _ptr, _stream_ptr
from .basic_layer import Layer
from .graphconv import SGC_LL


class _GraphPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, L, batch, with_grad):
        X = X.contiguous()
        R, F = X.shape
        assert R == batch.total_nodes and L.numel() == batch.total_lap and X.is_cuda
        Y = torch.empty_like(X)
        arg = torch.empty(R, F, dtype=torch.int32, device=X.device) if with_grad else None
        with torch.cuda.device(X.device):
            _lib.check(_lib.lib().agcn_graph_pool(batch.handle, _ptr(X), _ptr(L.contiguous()), _ptr(Y), _ptr(arg), F,
                                                  _stream_ptr(X.device)))
        ctx.batch, ctx.F, ctx.with_grad = batch, F, with_grad
        if with_grad:
            ctx.save_for_backward(arg)
        return Y

    @staticmethod
    def backward(ctx, dY):
        if not ctx.with_grad:
            return None, None, None, None      # tf.py_func has no gradient (graphpool.py:107)
        arg, = ctx.saved_tensors
        dX = torch.empty_like(dY)
        with torch.cuda.device(dY.device):
            _lib.check(_lib.lib().agcn_graph_pool_backward(ctx.batch.handle, _ptr(dY.contiguous()), _ptr(arg), _ptr(dX),
                                                           ctx.F, _stream_ptr(dY.device)))
        return dX, None, None, None


def graph_pool_packed(X, L, batch, pool_grad="reference"):
    """X [R,F] packed nodes, L packed Laplacians -> pooled [R,F]."""
    if pool_grad not in ("reference", "argmax"):
        raise ValueError("pool_grad must be 'reference' or 'argmax'")
    return _GraphPool.apply(X, L, batch, pool_grad == "argmax")


class GraphPoolMol(Layer):
    """GraphPoolMol(batch_size, **kwargs)                                   graphpool.py:14-27"""

    def __init__(self, batch_size, **kwargs):
        self.pool_grad = kwargs.pop('pool_grad', 'reference')
        super(GraphPoolMol, self).__init__(**kwargs)
        self.batch_size = batch_size
        self.sparse_inputs = True

    def get_output_shape_for(self, input_shape):
        return input_shape[0]

    def call(self, x, mask=None):
        """graphpool.py:38-53."""
        node_features = x['node_features']
        batch = SGC_LL._resolve_batch(x, node_features)
        X = SGC_LL._packed_nodes(node_features, batch)
        L = SGC_LL._packed_laps(x['original_laplacian'], batch, x, '_packed_laplacian')
        return PackedNodes(graph_pool_packed(X, L, batch, self.pool_grad), batch)
