"""MLP -- multi-layer perceptron on the node features: x W_0 + b_0, ..., x W_k + b_k (no activation in between), relu
after the zero padding.  Mirrors models/layers/MLP.py (constructor :19-41, build :43-56, call :58-67, MLP :69-83).
Each `tf.matmul(x, W) + b` of the per-graph loop is one product over the packed rows (agcn_node_gemm); the relu is
the epilogue of the last one."""
import torch

from ..batch import PackedNodes
from ..functional import node_linear
from ..operators import activations
from .basic_layer import Layer
from .graphconv import SGC_LL, _device, glorot


class MLP(Layer):
    def __init__(self, output_dim, hidden_dims, input_dim, batch_size, init='glorot_uniform', activation="relu",
                 bias=True, max_atom=128, **kwargs):
        super(MLP, self).__init__(**kwargs)
        assert type(hidden_dims) == list
        if init != 'glorot_uniform':
            raise ValueError('Invalid initialization: ' + str(init))      # the only one the reference's build() uses
        self.init = init
        self.activation = activations.get(activation)
        self.output_dim = output_dim
        self.input_dim = input_dim
        self.bias = bias
        self.max_atom = max_atom
        self.batch_size = batch_size
        self.hidden_dims = hidden_dims
        self.vars = {}

    def build(self):
        if self.vars:
            return
        dims = [self.input_dim] + self.hidden_dims + [self.output_dim]
        self.vars['weight'] = [glorot([a, b], name='trans_feature_%d' % i)
                               for i, (a, b) in enumerate(zip(dims[:-1], dims[1:]))]
        self.vars['bias'] = [torch.zeros(b, dtype=torch.float32, device=_device()).requires_grad_(True)
                             for b in dims[1:]]

    def parameters(self):
        self.build()
        return self.vars['weight'] + self.vars['bias']

    def call(self, x):
        self.build()
        node_features = x['node_features']
        batch = SGC_LL._resolve_batch(x, node_features)
        X = SGC_LL._packed_nodes(node_features, batch)
        n = len(self.vars['weight'])
        for i, (w, b) in enumerate(zip(self.vars['weight'], self.vars['bias'])):
            X = node_linear(X, w, bias=b, activation="relu" if i + 1 == n else "linear")    # MLP.py:76-82
        return PackedNodes(X, batch)
