"""BlockEnd -- last layer of a residual block: act(x_res W + x).  Mirrors models/layers/blockend.py (constructor
:25-40, build :42-44, call :46-65, add_residuals :67-86); the reference's per-graph loop of slice / matmul / add /
pad is ONE product over the packed rows (agcn_node_gemm, include/agcn_sgcll.h) with the residual add and the
activation in its epilogue.  Padding rows stay +0.0: they are not stored in the packed layout."""
from ..batch import PackedNodes
from ..functional import node_linear
from ..operators import activations
from .basic_layer import Layer
from .graphconv import SGC_LL, glorot


def fused_activation(fn):
    """(name the kernel epilogue applies, host function applied afterwards or None)."""
    if fn is activations.relu:
        return "relu", None
    if fn is activations.linear:
        return "linear", None
    return "linear", fn


class BlockEnd(Layer):
    def __init__(self, block_id, res_n_features, n_features, activation='relu', max_atom=128, batch_size=256, **kwargs):
        super(BlockEnd, self).__init__(**kwargs)
        self.max_atom = max_atom
        self.batch_size = batch_size
        self.block_id = block_id
        self.activation = activations.get(activation)
        self.res_n_features = res_n_features
        self.n_features = n_features
        self.vars = {}

    def build(self):
        if not self.vars:
            self.vars['weight'] = glorot([self.res_n_features, self.n_features], name='trans_feature')

    def call(self, x):
        """blockend.py:46-65: keys node_features, data_slice, block_outputs (the saved output of the last block)."""
        self.build()
        node_features = x['node_features']
        batch = SGC_LL._resolve_batch(x, node_features)
        X = SGC_LL._packed_nodes(node_features, batch)
        res = x['block_outputs']
        assert len(res) == len(node_features)                      # blockend.py:68
        Xres = SGC_LL._packed_nodes(res, batch)
        act, host_act = fused_activation(self.activation)
        Y = node_linear(Xres, self.vars['weight'], add=X, activation=act)       # x_res W + x   (:78)
        if host_act is not None:
            Y = host_act(Y)
        return PackedNodes(Y, batch)
