"""DenseBlockEnd -- end of a densely connected block: act(x + beta1 sum_l a_l W_l + beta2 sum_b o_b W_b).  Mirrors
models/layers/densenet_block.py (constructor :21-47, build :49-70, call :72-96, add_l_residuals :98-131).  Every
`x_res W` of the reference's per-graph loop is one product over the packed rows (agcn_node_gemm) that accumulates
into the running sum with the learned beta as a device scalar; the last one carries the activation."""
import torch

from ..batch import PackedNodes
from ..functional import node_linear
from ..operators import activations
from .basic_layer import Layer
from .blockend import fused_activation
from .graphconv import SGC_LL, _device, glorot


def _split_saved(saved, n_groups, batch):
    """The reference keeps saved activations as ONE flat list: entry layer_id * batch_size + graph_id
    (densenet_block.py:118,124).  Accepts that list, a list of PackedNodes, or a list of per-layer lists."""
    if n_groups == 0:
        return []
    if len(saved) == n_groups and all(isinstance(s, (PackedNodes, list, tuple)) or hasattr(s, 'dim') and s.dim() == 3
                                      for s in saved):
        groups = list(saved)
    else:
        B = batch.batch_size
        assert len(saved) == n_groups * B                          # densenet_block.py:109
        groups = [saved[i * B:(i + 1) * B] for i in range(n_groups)]
    return [SGC_LL._packed_nodes(g, batch) for g in groups]


class DenseBlockEnd(Layer):
    def __init__(self, block_id, res_n_features_list, output_n_features, activation='relu', K=2, max_atom=128,
                 batch_size=256, **kwargs):
        super(DenseBlockEnd, self).__init__(**kwargs)
        self.max_atom = max_atom
        self.batch_size = batch_size
        self.block_id = block_id
        self.activation = activations.get(activation)
        self.K = K
        assert type(res_n_features_list) == list
        self.res_n_features_list = res_n_features_list
        self.output_n_features = output_n_features
        self.vars = {}
        self.inblock_activations = []
        self.inblock_activations_dim = []
        self.preceding_blocks = []
        self.preceding_blocks_dim = []

    def build(self):
        """densenet_block.py:49-70 (the reference re-creates the variables on every call; here once)."""
        if self.vars:
            return
        self.vars['weight_outblock'] = [glorot([f, self.output_n_features], name='trans_outblock_feature_%d' % i)
                                        for i, f in enumerate(self.preceding_blocks_dim)]
        self.vars['weight_inblock'] = [glorot([f, self.output_n_features], name='trans_inblock_feature_%d' % i)
                                       for i, f in enumerate(self.inblock_activations_dim)]
        self.vars['beta_inblock'] = torch.ones(1, dtype=torch.float32, device=_device()).requires_grad_(True)
        self.vars['beta_outblock'] = torch.ones(1, dtype=torch.float32, device=_device()).requires_grad_(True)

    def parameters(self):
        self.build()
        return (self.vars['weight_outblock'] + self.vars['weight_inblock'] +
                [self.vars['beta_inblock'], self.vars['beta_outblock']])

    def call(self, x):
        """densenet_block.py:72-96."""
        node_features = x['node_features']
        self.inblock_activations = x['inblock_activations']
        self.inblock_activations_dim = list(x['inblock_activations_dim'])
        self.preceding_blocks = x['block_outputs']
        self.preceding_blocks_dim = list(x['block_outputs_dim'])
        self.build()
        batch = SGC_LL._resolve_batch(x, node_features)
        Y = SGC_LL._packed_nodes(node_features, batch)
        terms = [(a, w, self.vars['beta_inblock']) for a, w in
                 zip(_split_saved(self.inblock_activations, len(self.inblock_activations_dim), batch),
                     self.vars['weight_inblock'])]
        terms += [(o, w, self.vars['beta_outblock']) for o, w in
                  zip(_split_saved(self.preceding_blocks, len(self.preceding_blocks_dim), batch),
                      self.vars['weight_outblock'])]
        act, host_act = fused_activation(self.activation)
        for i, (a, w, beta) in enumerate(terms):                   # x += (x_res W) * beta   (:119-126)
            last = i + 1 == len(terms)
            Y = node_linear(a, w, scale=beta, add=Y, activation=act if last else "linear")
        if not terms and act == "relu":
            Y = torch.relu(Y)
        if host_act is not None:
            Y = host_act(Y)
        return PackedNodes(Y, batch)
