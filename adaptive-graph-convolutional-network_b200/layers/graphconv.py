"""SGC_LL -- Spectral Graph Convolution with Laplacian Learning, B200 build.

Mirrors models/layers/graphconv.py of the reference: same constructor, same ``call(dict)`` contract
(keys ``node_features``, ``original_laplacian``, ``data_slice``, ``lap_slice``), same return tuple
``(activated_nodes, res_L, res_W)``.  The per-graph Python/TensorFlow loop (graphconv.py:145-251) is
replaced by batch-wide CUDA kernels behind the C ABI of include/agcn_sgcll.h; there is no CPU path.

Two things the reference code does not say in its signature are explicit here:
``laplacian`` in {"reference_literal", "paper"} and ``metric_grad`` in {"reference", "full"};
the defaults reproduce the reference as written.
"""
import numpy as np
import torch

from ..batch import GraphBatch, PackedLaplacians, PackedNodes
from ..functional import pack_nodes, sgc_ll_packed
from ..operators import activations
from .basic_layer import Layer
from .dropout import Dropout

# process-wide default semantics; a layer can override them with the `laplacian=` / `metric_grad=`
# keyword arguments or by assigning the attributes before the first call.
SEMANTICS = {"laplacian": "reference_literal", "metric_grad": "reference"}
DEFAULT_DEVICE = ["cuda"]


def _device():
    return torch.device(DEFAULT_DEVICE[0])


def truncate_normal(shape, stddev=1e-3, name=None):
    """graphconv.py:14-16: truncated normal (|z| <= 2 sigma), fp32 variable."""
    t = torch.empty(*shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(t, mean=0.0, std=stddev, a=-2 * stddev, b=2 * stddev)
    return t.to(_device()).requires_grad_(True)


def glorot(shape, name=None):
    """graphconv.py:19-23: Glorot & Bengio uniform init."""
    init_range = np.sqrt(6.0 / (shape[0] + shape[1]))
    t = (torch.rand(*shape, dtype=torch.float32) * 2 - 1) * init_range
    return t.to(_device()).requires_grad_(True)


def zeros(shape, name=None):
    """graphconv.py:26-29."""
    return torch.zeros(*shape, dtype=torch.float32, device=_device()).requires_grad_(True)


class LazyLaplacians(object):
    """res_L / res_W lists computed only when somebody reads them."""

    def __init__(self, compute, key, batch):
        self._compute, self._key, self.batch = compute, key, batch

    def _packed(self):
        return self._compute()[self._key]

    def __len__(self):
        return self.batch.batch_size

    def __getitem__(self, g):
        return PackedLaplacians(self._packed(), self.batch)[g]

    def __iter__(self):
        return iter(PackedLaplacians(self._packed(), self.batch))


class SGC_LL(Layer):
    """SGC_LL(output_dim, input_dim, batch_size, activation='relu', dropout=None, K=2,
    save_lap=False, save_output=False, **kwargs)            graphconv.py:40-64"""

    variant = "SGC_LL"

    def __init__(self, output_dim, input_dim, batch_size, activation='relu', dropout=None, K=2, save_lap=False,
                 save_output=False, **kwargs):
        self.laplacian = kwargs.pop('laplacian', None)
        self.metric_grad = kwargs.pop('metric_grad', None)
        super(SGC_LL, self).__init__(**kwargs)
        self.dropout = dropout
        self.activation = activations.get(activation)           # callables pass through (activations.py:15-53)
        # the kernel epilogue fuses relu / linear; every other activation (by name or as a callable) is applied
        # by the host on the kernel's linear output (graphconv.py:120)
        if self.activation is activations.relu:
            self.activation_name = 'relu'
        elif self.activation is activations.linear:
            self.activation_name = 'linear'
        else:
            self.activation_name = None
        self.batch_size = batch_size
        self.nb_filter = output_dim
        self.n_atom_feature = input_dim
        self.vars = {}
        self.bias = True
        self.K = K
        self.save_lap = save_lap
        self.early_laps = None
        self.save_output = save_output

    # -------------------------------------------------------------------------------------------
    def build(self):
        """graphconv.py:66-83.  The reference re-creates the variables on every call(); here they
        are created once and reused."""
        if self.vars:
            return
        self.vars['weight'] = glorot([self.n_atom_feature * self.K, self.nb_filter], name='weights_feature')
        if self.bias:
            self.vars['bias'] = zeros([self.nb_filter], name='bias')
        self.vars['M_L'] = glorot([self.n_atom_feature, self.n_atom_feature], name='Maha_dist')
        self.vars['alpha'] = torch.ones(1, dtype=torch.float32, device=_device()).requires_grad_(True)

    def _semantics(self):
        return (self.laplacian or SEMANTICS["laplacian"], self.metric_grad or SEMANTICS["metric_grad"])

    # -------------------------------------------------------------------------------------------
    @staticmethod
    def _resolve_batch(x, node_features):
        batch = x.get('_batch')
        if batch is None and isinstance(node_features, PackedNodes):
            batch = node_features.batch
        if batch is None:
            ds = x['data_slice']
            ds = ds.detach().cpu().numpy() if isinstance(ds, torch.Tensor) else np.asarray(ds)
            first = node_features[0]
            batch = GraphBatch(ds[:, 0], int(first.shape[0]), device=first.device)
            x['_batch'] = batch  # later layers of the same step reuse the plan
        return batch

    @staticmethod
    def _packed_nodes(node_features, batch):
        if isinstance(node_features, PackedNodes):
            return node_features.data
        if isinstance(node_features, (list, tuple)):
            node_features = torch.stack(list(node_features), 0)  # list of B [max_atom, F]
        return pack_nodes(node_features, batch)

    @staticmethod
    def _packed_laps(laps, batch, x=None, cache_key=None):
        if isinstance(laps, PackedLaplacians):
            return laps.data
        if x is not None and cache_key in x:
            return x[cache_key]
        if isinstance(laps, (list, tuple)):
            if len(laps) == 0:
                return None
            if laps[0].shape[0] == batch.max_atom and all(l.shape == laps[0].shape for l in laps):
                packed = batch.pack_lap(torch.stack(list(laps), 0))
            else:  # list of unpadded [n_g, n_g]
                packed = torch.cat([l.reshape(-1) for l in laps])
        elif laps.dim() == 3:
            packed = batch.pack_lap(laps)
        else:
            packed = laps
        if x is not None and cache_key is not None:
            x[cache_key] = packed
        return packed

    def _cfg(self, fused_act):
        lap, mg = self._semantics()
        return {"F": self.n_atom_feature, "Fo": self.nb_filter, "K": self.K, "variant": self.variant,
                "laplacian": lap, "metric_grad": mg, "activation": fused_act}

    def _fused_activation(self):
        """(fused?, activation the kernel applies).  Not fused: the kernel runs linear and _finish applies
        self.activation, whatever it is (a registry name or a caller's own callable)."""
        if self.activation_name in ('relu', 'linear'):
            return True, self.activation_name
        return False, 'linear'

    def _finish(self, Y, fused):
        if not fused:
            Y = self.activation(Y)                      # graphconv.py:120
        if self.dropout is not None:
            Y = Dropout(self.dropout)(Y)                # graphconv.py:121-122
        return Y

    def call(self, x):
        """graphconv.py:85-125."""
        self.build()
        node_features = x['node_features']
        batch = self._resolve_batch(x, node_features)
        X = self._packed_nodes(node_features, batch)
        Lint = self._packed_laps(x['original_laplacian'], batch, x, '_packed_laplacian')
        fused, fused_act = self._fused_activation()
        cfg = self._cfg(fused_act)
        Y, _, _, _ = sgc_ll_packed(X, Lint, None, self.vars, batch, cfg)
        Y = self._finish(Y, fused)

        cache = {}
        # evaluated with the parameter values at the time of the first read (read them before the
        # optimizer step, like the reference's sess.run fetches them together with the train op)
        Xd, params = X.detach(), {k: v.detach() for k, v in self.vars.items()}

        def lazy():
            if not cache:
                with torch.no_grad():
                    c2 = dict(cfg, want_resL=True, want_resW=True)
                    _, rl, rw, _ = sgc_ll_packed(Xd, Lint, None, params, batch, c2)
                cache['res_L'], cache['res_W'] = rl, rw
            return cache

        return PackedNodes(Y, batch), LazyLaplacians(lazy, 'res_L', batch), LazyLaplacians(lazy, 'res_W', batch)
