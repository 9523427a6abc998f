"""SGC_LL_Reslap: models/layers/graphconv_reslap.py -- SGC_LL plus the previous layer's Laplacian
(`res_lap`, weighted by beta) and a second norm clip + leaky rectifier."""
import torch

from ..batch import PackedLaplacians, PackedNodes
from ..functional import sgc_ll_packed
from .graphconv import SGC_LL, LazyLaplacians, _device


class SGC_LL_Reslap(SGC_LL):
    variant = "SGC_LL_Reslap"

    def __init__(self, *args, **kwargs):
        super(SGC_LL_Reslap, self).__init__(*args, **kwargs)
        self.early_laps = None

    def build(self):
        """graphconv_reslap.py:21-43."""
        if self.vars:
            return
        super(SGC_LL_Reslap, self).build()
        self.vars['beta'] = torch.ones(1, dtype=torch.float32, device=_device()).requires_grad_(True)

    def call(self, x):
        """graphconv_reslap.py:45-89: returns (activated_nodes, res_L, res_W, L_all)."""
        self.build()
        node_features = x['node_features']
        batch = self._resolve_batch(x, node_features)
        X = self._packed_nodes(node_features, batch)
        Lint = self._packed_laps(x['original_laplacian'], batch, x, '_packed_laplacian')
        self.early_laps = x['res_lap']
        Lprev = None
        if self.early_laps is not None and len(self.early_laps) > 0:     # graphconv_reslap.py:188
            Lprev = self._packed_laps(self.early_laps, batch)
        fused, fused_act = self._fused_activation()
        cfg = self._cfg(fused_act)
        Y, _, _, Lall = sgc_ll_packed(X, Lint, Lprev, self.vars, batch, cfg)
        Y = self._finish(Y, fused)

        cache = {}
        # evaluated with the parameter values at the time of the first read (read them before the
        # optimizer step, like the reference's sess.run fetches them together with the train op)
        Xd, params = X.detach(), {k: v.detach() for k, v in self.vars.items()}
        Lpd = None if Lprev is None else Lprev.detach()

        def lazy():
            if not cache:
                with torch.no_grad():
                    c2 = dict(cfg, want_resL=True, want_resW=True)
                    _, rl, rw, _ = sgc_ll_packed(Xd, Lint, Lpd, params, batch, c2)
                cache['res_L'], cache['res_W'] = rl, rw
            return cache

        return (PackedNodes(Y, batch), LazyLaplacians(lazy, 'res_L', batch), LazyLaplacians(lazy, 'res_W', batch),
                PackedLaplacians(Lall, batch))
