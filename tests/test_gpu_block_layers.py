"""The layers around SGC-LL that are per-graph matmuls in the reference and Dropout
(section 8a), through the C ABI (agcn_node_gemm / agcn_gemm_tn / agcn_dropout) against oracle/network_oracle.py."""
import numpy as np
import pytest
import torch

from oracle import network_oracle as NO
from oracle import sgcll_oracle as O
from util_cases import TOL

pytestmark = pytest.mark.gpu

SIZES = [132, 4, 5, 18, 33, 64, 65, 17, 96, 31]


def _padded(n, F, Nmax, seed, relu=True):
    rng = np.random.default_rng(seed)
    X = np.zeros((len(n), Nmax, F), np.float32)
    for g, k in enumerate(n):
        v = rng.standard_normal((k, F)).astype(np.float32) * 0.5
        X[g, :k] = np.maximum(v, 0) if relu else v
    return X


def _x(batch, dev, **tensors):
    n = batch.n_nodes
    d = {"data_slice": np.stack([[k, -1] for k in n]).astype(np.int32), "_batch": batch}
    d.update(tensors)
    return d


def _rows(Xp, n):
    return [torch.tensor(Xp[g, :k], dtype=torch.float64, requires_grad=True) for g, k in enumerate(n)]


@pytest.mark.parametrize("Fres,F", [(64, 128), (75, 64), (128, 24)])
def test_block_end_matches_oracle(Fres, F):
    import agcn_b200
    from agcn_b200.layers import BlockEnd
    dev = torch.device("cuda:0")
    n = np.asarray(SIZES, np.int32)
    batch = agcn_b200.GraphBatch(n, 132, device=dev)
    Xp, Rp = _padded(n, F, 132, 1, relu=False), _padded(n, Fres, 132, 2)
    Xd = torch.tensor(Xp, device=dev, requires_grad=True)
    Rd = torch.tensor(Rp, device=dev, requires_grad=True)
    layer = BlockEnd(0, Fres, F, 'relu', max_atom=132, batch_size=len(n))
    out = layer(_x(batch, dev, node_features=Xd, block_outputs=Rd))
    cot = torch.randn(out.data.shape, generator=torch.Generator().manual_seed(3)).to(dev)
    (out.data * cot).sum().backward()
    # oracle
    W = layer.vars['weight'].detach().double().cpu().requires_grad_(True)
    xs, rs = _rows(Xp, n), _rows(Rp, n)
    ys = [NO.block_end(x, r, W) for x, r in zip(xs, rs)]
    cots = torch.split(cot.double().cpu(), [int(k) for k in n])
    sum((y * c).sum() for y, c in zip(ys, cots)).backward()
    Y = out.padded().detach().cpu()
    for g, k in enumerate(n):
        assert O.rel_err(Y[g, :k], ys[g].detach()) <= TOL
        assert float(Y[g, k:].abs().max() if k < 132 else 0.0) == 0.0       # padding rows stay +0.0
    assert O.rel_err(layer.vars['weight'].grad.cpu(), W.grad) <= TOL
    assert O.rel_err(torch.cat([Xd.grad[g, :k] for g, k in enumerate(n)]).cpu(), torch.cat([x.grad for x in xs])) <= TOL
    assert O.rel_err(torch.cat([Rd.grad[g, :k] for g, k in enumerate(n)]).cpu(), torch.cat([r.grad for r in rs])) <= TOL


def test_dense_block_end_matches_oracle():
    """Two in-block activations (64 and 128 wide) and one preceding block output, flat lists like the reference's
    containers keep them (tf_graphs.py:269-271,357-361: index layer_id * batch_size + graph_id)."""
    import agcn_b200
    from agcn_b200.layers import DenseBlockEnd
    dev = torch.device("cuda:0")
    n = np.asarray(SIZES, np.int32)
    B = len(n)
    batch = agcn_b200.GraphBatch(n, 132, device=dev)
    F = 128
    Xp = _padded(n, F, 132, 5, relu=False)
    ins = [_padded(n, 64, 132, 6), _padded(n, 128, 132, 7)]
    outs = [_padded(n, 96, 132, 8)]
    dX = torch.tensor(Xp, device=dev, requires_grad=True)
    flat_in = [torch.tensor(a[g], device=dev) for a in ins for g in range(B)]
    flat_out = [torch.tensor(a[g], device=dev) for a in outs for g in range(B)]
    layer = DenseBlockEnd(0, [64, 128], F, 'relu', max_atom=132, batch_size=B)
    x = _x(batch, dev, node_features=dX, inblock_activations=flat_in, inblock_activations_dim=[64, 128],
           block_outputs=flat_out, block_outputs_dim=[96])
    layer.preceding_blocks_dim, layer.inblock_activations_dim = [96], [64, 128]
    layer.build()
    with torch.no_grad():
        layer.vars['beta_inblock'].fill_(0.7)
        layer.vars['beta_outblock'].fill_(1.3)
    out = layer(x)
    cot = torch.randn(out.data.shape, generator=torch.Generator().manual_seed(4)).to(dev)
    (out.data * cot).sum().backward()
    p = {k: ([w.detach().double().cpu().requires_grad_(True) for w in v] if isinstance(v, list)
             else v.detach().double().cpu().requires_grad_(True)) for k, v in layer.vars.items()}
    xs = _rows(Xp, n)
    ys = []
    for g, k in enumerate(n):
        a = [torch.tensor(t[g, :k], dtype=torch.float64) for t in ins]
        o = [torch.tensor(t[g, :k], dtype=torch.float64) for t in outs]
        ys.append(NO.dense_block_end(xs[g], a, p['weight_inblock'], o, p['weight_outblock'], p['beta_inblock'],
                                     p['beta_outblock']))
    cots = torch.split(cot.double().cpu(), [int(k) for k in n])
    sum((y * c).sum() for y, c in zip(ys, cots)).backward()
    assert O.rel_err(out.data.detach().cpu(), torch.cat([y.detach() for y in ys])) <= TOL
    for a, b in zip(layer.vars['weight_inblock'] + layer.vars['weight_outblock'], p['weight_inblock'] + p['weight_outblock']):
        assert O.rel_err(a.grad.cpu(), b.grad) <= TOL
    assert O.rel_err(layer.vars['beta_inblock'].grad.cpu(), p['beta_inblock'].grad) <= TOL
    assert O.rel_err(layer.vars['beta_outblock'].grad.cpu(), p['beta_outblock'].grad) <= TOL
    assert O.rel_err(torch.cat([dX.grad[g, :k] for g, k in enumerate(n)]).cpu(), torch.cat([t.grad for t in xs])) <= TOL


@pytest.mark.parametrize("Fin,hidden,Fout", [(75, [128], 64), (3, [32, 64], 64)])
def test_mlp_matches_oracle(Fin, hidden, Fout):
    import agcn_b200
    from agcn_b200.layers import MLP
    dev = torch.device("cuda:0")
    n = np.asarray(SIZES, np.int32)
    batch = agcn_b200.GraphBatch(n, 132, device=dev)
    Xp = _padded(n, Fin, 132, 9, relu=False)
    Xd = torch.tensor(Xp, device=dev, requires_grad=True)
    layer = MLP(Fout, hidden, Fin, len(n), max_atom=132)
    layer.build()
    with torch.no_grad():
        for b in layer.vars['bias']:
            b.copy_(torch.randn(b.shape, generator=torch.Generator().manual_seed(1)).to(dev) * 0.1)
    out = layer(_x(batch, dev, node_features=Xd))
    cot = torch.randn(out.data.shape, generator=torch.Generator().manual_seed(5)).to(dev)
    (out.data * cot).sum().backward()
    Ws = [w.detach().double().cpu().requires_grad_(True) for w in layer.vars['weight']]
    bs = [b.detach().double().cpu().requires_grad_(True) for b in layer.vars['bias']]
    xs = _rows(Xp, n)
    ys = [NO.mlp(x, Ws, bs) for x in xs]
    cots = torch.split(cot.double().cpu(), [int(k) for k in n])
    sum((y * c).sum() for y, c in zip(ys, cots)).backward()
    assert O.rel_err(out.data.detach().cpu(), torch.cat([y.detach() for y in ys])) <= TOL
    for a, b in zip(layer.vars['weight'] + layer.vars['bias'], Ws + bs):
        assert O.rel_err(a.grad.cpu(), b.grad) <= TOL
    assert O.rel_err(torch.cat([Xd.grad[g, :k] for g, k in enumerate(n)]).cpu(), torch.cat([t.grad for t in xs])) <= TOL


def test_dropout_mask_statistics_determinism_and_gradient():
    """tf.nn.dropout semantics (dropout.py:34-41): kept elements are scaled by 1 / (1 - p), the keep rate is 1 - p, a
    seed fixes the mask, the gradient passes through the same mask, and the test phase is the identity."""
    from agcn_b200.layers import Dropout, SGC_LL
    from agcn_b200.operators import model_operatos as model_ops
    dev = torch.device("cuda:0")
    x = (torch.rand(4001, 64, device=dev) + 0.5).requires_grad_(True)
    for p in (0.1, 0.5, 0.9):
        model_ops.set_learning_phase(1)
        y = Dropout(p, seed=7)(x)
        keep = y != 0
        assert abs(float(keep.float().mean()) - (1 - p)) < 0.01
        assert torch.allclose(y[keep], x.detach()[keep] / (1 - p), rtol=1e-6)
        assert torch.equal(y, Dropout(p, seed=7)(x)) and not torch.equal(y, Dropout(p, seed=8)(x))
        g, = torch.autograd.grad(y.sum(), x)
        assert torch.equal(g != 0, keep) and torch.allclose(g[keep], torch.full_like(g[keep], 1 / (1 - p)))
        # rows / columns are not correlated: per-column keep rates are all close to 1 - p
        assert float((keep.float().mean(0) - (1 - p)).abs().max()) < 0.05
        assert not torch.equal(Dropout(p)(x), Dropout(p)(x))              # unseeded: a fresh stream per call
    model_ops.set_learning_phase(0)
    assert Dropout(0.5, seed=7)(x) is x
    model_ops.set_learning_phase(1)
    # through the layer (graphconv.py:121-122)
    import agcn_b200
    n = np.asarray([20, 4, 9], np.int32)
    batch = agcn_b200.GraphBatch(n, 20, device=dev)
    X = torch.tensor(_padded(n, 16, 20, 1), device=dev)
    L = torch.zeros(3, 20, 20, device=dev)
    layer = SGC_LL(32, 16, 3, K=2, dropout=0.5)
    out, _, _ = layer(_x(batch, dev, node_features=X, original_laplacian=L, lap_slice=None))
    frac = float((out.data == 0).float().mean())
    assert 0.4 < frac < 0.95
