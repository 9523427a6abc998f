"""Parity of the code path bench.py times: SimpleAGCNStep (4 chained SGC_LL layers, DenseMol + GraphGatherMol +
heads + loss, flat gradient buffer, Adam) against the CPU oracle network (oracle/network_oracle.py), at the full
sizes of BASELINE.json's C1 (B = 256, 12 tasks) and C2 (B = 1024, 617 tasks) configurations.

Checked: the loss, EVERY parameter gradient of the flat buffer (two-stage reductions over all tiles, split-K head),
the parameters after one Adam step, engine "stack" (one library call, what the bench runs) == engine "autograd"
(the layer classes), and eager launches == CUDA-graph replay, bit for bit.
"""
import numpy as np
import pytest
import torch

from oracle import network_oracle as NO
from oracle import sgcll_oracle as O
from util_cases import TOL

pytestmark = pytest.mark.gpu


def _oracle_params(model):
    layer_p = [{k: v.detach().double().cpu().clone().requires_grad_(True) for k, v in l.vars.items()}
               for l in model.layers]
    head_p = {k: getattr(model, k).detach().double().cpu().clone().requires_grad_(True)
              for k in ("dense_W", "dense_b", "head_W", "head_b")}
    return layer_p, head_p


def _flat_oracle_grad(layer_p, head_p):
    parts = []
    for p in layer_p:
        for k in ("weight", "bias", "M_L", "alpha"):
            g = p[k].grad
            parts.append((torch.zeros_like(p[k]) if g is None else g).reshape(-1))
    for k in ("dense_W", "dense_b", "head_W", "head_b"):
        parts.append(head_p[k].grad.reshape(-1))
    return parts


def _names(model):
    out = []
    for i in range(len(model.layers)):
        out += ["layer%d.%s" % (i, k) for k in ("weight", "bias", "M_L", "alpha")]
    return out + ["dense_W", "dense_b", "head_W", "head_b"]


def _split(model, flat):
    out, off = [], 0
    for p in model.params:
        out.append(flat[off:off + p.numel()])
        off += p.numel()
    return out


def _perturb(model, seed):
    """Move bias / alpha / dense_b / head_b off their zero / one initial values so every gradient path is live."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for l in model.layers:
            l.vars["bias"].copy_((torch.randn(l.vars["bias"].shape, generator=g) * 0.05).to(l.vars["bias"].device))
            l.vars["alpha"].fill_(0.7)
        model.dense_b.copy_((torch.randn(model.dense_b.shape, generator=g) * 0.02).to(model.dense_b.device))
        model.head_b.copy_((torch.randn(model.head_b.shape, generator=g) * 0.05).to(model.head_b.device))


# Tolerances.  The layer meets the 1e-4 budget (tests/test_gpu_parity.py, <= 2e-6 measured).  At network level the
# gradients pass through GraphGatherMol's tanh of a SUM over the atoms of a molecule (graphgather.py:68-77), which the
# reference's own initialisation saturates (|a| ~ 50): tanh'(a) = sech^2(a) has relative sensitivity 2 |da|, so the
# absolute rounding error of the four chained layers is amplified ~100x.  The same oracle run in fp32 on the CPU
# (plain FMA arithmetic) is 2.5e-5 .. 4e-5 away from its fp64 self on these gradients; tensor-core accumulation
# (truncating adds) is ~6x coarser.  Hence two variants per shape: the network as initialised (tolerance 1e-3, and
# for C1 at most 12x the fp32 CPU oracle's own error) and the same network with DenseMol's weights scaled by 0.02 so
# that the tanh is not saturated (tolerance 1e-4: every gradient path at the layer's own budget).
CASES = [
    # name, B, n_tasks, loss, laplacian, metric_grad, dense_scale, tolerance
    ("C1_tox21_literal", 256, 12, "sigmoid_ce", "reference_literal", "reference", 1.0, 1e-3),
    ("C1_tox21_literal_unsaturated", 256, 12, "sigmoid_ce", "reference_literal", "reference", 0.02, TOL),
    ("C2_toxcast_literal", 1024, 617, "sigmoid_ce", "reference_literal", "reference", 1.0, 1e-3),
    ("C2_toxcast_literal_unsaturated", 1024, 617, "sigmoid_ce", "reference_literal", "reference", 0.02, TOL),
    ("C1_tox21_paper_full", 256, 12, "sigmoid_ce", "paper", "full", 1.0, 1e-3),
    ("C2_toxcast_paper_full_unsaturated", 1024, 617, "sigmoid_ce", "paper", "full", 0.02, TOL),
    ("softmax_head_literal_unsaturated", 96, 40, "softmax_ce", "reference_literal", "reference", 0.02, TOL),
]


@pytest.mark.parametrize("name,B,n_tasks,loss_kind,lap,mg,dense_scale,tol", CASES, ids=[c[0] for c in CASES])
def test_simple_agcn_step_matches_oracle_network(name, B, n_tasks, loss_kind, lap, mg, dense_scale, tol):
    import agcn_b200
    from agcn_b200.simple_agcn import SimpleAGCNStep, synthetic_labels
    dev = torch.device("cuda:0")
    X, L, n = O.synthetic_molecule_batch(B, 132, seed=1235)
    batch = agcn_b200.GraphBatch(n, 132, device=dev)
    Xd = batch.pack_nodes(torch.from_numpy(X).to(dev))
    Ld = batch.pack_lap(torch.from_numpy(L).to(dev))
    tg, w = synthetic_labels(B, n_tasks, 7, dev, loss_kind)
    w = w * (0.5 + torch.rand(w.shape, generator=torch.Generator().manual_seed(3)).to(dev))     # non-trivial weights

    model = SimpleAGCNStep(75, (64, 128, 128, 64), 256, n_tasks, 3, B, device=dev, laplacian=lap, metric_grad=mg,
                           loss=loss_kind, engine="stack", seed=11)
    _perturb(model, 5)
    with torch.no_grad():
        model.dense_W.mul_(dense_scale)
        if lap == "paper":
            # Glorot weights grow the activations ~4x per layer (|H_3| ~ 50 |X|): pairwise distances of ~100 put
            # exp(-dist) into fp32 denormals, where neither the reference (float32 numpy) nor any fp32 implementation
            # has significant digits left.  A smaller feature transform keeps the learned metric in range.
            for l in model.layers:
                l.vars["weight"].mul_(0.3)
    layer_p, head_p = _oracle_params(model)
    p_before = model.flat_params.flat.detach().clone()

    # ---- which outputs the CUDA path finds positive, layer by layer (see NO.simple_agcn_features: the oracle then
    # differentiates the same piecewise-linear branch; forward values are still compared through the loss)
    masks = []
    with torch.no_grad():
        from agcn_b200.batch import PackedLaplacians, PackedNodes
        xin = {'node_features': PackedNodes(Xd, batch), 'original_laplacian': PackedLaplacians(Ld, batch),
               'data_slice': None, 'lap_slice': None, '_batch': batch}
        for layer in model.layers:
            out, _, _ = layer(xin)
            masks.append((out.data > 0).cpu())
            xin = dict(xin, node_features=out)

    # ---- oracle: loss and gradients in fp64
    w_o = w.double().cpu()
    tg_o = tg.double().cpu()
    X64, L64 = torch.tensor(X, dtype=torch.float64), torch.tensor(L, dtype=torch.float64)
    if loss_kind == "softmax_ce":
        H = NO.simple_agcn_features(X64, L64, n, layer_p, 3, lap, mg, masks)
        mol = torch.tanh(torch.stack([(h @ head_p["dense_W"] + head_p["dense_b"]).sum(0) for h in H]))
        logits = mol @ head_p["head_W"] + head_p["head_b"]
        ce = torch.logsumexp(logits, 1) - (logits * tg_o).sum(1)
        loss_o = (ce * w_o).sum() / B
    else:
        loss_o = NO.simple_agcn_loss(X64, L64, n, layer_p, head_p, tg_o, w_o, B, 3, lap, mg, masks)
    loss_o.backward()
    grads_o = _flat_oracle_grad(layer_p, head_p)

    # ---- the path bench.py times: one agcn_stack_loss_grad call
    loss = model.loss_and_grads(Xd, Ld, batch, tg, w)
    torch.cuda.synchronize()
    g_stack = model.flat_grad.detach().clone()
    assert abs(float(loss) - float(loss_o)) <= TOL * abs(float(loss_o)), (float(loss), float(loss_o))
    worst = {}
    for nm, a, b in zip(_names(model), _split(model, g_stack.cpu()), grads_o):
        worst[nm] = O.rel_err(a, b) if float(b.abs().max()) > 0 else float(a.abs().max())
    print(name, {k: "%.1e" % v for k, v in worst.items()})
    bad = {k: v for k, v in worst.items() if not v <= tol}
    assert not bad, "gradient mismatch vs oracle network: %s" % bad
    if name == "C1_tox21_literal":
        # the same oracle in fp32 on the CPU: how far plain fp32 arithmetic is from fp64 on this network
        lp32 = [{k: v.detach().float().requires_grad_(True) for k, v in p.items()} for p in layer_p]
        hp32 = {k: v.detach().float().requires_grad_(True) for k, v in head_p.items()}
        NO.simple_agcn_loss(torch.tensor(X), torch.tensor(L), n, lp32, hp32, tg_o.float(), w_o.float(), B, 3, lap,
                            mg, masks).backward()
        e32 = [O.rel_err(a, b) for a, b in zip(_flat_oracle_grad(lp32, hp32), grads_o) if float(b.abs().max()) > 0]
        print("fp32 CPU oracle vs fp64: max %.1e; CUDA vs fp64: max %.1e" % (max(e32), max(worst.values())))
        assert max(worst.values()) <= 12 * max(max(e32), 1e-5)

    # ---- the drop-in path (layer classes + autograd) must produce the same numbers, bit for bit
    model.flat_grad.zero_()
    model.engine = "autograd"
    loss_a = model.loss_and_grads(Xd, Ld, batch, tg, w)
    torch.cuda.synchronize()
    model.engine = "stack"
    assert float(loss_a) == float(loss)
    assert torch.equal(model.flat_grad, g_stack), "stack and autograd engines disagree"

    # ---- CUDA-graph replay of the same call == eager launches, bit for bit
    model.flat_grad.zero_()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model.loss_and_grads(Xd, Ld, batch, tg, w)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        model.loss_and_grads(Xd, Ld, batch, tg, w)
    model.flat_grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(model.flat_grad, g_stack), "CUDA-graph replay differs from eager launches"
    assert float(model._loss) == float(loss)
    del graph

    # ---- Adam (tf.train.AdamOptimizer's rule) on the flat buffers: two steps against the oracle formula
    p_o = p_before.double().cpu()
    m_o = torch.zeros_like(p_o)
    v_o = torch.zeros_like(p_o)
    for t in (1, 2):
        model.apply_adam()
        p_o, m_o, v_o = NO.adam_tf(p_o, g_stack.double().cpu(), m_o, v_o, t, lr=model.lr, eps=model.epsilon)
    torch.cuda.synchronize()
    assert int(model.adam_step) == 2
    assert O.rel_err(model.flat_params.flat.detach().cpu(), p_o) <= 1e-6
    assert O.rel_err((model.flat_params.flat.detach() - p_before).cpu(), p_o - p_before.double().cpu()) <= 1e-4


def test_step_decreases_loss_and_is_reproducible():
    """Ten full steps (loss_and_grads + Adam) on one batch: the loss goes down, and a second model with the same seed
    follows the same trajectory bit for bit (deterministic reductions, no atomics)."""
    import agcn_b200
    from agcn_b200.simple_agcn import SimpleAGCNStep, synthetic_labels
    dev = torch.device("cuda:0")
    B = 128
    X, L, n = O.synthetic_molecule_batch(B, 132, seed=77)
    batch = agcn_b200.GraphBatch(n, 132, device=dev)
    Xd = batch.pack_nodes(torch.from_numpy(X).to(dev))
    Ld = batch.pack_lap(torch.from_numpy(L).to(dev))
    tg, w = synthetic_labels(B, 12, 3, dev)
    traj = []
    for rep in range(2):
        model = SimpleAGCNStep(75, (64, 128, 128, 64), 256, 12, 3, B, device=dev, seed=5)
        losses = [float(model.step(Xd, Ld, batch, tg, w)) for _ in range(10)]
        traj.append(losses)
        assert np.isfinite(losses).all() and losses[-1] < losses[0]
    assert traj[0] == traj[1]


def test_graphed_step_follows_the_eager_step_bit_for_bit_over_changing_batches():
    """step_graphed (the step captured anew for every batch, the executable graph of the previous step updated in place
    and launched once) against step() (eager launches) over batches of different composition and size, some repeated:
    same losses and same parameters bit for bit; batches of the same bucket structure update the graph in place, and
    a batch with another node topology (point-cloud sizes: other kernels) falls back to a fresh instantiation."""
    import agcn_b200
    from agcn_b200.simple_agcn import SimpleAGCNStep, synthetic_labels
    dev = torch.device("cuda:0")
    B = 96
    rng = np.random.default_rng(8)
    batches = []
    for seed, nmax in ((1, 132), (2, 132), (3, 40), (4, 132)):
        X, L, n = O.synthetic_molecule_batch(B, 132, seed=seed)
        if nmax < 132:                                  # a batch of small molecules only: other size buckets
            n = np.minimum(n, nmax).astype(np.int32)
        b = agcn_b200.GraphBatch(n, 132, device=dev)
        batches.append((b, b.pack_nodes(torch.from_numpy(X).to(dev)), b.pack_lap(torch.from_numpy(L).to(dev))))
    big_n = np.asarray([200, 150] + [12] * (B - 2), np.int32)     # two graphs above 144 nodes: row-tiled kernels join in
    Xb = (rng.standard_normal((B, 200, 75)) * 0.3).astype(np.float32)
    Lb = (rng.standard_normal((B, 200, 200)) * 0.02).astype(np.float32)
    bb = agcn_b200.GraphBatch(big_n, 200, device=dev)
    batches.append((bb, bb.pack_nodes(torch.from_numpy(Xb).to(dev)), bb.pack_lap(torch.from_numpy(Lb).to(dev))))
    tg, w = synthetic_labels(B, 12, 3, dev)
    order = [0, 1, 0, 2, 3, 4, 1, 4, 0]
    runs = []
    for graphed in (False, True):
        model = SimpleAGCNStep(75, (64, 128, 128, 64), 256, 12, 3, B, device=dev, seed=5)
        losses = []
        for k in order:
            b, Xd, Ld = batches[k]
            fn = model.step_graphed if graphed else model.step
            losses.append(float(fn(Xd, Ld, b, tg, w)))
        torch.cuda.synchronize()
        runs.append((losses, model.flat_params.flat.detach().clone(), model.step_graph_updates))
    assert np.isfinite(runs[0][0]).all()
    assert runs[0][0] == runs[1][0]
    assert torch.equal(runs[0][1], runs[1][1])
    assert runs[0][2] == 0 and runs[1][2] >= 3          # at least the repeated molecule batches were updated in place


def test_plans_created_and_destroyed_back_to_back_behind_a_long_kernel():
    """A plan per batch with the host far ahead of the device: plans are destroyed while their table upload is still
    queued, so their pinned staging buffers go back to the pool in flight.  Re-acquiring them must neither reuse a
    buffer that is still being read nor leave cudaErrorNotReady behind for the next launch check."""
    import agcn_b200
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    big = torch.empty(64 * 1024 * 1024, device=dev)
    X = torch.randn(4000, 16, device=dev)
    torch.cuda.synchronize()
    expected = []
    outs = []
    for it in range(24):
        for _ in range(4):
            big.normal_()                              # keeps the stream busy: the host runs ahead
        n = rng.integers(4, 40, size=50).astype(np.int32)
        b = agcn_b200.GraphBatch(n, 40, device=dev)
        padded = torch.zeros(50, 40, 16, device=dev)
        off = 0
        for g, k in enumerate(n):
            padded[g, :k] = X[off:off + k]
            off += k
        outs.append((b.pack_nodes(padded), int(n.sum())))
        del b                                          # destroyed right away, upload possibly still queued
    torch.cuda.synchronize()
    for packed, R in outs:
        assert torch.equal(packed, X[:R])


@pytest.mark.gpu
def test_expand_labels_from_pinned_host_matches_one_hot():
    """agcn_expand_labels: tf.one_hot(label, 2) + per-logit weights (multitask_classifier.py:196-199) from the compact
    labels the reference feeds, read in place from pinned host memory; bit-exact against the numpy construction."""
    from agcn_b200.simple_agcn import expand_labels
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(5)
    for B, T in ((7, 12), (1024, 617), (3, 1)):
        y = rng.random((B, T)) < 0.3
        w = rng.random((B, T)).astype(np.float32)
        want_t = np.stack([1.0 - y, y], -1).astype(np.float32).reshape(B, 2 * T)
        want_w = np.repeat(w, 2, axis=1)
        tg = torch.full((B, 2 * T), -7.0, device=dev)
        ww = torch.full((B, 2 * T), -7.0, device=dev)
        expand_labels(torch.from_numpy(y).pin_memory(), torch.from_numpy(w).pin_memory(), tg, ww)
        torch.cuda.synchronize()
        assert np.array_equal(tg.cpu().numpy(), want_t) and np.array_equal(ww.cpu().numpy(), want_w)
        tg.fill_(-7.0)
        expand_labels(torch.from_numpy(y.astype(np.uint8)).to(dev), torch.from_numpy(w).to(dev), tg, ww)   # device inputs
        torch.cuda.synchronize()
        assert np.array_equal(tg.cpu().numpy(), want_t)
    with pytest.raises(ValueError):
        expand_labels(torch.zeros(2, 3, dtype=torch.uint8), torch.zeros(2, 3), tg, ww)   # pageable host memory
