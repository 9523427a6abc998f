"""Determinism / correctness of one layer's backward on the C1 batch (debug aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import agcn_b200
from agcn_b200.functional import sgc_ll_packed
from oracle import sgcll_oracle as O
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Xp, Lp, n = O.synthetic_molecule_batch(B, 132, seed=1235)
batch = agcn_b200.GraphBatch(n, 132, device=dev)
Ld = torch.from_numpy(np.concatenate([Lp[g, :k, :k].reshape(-1) for g, k in enumerate(n)])).to(dev)
for F, Fo in ((64, 128), (128, 128), (128, 64)):
    torch.manual_seed(F + Fo)
    X0 = torch.randn(batch.total_nodes, F, device=dev)
    p = {"weight": (torch.randn(F * 3, Fo, device=dev) * 0.05).requires_grad_(True), "bias": torch.zeros(Fo, device=dev, requires_grad=True),
         "M_L": (torch.randn(F, F, device=dev) * 0.05).requires_grad_(True), "alpha": torch.ones(1, device=dev, requires_grad=True)}
    cfg = {"F": F, "Fo": Fo, "K": 3, "variant": "SGC_LL", "laplacian": "reference_literal", "metric_grad": "reference", "activation": "relu"}
    cot = torch.randn(batch.total_nodes, Fo, device=dev)
    outs = []
    for it in range(4):
        X = X0.clone().requires_grad_(True)
        for v in p.values():
            v.grad = None
        Y = sgc_ll_packed(X, Ld, None, p, batch, cfg)
        Y = Y[0] if isinstance(Y, (tuple, list)) else Y
        (Y * cot).sum().backward()
        torch.cuda.synchronize()
        outs.append((X.grad.clone(), p["weight"].grad.clone(), Y.detach().clone()))
    for it in range(1, 4):
        print("F=%d Fo=%d run %d: dX equal %s (max diff %.3e, max |dX| %.3e), dW equal %s, Y equal %s" % (
            F, Fo, it, torch.equal(outs[it][0], outs[0][0]), float((outs[it][0] - outs[0][0]).abs().max()),
            float(outs[0][0].abs().max()), torch.equal(outs[it][1], outs[0][1]), torch.equal(outs[it][2], outs[0][2])))
