"""Eager vs CUDA-graph replay of one stack step, per parameter tensor (debug aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import agcn_b200
from agcn_b200.simple_agcn import SimpleAGCNStep, synthetic_labels
from oracle import sgcll_oracle as O
from test_gpu_network import _names, _split

B, n_tasks = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 12
dev = torch.device("cuda:0")
X, L, n = O.synthetic_molecule_batch(B, 132, seed=1235)
batch = agcn_b200.GraphBatch(n, 132, device=dev)
Xd = batch.pack_nodes(torch.from_numpy(X).to(dev)); Ld = batch.pack_lap(torch.from_numpy(L).to(dev))
tg, w = synthetic_labels(B, n_tasks, 7, dev, "sigmoid_ce")
model = SimpleAGCNStep(75, (64, 128, 128, 64), 256, n_tasks, 3, B, device=dev, seed=11)
def run():
    model.flat_grad.zero_()
    model.loss_and_grads(Xd, Ld, batch, tg, w)
    torch.cuda.synchronize()
    return model.flat_grad.detach().clone()
e1 = run(); e2 = run()
print("eager == eager:", torch.equal(e1, e2))
side = torch.cuda.Stream(device=dev)
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    model.loss_and_grads(Xd, Ld, batch, tg, w)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    model.loss_and_grads(Xd, Ld, batch, tg, w)
for it in range(3):
    model.flat_grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    r = model.flat_grad.detach().clone()
    print("replay %d == eager: %s" % (it, torch.equal(r, e1)))
    for nm, a, b in zip(_names(model), _split(model, r.cpu()), _split(model, e1.cpu())):
        d = float((a - b).abs().max())
        if d > 0:
            print("   %-22s max |diff| %.3e (max |eager| %.3e)" % (nm, d, float(b.abs().max())))
e3 = run()
print("eager after == eager:", torch.equal(e3, e1))
