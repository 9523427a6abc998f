"""20-second probe: do the pack kernels read pinned host memory in place, and how fast?  (needs a GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import agcn_b200

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
B, N, F = 1024, 132, 75
n = np.clip(np.round(rng.lognormal(np.log(17), 0.55, B)), 4, N).astype(np.int32)
Xh = torch.from_numpy(rng.standard_normal((B, N, F)).astype(np.float32)).pin_memory()
Lh = torch.from_numpy(rng.standard_normal((B, N, N)).astype(np.float32)).pin_memory()
b = agcn_b200.GraphBatch(n, N, device=dev)
Xa, La = b.pack_nodes(Xh), b.pack_lap(Lh)
Xb, Lb = b.pack_nodes(Xh.to(dev)), b.pack_lap(Lh.to(dev))
torch.cuda.synchronize()
print("bit exact:", torch.equal(Xa, Xb), torch.equal(La, Lb))
for name, fn in (("zero-copy", lambda: (b.pack_nodes(Xh), b.pack_lap(Lh))),
                 ("copy+pack", lambda: (b.pack_nodes(Xh.to(dev, non_blocking=True)), b.pack_lap(Lh.to(dev, non_blocking=True))))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%s: %.3f ms per batch (real bytes %.1f MB, wire bytes %.1f MB)" %
          (name, e0.elapsed_time(e1) / 5, (Xa.numel() + La.numel()) * 4 / 1e6, (Xh.numel() + Lh.numel()) * 4 / 1e6))
