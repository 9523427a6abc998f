"""Time of the zero-copy pack kernels alone (pinned host arrays in the reference's padded layout -> packed device arrays)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import agcn_b200, bench
cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
dev = torch.device("cuda:0")
r = bench.Runner(cfg, dev, 0, 1)
b = agcn_b200.GraphBatch(r.n_nodes, cfg["Nmax"], device=dev)
for name, fn in (("nodes", lambda: b.pack_nodes(r.Xpad_h)), ("lap", lambda: b.pack_lap(r.Lpad_h))):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%s: %.3f ms per call, %.1f MB packed, %.1f GB/s" % (name, ms, out.numel() * 4 / 1e6, out.numel() * 4 / ms / 1e6))
os._exit(0)
