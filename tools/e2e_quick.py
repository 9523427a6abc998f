"""ms per step of bench.py's e2e loop (zero-copy pack pipeline) for one workload:  python tools/e2e_quick.py [C2] [steps]"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
r = bench.Runner(cfg, torch.device("cuda:0"), 0, 1)
for _ in range(3):
    print("%s e2e %.4f ms per step" % (cfg["key"], r.timed_e2e(steps, 5, "zero_copy")))
os._exit(0)
