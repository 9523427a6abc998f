"""ms per step of bench.py's e2e loop (zero-copy pack pipeline) for one workload:  python tools/e2e_quick.py [C2] [steps] [mode,mode,..]"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
r = bench.Runner(cfg, torch.device("cuda:0"), 0, 1)
for mode in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["zero_copy"] * 3):
    print("%s e2e %-10s %.4f ms per step" % (cfg["key"], mode, r.timed_e2e(steps, int(os.environ.get("WARMUP", "5")), mode)))
print("steps whose executable graph was updated in place:", r.model.step_graph_updates, "refusals:", r.model.step_graph_refusals[:8], "disabled:", r.model._graph_disabled)
os._exit(0)
