"""Does H2D traffic beside the step slow a REPLAYED CUDA graph of the step as it slows the eager launches?
(tools/e2e_host_split.py BISECT=8: a plain 8 MB cudaMemcpyAsync per step costs the eager ToxCast step 0.07 ms.)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
dev = torch.device("cuda:0")
r = bench.Runner(cfg, dev, 0, 1)
for _ in range(3):
    r.resident_step()
torch.cuda.synchronize()
side = torch.cuda.Stream(device=dev)
H = torch.empty(int(sys.argv[2]) << 18 if len(sys.argv) > 2 else 2 << 20).pin_memory()
D = torch.empty_like(H, device=dev)
def loop(fn, n, dma):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        if dma:
            with torch.cuda.stream(side):
                D.copy_(H, non_blocking=True)
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager : %.4f ms per step alone, %.4f with %d MB arriving beside every step" % (
    loop(r.resident_step, 200, False), loop(r.resident_step, 200, True), H.numel() * 4 >> 20))
print("capture:", r.capture())
print("replay: %.4f ms per step alone, %.4f with the copies" % (loop(r.replay, 200, False), loop(r.replay, 200, True)))
# host cost of capturing + instantiating the step anew (what a per-batch graph would pay every step)
r.release()
ts = []
for i in range(12):
    g = torch.cuda.CUDAGraph()
    t0 = time.perf_counter()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        r.resident_step()
    t1 = time.perf_counter()
    g.replay()
    torch.cuda.synchronize()
    ts.append((t1 - t0) * 1e3)
    del g
print("capture + instantiate per step (host): " + " ".join("%.2f" % t for t in ts) + " ms")
os._exit(0)
