"""Minimal 2-rank NCCL probe (torchrun): init, one all-reduce, one all-reduce captured in a CUDA graph."""
import os, sys, time
import torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
t0 = time.time()
dist.init_process_group("nccl", device_id=dev)
print("[probe %d] init %.1fs" % (rank, time.time() - t0), flush=True)
x = torch.ones(1 << 19, device=dev)
dist.all_reduce(x); torch.cuda.synchronize()
print("[probe %d] eager all_reduce ok %g" % (rank, float(x[0])), flush=True)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        dist.all_reduce(x)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dist.all_reduce(x)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
print("[probe %d] graph all_reduce ok" % rank, flush=True)
dist.barrier(); torch.cuda.synchronize()
print("[probe %d] done (leaving without destroy_process_group: it hangs while a graph with captured NCCL is alive)" % rank, flush=True)
os._exit(0)
