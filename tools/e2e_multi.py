"""e2e loop of bench.py on N GPUs (torchrun), eager launches against the graph-launched step with the NCCL all-reduce
captured too (SimpleAGCNStep.graph_collectives):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/e2e_multi.py C2 100"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
r = bench.Runner(cfg, dev, rank, world)
r.model.graph_collectives = True
for mode in ("zero_copy", "graph", "zero_copy", "graph"):
    ms = r.timed_e2e(steps, 8, mode)
    if rank == 0:
        print("N=%d %s e2e %-10s %.4f ms per step = %.0f graphs/s" % (world, cfg["key"], mode, ms, world * r.B / ms * 1e3), flush=True)
if rank == 0:
    print("updated in place:", r.model.step_graph_updates, "refusals:", r.model.step_graph_refusals[:6], flush=True)
# parameters identical on every rank after the graph-launched steps
p = r.model.flat_params.flat.detach().clone()
q = p.clone()
dist.broadcast(q, 0)
print("rank %d parameters equal to rank 0: %s" % (rank, bool(torch.equal(p, q))), flush=True)
torch.cuda.synchronize()
os._exit(0)
