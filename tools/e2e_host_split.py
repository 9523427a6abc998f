"""Host time per call inside the pipelined e2e loop of bench.py (C2): plan + pack + labels ("stage") against the one
library call of the step, and the loop's wall time per step.  Tells a host-bound loop from a GPU-bound one."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import agcn_b200, bench
cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
dev = torch.device("cuda:0")
r = bench.Runner(cfg, dev, 0, 1)
agcn, model = agcn_b200, r.model
side = torch.cuda.Stream(device=dev, priority=0)
main = torch.cuda.Stream(device=dev, priority=-1)
lab = [(torch.empty_like(r.tg_h, device=dev), torch.empty_like(r.w_h, device=dev)) for _ in range(2)]
ready = [torch.cuda.Event() for _ in range(2)]
consumed = [torch.cuda.Event() for _ in range(2)]
T = {"plan": 0.0, "pack": 0.0, "labels": 0.0, "step": 0.0, "wait": 0.0}
def stage(i):
    slot = i % 2
    with torch.cuda.stream(side):
        side.wait_event(consumed[slot])
        t0 = time.perf_counter()
        b = agcn.GraphBatch(r.n_nodes, cfg["Nmax"], device=dev)
        t1 = time.perf_counter()
        X, L = b.pack_nodes(r.Xpad_h), b.pack_lap(r.Lpad_h)
        t2 = time.perf_counter()
        r.labels_to_device(lab[slot])
        ready[slot].record(side)
        t3 = time.perf_counter()
    T["plan"] += t1 - t0; T["pack"] += t2 - t1; T["labels"] += t3 - t2
    return b, X, L
with torch.cuda.stream(main):
    for ev in consumed:
        ev.record(main)
    host = [torch.empty(1).pin_memory() for _ in range(2)]
    N = 300
    for rep in range(2):
        for k in T: T[k] = 0.0
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        nxt = stage(0); pending = None
        for i in range(N):
            cur = nxt
            if i + 1 < N: nxt = stage(i + 1)
            slot = i % 2
            main.wait_event(ready[slot])
            b, X, L = cur
            t0 = time.perf_counter()
            loss = model.step(X, L, b, lab[slot][0], lab[slot][1])
            host[slot].copy_(loss, non_blocking=True)
            consumed[slot].record(main)
            done = torch.cuda.Event(); done.record(main)
            t1 = time.perf_counter()
            if pending is not None:
                pending[1].synchronize(); float(pending[0])
            t2 = time.perf_counter()
            T["step"] += t1 - t0; T["wait"] += t2 - t1
            pending = (host[slot], done)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
    print("wall per step %.3f ms; host per step: %s; host busy total %.3f ms" % (
        (w1 - w0) / N * 1e3, {k: round(v / N * 1e3, 3) for k, v in T.items()},
        sum(v for k, v in T.items() if k != "wait") / N * 1e3))
os._exit(0)
