"""Host-side cost of one end-to-end step (tuning aid; needs a GPU): where the CPU time of the e2e path goes.
    python tools/host_profile.py [C1|C2|C3|C4]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import agcn_b200
from agcn_b200 import _lib
import bench

cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C2"]
dev = torch.device("cuda:0")
r = bench.Runner(cfg, dev, 0, 1)
m = r.model
for _ in range(5):
    r.resident_step()
torch.cuda.synchronize()

def timeit(fn, n=50, sync=True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3

l0 = _lib.launch_count()
h, tot = timeit(r.resident_step)
print("resident step: host %.3f ms, host+drain %.3f ms, launches/step %.1f" % (h, tot, (_lib.launch_count() - l0) / 50.0))
h, tot = timeit(lambda: m.loss_and_grads(r.Xd, r.Ld, r.batch, r.tg_d, r.w_d))
print("  loss_and_grads: host %.3f ms (total %.3f)" % (h, tot))
h, tot = timeit(m.apply_adam)
print("  adam: host %.3f ms" % h)
h, tot = timeit(lambda: agcn_b200.GraphBatch(r.n_nodes, cfg["Nmax"], device=dev))
print("plan (GraphBatch): host %.3f ms" % h)
b = agcn_b200.GraphBatch(r.n_nodes, cfg["Nmax"], device=dev)
h, tot = timeit(lambda: (b.pack_nodes(r.Xpad_h), b.pack_lap(r.Lpad_h)))
print("pack (zero-copy): host %.3f ms, total %.3f ms" % (h, tot))
h, tot = timeit(lambda: (r.tg_h.to(dev, non_blocking=True), r.w_h.to(dev, non_blocking=True)))
print("labels H2D: host %.3f ms, total %.3f ms (%d bytes)" % (h, tot, r.tg_h.numel() * 4 + r.w_h.numel() * 4))
for mode in ("serial", "zero_copy"):
    print("e2e %s: %.3f ms/step" % (mode, r.timed_e2e(30, 5, mode)))

# ---- which part of the pipelined e2e loop costs the overlap: variants of the loop with pieces removed
def loop(n_steps, plan, pack, labels):
    side = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    lab = [(torch.empty_like(r.tg_h, device=dev), torch.empty_like(r.w_h, device=dev)) for _ in range(2)]
    for l in lab:
        l[0].copy_(r.tg_h); l[1].copy_(r.w_h)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    host = [torch.empty(1).pin_memory() for _ in range(2)]
    for ev in consumed:
        ev.record(main)

    def stage(i):
        slot = i % 2
        with torch.cuda.stream(side):
            side.wait_event(consumed[slot])
            b = agcn_b200.GraphBatch(r.n_nodes, cfg["Nmax"], device=dev) if plan else r.batch
            if pack:
                X, L = b.pack_nodes(r.Xpad_h), b.pack_lap(r.Lpad_h)
            else:
                X, L = r.Xd, r.Ld
            if labels:
                lab[slot][0].copy_(r.tg_h, non_blocking=True)
                lab[slot][1].copy_(r.w_h, non_blocking=True)
            ready[slot].record(side)
        return b, X, L

    def run(n):
        pending = None
        nxt = stage(0)
        for i in range(n):
            cur = nxt
            if i + 1 < n:
                nxt = stage(i + 1)
            slot = i % 2
            main.wait_event(ready[slot])
            b, X, L = cur
            loss = m.step(X, L, b, lab[slot][0], lab[slot][1])
            host[slot].copy_(loss, non_blocking=True)
            consumed[slot].record(main)
            done = torch.cuda.Event(); done.record(main)
            if pending is not None:
                pending.synchronize()
            pending = done
        pending.synchronize()

    run(5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(n_steps)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n_steps * 1e3

for name, args in (("no staging", (False, False, False)), ("plan only", (True, False, False)),
                   ("plan + pack", (True, True, False)), ("plan + pack + labels", (True, True, True)),
                   ("pack + labels, plan reused", (False, True, True))):
    print("pipelined loop, %s: %.3f ms/step" % (name, loop(40, *args)))
