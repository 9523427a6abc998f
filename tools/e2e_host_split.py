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
BISECT = int(os.environ.get("BISECT", "7"))   # bit mask of what is staged per step: 1 plan, 2 pack, 4 labels (7 = the loop of bench.py)
KEEP = []
DUMMY_H = torch.empty(2 << 20).pin_memory()
DUMMY_D = torch.empty(2 << 20, device=dev)
T = {"plan": 0.0, "pack": 0.0, "labels": 0.0, "step": 0.0, "wait": 0.0}
def stage(i):
    slot = i % 2
    with torch.cuda.stream(side):
        side.wait_event(consumed[slot])
        t0 = time.perf_counter()
        first = not KEEP
        if first or (BISECT & 1):
            b = agcn.GraphBatch(r.n_nodes, cfg["Nmax"], device=dev)
        else:
            b = KEEP[0][0]
        t1 = time.perf_counter()
        if first or (BISECT & 2):
            X, L = b.pack_nodes(r.Xpad_h), b.pack_lap(r.Lpad_h)
        else:
            X, L = KEEP[0][1], KEEP[0][2]
        t2 = time.perf_counter()
        if BISECT & 8:                    # a dummy 8 MB host -> device copy on the copy engine beside the step
            DUMMY_D.copy_(DUMMY_H, non_blocking=True)
        if (BISECT & 16) and not first:   # no transfer at all, but the step alternates between two copies of the batch
            if len(KEEP) < 2:
                KEEP.append((b, KEEP[0][1].clone(), KEEP[0][2].clone()))
            X, L = KEEP[i % 2][1], KEEP[i % 2][2]
        if first:
            KEEP.append((b, X, L))
            r.labels_to_device(lab[0]); r.labels_to_device(lab[1])
        elif BISECT & 4:
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
    evs = []
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
            tl_now = rep == 1 and i == 150 and os.environ.get("E2E_TIMELINE")
            if tl_now:   # one step of the loop under the library's event timeline (pack of the next step beside it)
                torch.cuda.synchronize()
                from agcn_b200 import _lib
                _lib.profile_enable(True)
            t0 = time.perf_counter()
            if rep == 1 and i >= 10:
                ev0 = torch.cuda.Event(enable_timing=True); ev0.record(main)
            loss = model.step(X, L, b, lab[slot][0], lab[slot][1])
            if rep == 1 and i >= 10:
                ev1 = torch.cuda.Event(enable_timing=True); ev1.record(main); evs.append((ev0, ev1))
            host[slot].copy_(loss, non_blocking=True)
            if tl_now:
                torch.cuda.synchronize()
                tl = _lib.profile_timeline()
                _lib.profile_enable(False)
                print("# one step inside the e2e loop, %d profiled launches" % len(tl))
                for name, a, b_ in sorted(tl, key=lambda x: x[1]):
                    print("  %9.1f %9.1f %8.1f  %s" % (a * 1e3, b_ * 1e3, (b_ - a) * 1e3, name))
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
    gpu = [a.elapsed_time(b) for a, b in evs]
    gaps = [evs[k + 1][0].elapsed_time(evs[k][1]) * -1.0 for k in range(len(evs) - 1)]
    gs = sorted(gpu)
    print("BISECT=%d GPU time of the step's launches on the main stream: mean %.3f ms (min %.3f, median %.3f, p90 %.3f, max %.3f); idle between steps: mean %.3f ms" % (
        BISECT, sum(gpu) / len(gpu), gs[0], gs[len(gs) // 2], gs[int(len(gs) * 0.9)], gs[-1], sum(gaps) / len(gaps)))
os._exit(0)
