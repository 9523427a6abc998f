"""Stress of the resident training step (the code path bench.py times): many back-to-back steps, eager or as CUDA-graph
replays, to flush out rare races / protocol bugs of the hand-written kernels.

    python tools/stress_step.py --workload C2 --mode replay --iters 3000
    CUDA_LAUNCH_BLOCKING=1 python tools/stress_step.py --mode eager --iters 600     # names the failing launch

Prints one JSON line: iterations completed, the first failure (if any) and the library's last error.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--mode", default="replay", choices=["replay", "eager"])
    ap.add_argument("--iters", type=int, default=3000)
    ap.add_argument("--sync-every", type=int, default=50)
    ap.add_argument("--paper", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--smi", action="store_true", help="poll nvidia-smi beside the loop like bench.py's ClockSampler")
    args = ap.parse_args()
    import torch
    from agcn_b200 import _lib
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    kw = dict(laplacian="paper", metric_grad="full") if args.paper else {}
    r = bench.Runner(bench.WORKLOADS[args.workload], dev, 0, 1, **kw)
    fn = r.resident_step
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if args.mode == "replay":
        print("capture:", r.capture(), file=sys.stderr)
        fn = r.replay
    out = {"workload": args.workload, "mode": args.mode, "paper": args.paper, "iters": args.iters, "completed": 0,
           "failure": None}
    sampler = None
    if args.smi:
        sampler = bench.ClockSampler(0)
        sampler.start()
    t0 = time.time()
    try:
        for i in range(args.iters):
            if not args.no_flush and (i % 7) == 0:
                r.flush.fill_(1.0)
            fn()
            if (i + 1) % args.sync_every == 0:
                torch.cuda.synchronize()
                out["completed"] = i + 1
        torch.cuda.synchronize()
        out["completed"] = args.iters
    except Exception as exc:
        out["failure"] = str(exc)[:300]
        try:
            out["lib_error"] = _lib.lib().agcn_last_error().decode()
        except Exception:
            pass
    out["seconds"] = time.time() - t0
    if sampler is not None:
        try:
            out["clocks"] = sampler.stop()
        except Exception as exc:
            out["clocks"] = str(exc)[:100]
    print(json.dumps(out))
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
