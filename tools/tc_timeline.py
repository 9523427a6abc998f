#!/usr/bin/env python
"""Per-k-block timeline of the first CTA of grouped_tc_kernel (agcn_debug_grouped_timeline): where a k-block's
1.3 us go -- waiting for the stage, splitting A, waiting for the TMA tile of B, splitting B, the MMAs.

    python tools/tc_timeline.py [--n 1024] [--graphs 32] [--F 128] [--trans 0]      # needs a GPU
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--graphs", type=int, default=32)
    ap.add_argument("--F", type=int, default=128)
    ap.add_argument("--trans", type=int, default=0)
    args = ap.parse_args()
    import agcn_b200
    from agcn_b200 import _lib
    from agcn_b200.batch import _ptr, _stream_ptr
    dev = torch.device("cuda:0")
    sizes = [args.n] * args.graphs
    batch = agcn_b200.GraphBatch(sizes, args.n, device=dev)
    L = torch.randn(batch.total_lap, device=dev) * 0.05
    X = torch.randn(batch.total_nodes, args.F, device=dev)
    out = torch.empty_like(X)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    lib = _lib.lib()

    def launch():
        _lib.check(lib.agcn_debug_grouped_product(batch.handle, _ptr(L), _ptr(X), _ptr(out), args.F, args.trans, 1, 2.0, 2,
                                                  _stream_ptr()))
    for _ in range(3):
        launch()
    flush.fill_(0.0)
    buf = torch.zeros(1024, dtype=torch.int64, device=dev)
    _lib.check(lib.agcn_debug_grouped_timeline(_ptr(buf)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    torch.cuda.synchronize()
    _lib.check(lib.agcn_debug_grouped_timeline(None))
    t = buf.cpu().numpy().astype(np.int64)
    t0 = t[520]
    kbs = min(64, (args.n + 31) // 32)
    print("kernel %.1f us (events); CTA 0: accumulator ready at %.2f us, done at %.2f us" %
          (e0.elapsed_time(e1) * 1e3, (t[521] - t0) / 1e3, (t[522] - t0) / 1e3))
    print("times in us from the start of CTA 0; w0 = first, w15 = last worker warp")
    print(" kb | tma issue | w0 stage free  A split done  B landed  arrive | w15 arrive | mma start  mma issued | k-block period")
    prev = None
    for kb in range(kbs):
        s = (t[8 * kb:8 * kb + 8] - t0) / 1e3
        period = "" if prev is None else "%6.2f" % (s[4] - prev)
        prev = s[4]
        print("%3d | %8.2f | %8.2f %12.2f %10.2f %8.2f | %9.2f | %8.2f %10.2f | %s" %
              (kb, s[6], s[0], s[1], s[2], s[3], s[7], s[4], s[5], period))
    d = (t[:8 * kbs].reshape(kbs, 8) - t0) / 1e3
    steady = slice(4, kbs)
    print("steady-state means (us): stage-free -> A done %.2f | A done -> B landed %.2f | B landed -> arrive %.2f | "
          "last arrive -> mma start %.2f | mma start -> next stage free (3 k-blocks later) %.2f | tma issue -> B landed %.2f" %
          ((d[steady, 1] - d[steady, 0]).mean(), (d[steady, 2] - d[steady, 1]).mean(), (d[steady, 3] - d[steady, 2]).mean(),
           (d[steady, 4] - np.maximum(d[steady, 3], d[steady, 7])).mean(),
           (d[7:kbs, 0] - d[4:kbs - 3, 4]).mean(), (d[steady, 2] - d[steady, 6]).mean()))


if __name__ == "__main__":
    main()
