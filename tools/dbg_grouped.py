#!/usr/bin/env python
"""Block-by-block comparison of the row-tiled product implementations (agcn_debug_grouped_product) against a
float64 torch product on the device: tells a layout / transposition / tile-edge bug apart at a glance.

    python tools/dbg_grouped.py            # needs a GPU
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def main():
    import agcn_b200
    from agcn_b200 import _lib
    from agcn_b200.batch import _ptr, _stream_ptr
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(3)
    worst = 0.0
    for sizes, F in (([300, 145, 257], 128), ([1024], 32), ([161, 450], 96), ([513], 256), ([200, 150], 4), ([333], 7)):
        batch = agcn_b200.GraphBatch(sizes, max(sizes), device=dev)
        R = batch.total_nodes
        L = torch.randn(batch.total_lap, device=dev, generator=gen) * 0.1
        X = torch.randn(R, F, device=dev, generator=gen)
        for transL in (0, 1):
            for add_identity in (0, 1):
                ref = torch.zeros(R, F, device=dev, dtype=torch.float64)
                for g, n in enumerate(sizes):
                    Lg = batch.lap_view(L, g).double()
                    if transL:
                        Lg = Lg.t()
                    if add_identity:
                        Lg = Lg + torch.eye(n, device=dev, dtype=torch.float64)
                    r0 = int(batch.node_off[g])
                    ref[r0:r0 + n] = 2.0 * (Lg @ X[r0:r0 + n].double())
                impls = [1] + ([3] if F <= 8 else []) + ([2] if F >= 16 and F % 4 == 0 else [])
                for impl in impls:
                    out = torch.full((R, F), float("nan"), device=dev)
                    rc = _lib.lib().agcn_debug_grouped_product(batch.handle, _ptr(L), _ptr(X), _ptr(out), F, transL,
                                                               add_identity, 2.0, impl, _stream_ptr())
                    _lib.check(rc)
                    torch.cuda.synchronize()
                    err = (out.double() - ref).abs()
                    bad = ~torch.isfinite(out)
                    rel = float(err[~bad].max() / ref.abs().max()) if (~bad).any() else float("nan")
                    worst = max(worst, rel if impl != 1 else 0.0)
                    print("sizes %-18s F %3d transL %d +I %d impl %d : rel err %.2e  non-finite %d" %
                          (sizes, F, transL, add_identity, impl, rel, int(bad.sum())))
                    if impl != 1 and (rel > 1e-5 or bad.any()):
                        # error map per (32-row block, 32-column block) of the first graph
                        n = sizes[0]
                        e = err[:n].clone()
                        e[bad[:n]] = 9.0
                        rows = (n + 31) // 32
                        cols = (F + 31) // 32
                        for rb in range(min(rows, 12)):
                            print("   rows %4d: " % (32 * rb) + " ".join(
                                "%.1e" % float(e[32 * rb:32 * rb + 32, 32 * cb:32 * cb + 32].max()) for cb in range(cols)))
    print("worst relative error of the new implementations: %.2e" % worst)


if __name__ == "__main__":
    main()
