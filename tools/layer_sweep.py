#!/usr/bin/env python
"""One-layer SGC-LL microbenchmark over BASELINE.json's other shapes (C1, C3, C4, C5).

    python tools/layer_sweep.py [--group c1,c3,c4,sweep,sweep_big] [--out gpurun_out/layer_sweep.jsonl]

For every case: forward and forward+backward of ONE layer (graphconv.py:127-252 through
agcn_b200.functional.sgc_ll_packed -> the C ABI), timed with CUDA events on the launching stream, L2 flushed
(256 MB fill) before every timed iteration, 3 warm-up + `--iters` timed iterations, median.  One JSON line per
case: graph-layers/s, algorithmic GB/s and TFLOP/s (SURVEY.md section 8d formulas, real n) as fractions of
MEASURED_PEAKS.json (tf32 dense taken as half the measured bf16 peak), and which of the two bounds the case.
bench.py stays the headline (C2); these lines are the per-shape evidence under profiles/.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p["hbm_gbs"], p["bf16_tflops"] / 2.0, "measured"
    return 6650.0, 795.0, "fallback"


def algorithmic(n, F, Fo, K, metric):
    """flops / bytes of one graph-layer summed over the batch."""
    n = n.astype(np.float64)
    f_fwd = 2 * K * n * n * F + 2 * n * K * F * Fo
    f_bwd = 4 * n * K * F * Fo + 4 * (K - 1) * n * n * F
    if metric:                     # projection + pairwise distances, and their gradients
        f_fwd = f_fwd + 2 * n * F * F
        f_bwd = f_bwd + 2 * n * n * F + 4 * n * F * F
    if K == 1:
        f_fwd = f_fwd - 2 * n * n * F   # no L*T product at K = 1
    b_fwd = 4 * (n * F + n * n + n * Fo)
    b_bwd = 4 * (2 * n * F + n * n + n * Fo)
    par = 4.0 * (F * F * (1 if metric else 0) + K * F * Fo + Fo + 2)
    return f_fwd.sum(), f_bwd.sum(), b_fwd.sum() + par, b_bwd.sum() + par


def knn_laplacians(n_nodes, dev, gen, deg=16):
    """Packed normalised Laplacians I - D^-1/2 A D^-1/2 of random symmetric graphs with ~deg neighbours."""
    out = []
    for n in n_nodes:
        n = int(n)
        k = min(deg, n - 1)
        idx = torch.randint(0, n, (n, k), device=dev, generator=gen)
        A = torch.zeros(n, n, device=dev)
        A.scatter_(1, idx, 1.0)
        A = torch.maximum(A, A.t())
        A.fill_diagonal_(0.0)
        d = A.sum(1).clamp_min(1.0).rsqrt()
        L = torch.eye(n, device=dev) - d[:, None] * A * d[None, :]
        out.append(L.reshape(-1))
    return torch.cat(out)


def threshold_laplacians(X, n_nodes, dev, quantile=None):
    """Point-cloud adjacency: d_ij < mean pairwise distance (meshloader.py:264-285) or the q-quantile rule
    (pointcloudloader.py:240-263), then the normalised Laplacian of A + I."""
    out, off = [], 0
    for n in n_nodes:
        n = int(n)
        P = X[off:off + n, :3]
        off += n
        D = torch.cdist(P, P)
        thr = D.mean() if quantile is None else torch.quantile(D.flatten()[:: max(1, D.numel() // 200000)], quantile)
        A = (D < thr).float()
        A.fill_diagonal_(1.0)
        d = A.sum(1).rsqrt()
        out.append((torch.eye(n, device=dev) - d[:, None] * A * d[None, :]).reshape(-1))
    return torch.cat(out)


def cases(groups):
    rng = np.random.default_rng(1234)
    cs = []
    if "c1" in groups:   # Tox21-shape: B = 256, Nmax = 132, lognormal n, first and hidden layers, K = 3
        n = np.clip(np.round(rng.lognormal(np.log(17), 0.55, 256)), 4, 132).astype(np.int32)
        n[0] = 132
        for F, Fo in ((75, 64), (128, 128)):
            for sem in ("literal", "paper"):
                cs.append(dict(name="C1 Tox21-shape B=256", n=n, F=F, Fo=Fo, K=3, lap="knn", sem=sem))
    if "c3" in groups:   # ModelNet40-shape: B = 32 clouds of 1024 points
        n = np.full(32, 1024, np.int32)
        for F, Fo in ((3, 64), (128, 128)):
            for sem in ("literal", "paper"):
                cs.append(dict(name="C3 ModelNet40-shape B=32 N=1024", n=n, F=F, Fo=Fo, K=3, lap="mean", sem=sem))
    if "c4" in groups:   # Sydney-shape: ragged loguniform[13, 1024], B = 128
        n = np.exp(rng.uniform(np.log(13), np.log(1024), 128)).round().astype(np.int32)
        for F, Fo in ((4, 64), (128, 128)):
            for sem in ("literal", "paper"):
                cs.append(dict(name="C4 Sydney-shape ragged B=128", n=n, F=F, Fo=Fo, K=3, lap="q10", sem=sem))
    if "sweep" in groups or "sweep_big" in groups:
        Ns = []
        if "sweep" in groups:
            Ns += [64, 128, 256, 512, 1024]
        if "sweep_big" in groups:
            Ns += [2048, 4096]
        for N in Ns:
            B = max(1, (64 << 20) // (N * N))          # B * N^2 * 4 bytes = 256 MB
            B = min(B, 4096)
            for F in (32, 128, 256):
                for K in (1, 3, 5):
                    if F == 256 and K == 5 and N <= 128:
                        continue
                    cs.append(dict(name="C5 sweep N=%d" % N, n=np.full(B, N, np.int32), F=F, Fo=F, K=K, lap="knn",
                                   sem="literal"))
            cs.append(dict(name="C5 sweep N=%d" % N, n=np.full(B, N, np.int32), F=128, Fo=128, K=3, lap="knn", sem="paper"))
    return cs


_LAP_CACHE = {}


def run_case(c, dev, iters, flush):
    import agcn_b200
    from agcn_b200.functional import sgc_ll_packed
    n, F, Fo, K = c["n"], c["F"], c["Fo"], c["K"]
    gen = torch.Generator(device=dev).manual_seed(7)
    batch = agcn_b200.GraphBatch(n, int(n.max()), device=dev)
    R = batch.total_nodes
    if F <= 4:
        X = torch.randn(R, F, device=dev, generator=gen)
        X = X / X.abs().max()
    else:
        X = torch.relu(torch.randn(R, F, device=dev, generator=gen)) * 0.5
    key = (c["lap"], n.tobytes())
    if key not in _LAP_CACHE:
        _LAP_CACHE.clear()
        if c["lap"] == "knn":
            _LAP_CACHE[key] = knn_laplacians(n, dev, gen)
        else:
            pts = torch.randn(R, 3, device=dev, generator=gen)
            _LAP_CACHE[key] = threshold_laplacians(pts, n, dev, None if c["lap"] == "mean" else 0.10)
    L = _LAP_CACHE[key]
    X.requires_grad_(True)
    def glorot(r, c):                          # graphconv.py:14-18
        lim = float(np.sqrt(6.0 / (r + c)))
        return ((torch.rand(r, c, device=dev, generator=gen) * 2 - 1) * lim).requires_grad_(True)

    p = {"weight": glorot(F * K, Fo), "bias": torch.zeros(Fo, device=dev, requires_grad=True), "M_L": glorot(F, F),
         "alpha": torch.ones(1, device=dev, requires_grad=True)}
    paper = c["sem"] == "paper"
    cfg = {"F": F, "Fo": Fo, "K": K, "variant": "SGC_LL", "laplacian": "paper" if paper else "reference_literal",
           "metric_grad": "full" if paper else "reference", "activation": "relu"}
    dY = torch.randn(R, Fo, device=dev, generator=gen)

    def fwd():
        with torch.no_grad():
            return sgc_ll_packed(X, L, None, p, batch, cfg)[0]

    def fwd_bwd():
        Y = sgc_ll_packed(X, L, None, p, batch, cfg)[0]
        torch.autograd.backward(Y, dY, inputs=[X] + list(p.values()))
        return Y

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    ms_f = timed(fwd)
    ms_fb = timed(fwd_bwd)
    y = fwd()
    assert bool(torch.isfinite(y).all()), "non-finite output"
    hbm, tf32, src = peaks()
    ff, fb, bf, bb = algorithmic(n, F, Fo, K, paper)
    B = len(n)
    gbs_f, gbs_fb = bf / ms_f / 1e6, (bf + bb) / ms_fb / 1e6
    tfs_f, tfs_fb = ff / ms_f / 1e9, (ff + fb) / ms_fb / 1e9
    ai = (ff + fb) / (bf + bb)
    ridge = tf32 * 1e12 / (hbm * 1e9)
    return {"case": c["name"], "B": B, "n_mean": float(n.mean()), "n_max": int(n.max()), "F": F, "Fo": Fo, "K": K,
            "semantics": "paper+full_metric_grad" if paper else "reference_literal",
            "ms_fwd": ms_f, "ms_fwd_bwd": ms_fb, "graph_layers_per_s": B / (ms_fb * 1e-3),
            "fwd": {"GBps": gbs_f, "hbm_frac": gbs_f / hbm, "TFLOPs": tfs_f, "tf32_frac": tfs_f / tf32},
            "fwd_bwd": {"GBps": gbs_fb, "hbm_frac": gbs_fb / hbm, "TFLOPs": tfs_fb, "tf32_frac": tfs_fb / tf32},
            "arith_intensity": ai, "bound": "tensor" if ai > ridge else "hbm", "peaks": src}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="c1,c3,c4,sweep")
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "layer_sweep.jsonl"))
    ap.add_argument("--budget-s", type=float, default=240.0, help="stop starting new cases after this many seconds")
    ap.add_argument("--match", default="", help="only cases whose key 'name F=.. K=.. sem' contains every one of these "
                                                "comma-separated substrings, e.g. 'N=1024 ,F=128 ,K=3'")
    ap.add_argument("--tag", default="", help="copied into every line (which build / environment produced it)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise RuntimeError("layer_sweep.py needs a CUDA device; there is no CPU path")
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    t0 = time.time()
    with open(args.out, "a") as fh:
        for c in cases(set(args.group.split(","))):
            key = "%s F=%d K=%d %s " % (c["name"] + " ", c["F"], c["K"], c["sem"])
            if any(m not in key for m in args.match.split(",") if m):
                continue
            if time.time() - t0 > args.budget_s:
                print("budget reached, stopping", file=sys.stderr)
                break
            try:
                line = run_case(c, dev, args.iters, flush)
            except Exception as exc:                           # keep the rest of the table
                line = {"case": c["name"], "F": c["F"], "Fo": c["Fo"], "K": c["K"], "semantics": c["sem"],
                        "error": repr(exc)[:300]}
                if "CUDA" in repr(exc) or "cuda" in repr(exc):
                    fh.write(json.dumps(line) + "\n")
                    print(json.dumps(line))
                    break
            if args.tag:
                line["tag"] = args.tag
            fh.write(json.dumps(line) + "\n")
            fh.flush()
            print(json.dumps(line))
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
