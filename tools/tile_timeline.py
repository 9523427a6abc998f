"""Per-tile timeline of the fused forward kernel (tuning aid; needs a GPU):
    python tools/tile_timeline.py [F Fo K]
Stamps (ns, relative to the earliest tile start): see FT_STAMP in csrc/agcn_fused_tile.cu."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import agcn_b200
from agcn_b200 import _lib
from agcn_b200.functional import sgc_ll_packed
from oracle import sgcll_oracle as O

F, Fo, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 128, 3)
dev = torch.device("cuda:0")
Xp, Lp, n = O.synthetic_molecule_batch(1024, 132, seed=1235)
batch = agcn_b200.GraphBatch(n, 132, device=dev)
Ld = torch.from_numpy(np.concatenate([Lp[g, :k, :k].reshape(-1) for g, k in enumerate(n)])).to(dev)
X = torch.relu(torch.randn(batch.total_nodes, F, device=dev))
p = {k: v.float().to(dev) for k, v in O.make_params(F, Fo, K, "SGC_LL", seed=3, dtype=torch.float32).items()}
cfg = {"F": F, "Fo": Fo, "K": K, "variant": "SGC_LL", "laplacian": "reference_literal", "metric_grad": "reference",
       "activation": "relu"}
flush = torch.empty(64 * 1024 * 1024, device=dev)
with torch.no_grad():
    for _ in range(3):
        sgc_ll_packed(X, Ld, None, p, batch, cfg)
    tiles = 400
    dbg = torch.zeros(tiles * 128, dtype=torch.int64, device=dev)
    flush.fill_(1.0)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().agcn_fused_debug_set(ctypes.c_void_p(dbg.data_ptr())))
    sgc_ll_packed(X, Ld, None, p, batch, cfg)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().agcn_fused_debug_set(None))
d = dbg.cpu().numpy().reshape(tiles, 128)
used = np.nonzero(d[:, 0] > 0)[0]
t0 = d[used, 0].min()
rel = lambda v: (v - t0) / 1e3 if v > 0 else float("nan")
dur = (d[used, 68] - d[used, 0]) / 1e3
print("tiles stamped:", len(used), "kernel span us: %.1f" % ((d[used, 68].max() - t0) / 1e3))
print("tile duration us: min %.1f median %.1f max %.1f" % (dur.min(), np.median(dur), dur.max()))
order = used[np.argsort(-dur)]
for name, t in (("slowest", order[0]), ("median", order[len(order) // 2]), ("fastest", order[-1]), ("tile0", used[0]),
                ("last", used[-1])):
    r = d[t]
    print("== %s tile %d: start %.1f end %.1f (dur %.1f)" % (name, t, rel(r[0]), rel(r[68]), (r[68] - r[0]) / 1e3))
    print("   prologue done +%.2f" % ((r[1] - r[0]) / 1e3))
    for c in range(8):
        if r[2 + 8 * c] == 0:
            break
        b = r[2 + 8 * c]
        print("   chunk %d: top +%.2f | emit0 +%.2f | mma1 +%.2f emit1 +%.2f | mma2 +%.2f emit2 +%.2f" % (
            c, (b - r[0]) / 1e3, (r[3 + 8 * c] - b) / 1e3, (r[4 + 8 * c] - b) / 1e3, (r[5 + 8 * c] - b) / 1e3,
            (r[6 + 8 * c] - b) / 1e3, (r[7 + 8 * c] - b) / 1e3))
    print("   workers done +%.2f | tmem_full +%.2f | end +%.2f" % ((r[66] - r[0]) / 1e3, (r[67] - r[0]) / 1e3,
                                                                   (r[68] - r[0]) / 1e3))
    print("   mma thread (B ready, A ready, issued) per k-block, us from tile start:")
    print("   " + " ".join("[%.1f %.1f %.1f]" % ((r[72 + 3 * k] - r[0]) / 1e3, (r[73 + 3 * k] - r[0]) / 1e3,
                                                   (r[74 + 3 * k] - r[0]) / 1e3) for k in range(16) if r[72 + 3 * k] > 0))
