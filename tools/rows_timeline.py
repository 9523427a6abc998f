"""Per-CTA timeline of the all-rows contraction pt::rows_gemm_kernel (forward of one layer; tuning aid, needs a GPU):
    python tools/rows_timeline.py [F Fo K]
Stamps (ns): RG_STAMP in csrc/agcn_pre_tile.cu."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import agcn_b200
from agcn_b200 import _lib
from agcn_b200.functional import sgc_ll_packed
from oracle.sgcll_oracle import synthetic_molecule_batch

F, Fo, K = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 128, 3)
dev = torch.device("cuda:0")
Xp, Lp, n = synthetic_molecule_batch(1024, 132, seed=1235)
batch = agcn_b200.GraphBatch(n, 132, device=dev)
Ld = torch.from_numpy(np.concatenate([Lp[g, :k, :k].reshape(-1) for g, k in enumerate(n)])).to(dev)
X = torch.relu(torch.randn(batch.total_nodes, F, device=dev))
torch.manual_seed(0)
p = {"weight": torch.randn(F * K, Fo, device=dev) * 0.05, "bias": torch.zeros(Fo, device=dev),
     "M_L": torch.randn(F, F, device=dev) * 0.05, "alpha": torch.ones(1, device=dev)}
cfg = {"F": F, "Fo": Fo, "K": K, "variant": "SGC_LL", "laplacian": "reference_literal", "metric_grad": "reference",
       "activation": "relu"}
flush = torch.empty(64 * 1024 * 1024, device=dev)
with torch.no_grad():
    for _ in range(3):
        sgc_ll_packed(X, Ld, None, p, batch, cfg)
    tiles = 1024
    dbg = torch.zeros(tiles * 128, dtype=torch.int64, device=dev)
    if "--cold" in sys.argv:
        flush.fill_(1.0)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().agcn_fused_debug_set(ctypes.c_void_p(dbg.data_ptr())))
    sgc_ll_packed(X, Ld, None, p, batch, cfg)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().agcn_fused_debug_set(None))
d = dbg.cpu().numpy().reshape(tiles, 128)
used = np.nonzero(d[:, 0] > 0)[0]
t0 = d[used, 0].min()
dur = (d[used, 3] - d[used, 0]) / 1e3
print("CTAs stamped:", len(used), "kernel span us: %.1f" % ((d[used, 3].max() - t0) / 1e3))
print("CTA duration us: min %.1f median %.1f max %.1f" % (dur.min(), np.median(dur), dur.max()))
order = used[np.argsort(-dur)]
for name, t in (("slowest", order[0]), ("median", order[len(order) // 2]), ("fastest", order[-1]), ("cta0", used[0]),
                ("last", used[-1])):
    r = d[t]
    u = lambda k: (r[k] - r[0]) / 1e3 if r[k] > 0 else float("nan")
    print("== %s CTA %d: start %.1f, setup +%.2f, accumulators complete +%.2f, end +%.2f" % (
        name, t, (r[0] - t0) / 1e3, u(1), u(2), u(3)))
    print("   worker warp 0 per item [raw tile landed, operand slot free, handed over]:")
    print("   " + " ".join("[%.2f %.2f %.2f]" % (u(4 + 3 * i), u(5 + 3 * i), u(6 + 3 * i)) for i in range(12) if r[4 + 3 * i] > 0))
    print("   MMA thread per item [operand ready, parameter tile ready, issued]:")
    print("   " + " ".join("[%.2f %.2f %.2f]" % (u(40 + 3 * i), u(41 + 3 * i), u(42 + 3 * i)) for i in range(12) if r[40 + 3 * i] > 0))

# ---- recurrence tiles (ct::cheb_tile_fwd_kernel), slots 100..: start, graph list, loads issued, landed, masks, step 1, step 2
cu = np.nonzero(d[:, 100] > 0)[0]
if len(cu):
    c0 = d[cu, 100].min()
    last = 104 + (K - 1)
    cdur = (d[cu, last] - d[cu, 100]) / 1e3
    print("cheb CTAs stamped:", len(cu), "kernel span us: %.1f" % ((d[cu, last].max() - c0) / 1e3),
          "CTA duration us: min %.1f median %.1f max %.1f" % (cdur.min(), np.median(cdur), cdur.max()))
    co = cu[np.argsort(-cdur)]
    for name, t in (("slowest", co[0]), ("median", co[len(co) // 2]), ("fastest", co[-1]), ("last started", cu[np.argmax(d[cu, 100])])):
        r = d[t]
        print("   %s CTA %d: start %.1f | graph list +%.2f | loads issued +%.2f | landed +%.2f | masks +%.2f | steps %s" % (
            name, t, (r[100] - c0) / 1e3, (r[101] - r[100]) / 1e3, (r[102] - r[100]) / 1e3, (r[103] - r[100]) / 1e3,
            (r[104] - r[100]) / 1e3, " ".join("+%.2f" % ((r[104 + s] - r[100]) / 1e3) for s in range(1, K))))
    starts = np.sort((d[cu, 100] - c0) / 1e3)
    print("   CTA start times us (deciles):", " ".join("%.1f" % starts[int(q * (len(starts) - 1) / 10)] for q in range(11)))
