"""Markdown table of a tools/layer_sweep.py result file, optionally beside an older one.

    python profiles/sweep_table.py profiles/r01_n_layer_sweep.jsonl [profiles/r01_l_layer_sweep_simt.jsonl]
"""
import json
import sys


def load(path):
    out = {}
    for line in open(path):
        d = json.loads(line)
        if "error" not in d:
            out[(d["case"], d["F"], d["K"], d["semantics"])] = d
    return out


new = load(sys.argv[1])
old = load(sys.argv[2]) if len(sys.argv) > 2 else {}
print("| shape | F→Fo | K | semantics | B | mean n | fwd ms | fwd+bwd ms | before | graph-layers/s | tf32 frac | HBM frac | bound |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for k, d in new.items():
    o = old.get(k)
    print("| %s | %d→%d | %d | %s | %d | %.0f | %.3f | %.3f | %s | %.3g | %.1f %% | %.1f %% | %s |" % (
        d["case"], d["F"], d["Fo"], d["K"], "literal" if d["semantics"].startswith("ref") else "paper+full", d["B"],
        d["n_mean"], d["ms_fwd"], d["ms_fwd_bwd"], "%.3f" % o["ms_fwd_bwd"] if o else "", d["graph_layers_per_s"],
        100 * d["fwd_bwd"]["tf32_frac"], 100 * d["fwd_bwd"]["hbm_frac"], d["bound"]))
