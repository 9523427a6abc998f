"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE.

Run once in the build container (needs /root/reference, which does not exist on
the GPU box; the produced .npz files are committed and travel):

    python tests/golden/make_golden.py

* ``metric_block.npz``: the reference's ``func`` closure (the learned-metric
  block that runs inside ``tf.py_func``) is pulled out of
  models/layers/graphconv.py:163-209 and graphconv_reslap.py:136-182 with ``ast``
  at run time -- no reference source is copied into this repository -- compiled
  as a free function and executed on seeded inputs.
* ``graph_pool.npz``: the same for the ``func`` closure of models/layers/graphpool.py:91-105.
* ``graph_laplacian.npz``: ``Graph(...).Laplacian`` from the reference's
  models/graph_structure.py, imported as a module from its file.

TensorFlow is not installable here, so the TF part of the layer has no
reference-generated golden; ``layer_regression.npz`` freezes the ORACLE's own
output (fp64) for drift detection only and is labelled as such.
"""
import ast
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def extract_func(path):
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "func":
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, path, "exec"), ns)
            return ns["func"]
    raise RuntimeError("func not found in " + path)


def extract_method(path, name):
    """A (static) method of a class in a reference file as a free function, pulled out with ast at run time."""
    tree = ast.parse(open(path).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise RuntimeError(name + " not found in " + path)


def point_graphs():
    """point_graph.npz: the reference's own graph construction for point clouds, executed here --
    get_adjacency of utils/data_loader/meshloader.py:264-285 (mean-distance rule, ModelNet40) and of
    utils/data_loader/pointcloudloader.py:240-263 (cut-off rule, Sydney), then Graph(...).Laplacian of
    models/graph_structure.py:85-130."""
    adj_mean = extract_method(os.path.join(REF, "utils/data_loader/meshloader.py"), "get_adjacency")
    adj_cut = extract_method(os.path.join(REF, "utils/data_loader/pointcloudloader.py"), "get_adjacency")
    spec = importlib.util.spec_from_file_location("ref_graph_structure", os.path.join(REF, "models/graph_structure.py"))
    gs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gs)
    rng = np.random.default_rng(20261018)
    out = {}
    cases = [("mean13", 13, 3, adj_mean), ("mean50", 50, 3, adj_mean), ("mean64", 64, 3, adj_mean),
             ("mean200", 200, 3, adj_mean), ("cut9", 9, 4, adj_cut), ("cut13", 13, 4, adj_cut),
             ("cut50", 50, 4, adj_cut), ("cut200", 200, 4, adj_cut)]
    for name, n, F, fn in cases:
        P = rng.standard_normal((n, F)).astype(np.float32) * rng.uniform(0.5, 2.0, F).astype(np.float32)
        if F == 4:
            P[:, 3] = rng.integers(0, 256, n) / 255.0
        adj_list, adj_matrix = fn(P)
        g = gs.Graph(P, adj_list, max_deg=n, min_deg=0)
        out[name + "/P"] = P
        out[name + "/A"] = (adj_matrix + adj_matrix.T).astype(np.uint8)      # the reference fills the upper triangle
        out[name + "/L"] = np.asarray(g.Laplacian.todense())
        print(name, "edges", int(adj_matrix.sum()), "of", n * (n - 1) // 2)
    np.savez_compressed(os.path.join(HERE, "point_graph.npz"), **out)


def main():
    from oracle import sgcll_oracle as O

    func_ll = extract_func(os.path.join(REF, "models/layers/graphconv.py"))
    func_rl = extract_func(os.path.join(REF, "models/layers/graphconv_reslap.py"))
    rng = np.random.default_rng(20261017)
    out = {}
    cases = [("tox4", 4, 75, "tox"), ("tox5", 5, 75, "tox"), ("tox18", 18, 75, "tox"),
             ("tox132", 132, 75, "tox"), ("pc50_f3", 50, 3, "xyz"), ("pc13_f4", 13, 4, "xyz"),
             ("hid20_f64", 20, 64, "relu"), ("far6_f8", 6, 8, "far")]
    for name, n, F, kind in cases:
        if kind == "tox":
            x = O.tox21_like_features(rng, n)
        elif kind == "xyz":
            x = rng.standard_normal((n, F)).astype(np.float32)
        elif kind == "relu":
            x = np.maximum(rng.standard_normal((n, F)), 0).astype(np.float32)
        else:  # rows so far apart that exp(-dist) underflows: degree-0 rows
            x = (rng.standard_normal((n, F)) * 400).astype(np.float32)
        lim = np.sqrt(6.0 / (2 * F))
        M = rng.uniform(-lim, lim, (F, F)).astype(np.float32)
        with np.errstate(all="ignore"):
            L1, W1 = func_ll(x, M)
            L2, W2 = func_rl(x, M)
        out[name + "/x"], out[name + "/M"] = x, M
        out[name + "/L_ll"], out[name + "/W_ll"] = L1, W1
        out[name + "/L_rl"], out[name + "/W_rl"] = L2, W2
        print(name, "L==I:", np.array_equal(L1, np.eye(n, dtype=np.float32)), "W max", W1.max())
    np.savez_compressed(os.path.join(HERE, "metric_block.npz"), **out)

    spec = importlib.util.spec_from_file_location("ref_graph_structure", os.path.join(REF, "models/graph_structure.py"))
    gs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gs)
    out = {}
    adjs = {
        "path5": [[1], [0, 2], [1, 3], [2, 4], [3]],
        "ring6": [[1, 5], [0, 2], [1, 3], [2, 4], [3, 5], [4, 0]],
        "star5": [[1, 2, 3, 4], [0], [0], [0], [0]],
        "asym4": [[1], [2], [3], []],                       # one-directional lists are symmetrised
        "isolated5": [[1], [0], [], [4], [3]],              # a node with no neighbour
        "mol18": O.molecule_like_adjacency(rng, 18),
        "mol132": O.molecule_like_adjacency(rng, 132),
    }
    for name, adj in adjs.items():
        n = len(adj)
        g = gs.MolGraph(np.zeros((n, 3), np.float32), adj)
        out[name + "/L"] = np.asarray(g.Laplacian.todense())
        out[name + "/deg"] = g.degree_list
        flat = np.array([len(a) for a in adj] + [v for a in adj for v in a], np.int64)
        out[name + "/adj_flat"] = flat
    g3 = gs.Graph(np.zeros((3, 2), np.float32), [[1], [0, 2], [1]], 4, 0)
    out["tiny3/has_Lap"] = np.array(g3.has_Lap)
    np.savez_compressed(os.path.join(HERE, "graph_laplacian.npz"), **out)

    # GraphPoolMol: the reference's py_func body (graphpool.py:91-105), executed on seeded inputs
    func_pool = extract_func(os.path.join(REF, "models/layers/graphpool.py"))
    out = {}
    pool_cases = {
        "mol18_f75": (O.tox21_like_features(rng, 18), gs.MolGraph(np.zeros((18, 3), np.float32), adjs["mol18"]).Laplacian),
        "isolated5_f4": (rng.standard_normal((5, 4)).astype(np.float32),
                         gs.MolGraph(np.zeros((5, 3), np.float32), adjs["isolated5"]).Laplacian),
        "mol132_f64": (np.maximum(rng.standard_normal((132, 64)), 0).astype(np.float32),
                       gs.MolGraph(np.zeros((132, 3), np.float32), adjs["mol132"]).Laplacian),
    }
    for name, (x, lap) in pool_cases.items():
        Ld = np.asarray(lap.todense()).astype(np.float32)
        out[name + "/x"], out[name + "/L"] = x, Ld
        out[name + "/y"] = func_pool(x, Ld)
    n = 50                                      # dense thresholded point-cloud graph with an all-zero row
    pts = rng.standard_normal((n, 3)).astype(np.float32)
    A = (np.linalg.norm(pts[:, None] - pts[None], axis=-1) < 1.2).astype(np.float32)
    d = 1.0 / np.sqrt(A.sum(1))
    Ld = (np.eye(n, dtype=np.float32) - d[:, None] * A * d[None, :]).astype(np.float32)
    Ld[7, :] = 0.0
    out["cloud50_f3/x"], out["cloud50_f3/L"] = pts, Ld
    out["cloud50_f3/y"] = func_pool(pts, Ld)
    np.savez_compressed(os.path.join(HERE, "graph_pool.npz"), **out)

    # oracle self-regression (NOT a reference output)
    import torch
    reg = {}
    X, L, n_nodes = O.synthetic_molecule_batch(4, 24, seed=7)
    n_nodes = np.minimum(n_nodes, 24)
    for variant in ("SGC_LL", "SGC_LL_Reslap"):
        for lap in ("reference_literal", "paper"):
            p = O.make_params(75, 8, 3, variant, seed=3, dtype=torch.float64)
            Xt, Lt = torch.tensor(X, dtype=torch.float64), torch.tensor(L, dtype=torch.float64)
            Y, RL, RW, LA = O.sgc_ll_batch(Xt, Lt, n_nodes, p, 3, variant, lap, "reference", None)
            key = variant + "/" + lap
            reg[key + "/Y"] = Y.numpy()
            reg[key + "/L_all0"] = LA[0].numpy()
    reg["X"], reg["L"], reg["n_nodes"] = X, L, n_nodes
    np.savez_compressed(os.path.join(HERE, "layer_regression.npz"), **reg)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "point_graphs":
        point_graphs()
    else:
        main()
        point_graphs()
