"""CPU-side tests: host logic of the drop-in boundary and the C-ABI library's exports."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_loads_and_exports_every_declared_symbol():
    from agcn_b200 import _lib
    header = open(os.path.join(ROOT, "include", "agcn_sgcll.h")).read()
    declared = set(re.findall(r"\b(agcn_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), "symbol missing from libagcn_sm100.so: " + name
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().agcn_version() >= 100


def test_compute_calls_fail_loudly_without_a_gpu():
    """No CPU fallback: without a device the plan cannot even be created."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from agcn_b200 import _lib
    n = np.array([4, 5], np.int32)
    handle = ctypes.c_void_p()
    rc = _lib.lib().agcn_plan_create(n.ctypes.data_as(ctypes.c_void_p), 2, 8, None, ctypes.byref(handle))
    assert rc != 0 and _lib.lib().agcn_last_error()
    with pytest.raises(_lib.AgcnError):
        _lib.check(rc)


def test_plan_rejects_bad_sizes():
    from agcn_b200 import _lib
    handle = ctypes.c_void_p()
    n = np.array([4, 0], np.int32)
    assert _lib.lib().agcn_plan_create(n.ctypes.data_as(ctypes.c_void_p), 2, 8, None, ctypes.byref(handle)) == -1
    n = np.array([4, 9], np.int32)
    assert _lib.lib().agcn_plan_create(n.ctypes.data_as(ctypes.c_void_p), 2, 8, None, ctypes.byref(handle)) == -1
    assert b"n_nodes" in _lib.lib().agcn_last_error()


def test_graph_matches_reference_golden():
    from agcn_b200 import Graph, MolGraph
    gold = np.load(os.path.join(GOLD, "graph_laplacian.npz"))
    names = sorted({k.split("/")[0] for k in gold.files} - {"tiny3"})
    for name in names:
        flat = gold[name + "/adj_flat"]
        n = len(gold[name + "/deg"])
        lens, vals = flat[:n], flat[n:]
        adj, pos = [], 0
        for k in lens:
            adj.append([int(v) for v in vals[pos:pos + k]])
            pos += k
        g = MolGraph(np.zeros((n, 3), np.float32), adj)
        assert g.has_Lap and g.n_node == n and g.n_feat == 3
        assert np.array_equal(g.degree_list, gold[name + "/deg"])
        assert np.abs(np.asarray(g.Laplacian.todense()) - gold[name + "/L"]).max() <= 1e-12, name
        assert g.Laplacian.dtype == np.float64 and g.Laplacian.format == "csr"
    tiny = Graph(np.zeros((3, 2), np.float32), [[1], [0, 2], [1]], 4, 0)
    assert tiny.Laplacian is None and not tiny.has_Lap and bool(gold["tiny3/has_Lap"]) is False
    assert MolGraph(np.zeros((4, 1), np.float32), [[1], [0], [3], [2]], smiles="CC").smiles == "CC"


def test_topology_padding_contract():
    from agcn_b200 import GraphTopologyMol, MolGraph
    topo = GraphTopologyMol(3, batch_size=2, max_atom=8, device="cpu")
    g = MolGraph(np.arange(15, dtype=np.float32).reshape(5, 3), [[1], [0, 2], [1, 3], [2, 4], [3]])
    feat, sl = topo.pad_data2sparse(g)
    assert feat.shape == (8, 3) and np.array_equal(feat[:5], g.node_features) and not feat[5:].any()
    assert sl.tolist() == [5, -1] and sl.dtype == np.int32
    Lp, ls = topo.pad_Lap2sparse(g)
    assert Lp.shape == (8, 8) and ls.tolist() == [5, 5] and not Lp[5:].any() and not Lp[:, 5:].any()
    assert np.allclose(Lp[:5, :5], g.Laplacian.todense())


def test_layer_constructor_contract():
    from agcn_b200.layers import SGC_LL, SGC_LL_Reslap, Layer, Dropout
    from agcn_b200.operators import model_operatos as model_ops
    model_ops.reset_uids()
    a = SGC_LL(64, 75, 256, K=3)
    b = SGC_LL(64, 75, 256)
    c = SGC_LL_Reslap(64, 75, 256, save_lap=True, name="mine")
    assert (a.name, b.name, c.name) == ("sgc_ll_1", "sgc_ll_2", "mine")
    assert type(a).__name__ == "SGC_LL" and type(c).__name__ == "SGC_LL_Reslap" and isinstance(c, SGC_LL)
    assert (a.nb_filter, a.n_atom_feature, a.K, b.K, a.save_lap, c.save_lap, a.save_output) == (64, 75, 3, 2, False, True, False)
    assert a.dropout is None and a.bias is True and a.trainable
    with pytest.raises(TypeError):
        SGC_LL(64, 75, 256, unknown_kwarg=1)
    with pytest.raises(ValueError):
        SGC_LL(64, 75, 256, activation="not_an_activation")
    assert Layer(name="x").name == "x" and Dropout(0.5).uses_learning_phase


def test_activations_and_leaky_relu():
    from agcn_b200.operators import activations, model_operatos as model_ops
    x = torch.tensor([-2.0, 0.0, 3.0])
    assert activations.get(None)(x) is x
    assert activations.get("relu")(x).tolist() == [0.0, 0.0, 3.0]
    assert model_ops.relu(x, alpha=0.5).tolist() == [-1.0, 0.0, 3.0]
    assert model_ops.relu(x, alpha=torch.tensor([1.0])).tolist() == [-2.0, 0.0, 3.0]   # alpha = 1: identity
    assert model_ops.relu(x, max_value=2.0).tolist() == [0.0, 0.0, 2.0]
    assert torch.allclose(activations.get("tanh")(x), torch.tanh(x))
    f = lambda t: t
    assert activations.get(f) is f


def test_learning_phase_and_dropout():
    from agcn_b200.layers import Dropout
    from agcn_b200.operators import model_operatos as model_ops
    x = torch.ones(1000)
    model_ops.set_learning_phase(0)
    assert Dropout(0.5)(x) is x
    model_ops.set_learning_phase(1)
    assert Dropout(0.0)(x) is x
    if not torch.cuda.is_available():
        from agcn_b200 import _lib
        with pytest.raises(_lib.AgcnError):          # the mask comes from the CUDA kernel: no CPU path
            Dropout(0.5, seed=1)(x)


def test_block_layers_constructor_contract():
    """BlockEnd / DenseBlockEnd / MLP keep the reference's constructors (blockend.py:25-40, densenet_block.py:21-47,
    MLP.py:19-41): positional order, defaults, kwargs check of the Layer base."""
    from agcn_b200.layers import BlockEnd, DenseBlockEnd, MLP
    b = BlockEnd(1, 64, 128)
    assert (b.block_id, b.res_n_features, b.n_features, b.max_atom, b.batch_size) == (1, 64, 128, 128, 256)
    d = DenseBlockEnd(0, [64, 128], 128, 'relu', max_atom=132, batch_size=8)
    assert d.res_n_features_list == [64, 128] and d.output_n_features == 128 and d.K == 2 and d.max_atom == 132
    with pytest.raises(AssertionError):
        DenseBlockEnd(0, (64, 128), 128)
    m = MLP(64, [32, 48], 75, 16)
    assert (m.output_dim, m.hidden_dims, m.input_dim, m.batch_size, m.bias, m.max_atom) == (64, [32, 48], 75, 16, True, 128)
    with pytest.raises(AssertionError):
        MLP(64, 32, 75, 16)
    with pytest.raises(TypeError):
        BlockEnd(0, 8, 8, bogus=1)
    with pytest.raises(ValueError):
        MLP(8, [8], 8, 4, activation="nope")


def test_initialisers():
    import agcn_b200.layers.graphconv as gc
    gc.DEFAULT_DEVICE[0] = "cpu"
    try:
        w = gc.glorot([225, 64])
        lim = np.sqrt(6.0 / (225 + 64))
        assert w.shape == (225, 64) and float(w.abs().max()) <= lim and w.requires_grad
        assert float(gc.zeros([64]).abs().max()) == 0.0
        t = gc.truncate_normal([1000, 4], stddev=1e-3)
        assert float(t.abs().max()) <= 2e-3 + 1e-9
        layer = gc.SGC_LL(64, 75, 8, K=3)
        layer.build()
        assert set(layer.vars) == {"weight", "bias", "M_L", "alpha"}
        assert layer.vars["weight"].shape == (225, 64) and layer.vars["M_L"].shape == (75, 75)
        assert layer.vars["alpha"].tolist() == [1.0] and layer.vars["alpha"].shape == (1,)
        from agcn_b200.layers import SGC_LL_Reslap
        r = SGC_LL_Reslap(64, 75, 8)
        r.build()
        assert r.vars["beta"].tolist() == [1.0]
    finally:
        gc.DEFAULT_DEVICE[0] = "cuda"


def test_fused_tile_packing_invariants():
    """Work decomposition of the fused tile kernels (agcn_fused_tiles_host, host only): every graph is
    covered exactly once, tiles hold at most 128 rows and AGCN_FUSE_LCAP floats of Laplacians, graphs
    above AGCN_FUSE_MAX_N are cut into 128-row ranges."""
    import ctypes
    from agcn_b200 import _lib
    from oracle import sgcll_oracle as O
    FUSE_MAX_N, LCAP = 64, 8320
    cases = [np.array([132, 4, 5, 18, 33, 64, 65, 17, 96, 31], np.int32),
             np.array([1] * 300, np.int32),
             np.array([1024, 700, 13, 145, 96, 97, 128, 129], np.int32),
             O.synthetic_molecule_batch(1024, 132, seed=1235)[2].astype(np.int32)]
    for n in cases:
        B = len(n)
        tiles, n_ent = ctypes.c_int32(), ctypes.c_int32()
        vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.check(_lib.lib().agcn_fused_tiles_host(vp(n), B, None, 0, None, 0, ctypes.byref(tiles), ctypes.byref(n_ent)))
        gstart = np.zeros(tiles.value + 1, np.int32)
        ent = np.zeros(n_ent.value * 4, np.int32)
        _lib.check(_lib.lib().agcn_fused_tiles_host(vp(n), B, vp(gstart), gstart.size, vp(ent), ent.size,
                                                    ctypes.byref(tiles), ctypes.byref(n_ent)))
        ent = ent.reshape(-1, 4)
        assert gstart[0] == 0 and gstart[-1] == n_ent.value and np.all(np.diff(gstart) >= 1)
        covered = np.zeros(B, np.int64)
        for t in range(tiles.value):
            e = ent[gstart[t]:gstart[t + 1]]
            if e[0, 3] < 0:                                   # row range of a big graph
                assert len(e) == 1 and n[e[0, 0]] > FUSE_MAX_N
                assert 1 <= e[0, 2] <= 128 and e[0, 1] % 128 == 0 and e[0, 1] + e[0, 2] <= n[e[0, 0]]
                covered[e[0, 0]] += e[0, 2]
            else:
                rows, lused = 0, 0
                for g, r0, ng, lbase in e:
                    assert ng == n[g] <= FUSE_MAX_N and r0 == rows and lbase == lused
                    rows += ng
                    lused += ng * (ng | 1)
                    covered[g] += ng
                assert rows <= 128 and lused <= LCAP and len(e) <= 128
        assert np.array_equal(covered, n.astype(np.int64))
        if B == 1024:   # ToxCast-shape batch: tiles are well filled (R / 128 is the lower bound)
            assert tiles.value <= int(np.ceil(n.sum() / 128.0 * 1.08)) + 2, (tiles.value, n.sum() / 128.0)


def test_batch_csr_matches_the_dense_laplacians():
    """Host side of the CSR input path: the batch CSR scattered by hand reproduces the
    dense per-graph matrices; duplicate stored entries are summed like scipy's todense()."""
    import scipy.sparse as sp
    import agcn_b200
    from agcn_b200.graph_topology import batch_csr
    from oracle import sgcll_oracle as O
    rng = np.random.default_rng(3)
    graphs = [agcn_b200.MolGraph(np.zeros((n, 2), np.float32), O.molecule_like_adjacency(rng, n)) for n in (4, 9, 31)]
    mats = [g.Laplacian for g in graphs]
    dup = sp.coo_matrix((np.array([1.0, 2.0, 0.5], np.float32), (np.array([0, 0, 2]), np.array([1, 1, 0]))), shape=(3, 3))
    mats.append(dup)
    indptr, indices, values = batch_csr(mats)
    assert indptr.dtype == np.int32 and indices.dtype == np.int32 and values.dtype == np.float32
    assert indptr[0] == 0 and indptr[-1] == len(indices) == len(values)
    row = 0
    for m in mats:
        n = m.shape[0]
        dense = np.zeros((n, n), np.float32)
        for i in range(n):
            cols = indices[indptr[row]:indptr[row + 1]]
            assert len(set(cols.tolist())) == len(cols)          # one entry per (row, column)
            dense[i, cols] = values[indptr[row]:indptr[row + 1]]
            row += 1
        assert np.array_equal(dense, np.asarray(m.todense(), np.float32))
    assert row == len(indptr) - 1


def test_callable_activation_is_applied_not_dropped():
    """activations.get passes callables through (activations.py:15-53) and the layer calls them (graphconv.py:120):
    a caller's own function must not silently become 'linear'."""
    from agcn_b200.layers import SGC_LL
    from agcn_b200.operators import activations
    calls = []

    def mine(x):
        calls.append(1)
        return x * 2

    layer = SGC_LL(8, 8, 4, activation=mine)
    assert layer.activation is mine and layer._fused_activation() == (False, 'linear')
    assert layer._finish(torch.ones(2, 2), False).tolist() == [[2.0, 2.0], [2.0, 2.0]] and calls
    assert SGC_LL(8, 8, 4, activation=activations.tanh)._fused_activation() == (False, 'linear')
    assert SGC_LL(8, 8, 4, activation='tanh')._fused_activation() == (False, 'linear')
    assert SGC_LL(8, 8, 4, activation=activations.relu)._fused_activation() == (True, 'relu')
    assert SGC_LL(8, 8, 4, activation='relu')._fused_activation() == (True, 'relu')
    assert SGC_LL(8, 8, 4, activation=None)._fused_activation() == (True, 'linear')
    assert SGC_LL(8, 8, 4, activation='linear')._fused_activation() == (True, 'linear')


def test_product_side_generator_matches_the_oracle_copy():
    """bench.py draws its inputs from agcn_b200.synthetic (the product harness never imports oracle/); the molecule
    batch must be the one the oracle-side generator of the tests produces."""
    from agcn_b200 import synthetic
    from oracle import sgcll_oracle as O
    X1, L1, n1 = synthetic.molecule_batch(24, 132, seed=1235)
    X2, L2, n2 = O.synthetic_molecule_batch(24, 132, seed=1235)
    assert np.array_equal(n1, n2) and np.array_equal(X1, X2)
    assert np.abs(L1 - L2).max() <= 1e-6


def test_point_cloud_adjacency_rules_restated_literally():
    """adjacency_mean_rule / adjacency_cutoff_rule against a literal double loop of meshloader.py:264-285 and
    pointcloudloader.py:240-263."""
    from agcn_b200 import synthetic
    rng = np.random.default_rng(0)
    for n, F in ((13, 3), (37, 4), (9, 4)):
        P = rng.standard_normal((n, F)).astype(np.float32)
        dist = [np.linalg.norm(P[i] - P[j]) for i in range(n) for j in range(i + 1)]
        for rule, d_lim in (("mean", np.mean(dist)), ("cut", np.sort(dist)[-int(n * 0.1)])):
            A = np.zeros((n, n), bool)
            for i in range(n):
                for j in range(i + 1, n):
                    if np.linalg.norm(P[i] - P[j]) < d_lim:
                        A[i, j] = A[j, i] = True
            got = synthetic.adjacency_mean_rule(P) if rule == "mean" else synthetic.adjacency_cutoff_rule(P)
            assert np.array_equal(got, A), (n, F, rule)
    # and the dense Laplacian helper against the Graph class (graph_structure.py:85-130)
    from agcn_b200 import Graph
    A = synthetic.adjacency_mean_rule(rng.standard_normal((20, 3)).astype(np.float32))
    adj = [list(np.nonzero(r)[0]) for r in A]
    ref = np.asarray(Graph(np.zeros((20, 3), np.float32), adj, 20, 0).Laplacian.todense())
    assert np.abs(synthetic.laplacian_from_dense_adjacency(A) - ref).max() <= 1e-12


def test_network_oracle_head_and_adam():
    from oracle import network_oracle as NO
    g = torch.Generator().manual_seed(0)
    H = [torch.randn(n, 8, generator=g, dtype=torch.float64) for n in (3, 5)]
    dW, db = torch.randn(8, 6, generator=g, dtype=torch.float64), torch.randn(6, generator=g, dtype=torch.float64)
    hW, hb = torch.randn(6, 4, generator=g, dtype=torch.float64), torch.randn(4, generator=g, dtype=torch.float64)
    t = torch.tensor([[1., 0, 0, 1], [0, 1, 1, 0]], dtype=torch.float64)
    w = torch.ones(2, 4, dtype=torch.float64)
    loss = NO.head_loss(H, dW, db, hW, hb, t, w, 0.5)
    mol = torch.tanh(torch.stack([(h @ dW).sum(0) + h.shape[0] * db for h in H]))     # gather commutes with the dense layer
    ref = torch.nn.functional.binary_cross_entropy_with_logits(mol @ hW + hb, t, reduction="sum") * 0.5
    assert abs(float(loss - ref)) < 1e-12
    # tf Adam: first step moves every parameter by ~lr * sign(g)
    p, m, v = NO.adam_tf(torch.zeros(3, dtype=torch.float64), torch.tensor([1., -2., 0.5], dtype=torch.float64),
                         torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64), 1, lr=0.1)
    assert torch.allclose(p, torch.tensor([-0.1, 0.1, -0.1], dtype=torch.float64), atol=1e-6)


def test_host_point_graph_construction_matches_reference_goldens():
    """agcn_b200.synthetic's vectorised adjacency rules + dense Laplacian against the outputs of the reference's own
    loaders and Graph class (tests/golden/point_graph.npz, generated by executing the reference)."""
    from agcn_b200 import synthetic
    gold = np.load(os.path.join(GOLD, "point_graph.npz"))
    names = sorted({k.split("/")[0] for k in gold.files})
    assert len(names) == 8
    for nm in names:
        P, A, L = gold[nm + "/P"], gold[nm + "/A"].astype(bool), gold[nm + "/L"]
        got = synthetic.adjacency_mean_rule(P) if nm.startswith("mean") else synthetic.adjacency_cutoff_rule(P)
        assert np.array_equal(got, A), nm
        assert np.abs(synthetic.laplacian_from_dense_adjacency(got) - L).max() <= 1e-12, nm
    assert not gold["cut9/A"].any()                       # int(9 * 0.1) == 0 -> threshold 0 -> no edge at all
