"""World-size-2 gloo test (CPU) of the data-parallel host logic: graph sharding and the single flat
gradient all-reduce, with the loss normalised by the GLOBAL batch (multitask_classifier.py:207-208)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sgcll_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _toy_loss(params, X, L, n_nodes, idx, global_batch):
    """Oracle SGC-LL layer on the graphs `idx`; sum of outputs / global batch."""
    loss = torch.zeros((), dtype=torch.float64)
    for g in idx:
        n = int(n_nodes[g])
        y, _, _, _ = O.sgc_ll_graph(torch.tensor(X[g, :n], dtype=torch.float64),
                                    torch.tensor(L[g, :n, :n], dtype=torch.float64), params, 3)
        loss = loss + torch.relu(y).sum()
    return loss / global_batch


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from agcn_b200.data_parallel import FlatGradBuffer, shard_graphs
    X, L, n_nodes = O.synthetic_molecule_batch(12, 40, seed=5)
    n_nodes = np.minimum(n_nodes, 40)
    params = {k: v.requires_grad_(True) for k, v in O.make_params(75, 8, 3, "SGC_LL", seed=1, dtype=torch.float64).items()}
    buf = FlatGradBuffer(list(params.values()))
    idx = shard_graphs(n_nodes, world, rank, balance="work")
    buf.zero()
    _toy_loss(params, X, L, n_nodes, idx, len(n_nodes)).backward()
    assert params["weight"].grad.data_ptr() == buf.flat.data_ptr()      # still a view after backward
    buf.all_reduce()
    out[rank] = (buf.flat.clone(), idx)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    from agcn_b200.data_parallel import FlatGradBuffer
    X, L, n_nodes = O.synthetic_molecule_batch(12, 40, seed=5)
    n_nodes = np.minimum(n_nodes, 40)
    params = {k: v.requires_grad_(True) for k, v in O.make_params(75, 8, 3, "SGC_LL", seed=1, dtype=torch.float64).items()}
    buf = FlatGradBuffer(list(params.values()))
    _toy_loss(params, X, L, n_nodes, range(12), 12).backward()
    g0, idx0 = out[0]
    g1, idx1 = out[1]
    assert torch.equal(g0, g1)                                           # every rank holds the same sum
    assert torch.allclose(g0, buf.flat, rtol=1e-12, atol=1e-14)          # == the single-process gradient
    assert sorted(list(idx0) + list(idx1)) == list(range(12))            # disjoint cover of the batch


def test_shard_graphs_balance():
    from agcn_b200.data_parallel import shard_graphs
    rng = np.random.default_rng(0)
    n = np.exp(rng.uniform(np.log(13), np.log(1024), 128)).astype(np.int64)     # Sydney-shape ragged sizes
    for world in (2, 4, 8):
        shards = [shard_graphs(n, world, r, "work") for r in range(world)]
        assert sorted(np.concatenate(shards).tolist()) == list(range(128))
        assert all(len(s) == 128 // world for s in shards)
        work = np.array([(n[s] ** 2).sum() for s in shards], dtype=np.float64)
        assert work.max() / work.mean() < 1.10                                  # within 10 % of perfect balance
        blocks = [shard_graphs(n, world, r, "count") for r in range(world)]
        assert np.concatenate(blocks).tolist() == list(range(128))


def test_flat_param_buffer_single_tensor_adam_matches_per_tensor_adam():
    """FlatParamBuffer + FlatGradBuffer: one-tensor Adam over the flat views == Adam over the separate
    parameters; parameters keep their identity and see the update."""
    from agcn_b200.data_parallel import FlatGradBuffer, FlatParamBuffer
    torch.manual_seed(0)
    shapes = [(5, 3), (3,), (1,), (4, 4)]
    ref = [torch.randn(*s, dtype=torch.float64, requires_grad=True) for s in shapes]
    ours = [t.detach().clone().requires_grad_(True) for t in ref]
    grads = FlatGradBuffer(ours, direct=ours[:2])
    flat = FlatParamBuffer(ours, grads)
    assert ours[0]._agcn_grad_out.data_ptr() == grads.flat.data_ptr() and not hasattr(ours[2], "_agcn_grad_out")
    opt_ref = torch.optim.Adam(ref, lr=1e-2, eps=1e-7)
    opt = torch.optim.Adam([flat.leaf], lr=1e-2, eps=1e-7)
    for step in range(3):
        gs = [torch.randn(*s, dtype=torch.float64) for s in shapes]
        for p, g in zip(ref, gs):
            p.grad = g.clone()
        grads.zero()
        for p, g in zip(ours, gs):
            p.grad.add_(g)                 # what autograd / the in-place SGC-LL backward do
        opt_ref.step()
        opt.step()
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=1e-12, atol=1e-14)
            assert a.data_ptr() >= flat.flat.data_ptr()
