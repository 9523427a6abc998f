"""Data-parallel parity on real GPUs: with the graphs of one batch sharded over 2 ranks, the
NCCL all-reduced flat gradient must equal the 1-rank gradient of the union batch, the bucketed all-reduce issued
while backward is still running must equal the single all-reduce bit for bit, and the replicas must stay identical
after the Adam step.  Skipped on boxes with fewer than 2 GPUs (the gloo twin of this test runs on the CPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import agcn_b200
    from agcn_b200.data_parallel import shard_graphs
    from agcn_b200.simple_agcn import SimpleAGCNStep, synthetic_labels
    from oracle import sgcll_oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    B = 192
    X, L, n = O.synthetic_molecule_batch(B, 132, seed=4321)
    tg, w = synthetic_labels(B, 20, 9, "cpu")
    mine = shard_graphs(n, world, rank, balance="work")
    per = len(mine)

    def run(idx, world_size, overlap, n_local):
        batch = agcn_b200.GraphBatch(n[idx], 132, device=dev)
        Xd = batch.pack_nodes(torch.from_numpy(X[idx]).to(dev))
        Ld = batch.pack_lap(torch.from_numpy(L[idx]).to(dev))
        model = SimpleAGCNStep(75, (64, 128, 128, 64), 256, 20, 3, n_local, device=dev, world_size=world_size, seed=3,
                               overlap_allreduce=overlap)
        loss = model.loss_and_grads(Xd, Ld, batch, tg[idx].to(dev), w[idx].to(dev))
        model._all_reduce()
        g = model.flat_grad.detach().clone()
        model.apply_adam()
        torch.cuda.synchronize()
        return float(loss), g, model.flat_params.flat.detach().clone()

    loss_r, g_plain, p_plain = run(mine, world, False, per)
    _, g_over, p_over = run(mine, world, True, per)
    res = {"bucketed_equals_single": bool(torch.equal(g_plain, g_over) and torch.equal(p_plain, p_over))}
    # replicas identical
    other = [torch.empty_like(p_plain) for _ in range(world)]
    dist.all_gather(other, p_plain)
    res["replicas_identical"] = bool(all(torch.equal(o, other[0]) for o in other))
    tl = torch.tensor([loss_r], device=dev, dtype=torch.float64)
    dist.all_reduce(tl)
    if rank == 0:
        # the union batch on one rank (global batch = B): same normalisation, no exchange
        loss_1, g_1, _ = run(np.arange(B), 1, False, B)
        res["loss_sum_ranks"] = float(tl)
        res["loss_union"] = loss_1
        res["grad_rel_err"] = float((g_plain - g_1).abs().max() / g_1.abs().max())
        torch.save(res, os.path.join(out_dir, "res.pt"))
    else:
        torch.save(res, os.path.join(out_dir, "res%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_allreduced_gradient_equals_union_batch(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = torch.load(os.path.join(str(tmp_path), "res.pt"))
    res1 = torch.load(os.path.join(str(tmp_path), "res1.pt"))
    assert res["bucketed_equals_single"] and res1["bucketed_equals_single"]
    assert res["replicas_identical"] and res1["replicas_identical"]
    assert abs(res["loss_sum_ranks"] - res["loss_union"]) <= 1e-5 * abs(res["loss_union"])
    assert res["grad_rel_err"] <= 1e-5, res
