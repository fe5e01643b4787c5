"""CPU, world_size 2 (gloo): the host-side logic of the row-sharded propagation - shard plan,
edge partition, rank-major layout, all-gather / reduce-scatter autograd, gradient all-reduce.
The CUDA kernels cannot run here; the per-rank aggregation is emulated with the oracle's own
scatter restatement so that 'shard -> aggregate owned rows -> all-gather' can be checked against
the unsharded oracle conv."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import random_edge_index


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


@pytest.mark.parametrize('n,world', [(10, 2), (11, 2), (7, 3), (8, 8), (3, 4)])
def test_shard_plan_is_a_partition(n, world):
    from graph_recsys_benchmark_b200.distributed import ShardPlan
    ids = torch.arange(n)
    seen = []
    for r in range(world):
        plan = ShardPlan(n, world, r)
        own = plan.local_global_ids()
        assert own.numel() == plan.rows_per_rank
        real = own[own >= 0]
        assert ((real % world) == r).all() and (plan.local_index(real) == torch.arange(real.numel())).all()
        seen.append(real)
    assert torch.equal(torch.sort(torch.cat(seen)).values, ids)
    plan = ShardPlan(n, world, 0)
    rm = plan.rank_major_to_global()
    assert rm.numel() == plan.padded
    pos = plan.to_rank_major(ids)
    assert torch.equal(rm[pos], ids) and pos.unique().numel() == n


def test_shard_coo_covers_every_edge_once():
    from graph_recsys_benchmark_b200.distributed import ShardPlan, shard_coo
    n, world = 50, 4
    ei = random_edge_index(n, 600, 3, self_loops=6, multi=40)
    kept = (ei[0] != ei[1])
    total = 0
    for r in range(world):
        plan = ShardPlan(n, world, r)
        src, dl = shard_coo(ei, plan)
        assert ((dl * world + r) < n).all()
        total += src.numel()
        src2, dl2 = shard_coo(ei, plan, explicit_self_loops=True)
        own = plan.local_global_ids()
        assert src2.numel() == src.numel() + int((own >= 0).sum())
    assert total == int(kept.sum())


def _sharded_gcn_via_oracle_math(rank, world):
    """Each rank aggregates its rows with plain torch scatter math (oracle style) over its shard,
    all-gathers, and rank 0 un-permutes; returns the full GCN output."""
    from graph_recsys_benchmark_b200.distributed import ShardPlan, shard_coo, all_gather_rows
    from oracle import pyg150
    torch.manual_seed(0)
    n, f = 37, 8
    ei = random_edge_index(n, 300, 5, self_loops=3, multi=25)
    conv = pyg150.GCNConv(f, 4).double()
    x = torch.randn(n, f, dtype=torch.float64, requires_grad=True)
    plan = ShardPlan(n, world, rank)
    src, dl = shard_coo(ei, plan, explicit_self_loops=True)
    nl = ei[0] != ei[1]
    deg = torch.bincount(ei[0][nl], minlength=n).double() + 1
    dis = deg.pow(-0.5)
    own = plan.local_global_ids().clamp(min=0)
    h = x @ conv.weight
    msg = (dis[src] * dis[own[dl]]).view(-1, 1) * h[src]
    local = torch.zeros(plan.rows_per_rank, 4, dtype=torch.float64).index_add_(0, dl, msg) + conv.bias
    full_rm = all_gather_rows(local)
    full = full_rm.index_select(0, plan.to_rank_major(torch.arange(n)))
    ref = conv(x.detach().clone().requires_grad_(True), ei)
    # gradient path: every rank backpropagates ITS OWN loss (a different weighting per rank); the
    # reduce-scatter in backward must deliver the sum over ranks to the owner of each row.
    w = torch.randn(n, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(100 + rank))
    (full * w).sum().backward()
    g = x.grad.clone()
    dist.all_reduce(g)
    return full.detach(), ref.detach(), g


def test_sharded_aggregation_matches_unsharded_oracle():
    outs = _spawn(_sharded_gcn_via_oracle_math, 2)
    from oracle import pyg150
    for full, ref, _ in outs:
        assert torch.allclose(full, ref, rtol=1e-12, atol=1e-12)
    # reference gradient of sum_r <out, w_r>
    torch.manual_seed(0)
    n, f = 37, 8
    ei = random_edge_index(n, 300, 5, self_loops=3, multi=25)
    conv = pyg150.GCNConv(f, 4).double()
    x = torch.randn(n, f, dtype=torch.float64, requires_grad=True)
    w = sum(torch.randn(n, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(100 + r)) for r in range(2))
    (conv(x, ei) * w).sum().backward()
    assert torch.allclose(outs[0][2], x.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(outs[1][2], x.grad, rtol=1e-10, atol=1e-12)


def _allreduce_grads(rank, world):
    from graph_recsys_benchmark_b200.distributed import allreduce_gradients
    ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(1))]
    ps[0].grad = torch.full((3, 2), float(rank + 1))
    ps[1].grad = torch.arange(5.) * (rank + 1)
    allreduce_gradients(ps)
    return [p.grad.clone() if p.grad is not None else None for p in ps]


def test_gradient_allreduce_sums_over_ranks():
    for grads in _spawn(_allreduce_grads, 2):
        assert torch.equal(grads[0], torch.full((3, 2), 3.0))
        assert torch.equal(grads[1], torch.arange(5.) * 3)
        assert grads[2] is None
