#!/usr/bin/env python
"""Isolates where a row-sharded PEAGCN (world = 1 over gloo: same code path, no real exchange) departs from the
unsharded kernels on the ML-25M-shaped 1/10 graph: per metapath, the first-step aggregate, the first projection and
their gradients for one fixed upstream gradient, then the second step (project -> gather -> aggregate)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from graph_recsys_benchmark_b200 import functional as F_                      # noqa: E402
from graph_recsys_benchmark_b200.datasets import SyntheticHIN                 # noqa: E402
from graph_recsys_benchmark_b200.distributed import ShardPlan, ShardedRelation, shard_aggregate, all_gather_rows   # noqa: E402
from graph_recsys_benchmark_b200.graph import get_graph                       # noqa: E402
from graph_recsys_benchmark_b200.utils.factory import build_model             # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29533')
    world = int(os.environ.get('DIAG_WORLD', '1'))
    dist.init_process_group('gloo', rank=0, world_size=1)
    ds = SyntheticHIN('ml-25m-lite', seed=1234)
    torch.manual_seed(2020)
    model = build_model(ds, 'gcn', device='cuda')
    n = ds.num_nodes
    x = model.x.detach()
    torch.manual_seed(1)
    up64 = torch.randn(n, 64, device='cuda')
    up16 = torch.randn(n, 16, device='cuda')
    for R in (1, 2, 8):
        print('==== shards of', R)
        for p, ch in enumerate(model.pea_channels):
            eil = model.meta_path_edge_index_list[p]
            l0, l1 = ch.gnn_layers
            # ---- unsharded
            xu = x.clone().requires_grad_(True)
            g0 = get_graph(eil[0], n)
            au = F_.gcn_aggregate(xu, g0)
            hu = F_.linear(au, l0.weight, l0.bias, w_is_out_in=False, relu=True)
            l0.weight.grad = l0.bias.grad = None
            (hu * up64).sum().backward()
            ref = dict(a=au.detach(), h=hu.detach(), dw=l0.weight.grad.clone(), db=l0.bias.grad.clone(), dx=xu.grad.clone())
            # ---- every shard of R, stitched back together
            got_a = torch.empty_like(ref['a']); got_h = torch.empty_like(ref['h'])
            dw = torch.zeros_like(ref['dw']); db = torch.zeros_like(ref['db']); dx = torch.zeros_like(ref['dx'])
            for r in range(R):
                plan = ShardPlan(n, R, r)
                srel = ShardedRelation(eil[0], plan, 'gcn')
                xs = x.clone().requires_grad_(True)
                a_s = shard_aggregate(xs, srel, 'orig')
                h_s = F_.linear(a_s, l0.weight, l0.bias, w_is_out_in=False, relu=True)
                own = plan.local_global_ids(device='cuda')
                ok = own >= 0
                l0.weight.grad = l0.bias.grad = None
                (h_s[ok] * up64[own[ok]]).sum().backward()
                got_a[own[ok]] = a_s.detach()[ok]; got_h[own[ok]] = h_s.detach()[ok]
                dw += l0.weight.grad; db += l0.bias.grad; dx += xs.grad
            print('path %2d step 1: A1 %.1e  H1 %.1e  dW1 %.1e  db1 %.1e  dx %.1e   (nnz %d, heavy fwd %d bwd %d)'
                  % (p, rel(got_a, ref['a']), rel(got_h, ref['h']), rel(dw, ref['dw']), rel(db, ref['db']), rel(dx, ref['dx']),
                     g0.nnz, g0.fwd.n_heavy, g0.bwd.n_heavy))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
