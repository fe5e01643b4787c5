"""Micro-benchmark of the GAT passes (row max, aggregate, backward target side, backward source side) of ONE layer on
the user2item relation of a synthetic HIN, both orientations, through the public autograd function.
Used for ncu captures:
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"Gat|RowMax" \\
      python tools/gat_microbench.py --iters 1 --feat 64"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graph_recsys_benchmark_b200 import functional as F_, _lib              # noqa: E402
from graph_recsys_benchmark_b200.datasets import SyntheticHIN                # noqa: E402
from graph_recsys_benchmark_b200.graph import RelationGraph                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='ml-25m')
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--feat', type=int, default=64)
ap.add_argument('--relation', default='user2item')
args = ap.parse_args()

dev = torch.device('cuda', 0)
ds = SyntheticHIN(args.workload, seed=1234)
ei = torch.from_numpy(np.asarray(ds.edge_index_nps[args.relation])).long().to(dev)
N, F = ds.num_nodes, args.feat
g0 = RelationGraph.from_edge_index(ei, N)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
torch.manual_seed(0)
for name, g in (('dst=item', g0), ('dst=user', g0.transposed())):
    H = (torch.randn(N, F, device=dev) * 0.1).requires_grad_(True)
    ai = (torch.randn(N, 1, device=dev) * 0.1).requires_grad_(True)
    aj = (torch.randn(N, 1, device=dev) * 0.1).requires_grad_(True)
    bias = torch.zeros(F, device=dev, requires_grad=True)
    w = torch.randn(N, F, device=dev)
    per = {}
    for it in range(args.iters + 1):
        flush.zero_()
        _lib.profile = []                  # every C-ABI launch is bracketed by events (graph_recsys_benchmark_b200/_lib.py)
        out = F_.gat_aggregate(H, ai, aj, g, 1, bias, relu=True)
        out.backward(w)
        torch.cuda.synchronize()
        if it > 0:
            for tag, _, e0, e1 in _lib.profile:
                per.setdefault(tag.split('_f')[0].split('_e')[0], []).append(e0.elapsed_time(e1))
        _lib.profile = None
        H.grad = ai.grad = aj.grad = bias.grad = None
    print('%s F=%d nnz=%d  ' % (name, F, g.nnz) + '  '.join('%s %.3f ms' % (k, float(np.median(v))) for k, v in per.items()))
