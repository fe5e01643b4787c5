"""Micro-benchmark of the aggregation kernel on the user2item relation of a synthetic HIN
(both orientations), for the widths the PEAGNN steps use.  Prints algorithmic GB/s per launch.
Used for ncu captures:  ncu --set full -k regex:csr_rows_kernel ... python tools/spmm_microbench.py --iters 2"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from graph_recsys_benchmark_b200 import functional as F_                     # noqa: E402
from graph_recsys_benchmark_b200.datasets import SyntheticHIN                # noqa: E402
from graph_recsys_benchmark_b200.graph import RelationGraph                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='ml-25m')
ap.add_argument('--iters', type=int, default=20)
ap.add_argument('--widths', default='64,16,112')
ap.add_argument('--relation', default='user2item')
args = ap.parse_args()

dev = torch.device('cuda', 0)
ds = SyntheticHIN(args.workload, seed=1234)
ei = torch.from_numpy(np.asarray(ds.edge_index_nps[args.relation])).long().to(dev)
N = ds.num_nodes
g = RelationGraph.from_edge_index(ei, N)
dis = g.gcn_dis
print('N=%d nnz=%d heavy fwd=%d (chunks %d) bwd=%d (chunks %d)' % (N, g.nnz, g.fwd.n_heavy, g.fwd.n_chunks,
                                                                    g.bwd.n_heavy, g.bwd.n_chunks))
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for F in [int(w) for w in args.widths.split(',')]:
    X = torch.randn(N, F, device=dev)
    out = torch.empty_like(X)
    for name, csr in (('fwd(dst=item)', g.fwd), ('bwd(dst=user)', g.bwd)):
        nbytes = F_.spmm_algorithmic_bytes(csr.nnz, N, F, True, True, True)
        for _ in range(2):
            F_.spmm_raw(csr, X, F, out, dis, dis, True)
        ts = []
        for _ in range(args.iters):
            flush.zero_()                       # evict L2 between iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            F_.spmm_raw(csr, X, F, out, dis, dis, True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts))
        print('F=%3d %-14s %8.3f ms  %7.1f GB/s algorithmic (%.2f of 6540.5)  [min %.3f ms]'
              % (F, name, t, nbytes / t / 1e6, nbytes / t / 1e6 / 6540.5, min(ts)))
