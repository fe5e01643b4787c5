"""Quick timing of the 64->64 / 64->16 projections under the current PEAGNN_DENSE mode."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_recsys_benchmark_b200 import functional as F_
N = 291120
dev = torch.device('cuda', 0)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
for K, M in ((64, 64), (64, 16)):
    X = torch.randn(N, K, device=dev); W = torch.randn(K, M, device=dev); b = torch.randn(M, device=dev)
    Y = torch.empty(N, M, device=dev)
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); F_.linear_raw(X, W, Y, False, b, True, False, None); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(os.environ.get('PEAGNN_DENSE', 'default'), K, M, '%.1f us' % (float(np.median(ts[2:])) * 1e3))
