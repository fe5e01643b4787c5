"""Quick timing of the hot projection shapes under the current PEAGNN_DENSE mode, at the row counts a step uses:
291,120 (full propagation), 162,541 / 62,423 (range passes of a demand-driven step), 12,288 (its list passes)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_recsys_benchmark_b200 import functional as F_
dev = torch.device('cuda', 0)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
rows = [int(a) for a in sys.argv[1:]] or [291120]
for N in rows:
    for K, M in ((64, 64), (64, 16), (16, 64)):
        X = torch.randn(N, K, device=dev); W = torch.randn(K, M, device=dev); b = torch.randn(M, device=dev)
        Y = torch.empty(N, M, device=dev)
        ts = []
        for it in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); F_.linear_raw(X, W, Y, False, b, True, False, None); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts[2:])) * 1e3
        print('%-8s N=%6d %2d->%2d %7.1f us  %6.0f GB/s' % (os.environ.get('PEAGNN_DENSE', 'default'), N, K, M, t, 4 * N * (K + M) / t / 1e3))
    for K, M in ((64, 64), (64, 16)):
        X = torch.randn(N, K, device=dev); dY = torch.randn(N, M, device=dev)
        dW = torch.empty(K, M, device=dev); db = torch.empty(M, device=dev)
        ts = []
        for it in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); F_.wgrad_raw(X, dY, K, M, False, dW, db, None); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts[2:])) * 1e3
        print('%-8s N=%6d wgrad %2dx%2d %7.1f us  %6.0f GB/s' % (os.environ.get('PEAGNN_DENSE', 'default'), N, K, M, t, 4 * N * (K + M) / t / 1e3))
