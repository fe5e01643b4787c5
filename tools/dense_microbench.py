"""Micro-benchmark of the projection kernels (peagnn_linear / peagnn_linear_wgrad) at the
ML-25M node count.  Prints time, HBM GB/s and FP32 TFLOP/s per launch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_recsys_benchmark_b200 import functional as F_      # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 291120
dev = torch.device('cuda', 0)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


A = torch.randn(N, 64, device=dev)
B = torch.empty_like(A)
t = timeit(lambda: B.copy_(A))
print('torch copy [N,64] fp32        %7.1f us  %7.1f GB/s' % (t * 1e3, 8 * N * 64 / t / 1e6))
t = timeit(lambda: A.sum())
print('torch sum  [N,64] fp32        %7.1f us  %7.1f GB/s' % (t * 1e3, 4 * N * 64 / t / 1e6))
for K, M, masked in [(64, 64, False), (64, 64, True), (64, 16, False), (16, 64, False), (64, 32, False)]:
    X = torch.randn(N, K, device=dev)
    W = torch.randn(K, M, device=dev)
    b = torch.randn(M, device=dev)
    Y = torch.empty(N, M, device=dev)
    mask = torch.randn(N, K, device=dev) if masked else None
    t = timeit(lambda: F_.linear_raw(X, W, Y, False, b, True, False, mask))
    by = 4 * N * (K + M + (K if masked else 0))
    print('linear K=%3d M=%3d mask=%d  %7.1f us  %7.1f GB/s  %6.2f TFLOP/s' % (K, M, masked, t * 1e3, by / t / 1e6, 2 * N * K * M / t / 1e9))
for K, M, masked in [(64, 64, False), (64, 64, True), (64, 16, False), (0, 64, False), (0, 112, False)]:
    X = torch.randn(N, K, device=dev) if K else None
    dY = torch.randn(N, M, device=dev)
    dW = torch.empty(max(K, 1), M, device=dev)
    db = torch.empty(M, device=dev)
    mask = torch.randn(N, M, device=dev) if masked else None
    t = timeit(lambda: F_.wgrad_raw(X, dY, K, M, False, dW if K else None, db, mask))
    by = 4 * N * (K + M + (M if masked else 0))
    print('wgrad  K=%3d M=%3d mask=%d  %7.1f us  %7.1f GB/s  %6.2f TFLOP/s' % (K, M, masked, t * 1e3, by / t / 1e6, 2 * N * K * M / t / 1e9))
