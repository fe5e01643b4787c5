"""How much of a demand-driven step's projection time is per-launch overhead?  13 projections of 12,288 rows each
(the list pass of 13 metapaths) replayed from a CUDA graph - on one stream and on 4 branches - against ONE launch
over the same 159,744 rows; the same for the 62,423-row range pass of 9 metapaths."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_recsys_benchmark_b200 import functional as F_
from graph_recsys_benchmark_b200.engine import _Fork
dev = torch.device('cuda', 0)


def bench(fn, reps=30):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for n, paths in ((12288, 13), (62423, 9)):
    for K, M in ((64, 64), (64, 16), (16, 64)):
        Xs = [torch.randn(n, K, device=dev) for _ in range(paths)]
        Ws = [torch.randn(K, M, device=dev) for _ in range(paths)]
        Ys = [torch.empty(n, M, device=dev) for _ in range(paths)]
        Xall = torch.randn(n * paths, K, device=dev); Yall = torch.empty(n * paths, M, device=dev)

        def serial():
            for X, W, Y in zip(Xs, Ws, Ys):
                F_.linear_raw(X, W, Y, False, None, True)

        def branched():
            with _Fork(dev) as fork:
                for k, (X, W, Y) in enumerate(zip(Xs, Ws, Ys)):
                    with fork.on(k):
                        F_.linear_raw(X, W, Y, False, None, True)

        def single():
            F_.linear_raw(Xall, Ws[0], Yall, False, None, True)
        a, b, c = bench(serial), bench(branched), bench(single)
        print('linear %2d->%2d  %2d x %6d rows: one stream %7.1f us, 4 branches %7.1f us, ONE launch of the same rows %7.1f us' % (K, M, paths, n, a, b, c))
    for K, M in ((64, 64), (64, 16)):
        Xs = [torch.randn(n, K, device=dev) for _ in range(paths)]
        Ds = [torch.randn(n, M, device=dev) for _ in range(paths)]
        dWs = [torch.empty(K, M, device=dev) for _ in range(paths)]
        dbs = [torch.empty(M, device=dev) for _ in range(paths)]
        Xall = torch.randn(n * paths, K, device=dev); Dall = torch.randn(n * paths, M, device=dev)

        def serial():
            for X, D, dW, db in zip(Xs, Ds, dWs, dbs):
                F_.wgrad_raw(X, D, K, M, False, dW, db)

        def branched():
            with _Fork(dev) as fork:
                for k, (X, D, dW, db) in enumerate(zip(Xs, Ds, dWs, dbs)):
                    with fork.on(k):
                        F_.wgrad_raw(X, D, K, M, False, dW, db)

        def single():
            F_.wgrad_raw(Xall, Dall, K, M, False, dWs[0], dbs[0])
        a, b, c = bench(serial), bench(branched), bench(single)
        print('wgrad  %2dx%2d  %2d x %6d rows: one stream %7.1f us, 4 branches %7.1f us, ONE launch of the same rows %7.1f us' % (K, M, paths, n, a, b, c))
