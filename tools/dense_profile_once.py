"""One launch of every hot projection shape (for an ncu capture of the dense kernels)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_recsys_benchmark_b200 import functional as F_      # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 291120
dev = torch.device('cuda', 0)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
for K, M in [(64, 64), (64, 16), (16, 64)]:
    X = torch.randn(N, K, device=dev)
    W = torch.randn(K, M, device=dev)
    b = torch.randn(M, device=dev)
    Y = torch.empty(N, M, device=dev)
    flush.zero_()
    F_.linear_raw(X, W, Y, False, b, True, False, None)
for K, M in [(64, 64), (64, 16)]:
    X = torch.randn(N, K, device=dev)
    dY = torch.randn(N, M, device=dev)
    dW = torch.empty(K, M, device=dev)
    db = torch.empty(M, device=dev)
    flush.zero_()
    F_.wgrad_raw(X, dY, K, M, False, dW, db, None)
torch.cuda.synchronize()
print('ok')
