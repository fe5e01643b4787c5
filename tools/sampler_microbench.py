"""Throughput of the device-side BPR row sampler (peagnn_bpr_rows) at the ML-25M shape."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_recsys_benchmark_b200.datasets import SyntheticHIN          # noqa: E402
from graph_recsys_benchmark_b200.sampling import DeviceBprSampler      # noqa: E402

for strategy, entity in (('random', False), ('unseen', False), ('unseen', True)):
    ds = SyntheticHIN('ml-25m', seed=1234, sampling_strategy=strategy, entity_aware=entity)
    t0 = time.perf_counter()
    smp = DeviceBprSampler(ds, 'cuda', seed=1)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    perm = smp.permutation(1)
    for B in (4096, 1 << 20):
        ids = perm[:B]
        for _ in range(3):
            smp.rows(ids, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            smp.rows(ids, 1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print('%-7s entity=%d  B=%8d  %8.3f ms/batch  %10.1f M rows/s   (tables built in %.1f s, epoch table %d rows)'
              % (strategy, entity, B, ms, B / ms / 1e3, t_build, len(smp)))
