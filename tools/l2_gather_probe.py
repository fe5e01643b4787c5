#!/usr/bin/env python
"""Random-row gather rate of this GPU (peagnn_probe_gather) over table size and row width: the measured
ceiling the aggregation kernels are reported against (bench.py: roofline.l2_gather_ceiling).

    python tools/l2_gather_probe.py            # prints one JSON line per (rows, width)

The ML-25M tables are 291 120 rows x {64, 16..112} fp32 (<= 130 MB): L2-resident on a B200 (126 MB) for
the 64- and 16-wide ones.  A 1.2 M-row table (307 MB at 64 floats) shows the same access pattern from HBM."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from graph_recsys_benchmark_b200 import _lib                       # noqa: E402
from graph_recsys_benchmark_b200.graph import _ptr, _stream        # noqa: E402


def measure(rows, feat, n_idx=23_000_000, zipf=False, reps=10):
    dev = torch.device('cuda')
    table = torch.randn(rows, feat, device=dev)
    g = torch.Generator(device='cpu').manual_seed(1)
    if zipf:
        w = 1.0 / (torch.arange(rows, dtype=torch.float64) + 30.0)
        idx = torch.multinomial(w / w.sum(), n_idx, replacement=True, generator=g).to(torch.int32)
        idx = torch.randperm(rows, generator=g).to(torch.int32)[idx.long()]
    else:
        idx = torch.randint(0, rows, (n_idx,), generator=g, dtype=torch.int32)
    idx = idx.to(dev)
    out = torch.empty(int(_lib.query('peagnn_probe_out_floats')), device=dev)
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)

    def launch():
        _lib.call('peagnn_probe_gather', _ptr(table), table.stride(0), feat, _ptr(idx), idx.numel(), _ptr(out), _stream())
    for _ in range(2):
        launch()
    ms = []
    for _ in range(reps):
        flush.zero_()                        # cold L2: the table has to be pulled from HBM once per launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    warm = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        warm.append(e0.elapsed_time(e1))
    nbytes = n_idx * (4 + 4 * feat)
    return {'rows': rows, 'feat': feat, 'table_mb': rows * feat * 4 / 1e6, 'gathers': n_idx, 'zipf': zipf,
            'cold_l2_ms': float(np.median(ms)), 'cold_l2_gbs': nbytes / np.median(ms) / 1e6,
            'warm_l2_ms': float(np.median(warm)), 'warm_l2_gbs': nbytes / np.median(warm) / 1e6}


if __name__ == '__main__':
    for rows, feat, zipf in ((291120, 64, False), (291120, 64, True), (291120, 16, False), (291120, 128, False),
                             (62423, 64, True), (1200000, 64, False)):
        print(json.dumps(measure(rows, feat, zipf=zipf)), flush=True)
