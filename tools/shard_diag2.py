#!/usr/bin/env python
"""Follow-up of tools/shard_diag.py for one metapath: where do the relu masks of the first projection differ between
the unsharded and the (single-shard) sharded aggregate, and what do those rows look like?"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_recsys_benchmark_b200 import functional as F_                      # noqa: E402
from graph_recsys_benchmark_b200.datasets import SyntheticHIN                 # noqa: E402
from graph_recsys_benchmark_b200.distributed import ShardPlan, ShardedRelation, shard_aggregate   # noqa: E402
from graph_recsys_benchmark_b200.graph import get_graph                       # noqa: E402
from graph_recsys_benchmark_b200.utils.factory import build_model             # noqa: E402

os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
os.environ.setdefault('MASTER_PORT', '29534')
dist.init_process_group('gloo', rank=0, world_size=1)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ds = SyntheticHIN('ml-25m-lite', seed=1234)
torch.manual_seed(2020)
model = build_model(ds, 'gcn', device='cuda')
n = ds.num_nodes
x = model.x.detach()
ei = model.meta_path_edge_index_list[P][0]
l0 = model.pea_channels[P].gnn_layers[0]
g0 = get_graph(ei, n)
au = F_.gcn_aggregate(x, g0)
srel = ShardedRelation(ei, ShardPlan(n, 1, 0), 'gcn')
a_s = shard_aggregate(x, srel, 'orig')
hu = F_.linear(au, l0.weight.detach(), l0.bias.detach(), w_is_out_in=False, relu=True)
hs = F_.linear(a_s, l0.weight.detach(), l0.bias.detach(), w_is_out_in=False, relu=True)
mism = (hu > 0) != (hs > 0)
rows = torch.nonzero(mism.any(dim=1)).flatten()
print('mask mismatches: %d entries in %d rows' % (int(mism.sum()), rows.numel()))
deg_in = torch.bincount(ei[1], minlength=n)
deg_out = torch.bincount(ei[0], minlength=n)
for r in rows[:12].tolist():
    print('row %6d  in-deg %4d out-deg %5d  |A1_u| %.3e |A1_s| %.3e  |A1_u - A1_s| %.3e  |x| %.3e  dis_u %.4e dis_s %.4e  max|h_u| %.3e max|h_s| %.3e'
          % (r, int(deg_in[r]), int(deg_out[r]), float(au[r].norm()), float(a_s[r].norm()), float((au[r] - a_s[r]).norm()), float(x[r].norm()),
             float(g0.gcn_dis[r]), float(srel.scale_orig[r]), float(hu[r].abs().max()), float(hs[r].abs().max())))
print('rows with A1 == 0 exactly: unsharded %d sharded %d' % (int((au.abs().sum(1) == 0).sum()), int((a_s.abs().sum(1) == 0).sum())))
print('finite: ', bool(torch.isfinite(au).all()), bool(torch.isfinite(a_s).all()))
# transposed aggregation of one fixed gradient
torch.manual_seed(3)
d = torch.randn(n, 64, device='cuda')
dxu = F_.spmm_raw(g0.bwd, d, 64, torch.empty_like(d), g0.gcn_dis, g0.gcn_dis, True)
dxs = F_.spmm_raw(srel.bwd('orig'), d, 64, torch.empty_like(d), srel.scale_orig, srel.scale_local, False)
diff = (dxu - dxs).abs().max(dim=1).values
bad = torch.nonzero(diff > 1e-4 * dxu.abs().max()).flatten()
print('transposed aggregation: rows that differ %d; heavy rows unsharded %d (chunks %d, threshold %d) sharded %d (chunks %d, threshold %d)'
      % (bad.numel(), g0.bwd.n_heavy, g0.bwd.n_chunks, g0.bwd.heavy_threshold, srel.bwd('orig').n_heavy, srel.bwd('orig').n_chunks,
         srel.bwd('orig').heavy_threshold))
for r in bad[:12].tolist():
    print('  row %6d out-deg %5d  |dx_u| %.3e |dx_s| %.3e  diff %.3e' % (r, int(deg_out[r]), float(dxu[r].norm()), float(dxs[r].norm()), float(diff[r])))
# reference by dense torch ops in fp64
src, dst = ei[0], ei[1]
dis = (deg_out.double() + 1).pow(-0.5)
ref = torch.zeros(n, 64, dtype=torch.float64, device='cuda')
ref.index_add_(0, src, d.double()[dst] * (dis[src] * dis[dst])[:, None])
ref += d.double() * (dis * dis)[:, None]
for name, got in (('unsharded', dxu), ('sharded', dxs)):
    print('  %s vs fp64 scatter: %.2e' % (name, float((got.double() - ref).abs().max() / ref.abs().max())))
dist.destroy_process_group()
