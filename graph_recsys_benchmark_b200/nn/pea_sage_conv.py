"""PEASageConv - drop-in for ``torch_geometric.nn.SAGEConv`` (1.5.0) as the reference uses it
(models/peasage.py:16-21): ``lin_rel(mean_{j->i} x_j) + lin_root(x_i)``, no self loops, rows
without in-edges aggregate to 0, normalize=False.  Parameters as pinned by the shipped
checkpoints: ``lin_rel.{weight[out,in], bias[out]}``, ``lin_root.weight[out,in]``; default
nn.Linear initialisation.  mean and lin_rel commute, so the mean runs on min(in, out) columns.
"""
import torch

from .. import functional as F_
from ..graph import get_graph


class PEASageConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels):
        super(PEASageConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.lin_rel = torch.nn.Linear(in_channels, out_channels, bias=True)
        self.lin_root = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x, edge_index, relu=False, graph=None, aggregated=None):
        g = graph if graph is not None else get_graph(edge_index, x.size(0), keep_self_loops=True)
        if self.in_channels <= self.out_channels:
            m = aggregated if aggregated is not None else F_.sage_mean_aggregate(x, g)
            rel = F_.linear(m, self.lin_rel.weight, self.lin_rel.bias, w_is_out_in=True)
        else:
            t = F_.linear(x, self.lin_rel.weight, None, w_is_out_in=True)
            rel = F_.sage_mean_aggregate(t, g, self.lin_rel.bias)
        return F_.linear_accumulate(x, self.lin_root.weight, rel, relu=relu)

    # -- split form (in > out): see PEAGCNConv ------------------------------------------------------
    @property
    def splits(self):
        return self.in_channels > self.out_channels

    def project(self, x):
        return F_.linear(x, self.lin_rel.weight, None, w_is_out_in=True)

    def batched_aggregate(self, t_cat, g, biases, relu):
        return F_.sage_mean_aggregate(t_cat, g, torch.cat(biases))      # relu comes after the root term

    def post_bias(self):
        return self.lin_rel.bias

    def finish(self, agg, x, relu):
        return F_.linear_accumulate(x, self.lin_root.weight, agg.contiguous(), relu=relu)

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)
