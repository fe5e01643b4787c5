"""PEAGATConv - drop-in for ``torch_geometric.nn.GATConv`` (1.5.0) as the reference uses it
(models/peagat.py:16-21): concat=True, negative_slope=0.2, self loops removed then one added per
node, softmax over each target's edge multiset with the 1e-16 guard.  Parameters as pinned by the
shipped checkpoints: ``lin.weight[heads*out, in]`` (no bias), ``att_i`` / ``att_j`` [1, heads, out],
``bias[heads*out]``; glorot / zeros.
"""
import torch
from torch.nn import Parameter

from .inits import glorot, zeros
from .. import functional as F_
from ..graph import get_graph


class PEAGATConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, dropout=0.):
        super(PEAGATConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.heads = heads
        self.negative_slope = 0.2
        self.dropout = dropout
        self.lin = torch.nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_i = Parameter(torch.Tensor(1, heads, out_channels))
        self.att_j = Parameter(torch.Tensor(1, heads, out_channels))
        self.bias = Parameter(torch.Tensor(heads * out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.lin.weight)
        glorot(self.att_i)
        glorot(self.att_j)
        zeros(self.bias)

    supports_active = True     # as a channel's last step, it can aggregate only the rows loss() reads
    supports_needed = True     # ... and as an earlier step, only the rows the following steps read (functional.NeededRows)

    def forward(self, x, edge_index, relu=False, graph=None, active=None, needed=None):
        if self.training and self.dropout > 0:
            # every shipped configuration trains with dropout 0 (experiments/peagat_solver_bpr.py:29);
            # attention dropout would need the reference's torch RNG stream reproduced per edge.
            raise NotImplementedError('attention dropout > 0 is not supported by the sm_100a GAT kernels')
        g = graph if graph is not None else get_graph(edge_index, x.size(0))
        h = F_.linear(x, self.lin.weight, None, w_is_out_in=True)
        a_i, a_j = F_.gat_scores(h, self.att_i.view(-1), self.att_j.view(-1), self.heads)
        return F_.gat_aggregate(h, a_i, a_j, g, self.heads, self.bias, relu=relu, active=active, needed=needed)

    def __repr__(self):
        return '{}({}, {}, heads={})'.format(self.__class__.__name__, self.in_channels, self.out_channels, self.heads)
