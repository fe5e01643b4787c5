"""PEAGCNConv - drop-in for ``torch_geometric.nn.GCNConv`` (1.5.0) as the reference uses it
(models/peagcn.py:16-21, called at models/base.py:137-139).

Same parameters (``weight[in, out]``, ``bias[out]``; glorot / zeros) and the same result
``D^-1/2 (A + I) D^-1/2 (X W) + b`` with the degree taken on the SOURCE index after
add_remaining_self_loops.  Aggregation and projection commute, so the layer aggregates on the
narrower side: aggregate-then-project when in <= out (the gathered table is then the layer
input - shared by every metapath in the first step), project-then-aggregate otherwise
(16-wide gathers in the last step).
"""
import torch
from torch.nn import Parameter

from .inits import glorot, zeros
from .. import functional as F_
from ..graph import get_graph


class PEAGCNConv(torch.nn.Module):
    def __init__(self, in_channels, out_channels):
        super(PEAGCNConv, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        self.bias = Parameter(torch.Tensor(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.weight)
        zeros(self.bias)

    def forward(self, x, edge_index, relu=False, graph=None, aggregated=None):
        """``relu=True`` fuses the channel's F.relu (models/base.py:138) into the epilogue.
        ``aggregated`` lets a caller that already holds A_hat @ x (shared across metapaths with the
        same first relation) skip the gather."""
        g = graph if graph is not None else get_graph(edge_index, x.size(0))
        if self.in_channels <= self.out_channels:
            a = aggregated if aggregated is not None else F_.gcn_aggregate(x, g)
            return F_.linear(a, self.weight, self.bias, w_is_out_in=False, relu=relu)
        h = F_.linear(x, self.weight, None, w_is_out_in=False)
        return F_.gcn_aggregate(h, g, self.bias, relu=relu)

    # -- split form of the project-then-aggregate order (in > out), used by the model to run ONE
    #    aggregation for all metapaths whose step shares a relation (columns concatenated) --------
    @property
    def splits(self):
        return self.in_channels > self.out_channels

    def project(self, x):
        return F_.linear(x, self.weight, None, w_is_out_in=False)

    def batched_aggregate(self, t_cat, g, biases, relu):
        return F_.gcn_aggregate(t_cat, g, torch.cat(biases), relu=relu)

    def post_bias(self):
        return self.bias

    def finish(self, agg, x, relu):
        return agg

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)
