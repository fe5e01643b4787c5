"""Restatement of the three torch-geometric==1.5.0 conv layers the reference calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  torch-geometric is an
un-vendored dependency (reference requirements.txt:46, torch-scatter 2.0.5 at
:47); its source is not under /root/reference.  What follows restates the
published 1.5.0 algorithm op by op (gather -> message -> scatter), on CPU,
in whatever dtype the inputs carry (fp32 for parity, fp64 for ground truth).
Parameter names/shapes are pinned by the shipped checkpoints
(experiments/checkpoint/weights/Movielenslatest-small/*/BPR/*/run_1/latest.pkl).

Call sites in the reference: models/peagcn.py:16-21, models/peagat.py:16-21,
models/peasage.py:16-21, invoked at models/base.py:137-139.
"""
import math

import torch
from torch import nn
from torch.nn import Parameter
import torch.nn.functional as F


# ---- torch_geometric.nn.inits (1.5.0) --------------------------------------
def glorot(tensor):
    if tensor is not None:
        stdv = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-stdv, stdv)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)


# ---- torch_scatter 2.0.5 reductions (CPU, sequential edge order) -----------
def scatter_add(src, index, dim_size):
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    return out.index_add_(0, index, src)


def scatter_max(src, index, dim_size):
    # every GAT target owns a self-loop, so no segment is ever empty
    out = torch.full((dim_size,) + tuple(src.shape[1:]), float('-inf'), dtype=src.dtype)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_reduce_(0, idx, src, reduce='amax', include_self=True)


# ---- torch_geometric.utils (1.5.0) -----------------------------------------
def remove_self_loops(edge_index):
    mask = edge_index[0] != edge_index[1]
    return edge_index[:, mask]


def add_self_loops(edge_index, num_nodes):
    loop = torch.arange(0, num_nodes, dtype=torch.long).unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop], dim=1)


def add_remaining_self_loops(edge_index, edge_weight, fill_value, num_nodes):
    row, col = edge_index
    mask = row != col
    inv_mask = ~mask
    loop_weight = torch.full((num_nodes,), fill_value, dtype=edge_weight.dtype)
    remaining = edge_weight[inv_mask]
    if remaining.numel() > 0:
        loop_weight[row[inv_mask]] = remaining
    edge_weight = torch.cat([edge_weight[mask], loop_weight], dim=0)
    loop = torch.arange(0, num_nodes, dtype=torch.long).unsqueeze(0).repeat(2, 1)
    edge_index = torch.cat([edge_index[:, mask], loop], dim=1)
    return edge_index, edge_weight


def segment_softmax(src, index, num_nodes):
    out = src - scatter_max(src, index, num_nodes)[index]
    out = out.exp()
    out = out / (scatter_add(out, index, num_nodes)[index] + 1e-16)
    return out


# ---- GCNConv ----------------------------------------------------------------
class GCNConv(nn.Module):
    """GCNConv(in, out) with PyG-1.5.0 defaults (improved=False, cached=False,
    bias=True, normalize=True).  Degree is taken on the SOURCE index
    (``row``) - 1.6.0 switched to the target index; ``deg_side`` keeps both."""

    deg_side = 'source'

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        self.bias = Parameter(torch.Tensor(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.weight)
        zeros(self.bias)

    @classmethod
    def norm(cls, edge_index, num_nodes, dtype):
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype)
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, 1, num_nodes)
        row, col = edge_index
        deg = scatter_add(edge_weight, row if cls.deg_side == 'source' else col, num_nodes)
        deg_inv_sqrt = deg.pow(-0.5)
        deg_inv_sqrt[deg_inv_sqrt == float('inf')] = 0
        return edge_index, deg_inv_sqrt[row] * edge_weight * deg_inv_sqrt[col]

    def forward(self, x, edge_index):
        x = torch.matmul(x, self.weight)
        edge_index, norm = self.norm(edge_index, x.size(0), x.dtype)
        x_j = x.index_select(0, edge_index[0])                 # messages flow row0 -> row1
        out = scatter_add(norm.view(-1, 1) * x_j, edge_index[1], x.size(0))
        return out + self.bias


# ---- GATConv ----------------------------------------------------------------
class GATConv(nn.Module):
    """GATConv(in, out, heads, dropout) with concat=True, negative_slope=0.2."""

    def __init__(self, in_channels, out_channels, heads=1, dropout=0.):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.dropout = 0.2, dropout
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_i = Parameter(torch.Tensor(1, heads, out_channels))
        self.att_j = Parameter(torch.Tensor(1, heads, out_channels))
        self.bias = Parameter(torch.Tensor(heads * out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.lin.weight)
        glorot(self.att_i)
        glorot(self.att_j)
        zeros(self.bias)

    def forward(self, x, edge_index):
        n = x.size(0)
        x = self.lin(x)
        edge_index = add_self_loops(remove_self_loops(edge_index), n)
        j, i = edge_index[0], edge_index[1]
        x_i = x.index_select(0, i).view(-1, self.heads, self.out_channels)
        x_j = x.index_select(0, j).view(-1, self.heads, self.out_channels)
        alpha = (x_i * self.att_i).sum(-1) + (x_j * self.att_j).sum(-1)
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = segment_softmax(alpha, i, n)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        msg = x_j * alpha.view(-1, self.heads, 1)
        out = scatter_add(msg, i, n).view(-1, self.heads * self.out_channels)
        return out + self.bias


# ---- SAGEConv ---------------------------------------------------------------
class SAGEConv(nn.Module):
    """SAGEConv(in, out) with normalize=False, bias=True, aggr='mean'."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_rel = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_root = nn.Linear(in_channels, out_channels, bias=False)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_rel.reset_parameters()
        self.lin_root.reset_parameters()

    def forward(self, x, edge_index):
        n = x.size(0)
        x_j = x.index_select(0, edge_index[0])
        summed = scatter_add(x_j, edge_index[1], n)
        count = scatter_add(torch.ones((edge_index.size(1),), dtype=x.dtype), edge_index[1], n)
        mean = summed / count.clamp(min=1).view(-1, 1)
        out = self.lin_rel(mean)
        out = out + self.lin_root(x)
        return out
