"""Dense-matrix closed forms of the three convs (TEST INFRASTRUCTURE).

An independent derivation used to pin oracle/pyg150.py
on small graphs: ``A[i, j]`` counts edges j -> i (row 0 = source j, row 1 =
target i), multiplicities kept.  O(N^2) memory - small graphs only.
"""
import torch
import torch.nn.functional as F


def adjacency(edge_index, n, dtype):
    a = torch.zeros(n, n, dtype=dtype)
    src, dst = edge_index
    a.index_put_((dst, src), torch.ones(src.numel(), dtype=dtype), accumulate=True)
    return a


def gcn_dense(x, edge_index, weight, bias, deg_side='source'):
    n = x.size(0)
    a = adjacency(edge_index, n, x.dtype)
    a_hat = a - torch.diag(torch.diag(a)) + torch.eye(n, dtype=x.dtype)
    deg = a_hat.sum(dim=0) if deg_side == 'source' else a_hat.sum(dim=1)
    dis = deg.pow(-0.5)
    return (dis.view(-1, 1) * a_hat * dis.view(1, -1)) @ (x @ weight) + bias


def sage_dense(x, edge_index, w_rel, b_rel, w_root):
    a = adjacency(edge_index, x.size(0), x.dtype)
    mean = (a @ x) / a.sum(dim=1).clamp(min=1).view(-1, 1)
    return mean @ w_rel.t() + b_rel + x @ w_root.t()


def gat_dense(x, edge_index, lin_weight, att_i, att_j, bias, heads, negative_slope=0.2):
    n = x.size(0)
    a = adjacency(edge_index, n, x.dtype)
    mult = a - torch.diag(torch.diag(a)) + torch.eye(n, dtype=x.dtype)      # edge multiplicity incl. one self loop
    h = (x @ lin_weight.t()).view(n, heads, -1)
    ai = (h * att_i).sum(-1)                                                # [n, H]
    aj = (h * att_j).sum(-1)
    outs = []
    for hd in range(heads):
        e = F.leaky_relu(ai[:, hd].view(-1, 1) + aj[:, hd].view(1, -1), negative_slope)   # e[i, j]
        e = e.masked_fill(mult == 0, float('-inf'))
        m = e.max(dim=1, keepdim=True).values
        p = torch.exp(e - m) * mult
        alpha = p / (p.sum(dim=1, keepdim=True) + 1e-16)
        outs.append(alpha @ h[:, hd, :])
    return torch.cat(outs, dim=1) + bias
