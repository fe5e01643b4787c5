"""Oracle restatement of the PEAGNN model (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows reference models/base.py:29-96 (loss / eval), :129-140 (channel),
:143-214 (model init, fusion, predict) and models/pea{gcn,gat,sage}.py:7-23
(layer stacks).  One class parametrised by the conv family replaces the three
near-identical reference subclasses; the state_dict keys are the reference's.
"""
import torch
from torch import nn
from torch.nn import Parameter
import torch.nn.functional as F

from .pyg150 import GCNConv, GATConv, SAGEConv, glorot


def _make_conv(kind, fin, fout, heads, dropout):
    if kind == 'gcn':
        return GCNConv(fin, fout)
    if kind == 'sage':
        return SAGEConv(fin, fout)
    if kind == 'gat':
        return GATConv(fin, fout, heads=heads, dropout=dropout)
    raise NotImplementedError(kind)


class OracleChannel(nn.Module):
    """One metapath channel: emb -> hidden -> ... -> repr (peagcn.py:14-21)."""

    def __init__(self, kind, num_steps, emb_dim, hidden_size, repr_dim, num_heads=1, dropout=0.):
        super().__init__()
        self.num_steps = num_steps
        h = num_heads if kind == 'gat' else 1
        layers = nn.ModuleList()
        if num_steps == 1:
            layers.append(_make_conv(kind, emb_dim, repr_dim, h, dropout))
        else:
            layers.append(_make_conv(kind, emb_dim, hidden_size, h, dropout))
            for _ in range(num_steps - 2):
                layers.append(_make_conv(kind, hidden_size * h, hidden_size, h, dropout))
            # the last GAT layer always has a single head (peagat.py:21)
            layers.append(_make_conv(kind, hidden_size * h, repr_dim, 1, dropout))
        self.gnn_layers = layers
        self.reset_parameters()                      # peagcn.py:23 (second draw)

    def reset_parameters(self):
        for m in self.gnn_layers:
            m.reset_parameters()

    def forward(self, x, edge_index_list):
        assert len(edge_index_list) == self.num_steps          # base.py:135
        for s in range(self.num_steps - 1):
            x = F.relu(self.gnn_layers[s](x, edge_index_list[s]))
        return self.gnn_layers[-1](x, edge_index_list[-1])


class OraclePEAModel(nn.Module):
    """PEABaseRecsysModel + GraphRecsysModel (base.py:29-96,143-214)."""

    def __init__(self, kind, num_nodes, meta_path_edge_index_list, meta_path_steps,
                 emb_dim=64, hidden_size=64, repr_dim=16, num_heads=1, dropout=0.,
                 channel_aggr='att', entity_aware=False, entity_aware_coff=0.1):
        super().__init__()
        assert len(meta_path_edge_index_list) == len(meta_path_steps)   # base.py:162
        self.kind = kind
        self.entity_aware, self.entity_aware_coff = entity_aware, entity_aware_coff
        self.channel_aggr = channel_aggr
        self.x = Parameter(torch.Tensor(num_nodes, emb_dim))            # base.py:156
        self.meta_path_edge_index_list = meta_path_edge_index_list
        self.pea_channels = nn.ModuleList(
            OracleChannel(kind, s, emb_dim, hidden_size, repr_dim, num_heads, dropout)
            for s in meta_path_steps)
        if channel_aggr == 'att':
            self.att = Parameter(torch.Tensor(1, len(meta_path_steps), repr_dim))
        self.fc1 = nn.Linear(2 * repr_dim, repr_dim)
        self.fc2 = nn.Linear(repr_dim, 1)
        self.cached_repr = None
        self.reset_parameters()

    def reset_parameters(self):                                         # base.py:181-189
        glorot(self.x)
        for ch in self.pea_channels:
            ch.reset_parameters()
        glorot(self.fc1.weight)
        glorot(self.fc2.weight)
        if self.channel_aggr == 'att':
            glorot(self.att)

    def channel_outputs(self):
        return [ch(self.x, self.meta_path_edge_index_list[p]) for p, ch in enumerate(self.pea_channels)]

    def forward(self, metapath_idx=None):                               # base.py:191-206
        zs = [z.unsqueeze(1) for z in self.channel_outputs()]
        if metapath_idx is not None:
            zs[metapath_idx] = torch.zeros_like(zs[metapath_idx])
        z = torch.cat(zs, dim=1)
        if self.channel_aggr == 'mean':
            return z.mean(dim=1)
        if self.channel_aggr == 'att':
            w = F.softmax(torch.sum(z * self.att, dim=-1), dim=-1).unsqueeze(-1)
            return torch.sum(z * w, dim=1)
        raise NotImplementedError('Other aggr methods not implemeted!')

    def predict(self, unids, inids):                                    # base.py:208-214
        u, i = self.cached_repr[unids], self.cached_repr[inids]
        return self.fc2(F.relu(self.fc1(torch.cat([u, i], dim=-1))))

    def loss(self, t):                                                  # base.py:43-80
        if self.training:
            self.cached_repr = self.forward()
        pos = self.predict(t[:, 0], t[:, 1])
        neg = self.predict(t[:, 0], t[:, 2])
        cf = -(pos - neg).sigmoid().log().sum()
        if not (self.entity_aware and self.training):
            return cf
        # op for op as base.py:56-76 (every x[...] is its own gather, so the gradient of x accumulates
        # in the reference's order)
        x = self.x
        item_pos_reg = (x[t[:, 1]] - x[t[:, 3]]) * (x[t[:, 1]] - x[t[:, 3]])
        item_neg_reg = (x[t[:, 1]] - x[t[:, 4]]) * (x[t[:, 1]] - x[t[:, 4]])
        item_pos_reg = item_pos_reg.sum(dim=-1)
        item_neg_reg = item_neg_reg.sum(dim=-1)
        user_pos_reg = (x[t[:, 0]] - x[t[:, 6]]) * (x[t[:, 0]] - x[t[:, 6]])
        user_neg_reg = (x[t[:, 0]] - x[t[:, 7]]) * (x[t[:, 0]] - x[t[:, 7]])
        user_pos_reg = user_pos_reg.sum(dim=-1)
        user_neg_reg = user_neg_reg.sum(dim=-1)
        item_reg = -((item_pos_reg - item_neg_reg) * t[:, 5]).sigmoid().log().sum()
        user_reg = -((user_pos_reg - user_neg_reg) * t[:, 8]).sigmoid().log().sum()
        reg = item_reg + user_reg
        return cf + self.entity_aware_coff * reg

    def eval(self, metapath_idx=None):                                  # base.py:88-96
        super().eval()
        with torch.no_grad():
            self.cached_repr = self.forward(metapath_idx)
        return self
