"""The three PEAGNN model families (reference models/peagcn.py, models/peagat.py,
models/peasage.py - three copies of one class pair upstream) as ONE table-driven channel:
a family is just the conv layer it stacks.

Layer stack of a channel with S steps (reference peagcn.py:14-21, peagat.py:14-21):
emb -> hidden -> ... -> hidden -> repr, every hidden output `heads * hidden` wide; a multi-step
GAT channel ends in a single-head layer (peagat.py:21), a one-step channel keeps its heads.
Each conv draws its parameters in its constructor and the channel re-draws them once more
(peagcn.py:23), which keeps the torch RNG stream aligned with the reference's constructors.
"""
import torch

from .. import functional as F_
from ..nn import PEAGATConv, PEAGCNConv, PEASageConv
from .base import PEABaseChannel, PEABaseRecsysModel


def channel_layer_dims(num_steps, emb_dim, hidden_size, repr_dim, heads=1):
    """[(in_width, out_channels, heads)] for the S conv layers of a channel."""
    dims, width = [], emb_dim
    for step in range(num_steps):
        is_last = step == num_steps - 1
        out = repr_dim if is_last else hidden_size
        h = 1 if (is_last and num_steps > 1) else heads
        dims.append((width, out, h))
        width = out * h
    return dims


class _GCNLayer(PEAGCNConv):
    shares_aggregate = 'gcn'       # A_hat @ x is weight-free: shared by channels with the same first relation

    @staticmethod
    def aggregate_input(x, g):
        return F_.gcn_aggregate(x, g)


class _SageLayer(PEASageConv):
    shares_aggregate = 'sage'
    keeps_self_loops = True        # SAGEConv treats a self-loop edge as an ordinary edge

    @staticmethod
    def aggregate_input(x, g):
        return F_.sage_mean_aggregate(x, g)


FAMILIES = {
    'gcn': lambda fin, fout, heads, dropout: _GCNLayer(fin, fout),
    'sage': lambda fin, fout, heads, dropout: _SageLayer(fin, fout),
    'gat': lambda fin, fout, heads, dropout: PEAGATConv(fin, fout, heads=heads, dropout=dropout),
}


class PEAChannel(PEABaseChannel):
    """One metapath channel of the given family; kwargs are the model kwargs plus num_steps."""
    family = None

    def __init__(self, **kwargs):
        super(PEAChannel, self).__init__()
        self.num_steps = kwargs['num_steps']
        self.num_nodes = kwargs['num_nodes']
        self.dropout = kwargs['dropout']
        make = FAMILIES[self.family]
        dims = channel_layer_dims(self.num_steps, kwargs['emb_dim'], kwargs['hidden_size'], kwargs['repr_dim'],
                                  kwargs.get('num_heads', 1) if self.family == 'gat' else 1)
        self.gnn_layers = torch.nn.ModuleList(make(fin, fout, h, self.dropout) for fin, fout, h in dims)
        self.reset_parameters()


def _family(name):
    channel = type('PEA%sChannel' % name, (PEAChannel,), {'family': name.lower(), '__doc__': PEAChannel.__doc__})

    class Model(PEABaseRecsysModel):
        channel_class = channel

        def __init__(self, **kwargs):
            kwargs['channel_class'] = self.channel_class
            super(Model, self).__init__(**kwargs)
    Model.__name__ = Model.__qualname__ = 'PEA%sRecsysModel' % name
    return channel, Model


PEAGCNChannel, PEAGCNRecsysModel = _family('GCN')
PEAGATChannel, PEAGATRecsysModel = _family('GAT')
PEASageChannel, PEASageRecsysModel = _family('Sage')
