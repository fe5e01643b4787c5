"""PEAGCN - drop-in for reference ``models/peagcn.py`` (channel :7-23, model :26-29)."""
import torch

from ..nn import PEAGCNConv
from .. import functional as F_
from .base import PEABaseChannel, PEABaseRecsysModel


class _GCNLayer(PEAGCNConv):
    shares_aggregate = 'gcn'

    @staticmethod
    def aggregate_input(x, g):
        return F_.gcn_aggregate(x, g)


class PEAGCNChannel(PEABaseChannel):
    def __init__(self, **kwargs):
        super(PEAGCNChannel, self).__init__()
        self.num_steps = kwargs['num_steps']
        self.num_nodes = kwargs['num_nodes']
        self.dropout = kwargs['dropout']

        self.gnn_layers = torch.nn.ModuleList()
        if kwargs['num_steps'] == 1:
            self.gnn_layers.append(_GCNLayer(kwargs['emb_dim'], kwargs['repr_dim']))
        else:
            self.gnn_layers.append(_GCNLayer(kwargs['emb_dim'], kwargs['hidden_size']))
            for i in range(kwargs['num_steps'] - 2):
                self.gnn_layers.append(_GCNLayer(kwargs['hidden_size'], kwargs['hidden_size']))
            self.gnn_layers.append(_GCNLayer(kwargs['hidden_size'], kwargs['repr_dim']))

        self.reset_parameters()


class PEAGCNRecsysModel(PEABaseRecsysModel):
    def __init__(self, **kwargs):
        kwargs['channel_class'] = PEAGCNChannel
        super(PEAGCNRecsysModel, self).__init__(**kwargs)
