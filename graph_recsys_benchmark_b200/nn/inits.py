"""Parameter initialisers the reference imports from torch_geometric.nn.inits (models/base.py:5):
glorot = U(-a, a), a = sqrt(6 / (size(-2) + size(-1))); zeros."""
import math


def glorot(tensor):
    if tensor is not None:
        stdv = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-stdv, stdv)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)
