from .pea_gcn_conv import PEAGCNConv
from .pea_gat_conv import PEAGATConv
from .pea_sage_conv import PEASageConv

__all__ = ['PEAGCNConv', 'PEAGATConv', 'PEASageConv']
