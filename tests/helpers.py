"""Shared test helpers: small random graphs, oracle/product model pairs."""
import numpy as np
import torch

from oracle import graph as ograph
from oracle.models import OraclePEAModel

KIND_OF = {'PEAGCNRecsysModel': 'gcn', 'PEAGATRecsysModel': 'gat', 'PEASageRecsysModel': 'sage'}


def random_edge_index(n, e, seed, self_loops=0, multi=0, src_range=None, dst_range=None):
    g = torch.Generator().manual_seed(seed)
    s_lo, s_hi = src_range or (0, n)
    d_lo, d_hi = dst_range or (0, n)
    src = torch.randint(s_lo, s_hi, (e,), generator=g)
    dst = torch.randint(d_lo, d_hi, (e,), generator=g)
    if multi:
        src = torch.cat([src, src[:multi]])
        dst = torch.cat([dst, dst[:multi]])
    if self_loops:
        l = torch.randint(0, n, (self_loops,), generator=g)
        src = torch.cat([src, l])
        dst = torch.cat([dst, l])
    return torch.stack([src, dst]).long()


def model_kwargs(dataset, kind, entity_aware=False, channel_aggr='att', num_heads=1, steps=None,
                 emb_dim=64, hidden=64, repr_dim=16):
    from graph_recsys_benchmark_b200.utils import metapath_table
    P = len(metapath_table({'dataset': dataset.dataset, 'name': dataset.name}))
    kw = {
        'model_type': 'Graph', 'if_use_features': False, 'emb_dim': emb_dim, 'hidden_size': hidden,
        'repr_dim': repr_dim, 'dropout': 0.0, 'meta_path_steps': steps or [2] * P, 'channel_aggr': channel_aggr,
        'entity_aware': entity_aware, 'entity_aware_coff': 0.1, 'num_nodes': dataset.num_nodes, 'dataset': dataset,
    }
    if kind == 'gat':
        kw['num_heads'] = num_heads
    return kw


def oracle_model_for(dataset, kind, dtype=torch.float32, **kw):
    """OraclePEAModel on the dataset's metapaths (CPU)."""
    eil = ograph.metapath_edge_index_list(dataset, dataset.dataset, dataset.name)
    mk = model_kwargs(dataset, kind, **kw)
    m = OraclePEAModel(kind, dataset.num_nodes, eil, mk['meta_path_steps'], emb_dim=mk['emb_dim'],
                       hidden_size=mk['hidden_size'], repr_dim=mk['repr_dim'], num_heads=mk.get('num_heads', 1),
                       dropout=0.0, channel_aggr=mk['channel_aggr'], entity_aware=mk['entity_aware'],
                       entity_aware_coff=mk['entity_aware_coff'])
    return m.to(dtype)


def product_model_for(dataset, kind, device='cuda', **kw):
    from graph_recsys_benchmark_b200 import models
    from graph_recsys_benchmark_b200.utils import update_pea_graph_input
    base = {'gcn': models.PEAGCNRecsysModel, 'gat': models.PEAGATRecsysModel, 'sage': models.PEASageRecsysModel}[kind]
    dargs = {'dataset': dataset.dataset, 'name': dataset.name}
    targs = {'device': device}

    class Model(base):
        def update_graph_input(self, ds):
            return update_pea_graph_input(dargs, targs, ds)
    Model.__name__ = base.__name__
    return Model(**model_kwargs(dataset, kind, **kw)).to(device)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
