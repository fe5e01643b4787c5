"""What the reference's experiment scripts do around the package before training starts
(experiments/peagcn_solver_bpr.py:67-84 model_args, :104-109 the model subclass that binds
``update_graph_input``), as two functions - used by bench.py, the experiment CLI and the tests."""
from .general_utils import metapath_table, update_pea_graph_input


def default_model_args(dataset, kind, entity_aware=False, channel_aggr='att', num_heads=1, steps=None,
                       emb_dim=64, hidden=64, repr_dim=16):
    """The ``model_args`` dict of the reference scripts (shipped hyper-parameters: emb 64, hidden 64,
    repr 16, every metapath 2 steps, attentive fusion)."""
    n_paths = len(metapath_table({'dataset': dataset.dataset, 'name': dataset.name}))
    args = {
        'model_type': 'Graph', 'if_use_features': False, 'emb_dim': emb_dim, 'hidden_size': hidden,
        'repr_dim': repr_dim, 'dropout': 0.0, 'meta_path_steps': steps or [2] * n_paths, 'channel_aggr': channel_aggr,
        'entity_aware': entity_aware, 'entity_aware_coff': 0.1, 'num_nodes': dataset.num_nodes, 'dataset': dataset,
    }
    if kind == 'gat':
        args['num_heads'] = num_heads
    return args


def build_model(dataset, kind, device='cuda', **kw):
    """PEA{GCN,GAT,Sage}RecsysModel on ``dataset``'s metapaths, on ``device``."""
    from .. import models
    base = {'gcn': models.PEAGCNRecsysModel, 'gat': models.PEAGATRecsysModel, 'sage': models.PEASageRecsysModel}[kind]
    dargs = {'dataset': dataset.dataset, 'name': dataset.name}
    targs = {'device': device}

    class Model(base):
        def update_graph_input(self, ds):
            return update_pea_graph_input(dargs, targs, ds)
    Model.__name__ = base.__name__
    return Model(**default_model_args(dataset, kind, **kw)).to(device)
