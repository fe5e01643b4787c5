"""Runs the REFERENCE's own source files, unmodified, from where they lie (TEST INFRASTRUCTURE).

``load()`` imports the reference package ``graph_recsys_benchmark`` straight out of the read-only
checkout (default /root/reference; nothing is copied) so that its hot-path code executes as
written:

  models/base.py            GraphRecsysModel.loss / eval (:43-96), PEABaseChannel.forward (:134-140),
                            PEABaseRecsysModel._init / reset_parameters / forward / predict (:147-214)
  models/pea{gcn,gat,sage}.py   the channel stacks (:7-23) and model classes
  utils/rec_utils.py        hit / ndcg / auc (:7-30)
  utils/general_utils.py    update_pea_graph_input (:280-395), save_model / load_model
  solvers.py                BaseSolver.generate_candidates / metrics (:21-104)
  datasets/movielens.py     MovieLens.cf_negative_sampling (:879-997) and __getitem__ (:1135-1182),
                            called unbound on a duck-typed dataset (the processed blobs are not shipped)

What has to be supplied for that to import under numpy 2 / python 3.12 without torch_geometric:

  * ``numpy.int / float / str`` (removed aliases the reference still spells) and
    ``collections.Iterable`` - restored as plain aliases, exactly what they used to be;
  * a stub ``torch_geometric``: ``nn.inits.glorot / zeros`` and ``nn.GCNConv / GATConv / SAGEConv``
    come from ``oracle/pyg150.py`` (torch-geometric==1.5.0 is an un-vendored dependency,
    requirements.txt:46 - those three layers are therefore STILL pinned only by the dense closed
    forms of ``oracle/dense.py``, never by PyG's own code); ``data.download_url / extract_*``
    raise (never called on this path);
  * ``models/__init__.py`` is skipped (it imports the KGAT / NGCF / NFM / MetaPath2Vec baselines,
    which need torch_sparse and torchfm); every other ``__init__`` runs as is.

Everything above the three conv classes - fusion, scoring, BPR and entity-aware loss, eval caching,
ablation, metapath tables, candidate generation, ranking metrics, negative sampling - is thus
checked against the reference's code itself, not against a restatement of it.
"""
import collections
import collections.abc
import importlib
import os
import sys
import types

import numpy as np

REF_ROOT = os.environ.get('PEAGNN_REFERENCE_ROOT', '/root/reference')
_PKG = 'graph_recsys_benchmark'
_loaded = None


def available():
    return os.path.isfile(os.path.join(REF_ROOT, _PKG, 'models', 'base.py'))


def _bare_package(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


def _install_shims():
    for alias, target in (('int', int), ('float', float), ('str', str), ('bool', bool), ('object', object)):
        if alias not in np.__dict__:
            setattr(np, alias, target)
    if not hasattr(collections, 'Iterable'):
        collections.Iterable = collections.abc.Iterable
    if 'torch_geometric' in sys.modules:
        return
    from . import pyg150

    def _no_network(*args, **kwargs):
        raise RuntimeError('raw-data download is outside the hot path (and there is no network)')

    tg = types.ModuleType('torch_geometric')
    tg.__path__ = []
    tg.__version__ = '1.5.0-oracle-stub'
    nn = types.ModuleType('torch_geometric.nn')
    nn.__path__ = []
    inits = types.ModuleType('torch_geometric.nn.inits')
    inits.glorot, inits.zeros = pyg150.glorot, pyg150.zeros
    nn.inits = inits
    nn.GCNConv, nn.GATConv, nn.SAGEConv = pyg150.GCNConv, pyg150.GATConv, pyg150.SAGEConv
    data = types.ModuleType('torch_geometric.data')
    data.download_url = data.extract_zip = data.extract_tar = data.extract_gz = _no_network
    tg.nn, tg.data = nn, data
    sys.modules.update({'torch_geometric': tg, 'torch_geometric.nn': nn, 'torch_geometric.nn.inits': inits,
                        'torch_geometric.data': data})


class Reference(object):
    """Handles to the reference's own objects."""

    def __init__(self):
        root = os.path.join(REF_ROOT, _PKG)
        _install_shims()
        if _PKG in sys.modules and not getattr(sys.modules[_PKG], '_oracle_ref_loader', False):
            raise RuntimeError('another %s is already imported' % _PKG)
        pkg = _bare_package(_PKG, root)               # top-level __init__ only re-exports BaseSolver
        pkg._oracle_ref_loader = True
        _bare_package(_PKG + '.models', os.path.join(root, 'models'))
        imp = importlib.import_module
        self.base = imp(_PKG + '.models.base')
        self.peagcn = imp(_PKG + '.models.peagcn')
        self.peagat = imp(_PKG + '.models.peagat')
        self.peasage = imp(_PKG + '.models.peasage')
        self.datasets = imp(_PKG + '.datasets')
        self.utils = imp(_PKG + '.utils')
        self.rec_utils = imp(_PKG + '.utils.rec_utils')
        self.general_utils = imp(_PKG + '.utils.general_utils')
        self.solvers = imp(_PKG + '.solvers')
        self.model_classes = {'gcn': self.peagcn.PEAGCNRecsysModel, 'gat': self.peagat.PEAGATRecsysModel,
                              'sage': self.peasage.PEASageRecsysModel}

    # -- the three things the experiment scripts do around the package (experiments/pea*_solver_bpr.py) --
    def model_class(self, kind, dataset_args, train_args, name=None):
        """The subclass the reference's experiment script defines (peagcn_solver_bpr.py:104-109):
        binds ``update_graph_input`` to the reference's ``update_pea_graph_input``."""
        base, update = self.model_classes[kind], self.general_utils.update_pea_graph_input

        class Model(base):
            def update_graph_input(self, dataset):
                return update(dataset_args, train_args, dataset)
        # the scripts name their subclasses PEAGCNRecsysModel / PEAGATRecsysModel / MPASAGERecsysModel
        Model.__name__ = name or {'gcn': 'PEAGCNRecsysModel', 'gat': 'PEAGATRecsysModel', 'sage': 'MPASAGERecsysModel'}[kind]
        return Model

    def solver(self, model_class, dataset_args, model_args, train_args):
        return self.solvers.BaseSolver(model_class, dataset_args, model_args, train_args)

    def cf_negative_sampling(self, dataset):
        """datasets/movielens.py:879-997 run on ``dataset`` (any object with the reference's attributes)."""
        return self.datasets.MovieLens.cf_negative_sampling(dataset)

    def getitem(self, dataset, idx):
        """datasets/movielens.py:1135-1182."""
        return self.datasets.MovieLens.__getitem__(dataset, idx)


def load():
    global _loaded
    if _loaded is None:
        if not available():
            raise RuntimeError('reference checkout not found under %s' % REF_ROOT)
        _loaded = Reference()
    return _loaded
