"""Command-line front end of the three PEAGNN experiments on the B200-native package.

The flag names, defaults and the three argument dicts handed to ``BaseSolver`` are those of the
reference's experiments/pea{gcn,gat,sage}_solver_bpr.py (flags :16-55, dicts :67-101, model
subclass :107-109), so the `.ps1` command lines of experiments/scripts/ run unchanged.
Additions: ``--synthetic <shape>`` (the reference downloads and preprocesses MovieLens / Yelp,
which is out of scope here - see datasets/synthetic.py) and ``--loss_sync_every``.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from graph_recsys_benchmark_b200 import models                                   # noqa: E402
from graph_recsys_benchmark_b200.solvers import BaseSolver                       # noqa: E402
from graph_recsys_benchmark_b200.utils import get_folder_path, update_pea_graph_input   # noqa: E402

# (flag, type, default, group); booleans travel as the strings 'true' / 'false' like upstream
FLAGS = [
    ('dataset', str, 'Movielens', 'data'), ('dataset_name', str, 'latest-small', 'data'),
    ('if_use_features', str, 'false', 'data'), ('num_core', int, 10, 'data'), ('num_feat_core', int, 10, 'data'),
    ('sampling_strategy', str, 'random', 'data'), ('entity_aware', str, 'false', 'data'),
    ('synthetic', str, None, 'data'),
    ('dropout', float, 0, 'model'), ('emb_dim', int, 64, 'model'), ('repr_dim', int, 16, 'model'),
    ('hidden_size', int, 64, 'model'), ('meta_path_steps', str, '2,2,2,2,2,2,2,2,2', 'model'),
    ('channel_aggr', str, 'att', 'model'), ('entity_aware_coff', float, 0.1, 'model'),
    ('init_eval', str, 'true', 'train'), ('num_negative_samples', int, 4, 'train'),
    ('num_neg_candidates', int, 99, 'train'), ('device', str, 'cuda', 'train'), ('gpu_idx', str, '0', 'train'),
    ('runs', int, 5, 'train'), ('epochs', int, 30, 'train'), ('batch_size', int, 1024, 'train'),
    ('num_workers', int, 12, 'train'), ('opt', str, 'adam', 'train'), ('lr', float, 0.001, 'train'),
    ('weight_decay', float, 0.001, 'train'), ('early_stopping', int, 20, 'train'),
    ('save_epochs', str, '5,10,15,20,25', 'train'), ('save_every_epoch', int, 26, 'train'),
    ('metapath_test', str, 'true', 'train'), ('loss_sync_every', int, 50, 'train'),
    # not upstream: whole-step CUDA graph replay, negatives / entity columns drawn on the GPU, processed-HIN pickle
    ('cuda_graph', str, 'false', 'train'), ('device_sampling', str, 'false', 'train'),
    ('processed_pickle', str, None, 'data'),
]
MODEL_ONLY_FLAGS = {'PEAGAT': [('num_heads', int, 1, 'model')]}


def _flag(value):
    return str(value).lower() == 'true'


def _ints(csv):
    return [int(tok) for tok in csv.split(',')]


def parse(model_name, argv=None):
    parser = argparse.ArgumentParser(description='%s + BPR on the sm_100a kernels' % model_name)
    for name, typ, default, _ in FLAGS + MODEL_ONLY_FLAGS.get(model_name, []):
        parser.add_argument('--' + name, type=typ, default=default, help='')
    return parser.parse_args(argv)


def argument_dicts(a, model_name):
    """(dataset_args, model_args, train_args) with the reference's keys."""
    if a.device == 'cpu' or not torch.cuda.is_available():
        raise RuntimeError('graph_recsys_benchmark_b200 needs a CUDA device (sm_100a); there is no CPU path')
    entity_aware = _flag(a.entity_aware)
    features = _flag(a.if_use_features)
    data_dir, weights_dir, logger_dir = get_folder_path(model=model_name, dataset=a.dataset + a.dataset_name,
                                                        loss_type='BPR')
    dataset_args = dict(root=data_dir, dataset=a.dataset, name=a.dataset_name, if_use_features=features,
                        num_negative_samples=a.num_negative_samples, num_core=a.num_core,
                        num_feat_core=a.num_feat_core, cf_loss_type='BPR', type='hete',
                        sampling_strategy=a.sampling_strategy, entity_aware=entity_aware, model=model_name,
                        synthetic=a.synthetic, processed_pickle=a.processed_pickle)
    model_args = dict(model_type='Graph', if_use_features=features, emb_dim=a.emb_dim, hidden_size=a.hidden_size,
                      repr_dim=a.repr_dim, dropout=a.dropout, meta_path_steps=_ints(a.meta_path_steps),
                      channel_aggr=a.channel_aggr, entity_aware=entity_aware, entity_aware_coff=a.entity_aware_coff)
    if hasattr(a, 'num_heads'):
        model_args['num_heads'] = a.num_heads
    # the checkpoint / logger folders are keyed by the model arguments, path count instead of the list
    tag = dict(model_args, meta_path_steps=len(model_args['meta_path_steps']))
    tag = str(tag)[:255]
    train_args = dict(init_eval=_flag(a.init_eval), num_negative_samples=a.num_negative_samples,
                      num_neg_candidates=a.num_neg_candidates, opt=a.opt, runs=a.runs, epochs=a.epochs,
                      batch_size=a.batch_size, weight_decay=a.weight_decay, device='cuda:{}'.format(a.gpu_idx),
                      lr=a.lr, num_workers=a.num_workers, weights_folder=os.path.join(weights_dir, tag),
                      logger_folder=os.path.join(logger_dir, tag), save_epochs=_ints(a.save_epochs),
                      save_every_epoch=a.save_every_epoch, metapath_test=_flag(a.metapath_test),
                      loss_sync_every=a.loss_sync_every, cuda_graph=_flag(a.cuda_graph),
                      device_sampling=_flag(a.device_sampling))
    for title, d in (('dataset', dataset_args), ('task', model_args), ('train', train_args)):
        print('{} params: {}'.format(title, d))
    return dataset_args, model_args, train_args


def run(model_name, base_class, argv=None):
    dataset_args, model_args, train_args = argument_dicts(parse(model_name, argv), model_name)

    def update_graph_input(self, dataset):
        return update_pea_graph_input(dataset_args, train_args, dataset)
    # the solver and base model test `__class__.__name__[:3] == 'PEA'`: keep the family's name
    model_class = type(base_class.__name__, (base_class,), {'update_graph_input': update_graph_input})
    solver = BaseSolver(model_class, dataset_args, model_args, train_args)
    solver.run()
    return solver


def main(model_name):
    run(model_name, getattr(models, model_name + 'RecsysModel'))
