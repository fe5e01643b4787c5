"""Host-side helpers the solver needs - mirrors of reference ``utils/general_utils.py``:
``update_pea_graph_input`` (:280-395, the metapath -> edge-list tables), ``get_opt_class``
(:30-38), ``save_model`` / ``load_model`` (:40-53, :87-104; same checkpoint dict layout so the
shipped ``latest.pkl`` files load), global logger pickles (:127-136, :182-191), ``instantwrite``
and ``clearcache`` (:270-277).  Dataset construction itself is out of scope (SURVEY.md section 2
rows 10-13); ``load_dataset`` hands back the synthetic HINs of ``datasets/synthetic.py`` or a
caller-supplied dataset object.
"""
import gc
import os
import os.path as osp
import pickle

import numpy as np
import torch

# (relation, flipped) per step; flipped = torch.flip(edge_index, dims=[0]) (general_utils.py:300)
_F, _R = False, True
METAPATHS = {
    ('Movielens', 'latest-small'): [                                   # general_utils.py:300-313
        [('user2item', _F), ('user2item', _R)], [('user2item', _R), ('user2item', _F)],
        [('year2item', _F), ('user2item', _R)], [('actor2item', _F), ('user2item', _R)],
        [('writer2item', _F), ('user2item', _R)], [('director2item', _F), ('user2item', _R)],
        [('genre2item', _F), ('user2item', _R)], [('tag2item', _F), ('user2item', _R)],
        [('tag2user', _F), ('user2item', _F)],
    ],
    ('Movielens', '25m'): [                                            # general_utils.py:335-356
        [('user2item', _F), ('user2item', _R)], [('year2item', _F), ('user2item', _R)],
        [('actor2item', _F), ('user2item', _R)], [('writer2item', _F), ('user2item', _R)],
        [('director2item', _F), ('user2item', _R)], [('genre2item', _F), ('user2item', _R)],
        [('genome_tag2item', _F), ('user2item', _R)], [('tag2user', _R), ('tag2user', _F)],
        [('tag2item', _R), ('tag2user', _F)], [('user2item', _R), ('user2item', _F)],
        [('tag2user', _F), ('user2item', _F)], [('tag2item', _R), ('tag2item', _F)],
        [('tag2user', _R), ('tag2item', _F)],
    ],
    ('Yelp', None): [                                                  # general_utils.py:377-392
        [('user2item', _F), ('user2item', _R)], [('user2item', _R), ('user2item', _F)],
        [('stars2item', _F), ('user2item', _R)], [('reviewcount2item', _F), ('user2item', _R)],
        [('attributes2item', _F), ('user2item', _R)], [('categories2item', _F), ('user2item', _R)],
        [('checkincount2item', _F), ('user2item', _R)], [('reviewcount2user', _F), ('user2item', _F)],
        [('friendcount2user', _F), ('user2item', _F)], [('fans2user', _F), ('user2item', _F)],
        [('stars2user', _F), ('user2item', _F)],
    ],
}


def metapath_table(dataset_args):
    if dataset_args['dataset'] == 'Movielens':
        key = ('Movielens', dataset_args['name'])
    elif dataset_args['dataset'] == 'Yelp':
        key = ('Yelp', None)
    else:
        raise NotImplementedError
    if key not in METAPATHS:
        raise NotImplementedError
    return METAPATHS[key]


def update_pea_graph_input(dataset_args, train_args, dataset):
    """list[P] of list[steps] of LongTensor[2, E] on train_args['device'] (row 0 = source,
    row 1 = target), exactly the lists of reference general_utils.py:280-395.  One tensor per
    relation is created (``.long()`` casts the float64 user2item array, :284) and reused; every
    flipped step is a fresh ``torch.flip`` copy as upstream."""
    table = metapath_table(dataset_args)
    base = {}
    out = []
    for path in table:
        steps = []
        for rel, flipped in path:
            if rel not in base:
                base[rel] = torch.from_numpy(np.asarray(dataset.edge_index_nps[rel])).long().to(train_args['device'])
            steps.append(torch.flip(base[rel], dims=[0]) if flipped else base[rel])
        out.append(steps)
    return out


def get_folder_path(model, dataset, loss_type):
    if dataset[:4] == "Yelp":
        dataset = "Yelp"
    data_folder = osp.join('checkpoint', 'data', dataset)
    weights_folder = osp.join('checkpoint', 'weights', dataset, model, loss_type)
    logger_folder = osp.join('checkpoint', 'loggers', dataset, model, loss_type)
    return tuple(osp.expanduser(osp.normpath(p)) for p in (data_folder, weights_folder, logger_folder))


def get_opt_class(opt):
    if opt.lower() == 'adam':
        return torch.optim.Adam
    elif opt.lower() == 'sgd':
        return torch.optim.SGD
    elif opt.lower() == 'sparseadam':
        return torch.optim.SparseAdam
    else:
        raise NotImplementedError('No such optims!')


def save_model(file_path, model, optim, epoch, rec_metrics, silent=False):
    states = {
        'epoch': epoch,
        'model_states': {'model': model.state_dict()},
        'optim_states': {'optim': optim.state_dict()},
        'rec_metrics': rec_metrics,
    }
    with open(file_path, mode='wb+') as f:
        torch.save(states, f)
    if not silent:
        print("Saved checkpoint_backup '{}'".format(file_path))


_LEGACY_KEYS = (('mpagcn_channels.', 'pea_channels.'), ('mpagat_channels.', 'pea_channels.'),
                ('mpasage_channels.', 'pea_channels.'), ('.gcn_layers.', '.gnn_layers.'),
                ('.gat_layers.', '.gnn_layers.'), ('.sage_layers.', '.gnn_layers.'))


def remap_legacy_state_dict(sd):
    """One shipped checkpoint (PEAGCN, entity_aware False) predates the PEA* renaming and uses
    ``mpagcn_channels.N.gcn_layers.M.*`` keys; map them onto the current names."""
    out = {}
    for k, v in sd.items():
        for old, new in _LEGACY_KEYS:
            k = k.replace(old, new)
        out[k] = v
    return out


def load_model(file_path, model, optim, device):
    if os.path.isfile(file_path):
        checkpoint = torch.load(file_path, map_location=device, weights_only=False)
        epoch = checkpoint['epoch']
        model.load_state_dict(remap_legacy_state_dict(checkpoint['model_states']['model']))
        optim.load_state_dict(checkpoint['optim_states']['optim'])
        rec_metrics = checkpoint['rec_metrics']
        for state in optim.state.values():
            for k, v in state.items():
                if isinstance(v, torch.Tensor):
                    state[k] = v.to(device)
        print("Loaded checkpoint_backup '{}'".format(file_path))
    else:
        print("No checkpoint_backup found at '{}'".format(file_path))
        epoch = 0
        rec_metrics = np.zeros((0, 16)), np.zeros((0, 16)), np.zeros((0, 1)), np.zeros((0, 1)), np.zeros((0, 1))
    return model, optim, epoch, rec_metrics


def save_global_logger(global_logger_filepath, HR_per_run, NDCG_per_run, AUC_per_run, train_loss_per_run,
                       eval_loss_per_run):
    with open(global_logger_filepath, 'wb') as f:
        pickle.dump([HR_per_run, NDCG_per_run, AUC_per_run, train_loss_per_run, eval_loss_per_run], f)


def load_global_logger(global_logger_filepath):
    if os.path.isfile(global_logger_filepath):
        with open(global_logger_filepath, 'rb') as f:
            HRs_per_run, NDCGs_per_run, AUC_per_run, train_loss_per_run, eval_loss_per_run = pickle.load(f)
    else:
        print("No loggers found at '{}'".format(global_logger_filepath))
        HRs_per_run, NDCGs_per_run, AUC_per_run, train_loss_per_run, eval_loss_per_run = \
            np.zeros((0, 16)), np.zeros((0, 16)), np.zeros((0, 1)), np.zeros((0, 1)), np.zeros((0, 1))
    return HRs_per_run, NDCGs_per_run, AUC_per_run, train_loss_per_run, eval_loss_per_run, HRs_per_run.shape[0]


def load_dataset(dataset_args):
    """The reference builds MovieLens / Yelp from raw downloads (out of scope).  Here a dataset is
    either handed over ready-made (``dataset_args['dataset_object']``), loaded from a processed
    pickle in the reference's schema (``dataset_args['processed_pickle']``) or synthesised with the
    reference's schema (``dataset_args['synthetic']`` = a ``datasets.synthetic`` shape name)."""
    if dataset_args.get('dataset_object') is not None:
        return dataset_args['dataset_object']
    if dataset_args.get('processed_pickle'):             # a reference-schema dataset_property_dict blob
        from ..datasets import ProcessedHIN
        return ProcessedHIN(dataset_args['processed_pickle'], dataset=dataset_args['dataset'],
                            name=dataset_args.get('name'),
                            num_negative_samples=dataset_args.get('num_negative_samples', 4),
                            sampling_strategy=dataset_args.get('sampling_strategy', 'random'),
                            entity_aware=dataset_args.get('entity_aware', False),
                            cf_loss_type=dataset_args.get('cf_loss_type', 'BPR'))
    from ..datasets import make_synthetic_dataset
    return make_synthetic_dataset(dataset_args)


def instantwrite(filename):
    filename.flush()
    os.fsync(filename.fileno())


def clearcache():
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
