"""Processed-HIN pickles in the reference's on-disk schema.

The reference's ``MovieLens`` / ``Yelp`` dataset classes end their preprocessing by pickling one
``dataset_property_dict`` (datasets/movielens.py:155-345 builds it, :612-615 loads it back and
sets every key as an attribute).  ``ProcessedHIN`` loads such a file and offers the interface the
solver and the models use (``edge_index_nps``, ``type_accs``, ``num_nodes``, the leave-one-out maps,
BPR negative sampling, entity-aware batches), so a real preprocessed MovieLens / Yelp blob plugs in
where the synthetic HINs do.  ``dump_reference_pickle`` writes any dataset object of this package
in the same schema (round-trip tests; also lets a synthetic graph be fed to the reference code).
The raw-data pipeline itself (CSV parsing, k-core filtering, OMDb scraping) stays out of scope.
"""
import pickle

import numpy as np

from .synthetic import SyntheticHIN

REFERENCE_KEYS = ('num_nodes', 'num_node_types', 'types', 'num_nodes_dict', 'type_accs', 'nid2e_dict',
                  'edge_index_nps', 'test_pos_unid_inid_map', 'neg_unid_inid_map', 'edge_type_dict',
                  'num_edge_types', 'rating_np', 'unique_uids', 'unique_iids')

# entity-aware feature order per target, as datasets/movielens.py:946-992 / datasets/yelp.py build it
_ITEM_FEATURES = {
    'Movielens': ('year2item', 'genre2item', 'actor2item', 'director2item', 'writer2item', 'tag2item',
                  'genome_tag2item'),
    'Yelp': ('stars2item', 'reviewcount2item', 'attributes2item', 'categories2item', 'checkincount2item'),
}
_USER_FEATURES = {
    'Movielens': ('tag2user',),
    'Yelp': ('reviewcount2user', 'friendcount2user', 'fans2user', 'stars2user'),
}


class ProcessedHIN(object):
    """A dataset object backed by a reference-schema ``dataset_property_dict`` pickle."""

    # batch assembly is shared with the synthetic datasets (same reference loops, same RNG use)
    cf_negative_sampling = SyntheticHIN.cf_negative_sampling
    _entity_columns = SyntheticHIN._entity_columns
    get_batch = SyntheticHIN.get_batch
    __getitem__ = SyntheticHIN.__getitem__
    __setitem__ = SyntheticHIN.__setitem__
    __len__ = SyntheticHIN.__len__

    def __init__(self, path, dataset='Movielens', name='latest-small', num_negative_samples=4,
                 sampling_strategy='random', entity_aware=False, cf_loss_type='BPR', **unused):
        if cf_loss_type != 'BPR':
            raise NotImplementedError('the PEAGNN hot path trains with BPR (experiments/pea*_solver_bpr.py)')
        with open(path, 'rb') as f:
            props = pickle.load(f)
        for k, v in props.items():                      # movielens.py:614-615
            setattr(self, k, v)
        self.dataset, self.name, self.path = dataset, name, path
        self.num_negative_samples = num_negative_samples
        self.sampling_strategy = sampling_strategy
        self.entity_aware = entity_aware
        self.cf_loss_type = cf_loss_type
        for t, c in self.num_nodes_dict.items():        # num_uids, num_iids, num_genres, ... (movielens.py:158-170)
            if not hasattr(self, 'num_' + t + 's'):
                setattr(self, 'num_' + t + 's', c)
        ei = self.edge_index_nps
        self.iid_feat_nids = SyntheticHIN._group_features(
            [np.asarray(ei[k]).astype(np.int64) for k in _ITEM_FEATURES[dataset] if k in ei],
            self.type_accs['iid'], self.num_iids)
        self.uid_feat_nids = SyntheticHIN._group_features(
            [np.asarray(ei[k]).astype(np.int64) for k in _USER_FEATURES[dataset] if k in ei],
            self.type_accs['uid'], self.num_uids)
        self.train_data, self.train_data_length = None, 0

    def __repr__(self):
        return 'ProcessedHIN({}, N={})'.format(self.path, self.num_nodes)


def dump_reference_pickle(ds, path):
    """Write ``ds`` (any dataset object of this package) as a reference-schema pickle."""
    props = {}
    for k in REFERENCE_KEYS:
        if hasattr(ds, k):
            props[k] = getattr(ds, k)
    n2e = props.get('nid2e_dict')
    if not isinstance(n2e, dict):                       # synthetic graphs keep it lazy; the schema wants a dict
        props['nid2e_dict'] = {nid: n2e[nid] for nid in range(ds.num_nodes)}
    neg = props.get('neg_unid_inid_map')
    if not isinstance(neg, dict):
        u0 = ds.type_accs['uid']
        props['neg_unid_inid_map'] = {u: list(map(int, neg[u])) for u in range(u0, u0 + ds.num_uids)}
    with open(path, 'wb') as f:
        pickle.dump(props, f)
    return path
