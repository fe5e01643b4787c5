"""Oracle graph preparation (TEST INFRASTRUCTURE).

* ``csr_by_key`` - the integer contract the CUDA CSR builder must match bit for
  bit: edges grouped by one endpoint, ties kept in COO order (stable), optional
  removal of self-loop edges (what PyG-1.5.0 ``add_remaining_self_loops`` /
  ``remove_self_loops`` do before a GCN / GAT layer re-adds one loop per node).
* ``metapath_tables`` - restates reference utils/general_utils.py:280-395: the
  per-dataset lists of (relation, flipped?) steps.  ``torch.flip(ei, dims=[0])``
  swaps source and target rows (general_utils.py:300).
"""
import numpy as np
import torch


def csr_by_key(key, val, num_nodes, drop_self_loops=False):
    """Group edges by ``key`` (int64 [E]); returns rowptr[int32 N+1], col (= val of the
    kept edges in grouped order), eid (original COO position of each kept edge)."""
    key = np.asarray(key, dtype=np.int64)
    val = np.asarray(val, dtype=np.int64)
    eid = np.arange(key.shape[0], dtype=np.int64)
    if drop_self_loops:
        keep = key != val
        key, val, eid = key[keep], val[keep], eid[keep]
    order = np.argsort(key, kind='stable')
    counts = np.bincount(key, minlength=num_nodes)[:num_nodes]
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr.astype(np.int32), val[order].astype(np.int32), eid[order].astype(np.int32)


_ML_SMALL = [('user2item', 0), ('user2item', 1)], [('user2item', 1), ('user2item', 0)], \
    [('year2item', 0), ('user2item', 1)], [('actor2item', 0), ('user2item', 1)], \
    [('writer2item', 0), ('user2item', 1)], [('director2item', 0), ('user2item', 1)], \
    [('genre2item', 0), ('user2item', 1)], [('tag2item', 0), ('user2item', 1)], \
    [('tag2user', 0), ('user2item', 0)]

_ML_25M = [('user2item', 0), ('user2item', 1)], [('year2item', 0), ('user2item', 1)], \
    [('actor2item', 0), ('user2item', 1)], [('writer2item', 0), ('user2item', 1)], \
    [('director2item', 0), ('user2item', 1)], [('genre2item', 0), ('user2item', 1)], \
    [('genome_tag2item', 0), ('user2item', 1)], [('tag2user', 1), ('tag2user', 0)], \
    [('tag2item', 1), ('tag2user', 0)], [('user2item', 1), ('user2item', 0)], \
    [('tag2user', 0), ('user2item', 0)], [('tag2item', 1), ('tag2item', 0)], \
    [('tag2user', 1), ('tag2item', 0)]

_YELP = [('user2item', 0), ('user2item', 1)], [('user2item', 1), ('user2item', 0)], \
    [('stars2item', 0), ('user2item', 1)], [('reviewcount2item', 0), ('user2item', 1)], \
    [('attributes2item', 0), ('user2item', 1)], [('categories2item', 0), ('user2item', 1)], \
    [('checkincount2item', 0), ('user2item', 1)], [('reviewcount2user', 0), ('user2item', 0)], \
    [('friendcount2user', 0), ('user2item', 0)], [('fans2user', 0), ('user2item', 0)], \
    [('stars2user', 0), ('user2item', 0)]


def metapath_tables(dataset_name, name):
    if dataset_name == 'Movielens' and name == 'latest-small':
        return _ML_SMALL                                    # general_utils.py:300-313
    if dataset_name == 'Movielens' and name == '25m':
        return _ML_25M                                      # general_utils.py:335-356
    if dataset_name == 'Yelp':
        return _YELP                                        # general_utils.py:377-392
    raise NotImplementedError


def metapath_edge_index_list(dataset, dataset_name, name):
    out = []
    for path in metapath_tables(dataset_name, name):
        steps = []
        for rel, flipped in path:
            ei = torch.from_numpy(dataset.edge_index_nps[rel]).long()
            steps.append(torch.flip(ei, dims=[0]) if flipped else ei)
        out.append(steps)
    return out
