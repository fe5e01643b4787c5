"""Device-side BPR batch assembly - opt-in replacement for the host loops of
``cf_negative_sampling`` / ``__getitem__`` (reference datasets/movielens.py:920-940, 1153-1177).

``DeviceBprSampler(dataset, device, seed)`` uploads the interaction list, every user's sorted train items
and (entity-aware runs) the feature lists once; ``rows(row_ids, epoch)`` then produces the ``[B, 3]`` /
``[B, 9]`` int64 batch for any set of row ids of the epoch's (never materialised) ``[E * k, cols]`` table
with one kernel (``peagnn_bpr_rows``, counter-based Philox draws).  Same sampling distributions as the
reference; the reference's host RNG stream itself is only reproduced by the default host sampler.
"""
import numpy as np
import torch

from . import _lib


def _csr_from_lists(lists, count):
    ptr = np.zeros(count + 1, dtype=np.int64)
    src = getattr(lists, '_src', None)
    if src is not None and getattr(lists, '_ptr', None) is not None:     # synthetic.py's lazy CSR-backed lists
        return np.asarray(lists._ptr, dtype=np.int64), np.asarray(src, dtype=np.int64)
    for i in range(count):
        ptr[i + 1] = ptr[i] + len(lists[i])
    nids = np.fromiter((n for i in range(count) for n in lists[i]), dtype=np.int64, count=int(ptr[-1]))
    return ptr, nids


class DeviceBprSampler(object):
    def __init__(self, dataset, device, seed=0):
        self.device = torch.device(device)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.num_neg = int(dataset.num_negative_samples)
        self.strategy = {'random': 0, 'unseen': 1}[dataset.sampling_strategy]
        self.cols = 9 if dataset.entity_aware else 3
        u2i = np.ascontiguousarray(np.asarray(dataset.edge_index_nps['user2item'], dtype=np.int64))
        self.E = int(u2i.shape[1])
        self.user_lo, self.item_lo = int(dataset.type_accs['uid']), int(dataset.type_accs['iid'])
        self.num_users, self.num_items = int(dataset.num_uids), int(dataset.num_iids)
        host = {'u2i': u2i}
        if self.strategy == 1:
            pairs = np.unique(u2i.T, axis=0)                   # sorted by (user, item), duplicates dropped
            counts = np.bincount(pairs[:, 0] - self.user_lo, minlength=self.num_users)
            host['seen_ptr'] = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
            host['seen_items'] = np.ascontiguousarray(pairs[:, 1])
        if self.cols == 9:
            host['ifeat_ptr'], host['ifeat_nids'] = _csr_from_lists(dataset.iid_feat_nids, self.num_items)
            host['ufeat_ptr'], host['ufeat_nids'] = _csr_from_lists(dataset.uid_feat_nids, self.num_users)
            starts = sorted(int(v) for v in dataset.type_accs.values())
            host['type_starts'] = np.asarray(starts + [int(dataset.num_nodes)], dtype=np.int64)
            self.num_types = len(starts)
        else:
            self.num_types = 0
        self.host = host
        self.tables = {k: torch.from_numpy(v).to(self.device) for k, v in host.items()}

    def __len__(self):
        return self.E * self.num_neg

    def permutation(self, epoch, rank=0, world=1):
        """The epoch's visiting order of the table's rows (the reference shuffles twice: randperm in
        cf_negative_sampling, then the DataLoader's sampler - one uniform permutation has the same law).
        Data-parallel runs: every rank draws the same permutation (same seed) and keeps every
        ``world``-th entry starting at ``rank`` - the ranks' slices partition the epoch."""
        g = torch.Generator(device=self.device)
        g.manual_seed((self.seed * 1000003 + int(epoch)) & 0x7FFFFFFFFFFFFFFF)
        order = torch.randperm(len(self), device=self.device, generator=g)
        return order if world == 1 else order[rank::world]

    def rows(self, row_ids, epoch):
        if self.device.type != 'cuda':
            raise RuntimeError('DeviceBprSampler.rows runs peagnn_bpr_rows on a CUDA device; there is no CPU path '
                               '(a CPU sampler object only holds the host tables)')
        row_ids = row_ids.to(self.device, dtype=torch.int64).contiguous()
        out = torch.empty(row_ids.numel(), self.cols, dtype=torch.int64, device=self.device)
        t = self.tables

        def ptr(name):
            return t[name].data_ptr() if name in t else None
        with torch.cuda.device(self.device):
            _lib.call('peagnn_bpr_rows', row_ids.data_ptr(), row_ids.numel(), ptr('u2i'), self.E, self.num_neg,
                      self.seed, int(epoch), self.strategy, self.user_lo, self.item_lo, self.num_items,
                      ptr('seen_ptr'), ptr('seen_items'), self.cols, ptr('ifeat_ptr'), ptr('ifeat_nids'),
                      ptr('ufeat_ptr'), ptr('ufeat_nids'), ptr('type_starts'), self.num_types, out.data_ptr(),
                      torch.cuda.current_stream(self.device).cuda_stream)
        return out
