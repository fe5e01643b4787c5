"""Synthetic heterogeneous information networks with the reference's dataset schema.

The reference builds its HINs from raw MovieLens / Yelp downloads plus OMDb scraping
(datasets/movielens.py:17-583, datasets/yelp.py:19-437 - out of scope).  What
the hot path consumes is only the processed schema, reproduced here from a seed:

  num_nodes, num_node_types, types, type_accs[type] (node ids contiguous by type,
  movielens.py:184-227), num_<type>s, edge_index_nps[name] -> np.ndarray[2, E] (row 0 = source,
  row 1 = target; ``user2item`` is float64 as upstream, movielens.py:294), test_pos_unid_inid_map
  (leave-one-out, the timestamp-last item, :297-307), neg_unid_inid_map (all items the user never
  touched), iid_feat_nids / uid_feat_nids (entity-aware features, :946-992), nid2e_dict.

Batch assembly mirrors datasets/movielens.py:879-997 (``cf_negative_sampling``, BPR branch) and
:1135-1182 (``__getitem__``) and consumes the global python / numpy / torch generators the same
way, so a seeded run draws the same triples as the reference would on this graph.
"""
import random as rd

import numpy as np
import torch

SHAPES = {
    # name: node counts by type (in id order), interactions, taggings, misc
    'tiny': dict(dataset='Movielens', name='latest-small',
                 counts=[('uid', 40), ('iid', 120), ('genre', 6), ('year', 4), ('director', 10), ('actor', 20),
                         ('writer', 12), ('tid', 15)],
                 interactions=1200, min_user=8, taggings=150),
    'ml-small': dict(dataset='Movielens', name='latest-small',
                     counts=[('uid', 610), ('iid', 9724), ('genre', 19), ('year', 8), ('director', 400),
                             ('actor', 1200), ('writer', 500), ('tid', 1500)],
                     interactions=100836, min_user=20, taggings=3683),
    'ml-small-ref': dict(dataset='Movielens', name='latest-small',       # N = 2933 like the shipped checkpoints
                         counts=[('uid', 608), ('iid', 1500), ('genre', 18), ('year', 8), ('director', 150),
                                 ('actor', 350), ('writer', 200), ('tid', 99)],
                         interactions=79500, min_user=20, taggings=2000),
    'ml-25m': dict(dataset='Movielens', name='25m',
                   counts=[('uid', 162541), ('iid', 62423), ('genre', 20), ('year', 8), ('director', 10000),
                           ('actor', 30000), ('writer', 15000), ('tid', 10000), ('genome_tid', 1128)],
                   interactions=25000095, min_user=20, taggings=1093360, genome_edges=700000),
    'ml-25m-lite': dict(dataset='Movielens', name='25m',                 # same schema, 1/10 of the edges
                        counts=[('uid', 16254), ('iid', 6242), ('genre', 20), ('year', 8), ('director', 1000),
                                ('actor', 3000), ('writer', 1500), ('tid', 1000), ('genome_tid', 1128)],
                        interactions=2500000, min_user=20, taggings=109336, genome_edges=70000),
    'yelp': dict(dataset='Yelp', name=None,
                 counts=[('uid', 200000), ('iid', 50000), ('user_reviewcount', 100), ('user_friendcount', 30),
                         ('user_fan', 100), ('user_star', 9), ('item_star', 9), ('item_reviewcount', 15),
                         ('item_attribute', 40), ('item_categorie', 1300), ('item_checkincount', 30)],
                 interactions=4000000, min_user=11, item_zipf=1.2),
    'yelp-lite': dict(dataset='Yelp', name=None,
                      counts=[('uid', 4000), ('iid', 1500), ('user_reviewcount', 30), ('user_friendcount', 10),
                              ('user_fan', 20), ('user_star', 9), ('item_star', 9), ('item_reviewcount', 15),
                              ('item_attribute', 20), ('item_categorie', 100), ('item_checkincount', 12)],
                      interactions=60000, min_user=11, item_zipf=1.2),
}


class _NidToEntity(object):
    """nid -> (type, index) without materialising a dict of N tuples (movielens.py:189-227)."""

    def __init__(self, types, accs, counts):
        self._types = types
        self._starts = np.array([accs[t] for t in types], dtype=np.int64)
        self._counts = counts

    def __getitem__(self, nid):
        k = int(np.searchsorted(self._starts, nid, side='right') - 1)
        return self._types[k], int(nid - self._starts[k])


class _LazyNegMap(object):
    """neg_unid_inid_map for graphs whose dense per-user lists would not fit in memory:
    u_nid -> np.ndarray of every item nid the user has no edge to (train or held-out)."""

    def __init__(self, ds):
        self._ds = ds

    def __getitem__(self, u_nid):
        return self._ds.unseen_items(int(u_nid))

    def __len__(self):
        return self._ds.num_uids


def _zipf_weights(n, offset, expo, rng):
    w = 1.0 / np.power(np.arange(n, dtype=np.float64) + offset, expo)
    w /= w.sum()
    return w[rng.permutation(n)]          # popularity is not correlated with the node id


def _sorted_unique(key):
    key = np.sort(key)
    keep = np.empty(key.shape[0], dtype=bool)
    keep[:1] = True
    np.not_equal(key[1:], key[:-1], out=keep[1:])
    return key[keep]


def _sample(weights, size, rng):
    cdf = np.cumsum(weights)
    cdf[-1] = 1.0
    return np.searchsorted(cdf, rng.random(size), side='right').astype(np.int64)


class SyntheticHIN(object):
    def __init__(self, shape='ml-small', seed=1234, num_negative_samples=4, sampling_strategy=None,
                 entity_aware=False, cf_loss_type='BPR', dense_neg_map=None, **unused):
        spec = SHAPES[shape]
        self.shape = shape
        self.dataset, self.name = spec['dataset'], spec['name']
        self.seed = seed
        self.num_negative_samples = num_negative_samples
        self.sampling_strategy = sampling_strategy or ('unseen' if spec['interactions'] < 10 ** 6 else 'random')
        self.entity_aware = entity_aware
        self.cf_loss_type = cf_loss_type
        if cf_loss_type != 'BPR':
            raise NotImplementedError('the PEAGNN hot path trains with BPR (experiments/pea*_solver_bpr.py)')
        rng = np.random.default_rng(seed)

        # ---- node id ranges ------------------------------------------------------------------
        self.types = [t for t, _ in spec['counts']]
        self.num_nodes_dict = dict(spec['counts'])
        self.type_accs, acc = {}, 0
        for t, c in spec['counts']:
            self.type_accs[t] = acc
            setattr(self, 'num_' + t + 's', c)
            acc += c
        self.num_nodes = acc
        self.num_node_types = len(self.types)
        self.nid2e_dict = _NidToEntity(self.types, self.type_accs, self.num_nodes_dict)
        U, I = self.num_uids, self.num_iids
        u0, i0 = self.type_accs['uid'], self.type_accs['iid']
        self.unique_uids = list(range(U))
        self.unique_iids = list(range(I))

        # ---- interactions --------------------------------------------------------------------
        act = rng.lognormal(mean=0.0, sigma=1.0, size=U)
        budget = spec['interactions'] - spec['min_user'] * U
        cnt = spec['min_user'] + np.floor(act / act.sum() * max(budget, 0)).astype(np.int64)
        cnt = np.minimum(cnt, int(0.6 * I))
        pop = _zipf_weights(I, 30.0 if I > 1000 else 5.0, spec.get('item_zipf', 1.0), rng)
        users = np.repeat(np.arange(U, dtype=np.int64), cnt)
        items = _sample(pop, users.shape[0], rng)
        key = _sorted_unique(users * I + items)                    # a user rates an item once
        users, items = key // I, key % I
        order = np.argsort(users + rng.random(users.shape[0]))     # "timestamp" order inside a user
        users, items = users[order], items[order]
        ucount = np.bincount(users, minlength=U)
        assert ucount.min() >= 2, 'every user needs a train and a held-out interaction'
        uptr = np.zeros(U + 1, dtype=np.int64)
        np.cumsum(ucount, out=uptr[1:])
        last = np.zeros(users.shape[0], dtype=bool)
        last[uptr[1:] - 1] = True                                  # leave-one-out: the latest item
        self._test_items = items[last] + i0                        # [U] held-out item nid per user
        tr_u, tr_i = users[~last] + u0, items[~last] + i0
        self.edge_index_nps = {}
        self.edge_index_nps['user2item'] = np.vstack([tr_u, tr_i]).astype(np.float64)   # float64 upstream
        self.rating_np = np.ones(tr_u.shape[0])
        self._train_ptr = uptr - np.arange(U + 1)                  # per-user slices of the train edges
        self._train_items_sorted = None
        self.test_pos_unid_inid_map = {int(u0 + u): [int(self._test_items[u])] for u in range(U)}
        dense = dense_neg_map if dense_neg_map is not None else (U * I <= 8 * 10 ** 6)
        if dense:
            self.neg_unid_inid_map = {}
            all_items = np.arange(i0, i0 + I, dtype=np.int64)
            for u in range(U):
                seen = items[uptr[u]:uptr[u + 1]] + i0
                self.neg_unid_inid_map[int(u0 + u)] = np.setdiff1d(all_items, seen, assume_unique=True).tolist()
        else:
            self.neg_unid_inid_map = _LazyNegMap(self)
        self._all_user_items = (uptr, items + i0)

        # ---- feature relations ---------------------------------------------------------------
        if self.dataset == 'Movielens':
            self._movielens_features(spec, rng)
        else:
            self._yelp_features(spec, rng)
        self.edge_type_dict = {k: i for i, k in enumerate(self.edge_index_nps.keys())}
        self.num_edge_types = len(self.edge_index_nps)
        self.train_data = None
        self.train_data_length = 0

    # ------------------------------------------------------------------------------------------
    def _rel(self, src_type, src_idx, dst_type, dst_idx):
        return np.vstack([src_idx + self.type_accs[src_type], dst_idx + self.type_accs[dst_type]]).astype(np.int64)

    def _per_item_people(self, kind, max_per_item, rng):
        I, n = self.num_iids, self.num_nodes_dict[kind]
        k = rng.integers(0, max_per_item + 1, size=I)
        it = np.repeat(np.arange(I, dtype=np.int64), k)
        who = _sample(_zipf_weights(n, 3.0, 1.0, rng), it.shape[0], rng)
        key = _sorted_unique(it * n + who)                              # listed once per item, item-major order
        return key % n, key // n

    def _movielens_features(self, spec, rng):
        I, U = self.num_iids, self.num_uids
        ei = self.edge_index_nps
        all_i = np.arange(I, dtype=np.int64)
        ei['year2item'] = self._rel('year', rng.integers(0, self.num_years, size=I), 'iid', all_i)
        ng = 1 + rng.binomial(3, 0.43, size=I)                     # ~2.3 genres per item
        gi = np.repeat(all_i, ng)
        gg = _sample(_zipf_weights(self.num_genres, 2.0, 0.8, rng), gi.shape[0], rng)
        key = _sorted_unique(gg * I + gi)                               # genre-major order (movielens.py:238-244)
        ei['genre2item'] = self._rel('genre', key // I, 'iid', key % I)
        feats = {'year': ei['year2item'], 'genre': ei['genre2item']}
        for kind, mx in (('director', 1), ('actor', 4), ('writer', 2)):
            who, it = self._per_item_people(kind, mx, rng)
            ei[kind + '2item'] = self._rel(kind, who, 'iid', it)
            feats[kind] = ei[kind + '2item']
        T = spec['taggings']
        taggers = _sample(_zipf_weights(U, 2.0, 1.1, rng), T, rng)  # few users tag a lot
        t_items = _sample(_zipf_weights(I, 30.0 if I > 1000 else 5.0, 1.0, rng), T, rng)
        t_tags = _sample(_zipf_weights(self.num_tids, 5.0, 1.0, rng), T, rng)
        ei['tag2user'] = self._rel('tid', t_tags, 'uid', taggers)   # duplicates kept on purpose
        ei['tag2item'] = self._rel('tid', t_tags, 'iid', t_items)
        feats['tid_item'] = ei['tag2item']
        if 'genome_tid' in self.type_accs:
            G = spec['genome_edges']
            n_items = max(1, min(I, G // 50))
            g_items = rng.choice(I, size=n_items, replace=False)
            gi = np.repeat(g_items.astype(np.int64), G // n_items)
            gt = rng.integers(0, self.num_genome_tids, size=gi.shape[0])
            key = _sorted_unique(gi * self.num_genome_tids + gt)
            ei['genome_tag2item'] = self._rel('genome_tid', key % self.num_genome_tids, 'iid', key // self.num_genome_tids)
            feats['genome'] = ei['genome_tag2item']
        # entity-aware feature lists, in the reference's order (movielens.py:946-992):
        # item: year, genres, actors, directors, writers, tags(, genome tags); user: tags
        order = ['year', 'genre', 'actor', 'director', 'writer', 'tid_item'] + (['genome'] if 'genome' in feats else [])
        self.iid_feat_nids = self._group_features([feats[k] for k in order], self.type_accs['iid'], I)
        self.uid_feat_nids = self._group_features([ei['tag2user']], self.type_accs['uid'], U)

    def _yelp_features(self, spec, rng):
        I, U = self.num_iids, self.num_uids
        ei = self.edge_index_nps
        all_i, all_u = np.arange(I, dtype=np.int64), np.arange(U, dtype=np.int64)
        one = lambda t, n: rng.integers(0, self.num_nodes_dict[t], size=n)
        ei['stars2item'] = self._rel('item_star', one('item_star', I), 'iid', all_i)
        ei['reviewcount2item'] = self._rel('item_reviewcount', one('item_reviewcount', I), 'iid', all_i)
        ei['checkincount2item'] = self._rel('item_checkincount', one('item_checkincount', I), 'iid', all_i)
        for name, t, mx in (('attributes2item', 'item_attribute', 6), ('categories2item', 'item_categorie', 4)):
            k = 1 + rng.integers(0, mx, size=I)
            it = np.repeat(all_i, k)
            n = self.num_nodes_dict[t]
            who = _sample(_zipf_weights(n, 3.0, 1.0, rng), it.shape[0], rng)
            key = _sorted_unique(it * n + who)
            ei[name] = self._rel(t, key % n, 'iid', key // n)
        ei['reviewcount2user'] = self._rel('user_reviewcount', one('user_reviewcount', U), 'uid', all_u)
        ei['friendcount2user'] = self._rel('user_friendcount', one('user_friendcount', U), 'uid', all_u)
        ei['fans2user'] = self._rel('user_fan', one('user_fan', U), 'uid', all_u)
        ei['stars2user'] = self._rel('user_star', one('user_star', U), 'uid', all_u)
        self.iid_feat_nids = self._group_features(
            [ei[k] for k in ('stars2item', 'reviewcount2item', 'attributes2item', 'categories2item', 'checkincount2item')],
            self.type_accs['iid'], I)
        self.uid_feat_nids = self._group_features(
            [ei[k] for k in ('reviewcount2user', 'friendcount2user', 'fans2user', 'stars2user')],
            self.type_accs['uid'], U)

    @staticmethod
    def _group_features(relations, lo, n):
        """list[n] of feature-nid lists: for each target, its sources relation by relation."""
        src = np.concatenate([r[0] for r in relations])
        dst = np.concatenate([r[1] for r in relations]) - lo
        order = np.argsort(dst, kind='stable')
        src, dst = src[order], dst[order]
        ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(dst, minlength=n), out=ptr[1:])
        if n > 200000:
            return _LazyFeatureLists(src, ptr)
        return [src[ptr[k]:ptr[k + 1]].tolist() for k in range(n)]

    # ------------------------------------------------------------------------------------------
    def unseen_items(self, u_nid):
        uptr, items = self._all_user_items
        u = u_nid - self.type_accs['uid']
        all_items = np.arange(self.type_accs['iid'], self.type_accs['iid'] + self.num_iids, dtype=np.int64)
        return np.setdiff1d(all_items, items[uptr[u]:uptr[u + 1]], assume_unique=True)

    def user_seen_csr(self):
        """(ptr[U+1], sorted item nids) of every item a user touched (train + held-out)."""
        uptr, items = self._all_user_items
        if self._train_items_sorted is None:
            u = np.repeat(np.arange(self.num_uids, dtype=np.int64), np.diff(uptr))
            span = np.int64(self.type_accs['iid'] + self.num_iids + 1)
            self._train_items_sorted = np.sort(u * span + items) % span
        return uptr, self._train_items_sorted

    def unseen_counts(self):
        """len(neg_unid_inid_map[u]) for every user, without building the lists."""
        uptr, _ = self._all_user_items
        return self.num_iids - np.diff(uptr)

    def kth_unseen(self, idx):
        """neg_unid_inid_map[u][idx[u, :]] for every user at once (the maps list unseen items in
        ascending id order): the j-th missing id of a sorted list s is j + #{t : s_t - t <= j}."""
        uptr, seen = self.user_seen_csr()
        i0 = self.type_accs['iid']
        U = self.num_uids
        cnt = np.diff(uptr)
        owner = np.repeat(np.arange(U, dtype=np.int64), cnt)
        rank = np.arange(seen.shape[0], dtype=np.int64) - np.repeat(uptr[:-1], cnt)
        big = np.int64(self.num_iids + 1)
        key = (seen - i0 - rank) + owner * big                  # nondecreasing inside a user, users in order
        q = idx.astype(np.int64) + np.arange(U, dtype=np.int64)[:, None] * big
        below = np.searchsorted(key, q.ravel(), side='right').reshape(idx.shape) - uptr[:-1][:, None]
        return idx + below + i0

    # ---- reference datasets/movielens.py:879-997 (BPR branch) ---------------------------------
    def cf_negative_sampling(self):
        pos = self.edge_index_nps['user2item'].T
        n = pos.shape[0]
        k = self.num_negative_samples
        train = np.repeat(pos, repeats=k, axis=0)
        if self.sampling_strategy == 'random':
            neg = np.random.randint(low=self.type_accs['iid'], high=self.type_accs['iid'] + self.num_iids,
                                    size=(n * k, 1))
        elif self.sampling_strategy == 'unseen':
            neg = np.empty((n * k, 1), dtype=np.int64)
            rnd = rd.random
            cache_u, pool, m = None, None, 0
            for row, u_nid in enumerate(pos[:, 0]):
                if u_nid != cache_u:                       # user2item is grouped by user
                    cache_u = u_nid
                    pool = self.test_pos_unid_inid_map[u_nid] + list(self.neg_unid_inid_map[u_nid])
                    m = len(pool)
                for j in range(k):                         # rd.choices(pool, k=k): floor(random() * m)
                    neg[row * k + j, 0] = pool[int(rnd() * m)]
        else:
            raise NotImplementedError
        train = np.hstack([train, neg])
        train_t = torch.from_numpy(train).long()
        shuffle_idx = torch.randperm(train_t.shape[0])
        self.train_data = train_t[shuffle_idx]
        self.train_data_length = train_t.shape[0]

    def __len__(self):
        return self.train_data_length

    def _entity_columns(self, feat_nids):
        if len(feat_nids) == 0:
            return 0, 0, 0
        pos_e = rd.choice(feat_nids)
        etype = self.nid2e_dict[pos_e][0]
        lo = self.type_accs.get(etype)
        hi = lo + getattr(self, 'num_' + etype + 's')
        return pos_e, rd.choice(range(lo, hi)), 1

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return getattr(self, idx, None)
        idx = idx.to_list() if torch.is_tensor(idx) else idx
        row = self.train_data[idx]
        if self.entity_aware:
            inid = row[1].item()
            ie = self._entity_columns(self.iid_feat_nids[int(inid - self.type_accs['iid'])])
            uid = row[0].item()
            ue = self._entity_columns(self.uid_feat_nids[int(uid - self.type_accs['uid'])])
            row = torch.cat([row, torch.tensor(list(ie) + list(ue), dtype=torch.long)], dim=-1)
        return row

    def get_batch(self, indices):
        """The batch DataLoader's default collate would build from ``[self[i] for i in indices]``."""
        if not self.entity_aware:
            return self.train_data[torch.as_tensor(indices, dtype=torch.long)]
        return torch.stack([self[i] for i in indices], dim=0)

    def __setitem__(self, key, value):
        if isinstance(key, str):
            setattr(self, key, value)
        else:
            raise NotImplementedError('Assignment can\'t be done outside of constructor')

    def __repr__(self):
        return 'SyntheticHIN-{}(N={}, E_u2i={})'.format(self.shape, self.num_nodes, self.edge_index_nps['user2item'].shape[1])


class _LazyFeatureLists(object):
    def __init__(self, src, ptr):
        self._src, self._ptr = src, ptr

    def __getitem__(self, k):
        return self._src[self._ptr[k]:self._ptr[k + 1]].tolist()

    def __len__(self):
        return self._ptr.shape[0] - 1


def make_synthetic_dataset(dataset_args):
    shape = dataset_args.get('synthetic')
    if shape is None:
        if dataset_args['dataset'] == 'Movielens':
            shape = 'ml-small' if dataset_args.get('name') == 'latest-small' else 'ml-25m'
        elif dataset_args['dataset'] == 'Yelp':
            shape = 'yelp'
        else:
            raise NotImplementedError
    return SyntheticHIN(shape=shape, seed=dataset_args.get('graph_seed', 1234),
                        num_negative_samples=dataset_args.get('num_negative_samples', 4),
                        sampling_strategy=dataset_args.get('sampling_strategy'),
                        entity_aware=dataset_args.get('entity_aware', False),
                        cf_loss_type=dataset_args.get('cf_loss_type', 'BPR'))
