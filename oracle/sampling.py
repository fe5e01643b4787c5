"""Oracle negative sampling + batch assembly (TEST INFRASTRUCTURE).

Restates, loop for loop, reference datasets/movielens.py:920-940 (BPR branch of
cf_negative_sampling), :994-997 (randperm shuffle) and :1135-1182 (__getitem__
with the entity-aware columns); datasets/yelp.py:746-766,824-827,970-1017 is
the same algorithm.  Works on any object exposing the reference dataset
attributes (edge_index_nps, type_accs, num_iids, test_pos_unid_inid_map,
neg_unid_inid_map, iid_feat_nids, uid_feat_nids, nid2e_dict, num_<type>s).
Consumes the global python / numpy / torch generators exactly as the
reference does.
"""
import random as rd

import numpy as np
import torch


def cf_negative_sampling_bpr(ds, num_negative_samples, sampling_strategy):
    pos = ds.edge_index_nps['user2item'].T
    n_inter = pos.shape[0]
    train = np.repeat(pos, repeats=num_negative_samples, axis=0)
    if sampling_strategy == 'random':
        neg = np.random.randint(low=ds.type_accs['iid'], high=ds.type_accs['iid'] + ds.num_iids,
                                size=(n_inter * num_negative_samples, 1))
    elif sampling_strategy == 'unseen':
        chunks = []
        for u_nid in pos[:, 0]:
            pool = ds.test_pos_unid_inid_map[u_nid] + ds.neg_unid_inid_map[u_nid]
            picks = rd.choices(pool, k=num_negative_samples)
            chunks.append(np.array(picks, dtype=np.int64).reshape(-1, 1))
        neg = np.vstack(chunks)
    else:
        raise NotImplementedError
    train = np.hstack([train, neg])
    train_t = torch.from_numpy(train).long()
    shuffle_idx = torch.randperm(train_t.shape[0])
    return train_t[shuffle_idx]


def getitem(ds, train_data, idx, entity_aware):
    row = train_data[idx]
    if not entity_aware:
        return row

    def sample(feat_nids):
        if len(feat_nids) == 0:
            return 0, 0, 0
        pos_e = rd.choice(feat_nids)
        etype = ds.nid2e_dict[pos_e][0]
        lo = ds.type_accs.get(etype)
        hi = lo + getattr(ds, 'num_' + etype + 's')
        neg_e = rd.choice(range(lo, hi))
        return pos_e, neg_e, 1

    inid = row[1].item()
    i_pos, i_neg, i_mask = sample(ds.iid_feat_nids[int(inid - ds.type_accs['iid'])])
    uid = row[0].item()
    u_pos, u_neg, u_mask = sample(ds.uid_feat_nids[int(uid - ds.type_accs['uid'])])
    extra = torch.tensor([i_pos, i_neg, i_mask, u_pos, u_neg, u_mask], dtype=torch.long)
    return torch.cat([row, extra], dim=-1)
