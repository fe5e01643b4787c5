"""Oracle solver pieces (TEST INFRASTRUCTURE) - restates reference solvers.py:
generate_candidates :21-31, metrics :33-104, the optimiser step :203-222 and
the seeding at :123-127.  pandas' inner merge on u_nid (:63-67) is restated as
the cartesian product pos x neg it computes for a single user.
"""
import random as rd

import numpy as np
import torch

from .rec_utils import hit, ndcg, auc


def seed_everything(run):                                   # solvers.py:123-127
    seed = 2019 + run
    rd.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


def generate_candidates(dataset, u_nid, num_neg_candidates):   # solvers.py:21-31
    pos = dataset.test_pos_unid_inid_map[u_nid]
    neg = list(np.random.choice(dataset.neg_unid_inid_map[u_nid], size=(num_neg_candidates,)))
    return pos, neg


def metrics(model, dataset, num_neg_candidates=99, return_per_user=False):
    """model must be in eval mode with cached_repr set (model.eval())."""
    HRs, NDCGs, AUC, losses = np.zeros((0, 16)), np.zeros((0, 16)), np.zeros((0, 1)), np.zeros((0, 1))
    ranks = []
    with torch.no_grad():
        for u_nid in list(dataset.test_pos_unid_inid_map.keys()):
            pos, neg = generate_candidates(dataset, u_nid, num_neg_candidates)
            if len(pos) == 0 or len(neg) == 0:
                raise ValueError("No pos or neg samples found in evaluation!")
            pair = torch.tensor([[u_nid, p, n] for p in pos for n in neg], dtype=torch.long)
            loss = model.loss(pair).detach().cpu().item()
            u_pos = torch.full((len(pos),), u_nid, dtype=torch.long)
            u_neg = torch.full((len(neg),), u_nid, dtype=torch.long)
            pos_pred = model.predict(u_pos, torch.tensor(pos, dtype=torch.long)).reshape(-1)
            neg_pred = model.predict(u_neg, torch.tensor(neg, dtype=torch.long)).reshape(-1)
            _, indices = torch.sort(torch.cat([pos_pred, neg_pred]), descending=True)
            hit_vec = (indices < len(pos)).numpy()
            ranks.append(int(np.argmax(hit_vec)))
            HRs = np.vstack([HRs, hit(hit_vec)])
            NDCGs = np.vstack([NDCGs, ndcg(hit_vec)])
            AUC = np.vstack([AUC, auc(pos_pred.numpy(), neg_pred.numpy())])
            losses = np.vstack([losses, loss])
    out = (np.mean(HRs, axis=0), np.mean(NDCGs, axis=0), np.mean(AUC, axis=0), np.mean(losses, axis=0))
    if return_per_user:
        return out, dict(ranks=np.array(ranks), HRs=HRs, NDCGs=NDCGs, AUC=AUC, losses=losses)
    return out


def train_step(model, optimizer, batch):                    # solvers.py:213-218
    optimizer.zero_grad()
    loss = model.loss(batch)
    loss.backward()
    optimizer.step()
    return loss.detach().cpu().item()
