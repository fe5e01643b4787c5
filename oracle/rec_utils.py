"""Oracle ranking metrics (TEST INFRASTRUCTURE) - restates reference
utils/rec_utils.py:7-30.  Note the reference's "ndcg" is not the textbook
NDCG (:23: hits-in-top-K divided by log2(first-hit-position + 2)); kept as is.
"""
import numpy as np
from itertools import product

K_RANGE = range(5, 21)            # 16 columns; column 5 is @10


def hit(hit_vec):
    return [1 if np.sum(hit_vec[:k]) > 0 else 0 for k in K_RANGE]


def ndcg(hit_vec):
    out = []
    for k in K_RANGE:
        top = np.array(hit_vec[:k], dtype=np.int64).reshape(1, -1)
        out.append(np.sum(top) / np.log2(np.argmax(top) + 2))
    return out


def auc(preds_pos, preds_neg):
    return np.mean([1 if p > n else 0 for p, n in product(preds_pos, preds_neg)])
