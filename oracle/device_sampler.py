"""TEST INFRASTRUCTURE ONLY - CPU mirror of the product's device-side BPR row sampler
(csrc/sample.cu, ``peagnn_bpr_rows``).

Unlike the rest of oracle/, this is not a restatement of reference code: the reference samples with
the host's numpy / random generators (datasets/movielens.py:920-940, 1153-1177), which a GPU cannot
replay.  The device sampler is counter-based instead (Philox4x32-10, Salmon et al., "Parallel random
numbers: as easy as 1, 2, 3", SC'11 - the published algorithm, pinned below by the Random123 known-answer
vectors), and this file recomputes its rows with numpy integer arithmetic so the kernel can be checked
bit for bit; the reference's DISTRIBUTIONS are checked separately (tests/test_host_logic.py).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds; all arguments uint32-valued (arrays or scalars)."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3)]
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> S32) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> S32) ^ c3 ^ np.uint64(k1)
        c1, c3, c0, c2 = p1 & MASK32, p0 & MASK32, n0, n2
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def bounded(lo, hi, n):
    """floor(((hi << 32) | lo) * n / 2^64) for n < 2^31 - the kernel's __umul64hi."""
    n = np.asarray(n, dtype=np.uint64)
    return (((hi * n) + ((lo * n) >> S32)) >> S32).astype(np.int64)


def bpr_rows(row_ids, u2i, num_neg, seed, epoch, strategy, user_lo, item_lo, num_items, seen_ptr=None,
             seen_items=None, cols=3, ifeat=None, ufeat=None, type_starts=None):
    """Rows ``row_ids`` of the epoch's unshuffled table.  ``ifeat`` / ``ufeat`` = (ptr, nids) CSR pairs."""
    r = np.asarray(row_ids, dtype=np.uint64)
    e = (r // np.uint64(num_neg)).astype(np.int64)
    u, pos = u2i[0][e], u2i[1][e]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    c0, c1, c2 = r & MASK32, r >> S32, np.uint64(epoch & 0xFFFFFFFF)
    d0 = philox4x32_10(c0, c1, c2, 0, k0, k1)
    if strategy == 0:
        neg = item_lo + bounded(d0[0], d0[1], num_items)
    else:
        s0 = seen_ptr[u - user_lo]
        deg = seen_ptr[u - user_lo + 1] - s0
        kth = bounded(d0[0], d0[1], num_items - deg)
        neg = np.empty_like(kth)
        for i in range(len(kth)):                       # plain loop: small cases only
            rel = seen_items[s0[i]:s0[i] + deg[i]] - item_lo - np.arange(deg[i])
            neg[i] = item_lo + kth[i] + np.searchsorted(rel, kth[i], side='right')
    out = np.stack([u, pos, neg], axis=1).astype(np.int64)
    if cols == 3:
        return out
    d1 = philox4x32_10(c0, c1, c2, 1, k0, k1)
    d2 = philox4x32_10(c0, c1, c2, 2, k0, k1)
    ts = np.asarray(type_starts, dtype=np.int64)

    def entity(table, local, a_lo, a_hi, b_lo, b_hi):
        ptr, nids = table
        f0 = ptr[local]
        cnt = ptr[local + 1] - f0
        has = cnt > 0
        pick = bounded(a_lo, a_hi, np.maximum(cnt, 1))
        pe = np.where(has, nids[np.minimum(f0 + pick, len(nids) - 1)] if len(nids) else 0, 0)
        t = np.searchsorted(ts[:-1], pe, side='right') - 1
        tlo, thi = ts[t], ts[t + 1]
        ne = tlo + bounded(b_lo, b_hi, thi - tlo)
        return np.stack([np.where(has, pe, 0), np.where(has, ne, 0), has.astype(np.int64)], axis=1)
    ie = entity(ifeat, pos - item_lo, d0[2], d0[3], d1[0], d1[1])
    ue = entity(ufeat, u - user_lo, d1[2], d1[3], d2[0], d2[1])
    return np.concatenate([out, ie, ue], axis=1).astype(np.int64)
