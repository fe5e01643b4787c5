"""CPU: the oracle restatement against independent closed forms and fixed known answers."""
import numpy as np
import pytest
import torch

from oracle import dense, pyg150, rec_utils, graph as ograph
from helpers import random_edge_index


def _tiny_graph():
    # 6 nodes; multi-edge (0->3 twice), pre-existing self loop (2->2), isolated node 5
    return torch.tensor([[0, 0, 1, 2, 2, 4, 0], [3, 3, 3, 2, 4, 1, 4]], dtype=torch.long)


@pytest.mark.parametrize('seed', [0, 1])
def test_gcn_matches_dense(seed):
    torch.manual_seed(seed)
    ei = _tiny_graph() if seed == 0 else random_edge_index(40, 300, seed, self_loops=3, multi=20)
    n = 6 if seed == 0 else 40
    conv = pyg150.GCNConv(8, 4).double()
    conv.bias.data.uniform_(-1, 1)
    x = torch.randn(n, 8, dtype=torch.float64)
    ref = dense.gcn_dense(x, ei, conv.weight.data, conv.bias.data)
    assert torch.allclose(conv(x, ei), ref, rtol=1e-12, atol=1e-12)


def test_gcn_source_degree_known_answer():
    # directed bipartite star: users 0,1 -> item 2.  deg (source side, + self loop): u0=2, u1=2, item=1
    ei = torch.tensor([[0, 1], [2, 2]])
    conv = pyg150.GCNConv(1, 1).double()
    conv.weight.data.fill_(1.0)
    conv.bias.data.fill_(0.0)
    x = torch.tensor([[1.0], [2.0], [4.0]], dtype=torch.float64)
    out = conv(x, ei).flatten()
    s = 1 / np.sqrt(2.0)
    # out[item] = x0/sqrt(2*1) + x1/sqrt(2*1) + x2/1 ; out[u] = x_u / 2   
    assert torch.allclose(out, torch.tensor([0.5, 1.0, s * 1 + s * 2 + 4.0], dtype=torch.float64))


@pytest.mark.parametrize('heads', [1, 2])
def test_gat_matches_dense(heads):
    ei = random_edge_index(30, 200, 3, self_loops=4, multi=15)
    conv = pyg150.GATConv(8, 4, heads=heads).double()
    conv.bias.data.uniform_(-1, 1)
    x = torch.randn(30, 8, dtype=torch.float64)
    ref = dense.gat_dense(x, ei, conv.lin.weight.data, conv.att_i.data, conv.att_j.data, conv.bias.data, heads)
    assert torch.allclose(conv(x, ei), ref, rtol=1e-10, atol=1e-12)


def test_sage_matches_dense():
    ei = random_edge_index(30, 200, 5, self_loops=4, multi=15)
    conv = pyg150.SAGEConv(8, 4).double()
    x = torch.randn(30, 8, dtype=torch.float64)
    ref = dense.sage_dense(x, ei, conv.lin_rel.weight.data, conv.lin_rel.bias.data, conv.lin_root.weight.data)
    assert torch.allclose(conv(x, ei), ref, rtol=1e-12, atol=1e-12)


def test_sage_isolated_rows_are_root_only():
    ei = torch.tensor([[0], [1]])
    conv = pyg150.SAGEConv(4, 3).double()
    x = torch.randn(3, 4, dtype=torch.float64)
    out = conv(x, ei)
    assert torch.allclose(out[2], conv.lin_rel.bias + conv.lin_root(x[2]))


def test_rec_utils_known_answers():
    hit_vec = np.zeros(100, dtype=bool)
    hit_vec[7] = True                       # positive ranked 8th
    hr = rec_utils.hit(hit_vec)
    nd = rec_utils.ndcg(hit_vec)
    assert hr[:3] == [0, 0, 0] and hr[3:] == [1] * 13          # K = 5,6,7 miss; K >= 8 hit
    assert nd[0] == 0.0 and abs(nd[5] - 1 / np.log2(9)) < 1e-15
    assert rec_utils.auc([0.5], [0.1, 0.5, 0.9, 0.2]) == 0.5   # strict >


def test_csr_by_key_is_stable_and_drops_self_loops():
    ei = _tiny_graph()
    rowptr, col, eid = ograph.csr_by_key(ei[1].numpy(), ei[0].numpy(), 6, drop_self_loops=True)
    assert rowptr.tolist() == [0, 0, 1, 1, 4, 6, 6]
    assert col.tolist() == [4, 0, 0, 1, 2, 0]
    assert eid.tolist() == [5, 0, 1, 2, 4, 6]


def test_oracle_reproduces_frozen_vectors():
    """tests/golden/oracle_vectors.pt was written by tests/golden/make_oracle_vectors.py; the oracle
    (and the seeded host logic feeding it) must keep producing exactly those numbers."""
    import os, sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, 'golden'))
    import make_oracle_vectors
    frozen = torch.load(os.path.join(here, 'golden', 'oracle_vectors.pt'), weights_only=False)
    fresh = make_oracle_vectors.build()
    assert set(frozen) == set(fresh)
    for kind in ('gcn', 'gat', 'sage'):
        a, b = frozen[kind], fresh[kind]
        assert torch.equal(a['batch'], b['batch']) and torch.equal(a['ranks'], b['ranks'])
        assert abs(a['loss'] - b['loss']) <= 1e-6 * abs(a['loss'])
        assert torch.allclose(a['repr_rows'], b['repr_rows'], rtol=1e-5, atol=1e-7)
        assert torch.allclose(a['x_grad_rows'], b['x_grad_rows'], rtol=1e-4, atol=1e-8)
        assert a['hr10'] == b['hr10'] and abs(a['ndcg10'] - b['ndcg10']) < 1e-12
    for k in ('rowptr', 'col', 'eid'):
        assert torch.equal(frozen['csr_user2item_by_target'][k], fresh['csr_user2item_by_target'][k])
    assert torch.equal(frozen['train_triples_first_64'], fresh['train_triples_first_64'])
    assert torch.equal(frozen['candidates_user0'], fresh['candidates_user0'])
