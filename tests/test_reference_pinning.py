"""Pins the oracle to the REFERENCE'S OWN CODE.

Two layers:
  * live (only where /root/reference exists, i.e. in the authoring container): the reference's
    models/base.py, models/pea*.py, utils/rec_utils.py, utils/general_utils.py, solvers.py and
    datasets/movielens.py are executed unmodified through ``oracle/ref_loader.py`` and the oracle must
    reproduce them BIT FOR BIT - parameters drawn from the same seed, triples, entity columns,
    losses over three Adam steps, representations, every gradient, candidate lists, ranks, HR / NDCG /
    AUC / eval loss;
  * frozen (everywhere, incl. the GPU box): ``tests/golden/reference_runs.pt`` was written by
    ``tests/golden/make_reference_fixtures.py`` from those same reference runs; the oracle must
    reproduce it (integers exactly, floating point to 1e-6 - BLAS blocking may differ between hosts).
What stays pinned only by dense closed forms (tests/test_oracle.py): the three PyG-1.5.0 conv
classes themselves - torch-geometric is not vendored in the reference and not installed here.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import oracle_run, product_model_for, rel_err, seed_all, state_sha     # noqa: E402
from oracle import ref_loader                                                       # noqa: E402
from graph_recsys_benchmark_b200.datasets import SyntheticHIN                       # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_runs.pt')
needs_reference = pytest.mark.skipif(not ref_loader.available(), reason='the reference checkout is not on this box')


def _fixtures():
    return torch.load(GOLDEN, weights_only=False)


def _reference_run(*a, **kw):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from make_reference_fixtures import reference_run
    return reference_run(*a, **kw)


LIVE = [('tiny', 7, 'gcn', False, 256), ('tiny', 7, 'gat', False, 256), ('tiny', 7, 'sage', True, 256),
        ('tiny', 7, 'gcn', True, 256), ('ml-small', 1234, 'gcn', False, 1024), ('ml-small', 1234, 'gat', False, 1024)]


@needs_reference
@pytest.mark.parametrize('shape,gseed,kind,ea,B', LIVE)
def test_oracle_equals_the_reference_code_bit_for_bit(shape, gseed, kind, ea, B):
    ds = SyntheticHIN(shape, seed=gseed, entity_aware=ea)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)        # multi-threaded CPU scatter-adds in autograd are not run-to-run reproducible
    try:
        ref = _reference_run(shape, gseed, kind, ea, B, evaluate=True, ds=ds)
        ora = oracle_run(ds, kind, ea, B, evaluate=True)
    finally:
        torch.set_num_threads(threads)
    assert ora['state_sha'] == ref['state_sha']                    # same init draws from the same seed
    for k in ('train_head', 'batches', 'rows', 'cand_head', 'ranks'):
        assert torch.equal(ora[k], ref[k]), k
    assert ora['train_sha'] == ref['train_sha'] and ora['cand_sha'] == ref['cand_sha']
    assert ora['losses'] == ref['losses']                          # three Adam steps, bitwise
    for k in ('repr_rows', 'x_grad_rows', 'eval_repr_rows'):
        assert torch.equal(ora[k], ref[k]), k
    for k, g in ref['grads'].items():
        assert torch.equal(ora['grads'][k], g), k
    for k in ('HR', 'NDCG', 'AUC', 'eval_loss'):
        assert np.array_equal(np.asarray(ora[k]), np.asarray(ref[k])), k


@needs_reference
def test_reference_rec_utils_and_metapath_tables():
    from oracle import rec_utils as orec, graph as ograph
    from graph_recsys_benchmark_b200.utils import update_pea_graph_input
    ref = ref_loader.load()
    rng = np.random.RandomState(3)
    for _ in range(200):
        hv = np.zeros(100, dtype=bool)
        hv[rng.randint(0, 100)] = True
        assert ref.rec_utils.hit(hv) == orec.hit(hv)
        assert np.array_equal(ref.rec_utils.ndcg(hv), orec.ndcg(hv))
    p, n = rng.rand(3).astype(np.float32), rng.rand(99).astype(np.float32)
    assert ref.rec_utils.auc(p, n) == orec.auc(p, n)
    for shape in ('ml-small', 'ml-25m-lite', 'yelp-lite'):        # general_utils.py:300-313, :335-356, :377-392
        ds = SyntheticHIN(shape, seed=1234)
        dargs, targs = {'dataset': ds.dataset, 'name': ds.name}, {'device': 'cpu'}
        theirs = ref.general_utils.update_pea_graph_input(dargs, targs, ds)
        ours = update_pea_graph_input(dargs, targs, ds)
        orac = ograph.metapath_edge_index_list(ds, ds.dataset, ds.name)
        assert len(theirs) == len(ours) == len(orac)
        for a, b, c in zip(theirs, ours, orac):
            assert len(a) == len(b) == len(c)
            for x, y, z in zip(a, b, c):
                assert torch.equal(x, y) and torch.equal(x, z)


@needs_reference
def test_product_sampling_and_init_follow_the_reference_streams():
    """The package's own host logic (datasets/synthetic.py, models/base.py::_init) against the reference's code:
    same triples, same entity columns, same initial parameters from the same seed."""
    ref = ref_loader.load()
    for shape, gseed, ea in (('tiny', 7, True), ('ml-small', 1234, False)):
        ds_a, ds_b = SyntheticHIN(shape, seed=gseed, entity_aware=ea), SyntheticHIN(shape, seed=gseed, entity_aware=ea)
        seed_all(11)
        ref.cf_negative_sampling(ds_a)
        rows_a = torch.stack([ref.getitem(ds_a, i) for i in range(300)])
        seed_all(11)
        ds_b.cf_negative_sampling()
        rows_b = torch.stack([ds_b[i] for i in range(300)])
        assert torch.equal(ds_a.train_data, ds_b.train_data)
        assert torch.equal(rows_a, rows_b)
    from helpers import model_kwargs
    ds = SyntheticHIN('tiny', seed=7)
    dargs, targs = {'dataset': ds.dataset, 'name': ds.name}, {'device': 'cpu'}
    for kind in ('gcn', 'gat', 'sage'):
        seed_all(2020)
        theirs = ref.model_class(kind, dargs, targs)(**model_kwargs(ds, kind))
        seed_all(2020)
        ours = product_model_for(ds, kind, device='cpu')
        assert list(theirs.state_dict().keys()) == list(ours.state_dict().keys())
        assert state_sha(theirs.state_dict()) == state_sha(ours.state_dict())


FROZEN = [('tiny', 'gcn', False, True), ('tiny', 'gat', False, True), ('tiny', 'sage', True, True),
          ('ml-small', 'gcn', False, True), ('ml-small', 'gat', False, True), ('ml-small', 'sage', True, True),
          ('ml-25m-lite', 'gcn', False, False), ('ml-25m-lite', 'gat', False, False)]


@pytest.mark.parametrize('shape,kind,ea,evaluate', FROZEN)
def test_oracle_reproduces_the_committed_reference_runs(shape, kind, ea, evaluate):
    fx = _fixtures()['%s/%s/%s' % (shape, kind, 'ea' if ea else 'plain')]
    f32, f64 = fx['f32'], fx['f64']
    ds = SyntheticHIN(shape, seed=fx['graph_seed'], entity_aware=ea)
    ora = oracle_run(ds, kind, ea, f32['B'], evaluate=evaluate, dtype=torch.float64)
    assert ora['state_sha'] == f32['state_sha'] == f64['state_sha']
    assert ora['train_sha'] == f32['train_sha'] and ora['train_rows'] == f32['train_rows']
    for k in ('train_head', 'batches', 'rows'):
        assert torch.equal(ora[k].to(torch.int32), f32[k]), k
    assert np.allclose(ora['losses'], f64['losses'], rtol=1e-9, atol=0)
    assert np.allclose(ora['losses'], f32['losses'], rtol=2e-5, atol=0)        # the fp32 run of the reference
    assert rel_err(ora['repr_rows'], f64['repr_rows']) < 1e-6
    assert rel_err(ora['x_grad_rows'], f64['x_grad_rows']) < 1e-6
    assert abs(ora['repr_abs_sum'] - f64['repr_abs_sum']) <= 1e-9 * f64['repr_abs_sum']
    for k, g in f64['grads'].items():
        if float(g.abs().max()) > 1e-12:
            assert rel_err(ora['grads'][k], g) < 1e-6, k
    if evaluate:
        assert ora['cand_sha'] == f32['cand_sha']
        assert torch.equal(ora['cand_head'].to(torch.int32), f32['cand_head'])
        assert torch.equal(ora['ranks'], f64['ranks'])
        for k in ('HR', 'NDCG', 'AUC'):
            assert np.array_equal(np.asarray(ora[k]), f64[k]), k
        assert np.allclose(ora['eval_loss'], f64['eval_loss'], rtol=1e-9)
