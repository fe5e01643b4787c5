"""GPU: the CUDA path against runs of the REFERENCE'S OWN CODE (tests/golden/reference_runs.pt, written by
tests/golden/make_reference_fixtures.py out of /root/reference - see oracle/ref_loader.py for what executes).

Per configuration - tiny x 3 families, the three ML-small BASELINE configs (PEAGCN, PEAGAT, PEASage +
entity-aware) at their stated size, and PEAGCN on the ML-25M-shaped 1/10 graph: same graph seed, parameters
drawn from the same seed (hash-checked against the reference's), the reference's own batches; then
  * step 0: loss, representation rows, every gradient vs the reference run in fp64 (1e-5 / 1e-4);
  * three Adam steps: each loss vs the reference's (1e-5 relative);
  * model.eval() + metrics(): candidate lists identical, per-user ranks vs the reference's, HR@K / NDCG@K.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import grad_bound, oracle_model_for, product_model_for, rel_err, seed_all, state_sha    # noqa: E402
from graph_recsys_benchmark_b200.datasets import SyntheticHIN                            # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_runs.pt')
CASES = ['tiny/gcn/plain', 'tiny/gat/plain', 'tiny/sage/ea', 'ml-small/gcn/plain', 'ml-small/gat/plain',
         'ml-small/sage/ea', 'ml-25m-lite/gcn/plain', 'ml-25m-lite/gat/plain']
_cache = {}


def _fixtures():
    if 'fx' not in _cache:
        _cache['fx'] = torch.load(GOLDEN, weights_only=False)
    return _cache['fx']


@pytest.mark.parametrize('key', CASES)
def test_cuda_path_matches_the_reference_runs(key):
    from graph_recsys_benchmark_b200.solvers import BaseSolver
    shape, kind, mode = key.split('/')
    ea = mode == 'ea'
    fx = _fixtures()[key]
    f32, f64 = fx['f32'], fx['f64']
    ds = SyntheticHIN(shape, seed=fx['graph_seed'], entity_aware=ea)
    seed_all(2019 + 1)                                             # reference solvers.py:123-127, run = 1
    oracle = oracle_model_for(ds, kind, entity_aware=ea)           # draws the reference's initial parameters ...
    assert state_sha(oracle.state_dict()) == f32['state_sha']      # ... bit for bit
    model = product_model_for(ds, kind, entity_aware=ea)
    model.load_state_dict(oracle.state_dict())
    del oracle
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    batches = f32['batches'].long().cuda()
    rows = f32['rows'].long()
    model.train()
    losses = []
    for s in range(batches.shape[0]):
        opt.zero_grad()
        loss = model.loss(batches[s])
        loss.backward()
        if s == 0:
            assert rel_err(model.cached_repr.detach()[rows.cuda()], f64['repr_rows']) < 1e-5
            assert abs(float(model.cached_repr.detach().double().abs().sum()) - f64['repr_abs_sum']) < 1e-5 * f64['repr_abs_sum']
            assert rel_err(model.x.grad[rows.cuda()], f64['x_grad_rows']) < 1e-4
            assert abs(float(model.x.grad.double().abs().sum()) - f64['x_grad_abs_sum']) < 1e-4 * f64['x_grad_abs_sum']
            named = dict(model.named_parameters())
            for name, g in f64['grads'].items():
                if float(g.abs().max()) > 1e-10:
                    assert rel_err(named[name].grad, g) < grad_bound(key, fx, name), name
        opt.step()
        losses.append(loss.item())
    for got, want64, want32 in zip(losses, f64['losses'], f32['losses']):
        assert abs(got - want64) < 1e-5 * abs(want64), (losses, f64['losses'])
        assert abs(got - want32) < 2e-5 * abs(want32)               # the reference's own fp32 run
    if 'HR' not in f64:
        return
    model.eval()
    assert rel_err(model.cached_repr[rows.cuda()], f64['eval_repr_rows']) < 1e-4      # after three Adam steps
    solver = BaseSolver(None, {}, {}, {'device': 'cuda', 'num_neg_candidates': 99, 'batch_size': 1024})
    np.random.seed(4000 + 1)
    users, cand, n_pos = solver.generate_all_candidates(ds)
    import hashlib
    assert hashlib.sha256(np.ascontiguousarray(cand).tobytes()).hexdigest() == f32['cand_sha']    # same 99 negatives per user
    np.random.seed(4000 + 1)
    (HR, NDCG, AUC, eloss), per_user = solver.metrics(1, 0, model, ds, return_per_user=True)
    ranks = per_user[:, 34].long().cpu()
    want = f64['ranks'].long()
    agree = (ranks == want)
    # a rank can only differ where the positive's score is within fp32 noise of a negative's
    clear = f64['rank_gap'] > 1e-4
    assert bool(agree[clear].all()), 'ranks differ away from near-ties'
    frac = float(agree.float().mean())
    print('%s: per-user rank agreement with the reference run %.6f (%d of %d users differ, all near-ties)'
          % (key, frac, int((~agree).sum()), ranks.numel()))
    assert frac >= 0.995
    U = float(ranks.numel())
    slack = float((~agree).sum()) / U + 1e-12
    assert np.abs(HR - f64['HR']).max() <= slack
    assert np.abs(NDCG - f64['NDCG']).max() <= slack
    assert abs(AUC[0] - f64['AUC'][0]) < 1e-3
    assert abs(eloss[0] - f64['eval_loss'][0]) < 1e-4 * abs(f64['eval_loss'][0])
    if frac == 1.0:
        assert np.array_equal(HR, f64['HR']) and np.allclose(NDCG, f64['NDCG'], rtol=0, atol=1e-12)
