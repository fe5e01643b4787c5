"""Diagnostic: PEAGAT gradients on the ML-25M-shaped 1/10 graph against the reference's fp64 run (tests/golden/reference_runs.pt).
Usage: python tools/gat_grad_diag.py [ROOT]   (ROOT = a checkout whose package / library to use; default: this one)"""
import os
import sys
import torch

root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
sys.path.insert(0, os.path.join(root, 'tests'))
from helpers import oracle_model_for, product_model_for, seed_all, state_sha, rel_err      # noqa: E402
from graph_recsys_benchmark_b200.datasets import SyntheticHIN                              # noqa: E402

fx = torch.load(os.path.join(root, 'tests', 'golden', 'reference_runs.pt'), weights_only=False)['ml-25m-lite/gat/plain']
ds = SyntheticHIN('ml-25m-lite', seed=fx['graph_seed'], entity_aware=False)
seed_all(2020)
oracle = oracle_model_for(ds, 'gat')
assert state_sha(oracle.state_dict()) == fx['f32']['state_sha']
batch = fx['f32']['batches'][0].long().cuda()
for dd in (False, True):
    model = product_model_for(ds, 'gat')
    model.load_state_dict(oracle.state_dict())
    model.demand_driven_loss = dd
    model.train()
    loss = model.loss(batch)
    loss.backward()
    named = dict(model.named_parameters())
    errs = sorted(((rel_err(named[n].grad, g), n) for n, g in fx['f64']['grads'].items() if float(g.abs().max()) > 1e-10),
                  reverse=True)
    print('%s demand_driven=%s loss rel %.2e' % (root, dd, abs(loss.item() - fx['f64']['losses'][0]) / fx['f64']['losses'][0]))
    for e, n in errs[:8]:
        print('   %.2e %s' % (e, n))
