#!/usr/bin/env python
"""Generates tests/golden/reference_runs.pt by EXECUTING THE REFERENCE'S OWN CODE (oracle/ref_loader.py:
models/base.py, models/pea*.py, utils/rec_utils.py, utils/general_utils.py, solvers.py and
datasets/movielens.py run unmodified out of /root/reference; only the three PyG-1.5.0 conv classes are
the oracle's restatement, see the loader's header).

    python tests/golden/make_reference_fixtures.py            # needs /root/reference; a few minutes

One entry per (graph shape, conv family, entity_aware): the run follows reference solvers.py:123-222 -
seed 2019+run for python / numpy / torch, model construction (the init RNG stream), Adam(lr 1e-3, wd 1e-3),
``cf_negative_sampling`` + ``__getitem__`` batches, three optimizer steps, ``model.eval()`` and
``BaseSolver.metrics`` - once in fp32 (what the reference computes; the oracle must agree BIT FOR BIT on CPU)
and once with the model cast to fp64 (ground truth for the CUDA path's 1e-5 bound).
Big tensors are stored at a fixed sample of rows plus an fp64 checksum so the file stays small.
"""
import hashlib
import os
import random as rd
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

OUT = os.path.join(HERE, 'reference_runs.pt')
RUN = 1
# (shape, graph seed, kind, entity_aware, batch size, evaluate?)
CONFIGS = [
    ('tiny', 7, 'gcn', False, 256, True), ('tiny', 7, 'gat', False, 256, True), ('tiny', 7, 'sage', True, 256, True),
    ('ml-small', 1234, 'gcn', False, 1024, True),       # BASELINE.json configs[0]
    ('ml-small', 1234, 'gat', False, 1024, True),       # configs[1]
    ('ml-small', 1234, 'sage', True, 1024, True),       # configs[2]
    ('ml-25m-lite', 1234, 'gcn', False, 4096, True),    # configs[3] schema at 1/10 of the edges
    ('ml-25m-lite', 1234, 'gat', False, 4096, False),   # the same graph through the edge-softmax path (train steps only)
]
N_STEPS = 3
SAMPLE_ROWS = 512


def sha(t):
    t = t.detach().cpu().contiguous()
    return hashlib.sha256(t.numpy().tobytes()).hexdigest()


def state_sha(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def seed_all(seed):                                          # reference solvers.py:123-127
    rd.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def sample_rows(n):
    if n <= SAMPLE_ROWS:
        return torch.arange(n)
    return torch.randperm(n, generator=torch.Generator().manual_seed(0))[:SAMPLE_ROWS].sort().values


def key_of(shape, kind, ea):
    return '%s/%s/%s' % (shape, kind, 'ea' if ea else 'plain')


def build_inputs(shape, graph_seed, ea):
    from graph_recsys_benchmark_b200.datasets import SyntheticHIN
    return SyntheticHIN(shape, seed=graph_seed, entity_aware=ea)


def model_args_for(ds, kind, ea):
    from helpers import model_kwargs
    return model_kwargs(ds, kind, entity_aware=ea)


def reference_run(shape, graph_seed, kind, ea, B, evaluate=True, dtype=torch.float32, ds=None, log=None):
    """The reference's code path end to end; returns a dict of plain tensors / numbers."""
    from oracle import ref_loader
    ref = ref_loader.load()
    ds = ds if ds is not None else build_inputs(shape, graph_seed, ea)
    dargs = {'dataset': ds.dataset, 'name': ds.name}
    targs = {'device': 'cpu', 'num_neg_candidates': 99}
    margs = model_args_for(ds, kind, ea)
    seed_all(2019 + RUN)
    model = ref.model_class(kind, dargs, targs)(**margs)
    out = {'state_sha': state_sha(model.state_dict()), 'num_nodes': ds.num_nodes, 'B': B}
    model = model.to(dtype)
    opt = ref.utils.get_opt_class('adam')(params=model.parameters(), lr=1e-3, weight_decay=1e-3)
    ref.cf_negative_sampling(ds)                              # datasets/movielens.py:879-997, the reference's code
    out['train_rows'] = int(ds.train_data.shape[0])
    out['train_head'] = ds.train_data[:1000].clone()
    out['train_sha'] = sha(ds.train_data)
    batches = [torch.stack([ref.getitem(ds, i) for i in range(s * B, (s + 1) * B)]) for s in range(N_STEPS)]
    out['batches'] = torch.stack(batches)
    rows = sample_rows(ds.num_nodes)
    out['rows'] = rows
    model.train()
    losses = []
    for s in range(N_STEPS):                                  # solvers.py:213-218
        opt.zero_grad()
        loss = model.loss(batches[s])
        loss.backward()
        if s == 0:
            out['repr_rows'] = model.cached_repr.detach()[rows].clone()
            out['repr_abs_sum'] = float(model.cached_repr.detach().double().abs().sum())
            grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            out['x_grad_rows'] = grads['x'][rows].clone()
            out['x_grad_abs_sum'] = float(grads['x'].double().abs().sum())
            out['grads'] = {k: g for k, g in grads.items() if k != 'x'}
        opt.step()
        losses.append(loss.detach().cpu().item())
    out['losses'] = losses
    if log:
        log('  train losses %s' % losses)
    if evaluate:
        model.eval()                                          # models/base.py:88-96: one no-grad propagation
        out['eval_repr_rows'] = model.cached_repr[rows].clone()
        solver = ref.solver(type(model), dargs, margs, targs)
        np.random.seed(4000 + RUN)
        t0 = time.time()
        HR, NDCG, AUC, eloss = solver.metrics(RUN, 0, model, ds)      # solvers.py:33-104, the reference's loop
        out['HR'], out['NDCG'], out['AUC'], out['eval_loss'] = (np.asarray(a, dtype=np.float64) for a in (HR, NDCG, AUC, eloss))
        # the same draws again, to keep the candidate lists and per-user ranks (solvers.py:21-31, :85-89)
        np.random.seed(4000 + RUN)
        cands, ranks, gaps = [], [], []
        with torch.no_grad():
            for u in list(ds.test_pos_unid_inid_map.keys()):
                pos, neg = solver.generate_candidates(ds, u)
                c = np.asarray(list(pos) + list(neg), dtype=np.int64)
                cands.append(c)
                # scores exactly as solvers.py:74-88 takes them: positives and negatives in separate predict() calls
                pos_pred = model.predict(torch.full((len(pos),), int(u), dtype=torch.long), torch.tensor(pos, dtype=torch.long)).reshape(-1)
                neg_pred = model.predict(torch.full((len(neg),), int(u), dtype=torch.long),
                                         torch.from_numpy(np.asarray(neg, dtype=np.int64))).reshape(-1)
                pred = torch.cat([pos_pred, neg_pred])
                _, idx = torch.sort(pred, descending=True)
                hit_vec = (idx < len(pos)).numpy()
                ranks.append(int(np.argmax(hit_vec)))
                gaps.append(float((neg_pred - pos_pred[0]).abs().min() / (pred.abs().max() + 1e-30)))
        cand = np.stack(cands)
        out['ranks'] = torch.tensor(ranks, dtype=torch.int16)
        out['rank_gap'] = torch.tensor(gaps, dtype=torch.float32)   # relative score gap to the nearest negative
        out['cand_sha'] = hashlib.sha256(cand.tobytes()).hexdigest()
        out['cand_head'] = torch.from_numpy(cand[:256].copy())
        if log:
            log('  eval %.1fs HR@10 %.4f NDCG@10 %.4f AUC %.4f loss %.4f' % (time.time() - t0, HR[5], NDCG[5], AUC[0], eloss[0]))
    return out


def compact(e):
    """Keeps the file small: integer streams once (int32), the fp32 run as scalars + metrics only (its tensors
    are compared live against the oracle whenever /root/reference is present, tests/test_reference_pinning.py),
    the fp64 run's tensors rounded to fp32 (6e-8 relative - far inside the 1e-5 bound they serve)."""
    f32, f64 = e['f32'], e['f64']
    for k in ('train_head', 'batches', 'rows', 'cand_head'):          # identical in both runs (integer streams)
        if k in f32:
            assert torch.equal(f32[k], f64[k])
            f64.pop(k, None)
            f32[k] = f32[k].to(torch.int32)
    # how far the reference's OWN fp32 run is from its fp64 run, per parameter gradient (max-norm relative): the floor
    # below which no fp32 implementation of this configuration can be judged
    f32['grad_err'] = {k: float((f32['grads'][k].double() - g.double()).abs().max() / g.double().abs().max())
                       if float(g.abs().max()) > 1e-10 else 0.0 for k, g in f64['grads'].items()}
    for k in ('repr_rows', 'x_grad_rows', 'grads', 'eval_repr_rows'):
        f32.pop(k, None)
    for k in ('repr_rows', 'x_grad_rows', 'eval_repr_rows'):
        if k in f64:
            f64[k] = f64[k].float()
    f64['grads'] = {k: g.float() for k, g in f64['grads'].items()}


def main():
    """``--only KEY [KEY ...]`` regenerates just those entries and merges them into the existing file."""
    from oracle import ref_loader
    assert ref_loader.available(), 'this script runs the reference out of /root/reference'
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[sys.argv.index('--only') + 1:] if '--only' in sys.argv else None
    if only:
        fixtures = torch.load(OUT, weights_only=False)
    else:
        fixtures = {'_meta': {'generator': 'tests/golden/make_reference_fixtures.py', 'run': RUN, 'n_steps': N_STEPS,
                              'torch': torch.__version__, 'numpy': np.__version__,
                              'note': 'reference code executed from /root/reference; convs = oracle/pyg150.py'}}
    for shape, gseed, kind, ea, B, evaluate in CONFIGS:
        k = key_of(shape, kind, ea)
        if only and k not in only:
            continue
        print(k, flush=True)
        ds = build_inputs(shape, gseed, ea)
        e = {'graph_seed': gseed}
        t0 = time.time()
        e['f32'] = reference_run(shape, gseed, kind, ea, B, evaluate, torch.float32, ds=ds, log=print)
        e['f64'] = reference_run(shape, gseed, kind, ea, B, evaluate, torch.float64, ds=ds, log=print)
        compact(e)
        print('  %.1fs' % (time.time() - t0), flush=True)
        fixtures[k] = e
    torch.save(fixtures, OUT)
    print('wrote %s (%.2f MB)' % (OUT, os.path.getsize(OUT) / 1e6))


if __name__ == '__main__':
    main()
