"""GPU parity of the drop-in models / solver against the oracle on the same seeded inputs
(BASELINE.json configs 0-2 at test size; the 25M / Yelp shapes are covered by properties in
test_gpu_scale.py)."""
import numpy as np
import pytest
import torch

from helpers import oracle_model_for, product_model_for, rel_err
from oracle import sampling as osampling, solver as osolver

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _dataset(shape='tiny', **kw):
    from graph_recsys_benchmark_b200.datasets import SyntheticHIN
    return SyntheticHIN(shape, seed=7, **kw)


def _batch(ds, B, entity_aware, seed=0):
    import random
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    ds.entity_aware = entity_aware
    ds.cf_negative_sampling()
    return ds.get_batch(list(range(B)))


@pytest.mark.parametrize('kind', ['gcn', 'gat', 'sage'])
@pytest.mark.parametrize('entity_aware', [False, True])
@pytest.mark.parametrize('aggr', ['att', 'mean'])
def test_model_loss_and_all_gradients(kind, entity_aware, aggr):
    ds = _dataset()
    batch = _batch(ds, 300, entity_aware)
    oracle = oracle_model_for(ds, kind, entity_aware=entity_aware, channel_aggr=aggr, dtype=torch.float64)
    model = product_model_for(ds, kind, entity_aware=entity_aware, channel_aggr=aggr)
    model.load_state_dict({k: v.float() for k, v in oracle.state_dict().items()})
    # make sure the fp64 oracle holds exactly the fp32 values the product sees
    oracle.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    oracle.train(); model.train()
    lo = oracle.loss(batch); lo.backward()
    lm = model.loss(batch.to(DEV)); lm.backward()
    assert abs(lm.item() - lo.item()) / abs(lo.item()) < 1e-5
    assert rel_err(model.cached_repr, oracle.cached_repr) < 1e-5
    og = dict(oracle.named_parameters())
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        ref = og[name].grad
        if float(ref.abs().max()) < 1e-12:
            assert float(p.grad.abs().max()) < 1e-6, name
        else:
            assert rel_err(p.grad, ref) < 1e-4, name


@pytest.mark.parametrize('kind', ['gcn', 'gat', 'sage'])
def test_forward_predict_eval_and_ablation(kind):
    ds = _dataset()
    oracle = oracle_model_for(ds, kind, dtype=torch.float64)
    model = product_model_for(ds, kind)
    model.load_state_dict({k: v.float() for k, v in oracle.state_dict().items()})
    oracle.load_state_dict({k: v.double() for k, v in model.state_dict().items()})
    for idx in (None, 0, 8):
        oracle.eval(idx); model.eval(idx)
        assert not model.training
        assert rel_err(model.cached_repr, oracle.cached_repr) < 1e-5
    u = torch.randint(0, ds.num_uids, (77,))
    i = torch.randint(ds.type_accs['iid'], ds.type_accs['iid'] + ds.num_iids, (77,))
    po, pm = oracle.predict(u, i), model.predict(u.to(DEV), i.to(DEV))
    assert pm.shape == (77, 1) and rel_err(pm, po) < 1e-5


def test_state_dict_keys_match_oracle_and_reference_naming():
    ds = _dataset()
    for kind in ('gcn', 'gat', 'sage'):
        o = oracle_model_for(ds, kind)
        m = product_model_for(ds, kind)
        assert list(o.state_dict().keys()) == list(m.state_dict().keys())
        assert {k: tuple(v.shape) for k, v in o.state_dict().items()} == {k: tuple(v.shape) for k, v in m.state_dict().items()}
    keys = set(product_model_for(ds, 'gat').state_dict().keys())
    assert {'x', 'att', 'fc1.weight', 'fc2.bias', 'pea_channels.0.gnn_layers.0.lin.weight',
            'pea_channels.8.gnn_layers.1.att_j', 'pea_channels.3.gnn_layers.1.bias'} <= keys


@pytest.mark.parametrize('kind,entity_aware', [('gcn', False), ('gat', False), ('sage', True)])
def test_training_steps_and_metrics_follow_the_oracle(kind, entity_aware):
    """BASELINE configs 0-2 in miniature: same seeds -> same triples, losses within 1e-5 for the
    first steps (1e-4 after Adam has compounded rounding), HR@10 / NDCG@10 identical."""
    import random
    from graph_recsys_benchmark_b200.solvers import BaseSolver
    ds = _dataset(entity_aware=entity_aware)
    oracle = oracle_model_for(ds, kind, entity_aware=entity_aware)
    model = product_model_for(ds, kind, entity_aware=entity_aware)
    model.load_state_dict(oracle.state_dict())
    opt_o = torch.optim.Adam(oracle.parameters(), lr=1e-3, weight_decay=1e-3)
    opt_m = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    osolver.seed_everything(1)
    ds.cf_negative_sampling()
    batches = [ds.get_batch(list(range(k * 128, (k + 1) * 128))) for k in range(6)]
    oracle.train(); model.train()
    for step, b in enumerate(batches):
        lo = osolver.train_step(oracle, opt_o, b)
        opt_m.zero_grad()
        lm = model.loss(b.to(DEV)); lm.backward(); opt_m.step()
        assert abs(lm.item() - lo) / abs(lo) < (1e-5 if step < 2 else 1e-4), step
    # evaluation: identical candidate draws, identical ranks / HR / NDCG
    solver = BaseSolver(None, {}, {}, {'device': DEV, 'num_neg_candidates': 99, 'batch_size': 128})
    oracle.eval(); model.eval()
    np.random.seed(99)
    (hr_o, nd_o, auc_o, l_o), per = osolver.metrics(oracle, ds, 99, return_per_user=True)
    np.random.seed(99)
    (hr_m, nd_m, auc_m, l_m), per_m = solver.metrics(1, 1, model, ds, return_per_user=True)
    ranks_m = per_m[:, 34].cpu().numpy().astype(np.int64)
    agree = (ranks_m == per['ranks']).mean()
    assert agree >= 0.98, agree                      # a fp32 near-tie may flip a neighbour pair
    if agree == 1.0:
        assert np.array_equal(hr_m, hr_o) and np.allclose(nd_m, nd_o, rtol=1e-12)
    assert abs(hr_m[5] - hr_o[5]) <= 2.0 / len(per['ranks']) and abs(nd_m[5] - nd_o[5]) <= 2.0 / len(per['ranks'])
    assert abs(auc_m[0] - auc_o[0]) < 1e-3 and abs(l_m[0] - l_o[0]) / abs(l_o[0]) < 1e-4


@pytest.mark.parametrize('extra', [[], ['--cuda_graph=true', '--device_sampling=true']])
def test_experiment_cli_trains_evaluates_checkpoints_and_resumes(tmp_path, monkeypatch, capsys, extra):
    """reference experiments/peagcn_solver_bpr.py flags -> BaseSolver.run(): init eval, 2 epochs,
    checkpoint + global logger written in the reference's layout, resume from latest.pkl."""
    import glob
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'experiments'))
    import pea_cli
    from graph_recsys_benchmark_b200 import models
    monkeypatch.chdir(tmp_path)
    argv = ['--dataset=Movielens', '--dataset_name=latest-small', '--synthetic=tiny', '--sampling_strategy=unseen',
            '--runs=1', '--epochs=2', '--batch_size=512', '--save_every_epoch=0', '--save_epochs=1',
            '--metapath_test=false', '--num_workers=0'] + extra
    pea_cli.run('PEAGCN', models.PEAGCNRecsysModel, argv=argv)
    out = capsys.readouterr().out
    assert 'Initial performance HR@5' in out and 'Run: 1, epoch: 2, HR@5' in out and 'Duration' in out
    ck = glob.glob(os.path.join(str(tmp_path), 'checkpoint', 'weights', 'Movielenslatest-small', 'PEAGCN', 'BPR', '*', 'run_1', 'latest.pkl'))
    assert len(ck) == 1
    state = torch.load(ck[0], map_location='cpu', weights_only=False)
    assert state['epoch'] == 2 and set(state) == {'epoch', 'model_states', 'optim_states', 'rec_metrics'}
    assert state['rec_metrics'][0].shape == (2, 16) and 'pea_channels.0.gnn_layers.0.weight' in state['model_states']['model']
    assert os.path.exists(os.path.join(os.path.dirname(ck[0]), '1.pkl'))
    logs = glob.glob(os.path.join(str(tmp_path), 'checkpoint', 'loggers', '*', 'PEAGCN', 'BPR', '*', 'logger_file.txt'))
    assert len(logs) == 1 and 'Run: 1, epoch: 1' in open(logs[0]).read()
    # a finished run is not repeated (global_logger.pkl holds it)
    pea_cli.run('PEAGCN', models.PEAGCNRecsysModel, argv=argv)
    assert 'epoch: 1' not in capsys.readouterr().out
    # a fresh logger folder with the old weights resumes at epoch 3
    os.remove(os.path.join(os.path.dirname(logs[0]), 'global_logger.pkl'))
    pea_cli.run('PEAGCN', models.PEAGCNRecsysModel, argv=[a if not a.startswith('--epochs') else '--epochs=3' for a in argv])
    out = capsys.readouterr().out
    assert "Loaded checkpoint_backup" in out and 'Run: 1, epoch: 3, HR@5' in out and 'Run: 1, epoch: 2, HR@5' not in out


@pytest.mark.parametrize('kind', ['gcn', 'sage'])
@pytest.mark.parametrize('aggr', ['att', 'mean'])
@pytest.mark.parametrize('entity_aware', [False, True])
def test_fused_engine_equals_layer_path(kind, aggr, entity_aware):
    """engine.py (two autograd nodes, planned buffers) and the per-layer modules run the same
    kernels: loss, representation and every gradient agree to rounding."""
    from graph_recsys_benchmark_b200.engine import GcnPlan
    ds = _dataset()
    batch = _batch(ds, 300, entity_aware).to(DEV)
    torch.manual_seed(11)
    model = product_model_for(ds, kind, entity_aware=entity_aware, channel_aggr=aggr)
    assert GcnPlan.applies(model, kind)
    model.train()
    res = {}
    for fused in (True, False):
        model.fused_engine = fused
        model.zero_grad()
        loss = model.loss(batch)
        loss.backward()
        res[fused] = (loss.item(), model.cached_repr.detach().clone(),
                      {n: p.grad.clone() for n, p in model.named_parameters()})
    assert abs(res[True][0] - res[False][0]) <= 1e-6 * abs(res[False][0])
    assert rel_err(res[True][1], res[False][1]) < 1e-6
    for n, g in res[False][2].items():
        if float(g.abs().max()) > 1e-12:
            # two fp32 schedules of the same sums (each within 1e-5 of the fp64 oracle, test_model_loss_and_all_gradients)
            assert rel_err(res[True][2][n], g) < 2e-5, n
    for idx in (0, 4):                                   # ablation goes through the engine too
        outs = []
        for fused in (True, False):
            model.fused_engine = fused
            model.eval(idx)
            outs.append(model.cached_repr.clone())
        assert rel_err(outs[0], outs[1]) < 1e-6


def test_engine_declines_other_shapes():
    from graph_recsys_benchmark_b200.engine import GcnPlan
    ds = _dataset()
    assert not GcnPlan.applies(product_model_for(ds, 'sage'), 'gcn')
    assert not GcnPlan.applies(product_model_for(ds, 'gcn'), 'sage')
    assert not GcnPlan.applies(product_model_for(ds, 'gat'), 'gcn') and not GcnPlan.applies(product_model_for(ds, 'gat'), 'sage')
    m = product_model_for(ds, 'gcn', hidden=16)          # hidden == repr: both steps aggregate first
    assert not GcnPlan.applies(m)
    m.eval()                                             # ... and the per-layer path serves it
    assert m.cached_repr.shape == (ds.num_nodes, 16) and bool(torch.isfinite(m.cached_repr).all())


@pytest.mark.parametrize('kind', ['gcn', 'gat', 'sage'])
def test_product_matches_frozen_golden_vectors(kind):
    """The committed golden vectors (tests/golden/oracle_vectors.pt, frozen oracle outputs on the
    seeded tiny HIN): same triples, loss / representation / gradient within 1e-5 / 1e-4, same ranks."""
    import os
    frozen = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'oracle_vectors.pt'),
                        weights_only=False)
    g = frozen[kind]
    ds = _dataset(entity_aware=g['entity_aware'])
    torch.manual_seed(2020)
    oracle = oracle_model_for(ds, kind, entity_aware=g['entity_aware'])       # same init draw as the script
    model = product_model_for(ds, kind, entity_aware=g['entity_aware'])
    model.load_state_dict(oracle.state_dict())
    osolver.seed_everything(1)
    ds.cf_negative_sampling()
    batch = ds.get_batch(list(range(128)))
    assert torch.equal(batch, g['batch'])                                       # integer: bit-exact
    model.train()
    loss = model.loss(batch.to(DEV))
    loss.backward()
    assert abs(loss.item() - g['loss']) <= 1e-5 * abs(g['loss'])
    assert rel_err(model.cached_repr[:8], g['repr_rows']) < 1e-5
    assert abs(float(model.cached_repr.detach().double().sum()) - g['repr_sum']) <= 1e-4 * abs(g['repr_sum']) + 1e-4
    assert rel_err(model.x.grad[:4], g['x_grad_rows']) < 1e-4
    from graph_recsys_benchmark_b200.solvers import BaseSolver
    solver = BaseSolver(None, {}, {}, {'device': DEV, 'num_neg_candidates': 99, 'batch_size': 128})
    model.eval()
    np.random.seed(99)
    (hr, nd, auc, el), per = solver.metrics(1, 1, model, ds, return_per_user=True)
    ranks = per[:, 34].long().cpu()
    assert float((ranks == g['ranks']).float().mean()) >= 0.975
    if torch.equal(ranks, g['ranks']):
        assert hr[5] == g['hr10'] and abs(nd[5] - g['ndcg10']) < 1e-12
    assert abs(auc[0] - g['auc']) < 2e-3 and abs(el[0] - g['eval_loss']) <= 1e-4 * abs(g['eval_loss'])


@pytest.mark.parametrize('kind', ['gcn', 'sage', 'gat'])
def test_cuda_graph_step_equals_eager_steps(kind):
    """graphed.GraphedTrainStep: replaying the captured step (loss, backward, Adam) leaves the same weights
    and losses as launching every kernel eagerly; a batch of another shape takes the eager path."""
    import copy
    from graph_recsys_benchmark_b200.graphed import GraphedTrainStep
    ds = _dataset()
    batches = [_batch(ds, 256, False, seed=i).to(DEV) for i in range(5)]
    torch.manual_seed(5)
    eager = product_model_for(ds, kind)
    replay = product_model_for(ds, kind)
    replay.load_state_dict(copy.deepcopy(eager.state_dict()))
    losses = {}
    for name, model in (('eager', eager), ('graph', replay)):
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3, fused=True, capturable=True)
        model.train()
        out = []

        def one(b):
            opt.zero_grad(set_to_none=True)
            loss = model.loss(b)
            loss.backward()
            opt.step()
            return loss.detach()
        out.append(float(one(batches[0])))                      # the eager step the capture needs first
        if name == 'graph':
            stepper = GraphedTrainStep(model, opt, batches[1])  # trains on batches[1] eagerly, then captures
            assert stepper.launches_per_replay > 20
            out.append(float(stepper.first_loss))
        else:
            stepper = one
            out.append(float(one(batches[1])))
        for b in batches[2:]:
            out.append(float(stepper(b)))
        out.append(float(stepper(batches[0][:100])))             # other shape -> eager fallback inside the wrapper
        losses[name] = out
    # the scoring kernel scatters d_repr with float atomics, and Adam turns a last-bit gradient difference
    # into an update difference of up to ~lr on near-zero gradients: weights agree to a fraction of one update
    for a, b in zip(losses['eager'], losses['graph']):
        assert abs(a - b) <= 1e-4 * abs(a), losses
    for (n, p), (_, q) in zip(eager.named_parameters(), replay.named_parameters()):
        assert float((p - q).detach().abs().max()) < 2e-3, n          # 6 steps x lr 1e-3 is the largest possible drift
        assert rel_err(q, p) < 5e-3, n


@pytest.mark.parametrize('shape,strategy,entity_aware', [('tiny', 'random', False), ('tiny', 'unseen', True),
                                                         ('ml-small', 'unseen', False), ('ml-small', 'random', True)])
def test_device_sampler_rows_equal_cpu_mirror(shape, strategy, entity_aware):
    """peagnn_bpr_rows against oracle/device_sampler.py: integer rows, bit-exact."""
    from oracle import device_sampler as ods
    from graph_recsys_benchmark_b200.datasets import SyntheticHIN
    from graph_recsys_benchmark_b200.sampling import DeviceBprSampler
    ds = SyntheticHIN(shape, seed=7, entity_aware=entity_aware, sampling_strategy=strategy)
    smp = DeviceBprSampler(ds, DEV, seed=(1 << 40) + 12345)
    h = smp.host
    n = len(smp)
    ids = np.unique(np.concatenate([np.arange(min(n, 3000)), np.random.RandomState(0).randint(0, n, 3000), [n - 1]]))
    got = smp.rows(torch.from_numpy(ids), epoch=5).cpu().numpy()
    want = ods.bpr_rows(ids, h['u2i'], num_neg=smp.num_neg, seed=smp.seed, epoch=5, strategy=smp.strategy,
                        user_lo=smp.user_lo, item_lo=smp.item_lo, num_items=smp.num_items,
                        seen_ptr=h.get('seen_ptr'), seen_items=h.get('seen_items'), cols=smp.cols,
                        ifeat=(h['ifeat_ptr'], h['ifeat_nids']) if entity_aware else None,
                        ufeat=(h['ufeat_ptr'], h['ufeat_nids']) if entity_aware else None,
                        type_starts=h.get('type_starts'))
    assert got.shape == want.shape == (len(ids), 9 if entity_aware else 3)
    assert np.array_equal(got, want)
    perm = smp.permutation(2)
    assert torch.equal(torch.sort(perm).values, torch.arange(n, device=DEV))      # a permutation of the table's rows
    assert torch.equal(perm, smp.permutation(2)) and not torch.equal(perm, smp.permutation(3))


def test_solver_trains_with_device_sampling_and_cuda_graph(tmp_path):
    """BaseSolver.train_epoch with train_args['device_sampling'] + ['cuda_graph']: the loss goes down."""
    from graph_recsys_benchmark_b200.solvers import BaseSolver
    ds = SyntheticHIN_small()
    torch.manual_seed(3)
    model = product_model_for(ds, 'gcn')
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=1e-3, capturable=True)
    solver = BaseSolver(None, {}, {}, {'device': 'cuda', 'batch_size': 128, 'quiet': True, 'device_sampling': True,
                                       'cuda_graph': True, 'loss_sync_every': 10})
    n_batches = -(-4 * ds.edge_index_nps['user2item'].shape[1] // 128)      # the last one is short -> eager fallback
    losses = []
    for epoch in (1, 2, 3):
        _, ep_losses = solver.train_epoch(1, epoch, model, opt, ds)
        assert len(ep_losses) == n_batches
        losses.extend(ep_losses)
    full = [l for i, l in enumerate(losses) if (i + 1) % n_batches]           # drop the short batches (sum loss scales with B)
    assert np.isfinite(losses).all()
    assert np.mean(full[-10:]) < 0.9 * np.mean(full[:10])


def SyntheticHIN_small():
    from graph_recsys_benchmark_b200.datasets import SyntheticHIN
    return SyntheticHIN('tiny', seed=7, sampling_strategy='unseen')


@pytest.mark.parametrize('kind', ['gcn', 'sage', 'gat', 'gcn+fused-first-projection'])
@pytest.mark.parametrize('entity_aware', [False, True])
@pytest.mark.parametrize('shape', ['tiny', 'ml-small'])
def test_demand_driven_loss_equals_full_propagation(kind, entity_aware, shape):
    """loss() with ``demand_driven_loss`` computes the last step only on the batch's user / item rows
    (reference models/base.py:209-210 reads nothing else): same loss, same gradients, and every
    representation row it does compute is the full propagation's row."""
    ds = _dataset(shape)
    batch = _batch(ds, 512, entity_aware).to(DEV)
    model = product_model_for(ds, kind.split('+')[0], entity_aware=entity_aware)
    model.fuse_first_projection = kind.endswith('fused-first-projection')     # peagnn_spmm_proj in the head (off by default)
    model.train()
    full = model.loss(batch)
    full.backward()
    ref_repr = model.cached_repr.detach().clone()
    ref_grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    model.zero_grad()
    model.demand_driven_loss = True
    lean = model.loss(batch)
    lean.backward()
    rows = torch.unique(batch[:, :3].reshape(-1))
    # (PEAGCN: the first projection rides in the aggregation's epilogue as plain fp32 FMAs there, as a 3xTF32 tensor-core
    # product on the full path - the same numbers to fp32 rounding, not bit for bit)
    assert rel_err(model.cached_repr.detach()[rows], ref_repr[rows]) < 2e-6
    assert abs(lean.item() - full.item()) <= 2e-6 * abs(full.item())
    for n, p in model.named_parameters():
        if float(ref_grads[n].abs().max()) > 1e-12:
            # two fp32 evaluations of the same sums in different orders (a bias gradient is summed over the [3B] batch rows
            # here, over all N rows - mostly zeros - there; observed up to 2.1e-5 on the ML-small GAT biases).  Each of the
            # two is judged against the reference's fp64 run on its own in tests/test_gpu_reference_fixtures.py
            assert rel_err(p.grad, ref_grads[n]) < 5e-5, n
    # and it is reproducible run to run
    model.zero_grad()
    again = model.loss(batch)
    again.backward()
    assert again.item() == lean.item()


def test_three_step_gat_channel_on_the_rows_its_batch_reads():
    """A 3-step PEAGAT channel (the reference's tables are all 2-step; ``meta_path_steps`` allows more): with the batch
    rows given, every earlier step runs on the rows the later ones read (source ranges accumulate step by step,
    PEABaseChannel._needed_rows) - the batch rows of the output and all gradients equal the full channel's."""
    from graph_recsys_benchmark_b200 import functional as F_
    from graph_recsys_benchmark_b200.models.families import PEAGATChannel
    from helpers import random_edge_index
    torch.manual_seed(11)
    n = 600
    # three relations over four "types": 0-99 -> 100-299 -> 300-449 -> 450-599
    e1 = random_edge_index(n, 2500, 1, src_range=(0, 100), dst_range=(100, 300))
    e2 = random_edge_index(n, 2500, 2, src_range=(100, 300), dst_range=(300, 450))
    e3 = random_edge_index(n, 2500, 3, src_range=(300, 450), dst_range=(450, 600))
    eil = [e.to(DEV) for e in (e1, e2, e3)]
    ch = PEAGATChannel(num_steps=3, num_nodes=n, dropout=0., emb_dim=64, hidden_size=64, repr_dim=16, num_heads=1).to(DEV)
    x = torch.randn(n, 64, device=DEV)
    batch_rows = torch.randint(450, 600, (40,), device=DEV)
    active = F_.active_rows(batch_rows, n)
    mask = torch.zeros(n, dtype=torch.bool, device=DEV)
    mask[batch_rows] = True
    w = torch.randn(n, 16, device=DEV) * mask[:, None]

    def run(act):
        ch.zero_grad()
        xs = x.clone().requires_grad_(True)
        y = ch.forward_split(xs, eil, None, False, act)
        (y * w).sum().backward()
        return y.detach(), [xs.grad.clone()] + [p.grad.clone() for p in ch.parameters()]
    y0, g0 = run(None)
    y1, g1 = run(active)
    needed = ch._needed_rows(eil, n, active)
    assert needed[0].ranges == [(100, 450)] and needed[1].ranges == [(300, 450)] and needed[2] is None
    assert rel_err(y1[mask], y0[mask]) < 2e-6 and float(y1[~mask].abs().max()) == 0.
    for a, b in zip(g1, g0):
        if float(b.abs().max()) > 1e-12:
            assert rel_err(a, b) < 5e-5
