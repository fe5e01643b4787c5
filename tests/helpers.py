"""Shared test helpers: small random graphs, oracle/product model pairs."""
import numpy as np
import torch

from oracle import graph as ograph
from oracle.models import OraclePEAModel

KIND_OF = {'PEAGCNRecsysModel': 'gcn', 'PEAGATRecsysModel': 'gat', 'PEASageRecsysModel': 'sage'}


def random_edge_index(n, e, seed, self_loops=0, multi=0, src_range=None, dst_range=None):
    g = torch.Generator().manual_seed(seed)
    s_lo, s_hi = src_range or (0, n)
    d_lo, d_hi = dst_range or (0, n)
    src = torch.randint(s_lo, s_hi, (e,), generator=g)
    dst = torch.randint(d_lo, d_hi, (e,), generator=g)
    if multi:
        src = torch.cat([src, src[:multi]])
        dst = torch.cat([dst, dst[:multi]])
    if self_loops:
        l = torch.randint(0, n, (self_loops,), generator=g)
        src = torch.cat([src, l])
        dst = torch.cat([dst, l])
    return torch.stack([src, dst]).long()


def model_kwargs(dataset, kind, **kw):
    from graph_recsys_benchmark_b200.utils.factory import default_model_args
    return default_model_args(dataset, kind, **kw)


def oracle_model_for(dataset, kind, dtype=torch.float32, **kw):
    """OraclePEAModel on the dataset's metapaths (CPU)."""
    eil = ograph.metapath_edge_index_list(dataset, dataset.dataset, dataset.name)
    mk = model_kwargs(dataset, kind, **kw)
    m = OraclePEAModel(kind, dataset.num_nodes, eil, mk['meta_path_steps'], emb_dim=mk['emb_dim'],
                       hidden_size=mk['hidden_size'], repr_dim=mk['repr_dim'], num_heads=mk.get('num_heads', 1),
                       dropout=0.0, channel_aggr=mk['channel_aggr'], entity_aware=mk['entity_aware'],
                       entity_aware_coff=mk['entity_aware_coff'])
    return m.to(dtype)


def product_model_for(dataset, kind, device='cuda', **kw):
    from graph_recsys_benchmark_b200.utils.factory import build_model
    return build_model(dataset, kind, device=device, **kw)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# Parameters whose gradient contains ONE relu decision that is a tie in fp32: with the fixture seed a first-layer
# pre-activation is zero to within fp32 rounding (GCN: metapath 2, row 21080 - profiles/r2_shard_relu_tie.txt; GAT:
# metapath 8, row 21851 column 26, -7.4e-10 in the reference's fp64 run against a row scale of 4e-2 -
# tools/gat_grad_diag.py), so an fp32 sum in another order lands on the other side of the relu than the fp64
# reference and that row's upstream gradient (9.2e-4 of the largest entry for the GAT case) enters or leaves the
# layer's weight / bias gradient.  A property of fp32, named here instead of widening every bound.
RELU_TIES = {
    'ml-25m-lite/gcn/plain': ('pea_channels.2.gnn_layers.0.weight', 'pea_channels.2.gnn_layers.0.bias'),
    'ml-25m-lite/gat/plain': ('pea_channels.8.gnn_layers.0.lin.weight', 'pea_channels.8.gnn_layers.0.bias',
                              'pea_channels.8.gnn_layers.0.att_i', 'pea_channels.8.gnn_layers.0.att_j'),
}


def grad_bound(fixture_key, fx, name, base=1e-4):
    """Bound on a parameter gradient's max-norm relative error against the reference's fp64 run: ``base`` (1e-4),
    unless the reference's OWN fp32 run is already further away than a third of that (``f32['grad_err']``, written
    by tests/golden/make_reference_fixtures.py: long fp32 column sums with cancellation - the GAT bias gradients on
    the 30 k-node graph are 4e-4 off in the reference itself), or the entry holds a relu tie."""
    if name in RELU_TIES.get(fixture_key, ()):
        return 1.5e-3
    own = fx['f32'].get('grad_err', {}).get(name, 0.0)
    return max(base, 3.0 * own)


# ---------------------------------------------------------------------------------------------
# the run tests/golden/make_reference_fixtures.py performs with the reference's code, restated
# with the oracle (same seeds, same order of RNG consumption, same outputs)
# ---------------------------------------------------------------------------------------------
def state_sha(sd):
    import hashlib
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def tensor_sha(t):
    import hashlib
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def sample_rows(n, k=512):
    if n <= k:
        return torch.arange(n)
    return torch.randperm(n, generator=torch.Generator().manual_seed(0))[:k].sort().values


def seed_all(seed):
    import random as rd
    rd.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def oracle_run(ds, kind, ea, B, evaluate=True, dtype=torch.float32, run=1, n_steps=3):
    """Oracle counterpart of ``reference_run`` (tests/golden/make_reference_fixtures.py): every step the
    reference's code takes there, taken with oracle/ instead."""
    import hashlib
    from oracle import sampling as osampling, solver as osolver
    seed_all(2019 + run)
    model = oracle_model_for(ds, kind, entity_aware=ea)
    out = {'state_sha': state_sha(model.state_dict()), 'state': {k: v.clone() for k, v in model.state_dict().items()}}
    model = model.to(dtype)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    train_data = osampling.cf_negative_sampling_bpr(ds, ds.num_negative_samples, ds.sampling_strategy)
    out['train_rows'] = int(train_data.shape[0])
    out['train_head'] = train_data[:1000].clone()
    out['train_sha'] = tensor_sha(train_data)
    batches = [torch.stack([osampling.getitem(ds, train_data, i, ea) for i in range(s * B, (s + 1) * B)])
               for s in range(n_steps)]
    out['batches'] = torch.stack(batches)
    rows = sample_rows(ds.num_nodes)
    out['rows'] = rows
    model.train()
    losses = []
    for s in range(n_steps):
        opt.zero_grad()
        loss = model.loss(batches[s])
        loss.backward()
        if s == 0:
            out['repr'] = model.cached_repr.detach().clone()
            out['repr_rows'] = out['repr'][rows].clone()
            out['repr_abs_sum'] = float(out['repr'].double().abs().sum())
            grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
            out['x_grad'] = grads['x']
            out['x_grad_rows'] = grads['x'][rows].clone()
            out['x_grad_abs_sum'] = float(grads['x'].double().abs().sum())
            out['grads'] = {k: g for k, g in grads.items() if k != 'x'}
        opt.step()
        losses.append(loss.detach().cpu().item())
    out['losses'] = losses
    if evaluate:
        model.eval()
        out['eval_repr_rows'] = model.cached_repr[rows].clone()
        np.random.seed(4000 + run)
        (HR, NDCG, AUC, eloss), per = osolver.metrics(model, ds, 99, return_per_user=True)
        out['HR'], out['NDCG'], out['AUC'], out['eval_loss'] = HR, NDCG, AUC, eloss
        out['ranks'] = torch.tensor(per['ranks'], dtype=torch.int16)
        np.random.seed(4000 + run)
        cand = np.stack([np.asarray(list(p) + list(n), dtype=np.int64) for p, n in
                         (osolver.generate_candidates(ds, u, 99) for u in list(ds.test_pos_unid_inid_map.keys()))])
        out['cand'] = cand
        out['cand_sha'] = hashlib.sha256(cand.tobytes()).hexdigest()
        out['cand_head'] = torch.from_numpy(cand[:256].copy())
    out['model'] = model
    return out
