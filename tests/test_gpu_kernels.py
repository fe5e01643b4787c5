"""GPU parity tests, kernel by kernel, through the C ABI (functional.py is a ctypes shim).
Oracle = oracle/ (CPU); tolerance: bit-exact for integers, 1e-5 relative (max-norm) for fp32 as
BASELINE.json's north_star states (gradients through long fp32 sums get 1e-4)."""
import numpy as np
import pytest
import torch

from helpers import random_edge_index, rel_err
from oracle import graph as ograph, pyg150

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = 'cuda'


def _graph(ei, n, **kw):
    from graph_recsys_benchmark_b200.graph import RelationGraph
    return RelationGraph.from_edge_index(ei.to(DEV), n, **kw)


@pytest.mark.parametrize('n,e,loops,multi', [(6, 0, 0, 0), (50, 400, 5, 30), (3000, 60000, 50, 1000), (10, 5000, 3, 0)])
def test_csr_build_bit_exact(n, e, loops, multi):
    ei = random_edge_index(n, e, 11, self_loops=loops, multi=multi)
    g = _graph(ei, n)
    for csr, key, val in ((g.fwd, ei[1], ei[0]), (g.bwd, ei[0], ei[1])):
        rp, col, eid = ograph.csr_by_key(key.numpy(), val.numpy(), n, drop_self_loops=True)
        assert np.array_equal(csr.rowptr.cpu().numpy(), rp)
        assert np.array_equal(csr.col.cpu().numpy(), col)
        assert np.array_equal(csr.eid.cpu().numpy(), eid)
    if g.fwd.nnz:
        perm = g.bwd_to_fwd.cpu().long()
        assert torch.equal(g.fwd.eid.cpu()[perm], g.bwd.eid.cpu())


def _conv_pair(kind, fin, fout, heads=1):
    from graph_recsys_benchmark_b200 import nn as pnn
    torch.manual_seed(3)
    if kind == 'gcn':
        o, p = pyg150.GCNConv(fin, fout), pnn.PEAGCNConv(fin, fout)
        o.bias.data.uniform_(-0.5, 0.5)
    elif kind == 'sage':
        o, p = pyg150.SAGEConv(fin, fout), pnn.PEASageConv(fin, fout)
    else:
        o, p = pyg150.GATConv(fin, fout, heads=heads), pnn.PEAGATConv(fin, fout, heads=heads)
        o.bias.data.uniform_(-0.5, 0.5)
    p.load_state_dict(o.state_dict())
    return o, p.to(DEV)


GRAPHS = {
    'small': dict(n=60, e=500, loops=4, multi=40),
    'isolated': dict(n=200, e=150, loops=0, multi=0),
    'heavy': dict(n=300, e=40000, loops=10, multi=500),      # rows far above the heavy threshold
    'bipartite': dict(n=400, e=6000, loops=0, multi=100, src_range=(0, 100), dst_range=(100, 400)),
}


@pytest.mark.parametrize('kind', ['gcn', 'sage', 'gat'])
@pytest.mark.parametrize('gname', list(GRAPHS))
@pytest.mark.parametrize('fin,fout', [(64, 64), (64, 16), (16, 64)])
@pytest.mark.parametrize('relu', [False, True])
def test_conv_forward_backward(kind, gname, fin, fout, relu):
    spec = dict(GRAPHS[gname])
    n = spec.pop('n')
    ei = random_edge_index(n, spec.pop('e'), 5, self_loops=spec.pop('loops'), multi=spec.pop('multi'), **spec)
    from graph_recsys_benchmark_b200 import graph as pgraph
    old = pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES
    if gname == 'heavy':
        pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES = 64, 96    # forces multi-chunk rows at test size
    try:
        pgraph.clear_cache()
        o, p = _conv_pair(kind, fin, fout)
        x = torch.randn(n, fin)
        xo = x.clone().double().requires_grad_(True)
        xp = x.clone().to(DEV).requires_grad_(True)
        o64 = o.double()
        yo = o64(xo, ei)
        if relu:
            yo = torch.relu(yo)
        yp = p(xp, ei.to(DEV), relu=relu)
        assert rel_err(yp, yo) < TOL
        w = torch.randn(n, yo.shape[1])
        (yo * w.double()).sum().backward()
        (yp * w.to(DEV)).sum().backward()
        assert rel_err(xp.grad, xo.grad) < 10 * TOL
        for (name, po), (_, pp) in zip(o64.named_parameters(), p.named_parameters()):
            assert rel_err(pp.grad, po.grad) < 10 * TOL, name
    finally:
        pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES = old
        pgraph.clear_cache()


@pytest.mark.parametrize('heads', [2, 4])
def test_gat_multi_head(heads):
    n = 120
    ei = random_edge_index(n, 1500, 9, self_loops=5, multi=60)
    o, p = _conv_pair('gat', 64, 16, heads=heads)
    x = torch.randn(n, 64)
    xo = x.clone().double().requires_grad_(True)
    xp = x.clone().to(DEV).requires_grad_(True)
    o64 = o.double()
    yo, yp = o64(xo, ei), p(xp, ei.to(DEV))
    assert yp.shape == (n, heads * 16) and rel_err(yp, yo) < TOL
    w = torch.randn(n, heads * 16)
    (yo * w.double()).sum().backward()
    (yp * w.to(DEV)).sum().backward()
    assert rel_err(xp.grad, xo.grad) < 10 * TOL
    for (name, po), (_, pp) in zip(o64.named_parameters(), p.named_parameters()):
        assert rel_err(pp.grad, po.grad) < 10 * TOL, name


def test_conv_fp32_oracle_agreement_is_rounding_level():
    """Against the fp32 oracle the gap is rounding only (different summation order)."""
    n = 500
    ei = random_edge_index(n, 20000, 2, multi=300)
    o, p = _conv_pair('gcn', 64, 64)
    x = torch.randn(n, 64)
    assert rel_err(p(x.to(DEV), ei.to(DEV)), o(x, ei)) < TOL


@pytest.mark.parametrize('P,D,mode', [(9, 16, 'att'), (13, 16, 'att'), (11, 16, 'mean'), (1, 16, 'att'), (9, 64, 'att'), (5, 8, 'att')])
def test_fuse(P, D, mode):
    from graph_recsys_benchmark_b200 import functional as F_
    n = 777
    z = torch.randn(n, P, D)
    att = torch.randn(1, P, D) * 0.5
    zo = z.clone().double().requires_grad_(True)
    ao = att.clone().double().requires_grad_(True)
    zp = z.clone().to(DEV).requires_grad_(True)
    ap = att.clone().to(DEV).requires_grad_(True)
    if mode == 'att':
        w = torch.softmax((zo * ao).sum(-1), dim=-1).unsqueeze(-1)
        ro = (zo * w).sum(1)
    else:
        ro = zo.mean(1)
    rp = F_.fuse_channels(zp, ap if mode == 'att' else None, mode)
    assert rel_err(rp, ro) < TOL
    g = torch.randn(n, D)
    (ro * g.double()).sum().backward()
    (rp * g.to(DEV)).sum().backward()
    assert rel_err(zp.grad, zo.grad) < 10 * TOL
    if mode == 'att':
        assert rel_err(ap.grad, ao.grad) < 10 * TOL


@pytest.mark.parametrize('skip', [0, 4, 8])
def test_fuse_ablation(skip):
    from graph_recsys_benchmark_b200 import functional as F_
    n, P, D = 300, 9, 16
    z = torch.randn(n, P, D)
    att = torch.randn(1, P, D)
    zz = z.clone().double()
    zz[:, skip] = 0
    w = torch.softmax((zz * att.double()).sum(-1), dim=-1).unsqueeze(-1)
    ro = (zz * w).sum(1)
    with torch.no_grad():
        rp = F_.fuse_channels(z.to(DEV), att.to(DEV), 'att', skip)
    assert rel_err(rp, ro) < TOL


def _fc(D):
    torch.manual_seed(4)
    return torch.nn.Linear(2 * D, D), torch.nn.Linear(D, 1)


@pytest.mark.parametrize('B', [1, 37, 1024, 4096])
def test_bpr_loss_and_grads(B):
    from graph_recsys_benchmark_b200 import functional as F_
    n, D = 500, 16
    fc1, fc2 = _fc(D)
    r = torch.randn(n, D)
    batch = torch.randint(0, n, (B, 3))
    ro = r.clone().double().requires_grad_(True)
    f1, f2 = torch.nn.Linear(2 * D, D).double(), torch.nn.Linear(D, 1).double()
    f1.load_state_dict(fc1.state_dict()); f2.load_state_dict(fc2.state_dict())

    def pred(u, i):
        return f2(torch.relu(f1(torch.cat([ro[u], ro[i]], dim=-1))))
    lo = -(pred(batch[:, 0], batch[:, 1]) - pred(batch[:, 0], batch[:, 2])).sigmoid().log().sum()
    lo.backward()
    rp = r.clone().to(DEV).requires_grad_(True)
    ps = [t.detach().clone().to(DEV).requires_grad_(True) for t in (fc1.weight, fc1.bias, fc2.weight, fc2.bias)]
    lp = F_.bpr_loss(rp, ps[0], ps[1], ps[2], ps[3], batch.to(DEV))
    lp.backward()
    assert abs(lp.item() - lo.item()) / abs(lo.item()) < TOL
    assert rel_err(rp.grad, ro.grad) < 10 * TOL
    for pp, po in zip(ps, (f1.weight, f1.bias, f2.weight, f2.bias)):
        assert rel_err(pp.grad, po.grad) < 10 * TOL or float(po.grad.abs().max()) < 1e-12
    # predict()
    sp = F_.predict_raw(rp.detach(), batch[:, 0].to(DEV), batch[:, 1].to(DEV), ps[0], ps[1], ps[2], ps[3])
    assert sp.shape == (B, 1) and rel_err(sp, pred(batch[:, 0], batch[:, 1])) < TOL


def test_entity_regulariser():
    from graph_recsys_benchmark_b200 import functional as F_
    n, E, B = 400, 64, 513
    x = torch.randn(n, E) * 0.3
    batch = torch.randint(0, n, (B, 9))
    batch[:, 5] = torch.randint(0, 2, (B,))
    batch[:, 8] = torch.randint(0, 2, (B,))
    xo = x.clone().double().requires_grad_(True)

    def sq(a, b):
        d = xo[a] - xo[b]
        return (d * d).sum(-1)
    it = (sq(batch[:, 1], batch[:, 3]) - sq(batch[:, 1], batch[:, 4])) * batch[:, 5]
    us = (sq(batch[:, 0], batch[:, 6]) - sq(batch[:, 0], batch[:, 7])) * batch[:, 8]
    lo = 0.1 * (-it.sigmoid().log().sum() - us.sigmoid().log().sum())
    lo.backward()
    xp = x.clone().to(DEV).requires_grad_(True)
    lp = F_.entity_reg(xp, batch.to(DEV), 0.1)
    lp.backward()
    assert abs(lp.item() - lo.item()) / abs(lo.item()) < TOL
    assert rel_err(xp.grad, xo.grad) < 10 * TOL


@pytest.mark.parametrize('U,C,n_pos', [(1, 100, 1), (613, 100, 1), (50, 20, 3), (5, 1000, 1)])
def test_eval_rank_matches_reference_metrics(U, C, n_pos):
    from graph_recsys_benchmark_b200 import functional as F_
    from oracle import rec_utils
    n, D = 2000, 16
    fc1, fc2 = _fc(D)
    r = torch.randn(n, D)
    users = torch.randint(0, n, (U,))
    cand = torch.randint(0, n, (U, C))
    cand[:, n_pos + 3] = cand[:, n_pos + 2]          # duplicated negatives tie with each other
    args = [t.detach().to(DEV) for t in (fc1.weight, fc1.bias, fc2.weight, fc2.bias)]
    per_user, means, scores = F_.eval_rank(r.to(DEV), users.to(DEV), cand.to(DEV), n_pos, *args, return_scores=True)
    s = scores.cpu()
    # scores against the fp64 MLP
    f1, f2 = fc1.double(), fc2.double()
    ref = f2(torch.relu(f1(torch.cat([r.double()[users].unsqueeze(1).expand(-1, C, -1), r.double()[cand]], dim=-1)))).squeeze(-1)
    assert rel_err(s, ref) < TOL
    # ranking / metrics: reference procedure (solvers.py:88-96) applied to the SAME fp32 scores -> bit-exact
    pu = per_user.cpu().numpy()
    for k in range(U):
        _, idx = torch.sort(s[k], descending=True, stable=True)
        hit_vec = (idx < n_pos).numpy()
        assert np.array_equal(pu[k, :16], np.array(rec_utils.hit(hit_vec), dtype=np.float64))
        assert np.allclose(pu[k, 16:32], np.array(rec_utils.ndcg(hit_vec)), rtol=1e-14, atol=0)
        assert pu[k, 32] == rec_utils.auc(s[k, :n_pos].numpy(), s[k, n_pos:].numpy())
        assert int(pu[k, 34]) == int(np.argmax(hit_vec))
        pair = -(s[k, :n_pos].double().view(-1, 1) - s[k, n_pos:].double().view(1, -1)).sigmoid().log().sum()
        assert abs(pu[k, 33] - pair.item()) <= 1e-5 * abs(pair.item())
    assert np.allclose(means.cpu().numpy(), pu.mean(axis=0), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize('feat', [4, 8, 12, 32, 100, 112, 128, 144, 208, 256, 320])
@pytest.mark.parametrize('mode', ['gcn', 'mean'])
def test_aggregation_every_width_class(feat, mode):
    """Slot geometries G = 4 / 8 / 16 / 32 and 1, 2, 4 float4 chunks per lane, light + heavy rows,
    against an independent torch index_add reduction (fp64)."""
    from graph_recsys_benchmark_b200 import functional as F_, graph as pgraph
    n = 700
    ei = random_edge_index(n, 9000, 21, multi=400, dst_range=(0, 90))            # 90 busy rows, 610 empty
    old = pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES
    pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES = 48, 64
    try:
        g = _graph(ei, n)
        assert g.fwd.n_heavy > 0 and g.fwd.n_chunks > g.fwd.n_heavy
        x = torch.randn(n, feat)
        xd = x.double()
        src, dst = ei[0], ei[1]
        nl = src != dst
        if mode == 'gcn':
            dis = (torch.bincount(src[nl], minlength=n).double() + 1).pow(-0.5)
            ref = torch.zeros(n, feat, dtype=torch.float64).index_add_(0, dst[nl], xd[src[nl]] * (dis[src[nl]] * dis[dst[nl]]).view(-1, 1))
            ref += xd * (dis * dis).view(-1, 1)
            got = F_.spmm_raw(g.fwd, x.to(DEV), feat, torch.empty(n, feat, device=DEV), g.gcn_dis, g.gcn_dis, True)
        else:
            cnt = torch.bincount(dst[nl], minlength=n).double().clamp(min=1)
            ref = torch.zeros(n, feat, dtype=torch.float64).index_add_(0, dst[nl], xd[src[nl]]) / cnt.view(-1, 1)
            got = F_.spmm_raw(g.fwd, x.to(DEV), feat, torch.empty(n, feat, device=DEV), g.inv_in_degree, None, False)
        assert rel_err(got, ref) < TOL
    finally:
        pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES = old
        pgraph.clear_cache()


@pytest.mark.parametrize('D', [8, 32])
def test_scoring_kernels_other_repr_dims(D):
    from graph_recsys_benchmark_b200 import functional as F_
    n, B = 300, 257
    fc1, fc2 = _fc(D)
    r = torch.randn(n, D)
    batch = torch.randint(0, n, (B, 3))
    f1, f2 = torch.nn.Linear(2 * D, D).double(), torch.nn.Linear(D, 1).double()
    f1.load_state_dict(fc1.state_dict()); f2.load_state_dict(fc2.state_dict())
    ro = r.clone().double().requires_grad_(True)

    def pred(u, i):
        return f2(torch.relu(f1(torch.cat([ro[u], ro[i]], dim=-1))))
    lo = -(pred(batch[:, 0], batch[:, 1]) - pred(batch[:, 0], batch[:, 2])).sigmoid().log().sum()
    lo.backward()
    rp = r.clone().to(DEV).requires_grad_(True)
    ps = [t.detach().clone().to(DEV).requires_grad_(True) for t in (fc1.weight, fc1.bias, fc2.weight, fc2.bias)]
    lp = F_.bpr_loss(rp, ps[0], ps[1], ps[2], ps[3], batch.to(DEV))
    lp.backward()
    assert abs(lp.item() - lo.item()) / abs(lo.item()) < TOL and rel_err(rp.grad, ro.grad) < 10 * TOL
    assert rel_err(ps[0].grad, f1.weight.grad) < 10 * TOL and rel_err(ps[2].grad, f2.weight.grad) < 10 * TOL
    users, cand = torch.randint(0, n, (40,)), torch.randint(0, n, (40, 50))
    _, _, scores = F_.eval_rank(r.to(DEV), users.to(DEV), cand.to(DEV), 1, *[p.detach() for p in ps], return_scores=True)
    ref = f2(torch.relu(f1(torch.cat([r.double()[users].unsqueeze(1).expand(-1, 50, -1), r.double()[cand]], dim=-1)))).squeeze(-1)
    assert rel_err(scores, ref) < TOL


# ---- projections called directly (tensor-core 3xTF32 kernels for the hot shapes, FFMA otherwise) ---------------
@pytest.mark.parametrize('K,M', [(64, 64), (64, 16), (16, 64), (64, 32), (32, 32), (16, 16), (32, 64), (64, 48), (128, 16)])
@pytest.mark.parametrize('n', [1, 127, 129, 5000])
@pytest.mark.parametrize('variant', ['plain', 'out_in+bias+relu', 'accumulate+gate', 'in_mask', 'narrow_out'])
def test_linear_direct(K, M, n, variant):
    """peagnn_linear against an fp64 matmul, every epilogue flag, ragged row counts, strided operands."""
    from graph_recsys_benchmark_b200 import functional as F_
    g = torch.Generator().manual_seed(1000 * K + M + n)
    Xw = torch.randn(n, K + 8, generator=g).cuda()
    X = Xw[:, 4:4 + K]                                   # leading dimension > K
    out_in = variant == 'out_in+bias+relu'
    W = torch.randn((M, K) if out_in else (K, M), generator=g).cuda()
    bias = torch.randn(M, generator=g).cuda() if variant != 'plain' else None
    Y0 = torch.randn(n, M, generator=g).cuda()
    gate = torch.randn(n, M, generator=g).cuda() if variant == 'accumulate+gate' else None
    mask = torch.randn(n, K, generator=g).cuda() if variant == 'in_mask' else None
    Y = Y0.clone()
    if variant == 'narrow_out':                          # rows only 16-byte aligned: the 128-bit epilogue path
        Y = torch.zeros(n, M + 4, device='cuda')[:, 4:]
    F_.linear_raw(X, W, Y, out_in, bias, relu=out_in, accumulate=variant == 'accumulate+gate', mask=mask, out_mask=gate)
    Xd = X.double() * (mask > 0).double() if mask is not None else X.double()
    want = Xd @ (W.double().t() if out_in else W.double())
    if bias is not None:
        want = want + bias.double()
    if variant == 'accumulate+gate':
        want = (want + Y0.double()) * (gate > 0).double()
    if out_in:
        want = want.clamp_min(0)
    assert rel_err(Y, want) < 1e-5                      # fp32 tolerance of the path
    if variant == 'accumulate+gate':
        assert bool((Y[gate <= 0] == 0).all())          # gated entries are exactly zero


@pytest.mark.parametrize('K,M', [(64, 64), (64, 16), (16, 64), (64, 32), (32, 32), (16, 16), (128, 64)])
@pytest.mark.parametrize('n', [1, 63, 4097, 40000])
@pytest.mark.parametrize('masked,out_in', [(False, False), (True, True)])
def test_wgrad_direct(K, M, n, masked, out_in):
    """peagnn_linear_wgrad: dW = X^T gate(dY), db = colsum(gate(dY)) against fp64; run twice -> bit-identical."""
    from graph_recsys_benchmark_b200 import functional as F_
    g = torch.Generator().manual_seed(77 * K + M + n)
    X = torch.randn(n, K, generator=g).cuda()
    dY = torch.randn(n, M, generator=g).cuda()
    mask = torch.randn(n, M, generator=g).cuda() if masked else None
    outs = []
    for _ in range(2):
        dW = torch.full((M, K) if out_in else (K, M), float('nan'), device='cuda')
        db = torch.full((M,), float('nan'), device='cuda')
        F_.wgrad_raw(X, dY, K, M, out_in, dW, db, mask)
        outs.append((dW, db))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    d = dY.double() * (mask > 0).double() if masked else dY.double()
    want = X.double().t() @ d
    assert rel_err(outs[0][0], want.t() if out_in else want) < 1e-5
    assert rel_err(outs[0][1], d.sum(0)) < 1e-5


# ---- demand-driven aggregation (peagnn_spmm_filtered) and the deterministic gradient scatter -----------------
@pytest.mark.parametrize('gname', ['small', 'heavy', 'bipartite'])
@pytest.mark.parametrize('feat', [16, 64, 112])
@pytest.mark.parametrize('few', [False, True])
def test_filtered_aggregation_equals_the_full_one_on_what_it_computes(gname, feat, few):
    """``few``: the schedule for a filter that marks a few percent of the rows (one warp per 32-row bitmap word,
    peagnn_csr_t.sparse_filter) / the one-row-per-warp schedule - same results."""
    from graph_recsys_benchmark_b200 import functional as F_
    spec = dict(GRAPHS[gname])
    n = spec.pop('n')
    ei = random_edge_index(n, spec.pop('e'), 5, self_loops=spec.pop('loops'), multi=spec.pop('multi'), **spec)
    g = _graph(ei, n)
    torch.manual_seed(0)
    X = torch.randn(n, feat, device=DEV)
    dis = g.gcn_dis
    full = F_.spmm_raw(g.fwd, X, feat, torch.empty(n, feat, device=DEV), dis, dis, True)
    ids = torch.randperm(n, device=DEV)[:max(1, n // 7)]
    bitmap = F_.mark_rows(ids, n)
    bitmap.marked_at_most = 0 if few else n                # the host-side hint only picks the schedule
    assert F_.few_rows_marked(bitmap, n) == few
    marked = torch.zeros(n, dtype=torch.bool, device=DEV)
    marked[ids] = True
    # rows: the marked rows equal the full launch bit for bit, the others are left untouched
    out = torch.full((n, feat), 7.0, device=DEV)
    F_.spmm_raw(g.fwd, X, feat, out, dis, dis, True, active_rows=bitmap)
    assert torch.equal(out[marked], full[marked])
    assert bool((out[~marked] == 7.0).all())
    # cols: skipping the edges that gather an unmarked node == aggregating a table whose unmarked rows are zero
    Xz = X * marked[:, None]
    want = F_.spmm_raw(g.bwd, Xz, feat, torch.empty(n, feat, device=DEV), dis, dis, True)
    got = F_.spmm_raw(g.bwd, Xz, feat, torch.empty(n, feat, device=DEV), dis, dis, True, active_cols=bitmap)
    assert rel_err(got, want) < 1e-6
    again = F_.spmm_raw(g.bwd, Xz, feat, torch.empty(n, feat, device=DEV), dis, dis, True, active_cols=bitmap)
    assert torch.equal(got, again)


def test_mark_rows_with_cyclic_ownership():
    from graph_recsys_benchmark_b200 import functional as F_
    ids = torch.tensor([0, 5, 5, 64, 99, 31, 32], device=DEV)
    bm = F_.mark_rows(ids, 100).cpu().numpy().view(np.uint32)
    want = np.zeros_like(bm)
    for i in (0, 5, 64, 99, 31, 32):
        want[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    assert np.array_equal(bm, want)
    bm = F_.mark_rows(ids, 34, mod=3, rem=2).cpu().numpy().view(np.uint32)       # ids 5 (-> 1) and 32 (-> 10)
    want = np.zeros_like(bm)
    for i in (1, 10):
        want[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    assert np.array_equal(bm, want)


def test_bpr_and_entity_gradients_are_bit_reproducible():
    """The row gradients are scattered by a stable sort + one writer per node (no float atomics)."""
    from graph_recsys_benchmark_b200 import functional as F_
    torch.manual_seed(4)
    n, D, B = 500, 16, 4096
    rep = torch.randn(n, D, device=DEV, requires_grad=True)
    fc1w, fc1b = torch.randn(D, 2 * D, device=DEV) * 0.3, torch.randn(D, device=DEV) * 0.1
    fc2w, fc2b = torch.randn(1, D, device=DEV) * 0.3, torch.randn(1, device=DEV) * 0.1
    batch = torch.randint(0, n, (B, 9), device=DEV)
    batch[:, 5] = (batch[:, 5] % 2)
    batch[:, 8] = (batch[:, 8] % 2)
    x = torch.randn(n, 64, device=DEV, requires_grad=True)
    grads = []
    for _ in range(3):
        rep.grad = x.grad = None
        (F_.bpr_loss(rep, fc1w, fc1b, fc2w, fc2b, batch) + F_.entity_reg(x, batch, 0.1)).backward()
        grads.append((rep.grad.clone(), x.grad.clone()))
    for a, b in grads[1:]:
        assert torch.equal(a, grads[0][0]) and torch.equal(b, grads[0][1])
    # and the loss stays finite where a literal sigmoid().log() overflows in fp32 (z < -88.7)
    big = torch.zeros(n, D, device=DEV)
    big[1] = -40.0
    big[2] = 40.0
    w1 = torch.cat([torch.zeros(D, D), torch.eye(D)], dim=1).to(DEV)
    loss = F_.bpr_loss(big, w1, torch.zeros(D, device=DEV), torch.ones(1, D, device=DEV), torch.zeros(1, device=DEV),
                       torch.tensor([[0, 1, 2]], device=DEV))
    assert torch.isfinite(loss) and abs(loss.item() - 640.0) < 1e-3        # softplus(640) = 640


@pytest.mark.parametrize('gname', ['small', 'heavy', 'bipartite', 'isolated'])
@pytest.mark.parametrize('n_proj', [1, 2])
def test_aggregation_with_the_first_projection_in_its_epilogue(gname, n_proj):
    """peagnn_spmm_proj (north_star (2)): relu((A_hat X) W + b) computed while the aggregated row is in registers
    equals the two-launch form, on every row path of the traversal engine (light, heavy / chunked, edge-less) and with
    the row filter."""
    from graph_recsys_benchmark_b200 import functional as F_
    spec = dict(GRAPHS[gname])
    n = spec.pop('n')
    ei = random_edge_index(n, spec.pop('e'), 9, self_loops=spec.pop('loops'), multi=spec.pop('multi'), **spec)
    g = _graph(ei, n)
    torch.manual_seed(2)
    X = torch.randn(n, 64, device=DEV)
    dis = g.gcn_dis
    Ws = [torch.randn(64, 64, device=DEV) * 0.2 for _ in range(n_proj)]
    bs = [torch.randn(64, device=DEV) * 0.1 for _ in range(n_proj)]
    agg = F_.spmm_raw(g.fwd, X, 64, torch.empty(n, 64, device=DEV), dis, dis, True)
    want = [torch.relu(agg.double() @ W.double() + b.double()) for W, b in zip(Ws, bs)]
    Hs = [torch.empty(n, 64, device=DEV) for _ in range(n_proj)]
    out = F_.spmm_proj_raw(g.fwd, X, torch.empty(n, 64, device=DEV), dis, dis, True, list(zip(Ws, bs, Hs)), relu=True)
    assert torch.equal(out, agg)
    for H, w in zip(Hs, want):
        assert rel_err(H, w) < 1e-5
    ids = torch.randperm(n, device=DEV)[:max(1, n // 6)]
    bm = F_.mark_rows(ids, n)
    marked = torch.zeros(n, dtype=torch.bool, device=DEV)
    marked[ids] = True
    Hf = [torch.full((n, 64), -5.0, device=DEV) for _ in range(n_proj)]
    outf = torch.full((n, 64), -5.0, device=DEV)
    F_.spmm_proj_raw(g.fwd, X, outf, dis, dis, True, list(zip(Ws, bs, Hf)), relu=True, active_rows=bm)
    assert torch.equal(outf[marked], agg[marked]) and bool((outf[~marked] == -5.0).all())
    for H, full in zip(Hf, Hs):
        assert torch.equal(H[marked], full[marked]) and bool((H[~marked] == -5.0).all())


@pytest.mark.parametrize('gname', ['small', 'heavy', 'bipartite'])
@pytest.mark.parametrize('covering', [False, True])
@pytest.mark.parametrize('fout', [64, 16])
def test_gat_step_on_the_rows_the_next_step_reads(gname, covering, fout):
    """An earlier GAT step of a demand-driven loss() (functional.NeededRows): the marked rows equal the full pass, the
    others are zero, and with an upstream gradient that is zero outside the marked rows every gradient equals
    the full pass's.  ``covering``: the static part of the filter holds every row that has an edge (the usual case -
    no per-edge slot stays unwritten) / it does not (edges into skipped rows must read as zero on the source side)."""
    from graph_recsys_benchmark_b200 import functional as F_, graph as pgraph
    spec = dict(GRAPHS[gname])
    n = spec.pop('n')
    ei = random_edge_index(n, spec.pop('e'), 11, self_loops=spec.pop('loops'), multi=spec.pop('multi'), **spec)
    old = pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES
    if gname == 'heavy':
        pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES = 64, 96
    try:
        pgraph.clear_cache()
        _, p = _conv_pair('gat', 64, fout)
        lo, hi = (0, n) if covering else (n // 2, n)
        if covering and gname == 'bipartite':
            lo, hi = 100, 400                                # the target type's id range
        static = F_.range_bitmap(lo, hi, n, torch.device(DEV))
        extra = torch.randperm(n, generator=torch.Generator().manual_seed(3))[:7].to(DEV)
        bitmap = torch.bitwise_or(static, F_.mark_rows(extra, n))
        needed = F_.NeededRows(static, bitmap)
        mask = torch.zeros(n, dtype=torch.bool, device=DEV)
        mask[lo:hi] = True
        mask[extra] = True
        x = torch.randn(n, 64, device=DEV)
        w = torch.randn(n, fout, device=DEV) * mask[:, None]
        xa = x.clone().requires_grad_(True)
        ya = p(xa, ei.to(DEV), relu=True)
        (ya * w).sum().backward()
        ga = [xa.grad.clone()] + [q.grad.clone() for q in p.parameters()]
        p.zero_grad()
        xb = x.clone().requires_grad_(True)
        yb = p(xb, ei.to(DEV), relu=True, needed=needed)
        (yb * w).sum().backward()
        gb = [xb.grad.clone()] + [q.grad.clone() for q in p.parameters()]
        assert needed.covers(pgraph.get_graph(ei.to(DEV), n).fwd) == covering
        assert rel_err(yb[mask], ya[mask]) < 1e-6 and float(yb[~mask].abs().max() if (~mask).any() else 0.) == 0.
        for a, b in zip(ga, gb):
            assert torch.isfinite(b).all() and rel_err(b, a) < 1e-6
    finally:
        pgraph.HEAVY_THRESHOLD, pgraph.CHUNK_EDGES = old
        pgraph.clear_cache()


@pytest.mark.parametrize('K,M,out_in', [(64, 64, False), (64, 16, False), (16, 64, True), (64, 64, True), (32, 32, False)])
@pytest.mark.parametrize('count', [1, 13, 40])
def test_grouped_projections_equal_single_launches(K, M, out_in, count):
    """peagnn_linear_grouped: many problems of one shape in one launch (row counts from 0 to several tiles, strided
    outputs, bias / relu, the relu-backward gate on the way out, accumulation) - the same tiles through the same kernel
    body as peagnn_linear, so the results are bit-identical; (32, 32) has no grouped kernel and runs as single launches;
    40 problems are cut into two launches."""
    from graph_recsys_benchmark_b200 import functional as F_
    torch.manual_seed(K + M + count)
    rows = [0, 1, 127, 128, 129, 700, 3000, 5000, 12288]
    for relu, accumulate, gated in ((True, False, False), (False, True, False), (False, False, True)):
        probs, want = [], []
        for k in range(count):
            n = rows[(k * 5 + count) % len(rows)]
            X = torch.randn(n, K, device=DEV)
            W = torch.randn((M, K) if out_in else (K, M), device=DEV) * 0.3
            bias = torch.randn(M, device=DEV) if (k % 2 == 0 and not gated) else None
            wide = torch.randn(n, 2 * M, device=DEV)                    # the output is a column slice: leading dimension 2M
            out = wide[:, M:]
            gate = torch.randn(n, M, device=DEV) if gated else None
            ref = out.clone()
            if n:
                F_.linear_raw(X, W, ref, out_in, bias, relu, accumulate, out_mask=gate)
            probs.append((X, W, bias, out, gate))
            want.append(ref)
        F_.linear_grouped_raw(probs, K, M, out_in, relu=relu, accumulate=accumulate)
        for (X, W, bias, out, gate), ref in zip(probs, want):
            if K == 64 and M == 64 and 0 < X.shape[0] < 4096:
                # a single 64 -> 64 launch below 4096 rows takes the SS-form kernel, the group always the TS form
                assert rel_err(out, ref) < 2e-6
            else:
                assert torch.equal(out, ref)


@pytest.mark.parametrize('K,M', [(64, 64), (64, 16), (16, 64)])
@pytest.mark.parametrize('count', [1, 26, 35])
def test_grouped_weight_gradients(K, M, count):
    """peagnn_linear_wgrad_grouped against fp64: d W = X^T d Y and d b = column sums per problem; a problem without rows
    gets zeros; run twice: bit-identical (fixed fold order)."""
    from graph_recsys_benchmark_b200 import functional as F_
    torch.manual_seed(K * M + count)
    rows = [0, 5, 256, 257, 1000, 4096, 12288, 20000]
    probs, want = [], []
    for k in range(count):
        n = rows[(k * 3 + count) % len(rows)]
        wide = torch.randn(n, K + 16, device=DEV)
        X = wide[:, :K] if K + 16 != K else wide                          # strided input (leading dimension K + 16)
        dY = torch.randn(n, M, device=DEV)
        dW = torch.full((K, M), 7.0, device=DEV)
        db = torch.full((M,), 7.0, device=DEV) if k % 2 == 0 else None
        probs.append((X, dY, dW, db))
        want.append((X.double().t() @ dY.double(), dY.double().sum(0)))
    F_.wgrad_grouped_raw(probs, K, M, False)
    first = [(q[2].clone(), q[3].clone() if q[3] is not None else None) for q in probs]
    for (X, dY, dW, db), (w, b) in zip(probs, want):
        if X.shape[0] == 0:
            assert float(dW.abs().max()) == 0. and (db is None or float(db.abs().max()) == 0.)
            continue
        assert rel_err(dW, w) < TOL
        if db is not None:
            assert rel_err(db, b) < TOL
    F_.wgrad_grouped_raw(probs, K, M, False)
    for (X, dY, dW, db), (w0, b0) in zip(probs, first):
        assert torch.equal(dW, w0) and (db is None or torch.equal(db, b0))
