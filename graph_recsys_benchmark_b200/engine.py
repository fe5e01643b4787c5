"""Fused propagation engine for the standard PEAGCN configuration (every metapath = 2 GCN steps,
emb <= hidden, hidden > repr - all shipped configurations, experiments/scripts/*.ps1).

The per-layer modules (nn/, models/) remain the definition of the arithmetic and the fallback for
any other configuration; this engine runs the SAME kernels with the data flow planned for the
whole model instead of layer by layer, as two autograd nodes instead of ~100:

  head   A1_r = A_hat_r @ x for each DISTINCT first-step relation r (weight-free, shared by the
         metapaths that start with r).  Backward: the transposed aggregations accumulate straight
         into one d_x buffer (no per-relation gradient tensors, no add kernels).
  body   per metapath p:  H1_p = relu(A1_r(p) W1_p + b1_p);  T2_p = H1_p W2_p is written directly
         into its 16-column slot of one [N, P*16] table laid out in last-step-relation order;
         ONE aggregation per distinct last-step relation over its column range (+ b2) into the
         [N, P*16] channel table; the fusion kernel reads that table in place.  Backward mirrors it:
         fusion -> one bias column-sum for all metapaths -> one transposed aggregation per relation
         -> per metapath d W2, dP1 = (dT2 W2^T) gated by H1 > 0 in the projection's epilogue,
         d W1 / d b1, and d A1_r accumulated across the metapaths that share r.
No concatenation, stacking or splitting copies remain, and relu backward never runs as its own pass.
"""
import torch

from . import functional as F_
from .graph import get_graph


class GcnPlan(object):
    """Static schedule of a model: which relation feeds which metapath, and the column layout."""

    def __init__(self, model):
        n = model.x.shape[0]
        eil = model.meta_path_edge_index_list
        self.P = len(model.pea_channels)
        self.first_graphs, self.rel_of_path = [], []
        seen = {}
        for p in range(self.P):
            g = get_graph(eil[p][0], n)
            if id(g) not in seen:
                seen[id(g)] = len(self.first_graphs)
                self.first_graphs.append(g)
            self.rel_of_path.append(seen[id(g)])
        groups, index = [], {}
        for p in range(self.P):
            g = get_graph(eil[p][1], n)
            if id(g) not in index:
                index[id(g)] = len(groups)
                groups.append((g, []))
            groups[index[id(g)]][1].append(p)
        self.groups = groups
        self.order = [p for _, members in groups for p in members]      # column slot -> metapath
        self.slot = {p: s for s, p in enumerate(self.order)}
        self.order_t = torch.tensor(self.order, dtype=torch.long, device=model.x.device)
        first = model.pea_channels[0].gnn_layers
        self.emb, self.hidden, self.repr = first[0].in_channels, first[0].out_channels, first[1].out_channels

    # ---- data movement of the two aggregation phases (overridden by the row-sharded plan) ----------
    def rows(self, x):
        return x.shape[0]

    def head_forward(self, x):
        outs = []
        for g in self.first_graphs:
            dis = g.gcn_dis
            outs.append(F_.spmm_raw(g.fwd, x, x.shape[1], torch.empty_like(x), dis, dis, True))
        return outs

    def head_backward(self, grads):
        dx = None
        for g, d in zip(self.first_graphs, grads):
            if d is None:
                continue
            d = F_._rows(d)
            dis = g.gcn_dis
            if dx is None:
                dx = F_.spmm_raw(g.bwd, d, d.shape[1], torch.empty_like(d), dis, dis, True)
            else:
                F_.spmm_raw(g.bwd, d, d.shape[1], dx, dis, dis, True, accumulate=True)
        return dx

    def last_forward(self, t2, z, bias_all):
        D, start = self.repr, 0
        for g, members in self.groups:
            width = len(members) * D
            dis = g.gcn_dis
            F_.spmm_raw(g.fwd, t2[:, start:start + width], width, z[:, start:start + width], dis, dis, True,
                        bias_all[start:start + width])
            start += width

    def last_backward(self, dz):
        dt2 = torch.empty_like(dz)
        D, start = self.repr, 0
        for g, members in self.groups:
            width = len(members) * D
            dis = g.gcn_dis
            F_.spmm_raw(g.bwd, dz[:, start:start + width], width, dt2[:, start:start + width], dis, dis, True)
            start += width
        return dt2

    @staticmethod
    def applies(model):
        from .models.families import _GCNLayer
        if getattr(model, 'channel_aggr', None) not in ('att', 'mean'):
            return False
        dims = None
        for ch in model.pea_channels:
            if ch.num_steps != 2 or not all(isinstance(l, _GCNLayer) for l in ch.gnn_layers):
                return False
            d = (ch.gnn_layers[0].in_channels, ch.gnn_layers[0].out_channels, ch.gnn_layers[1].in_channels,
                 ch.gnn_layers[1].out_channels)
            if dims is None:
                dims = d
            if d != dims or not (d[0] <= d[1] and d[2] > d[3] and d[1] == d[2]):
                return False
        p, r = len(model.pea_channels), dims[3]
        return p <= 32 and r % 4 == 0 and p * r <= 256 and (r // 4) & (r // 4 - 1) == 0


class _GcnHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan):
        x = F_._rows(F_._req(x, 'x'))
        ctx.plan = plan
        return tuple(plan.head_forward(x))

    @staticmethod
    def backward(ctx, *grads):
        return ctx.plan.head_backward(grads), None


class _GcnBody(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, att, mode, skip, n_rel, *tensors):
        A1 = [F_._rows(t) for t in tensors[:n_rel]]
        params = tensors[n_rel:]
        P, D, H = plan.P, plan.repr, plan.hidden
        W1 = [params[4 * p].contiguous() for p in range(P)]
        b1 = [params[4 * p + 1].contiguous() for p in range(P)]
        W2 = [params[4 * p + 2].contiguous() for p in range(P)]
        b2 = [params[4 * p + 3] for p in range(P)]
        dev = A1[0].device
        n = A1[0].shape[0]
        wide = P * D
        t2 = torch.empty(n, wide, dtype=torch.float32, device=dev)
        h1 = []
        for p in range(P):
            h = torch.empty(n, H, dtype=torch.float32, device=dev)
            F_.linear_raw(A1[plan.rel_of_path[p]], W1[p], h, False, b1[p], True)
            s = plan.slot[p]
            F_.linear_raw(h, W2[p], t2[:, s * D:(s + 1) * D], False)
            h1.append(h)
        z = torch.empty(n, wide, dtype=torch.float32, device=dev)
        bias_all = torch.cat([b2[p] for p in plan.order])
        plan.last_forward(t2, z, bias_all)
        del t2
        att_perm = att.reshape(P, D).index_select(0, plan.order_t).contiguous() if att is not None else None
        out = torch.empty(n, D, dtype=torch.float32, device=dev)
        skip_slot = plan.slot[skip] if skip >= 0 else -1
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_forward', F_._ptr(z), wide, n, P, D, F_._ptr(att_perm), mode, skip_slot,
                         F_._ptr(out), D, F_._stream())
        ctx.plan, ctx.mode, ctx.skip, ctx.n_rel = plan, mode, skip, n_rel
        ctx.att_shape = att.shape if att is not None else None
        ctx.save_for_backward(z, att_perm, *A1, *h1, *W1, *W2)
        return out

    @staticmethod
    def backward(ctx, dout):
        plan, n_rel = ctx.plan, ctx.n_rel
        if ctx.skip >= 0:
            raise RuntimeError('metapath ablation (metapath_idx) is an evaluation-only path (models/base.py:88-96)')
        P, D, H, E = plan.P, plan.repr, plan.hidden, plan.emb
        saved = ctx.saved_tensors
        z, att_perm = saved[0], saved[1]
        A1 = saved[2:2 + n_rel]
        h1 = saved[2 + n_rel:2 + n_rel + P]
        W1 = saved[2 + n_rel + P:2 + n_rel + 2 * P]
        W2 = saved[2 + n_rel + 2 * P:2 + n_rel + 3 * P]
        dev = z.device
        n = z.shape[0]
        wide = P * D
        dout = F_._rows(dout)
        dz = torch.empty(n, wide, dtype=torch.float32, device=dev)
        d_att_perm = torch.empty(P, D, dtype=torch.float32, device=dev) if ctx.mode == 0 else None
        need = int(F_._lib.query('peagnn_fuse_workspace_floats', n, P, D)) if ctx.mode == 0 else 0
        ws = F_._ws(need, dev) if ctx.mode == 0 else None
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_backward', F_._ptr(z), wide, n, P, D, F_._ptr(att_perm), ctx.mode, F_._ptr(dout),
                         dout.stride(0), F_._ptr(dz), wide, F_._ptr(d_att_perm), F_._ptr(ws), need, F_._stream())
        db2_all = torch.empty(wide, dtype=torch.float32, device=dev)
        F_.wgrad_raw(None, dz, 0, wide, 0, None, db2_all)                     # every metapath's d b2 at once
        dt2 = plan.last_backward(dz)
        dA1 = [None] * n_rel
        grads = []
        dp1 = torch.empty(n, H, dtype=torch.float32, device=dev)
        for p in range(P):
            s = plan.slot[p]
            d_t2 = dt2[:, s * D:(s + 1) * D]
            dW2 = torch.empty_like(W2[p])
            F_.wgrad_raw(h1[p], d_t2, H, D, False, dW2, None)
            F_.linear_raw(d_t2, W2[p], dp1, True, out_mask=h1[p])             # (dT2 W2^T) gated by relu
            r = plan.rel_of_path[p]
            dW1 = torch.empty_like(W1[p])
            db1 = torch.empty(H, dtype=torch.float32, device=dev)
            F_.wgrad_raw(A1[r], dp1, E, H, False, dW1, db1)
            if dA1[r] is None:
                dA1[r] = torch.empty(n, E, dtype=torch.float32, device=dev)
                F_.linear_raw(dp1, W1[p], dA1[r], True)
            else:
                F_.linear_raw(dp1, W1[p], dA1[r], True, accumulate=True)
            grads.extend([dW1, db1, dW2, db2_all[s * D:(s + 1) * D]])
        d_att = None
        if d_att_perm is not None:
            d_att = torch.empty_like(d_att_perm)
            d_att.index_copy_(0, plan.order_t, d_att_perm)
            d_att = d_att.reshape(ctx.att_shape)
        return (None, d_att, None, None, None) + tuple(dA1) + tuple(grads)


def gcn_forward(model, metapath_idx=None, plan=None):
    """model.forward() through the fused engine (same result as the per-layer path).  With a
    row-sharded ``plan`` (distributed.ShardedGcnPlan) the result is this rank's rows."""
    if plan is None:
        plan = getattr(model, '_gcn_plan', None)
        if plan is None:
            plan = model._gcn_plan = GcnPlan(model)
    a1 = _GcnHead.apply(model.x, plan)
    params = []
    for ch in model.pea_channels:
        l0, l1 = ch.gnn_layers
        params.extend([l0.weight, l0.bias, l1.weight, l1.bias])
    att = model.att if model.channel_aggr == 'att' else None
    mode = 0 if model.channel_aggr == 'att' else 1
    skip = -1 if metapath_idx is None else int(metapath_idx)
    return _GcnBody.apply(plan, att, mode, skip, len(a1), *a1, *params)
