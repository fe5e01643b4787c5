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


_side_streams = {}
PARALLEL_BRANCHES = True      # False: _Fork runs its branches in order on the current stream (per-kernel timing builds)


GROUPED_PROJECTIONS = __import__('os').environ.get('PEAGNN_GROUPED', '1') != '0'   # demand-driven PEAGCN step: one launch per
                                                                                   # projection shape across the metapaths

SAGE_LEAN = __import__('os').environ.get('PEAGNN_SAGE_LEAN', '1') != '0'    # demand-driven PEASage step: range / list projection passes

N_BRANCHES = int(__import__('os').environ.get('PEAGNN_BRANCHES', '4'))     # side streams per device


def fork_streams(device, n_streams=None):
    n_streams = n_streams or N_BRANCHES
    key = (device.index, n_streams)
    if key not in _side_streams:
        _side_streams[key] = [torch.cuda.Stream(device=device) for _ in range(n_streams)]
    return _side_streams[key]


class _Fork(object):
    """Independent kernel chains on side streams: fork from the current stream on entry, join back on exit.
    The 13 per-metapath projection chains of a demand-driven step are many small launches (a few hundred rows
    to 62 k rows each); issued back to back on one stream every one of them pays its own ramp-up and tail, on
    parallel branches they fill the GPU together.  Inside a CUDA-graph capture the branches become parallel
    paths of the graph.  Buffers that outlive the block are allocated by the caller BEFORE entering it."""

    def __init__(self, device, n_streams=None):
        self.streams = fork_streams(device, n_streams)
        self.device = device

    def __enter__(self):
        self.parallel = PARALLEL_BRANCHES
        if self.parallel:
            self.main = torch.cuda.current_stream(self.device)
            ev = self.main.record_event()
            for st in self.streams:
                st.wait_event(ev)
        return self

    def on(self, k):
        if not self.parallel:
            import contextlib
            return contextlib.nullcontext()
        return torch.cuda.stream(self.streams[k % len(self.streams)])

    def record(self):
        """An event on the current branch (None when the branches run in order on one stream)."""
        if not self.parallel:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return ev

    def wait(self, ev):
        """The current branch continues after ``ev`` (a cross-branch dependency)."""
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def __exit__(self, *exc):
        if self.parallel:
            for st in self.streams:
                self.main.wait_stream(st)
        return False


ActiveSet = F_.ActiveRows      # bitmap + sorted id list + first-occurrence mask of the rows a loss() call reads


class GcnPlan(object):
    """Static schedule of a model: which relation feeds which metapath, and the column layout.
    ``kind`` 'gcn' (normalised sum with self loops) or 'sage' (mean over in-edges, no self loops)."""
    lean_projections = True      # demand-driven loss(): projections on the last step's source range + the active rows
    head_params = None           # set per call by gcn_forward: [(W1_p, b1_p)] when the first projections are fused into the head
    head_h1 = None               # ... and the activations that epilogue wrote, per metapath [N, hidden]
    gather_bf16 = False          # opt-in (model.gather_dtype = 'bf16'): first-step tables are gathered from a bf16 copy

    def __init__(self, model, kind='gcn'):
        n = model.x.shape[0]
        self.kind, self.num_nodes = kind, n
        eil = model.meta_path_edge_index_list
        self.P = len(model.pea_channels)
        self.first_graphs, self.rel_of_path = [], []
        seen = {}
        keep = kind == 'sage'
        for p in range(self.P):
            g = get_graph(eil[p][0], n, keep_self_loops=keep)
            if id(g) not in seen:
                seen[id(g)] = len(self.first_graphs)
                self.first_graphs.append(g)
            self.rel_of_path.append(seen[id(g)])
        groups, index = [], {}
        for p in range(self.P):
            g = get_graph(eil[p][1], n, keep_self_loops=keep)
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
        # the shapes the grouped projection kernels exist for (every shipped configuration); any other width keeps the
        # single launches on parallel branches
        self.grouped_shapes = (self.emb, self.hidden, self.repr) == (64, 64, 16)
        self._src_range = None

    def source_range_tensors(self, dev):
        if getattr(self, '_range_t', None) is None:
            ranges = self.source_ranges()
            self._range_t = (torch.tensor([ranges[p][0] for p in self.order], dtype=torch.long, device=dev),
                             torch.tensor([ranges[p][1] for p in self.order], dtype=torch.long, device=dev))
        return self._range_t

    # ---- data movement of the two aggregation phases (overridden by the row-sharded plan) ----------
    def _scales(self, g, transposed):
        """(row scale, column scale, implicit self loop) of the aggregation operator or its transpose."""
        if self.kind == 'gcn':
            return g.gcn_dis, g.gcn_dis, True
        return (None, g.inv_in_degree, False) if transposed else (g.inv_in_degree, None, False)

    def head_row_bitmaps(self, active):
        """Per first-step relation: the rows of its aggregate a demand-driven step reads - the source ranges of the
        metapaths that start with it (static) plus the active rows.  [n_rel, words] int32, one OR per step."""
        if getattr(self, '_static_rows', None) is None:
            ranges = self.source_ranges()
            words = active.bitmap.numel()
            dev = active.bitmap.device
            bit = torch.arange(words * 32, device=dev)
            rows = []
            for r in range(len(self.first_graphs)):
                on = torch.zeros(words * 32, dtype=torch.bool, device=dev)
                for p in range(self.P):
                    if self.rel_of_path[p] == r:
                        lo, hi = ranges[p]
                        on |= (bit >= lo) & (bit < hi)
                w = (on.view(words, 32).to(torch.int64) << torch.arange(32, device=dev)).sum(dim=1)
                rows.append(torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32))
            self._static_rows = torch.stack(rows).contiguous()
        return torch.bitwise_or(self._static_rows, active.bitmap[None, :])

    def head_forward(self, x, row_bitmaps=None):
        """One aggregation of the embedding table per distinct first-step relation; the launches are independent
        (most are small, mostly-empty relations bound by their row epilogues), so they run on parallel branches.
        ``row_bitmaps`` [n_rel, words]: only the marked rows of each aggregate are computed (the rest stay unwritten)."""
        outs = [torch.empty_like(x) for _ in self.first_graphs]
        bm = (lambda k: row_bitmaps[k]) if row_bitmaps is not None else (lambda k: None)
        bf16 = self.gather_bf16 and x.shape[1] == 64
        # north_star (2): the first projection of every metapath rides in the epilogue of its relation's aggregation
        # (peagnn_spmm_proj) - demand-driven GCN steps with 64 -> 64 first layers; the activations land in head_h1
        fused = self.head_params is not None and not bf16 and x.shape[1] == 64 and self.hidden == 64 and self.kind == 'gcn'
        self.head_h1 = None
        if fused:
            by_rel = [[p for p in range(self.P) if self.rel_of_path[p] == k] for k in range(len(self.first_graphs))]
            fused = all(1 <= len(ps) <= 2 for ps in by_rel)
        if fused:
            self.head_h1 = [torch.empty(x.shape[0], self.hidden, dtype=torch.float32, device=x.device) for _ in range(self.P)]
            proj = [[(self.head_params[p][0], self.head_params[p][1], self.head_h1[p]) for p in ps] for ps in by_rel]

        if fused:
            scales = [self._scales(g, False) for g in self.first_graphs]
            for g in self.first_graphs:
                g.fwd.view(64)
            parallel = x.shape[0] >= 50000
            with _Fork(x.device) as fork:
                for k, (g, out) in enumerate(zip(self.first_graphs, outs)):
                    rs, cs, loop = scales[k]
                    if parallel:
                        with fork.on(k):
                            F_.spmm_proj_raw(g.fwd, x, out, rs, cs, loop, proj[k], True, bm(k))
                    else:
                        F_.spmm_proj_raw(g.fwd, x, out, rs, cs, loop, proj[k], True, bm(k))
            return outs
        table = F_.to_bf16(x) if bf16 else x                     # opt-in: the gathered copy of the embedding table in bf16
        agg = F_.spmm_bf16_raw if bf16 else F_.spmm_raw
        if x.shape[0] < 50000:                                   # tiny graphs: not worth the fork / join
            for k, (g, out) in enumerate(zip(self.first_graphs, outs)):
                rs, cs, loop = self._scales(g, False)
                agg(g.fwd, table, x.shape[1], out, rs, cs, loop, active_rows=bm(k))
            return outs
        # everything that is built lazily (degree scalings, chunk-partial workspaces) has to exist BEFORE the fork: a
        # kernel launched on this stream after the fork point is not ordered before the branches
        scales = [self._scales(g, False) for g in self.first_graphs]
        for g in self.first_graphs:
            g.fwd.view(x.shape[1])
        with _Fork(x.device) as fork:
            for k, (g, out) in enumerate(zip(self.first_graphs, outs)):
                rs, cs, loop = scales[k]
                with fork.on(k):
                    agg(g.fwd, table, x.shape[1], out, rs, cs, loop, active_rows=bm(k))
        return outs

    def head_backward_rows(self, row_bitmaps):
        """Rows a transposed first-step aggregation of a demand-driven step can write anything but zero to: the sources
        of its relation (rows with an edge in the transposed structure) and - through the self loop - the rows its
        gradient table is non-zero on (``row_bitmaps``, the rows the forward computed).  [n_rel, words]."""
        if getattr(self, '_bwd_sources', None) is None:
            self._bwd_sources = torch.stack([g.bwd.nonempty_row_bitmap() for g in self.first_graphs]).contiguous()
        return torch.bitwise_or(self._bwd_sources, row_bitmaps)

    def head_backward(self, grads, row_bitmaps=None):
        """d x = sum over the first-step relations of A_hat_r^T d A1_r.  The transposed aggregations are independent;
        each branch accumulates its relations into its own buffer (fixed assignment and order: deterministic) and the
        few partial tables are added at the end.  ``row_bitmaps`` (demand-driven steps): every launch but the first of
        a buffer visits only the rows it can contribute to (head_backward_rows) instead of re-reading and re-writing
        the whole [N, emb] partial table for a relation with a handful of sources."""
        todo = [(g, F_._rows(d), k) for k, (g, d) in enumerate(zip(self.first_graphs, grads)) if d is not None]
        if not todo:
            return None
        rows = self.head_backward_rows(row_bitmaps) if row_bitmaps is not None and hasattr(todo[0][0].bwd, 'nonempty_row_bitmap') else None
        only = (lambda k, accumulate: rows[k] if (rows is not None and accumulate) else None)
        n_br = 1 if todo[0][1].shape[0] < 50000 else min(len(todo), len(fork_streams(todo[0][1].device)))
        parts = [torch.empty_like(todo[0][1]) for _ in range(n_br)]
        # opt-in bf16 gathers: worth the conversion pass only where the gather dominates (the relations with many edges)
        use_bf16 = [self.gather_bf16 and d.shape[1] == 64 and g.bwd.nnz >= 8 * d.shape[0] for g, d, _ in todo]
        tables = [F_.to_bf16(d) if b else d for (g, d, _), b in zip(todo, use_bf16)]
        aggs = [F_.spmm_bf16_raw if b else F_.spmm_raw for b in use_bf16]
        if n_br == 1:
            for j, (g, d, k) in enumerate(todo):
                rs, cs, loop = self._scales(g, True)
                aggs[j](g.bwd, tables[j], d.shape[1], parts[0], rs, cs, loop, accumulate=j > 0, active_rows=only(k, j > 0))
            return parts[0]
        # the two largest relations go to different branches; the rest are dealt round-robin
        order = sorted(range(len(todo)), key=lambda j: -todo[j][0].bwd.nnz)
        scales = [self._scales(g, True) for g, _, _ in todo]      # lazily built tensors and workspaces: before the fork
        for g, d, _ in todo:
            g.bwd.view(d.shape[1])
        seen = [False] * n_br
        with _Fork(parts[0].device) as fork:
            for pos, j in enumerate(order):
                g, d, k = todo[j]
                b = pos % n_br
                rs, cs, loop = scales[j]
                with fork.on(b):
                    aggs[j](g.bwd, tables[j], d.shape[1], parts[b], rs, cs, loop, accumulate=seen[b], active_rows=only(k, seen[b]))
                seen[b] = True
        dx = parts[0]
        for extra in parts[1:]:
            dx.add_(extra)
        return dx

    def active_bitmap(self, ids):
        """Rows of the final representation a loss on the node ids ``ids`` reads (models/base.py:209-210)."""
        return F_.active_rows(ids, self.num_nodes)

    def source_ranges(self):
        """Per metapath: [lo, hi) node-id range holding every SOURCE of its last-step relation (node ids are
        contiguous by type upstream, datasets/movielens.py:184-227, so this is the source type's range; any
        superset is still exact).  The last step reads the projected table only there and on the active rows."""
        if self._src_range is None:
            per_group = []
            for g, _ in self.groups:
                col = g.fwd.col
                per_group.append((int(col.min().item()), int(col.max().item()) + 1) if col.numel() else (0, 0))
            self._src_range = {}
            for (g, members), rng in zip(self.groups, per_group):
                for p in members:
                    self._src_range[p] = rng
        return self._src_range

    def last_forward(self, t2, z, bias_all, active=None):
        """``active``: only the marked destination rows of the last step are aggregated (the rest of z is not written).
        The relation groups own disjoint column ranges: parallel branches."""
        D, start = self.repr, 0
        jobs = []
        for g, members in self.groups:
            width = len(members) * D
            g.fwd.view(width)                                    # structures / workspaces before the fork
            jobs.append((g, start, width, self._scales(g, False)))
            start += width
        with _Fork(t2.device) as fork:
            for k, (g, start, width, (rs, cs, loop)) in enumerate(jobs):
                with fork.on(k):
                    F_.spmm_raw(g.fwd, t2[:, start:start + width], width, z[:, start:start + width], rs, cs, loop,
                                bias_all[start:start + width], active_rows=active.bitmap if active is not None else None)

    def last_backward(self, dz, active=None):
        """``active``: dz is zero outside the marked rows, so the transposed pass skips every other edge.
        (The implicit self-loop term of row i reads dz[i], which is zero wherever i is not marked: exact.)"""
        dt2 = torch.empty_like(dz)
        D, start = self.repr, 0
        jobs = []
        for g, members in self.groups:
            width = len(members) * D
            g.bwd.view(width)
            jobs.append((g, start, width, self._scales(g, True)))
            start += width
        with _Fork(dz.device) as fork:
            for k, (g, start, width, (rs, cs, loop)) in enumerate(jobs):
                with fork.on(k):
                    F_.spmm_raw(g.bwd, dz[:, start:start + width], width, dt2[:, start:start + width], rs, cs, loop,
                                active_cols=active.bitmap if active is not None else None)
        return dt2

    @staticmethod
    def applies(model, kind='gcn'):
        from .models.families import _GCNLayer, _SageLayer
        layer_class = _GCNLayer if kind == 'gcn' else _SageLayer
        if getattr(model, 'channel_aggr', None) not in ('att', 'mean'):
            return False
        dims = None
        for ch in model.pea_channels:
            if ch.num_steps != 2 or not all(isinstance(l, layer_class) for l in ch.gnn_layers):
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
    def forward(ctx, x, plan, row_bitmaps=None):
        x = F_._rows(F_._req(x, 'x'))
        ctx.plan, ctx.row_bitmaps = plan, row_bitmaps
        if row_bitmaps is not None:
            return tuple(plan.head_forward(x, row_bitmaps))
        return tuple(plan.head_forward(x))

    @staticmethod
    def backward(ctx, *grads):
        if ctx.row_bitmaps is not None:
            return ctx.plan.head_backward(grads, ctx.row_bitmaps), None, None
        return ctx.plan.head_backward(grads), None, None


class _GcnBody(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, att, mode, skip, n_rel, active, *tensors):
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
        h1 = [torch.empty(n, H, dtype=torch.float32, device=dev) for _ in range(P)]
        if GROUPED_PROJECTIONS and plan.grouped_shapes:                                # one launch per projection shape across the metapaths
            rel = plan.rel_of_path
            F_.linear_grouped_raw([(A1[rel[p]], W1[p], b1[p], h1[p], None) for p in range(P)], plan.emb, H, False, relu=True)
            F_.linear_grouped_raw([(h1[p], W2[p], None, t2[:, plan.slot[p] * D:(plan.slot[p] + 1) * D], None) for p in range(P)],
                                  H, D, False)
        else:
          with _Fork(dev) as fork:                             # the per-metapath chains are independent: parallel branches
            for p in range(P):
                r, s = plan.rel_of_path[p], plan.slot[p]
                with fork.on(r):
                    F_.linear_raw(A1[r], W1[p], h1[p], False, b1[p], True)
                    F_.linear_raw(h1[p], W2[p], t2[:, s * D:(s + 1) * D], False)
        z = (torch.zeros if active is not None else torch.empty)(n, wide, dtype=torch.float32, device=dev)
        bias_all = torch.cat([b2[p] for p in plan.order])
        plan.last_forward(t2, z, bias_all, active)
        del t2
        att_perm = att.reshape(P, D).index_select(0, plan.order_t).contiguous() if att is not None else None
        out = torch.empty(n, D, dtype=torch.float32, device=dev)
        skip_slot = plan.slot[skip] if skip >= 0 else -1
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_forward', F_._ptr(z), wide, n, P, D, F_._ptr(att_perm), mode, skip_slot,
                         F_._ptr(out), D, F_._stream())
        ctx.plan, ctx.mode, ctx.skip, ctx.n_rel, ctx.active = plan, mode, skip, n_rel, active
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
        dt2 = plan.last_backward(dz, ctx.active)
        dA1 = [None] * n_rel
        new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        dW2, dW1, db1 = [new(H, D) for _ in range(P)], [new(E, H) for _ in range(P)], [new(H) for _ in range(P)]
        first_of = {}
        for p in range(P):
            r = plan.rel_of_path[p]
            if dA1[r] is None:
                dA1[r] = new(n, E)
                first_of[r] = p
        n_branches = len(fork_streams(dev))
        if GROUPED_PROJECTIONS and plan.grouped_shapes:
            rel = plan.rel_of_path
            dslot = lambda p: dt2[:, plan.slot[p] * D:(plan.slot[p] + 1) * D]
            dp1 = [new(n, H) for _ in range(P)]
            rounds, seen = [], {}
            for p in range(P):                                 # metapaths that share d A1[r]: successive launches, fixed order
                k = seen.get(rel[p], 0)
                seen[rel[p]] = k + 1
                while len(rounds) <= k:
                    rounds.append([])
                rounds[k].append(p)
            with _Fork(dev) as fork:
                with fork.on(0):                               # d P1 = (d T2 W2^T) gated by relu -> d A1
                    F_.linear_grouped_raw([(dslot(p), W2[p], None, dp1[p], h1[p]) for p in range(P)], D, H, True)
                    have_dp1 = fork.record()
                    for k, members in enumerate(rounds):
                        F_.linear_grouped_raw([(dp1[p], W1[p], None, dA1[rel[p]], None) for p in members], H, E, True,
                                              accumulate=k > 0)
                with fork.on(1):                               # d W2 needs nothing of the above; d W1, d b1 need d P1
                    F_.wgrad_grouped_raw([(h1[p], dslot(p), dW2[p], None) for p in range(P)], H, D, False)
                    fork.wait(have_dp1)
                    F_.wgrad_grouped_raw([(A1[rel[p]], dp1[p], dW1[p], db1[p]) for p in range(P)], E, H, False)
            del dp1
        else:
          dp1 = [new(n, H) for _ in range(min(P, n_branches))]   # one gated-gradient buffer per branch
          with _Fork(dev) as fork:
            for p in range(P):
                r, s = plan.rel_of_path[p], plan.slot[p]
                with fork.on(r):                               # metapaths sharing dA1[r] stay on one branch, in order
                    buf = dp1[r % len(dp1)]
                    d_t2 = dt2[:, s * D:(s + 1) * D]
                    F_.wgrad_raw(h1[p], d_t2, H, D, False, dW2[p], None)
                    F_.linear_raw(d_t2, W2[p], buf, True, out_mask=h1[p])      # (dT2 W2^T) gated by relu
                    F_.wgrad_raw(A1[r], buf, E, H, False, dW1[p], db1[p])
                    F_.linear_raw(buf, W1[p], dA1[r], True, accumulate=first_of[r] != p)
        grads = []
        for p in range(P):
            s = plan.slot[p]
            grads.extend([dW1[p], db1[p], dW2[p], db2_all[s * D:(s + 1) * D]])
        d_att = None
        if d_att_perm is not None:
            d_att = torch.empty_like(d_att_perm)
            d_att.index_copy_(0, plan.order_t, d_att_perm)
            d_att = d_att.reshape(ctx.att_shape)
        return (None, d_att, None, None, None, None) + tuple(dA1) + tuple(grads)


class _GcnBodyLean(torch.autograd.Function):
    """_GcnBody for a demand-driven loss(): besides aggregating the last step on the active rows only, the
    per-metapath projections run only where that step reads their result -
      * the RANGE pass: the node-id range of the last relation's sources (plan.source_ranges(): e.g. the 62 k item
        rows for the nine ML-25M metapaths that end with item -> user, instead of all 291 k rows), and
      * the LIST pass: the active rows themselves (the self-loop term), gathered into a [3B, .] buffer.
    Rows of the range pass that are also on the list are computed twice with the same result; the backward counts each
    row once (``first`` occurrence, outside the range).  Everything is existing kernels on row slices / gathered
    rows, fixed shapes, no host sync - the step stays graph-capturable.  Every value that is computed, the loss and
    all gradients equal _GcnBody's (tests/test_gpu_model.py::test_demand_driven_loss_equals_full_propagation)."""

    @staticmethod
    def forward(ctx, plan, att, mode, n_rel, active, *tensors):
        A1 = [F_._rows(t) for t in tensors[:n_rel]]
        params = tensors[n_rel:]
        P, D, H = plan.P, plan.repr, plan.hidden
        W1 = [params[4 * p].contiguous() for p in range(P)]
        b1 = [params[4 * p + 1].contiguous() for p in range(P)]
        W2 = [params[4 * p + 2].contiguous() for p in range(P)]
        b2 = [params[4 * p + 3] for p in range(P)]
        dev, n, wide = A1[0].device, A1[0].shape[0], P * D
        ranges = plan.source_ranges()
        ids, nl = active.ids, int(active.ids.numel())
        t2 = torch.empty(n, wide, dtype=torch.float32, device=dev)
        a1c = [None] * n_rel                                   # the active rows of each first-step aggregate
        for p in range(P):
            r = plan.rel_of_path[p]
            if a1c[r] is None:
                a1c[r] = A1[r].index_select(0, ids)
        pre = active.h1_full                                   # first-layer activations from the head's fused epilogue
        if pre is not None:
            h1r = [pre[p][ranges[p][0]:ranges[p][1]] for p in range(P)]
            h1c = [pre[p].index_select(0, ids) for p in range(P)]
        else:
            h1r = [torch.empty(max(ranges[p][1] - ranges[p][0], 0), H, dtype=torch.float32, device=dev) for p in range(P)]
            h1c = [torch.empty(nl, H, dtype=torch.float32, device=dev) for p in range(P)]
        t2c = torch.empty(nl, wide, dtype=torch.float32, device=dev)
        if GROUPED_PROJECTIONS and plan.grouped_shapes:
            # every metapath's range and list problems of one shape as ONE launch (peagnn_linear_grouped): 26 problems of
            # 12 k - 160 k rows pay one prologue / tail together instead of 26, on one stream
            slot = lambda t, p, a, b: t[a:b, plan.slot[p] * D:(plan.slot[p] + 1) * D]
            rel = plan.rel_of_path
            has = [p for p in range(P) if ranges[p][1] > ranges[p][0]]
            with _Fork(dev) as fork:                           # the range chain and the list chain are independent
                with fork.on(0):
                    if pre is None:
                        F_.linear_grouped_raw([(A1[rel[p]][ranges[p][0]:ranges[p][1]], W1[p], b1[p], h1r[p], None) for p in has],
                                              plan.emb, H, False, relu=True)
                    F_.linear_grouped_raw([(h1r[p], W2[p], None, slot(t2, p, ranges[p][0], ranges[p][1]), None) for p in has],
                                          H, D, False)
                with fork.on(1):
                    if pre is None:
                        F_.linear_grouped_raw([(a1c[rel[p]], W1[p], b1[p], h1c[p], None) for p in range(P)], plan.emb, H, False,
                                              relu=True)
                    F_.linear_grouped_raw([(h1c[p], W2[p], None, slot(t2c, p, 0, nl), None) for p in range(P)], H, D, False)
        else:
          with _Fork(dev) as fork:
            for p in range(P):
                r, s = plan.rel_of_path[p], plan.slot[p]
                lo, hi = ranges[p]
                with fork.on(r):                               # metapaths sharing a first relation share a branch
                    if hi > lo:
                        if pre is None:
                            F_.linear_raw(A1[r][lo:hi], W1[p], h1r[p], False, b1[p], True)
                        F_.linear_raw(h1r[p], W2[p], t2[lo:hi, s * D:(s + 1) * D], False)
                    if pre is None:
                        F_.linear_raw(a1c[r], W1[p], h1c[p], False, b1[p], True)
                    F_.linear_raw(h1c[p], W2[p], t2c[:, s * D:(s + 1) * D], False)
        t2.index_copy_(0, ids, t2c)                            # duplicates / range rows rewrite identical values
        del t2c
        z = torch.empty(n, wide, dtype=torch.float32, device=dev)     # only the active rows are written, and only they are read
        bias_all = torch.cat([b2[p] for p in plan.order])
        plan.last_forward(t2, z, bias_all, active)
        del t2
        att_perm = att.reshape(P, D).index_select(0, plan.order_t).contiguous() if att is not None else None
        # fusion on the active rows only: gathered into [3B, P*D], fused, scattered back (duplicates rewrite the same row)
        z = z.index_select(0, ids)
        out_c = torch.empty(nl, D, dtype=torch.float32, device=dev)
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_forward', F_._ptr(z), wide, nl, P, D, F_._ptr(att_perm), mode, -1,
                         F_._ptr(out_c), D, F_._stream())
        out = torch.zeros(n, D, dtype=torch.float32, device=dev)
        out.index_copy_(0, ids, out_c)
        ctx.n_nodes = n
        ctx.plan, ctx.mode, ctx.n_rel, ctx.active = plan, mode, n_rel, active
        ctx.att_shape = att.shape if att is not None else None
        used = [k for k in range(n_rel) if a1c[k] is not None]
        ctx.used = used
        ctx.save_for_backward(z, att_perm, *A1, *[a1c[k] for k in used], *h1r, *h1c, *W1, *W2)
        return out

    @staticmethod
    def backward(ctx, dout):
        plan, n_rel, active = ctx.plan, ctx.n_rel, ctx.active
        P, D, H, E = plan.P, plan.repr, plan.hidden, plan.emb
        saved = list(ctx.saved_tensors)
        z, att_perm = saved[0], saved[1]
        pos = 2
        A1 = saved[pos:pos + n_rel]; pos += n_rel
        a1c = {k: t for k, t in zip(ctx.used, saved[pos:pos + len(ctx.used)])}; pos += len(ctx.used)
        h1r = saved[pos:pos + P]; pos += P
        h1c = saved[pos:pos + P]; pos += P
        W1 = saved[pos:pos + P]; pos += P
        W2 = saved[pos:pos + P]
        dev, n, wide = z.device, ctx.n_nodes, P * D
        ranges = plan.source_ranges()
        ids, first = active.ids, active.first
        nl = int(ids.numel())
        # z holds the gathered active rows; each node's gradient enters once, at its first occurrence on the list
        dout_c = dout.index_select(0, ids) * first[:, None].to(torch.float32)
        dz_c = torch.empty(nl, wide, dtype=torch.float32, device=dev)
        d_att_perm = torch.empty(P, D, dtype=torch.float32, device=dev) if ctx.mode == 0 else None
        need = int(F_._lib.query('peagnn_fuse_workspace_floats', nl, P, D)) if ctx.mode == 0 else 0
        ws = F_._ws(need, dev) if ctx.mode == 0 else None
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_backward', F_._ptr(z), wide, nl, P, D, F_._ptr(att_perm), ctx.mode, F_._ptr(dout_c),
                         dout_c.stride(0), F_._ptr(dz_c), wide, F_._ptr(d_att_perm), F_._ptr(ws), need, F_._stream())
        db2_all = torch.empty(wide, dtype=torch.float32, device=dev)
        F_.wgrad_raw(None, dz_c, 0, wide, 0, None, db2_all)
        dz = torch.zeros(n, wide, dtype=torch.float32, device=dev)
        dz.index_add_(0, ids, dz_c)                            # later occurrences add exact zeros: order-independent
        dt2 = plan.last_backward(dz, active)
        # per slot: the active rows' gradients, each node once (first occurrence) and only where the range pass of that
        # metapath does not already cover it
        lo_s, hi_s = plan.source_range_tensors(dev)             # per slot, built once (no host copy inside a capture)
        keep = first[:, None] & ((ids[:, None] < lo_s[None, :]) | (ids[:, None] >= hi_s[None, :]))      # [3B, P]
        dt2c = dt2.index_select(0, ids).view(nl, P, D) * keep[:, :, None].to(torch.float32)          # [3B, P, D]
        dA1 = [None] * n_rel
        for p in range(P):
            r = plan.rel_of_path[p]
            if dA1[r] is None:
                dA1[r] = torch.zeros(n, E, dtype=torch.float32, device=dev)
        new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        dW2a, dW1a, db1a = [new(H, D) for _ in range(P)], [new(E, H) for _ in range(P)], [new(H) for _ in range(P)]
        dW2b, dW1b, db1b = [new(H, D) for _ in range(P)], [new(E, H) for _ in range(P)], [new(H) for _ in range(P)]
        dp1r = [new(max(ranges[p][1] - ranges[p][0], 0), H) for p in range(P)]
        dp1c, dac = [new(nl, H) for _ in range(P)], [new(nl, E) for _ in range(P)]
        if GROUPED_PROJECTIONS and plan.grouped_shapes:
            rel = plan.rel_of_path
            has = [p for p in range(P) if ranges[p][1] > ranges[p][0]]
            dslot = lambda p: dt2[ranges[p][0]:ranges[p][1], plan.slot[p] * D:(plan.slot[p] + 1) * D]
            a1r = lambda p: A1[rel[p]][ranges[p][0]:ranges[p][1]]
            rounds, seen = [], {}
            for p in has:                                      # metapaths that share d A1[r]: successive launches, fixed order
                k = seen.get(rel[p], 0)
                seen[rel[p]] = k + 1
                while len(rounds) <= k:
                    rounds.append([])
                rounds[k].append(p)
            tc = lambda p: dt2c[:, plan.slot[p], :]
            with _Fork(dev) as fork:                           # four independent chains of grouped launches
                with fork.on(0):                               # range: d P1 = (d T2 W2^T) gated by relu -> d A1
                    F_.linear_grouped_raw([(dslot(p), W2[p], None, dp1r[p], h1r[p]) for p in has], D, H, True)
                    have_dp1r = fork.record()
                    for members in rounds:
                        F_.linear_grouped_raw([(dp1r[p], W1[p], None, dA1[rel[p]][ranges[p][0]:ranges[p][1]], None)
                                               for p in members], H, E, True, accumulate=True)
                with fork.on(1):                               # range: d W2 (a metapath without a range pass gets zeros),
                    F_.wgrad_grouped_raw([(h1r[p], dslot(p), dW2a[p], None) for p in range(P)], H, D, False)
                    fork.wait(have_dp1r)                       # ... then d W1, d b1 next to the d A1 launches
                    F_.wgrad_grouped_raw([(a1r(p), dp1r[p], dW1a[p], db1a[p]) for p in range(P)], E, H, False)
                with fork.on(2):                               # list: the same on the gathered batch rows
                    F_.linear_grouped_raw([(tc(p), W2[p], None, dp1c[p], h1c[p]) for p in range(P)], D, H, True)
                    have_dp1c = fork.record()
                    F_.linear_grouped_raw([(dp1c[p], W1[p], None, dac[p], None) for p in range(P)], H, E, True)
                with fork.on(3):
                    F_.wgrad_grouped_raw([(h1c[p], tc(p), dW2b[p], None) for p in range(P)], H, D, False)
                    fork.wait(have_dp1c)
                    F_.wgrad_grouped_raw([(a1c[rel[p]], dp1c[p], dW1b[p], db1b[p]) for p in range(P)], E, H, False)
            for p in range(P):
                dA1[rel[p]].index_add_(0, ids, dac[p])         # masked-out entries add exact zeros: order-independent
        else:
          with _Fork(dev) as fork:
            for p in range(P):
                r, s = plan.rel_of_path[p], plan.slot[p]
                lo, hi = ranges[p]
                with fork.on(r):                               # same branch for metapaths that accumulate into one dA1[r]
                    if hi > lo:                                # ---- range pass
                        d_t2 = dt2[lo:hi, s * D:(s + 1) * D]
                        F_.wgrad_raw(h1r[p], d_t2, H, D, False, dW2a[p], None)
                        F_.linear_raw(d_t2, W2[p], dp1r[p], True, out_mask=h1r[p])
                        F_.wgrad_raw(A1[r][lo:hi], dp1r[p], E, H, False, dW1a[p], db1a[p])
                        F_.linear_raw(dp1r[p], W1[p], dA1[r][lo:hi], True, accumulate=True)
                    else:
                        dW2a[p].zero_(); dW1a[p].zero_(); db1a[p].zero_()
                    # ---- list pass
                    d_t2c = dt2c[:, s, :]
                    F_.wgrad_raw(h1c[p], d_t2c, H, D, False, dW2b[p], None)
                    F_.linear_raw(d_t2c, W2[p], dp1c[p], True, out_mask=h1c[p])
                    F_.wgrad_raw(a1c[r], dp1c[p], E, H, False, dW1b[p], db1b[p])
                    F_.linear_raw(dp1c[p], W1[p], dac[p], True)
                    dA1[r].index_add_(0, ids, dac[p])          # masked-out entries add exact zeros: order-independent
        torch._foreach_add_(dW1a + db1a + dW2a, dW1b + db1b + dW2b)     # range + list parts: one launch for all 39 sums
        grads = []
        for p in range(P):
            s = plan.slot[p]
            grads.extend([dW1a[p], db1a[p], dW2a[p], db2_all[s * D:(s + 1) * D]])
        d_att = None
        if d_att_perm is not None:
            d_att = torch.empty_like(d_att_perm)
            d_att.index_copy_(0, plan.order_t, d_att_perm)
            d_att = d_att.reshape(ctx.att_shape)
        return (None, d_att, None, None, None) + tuple(dA1) + tuple(grads)


def gcn_forward(model, metapath_idx=None, plan=None, active=None):
    """model.forward() through the fused engine (same result as the per-layer path).  With a
    row-sharded ``plan`` (distributed.ShardedGcnPlan) the result is this rank's rows.  With ``active`` (a bitmap
    from ``plan.active_bitmap``) only the marked rows of the result are computed - the demand-driven form
    ``loss()`` uses: the last step's aggregation and its transpose touch the batch's rows only."""
    if plan is None:
        plan = getattr(model, '_gcn_plan', None)
        if plan is None:
            plan = model._gcn_plan = GcnPlan(model)
    plan.gather_bf16 = getattr(model, 'gather_dtype', 'fp32') == 'bf16'
    lean = active is not None and active.ids is not None and (metapath_idx is None) and getattr(plan, 'lean_projections', False)
    params = []
    for ch in model.pea_channels:
        l0, l1 = ch.gnn_layers
        params.extend([l0.weight, l0.bias, l1.weight, l1.bias])
    plan.head_params = [(params[4 * p].detach(), params[4 * p + 1].detach()) for p in range(plan.P)] \
        if (lean and getattr(model, 'fuse_first_projection', False)) else None
    a1 = _GcnHead.apply(model.x, plan, plan.head_row_bitmaps(active) if lean else None)
    h1_full = plan.head_h1 if lean else None
    plan.head_params = None
    att = model.att if model.channel_aggr == 'att' else None
    mode = 0 if model.channel_aggr == 'att' else 1
    skip = -1 if metapath_idx is None else int(metapath_idx)
    if lean:
        active.h1_full = h1_full
        return _GcnBodyLean.apply(plan, att, mode, len(a1), active, *a1, *params)
    return _GcnBody.apply(plan, att, mode, skip, len(a1), active, *a1, *params)


class _SageBody(torch.autograd.Function):
    """PEASage counterpart of _GcnBody.  Per metapath p (nn.Linear weights are [out, in]):
         H1 = relu(M1_r Wrel1^T + brel1 + x Wroot1^T);   T2 = H1 Wrel2^T  -> its column slot;
         Z  = mean_agg(T2) + brel2  (one launch per last-step relation)  + H1 Wroot2^T (accumulated)."""

    @staticmethod
    def forward(ctx, plan, att, mode, skip, n_rel, active, x, *tensors):
        x = F_._rows(x)
        M1 = [F_._rows(t) for t in tensors[:n_rel]]
        params = [t.contiguous() for t in tensors[n_rel:]]
        P, D, H = plan.P, plan.repr, plan.hidden
        par = [params[6 * p:6 * p + 6] for p in range(P)]      # Wrel1, brel1, Wroot1, Wrel2, brel2, Wroot2
        dev, n, wide = x.device, x.shape[0], P * D
        t2 = torch.empty(n, wide, dtype=torch.float32, device=dev)
        h1 = []
        for p in range(P):
            wrel1, brel1, wroot1, wrel2, _, _ = par[p]
            h = torch.empty(n, H, dtype=torch.float32, device=dev)
            F_.linear_raw(M1[plan.rel_of_path[p]], wrel1, h, True, brel1)
            F_.linear_raw(x, wroot1, h, True, None, True, True)                    # += root term, then relu
            s = plan.slot[p]
            F_.linear_raw(h, wrel2, t2[:, s * D:(s + 1) * D], True)
            h1.append(h)
        z = (torch.zeros if active is not None else torch.empty)(n, wide, dtype=torch.float32, device=dev)
        plan.last_forward(t2, z, torch.cat([par[p][4] for p in plan.order]), active)
        del t2
        for p in range(P):
            s = plan.slot[p]
            F_.linear_raw(h1[p], par[p][5], z[:, s * D:(s + 1) * D], True, None, False, True)   # += H1 Wroot2^T
        att_perm = att.reshape(P, D).index_select(0, plan.order_t).contiguous() if att is not None else None
        out = torch.empty(n, D, dtype=torch.float32, device=dev)
        skip_slot = plan.slot[skip] if skip >= 0 else -1
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_forward', F_._ptr(z), wide, n, P, D, F_._ptr(att_perm), mode, skip_slot,
                         F_._ptr(out), D, F_._stream())
        ctx.plan, ctx.mode, ctx.skip, ctx.n_rel, ctx.active = plan, mode, skip, n_rel, active
        ctx.att_shape = att.shape if att is not None else None
        ctx.save_for_backward(z, att_perm, x, *M1, *h1, *[w for p in range(P) for w in (par[p][0], par[p][2], par[p][3], par[p][5])])
        return out

    @staticmethod
    def backward(ctx, dout):
        plan, n_rel = ctx.plan, ctx.n_rel
        if ctx.skip >= 0:
            raise RuntimeError('metapath ablation (metapath_idx) is an evaluation-only path (models/base.py:88-96)')
        P, D, H, E = plan.P, plan.repr, plan.hidden, plan.emb
        saved = ctx.saved_tensors
        z, att_perm, x = saved[0], saved[1], saved[2]
        M1 = saved[3:3 + n_rel]
        h1 = saved[3 + n_rel:3 + n_rel + P]
        ws = saved[3 + n_rel + P:]
        dev, n, wide = z.device, z.shape[0], P * D
        dout = F_._rows(dout)
        dz = torch.empty(n, wide, dtype=torch.float32, device=dev)
        d_att_perm = torch.empty(P, D, dtype=torch.float32, device=dev) if ctx.mode == 0 else None
        need = int(F_._lib.query('peagnn_fuse_workspace_floats', n, P, D)) if ctx.mode == 0 else 0
        wsp = F_._ws(need, dev) if ctx.mode == 0 else None
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_backward', F_._ptr(z), wide, n, P, D, F_._ptr(att_perm), ctx.mode, F_._ptr(dout),
                         dout.stride(0), F_._ptr(dz), wide, F_._ptr(d_att_perm), F_._ptr(wsp), need, F_._stream())
        db2_all = torch.empty(wide, dtype=torch.float32, device=dev)
        F_.wgrad_raw(None, dz, 0, wide, 0, None, db2_all)
        dt2 = plan.last_backward(dz, ctx.active)
        dM1 = [None] * n_rel
        dx = None
        grads = []
        dp1 = torch.empty(n, H, dtype=torch.float32, device=dev)
        for p in range(P):
            wrel1, wroot1, wrel2, wroot2 = ws[4 * p:4 * p + 4]
            s = plan.slot[p]
            d_z, d_t2 = dz[:, s * D:(s + 1) * D], dt2[:, s * D:(s + 1) * D]
            dwroot2 = torch.empty_like(wroot2)
            F_.wgrad_raw(h1[p], d_z, H, D, True, dwroot2, None)
            dwrel2 = torch.empty_like(wrel2)
            F_.wgrad_raw(h1[p], d_t2, H, D, True, dwrel2, None)
            F_.linear_raw(d_z, wroot2, dp1, False)                                   # dH1 = dZ Wroot2 ...
            F_.linear_raw(d_t2, wrel2, dp1, False, None, False, True, out_mask=h1[p])  # ... + dT2 Wrel2, relu-gated
            r = plan.rel_of_path[p]
            dwrel1, dbrel1 = torch.empty_like(wrel1), torch.empty(H, dtype=torch.float32, device=dev)
            F_.wgrad_raw(M1[r], dp1, E, H, True, dwrel1, dbrel1)
            dwroot1 = torch.empty_like(wroot1)
            F_.wgrad_raw(x, dp1, E, H, True, dwroot1, None)
            if dM1[r] is None:
                dM1[r] = torch.empty(n, E, dtype=torch.float32, device=dev)
                F_.linear_raw(dp1, wrel1, dM1[r], False)
            else:
                F_.linear_raw(dp1, wrel1, dM1[r], False, accumulate=True)
            if dx is None:
                dx = torch.empty(n, E, dtype=torch.float32, device=dev)
                F_.linear_raw(dp1, wroot1, dx, False)
            else:
                F_.linear_raw(dp1, wroot1, dx, False, accumulate=True)
            grads.extend([dwrel1, dbrel1, dwroot1, dwrel2, db2_all[s * D:(s + 1) * D], dwroot2])
        d_att = None
        if d_att_perm is not None:
            d_att = torch.empty_like(d_att_perm)
            d_att.index_copy_(0, plan.order_t, d_att_perm)
            d_att = d_att.reshape(ctx.att_shape)
        return (None, d_att, None, None, None, None, dx) + tuple(dM1) + tuple(grads)


class _SageBodyLean(torch.autograd.Function):
    """_SageBody for a demand-driven loss() (the PEASage counterpart of _GcnBodyLean).  SAGEConv has no self loops, so the
    last step's mean aggregation on the batch rows reads T2 only on the node-id range of its relation's sources (the RANGE
    pass), and the root term ``H1 Wroot2^T`` is wanted on the batch rows themselves (the LIST pass, on gathered [3B, .]
    rows).  A row that is in the range and on the list has its first layer computed twice with the same result; its two
    uses (T2 / root term) are different terms of Z, so both passes' gradients count.  Grouped launches (one per shape and
    pass), the range chain and the list chain on parallel branches; fixed accumulation order: bit-reproducible."""

    @staticmethod
    def forward(ctx, plan, att, mode, n_rel, active, x, *tensors):
        x = F_._rows(x)
        M1 = [F_._rows(t) for t in tensors[:n_rel]]
        params = [t.contiguous() for t in tensors[n_rel:]]
        P, D, H, E = plan.P, plan.repr, plan.hidden, plan.emb
        par = [params[6 * p:6 * p + 6] for p in range(P)]      # Wrel1, brel1, Wroot1, Wrel2, brel2, Wroot2
        dev, n, wide = x.device, x.shape[0], P * D
        ranges, rel = plan.source_ranges(), plan.rel_of_path
        ids, nl = active.ids, int(active.ids.numel())
        has = [p for p in range(P) if ranges[p][1] > ranges[p][0]]
        new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        rng = lambda t, p: t[ranges[p][0]:ranges[p][1]]
        slot = lambda t, p: t[:, plan.slot[p] * D:(plan.slot[p] + 1) * D]
        m1c = [M1[r].index_select(0, ids) for r in range(n_rel)]
        xc = x.index_select(0, ids)
        h1r = [new(max(ranges[p][1] - ranges[p][0], 0), H) for p in range(P)]
        h1c = [new(nl, H) for _ in range(P)]
        t2, zroot = new(n, wide), new(nl, wide)
        with _Fork(dev) as fork:
            with fork.on(0):                                   # range: H1 = relu(M1 Wrel1^T + brel1 + x Wroot1^T), T2 = H1 Wrel2^T
                F_.linear_grouped_raw([(rng(M1[rel[p]], p), par[p][0], par[p][1], h1r[p], None) for p in has], E, H, True)
                F_.linear_grouped_raw([(rng(x, p), par[p][2], None, h1r[p], None) for p in has], E, H, True, relu=True,
                                      accumulate=True)
                F_.linear_grouped_raw([(h1r[p], par[p][3], None, slot(rng(t2, p), p), None) for p in has], H, D, True)
            with fork.on(1):                                   # list: the same first layer, then the root term H1 Wroot2^T
                F_.linear_grouped_raw([(m1c[rel[p]], par[p][0], par[p][1], h1c[p], None) for p in range(P)], E, H, True)
                F_.linear_grouped_raw([(xc, par[p][2], None, h1c[p], None) for p in range(P)], E, H, True, relu=True,
                                      accumulate=True)
                F_.linear_grouped_raw([(h1c[p], par[p][5], None, slot(zroot, p), None) for p in range(P)], H, D, True)
        z = new(n, wide)                                       # only the batch rows are written, and only they are read
        plan.last_forward(t2, z, torch.cat([par[p][4] for p in plan.order]), active)
        del t2
        z_c = z.index_select(0, ids)
        z_c.add_(zroot)
        att_perm = att.reshape(P, D).index_select(0, plan.order_t).contiguous() if att is not None else None
        out_c = new(nl, D)
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_forward', F_._ptr(z_c), wide, nl, P, D, F_._ptr(att_perm), mode, -1,
                         F_._ptr(out_c), D, F_._stream())
        out = torch.zeros(n, D, dtype=torch.float32, device=dev)
        out.index_copy_(0, ids, out_c)                         # duplicates rewrite the same row
        ctx.plan, ctx.mode, ctx.n_rel, ctx.active, ctx.n_nodes = plan, mode, n_rel, active, n
        ctx.att_shape = att.shape if att is not None else None
        ctx.save_for_backward(z_c, att_perm, x, xc, *M1, *m1c, *h1r, *h1c,
                              *[w for p in range(P) for w in (par[p][0], par[p][2], par[p][3], par[p][5])])
        return out

    @staticmethod
    def backward(ctx, dout):
        plan, n_rel, active = ctx.plan, ctx.n_rel, ctx.active
        P, D, H, E = plan.P, plan.repr, plan.hidden, plan.emb
        saved = list(ctx.saved_tensors)
        z_c, att_perm, x, xc = saved[:4]
        pos = 4
        M1 = saved[pos:pos + n_rel]; pos += n_rel
        m1c = saved[pos:pos + n_rel]; pos += n_rel
        h1r = saved[pos:pos + P]; pos += P
        h1c = saved[pos:pos + P]; pos += P
        ws = saved[pos:]
        wrel1, wroot1 = [ws[4 * p] for p in range(P)], [ws[4 * p + 1] for p in range(P)]
        wrel2, wroot2 = [ws[4 * p + 2] for p in range(P)], [ws[4 * p + 3] for p in range(P)]
        dev, n, wide = z_c.device, ctx.n_nodes, P * D
        ranges, rel = plan.source_ranges(), plan.rel_of_path
        ids, first = active.ids, active.first
        nl = int(ids.numel())
        has = [p for p in range(P) if ranges[p][1] > ranges[p][0]]
        new = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=dev)
        rng = lambda t, p: t[ranges[p][0]:ranges[p][1]]
        slot = lambda t, p: t[:, plan.slot[p] * D:(plan.slot[p] + 1) * D]
        # each node's gradient enters once, at its first occurrence on the list
        dout_c = dout.index_select(0, ids) * first[:, None].to(torch.float32)
        dz_c = new(nl, wide)
        d_att_perm = new(P, D) if ctx.mode == 0 else None
        need = int(F_._lib.query('peagnn_fuse_workspace_floats', nl, P, D)) if ctx.mode == 0 else 0
        wsp = F_._ws(need, dev) if ctx.mode == 0 else None
        with F_._on(dev):
            F_._lib.call('peagnn_fuse_backward', F_._ptr(z_c), wide, nl, P, D, F_._ptr(att_perm), ctx.mode, F_._ptr(dout_c),
                         dout_c.stride(0), F_._ptr(dz_c), wide, F_._ptr(d_att_perm), F_._ptr(wsp), need, F_._stream())
        db2_all = new(wide)
        F_.wgrad_raw(None, dz_c, 0, wide, 0, None, db2_all)                   # every metapath's d brel2 at once
        dz = torch.zeros(n, wide, dtype=torch.float32, device=dev)
        dz.index_add_(0, ids, dz_c)                            # later occurrences add exact zeros: order-independent
        dt2 = plan.last_backward(dz, active)
        dM1 = [torch.zeros(n, E, dtype=torch.float32, device=dev) for _ in range(n_rel)]
        dx = torch.zeros(n, E, dtype=torch.float32, device=dev)
        # range-pass (a) and list-pass (b) parts of every parameter gradient, summed at the end
        dwrel1a, dbrel1a, dwroot1a = [new(H, E) for _ in range(P)], [new(H) for _ in range(P)], [new(H, E) for _ in range(P)]
        dwrel1b, dbrel1b, dwroot1b = [new(H, E) for _ in range(P)], [new(H) for _ in range(P)], [new(H, E) for _ in range(P)]
        dwrel2, dwroot2 = [new(D, H) for _ in range(P)], [new(D, H) for _ in range(P)]
        dp1r = [new(max(ranges[p][1] - ranges[p][0], 0), H) for p in range(P)]
        dp1c, dm1c, dxc = [new(nl, H) for _ in range(P)], [new(nl, E) for _ in range(P)], [new(nl, E) for _ in range(P)]

        def rounds_by(key):
            """has-metapaths that accumulate into one buffer region (same key) go into successive launches."""
            out, seen = [], {}
            for p in has:
                k = seen.get(key(p), 0)
                seen[key(p)] = k + 1
                while len(out) <= k:
                    out.append([])
                out[k].append(p)
            return out
        with _Fork(dev) as fork:
            with fork.on(0):                                   # range: d P1 = (d T2 Wrel2) gated by relu -> d M1, d x
                F_.linear_grouped_raw([(slot(rng(dt2, p), p), wrel2[p], None, dp1r[p], h1r[p]) for p in has], D, H, False)
                have_dp1r = fork.record()
                for members in rounds_by(lambda p: rel[p]):
                    F_.linear_grouped_raw([(dp1r[p], wrel1[p], None, rng(dM1[rel[p]], p), None) for p in members], H, E, False,
                                          accumulate=True)
                for members in rounds_by(lambda p: ranges[p]):
                    F_.linear_grouped_raw([(dp1r[p], wroot1[p], None, rng(dx, p), None) for p in members], H, E, False,
                                          accumulate=True)
            with fork.on(1):                                   # range: weight gradients
                F_.wgrad_grouped_raw([(h1r[p], slot(rng(dt2, p), p), dwrel2[p], None) for p in range(P)], H, D, True)
                fork.wait(have_dp1r)
                F_.wgrad_grouped_raw([(rng(M1[rel[p]], p), dp1r[p], dwrel1a[p], dbrel1a[p]) for p in range(P)], E, H, True)
                F_.wgrad_grouped_raw([(rng(x, p), dp1r[p], dwroot1a[p], None) for p in range(P)], E, H, True)
            with fork.on(2):                                   # list: d P1 = (d Z Wroot2) gated by relu -> d M1, d x rows
                F_.linear_grouped_raw([(slot(dz_c, p), wroot2[p], None, dp1c[p], h1c[p]) for p in range(P)], D, H, False)
                have_dp1c = fork.record()
                F_.linear_grouped_raw([(dp1c[p], wrel1[p], None, dm1c[p], None) for p in range(P)], H, E, False)
                F_.linear_grouped_raw([(dp1c[p], wroot1[p], None, dxc[p], None) for p in range(P)], H, E, False)
            with fork.on(3):                                   # list: weight gradients
                F_.wgrad_grouped_raw([(h1c[p], slot(dz_c, p), dwroot2[p], None) for p in range(P)], H, D, True)
                fork.wait(have_dp1c)
                F_.wgrad_grouped_raw([(m1c[rel[p]], dp1c[p], dwrel1b[p], dbrel1b[p]) for p in range(P)], E, H, True)
                F_.wgrad_grouped_raw([(xc, dp1c[p], dwroot1b[p], None) for p in range(P)], E, H, True)
        for p in range(P):                                     # non-first list entries carry exact zeros: order-independent
            dM1[rel[p]].index_add_(0, ids, dm1c[p])
            dx.index_add_(0, ids, dxc[p])
        torch._foreach_add_(dwrel1a + dbrel1a + dwroot1a, dwrel1b + dbrel1b + dwroot1b)
        grads = []
        for p in range(P):
            s = plan.slot[p]
            grads.extend([dwrel1a[p], dbrel1a[p], dwroot1a[p], dwrel2[p], db2_all[s * D:(s + 1) * D], dwroot2[p]])
        d_att = None
        if d_att_perm is not None:
            d_att = torch.empty_like(d_att_perm)
            d_att.index_copy_(0, plan.order_t, d_att_perm)
            d_att = d_att.reshape(ctx.att_shape)
        return (None, d_att, None, None, None, dx) + tuple(dM1) + tuple(grads)


def sage_forward(model, metapath_idx=None, active=None):
    """model.forward() of a standard PEASage model through the fused engine."""
    plan = getattr(model, '_sage_plan', None)
    if plan is None:
        plan = model._sage_plan = GcnPlan(model, 'sage')
    lean = (SAGE_LEAN and active is not None and active.ids is not None and metapath_idx is None and plan.grouped_shapes)
    m1 = _GcnHead.apply(model.x, plan, plan.head_row_bitmaps(active) if lean else None)
    params = []
    for ch in model.pea_channels:
        l0, l1 = ch.gnn_layers
        params.extend([l0.lin_rel.weight, l0.lin_rel.bias, l0.lin_root.weight,
                       l1.lin_rel.weight, l1.lin_rel.bias, l1.lin_root.weight])
    att = model.att if model.channel_aggr == 'att' else None
    mode = 0 if model.channel_aggr == 'att' else 1
    if lean:
        return _SageBodyLean.apply(plan, att, mode, len(m1), active, model.x, *m1, *params)
    skip = -1 if metapath_idx is None else int(metapath_idx)
    return _SageBody.apply(plan, att, mode, skip, len(m1), active, model.x, *m1, *params)
