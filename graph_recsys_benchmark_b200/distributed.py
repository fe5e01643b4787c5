"""1-D destination-row sharding of the metapath propagation across the GPUs of one box
(BASELINE.json north_star).  The reference has no distributed code at all -
this is the multi-GPU form of the same arithmetic:

  * rank r of R owns the destination rows i with i % R == r (cyclic, so popular items / active
    users / hot tags spread evenly; node ids are contiguous by type upstream, a blocked split
    would put every item on one rank);
  * an aggregation step computes only the owned rows; what the next step gathers from is
    re-assembled with ONE NCCL all-gather per step (its backward is a reduce-scatter);
  * gathered tables live in "rank-major" order: position pi(i) = (i % R) * rows_per_rank + i // R,
    which is exactly what all_gather_into_tensor produces, so no re-shuffle is needed - the shard's
    column indices are stored pre-permuted;
  * self loops (GCN) are stored as explicit edges of the owning rank, so the transposed pass needs
    no special case;
  * BPR minibatches are data-parallel (each rank scores its own triples against the full fused
    representation); parameter gradients - the embedding table's partial d_x included - are summed
    with one all-reduce per step.
One process per GPU; torch.distributed (NCCL over NVLink 5 / NVSwitch) owns the communicator.
"""
import torch
import torch.distributed as dist

from . import functional as F_
from .graph import build_csr


class ShardPlan(object):
    def __init__(self, num_nodes, world_size, rank):
        self.num_nodes, self.world, self.rank = int(num_nodes), int(world_size), int(rank)
        self.rows_per_rank = (self.num_nodes + self.world - 1) // self.world
        self.padded = self.rows_per_rank * self.world

    def owner(self, ids):
        return ids % self.world

    def local_index(self, ids):
        return ids // self.world

    def to_rank_major(self, ids):
        """pi(i): position of node i in an all-gathered (rank-major) table."""
        return (ids % self.world) * self.rows_per_rank + ids // self.world

    def local_global_ids(self, rank=None, device=None):
        """Global ids of this rank's rows, padded with -1 up to rows_per_rank."""
        r = self.rank if rank is None else rank
        ids = torch.arange(self.rows_per_rank, device=device, dtype=torch.long) * self.world + r
        return torch.where(ids < self.num_nodes, ids, torch.full_like(ids, -1))

    def rank_major_to_global(self, device=None):
        """[padded] global id stored at each rank-major position (-1 for padding)."""
        return torch.cat([self.local_global_ids(r, device) for r in range(self.world)])


def shard_coo(edge_index, plan, rank=None, explicit_self_loops=False, drop_self_loops=True):
    """The edges a rank owns (target % R == rank), as (src_global, dst_local) int64 tensors.
    With ``explicit_self_loops`` one loop per owned row is appended (GCNConv's add_remaining_self_loops
    after dropping existing loops)."""
    r = plan.rank if rank is None else rank
    src, dst = edge_index[0], edge_index[1]
    keep = (dst % plan.world) == r
    if drop_self_loops:
        keep &= src != dst
    src, dst = src[keep], dst[keep]
    if explicit_self_loops:
        own = plan.local_global_ids(r, edge_index.device)
        own = own[own >= 0]
        src = torch.cat([src, own])
        dst = torch.cat([dst, own])
    return src, dst // plan.world


class ShardedRelation(object):
    """One rank's part of a relation, for one conv family ('gcn' or 'sage').

    fwd(layout)  CSR over the owned rows gathering from a table in ``layout`` ('orig' = the
                 embedding table / any [N, F] tensor in node-id order, 'rm' = an all-gathered
                 rank-major table);
    bwd(layout)  its transpose: rows = the table's rows, gathering from the local [rows_per_rank, F]
                 gradient."""

    def __init__(self, edge_index, plan, kind):
        self.plan, self.kind = plan, kind
        dev = edge_index.device
        n = plan.num_nodes
        src, dst = edge_index[0], edge_index[1]
        loops = kind in ('gcn', 'gat')      # these convs drop self-loop edges and add one loop per node
        nl = src != dst if loops else torch.ones_like(src, dtype=torch.bool)
        if kind == 'gcn':
            deg = torch.bincount(src[nl], minlength=n).float() + 1.0       # source side + self loop
            scale = deg.pow(-0.5)
        elif kind == 'sage':
            cnt = torch.bincount(dst, minlength=n).float()
            scale = 1.0 / cnt.clamp(min=1.0)
        else:
            scale = torch.ones(n, device=dev)
        own = plan.local_global_ids(device=dev)
        valid = own >= 0
        self.scale_orig = scale.contiguous()
        rm_ids = plan.rank_major_to_global(dev)
        self.scale_rm = torch.where(rm_ids >= 0, scale[rm_ids.clamp(min=0)], torch.zeros_like(rm_ids, dtype=scale.dtype)).contiguous()
        self.scale_local = torch.where(valid, scale[own.clamp(min=0)], torch.zeros_like(own, dtype=scale.dtype)).contiguous()
        self.src, self.dst_local = shard_coo(edge_index, plan, explicit_self_loops=loops, drop_self_loops=loops)
        self.explicit_self_loops = loops
        self.n_loops = int(valid.sum().item()) if loops else 0          # the explicit self loops are the LAST n_loops edges
        # node-id range holding every source of the relation (ids are contiguous by type upstream): what a step that
        # gathers through this relation actually reads of an exchanged table
        real = src[nl]
        self.src_lo = int(real.min().item()) if real.numel() else 0
        self.src_hi = int(real.max().item()) + 1 if real.numel() else 0
        self._fwd, self._bwd, self._perm = {}, {}, {}
        self._src_layout = None

    def src_layout(self):
        """Compact exchange layout: every rank contributes only its rows inside the source range
        [src_lo, src_hi) - a contiguous local slice [start_q, start_q + cmax) - and the table a shard gathers from is
        [R * cmax exchanged rows | this rank's own rows (the self-loop targets, no exchange needed)]."""
        if self._src_layout is None:
            R, rpr, n = self.plan.world, self.plan.rows_per_rank, self.plan.num_nodes
            lo_q = [max(0, -((q - self.src_lo) // R)) for q in range(R)]
            hi_q = [min(rpr, max(0, -((q - self.src_hi) // R))) for q in range(R)]
            cmax = max(1, max(h - l for l, h in zip(lo_q, hi_q)))
            start = [max(0, min(l, rpr - cmax)) for l in lo_q]
            dev = self.src.device
            start_t = torch.tensor(start, dtype=torch.long, device=dev)
            n_real = self.src.numel() - self.n_loops
            s_real, loops_local = self.src[:n_real], self.dst_local[n_real:]
            owner = s_real % R
            pos = owner * cmax + (s_real // R - start_t[owner])
            assert n_real == 0 or (int(pos.min()) >= 0 and bool(((s_real // R - start_t[owner]) < cmax).all()))
            cols = torch.cat([pos, R * cmax + loops_local])
            # column scale of every table row: exchanged part by the node it holds (0 for slots beyond N), then the own rows
            q = torch.arange(R, device=dev).repeat_interleave(cmax)
            j = torch.arange(cmax, device=dev).repeat(R)
            gid = (start_t[q] + j) * R + q
            ok = gid < n
            cs_x = torch.where(ok, self.scale_orig[gid.clamp(max=n - 1)], torch.zeros_like(gid, dtype=self.scale_orig.dtype))
            self._src_layout = dict(cmax=cmax, start=start, rows=R * cmax + rpr, cols=cols,
                                    scale=torch.cat([cs_x, self.scale_local]).contiguous())
        return self._src_layout

    def _cols(self, layout):
        if layout == 'src':
            return self.src_layout()['cols']
        return self.src if layout == 'orig' else self.plan.to_rank_major(self.src)

    def fwd(self, layout):
        if layout not in self._fwd:
            csr = build_csr(self.dst_local, self._cols(layout), self.plan.rows_per_rank, False)
            csr.explicit_self_loops = self.explicit_self_loops
            self._fwd[layout] = csr
        return self._fwd[layout]

    def bwd(self, layout):
        if layout not in self._bwd:
            rows = self.plan.num_nodes if layout == 'orig' else \
                (self.src_layout()['rows'] if layout == 'src' else self.plan.padded)
            csr = build_csr(self._cols(layout), self.dst_local, rows, False)
            csr.explicit_self_loops = self.explicit_self_loops
            self._bwd[layout] = csr
        return self._bwd[layout]

    def bwd_to_fwd(self, layout):
        """perm[k] = position in the target-grouped shard arrays of the k-th source-grouped edge."""
        if layout not in self._perm:
            f, b = self.fwd(layout), self.bwd(layout)
            inv = torch.empty(max(int(self.src.numel()), 1), dtype=torch.int32, device=f.eid.device)
            inv[f.eid.long()] = torch.arange(f.nnz, dtype=torch.int32, device=inv.device)
            self._perm[layout] = inv[b.eid.long()].contiguous()
        return self._perm[layout]

    def table_scale(self, layout):
        if layout == 'src':
            return self.src_layout()['scale']
        return self.scale_orig if layout == 'orig' else self.scale_rm


class _ShardAggregate(torch.autograd.Function):
    """Owned rows of  diag(rs) A diag(cs) X (+ bias)  over a ShardedRelation; X is a full table."""

    @staticmethod
    def forward(ctx, X, bias, rel, layout):
        X = F_._rows(F_._req(X, 'table'))
        feat = X.shape[1]
        out = torch.empty(rel.plan.rows_per_rank, feat, dtype=torch.float32, device=X.device)
        if rel.kind == 'gcn':
            rs, cs = rel.scale_local, rel.table_scale(layout)
        else:
            rs, cs = rel.scale_local, None
        F_.spmm_raw(rel.fwd(layout), X, feat, out, rs, cs, False, bias, False)
        ctx.rel, ctx.layout, ctx.rows = rel, layout, X.shape[0]
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        rel, layout = ctx.rel, ctx.layout
        dout = F_._rows(dout)
        feat = dout.shape[1]
        dX = db = None
        if ctx.needs_input_grad[0]:
            dX = torch.empty(ctx.rows, feat, dtype=torch.float32, device=dout.device)
            if rel.kind == 'gcn':
                rs, cs = rel.table_scale(layout), rel.scale_local
            else:
                rs, cs = None, rel.scale_local
            F_.spmm_raw(rel.bwd(layout), dout, feat, dX, rs, cs, False)
        if ctx.has_bias and ctx.needs_input_grad[1]:
            # padding rows carry the bias too; their upstream gradient is zero by construction
            db = torch.empty(feat, dtype=torch.float32, device=dout.device)
            F_.wgrad_raw(None, dout, 0, feat, 0, None, db)
        return dX, db, None, None


def shard_aggregate(X, rel, layout, bias=None):
    return _ShardAggregate.apply(X, bias, rel, layout)


class _ShardGatAggregate(torch.autograd.Function):
    """Owned rows of the GAT edge-softmax aggregate.  H [padded, heads*F] and a_j [padded, heads] are
    rank-major tables, a_i [rows_per_rank, heads] is local; the shard stores self loops as edges."""

    @staticmethod
    def forward(ctx, H, ai, aj, bias, rel, heads, relu):
        import ctypes as C
        from . import _lib
        from .graph import _ptr, _stream
        H = F_._rows(F_._req(H, 'h'))
        rpr = rel.plan.rows_per_rank
        feat = H.shape[1] // heads
        dev = H.device
        ai, aj = ai.contiguous(), aj.contiguous()
        rowmax = torch.empty(rpr, heads, dtype=torch.float32, device=dev)
        denom = torch.empty_like(rowmax)
        out = torch.empty(rpr, heads * feat, dtype=torch.float32, device=dev)
        view = rel.fwd('rm').view(feat, heads)
        with F_._on(dev):
            _lib.call('peagnn_gat_rowmax', C.byref(view), _ptr(ai), _ptr(aj), heads, F_.NEG_SLOPE, _ptr(rowmax), _stream())
            _lib.call('peagnn_gat_aggregate', C.byref(view), _ptr(H), H.stride(0), feat, heads, _ptr(ai), _ptr(aj),
                      F_.NEG_SLOPE, _ptr(rowmax), _ptr(denom), _ptr(out), out.stride(0), _ptr(bias), int(relu), _stream())
        ctx.rel, ctx.heads, ctx.relu, ctx.has_bias = rel, heads, relu, bias is not None
        ctx.save_for_backward(H, ai, aj, rowmax, denom, out, bias)
        return out

    @staticmethod
    def backward(ctx, dout):
        import ctypes as C
        from . import _lib
        from .graph import _ptr, _stream
        H, ai, aj, rowmax, denom, out, bias = ctx.saved_tensors
        rel, heads = ctx.rel, ctx.heads
        feat = H.shape[1] // heads
        dev = H.device
        dout = F_._rows(dout)
        if ctx.relu:
            dout = F_.relu_backward_raw(dout, out)
        db = None
        if ctx.has_bias and ctx.needs_input_grad[3]:
            db = torch.empty(heads * feat, dtype=torch.float32, device=dev)
            F_.wgrad_raw(None, dout, 0, heads * feat, 0, None, db)
        fwd, bwd = rel.fwd('rm'), rel.bwd('rm')
        ads_e = torch.empty(max(fwd.nnz, 1), heads, 2, dtype=torch.float32, device=dev)     # (alpha, ds) per edge
        d_ai = torch.empty_like(ai)
        d_aj = torch.empty_like(aj)
        dH = torch.empty(H.shape[0], heads * feat, dtype=torch.float32, device=dev)
        vf, vb = fwd.view(feat, heads), bwd.view(feat, heads)
        perm = rel.bwd_to_fwd('rm')
        with F_._on(dev):
            _lib.call('peagnn_gat_backward_dst', C.byref(vf), _ptr(H), H.stride(0), feat, heads, _ptr(ai), _ptr(aj),
                      F_.NEG_SLOPE, _ptr(rowmax), _ptr(denom), _ptr(out), out.stride(0), _ptr(bias), _ptr(dout),
                      dout.stride(0), _ptr(ads_e), _ptr(None), _ptr(None), _ptr(d_ai), _stream())
            _lib.call('peagnn_gat_backward_src', C.byref(vb), _ptr(perm), _ptr(ads_e), _ptr(None),
                      _ptr(None), _ptr(dout), dout.stride(0), feat, heads, _ptr(dH), dH.stride(0), _ptr(d_aj), _stream())
        return dH, d_ai, d_aj, db, None, None, None


def shard_gat_aggregate(H, ai, aj, rel, heads, bias=None, relu=False):
    return _ShardGatAggregate.apply(H, ai, aj, bias, rel, heads, relu)


def raw_all_gather(local, group=None, out=None):
    """[rows_per_rank, F] -> [R * rows_per_rank, F] (rank-major).  NCCL all_gather_into_tensor; gloo
    (CPU / single-GPU test rigs) goes through the list form.  ``out`` lets the caller keep one
    persistent receive buffer (buffers that cross to NCCL's stream are slow to return to the
    caching allocator; re-allocating 100s of MB per step ends in cudaMalloc / cudaFree stalls)."""
    local = local.contiguous()
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty(world * local.shape[0], local.shape[1], dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == 'gloo':
        dist.all_gather(list(out.chunk(world, dim=0)), local, group=group)
    else:
        dist.all_gather_into_tensor(out, local, group=group)
    return out


def raw_reduce_scatter(full, group=None, out=None):
    """Sum over ranks of [R * rows_per_rank, F], each rank keeping its own row block."""
    full = full.contiguous()
    world = dist.get_world_size(group)
    rows = full.shape[0] // world
    if dist.get_backend(group) == 'gloo':               # gloo has no reduce-scatter: all-reduce, keep own slice
        full = full.clone()
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
        r = dist.get_rank(group)
        return full[r * rows:(r + 1) * rows].contiguous()
    local = out if out is not None else torch.empty(rows, full.shape[1], dtype=full.dtype, device=full.device)
    dist.reduce_scatter_tensor(local, full, op=dist.ReduceOp.SUM, group=group)
    return local


class _AllGatherRows(torch.autograd.Function):
    """[rows_per_rank, F] per rank -> [R * rows_per_rank, F] rank-major on every rank.
    Backward: every rank holds a gradient for the whole table; the owner needs their sum ->
    reduce-scatter."""

    @staticmethod
    def forward(ctx, local, group):
        ctx.group = group
        return raw_all_gather(local, group)

    @staticmethod
    def backward(ctx, dfull):
        return raw_reduce_scatter(dfull, ctx.group), None


def all_gather_rows(local, group=None):
    return _AllGatherRows.apply(local, group)


_flat_buffers = {}


def allreduce_gradients(params, group=None):
    """Sum the parameter gradients over ranks with one flat all-reduce (data-parallel BPR batches +
    row-sharded propagation both leave per-rank partial sums)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    total = sum(g.numel() for g in grads)
    key = (total, grads[0].device, grads[0].dtype)
    flat = _flat_buffers.get(key)
    if flat is None:                                   # one persistent bucket per model (see raw_all_gather)
        flat = _flat_buffers[key] = torch.empty(total, dtype=grads[0].dtype, device=grads[0].device)
    torch.cat([g.reshape(-1) for g in grads], out=flat)            # one launch in ...
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g))
        off += n
    torch._foreach_copy_(grads, views)                             # ... and one multi-tensor launch out


class ShardedPropagation(object):
    """Row-sharded forward of a PEAGCN / PEASage model (attached by ``shard_model``)."""

    def __init__(self, model, plan, group=None):
        self.model, self.plan, self.group = model, plan, group
        self.kind = None
        self._rels = {}
        first = model.pea_channels[0].gnn_layers[0]
        tag = getattr(first, 'shares_aggregate', None)
        if tag is None and hasattr(first, 'att_i'):
            tag = 'gat'
        if tag not in ('gcn', 'sage', 'gat'):
            raise NotImplementedError('row-sharded propagation covers the PEAGCN / PEASage / PEAGAT families')
        self.kind = tag
        dev = model.x.device
        rm_ids = plan.rank_major_to_global(dev)
        self.rm_ids = rm_ids.clamp(min=0)                   # padding positions read node 0 (never gathered)
        own = plan.local_global_ids(device=dev)
        self.own_ids = own.clamp(min=0)
        self.own_valid = (own >= 0)
        self.rm_of_global = plan.to_rank_major(torch.arange(plan.num_nodes, device=dev))
        self._plan, self._applies = None, None

    def _engine_applies(self):
        if self._applies is None:
            from .engine import GcnPlan
            self._applies = GcnPlan.applies(self.model)
        return self._applies

    def relation(self, edge_index):
        key = (edge_index.data_ptr(), int(edge_index.shape[1]))
        hit = self._rels.get(key)
        if hit is None:
            for (_, e), (other, rel) in self._rels.items():
                if e == edge_index.shape[1] and torch.equal(other, edge_index):
                    hit = (edge_index, rel)
                    break
            if hit is None:
                hit = (edge_index, ShardedRelation(edge_index, self.plan, self.kind))
            self._rels[key] = hit
        return hit[1]

    def _local_rows(self, state):
        layout, t = state
        if layout == 'local':
            return t
        if layout == 'orig':
            return t.index_select(0, self.own_ids)      # padding rows read node 0; never used downstream
        raise AssertionError(layout)

    def _table(self, state):
        layout, t = state
        if layout == 'local':
            return 'rm', all_gather_rows(t, self.group)
        return layout, t

    def _gat_gather(self, parts):
        """One all-gather for a list of (H_local [rpr, HF], a_j_local [rpr, heads]) pairs; returns the
        rank-major (H_full view, a_j_full) per pair."""
        cols = [h for h, _ in parts] + [a for _, a in parts]       # feature blocks first: 16-byte aligned slices
        width = sum(t.shape[1] for t in cols)
        pad = (-width) % 4
        if pad:
            cols.append(cols[0].new_zeros(cols[0].shape[0], pad))
        full = all_gather_rows(torch.cat(cols, dim=1), self.group)
        out = []
        off_h, off_a = 0, sum(h.shape[1] for h, _ in parts)
        for h_loc, aj_loc in parts:
            hf, hd = h_loc.shape[1], aj_loc.shape[1]
            out.append((full[:, off_h:off_h + hf], full[:, off_a:off_a + hd].contiguous()))
            off_h += hf
            off_a += hd
        return out

    def _gat_channel_outputs(self):
        model = self.model
        plan, r = self.plan, self.plan.rank
        rpr = plan.rows_per_rank
        x_rm = model.x.index_select(0, self.rm_ids)            # embedding table in rank-major order, once
        pending = []
        for idx, channel in enumerate(model.pea_channels):
            eil = model.meta_path_edge_index_list[idx]
            assert len(eil) == channel.num_steps
            local = None
            for s, layer in enumerate(channel.gnn_layers):
                if layer.training and layer.dropout > 0:
                    raise NotImplementedError('attention dropout > 0 is not supported by the sm_100a GAT kernels')
                rel = self.relation(eil[s])
                last = s == channel.num_steps - 1
                att_i, att_j = layer.att_i.view(-1), layer.att_j.view(-1)
                if local is None:
                    # first step: every rank projects the whole (replicated) table - cheaper than
                    # gathering a [N, heads*hidden] table per metapath over NVLink
                    h_full = F_.linear(x_rm, layer.lin.weight, None, w_is_out_in=True)
                    ai_full, aj_full = F_.gat_scores(h_full, att_i, att_j, layer.heads)
                    ai_loc = ai_full[r * rpr:(r + 1) * rpr]
                else:
                    h_loc = F_.linear(local, layer.lin.weight, None, w_is_out_in=True)
                    ai_loc, aj_loc = F_.gat_scores(h_loc, att_i, att_j, layer.heads)
                    if last:
                        pending.append((layer, rel, h_loc, ai_loc, aj_loc))
                        break
                    (h_full, aj_full), = self._gat_gather([(h_loc, aj_loc)])
                local = shard_gat_aggregate(h_full, ai_loc, aj_full, rel, layer.heads, layer.bias, relu=not last)
            else:
                pending.append(local)
        todo = [k for k, o in enumerate(pending) if isinstance(o, tuple)]
        if todo:
            gathered = self._gat_gather([(pending[k][2], pending[k][4]) for k in todo])   # ONE all-gather
            for k, (h_full, aj_full) in zip(todo, gathered):
                layer, rel, _, ai_loc, _ = pending[k]
                pending[k] = shard_gat_aggregate(h_full, ai_loc, aj_full, rel, layer.heads, layer.bias, relu=False)
        return pending

    def channel_outputs(self):
        if self.kind == 'gat':
            return self._gat_channel_outputs()
        model = self.model
        shared, pending = {}, []
        for idx, channel in enumerate(model.pea_channels):
            eil = model.meta_path_edge_index_list[idx]
            assert len(eil) == channel.num_steps
            state = ('orig', model.x)
            result = None
            for s, layer in enumerate(channel.gnn_layers):
                rel = self.relation(eil[s])
                last = s == channel.num_steps - 1
                if layer.in_channels <= layer.out_channels:           # aggregate, then project
                    layout, table = self._table(state)
                    key = (id(rel), layout, id(table))
                    if key not in shared:
                        shared[key] = shard_aggregate(table, rel, layout)
                    agg = shared[key]
                    if self.kind == 'gcn':
                        out = F_.linear(agg, layer.weight, layer.bias, w_is_out_in=False, relu=not last)
                    else:
                        relp = F_.linear(agg, layer.lin_rel.weight, layer.lin_rel.bias, w_is_out_in=True)
                        out = F_.linear_accumulate(self._local_rows(state), layer.lin_root.weight, relp, relu=not last)
                    state = ('local', out)
                else:                                                  # project, then aggregate
                    loc = self._local_rows(state)
                    t = layer.project(loc)
                    if last and model.batch_last_step:
                        result = (layer, rel, t, loc)
                        break
                    table = all_gather_rows(t, self.group)
                    agg = shard_aggregate(table, rel, 'rm', layer.post_bias())
                    if self.kind == 'sage':
                        agg = F_.linear_accumulate(loc, layer.lin_root.weight, agg, relu=not last)
                    elif not last:
                        raise NotImplementedError('relu after a narrowing GCN step inside a sharded channel')
                    state = ('local', agg)
            pending.append(result if result is not None else state[1])
        groups = {}
        for idx, o in enumerate(pending):
            if isinstance(o, tuple):
                groups.setdefault(id(o[1]), []).append(idx)
        for members in groups.values():
            rel = pending[members[0]][1]
            ts = [pending[m][2] for m in members]
            t_cat = ts[0] if len(ts) == 1 else torch.cat(ts, dim=1)
            table = all_gather_rows(t_cat, self.group)                 # ONE all-gather per relation group
            bias = torch.cat([pending[m][0].post_bias() for m in members])
            agg = shard_aggregate(table, rel, 'rm', bias)
            parts = torch.split(agg, [t.shape[1] for t in ts], dim=1) if len(ts) > 1 else (agg,)
            for m, part in zip(members, parts):
                layer, _, _, loc = pending[m]
                pending[m] = layer.finish(part, loc, False)
        return pending

    def supports_demand_driven(self):
        return self.kind == 'gcn' and self.model.fused_engine and self._engine_applies()

    def active_bitmap(self, ids):
        """Bitmap over this rank's LOCAL rows of the representation rows ANY rank's batch reads: the batches are
        data-parallel, but a rank computes its owned rows for everybody, so the node ids are all-gathered first
        (3 x B int64 per rank - one small collective)."""
        ids = ids.reshape(-1).contiguous()
        world = self.plan.world
        everyone = torch.empty(world * ids.numel(), dtype=ids.dtype, device=ids.device)
        if dist.get_backend(self.group) == 'gloo':
            dist.all_gather(list(everyone.chunk(world)), ids, group=self.group)
        else:
            dist.all_gather_into_tensor(everyone, ids, group=self.group)
        from .engine import ActiveSet
        return ActiveSet(F_.mark_rows(everyone, self.plan.rows_per_rank, mod=world, rem=self.plan.rank))

    def forward(self, metapath_idx=None, active_ids=None):
        model = self.model
        if self.kind == 'gcn' and model.fused_engine and self._engine_applies():
            from .engine import gcn_forward
            if self._plan is None:
                self._plan = ShardedGcnPlan(model, self)
            active = self.active_bitmap(active_ids) if active_ids is not None else None
            fused_local = gcn_forward(model, metapath_idx, plan=self._plan, active=active)
        else:
            z = torch.stack(self.channel_outputs(), dim=1)             # [rows_per_rank, P, repr]
            att = model.att if model.channel_aggr == 'att' else None
            fused_local = F_.fuse_channels(z, att, model.channel_aggr, metapath_idx)
        fused_rm = all_gather_rows(fused_local, self.group)            # [padded, repr], rank-major
        return fused_rm.index_select(0, self.rm_of_global)             # node-id order, as the API promises


class ShardedGcnPlan(object):
    """engine.GcnPlan for row shards: same column layout, the two aggregation phases run on the
    owned rows with ONE all-gather (forward) / reduce-scatter (backward) of the [rows, P*repr] table."""
    lean_projections = False     # the range / list projection passes are single-GPU only for now

    def __init__(self, model, sp):
        self.sp = sp
        eil = model.meta_path_edge_index_list
        self.P = len(model.pea_channels)
        self.first_rels, self.rel_of_path, seen = [], [], {}
        for p in range(self.P):
            rel = sp.relation(eil[p][0])
            if id(rel) not in seen:
                seen[id(rel)] = len(self.first_rels)
                self.first_rels.append(rel)
            self.rel_of_path.append(seen[id(rel)])
        groups, index = [], {}
        for p in range(self.P):
            rel = sp.relation(eil[p][1])
            if id(rel) not in index:
                index[id(rel)] = len(groups)
                groups.append((rel, []))
            groups[index[id(rel)]][1].append(p)
        self.groups = groups
        self.order = [p for _, members in groups for p in members]
        self.slot = {p: s for s, p in enumerate(self.order)}
        self.order_t = torch.tensor(self.order, dtype=torch.long, device=model.x.device)
        first = model.pea_channels[0].gnn_layers
        self.emb, self.hidden, self.repr = first[0].in_channels, first[0].out_channels, first[1].out_channels
        self.grouped_shapes = (self.emb, self.hidden, self.repr) == (64, 64, 16)
        self._table = self._dtab = None
        self._xbuf = {}

    def head_forward(self, x):
        from .engine import _Fork
        rpr = self.sp.plan.rows_per_rank
        outs = [torch.empty(rpr, x.shape[1], dtype=torch.float32, device=x.device) for _ in self.first_rels]
        for rel in self.first_rels:
            rel.fwd('orig').view(x.shape[1])                     # structures / workspaces are built on this stream, before the fork
        with _Fork(x.device) as fork:                            # independent launches on this rank's shards: parallel branches
            for k, (rel, out) in enumerate(zip(self.first_rels, outs)):
                with fork.on(k):
                    F_.spmm_raw(rel.fwd('orig'), x, x.shape[1], out, rel.scale_local, rel.scale_orig, False)
        return outs

    def head_backward(self, grads):
        dx = None
        for rel, d in zip(self.first_rels, grads):
            if d is None:
                continue
            d = F_._rows(d)
            if dx is None:
                dx = torch.empty(self.sp.plan.num_nodes, d.shape[1], dtype=torch.float32, device=d.device)
                F_.spmm_raw(rel.bwd('orig'), d, d.shape[1], dx, rel.scale_orig, rel.scale_local, False)
            else:
                # an accumulating launch only visits the table rows this rank's part of the relation reaches (its edges'
                # sources and, as explicit self-loop edges, the rows it owns): re-reading and re-writing all N rows of the
                # partial table for a relation with a handful of sources cost more than the gather itself
                t = rel.bwd('orig')
                F_.spmm_raw(t, d, d.shape[1], dx, rel.scale_orig, rel.scale_local, False, accumulate=True,
                            active_rows=t.nonempty_row_bitmap())
        return dx

    compact_exchange = True      # exchange only each last-step relation's source-type rows (False: the whole [N, P*repr] table)

    def _exchange_buffers(self, k, rows, width, dev):
        key = (k, rows, width)
        if key not in self._xbuf:
            self._xbuf[key] = (torch.empty(rows, width, dtype=torch.float32, device=dev),
                               torch.empty(rows, width, dtype=torch.float32, device=dev))
        return self._xbuf[key]

    def last_forward(self, t2, z, bias_all, active=None):
        rows_bm = active.bitmap if active is not None else None
        D, start = self.repr, 0
        if self.compact_exchange:
            R, rpr, group = self.sp.plan.world, self.sp.plan.rows_per_rank, self.sp.group
            for k, (rel, members) in enumerate(self.groups):
                width = len(members) * D
                lay = rel.src_layout()
                cmax, s0 = lay['cmax'], lay['start'][self.sp.plan.rank]
                table, _ = self._exchange_buffers(k, lay['rows'], width, t2.device)
                raw_all_gather(t2[s0:s0 + cmax, start:start + width], group, out=table[:R * cmax])   # source-range rows only
                table[R * cmax:].copy_(t2[:, start:start + width])                                   # own rows: self loops
                F_.spmm_raw(rel.fwd('src'), table, width, z[:, start:start + width], rel.scale_local, lay['scale'], False,
                            bias_all[start:start + width], active_rows=rows_bm)
                start += width
            return
        if self._table is None or self._table.shape != (self.sp.plan.padded, t2.shape[1]):
            self._table = torch.empty(self.sp.plan.padded, t2.shape[1], dtype=torch.float32, device=t2.device)
        table = raw_all_gather(t2, self.sp.group, out=self._table)     # the step's only all-gather
        for rel, members in self.groups:
            width = len(members) * D
            F_.spmm_raw(rel.fwd('rm'), table[:, start:start + width], width, z[:, start:start + width],
                        rel.scale_local, rel.scale_rm, False, bias_all[start:start + width], active_rows=rows_bm)
            start += width

    def last_backward(self, dz, active=None):
        cols_bm = active.bitmap if active is not None else None
        D, start = self.repr, 0
        if self.compact_exchange:
            R, rpr, group = self.sp.plan.world, self.sp.plan.rows_per_rank, self.sp.group
            dt2 = torch.empty_like(dz)
            for k, (rel, members) in enumerate(self.groups):
                width = len(members) * D
                lay = rel.src_layout()
                cmax, s0 = lay['cmax'], lay['start'][self.sp.plan.rank]
                _, dtab = self._exchange_buffers(k, lay['rows'], width, dz.device)
                F_.spmm_raw(rel.bwd('src'), dz[:, start:start + width], width, dtab, lay['scale'], rel.scale_local, False,
                            active_cols=cols_bm)
                dt2[:, start:start + width].copy_(dtab[R * cmax:])                       # the self-loop part stays local
                mine = raw_reduce_scatter(dtab[:R * cmax], group)                        # [cmax, width] summed over ranks
                dt2[s0:s0 + cmax, start:start + width] += mine
                start += width
            return dt2
        if self._dtab is None or self._dtab.shape != (self.sp.plan.padded, dz.shape[1]):
            self._dtab = torch.empty(self.sp.plan.padded, dz.shape[1], dtype=torch.float32, device=dz.device)
        dtab = self._dtab
        for rel, members in self.groups:
            width = len(members) * D
            F_.spmm_raw(rel.bwd('rm'), dz[:, start:start + width], width, dtab[:, start:start + width],
                        rel.scale_rm, rel.scale_local, False, active_cols=cols_bm)
            start += width
        return raw_reduce_scatter(dtab, self.sp.group)                 # ... and its only reduce-scatter


def shard_model(model, world_size=None, rank=None, group=None):
    """Switch a PEAGCN / PEASage model to row-sharded propagation.  ``model.forward`` keeps its
    signature and still returns the full [N, repr] representation in node-id order."""
    world_size = dist.get_world_size(group) if world_size is None else world_size
    rank = dist.get_rank(group) if rank is None else rank
    plan = ShardPlan(model.x.shape[0], world_size, rank)
    model._sharded = ShardedPropagation(model, plan, group)
    return model
