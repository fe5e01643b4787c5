"""PEAGNN model core - drop-in for reference ``graph_recsys_benchmark/models/base.py``:
``GraphRecsysModel`` (:29-96: loss / eval), ``PEABaseChannel`` (:129-140) and
``PEABaseRecsysModel`` (:143-214: embedding table, P channels, fusion, pair-MLP predict).

Same constructor kwargs, attributes, state_dict keys and method semantics; every tensor
operation is a libpeagnn_sm100 kernel (functional.py).  What changes underneath:
  * each distinct relation is turned into CSR/CSC once (graph.py) instead of being re-derived by
    PyG on every conv call;
  * first-step aggregations of the raw embedding table are shared between metapaths that start
    with the same relation (A_hat @ x does not depend on the channel's weights);
  * scoring + BPR loss (+ the entity-aware term) are one fused call.
"""
import torch
from torch.nn import Parameter

from ..nn.inits import glorot
from .. import functional as F_
from ..graph import get_graph


class GraphRecsysModel(torch.nn.Module):
    def __init__(self, **kwargs):
        super(GraphRecsysModel, self).__init__()
        self._init(**kwargs)
        self.reset_parameters()

    def _init(self, **kwargs):
        raise NotImplementedError

    def reset_parameters(self):
        raise NotImplementedError

    def loss(self, pos_neg_pair_t):
        """reference models/base.py:43-80 (BPR sum + entity-aware regulariser on raw x)."""
        if self.training:
            sharded = getattr(self, '_sharded', None)
            if self.demand_driven_loss and sharded is not None and sharded.supports_demand_driven():
                self.cached_repr = sharded.forward(active_ids=pos_neg_pair_t[:, :3])
            elif self.demand_driven_loss and sharded is None and self._engine_kind():
                # only the batch's user / item rows of the representation are read below (models/base.py:209-210):
                # the last step's aggregation and its transpose run on those rows only; every row that IS computed
                # equals the full propagation's.  cached_repr is then valid on the batch's rows only.
                ids = pos_neg_pair_t[:, :3]
                self.cached_repr = self.forward(_active=self._plan().active_bitmap(ids))
            elif self.demand_driven_loss and sharded is None and all(ch.supports_active() for ch in self.pea_channels):
                # per-layer path (PEAGAT): the channels' last conv aggregates the batch rows, the earlier ones the rows
                # that feeds on, and the fusion runs on the batch rows
                self.cached_repr = self.forward(_active=F_.active_rows(pos_neg_pair_t[:, :3], self.x.shape[0]))
            else:
                self.cached_repr = self.forward()
        cf_loss = F_.bpr_loss(self.cached_repr, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                              pos_neg_pair_t)
        if self.entity_aware and self.training:
            loss = cf_loss + F_.entity_reg(self.x, pos_neg_pair_t, self.entity_aware_coff)
        else:
            loss = cf_loss
        return loss

    def update_graph_input(self, dataset):
        raise NotImplementedError

    def predict(self, unids, inids):
        raise NotImplementedError

    def eval(self, metapath_idx=None):
        """reference models/base.py:88-96: nn.Module.eval() + one no-grad propagation."""
        torch.nn.Module.eval(self)
        # upstream only forwards ``metapath_idx`` to classes whose NAME starts with 'PEA' (its PEASage
        # script names the class MPASAGE..., silently disabling the ablation) - same rule here
        ablate = metapath_idx if type(self).__name__.startswith('PEA') else None
        with torch.no_grad():
            self.cached_repr = self.forward() if ablate is None else self.forward(ablate)
        return self


class PEABaseChannel(torch.nn.Module):
    def reset_parameters(self):
        for module in self.gnn_layers:
            module.reset_parameters()

    def forward(self, x, edge_index_list, shared=None):
        """relu(conv_s(x)) for every step but the last (reference models/base.py:134-140); the
        relu is fused into the conv's epilogue.  ``shared`` memoises first-step aggregations."""
        out = self.forward_split(x, edge_index_list, shared, allow_split=False)
        return out

    def supports_active(self):
        return bool(getattr(self.gnn_layers[-1], 'supports_active', False)) and self.num_steps > 1

    def forward_split(self, x, edge_index_list, shared=None, allow_split=True, active=None):
        """Runs the channel; if the last layer is of the project-then-aggregate kind and
        ``allow_split``, stops after its projection and returns ``(layer, graph, projected, layer_input)``
        so the model can aggregate all channels that share that relation in one launch."""
        assert len(edge_index_list) == self.num_steps
        n = x.size(0)
        needed = self._needed_rows(edge_index_list, n, active) if active is not None else None
        for step_idx in range(self.num_steps):
            layer = self.gnn_layers[step_idx]
            ei = edge_index_list[step_idx]
            last = step_idx == self.num_steps - 1
            g = get_graph(ei, n, keep_self_loops=getattr(layer, 'keeps_self_loops', False))
            if last and allow_split and getattr(layer, 'splits', False):
                return layer, g, layer.project(x), x
            kw = {}
            if step_idx == 0 and shared is not None and getattr(layer, 'shares_aggregate', None):
                key = (layer.shares_aggregate, id(g))
                if layer.in_channels <= layer.out_channels:
                    if key not in shared:
                        shared[key] = layer.aggregate_input(x, g)
                    kw['aggregated'] = shared[key]
            if last and active is not None and getattr(layer, 'supports_active', False):
                kw['active'] = active                       # demand-driven loss(): only the rows its batch reads
            elif not last and needed is not None and needed[step_idx] is not None:
                kw['needed'] = needed[step_idx]             # ... and of an earlier step, only the rows the later steps read
            x = layer(x, ei, relu=not last, graph=g, **kw)
        return x

    def _needed_rows(self, edge_index_list, n, active):
        """Per step but the last: the rows of its output a demand-driven loss() reads, or None (step computes all
        rows).  Step s+1 gathers the sources of its relation - node ids are contiguous by type upstream
        (datasets/movielens.py:184-227), so the [min, max] id range of its source column is that type - and its self
        loop reads the rows it produces itself: needed(s) = source range of relation s+1 OR needed(s+1), with
        needed(last) = the batch rows.  Any superset is exact; rows outside come out as zero and receive no gradient."""
        layers = self.gnn_layers
        S = self.num_steps
        if S < 2 or not getattr(layers[-1], 'supports_active', False):
            return None
        out = [None] * S
        cache = getattr(active, '_needed', None)
        if cache is None:
            cache = active._needed = {}
        statics = getattr(self, '_static_bitmaps', None)
        if statics is None:
            statics = self._static_bitmaps = {}
        ranges = []                                         # static part of the rows read of step s+1's output
        for s in range(S - 2, -1, -1):
            if not getattr(layers[s], 'supports_needed', False):
                break
            g_next = get_graph(edge_index_list[s + 1], n, keep_self_loops=getattr(layers[s + 1], 'keeps_self_loops', False))
            ranges = F_.merge_ranges(ranges + [g_next.source_range()])
            key = tuple(ranges)
            static = statics.get(key)
            if static is None:
                static = F_.range_bitmap(0, 0, n, active.device)
                for lo, hi in ranges:
                    static = torch.bitwise_or(static, F_.range_bitmap(lo, hi, n, active.device))
                statics[key] = static
            bm = cache.get(key)
            if bm is None:
                bm = cache[key] = torch.bitwise_or(static, active.bitmap)   # shared by the channels with the same later steps
            out[s] = F_.NeededRows(static, bm, ranges, active)
        return out

class PEABaseRecsysModel(GraphRecsysModel):
    def __init__(self, **kwargs):
        super(PEABaseRecsysModel, self).__init__(**kwargs)

    def _init(self, **kwargs):
        """Reads the reference's kwargs (models/base.py:148-179).  Module creation order is kept
        (x, channels, att, fc1, fc2): it fixes both the state_dict key order of the shipped
        checkpoints and the order in which the torch RNG is consumed at construction."""
        for key in ('entity_aware', 'entity_aware_coff', 'meta_path_steps', 'if_use_features', 'channel_aggr'):
            setattr(self, key, kwargs[key])
        if self.if_use_features:
            raise NotImplementedError('Feature not implemented!')
        steps, repr_dim = self.meta_path_steps, kwargs['repr_dim']
        self.x = Parameter(torch.Tensor(kwargs['dataset']['num_nodes'], kwargs['emb_dim']))

        # one COO list per (metapath, step), exactly as upstream keeps them; CSR is derived lazily
        self.meta_path_edge_index_list = self.update_graph_input(kwargs['dataset'])
        assert len(self.meta_path_edge_index_list) == len(steps)

        make_channel = kwargs['channel_class']
        self.pea_channels = torch.nn.ModuleList(make_channel(**dict(kwargs, num_steps=s)) for s in steps)
        if self.channel_aggr == 'att':
            self.att = Parameter(torch.Tensor(1, len(steps), repr_dim))
        fc1_in = 2 * repr_dim * (len(steps) if self.channel_aggr == 'cat' else 1)
        self.fc1 = torch.nn.Linear(fc1_in, repr_dim)
        self.fc2 = torch.nn.Linear(repr_dim, 1)
        self.cached_repr = None

    def reset_parameters(self):
        """glorot on x, channel re-draws, glorot on fc1 / fc2 weights (biases keep nn.Linear's
        init) and on att - the order of models/base.py:181-189."""
        glorot(self.x)
        for channel in self.pea_channels:
            channel.reset_parameters()
        for tensor in (self.fc1.weight, self.fc2.weight, getattr(self, 'att', None)):
            glorot(tensor)

    batch_last_step = True     # one aggregation per distinct last-step relation (columns concatenated)
    fused_engine = True        # engine.py: head / body autograd nodes instead of one node per kernel
    demand_driven_loss = False  # loss() computes only the representation rows its batch reads (BaseSolver turns it on)
    fuse_first_projection = False   # True (PEAGCN, demand-driven steps): relu(A_hat x W1 + b1) is computed in the EPILOGUE of the
                                    # first-step aggregation (peagnn_spmm_proj, north_star (2)) instead of by the tensor-core
                                    # projection kernels.  Parity-tested, but measured slower at every shape (ML-25M step 6.2 ->
                                    # 7.3 ms: 8 kFLOP of dependent fp32 FMAs per row in a latency-bound epilogue), so it is off.
    gather_dtype = 'fp32'       # 'bf16' (opt-in, PEAGCN engine): the first-step aggregations gather a bf16 copy of their
                                # table - fp32 accumulation and outputs, results within the tolerance of tests/test_gpu_bf16.py

    def _engine_kind(self):
        ok = getattr(self, '_engine_ok', None)
        if ok is None:
            from ..engine import GcnPlan
            ok = self._engine_ok = 'gcn' if GcnPlan.applies(self, 'gcn') else 'sage' if GcnPlan.applies(self, 'sage') else ''
        return ok if self.fused_engine else ''

    def _plan(self):
        from ..engine import GcnPlan
        kind = self._engine_kind()
        attr = '_gcn_plan' if kind == 'gcn' else '_sage_plan'
        plan = getattr(self, attr, None)
        if plan is None:
            plan = GcnPlan(self, kind)
            setattr(self, attr, plan)
        return plan

    def channel_outputs(self, active=None):
        x = self.x
        shared = {}
        outs = [module.forward_split(x, self.meta_path_edge_index_list[idx], shared, self.batch_last_step, active)
                for idx, module in enumerate(self.pea_channels)]
        groups = {}
        for idx, o in enumerate(outs):
            if isinstance(o, tuple):
                groups.setdefault(id(o[1]), []).append(idx)
        for members in groups.values():
            layer0, g = outs[members[0]][0], outs[members[0]][1]
            if len(members) == 1:
                t_cat = outs[members[0]][2]
            else:
                t_cat = torch.cat([outs[m][2] for m in members], dim=1)
            agg = layer0.batched_aggregate(t_cat, g, [outs[m][0].post_bias() for m in members], False)
            widths = [outs[m][2].shape[1] for m in members]
            parts = torch.split(agg, widths, dim=1) if len(members) > 1 else (agg,)
            for m, part in zip(members, parts):
                layer, _, _, x_in = outs[m]
                outs[m] = layer.finish(part, x_in, False)
        return outs

    def forward(self, metapath_idx=None, _active=None):
        """reference models/base.py:191-206.  'att' and 'mean' are the fusions that work upstream
        ('cat' / 'concat' disagree between _init and forward there and raise)."""
        if self.channel_aggr not in ('att', 'mean'):
            raise NotImplementedError('Other aggr methods not implemeted!')
        if getattr(self, '_sharded', None) is not None:      # distributed.shard_model(): row-sharded propagation
            return self._sharded.forward(metapath_idx)
        ok = self._engine_kind()                             # whole-model schedule for the standard PEAGCN / PEASage shape
        if ok:
            from .. import engine
            if ok == 'gcn':
                return engine.gcn_forward(self, metapath_idx, plan=self._plan(), active=_active)
            return engine.sage_forward(self, metapath_idx, active=_active)
        outs = self.channel_outputs(_active)
        att = self.att if self.channel_aggr == 'att' else None
        if _active is not None and getattr(_active, 'ids', None) is not None and metapath_idx is None:
            # demand-driven loss(): only the batch rows of the channel outputs are non-zero - fuse those [3B, P, repr]
            # rows and put the result back into an otherwise zero table instead of stacking and fusing all N rows
            z = torch.stack([F_.gather_active(o, _active) for o in outs], dim=1)
            return F_.scatter_active(F_.fuse_channels(z, att, self.channel_aggr, None), _active, self.x.shape[0])
        z = torch.stack(outs, dim=1)                                      # [N, P, repr]
        return F_.fuse_channels(z, att, self.channel_aggr, metapath_idx)

    def predict(self, unids, inids):
        """reference models/base.py:208-214: fc2(relu(fc1([repr[u] || repr[i]]))) -> [B, 1]."""
        return F_.predict_raw(self.cached_repr, unids, inids, self.fc1.weight, self.fc1.bias,
                              self.fc2.weight, self.fc2.bias)
