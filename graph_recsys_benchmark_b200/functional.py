"""Autograd bindings of the C-ABI kernels (``include/peagnn.h``).

Each ``torch.autograd.Function`` below is a thin shim: validate on the Python side (dtype,
device, contiguity - the reference raises Python exceptions), hand raw device
pointers to ``libpeagnn_sm100.so`` on the current stream, keep what backward needs.  No tensor
math happens in torch on these paths; there is no CPU path at all.
"""
import ctypes as C

import torch

from . import _lib
from .graph import _ptr, _stream, _on

NEG_SLOPE = 0.2
SHARE_FILTERED_WALK = True     # demand-driven GAT backward: per-step sub-structure of the transposed relation (Csr.filtered)


def _req(t, name, dim=2):
    if not t.is_cuda:
        raise RuntimeError('%s must live on a CUDA device (no CPU path in this package)' % name)
    if t.dtype != torch.float32:
        raise TypeError('%s must be float32, got %s' % (name, t.dtype))
    if t.dim() != dim:
        raise ValueError('%s must be %d-d' % (name, dim))
    return t


def _rows(t):
    """Row-major 2-d view with a unit inner stride and a 4-element aligned leading dimension."""
    if t.stride(1) != 1 or t.stride(0) % 4 != 0 or t.stride(0) < t.shape[1] or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


def _ws(n, device):
    return torch.empty(max(int(n), 1), dtype=torch.float32, device=device)


# ---------------------------------------------------------------------------------------------
# raw launches
# ---------------------------------------------------------------------------------------------
def spmm_algorithmic_bytes(nnz, nrows, feat, has_rs, has_cs, self_loop, accumulate=False):
    """every gathered row counted as if it came from HBM, int32 indices,
    fp32 rows: E*(4 col + 4 cs[src] + F*4) + N*(F*4 own row + 4 rs[i] + F*4 write) + (N+1)*4."""
    per_edge = 4 + (4 if has_cs else 0) + 4 * feat
    per_row = (4 * feat if self_loop else 0) + (4 if has_rs else 0) + 4 * feat + (4 * feat if accumulate else 0)
    return nnz * per_edge + nrows * per_row + (nrows + 1) * 4


def spmm_raw(csr, X, feat, out, rs=None, cs=None, self_loop=False, bias=None, relu=False, accumulate=False,
             active_rows=None, active_cols=None):
    """One aggregation launch.  ``active_rows`` / ``active_cols`` (bitmaps from ``mark_rows``) switch to the
    demand-driven entry point: only the marked rows are written / only edges gathering a marked node are read."""
    view = csr.view(feat)
    if active_rows is not None and few_rows_marked(active_rows, view.nrows):
        view = _lib.CsrView.from_buffer_copy(view)
        view.sparse_filter = 1
    tag, nbytes = None, 0
    filtered = active_rows is not None or active_cols is not None
    if _lib.profile is not None:
        nnz = csr.nnz if hasattr(csr, 'nnz') else int(csr.rowptr[-1].item() - csr.rowptr[0].item())
        tag = 'spmm%s_f%d_e%d_n%d_h%d' % ('_filtered' if filtered else '', feat, nnz, view.nrows, view.n_heavy)
        # a filtered launch still walks every index (4 B per edge) but gathers rows only for the marked part, which
        # is data dependent: its algorithmic bytes are counted as the index walk + one row write per row
        nbytes = (nnz * 4 + view.nrows * 4 * feat + (view.nrows + 1) * 4) if filtered else \
            spmm_algorithmic_bytes(nnz, view.nrows, feat, rs is not None, cs is not None, self_loop, accumulate)
    with _on(X.device):
        if filtered:
            _lib.call('peagnn_spmm_filtered', C.byref(view), _ptr(X), X.stride(0), feat, _ptr(out), out.stride(0),
                      _ptr(rs), _ptr(cs), int(self_loop), _ptr(bias), int(relu), int(accumulate),
                      _ptr(active_rows), _ptr(active_cols), _stream(), tag=tag, nbytes=nbytes)
        else:
            _lib.call('peagnn_spmm', C.byref(view), _ptr(X), X.stride(0), feat, _ptr(out), out.stride(0),
                      _ptr(rs), _ptr(cs), int(self_loop), _ptr(bias), int(relu), int(accumulate), _stream(),
                      tag=tag, nbytes=nbytes)
    return out


def spmm_proj_raw(csr, X, out, rs, cs, self_loop, projections, relu=True, active_rows=None):
    """Aggregation of a 64-wide table with up to two first projections fused into its epilogue.
    ``projections``: [(W [64, 64] (in, out), b or None, H out tensor [N, 64])]."""
    view = csr.view(64)
    tag, nbytes = None, 0
    if _lib.profile is not None:
        filtered = active_rows is not None
        tag = 'spmm%s_proj%d_f64_e%d_n%d_h%d' % ('_filtered' if filtered else '', len(projections), csr.nnz, view.nrows, view.n_heavy)
        nbytes = (csr.nnz * 4 + view.nrows * 256) if filtered else \
            spmm_algorithmic_bytes(csr.nnz, view.nrows, 64, rs is not None, cs is not None, self_loop) + len(projections) * view.nrows * 256
    (W0, b0, H0) = projections[0]
    (W1, b1, H1) = projections[1] if len(projections) > 1 else (None, None, None)
    with _on(X.device):
        _lib.call('peagnn_spmm_proj', C.byref(view), _ptr(X), X.stride(0), _ptr(out), out.stride(0), _ptr(rs), _ptr(cs),
                  int(self_loop), _ptr(active_rows), len(projections), _ptr(W0), _ptr(b0), _ptr(H0), _ptr(W1), _ptr(b1), _ptr(H1),
                  H0.stride(0), int(relu), _stream(), tag=tag, nbytes=nbytes)
    return out


def to_bf16(X):
    """bf16 copy (int16 bit patterns, same shape) of an fp32 table whose width is a multiple of 8."""
    X = _rows(_req(X, 'table'))
    out = torch.empty(X.shape[0], X.shape[1], dtype=torch.int16, device=X.device)
    with _on(X.device):
        _lib.call('peagnn_to_bf16', _ptr(X), X.stride(0), X.shape[0], X.shape[1], _ptr(out), out.stride(0), _stream())
    return out


def spmm_bf16_raw(csr, Xb, feat, out, rs=None, cs=None, self_loop=False, bias=None, relu=False, accumulate=False,
                  active_rows=None, active_cols=None):
    """spmm_raw gathering a bf16 table (``to_bf16``); fp32 accumulation and output."""
    view = csr.view(feat)
    tag, nbytes = None, 0
    if _lib.profile is not None:
        filtered = active_rows is not None or active_cols is not None
        tag = 'spmm%s_bf16_f%d_e%d_n%d_h%d' % ('_filtered' if filtered else '', feat, csr.nnz, view.nrows, view.n_heavy)
        # gathered rows are 2 bytes per element, the own row (self loop) too, the output row 4
        per_edge = 4 + (4 if cs is not None else 0) + 2 * feat
        per_row = (2 * feat if self_loop else 0) + (4 if rs is not None else 0) + 4 * feat + (4 * feat if accumulate else 0)
        nbytes = (csr.nnz * 4 + view.nrows * 4 * feat) if filtered else csr.nnz * per_edge + view.nrows * per_row + (view.nrows + 1) * 4
    with _on(out.device):
        _lib.call('peagnn_spmm_bf16', C.byref(view), _ptr(Xb), Xb.stride(0), feat, _ptr(out), out.stride(0), _ptr(rs), _ptr(cs),
                  int(self_loop), _ptr(bias), int(relu), int(accumulate), _ptr(active_rows), _ptr(active_cols), _stream(),
                  tag=tag, nbytes=nbytes)
    return out


def mark_rows(ids, n_bits, mod=0, rem=0):
    """Bitmap (int32 words) with bit ``id`` set for every id in ``ids`` (int64, any shape); with ``mod`` > 1 only ids
    with ``id % mod == rem`` count and their bit index is ``id // mod`` (the local row id under cyclic sharding)."""
    ids = ids.reshape(-1).contiguous()
    if ids.dtype != torch.long or not ids.is_cuda:
        raise TypeError('ids must be a CUDA LongTensor')
    bitmap = torch.zeros((int(n_bits) + 31) // 32 + 1, dtype=torch.int32, device=ids.device)
    with _on(ids.device):
        _lib.call('peagnn_mark_rows', _ptr(ids), ids.numel(), int(mod), int(rem), _ptr(bitmap), _stream())
    bitmap.marked_at_most = int(ids.numel())      # host-side bound: lets a launch over the marked rows pick its schedule
    return bitmap


def few_rows_marked(bitmap, n_rows):
    """True when ``bitmap`` (from ``mark_rows``) is known to mark only a few percent of ``n_rows`` rows - a training
    batch on a large graph: row-filtered launches then put one warp on each 32-row bitmap word (``sparse_filter``)."""
    m = getattr(bitmap, 'marked_at_most', None)
    return m is not None and 8 * m <= n_rows


class ActiveRows(object):
    """The rows of the final representation a loss() call reads (its batch's users and items, models/base.py:209-210):
    ``bitmap``  one bit per node (per LOCAL row on a shard) for the row / column filters of the aggregations;
    ``ids``     the same node ids sorted, duplicates kept ([3B]: a fixed size, so a step stays graph-capturable);
    ``first``   True at the first occurrence of each id - row lists that must count every node once."""

    def __init__(self, bitmap, ids=None, first=None):
        self.bitmap, self.ids, self.first = bitmap, ids, first
        self.h1_full = None      # (PEAGCN engine) first-layer activations written by the head's fused epilogue
        self.device = bitmap.device


def active_rows(ids, n_nodes):
    """ActiveRows for the node ids ``ids`` (any shape, int64, CUDA): one mark pass + one sort, no host sync."""
    flat = ids.reshape(-1).contiguous()
    srt = torch.sort(flat).values
    first = torch.ones_like(srt, dtype=torch.bool)
    first[1:] = srt[1:] != srt[:-1]
    return ActiveRows(mark_rows(flat, n_nodes), srt, first)


def merge_ranges(ranges):
    """Sorted, disjoint form of a list of [lo, hi) id ranges (overlapping / touching ranges are joined)."""
    out = []
    for lo, hi in sorted((int(lo), int(hi)) for lo, hi in ranges if hi > lo):
        if out and lo <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], hi))
        else:
            out.append((lo, hi))
    return out


class RowSets(object):
    """Rows a row-wise op of a demand-driven step has to produce: a few node-id ranges (node types, static) plus the
    batch rows (``active``: sorted ids with duplicates, a fixed-size list).  ``keep()`` marks the list entries that
    stand for a row not yet covered - first occurrence of the id, outside every range - so gradients count each
    row once."""

    def __init__(self, ranges, active):
        self.ranges, self.active = merge_ranges(ranges), active
        self.ids = active.ids

    def keep(self):
        cache = getattr(self.active, '_keep', None)
        if cache is None:
            cache = self.active._keep = {}
        key = tuple(self.ranges)
        k = cache.get(key)
        if k is None:
            k = self.active.first
            for lo, hi in self.ranges:
                k = k & ((self.ids < lo) | (self.ids >= hi))
            cache[key] = k
        return k

    def n_rows(self):
        return sum(hi - lo for lo, hi in self.ranges) + int(self.ids.numel())


class _GatherActive(torch.autograd.Function):
    """rows ``active.ids`` of a table whose gradient is wanted once per node: forward = index_select, backward = the
    list gradient added back into a zero table (later occurrences of a node carry exact zeros upstream)."""

    @staticmethod
    def forward(ctx, table, active):
        ctx.active, ctx.n = active, table.shape[0]
        return table.index_select(0, active.ids)

    @staticmethod
    def backward(ctx, d):
        out = torch.zeros((ctx.n,) + tuple(d.shape[1:]), dtype=d.dtype, device=d.device)
        out.index_add_(0, ctx.active.ids, d)
        return out, None


class _ScatterActive(torch.autograd.Function):
    """[3B, D] rows back into a zero [N, D] table (duplicates rewrite the same row); the gradient of every node is taken
    once, at its first occurrence on the list."""

    @staticmethod
    def forward(ctx, rows, active, n):
        ctx.active = active
        out = torch.zeros((n,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        out.index_copy_(0, active.ids, rows)
        return out

    @staticmethod
    def backward(ctx, dout):
        a = ctx.active
        return dout.index_select(0, a.ids) * a.first[:, None].to(dout.dtype), None, None


def gather_active(table, active):
    return _GatherActive.apply(table, active)


def scatter_active(rows, active, n):
    return _ScatterActive.apply(rows, active, n)


def column_sum_where_nonzero(dout, active=None, needed=None):
    """Column sums of a gradient table that is known to be zero outside a few rows: the batch rows (``active``, a
    last step) or a node-id range plus the batch rows (``needed``, an earlier step) - the conv's bias gradient without
    streaming all N rows.  Falls back to the whole table when neither is given."""
    M = dout.shape[1]
    out = torch.empty(M, dtype=torch.float32, device=dout.device)
    if active is not None and active.ids is not None:
        rows = dout.index_select(0, active.ids) * active.first[:, None].to(dout.dtype)
        wgrad_raw(None, rows, 0, M, 0, None, out)
        return out
    if needed is not None and needed.ranges is not None and needed.active is not None and needed.active.ids is not None:
        sets = RowSets(needed.ranges, needed.active)
        rows = dout.index_select(0, sets.ids) * sets.keep()[:, None].to(dout.dtype)
        wgrad_raw(None, rows, 0, M, 0, None, out)
        for lo, hi in sets.ranges:
            part = torch.empty_like(out)
            wgrad_raw(None, dout[lo:hi], 0, M, 0, None, part)
            out.add_(part)
        return out
    wgrad_raw(None, dout, 0, M, 0, None, out)
    return out


def range_bitmap(lo, hi, n_bits, device):
    """``mark_rows`` layout with the bits lo <= id < hi set (a node type's id range)."""
    words = (int(n_bits) + 31) // 32 + 1
    bit = torch.arange(words * 32, device=device)
    on = (bit >= lo) & (bit < hi)
    w = (on.view(words, 32).to(torch.int64) << torch.arange(32, device=device)).sum(dim=1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).contiguous()


def linear_algorithmic_bytes(n, K, M, accumulate=False, gated=False):
    """X once, Y once, W once (+ Y read when accumulating, + the gate's rows)."""
    return 4 * n * (K + M) + 4 * K * M + (4 * n * M if accumulate else 0) + (4 * n * M if gated else 0)


def linear_raw(X, W, out, w_is_out_in, bias=None, relu=False, accumulate=False, mask=None, out_mask=None):
    n, K = X.shape
    M = out.shape[1]
    tag, nbytes = None, 0
    if _lib.profile is not None:
        tag = 'linear_%dto%d' % (K, M)
        nbytes = linear_algorithmic_bytes(n, K, M, accumulate, mask is not None or out_mask is not None)
    with _on(X.device):
        _lib.call('peagnn_linear', _ptr(X), X.stride(0), _ptr(mask), mask.stride(0) if mask is not None else 0,
                  n, K, M, _ptr(W), int(w_is_out_in), _ptr(bias), int(relu), int(accumulate),
                  _ptr(out), out.stride(0), _ptr(out_mask), out_mask.stride(0) if out_mask is not None else 0,
                  _stream(), tag=tag, nbytes=nbytes)
    return out


def wgrad_raw(X, dY, K, M, w_is_out_in, dW, db, mask=None):
    n = dY.shape[0]
    need = int(_lib.query('peagnn_wgrad_workspace_floats', n, K, M))
    ws = _ws(need, dY.device)
    tag, nbytes = None, 0
    if _lib.profile is not None:
        tag = 'wgrad_%dx%d' % (K, M)
        nbytes = 4 * n * (K + M) + 4 * K * M + (4 * n * M if mask is not None else 0)
    with _on(dY.device):
        _lib.call('peagnn_linear_wgrad', _ptr(X), X.stride(0) if X is not None else 0, _ptr(dY), dY.stride(0),
                  _ptr(mask), mask.stride(0) if mask is not None else 0, n, K, M, int(w_is_out_in),
                  _ptr(dW), _ptr(db), _ptr(ws), need, _stream(), tag=tag, nbytes=nbytes)


def linear_grouped_raw(problems, K, M, w_is_out_in, relu=False, accumulate=False):
    """``peagnn_linear_grouped``: the projections ``(X, W, bias, out, out_mask)`` of ONE shape as a single launch (the
    per-metapath projections of a step: many small problems).  Outputs must not overlap."""
    problems = [q for q in problems if q[0].shape[0] > 0]
    if not problems:
        return
    arr = (_lib.LinearProblem * len(problems))()
    nbytes = 0
    for k, (X, W, bias, out, out_mask) in enumerate(problems):
        assert X.shape[1] == K and out.shape[1] == M and out.shape[0] == X.shape[0]
        a = arr[k]
        a.X, a.ldx, a.n, a.W, a.bias = X.data_ptr(), X.stride(0), X.shape[0], W.data_ptr(), (bias.data_ptr() if bias is not None else None)
        a.Y, a.ldy = out.data_ptr(), out.stride(0)
        a.out_mask, a.ldom = (out_mask.data_ptr(), out_mask.stride(0)) if out_mask is not None else (None, 0)
        nbytes += linear_algorithmic_bytes(X.shape[0], K, M, accumulate, out_mask is not None)
    dev = problems[0][0].device
    with _on(dev):
        _lib.call('peagnn_linear_grouped', C.cast(arr, C.c_void_p), len(problems), K, M, int(w_is_out_in), int(relu),
                  int(accumulate), _stream(), tag='linear_%dto%d_grouped' % (K, M) if _lib.profile is not None else None,
                  nbytes=nbytes)


def wgrad_grouped_raw(problems, K, M, w_is_out_in):
    """``peagnn_linear_wgrad_grouped``: the weight (and bias) gradients ``(X, dY, dW, db)`` of ONE shape as two launches.
    A problem without rows gets zeros."""
    if not problems:
        return
    arr = (_lib.WgradProblem * len(problems))()
    nbytes = 0
    for k, (X, dY, dW, db) in enumerate(problems):
        assert X.shape[1] == K and dY.shape[1] == M and X.shape[0] == dY.shape[0]
        a = arr[k]
        a.X, a.ldx, a.dY, a.ldd, a.n = X.data_ptr(), X.stride(0), dY.data_ptr(), dY.stride(0), X.shape[0]
        a.dW, a.db = (dW.data_ptr() if dW is not None else None), (db.data_ptr() if db is not None else None)
        nbytes += 4 * X.shape[0] * (K + M) + 4 * K * M
    dev = problems[0][0].device
    need = int(_lib.query('peagnn_wgrad_grouped_workspace_floats', len(problems), K, M))
    ws = _ws(need, dev)
    with _on(dev):
        _lib.call('peagnn_linear_wgrad_grouped', C.cast(arr, C.c_void_p), len(problems), K, M, int(w_is_out_in), _ptr(ws),
                  need, _stream(), tag='wgrad_%dx%d_grouped' % (K, M) if _lib.profile is not None else None, nbytes=nbytes)


def relu_backward_raw(dy, act):
    out = torch.empty_like(act)
    with _on(dy.device):
        _lib.call('peagnn_relu_backward', _ptr(dy), dy.stride(0), _ptr(act), act.stride(0), dy.shape[0],
                  dy.shape[1], _ptr(out), out.stride(0), _stream())
    return out


# ---------------------------------------------------------------------------------------------
# K1/K2  aggregation
# ---------------------------------------------------------------------------------------------
class _Aggregate(torch.autograd.Function):
    """out = rs * (A (cs * X) [+ cs_i X_i]) [+ bias] [relu]   over graph.fwd; backward over graph.bwd
    with rs / cs swapped (the transpose of diag(rs) A diag(cs) is diag(cs) A^T diag(rs))."""

    @staticmethod
    def forward(ctx, X, bias, graph, rs, cs, self_loop, relu):
        X = _rows(_req(X, 'x'))
        feat = X.shape[1]
        out = torch.empty(graph.num_nodes, feat, dtype=torch.float32, device=X.device)
        spmm_raw(graph.fwd, X, feat, out, rs, cs, self_loop, bias, relu)
        ctx.graph, ctx.rs, ctx.cs, ctx.self_loop, ctx.relu = graph, rs, cs, self_loop, relu
        ctx.has_bias = bias is not None
        if relu:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = _rows(dout)
        if ctx.relu:
            (out,) = ctx.saved_tensors
            dout = relu_backward_raw(dout, out)
        feat = dout.shape[1]
        dX = db = None
        if ctx.needs_input_grad[0]:
            dX = torch.empty_like(dout)
            spmm_raw(ctx.graph.bwd, dout, feat, dX, ctx.cs, ctx.rs, ctx.self_loop)
        if ctx.has_bias and ctx.needs_input_grad[1]:
            db = torch.empty(feat, dtype=torch.float32, device=dout.device)
            wgrad_raw(None, dout, 0, feat, 0, None, db)
        return dX, db, None, None, None, None, None


def gcn_aggregate(X, graph, bias=None, relu=False):
    """D^-1/2 (A + I) D^-1/2 X (+ bias)(relu)  with PyG-1.5.0's source-side degree."""
    dis = graph.gcn_dis
    return _Aggregate.apply(X, bias, graph, dis, dis, True, relu)


def sage_mean_aggregate(X, graph, bias=None):
    """mean_{j -> i} X_j (+ bias)  (0 for rows without in-edges; no self loops)."""
    return _Aggregate.apply(X, bias, graph, graph.inv_in_degree, None, False, False)


# ---------------------------------------------------------------------------------------------
# K4  projections
# ---------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    """Y = act(X @ W (+ b)) with W either [in, out] (GCNConv.weight) or [out, in] (nn.Linear)."""

    @staticmethod
    def forward(ctx, X, W, bias, w_is_out_in, relu):
        X = _rows(_req(X, 'x'))
        W = _req(W, 'weight').contiguous()
        K = X.shape[1]
        M = W.shape[0] if w_is_out_in else W.shape[1]
        if (W.shape[1] if w_is_out_in else W.shape[0]) != K:
            raise ValueError('weight shape %s does not match input width %d' % (tuple(W.shape), K))
        Y = torch.empty(X.shape[0], M, dtype=torch.float32, device=X.device)
        linear_raw(X, W, Y, w_is_out_in, bias, relu)
        ctx.w_is_out_in, ctx.relu, ctx.has_bias = w_is_out_in, relu, bias is not None
        ctx.save_for_backward(X, W, Y if relu else None)
        return Y

    @staticmethod
    def backward(ctx, dY):
        X, W, Y = ctx.saved_tensors
        dY = _rows(dY)
        K, M = X.shape[1], dY.shape[1]
        dX = dW = db = None
        if ctx.needs_input_grad[0]:
            dX = torch.empty_like(X)
            # dX = gate(dY) @ W^T : same kernel with the opposite weight layout
            linear_raw(dY, W, dX, not ctx.w_is_out_in, None, False, False, mask=Y)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dW = torch.empty_like(W)
            db = torch.empty(M, dtype=torch.float32, device=dY.device) if ctx.has_bias else None
            wgrad_raw(X, dY, K, M, ctx.w_is_out_in, dW, db, mask=Y)
        return dX, dW, db, None, None


def linear(X, W, bias=None, w_is_out_in=True, relu=False):
    return _Linear.apply(X, W, bias, w_is_out_in, relu)


class _LinearAccumulate(torch.autograd.Function):
    """Y = act(base + X @ W^T), W [out, in]; written in place over ``base`` (SAGEConv's
    ``lin_rel(mean) + lin_root(x)``: the second projection accumulates onto the first)."""

    @staticmethod
    def forward(ctx, X, W, base, relu):
        X = _rows(_req(X, 'x'))
        W = _req(W, 'weight').contiguous()
        if not base.is_contiguous():
            raise ValueError('accumulation target must be contiguous')
        linear_raw(X, W, base, True, None, relu, True)
        ctx.mark_dirty(base)
        ctx.relu = relu
        ctx.save_for_backward(X, W, base if relu else None)
        return base

    @staticmethod
    def backward(ctx, dY):
        X, W, Y = ctx.saved_tensors
        dY = _rows(dY)
        if ctx.relu:
            dY = relu_backward_raw(dY, Y)
        K, M = X.shape[1], dY.shape[1]
        dX = dW = None
        if ctx.needs_input_grad[0]:
            dX = torch.empty_like(X)
            linear_raw(dY, W, dX, False)
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(W)
            wgrad_raw(X, dY, K, M, True, dW, None)
        return dX, dW, dY, None


def linear_accumulate(X, W, base, relu=False):
    return _LinearAccumulate.apply(X, W, base, relu)


# ---------------------------------------------------------------------------------------------
# K3  GAT
# ---------------------------------------------------------------------------------------------
class _GatScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, att_i, att_j, heads):
        H = _rows(_req(H, 'h'))
        n = H.shape[0]
        feat = H.shape[1] // heads
        ai = torch.empty(n, heads, dtype=torch.float32, device=H.device)
        aj = torch.empty_like(ai)
        att_i, att_j = att_i.contiguous(), att_j.contiguous()
        with _on(H.device):
            _lib.call('peagnn_gat_scores', _ptr(H), H.stride(0), n, feat, heads, _ptr(att_i), _ptr(att_j),
                      _ptr(ai), _ptr(aj), _stream())
        ctx.heads = heads
        ctx.save_for_backward(H, att_i, att_j)
        return ai, aj

    @staticmethod
    def backward(ctx, d_ai, d_aj):
        H, att_i, att_j = ctx.saved_tensors
        heads = ctx.heads
        n = H.shape[0]
        feat = H.shape[1] // heads
        d_ai, d_aj = d_ai.contiguous(), d_aj.contiguous()
        dH = torch.empty_like(H)
        d_att_i, d_att_j = torch.empty_like(att_i), torch.empty_like(att_j)
        need = int(_lib.query('peagnn_wgrad_workspace_floats', n, heads * feat, 4))
        ws = _ws(need, H.device)
        with _on(H.device):
            _lib.call('peagnn_gat_scores_backward', _ptr(H), H.stride(0), n, feat, heads, _ptr(att_i), _ptr(att_j),
                      _ptr(d_ai), _ptr(d_aj), _ptr(dH), dH.stride(0), 0, _ptr(d_att_i), _ptr(d_att_j),
                      _ptr(ws), need, _stream())
        return dH, d_att_i, d_att_j, None


def _filtered_view(view, active_rows=None, active_cols=None):
    """A copy of a ``peagnn_csr_t`` carrying the demand-driven filters (bitmaps from ``mark_rows``)."""
    v = _lib.CsrView.from_buffer_copy(view)
    v.active_rows = active_rows.data_ptr() if active_rows is not None else None
    v.active_cols = active_cols.data_ptr() if active_cols is not None else None
    v.sparse_filter = int(active_rows is not None and few_rows_marked(active_rows, v.nrows))
    return v


class NeededRows(object):
    """Rows of an intermediate step's output that the rest of a demand-driven loss() reads: ``static`` (a bitmap fixed
    by the metapath: the node-id range of the next relation's sources) OR the step's batch rows -> ``bitmap``."""

    def __init__(self, static, bitmap, ranges=None, active=None):
        self.static, self.bitmap = static, bitmap
        self.ranges, self.active = ranges, active    # the static part as disjoint id ranges [(lo, hi)]; the batch rows

    def covers(self, csr):
        """True when every row of ``csr`` that has an edge is marked by the static part (decided once per structure,
        outside any capture): a row-filtered pass then still writes every per-edge slot."""
        cache = getattr(self.static, '_covers', None)
        if cache is None:
            cache = self.static._covers = {}
        hit = cache.get(id(csr))
        if hit is None:
            has = csr.nonempty_row_bitmap()
            hit = cache[id(csr)] = bool((torch.bitwise_and(has, torch.bitwise_not(self.static)) == 0).all().item())
        return hit


class _GatAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, ai, aj, bias, graph, heads, relu, active=None, needed=None):
        """Demand-driven forms (bitmaps from ``mark_rows``); the rows that are not marked come out as zero:
        ``active``  a channel's LAST step - the few rows a batch reads: the target-side passes visit those rows, the
                    source-side pass walks a per-step sub-structure holding only the edges into them;
        ``needed``  an EARLIER step - the rows the next step reads (its relation's source type plus the batch rows):
                    nearly every edge survives, what is skipped is the per-row work on the other node types."""
        H = _rows(_req(H, 'h'))
        n = graph.num_nodes
        feat = H.shape[1] // heads
        dev = H.device
        ai, aj = ai.contiguous(), aj.contiguous()
        active_bm = active.bitmap if isinstance(active, ActiveRows) else active
        rows = active_bm if active is not None else (needed.bitmap if needed is not None else None)
        alloc = torch.zeros if rows is not None else torch.empty
        rowmax = alloc(n, heads, dtype=torch.float32, device=dev)
        denom = alloc(n, heads, dtype=torch.float32, device=dev)
        out = alloc(n, heads * feat, dtype=torch.float32, device=dev)
        view = graph.fwd.view(feat, heads)
        if rows is not None:
            view = _filtered_view(view, active_rows=rows)
        # E*(4 col + 4 a_j[src] + F*4) + N*(F*4 own row + 8 + F*4 write) + (N+1)*4, per head
        nnz = graph.fwd.nnz
        with _on(dev):
            _lib.call('peagnn_gat_rowmax', C.byref(view), _ptr(ai), _ptr(aj), heads, NEG_SLOPE, _ptr(rowmax), _stream())
            _lib.call('peagnn_gat_aggregate', C.byref(view), _ptr(H), H.stride(0), feat, heads, _ptr(ai), _ptr(aj),
                      NEG_SLOPE, _ptr(rowmax), _ptr(denom), _ptr(out), out.stride(0), _ptr(bias), int(relu), _stream(),
                      tag='gat_agg_f%d_e%d_n%d' % (feat, nnz, n),
                      nbytes=heads * (nnz * (8 + 4 * feat) + n * (8 * feat + 8) + (n + 1) * 4))
        ctx.graph, ctx.heads, ctx.relu, ctx.has_bias, ctx.active, ctx.needed = graph, heads, relu, bias is not None, active, needed
        ctx.save_for_backward(H, ai, aj, rowmax, denom, out, bias)
        return out

    @staticmethod
    def backward(ctx, dout):
        H, ai, aj, rowmax, denom, out, bias = ctx.saved_tensors
        graph, heads, active_obj, needed = ctx.graph, ctx.heads, ctx.active, ctx.needed
        active = active_obj.bitmap if isinstance(active_obj, ActiveRows) else active_obj
        n = graph.num_nodes
        feat = H.shape[1] // heads
        dev = H.device
        dout = _rows(dout)
        if ctx.relu:
            dout = relu_backward_raw(dout, out)
        # the kernel needs <dout, aggregate-before-bias>; it subtracts the bias itself, and where a
        # relu clamped the output dout is already 0, so passing the forward output is exact.
        db = None
        if ctx.has_bias and ctx.needs_input_grad[3]:
            # dout is zero outside the rows this step computed: sum only where it can be anything else
            db = column_sum_where_nonzero(dout, active_obj if isinstance(active_obj, ActiveRows) else None,
                                          needed if active_obj is None else None)
        nnz = graph.fwd.nnz
        rows = active if active is not None else (needed.bitmap if needed is not None else None)
        # (alpha, ds) per edge and head, interleaved.  `needed`: the target-side pass skips rows, the source-side pass
        # walks every edge - an edge into a skipped row has to read as (0, 0) (no such edge when the static part of
        # the filter covers every row that has one: the usual case, the next relation's sources ARE this one's targets)
        unwritten = needed is not None and active is None and not needed.covers(graph.fwd)
        ads_e = (torch.zeros if unwritten else torch.empty)(max(nnz, 1), heads, 2, dtype=torch.float32, device=dev)
        # demand-driven: the target-side pass visits the marked rows only (dout is zero elsewhere), the source-side
        # pass skips every edge into another target; per-node outputs of skipped rows must read as zero
        alloc = torch.zeros if rows is not None else torch.empty
        alpha_s = alloc(n, heads, dtype=torch.float32, device=dev)
        ds_s = alloc(n, heads, dtype=torch.float32, device=dev)
        d_ai = alloc(n, heads, dtype=torch.float32, device=dev)
        d_aj = alloc(n, heads, dtype=torch.float32, device=dev)
        dH = torch.zeros_like(H) if rows is not None else torch.empty_like(H)
        vf = graph.fwd.view(feat, heads)
        vb = graph.bwd.view(feat, heads)
        perm = graph.bwd_to_fwd
        if active is not None:
            vf = _filtered_view(vf, active_rows=active)
            if SHARE_FILTERED_WALK and not graph.bwd.explicit_self_loops and hasattr(graph.bwd, 'filtered'):
                # one pass over the relation's index array per step (shared by every metapath that ends with it)
                # instead of one column-filtered walk per metapath
                sub = graph.bwd.filtered(active, perm)
                # ... and only the rows that can receive a gradient are visited: the relation's sources (edges) and the
                # batch rows (self loop); d H and d a_j are zero elsewhere
                rows_bm = torch.bitwise_or(graph.bwd.nonempty_row_bitmap(), active)
                vb, perm = _filtered_view(sub.view(feat, heads), active_rows=rows_bm), sub.perm
            else:
                vb = _filtered_view(vb, active_cols=active)
        elif needed is not None:
            vf = _filtered_view(vf, active_rows=needed.bitmap)
            # rows that can receive a gradient: the relation's sources (edges) and the needed rows (self loop)
            vb = _filtered_view(vb, active_rows=torch.bitwise_or(graph.bwd.nonempty_row_bitmap(), needed.bitmap))
        with _on(dev):
            kind = ('_filtered' if active is not None else '_needed' if needed is not None else '') + '_f%d_e%d' % (feat, nnz)
            _lib.call('peagnn_gat_backward_dst', C.byref(vf), _ptr(H), H.stride(0), feat, heads, _ptr(ai), _ptr(aj),
                      NEG_SLOPE, _ptr(rowmax), _ptr(denom), _ptr(out), out.stride(0), _ptr(bias), _ptr(dout), dout.stride(0),
                      _ptr(ads_e), _ptr(alpha_s), _ptr(ds_s), _ptr(d_ai), _stream(),
                      tag='gat_backward_dst' + kind if _lib.profile is not None else None)
            _lib.call('peagnn_gat_backward_src', C.byref(vb), _ptr(perm), _ptr(ads_e), _ptr(alpha_s),
                      _ptr(ds_s), _ptr(dout), dout.stride(0), feat, heads, _ptr(dH), dH.stride(0), _ptr(d_aj),
                      _stream(), tag='gat_backward_src' + kind if _lib.profile is not None else None)
        return dH, d_ai, d_aj, db, None, None, None, None, None


def gat_scores(H, att_i, att_j, heads):
    return _GatScores.apply(H, att_i, att_j, heads)


def gat_aggregate(H, ai, aj, graph, heads, bias=None, relu=False, active=None, needed=None):
    return _GatAggregate.apply(H, ai, aj, bias, graph, heads, relu, active, needed)


# ---------------------------------------------------------------------------------------------
# K5  fusion across metapaths
# ---------------------------------------------------------------------------------------------
class _Fuse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Z, att, mode, skip):
        # Z: [N, P, D] contiguous
        Z = _req(Z, 'channel outputs', 3).contiguous()
        n, P, D = Z.shape
        out = torch.empty(n, D, dtype=torch.float32, device=Z.device)
        att2 = att.reshape(P, D).contiguous() if att is not None else None
        with _on(Z.device):
            _lib.call('peagnn_fuse_forward', _ptr(Z), P * D, n, P, D, _ptr(att2), mode, skip, _ptr(out), D, _stream())
        ctx.mode, ctx.skip = mode, skip
        ctx.att_shape = att.shape if att is not None else None
        ctx.save_for_backward(Z, att2)
        return out

    @staticmethod
    def backward(ctx, dout):
        Z, att2 = ctx.saved_tensors
        if ctx.skip >= 0:
            raise RuntimeError('metapath ablation (metapath_idx) is an evaluation-only path (models/base.py:88-96)')
        n, P, D = Z.shape
        dout = _rows(dout)
        dZ = torch.empty_like(Z)
        d_att = torch.empty(P, D, dtype=torch.float32, device=Z.device) if ctx.mode == 0 else None
        need = int(_lib.query('peagnn_fuse_workspace_floats', n, P, D)) if ctx.mode == 0 else 0
        ws = _ws(need, Z.device) if ctx.mode == 0 else None
        with _on(Z.device):
            _lib.call('peagnn_fuse_backward', _ptr(Z), P * D, n, P, D, _ptr(att2), ctx.mode, _ptr(dout), dout.stride(0),
                      _ptr(dZ), P * D, _ptr(d_att), _ptr(ws), need, _stream())
        return dZ, (d_att.reshape(ctx.att_shape) if d_att is not None else None), None, None


def fuse_channels(Z, att, mode='att', skip=None):
    return _Fuse.apply(Z, att, 0 if mode == 'att' else 1, -1 if skip is None else int(skip))


# ---------------------------------------------------------------------------------------------
# K6  scoring + BPR (+ entity-aware term)
# ---------------------------------------------------------------------------------------------
def predict_raw(repr_, unids, inids, fc1_w, fc1_b, fc2_w, fc2_b):
    repr_ = _rows(_req(repr_, 'cached_repr'))
    unids, inids = unids.contiguous(), inids.contiguous()
    if unids.dtype != torch.long or inids.dtype != torch.long:
        raise TypeError('node ids must be int64 (torch.long)')
    B = int(unids.numel())
    out = torch.empty(B, 1, dtype=torch.float32, device=repr_.device)
    with _on(repr_.device):
        _lib.call('peagnn_predict', _ptr(repr_), repr_.stride(0), repr_.shape[1], _ptr(unids), _ptr(inids), B,
                  _ptr(fc1_w.contiguous()), _ptr(fc1_b.contiguous()), _ptr(fc2_w.contiguous()), _ptr(fc2_b.contiguous()),
                  _ptr(out), _stream())
    return out


class _Predict(torch.autograd.Function):
    """predict() outside loss(): forward-only on the kernel path.  (Training goes through
    _BprLoss, which fuses scoring and loss; upstream never differentiates predict() directly.)"""

    @staticmethod
    def forward(ctx, repr_, unids, inids, fc1_w, fc1_b, fc2_w, fc2_b):
        return predict_raw(repr_, unids, inids, fc1_w, fc1_b, fc2_w, fc2_b)

    @staticmethod
    def backward(ctx, g):
        raise RuntimeError('predict() is forward-only here; differentiate loss() instead')


class _BprLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, repr_, fc1_w, fc1_b, fc2_w, fc2_b, batch):
        repr_ = _rows(_req(repr_, 'cached_repr'))
        batch = batch.contiguous()
        if batch.dtype != torch.long or batch.dim() != 2 or batch.shape[1] < 3:
            raise ValueError('batch must be a LongTensor [B, >=3]')
        B, cols = int(batch.shape[0]), int(batch.shape[1])
        if repr_.shape[0] >= 2 ** 31 - 1:
            raise ValueError('node ids must fit int32 (the gradient scatter sorts 31-bit keys)')
        D = repr_.shape[1]
        dev = repr_.device
        need_grad = any(ctx.needs_input_grad[:5])
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        need = int(_lib.query('peagnn_bpr_workspace_floats', B, D))
        ws = _ws(need, dev)
        fc1_w, fc1_b, fc2_w, fc2_b = (t.contiguous() for t in (fc1_w, fc1_b, fc2_w, fc2_b))
        if need_grad:
            d_repr = torch.zeros_like(repr_)
            g1w, g1b, g2w, g2b = (torch.empty_like(t) for t in (fc1_w, fc1_b, fc2_w, fc2_b))
        else:
            d_repr = g1w = g1b = g2w = g2b = None
        with _on(dev):
            _lib.call('peagnn_bpr_loss', _ptr(repr_), repr_.stride(0), D, _ptr(batch), cols, B,
                      _ptr(fc1_w), _ptr(fc1_b), _ptr(fc2_w), _ptr(fc2_b), _ptr(loss), int(need_grad),
                      _ptr(d_repr), d_repr.stride(0) if need_grad else 0, _ptr(g1w), _ptr(g1b), _ptr(g2w), _ptr(g2b),
                      _ptr(ws), need, _stream())
        if need_grad:
            ctx.save_for_backward(d_repr, g1w, g1b, g2w, g2b)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        d_repr, g1w, g1b, g2w, g2b = ctx.saved_tensors
        return d_repr * g, g1w * g, g1b * g, g2w * g, g2b * g, None


class _EntityReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, batch, coff):
        x = _rows(_req(x, 'x'))
        batch = batch.contiguous()
        if batch.dtype != torch.long or batch.dim() != 2 or batch.shape[1] != 9:
            raise ValueError('entity-aware batches are LongTensor [B, 9] (datasets/movielens.py:1179)')
        B = int(batch.shape[0])
        dev = x.device
        need_grad = ctx.needs_input_grad[0]
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        dx = torch.zeros_like(x) if need_grad else None
        need = int(_lib.query('peagnn_entity_workspace_floats', B, x.shape[1], int(need_grad)))
        ws = _ws(need, dev)
        with _on(dev):
            _lib.call('peagnn_entity_reg', _ptr(x), x.stride(0), x.shape[1], _ptr(batch), B, float(coff), _ptr(loss),
                      int(need_grad), _ptr(dx), dx.stride(0) if need_grad else 0, _ptr(ws), need, _stream())
        if need_grad:
            ctx.save_for_backward(dx)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None, None


def bpr_loss(repr_, fc1_w, fc1_b, fc2_w, fc2_b, batch):
    return _BprLoss.apply(repr_, fc1_w, fc1_b, fc2_w, fc2_b, batch)


def entity_reg(x, batch, coff):
    """coff * (item_reg + user_reg) of models/base.py:50-76 (already scaled by coff)."""
    return _EntityReg.apply(x, batch, coff)


# ---------------------------------------------------------------------------------------------
# K7  evaluation
# ---------------------------------------------------------------------------------------------
def eval_rank(repr_, users, cand, n_pos, fc1_w, fc1_b, fc2_w, fc2_b, return_scores=False):
    """per-user fp64 [U, 36] rows (HR[16] | NDCG[16] | AUC | loss | first-hit rank | 0) and their
    column means, all on the device; optionally the raw scores [U, C]."""
    repr_ = _rows(_req(repr_, 'cached_repr'))
    users, cand = users.contiguous(), cand.contiguous()
    U, Cn = int(cand.shape[0]), int(cand.shape[1])
    dev = repr_.device
    per_user = torch.empty(U, 36, dtype=torch.float64, device=dev)
    scores = torch.empty(U, Cn, dtype=torch.float32, device=dev) if return_scores else None
    means = torch.empty(36, dtype=torch.float64, device=dev)
    ws = torch.empty(148 * 36, dtype=torch.float64, device=dev)
    with _on(dev):
        _lib.call('peagnn_eval_rank', _ptr(repr_), repr_.stride(0), repr_.shape[1], _ptr(users), _ptr(cand), U, Cn,
                  int(n_pos), _ptr(fc1_w.contiguous()), _ptr(fc1_b.contiguous()), _ptr(fc2_w.contiguous()),
                  _ptr(fc2_b.contiguous()), _ptr(per_user), _ptr(scores), _stream())
        _lib.call('peagnn_column_mean', _ptr(per_user), 36, U, 36, _ptr(means), _ptr(ws), ws.numel(), _stream())
    return per_user, means, scores
