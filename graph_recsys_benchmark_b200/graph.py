"""Device-resident graph structures for the metapath steps.

The reference hands every conv a COO ``edge_index`` (LongTensor [2, E], row 0 = source, row 1 =
target; utils/general_utils.py:280-395) and lets PyG re-derive self loops, degrees and the
scatter layout on every call.  Here each distinct relation is converted ONCE into

* ``fwd``  - CSR grouped by target (gather side of the forward pass),
* ``bwd``  - CSR grouped by source (the transposed structure the backward pass gathers over),

both int32, both stable in COO order (multi-edges kept - ``tag2item`` / ``tag2user`` carry
duplicates on purpose, datasets/movielens.py:282-286), self-loop edges dropped (GCNConv /
GATConv re-add exactly one per node; none of the PEAGNN relations has any).  ``flip(ei)`` of a
known relation is recognised and served by swapping the two structures.
"""
import ctypes as C

import torch

from . import _lib

HEAVY_THRESHOLD = 256      # rows with more edges than this are cut into chunks (a single warp walking a longer
                           # row becomes the critical path of the launch) ...
CHUNK_EDGES = 4096         # ... of at most this many edges, one CTA each
MIN_CHUNK_EDGES = 256
SPARSE_HEAVY_THRESHOLD = 96


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Here(object):
    """No-op context: the tensor's device is already current (the common case - switching the
    device around every launch costs more host time than the launch itself)."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_HERE = _Here()


def _on(device):
    """Context that makes ``device`` current for a launch."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _HERE
    return torch.cuda.device(device)


class Csr(object):
    """One CSR structure (rows gather from ``col``) plus its heavy-row work list."""

    def __init__(self, rowptr, col, eid, num_nodes, heavy_threshold=None, chunk_edges=None):
        self.rowptr, self.col, self.eid = rowptr, col, eid
        self.num_nodes = num_nodes
        self.nnz = int(col.numel())
        if heavy_threshold is None:      # sparse views pack 8 rows per warp: keep those rows short
            heavy_threshold = min(HEAVY_THRESHOLD, SPARSE_HEAVY_THRESHOLD) if self.nnz < 8 * num_nodes else HEAVY_THRESHOLD
        self.heavy_threshold = heavy_threshold
        self.chunk_edges = CHUNK_EDGES if chunk_edges is None else chunk_edges
        self.explicit_self_loops = False     # set by row shards that store self loops as edges
        self._build_work_list()
        self._partial = None
        self._views = {}

    def _build_work_list(self):
        dev = self.rowptr.device
        rp = self.rowptr.long()
        deg = rp[1:] - rp[:-1]
        heavy = torch.nonzero(deg > self.heavy_threshold).flatten()
        self.n_heavy = int(heavy.numel())
        if self.n_heavy == 0:
            self.heavy_rows = self.heavy_chunk_ptr = None
            self.chunk_row = self.chunk_begin = self.chunk_end = None
            self.n_chunks = 0
            return
        hdeg = deg[heavy]
        # few very long rows (a genre / year / tag node as a target): shrink the chunks until there
        # are about two CTAs per SM, so the row does not sit on a handful of SMs
        heavy_edges = int(hdeg.sum().item())
        fill = max(MIN_CHUNK_EDGES, (heavy_edges // (2 * 148) + 31) // 32 * 32)
        self.chunk_edges = min(self.chunk_edges, fill)
        nch = (hdeg + self.chunk_edges - 1) // self.chunk_edges
        cptr = torch.zeros(self.n_heavy + 1, dtype=torch.long, device=dev)
        cptr[1:] = torch.cumsum(nch, 0)
        self.n_chunks = int(cptr[-1].item())
        owner = torch.repeat_interleave(torch.arange(self.n_heavy, device=dev), nch)
        within = torch.arange(self.n_chunks, device=dev) - cptr[owner]
        begin = rp[heavy][owner] + within * self.chunk_edges
        end = torch.minimum(begin + self.chunk_edges, rp[heavy + 1][owner])
        self.heavy_rows = heavy.int().contiguous()
        self.heavy_chunk_ptr = cptr.int().contiguous()
        self.chunk_row = heavy[owner].int().contiguous()
        self.chunk_begin = begin.int().contiguous()
        self.chunk_end = end.int().contiguous()

    def partial(self, feat, heads=1):
        """Workspace for the heavy-row chunk partials (grown on demand, reused)."""
        if self.n_chunks == 0:
            return None
        need = int(_lib.query('peagnn_partial_floats', self.n_chunks, feat, heads))
        if self._partial is None or self._partial.numel() < need:
            self._partial = torch.empty(need, dtype=torch.float32, device=self.rowptr.device)
        return self._partial

    def view(self, feat, heads=1):
        """``peagnn_csr_t`` for the whole structure (ctypes struct; keeps tensors alive via self).
        Cached per (feat, heads): the struct only changes when the partial workspace is regrown."""
        part = self.partial(feat, heads)
        key = (feat, heads, part.data_ptr() if part is not None else 0, self.explicit_self_loops)
        hit = self._views.get(key)
        if hit is not None:
            return hit
        v = self._views[key] = _lib.CsrView()
        v.rowptr = self.rowptr.data_ptr()
        v.col = self.col.data_ptr() if self.nnz else 0
        v.nrows = self.num_nodes
        v.nnz = self.nnz
        v.explicit_self_loops = int(self.explicit_self_loops)
        v.row_offset = 0
        v.heavy_threshold = self.heavy_threshold
        v.n_heavy = self.n_heavy
        v.n_chunks = self.n_chunks
        if self.n_heavy:
            v.heavy_rows = self.heavy_rows.data_ptr()
            v.heavy_chunk_ptr = self.heavy_chunk_ptr.data_ptr()
            v.chunk_row = self.chunk_row.data_ptr()
            v.chunk_begin = self.chunk_begin.data_ptr()
            v.chunk_end = self.chunk_end.data_ptr()
            v.partial = part.data_ptr()
        return v

    def nonempty_row_bitmap(self):
        """Bitmap (``functional.mark_rows`` layout) of the rows that have at least one edge - the node ids that can
        receive anything through this structure (one node type, contiguous upstream)."""
        if getattr(self, '_row_bm', None) is None:
            rp = self.rowptr.long()
            words = (self.num_nodes + 31) // 32 + 1
            on = torch.zeros(words * 32, dtype=torch.bool, device=rp.device)
            on[:self.num_nodes] = (rp[1:] - rp[:-1]) > 0
            w = (on.view(words, 32).to(torch.int64) << torch.arange(32, device=rp.device)).sum(dim=1)
            self._row_bm = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32).contiguous()
        return self._row_bm

    def filtered(self, bitmap, perm=None):
        """Sub-structure holding only the edges that gather a node marked in ``bitmap`` (``functional.mark_rows``),
        original order kept; built once per (structure, bitmap) - every metapath that ends with this relation in a
        demand-driven step shares it - into persistent buffers (rowptr / col / perm at full capacity)."""
        cache = getattr(bitmap, '_sub_csr', None)
        if cache is None:
            cache = bitmap._sub_csr = {}
        hit = cache.get(id(self))
        if hit is not None:
            return hit
        dev = self.rowptr.device
        if getattr(self, '_filt_buf', None) is None:
            ws_bytes = int(_lib.query('peagnn_csr_filter_workspace_bytes', self.nnz))
            self._filt_buf = (torch.empty(self.num_nodes + 1, dtype=torch.int32, device=dev),
                              torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev),
                              torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev),
                              torch.empty(ws_bytes, dtype=torch.uint8, device=dev), ws_bytes)
        rowptr, col, perm_out, ws, ws_bytes = self._filt_buf
        view = self.view(4)
        with _on(dev):
            _lib.call('peagnn_csr_filter', C.byref(view), _ptr(bitmap), _ptr(perm), _ptr(rowptr), _ptr(col), _ptr(perm_out),
                      _ptr(ws), ws_bytes, _stream())
        sub = cache[id(self)] = FilteredCsr(self, rowptr, col, perm_out)
        return sub

    def row_shard(self, lo, hi):
        """The rows [lo, hi) as their own structure (1D destination-row sharding).
        ``col`` stays global; offsets stay absolute into the shared ``col`` array."""
        return CsrShard(self, lo, hi)


class FilteredCsr(object):
    """What ``Csr.filtered`` returns: the kept edges of a structure for one step.  Its edge count lives on the device
    only, so there is no heavy-row list - a kept row is short (a source's neighbours inside one batch)."""

    def __init__(self, parent, rowptr, col, perm):
        self.parent, self.rowptr, self.col, self.perm = parent, rowptr, col, perm
        self.num_nodes = parent.num_nodes
        self._views = {}

    def view(self, feat, heads=1):
        v = self._views.get((feat, heads))
        if v is None:
            v = self._views[(feat, heads)] = _lib.CsrView()
            v.rowptr = self.rowptr.data_ptr()
            v.col = self.col.data_ptr()
            v.nrows = self.num_nodes
            v.nnz = self.num_nodes                      # scheduling hint only: sparse, several rows per warp
            v.row_offset = 0
            v.heavy_threshold = 2 ** 31 - 1
            v.n_heavy = v.n_chunks = 0
            v.explicit_self_loops = int(self.parent.explicit_self_loops)
        return v


class CsrShard(object):
    """Rows [lo, hi) of a Csr: same col array, sliced rowptr, its own heavy-row work list."""

    def __init__(self, parent, lo, hi):
        self.parent, self.lo, self.hi = parent, lo, hi
        self.rowptr = parent.rowptr[lo:hi + 1].contiguous()
        self.col = parent.col
        self.num_rows = hi - lo
        self.heavy_threshold = parent.heavy_threshold
        self.n_heavy = self.n_chunks = 0
        if parent.n_heavy:
            hr = parent.heavy_rows.long()
            sel = torch.nonzero((hr >= lo) & (hr < hi)).flatten()
            self.n_heavy = int(sel.numel())
            if self.n_heavy:
                cp = parent.heavy_chunk_ptr.long()
                c0, c1 = int(cp[sel[0]].item()), int(cp[sel[-1] + 1].item())
                self.heavy_rows = (hr[sel] - lo).int().contiguous()
                self.heavy_chunk_ptr = (cp[sel[0]:sel[-1] + 2] - c0).int().contiguous()
                self.chunk_row = (parent.chunk_row[c0:c1] - lo).contiguous()
                self.chunk_begin = parent.chunk_begin[c0:c1].contiguous()
                self.chunk_end = parent.chunk_end[c0:c1].contiguous()
                self.n_chunks = c1 - c0
        self._partial = None

    def view(self, feat, heads=1):
        v = _lib.CsrView()
        v.rowptr = self.rowptr.data_ptr()
        v.col = self.col.data_ptr() if self.col.numel() else 0
        v.nrows = self.num_rows
        v.row_offset = self.lo
        v.heavy_threshold = self.heavy_threshold
        v.n_heavy = self.n_heavy
        v.n_chunks = self.n_chunks
        if self.n_heavy:
            need = int(_lib.query('peagnn_partial_floats', self.n_chunks, feat, heads))
            if self._partial is None or self._partial.numel() < need:
                self._partial = torch.empty(need, dtype=torch.float32, device=self.rowptr.device)
            v.heavy_rows = self.heavy_rows.data_ptr()
            v.heavy_chunk_ptr = self.heavy_chunk_ptr.data_ptr()
            v.chunk_row = self.chunk_row.data_ptr()
            v.chunk_begin = self.chunk_begin.data_ptr()
            v.chunk_end = self.chunk_end.data_ptr()
            v.partial = self._partial.data_ptr()
        return v


def build_csr(key, val, num_nodes, drop_self_loops=True, heavy_threshold=None, chunk_edges=None):
    """COO (int64, device) -> Csr grouped by ``key`` via peagnn_csr_build."""
    assert key.is_cuda and key.dtype == torch.long and val.dtype == torch.long
    key, val = key.contiguous(), val.contiguous()
    E = int(key.numel())
    dev = key.device
    rowptr = torch.empty(num_nodes + 1, dtype=torch.int32, device=dev)
    col = torch.empty(E, dtype=torch.int32, device=dev)
    eid = torch.empty(E, dtype=torch.int32, device=dev)
    ws_bytes = int(_lib.query('peagnn_csr_workspace_bytes', E, num_nodes))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with _on(dev):
        _lib.call('peagnn_csr_build', _ptr(key), _ptr(val), E, num_nodes, int(drop_self_loops),
                  _ptr(rowptr), _ptr(col), _ptr(eid), _ptr(ws), ws_bytes, _stream())
        kept = int(rowptr[-1].item())       # one-off sync at graph-build time
    del ws
    return Csr(rowptr, col[:kept].contiguous(), eid[:kept].contiguous(), num_nodes, heavy_threshold, chunk_edges)


class RelationGraph(object):
    """Both structures of one relation plus the per-node scalings the three conv families need."""

    def __init__(self, fwd, bwd, num_nodes, num_edges_coo):
        self.fwd, self.bwd = fwd, bwd           # grouped by target / grouped by source
        self.num_nodes = num_nodes
        self.num_edges_coo = num_edges_coo
        self._dis = self._inv_in = None
        self._bwd2fwd = None
        self._t = None

    @classmethod
    def from_edge_index(cls, edge_index, num_nodes, heavy_threshold=None, chunk_edges=None, drop_self_loops=True):
        src, dst = edge_index[0], edge_index[1]
        fwd = build_csr(dst, src, num_nodes, drop_self_loops, heavy_threshold, chunk_edges)
        bwd = build_csr(src, dst, num_nodes, drop_self_loops, heavy_threshold, chunk_edges)
        return cls(fwd, bwd, num_nodes, int(edge_index.shape[1]))

    @property
    def has_no_self_loops(self):
        """True when no edge was dropped at build time, i.e. the relation has no self-loop edges
        (all PEAGNN relations are bipartite) and one structure serves all three conv families."""
        return self.fwd.nnz == self.num_edges_coo

    def transposed(self):
        """The graph of ``torch.flip(edge_index, dims=[0])``: same structures, roles swapped."""
        if self._t is None:
            self._t = RelationGraph(self.bwd, self.fwd, self.num_nodes, self.num_edges_coo)
            self._t._t = self
        return self._t

    @property
    def nnz(self):
        return self.fwd.nnz

    def source_range(self):
        """[lo, hi) node-id range holding every source of the relation (ids are contiguous by type upstream,
        datasets/movielens.py:184-227: this is the source type's range; one host sync, at first use)."""
        if getattr(self, '_src_range', None) is None:
            col = self.fwd.col
            self._src_range = (int(col.min().item()), int(col.max().item()) + 1) if col.numel() else (0, 0)
        return self._src_range

    def _scale(self, csr, add, power, clamp):
        out = torch.empty(self.num_nodes, dtype=torch.float32, device=csr.rowptr.device)
        with _on(out.device):
            _lib.call('peagnn_degree_scale', _ptr(csr.rowptr), self.num_nodes, float(add), float(power),
                      int(clamp), _ptr(out), _stream())
        return out

    @property
    def gcn_dis(self):
        """deg^-1/2 with deg = (#non-loop edges leaving the node) + 1: PyG-1.5.0 GCNConv.norm
        sums the weights over the SOURCE index after add_remaining_self_loops."""
        if self._dis is None:
            self._dis = self._scale(self.bwd, 1.0, -0.5, 0)
        return self._dis

    @property
    def inv_in_degree(self):
        """1 / max(in-degree, 1): SAGEConv mean aggregation (scatter_mean semantics)."""
        if self._inv_in is None:
            self._inv_in = self._scale(self.fwd, 0.0, -1.0, 1)
        return self._inv_in

    @property
    def bwd_to_fwd(self):
        """perm[k] = position in the target-grouped edge arrays of the k-th source-grouped edge
        (GAT backward reads per-edge alpha written in target order)."""
        if self._bwd2fwd is None:
            inv = torch.empty(max(self.num_edges_coo, 1), dtype=torch.int32, device=self.fwd.eid.device)
            inv[self.fwd.eid.long()] = torch.arange(self.fwd.nnz, dtype=torch.int32, device=inv.device)
            self._bwd2fwd = inv[self.bwd.eid.long()].contiguous()
        return self._bwd2fwd


# ---- relation cache -------------------------------------------------------------------------
_by_ptr = {}       # (data_ptr, E, N, device) -> RelationGraph   (fast path for repeated calls)
_by_sig = {}       # content signature -> [(edge_index, RelationGraph)]


def _signature(edge_index, num_nodes):
    s, d = edge_index[0], edge_index[1]
    e = int(edge_index.shape[1])
    if e == 0:
        return (0, num_nodes, 0, 0, 0, str(edge_index.device))
    w = torch.arange(1, e + 1, device=edge_index.device, dtype=torch.long) % 1000003
    return (e, num_nodes, int(s.sum().item()), int(d.sum().item()),
            int(((s * 31 + d * 17) * w).sum().item()), str(edge_index.device))


def _flipped_signature(edge_index, num_nodes):
    s, d = edge_index[1], edge_index[0]
    e = int(edge_index.shape[1])
    if e == 0:
        return (0, num_nodes, 0, 0, 0, str(edge_index.device))
    w = torch.arange(1, e + 1, device=edge_index.device, dtype=torch.long) % 1000003
    return (e, num_nodes, int(s.sum().item()), int(d.sum().item()),
            int(((s * 31 + d * 17) * w).sum().item()), str(edge_index.device))


def get_graph(edge_index, num_nodes, keep_self_loops=False):
    """RelationGraph for a COO edge_index; built on first sight, then served from the cache.
    The caller must not mutate ``edge_index`` in place afterwards (the reference never does).
    GCNConv / GATConv drop self-loop edges and add exactly one loop per node themselves; SAGEConv
    keeps such edges as ordinary ones (``keep_self_loops=True``)."""
    if keep_self_loops:
        g = get_graph(edge_index, num_nodes)
        if g.has_no_self_loops:
            return g
        key = ('keep', edge_index.data_ptr(), int(edge_index.shape[1]), num_nodes, str(edge_index.device))
        hit = _by_ptr.get(key)
        if hit is not None and hit[0]() is not None:
            return hit[1]
        import weakref
        gk = RelationGraph.from_edge_index(edge_index, num_nodes, drop_self_loops=False)
        _by_ptr[key] = (weakref.ref(edge_index), gk)
        return gk
    if not edge_index.is_cuda:
        raise RuntimeError('graph_recsys_benchmark_b200 runs on CUDA only (sm_100a); got a %s edge_index'
                           % edge_index.device)
    if edge_index.dtype != torch.long or edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError('edge_index must be a LongTensor of shape [2, E]')
    key = (edge_index.data_ptr(), int(edge_index.shape[1]), num_nodes, str(edge_index.device),
           edge_index.stride())
    hit = _by_ptr.get(key)
    if hit is not None and hit[0]() is not None:
        return hit[1]
    sig = _signature(edge_index, num_nodes)
    g = None
    for other, cand in _by_sig.get(sig, []):
        if torch.equal(other, edge_index):
            g = cand
            break
    if g is None:
        fsig = _flipped_signature(edge_index, num_nodes)
        for other, cand in _by_sig.get(fsig, []):
            if torch.equal(other[0], edge_index[1]) and torch.equal(other[1], edge_index[0]):
                g = cand.transposed()
                break
    if g is None:
        g = RelationGraph.from_edge_index(edge_index, num_nodes)
        # one strong reference per DISTINCT relation (needed to recognise its copies and its flip later); content
        # hits are not appended again, so rebuilding a model on new device copies does not pin those copies
        _by_sig.setdefault(sig, []).append((edge_index, g))
    elif not any(c is g for _, c in _by_sig.get(sig, [])):
        _by_sig.setdefault(sig, []).append((edge_index, g))       # the flipped orientation of a known relation, once
    import weakref
    _by_ptr[key] = (weakref.ref(edge_index), g)
    return g


def clear_cache():
    _by_ptr.clear()
    _by_sig.clear()
