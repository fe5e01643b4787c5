// One traversal engine for every CSR-by-destination kernel in the library (GCN / SAGE sums,
// GAT row max, GAT softmax aggregate, GAT backward passes).
//
// Scheduling (B200: 148 SMs, 64 resident warps each; the gathered tables are mostly L2-resident,
// so the kernels are bound by instruction issue and L2 bandwidth, not by DRAM):
//   * one WARP owns one destination row.  A row of F floats is held as F/4 float4 chunks, one per
//     lane of a G-lane "slot" (G = F/4 rounded up to a power of two, <= 32); the warp's 32/G
//     slots work on 32/G DIFFERENT EDGES of the same row at once, so every gather instruction is
//     a full-warp 128-bit load and all shuffles use the full mask (sub-warp collectives compile
//     to WARPSYNC / predicated replay sequences that tripled the instruction count in v1);
//   * neighbour ids and the per-edge scalars an op needs are fetched 32 at a time (one coalesced
//     load per lane, next batch prefetched) and handed to the slots with shuffles;
//   * the slots' partial sums are folded with xor-shuffles in a fixed order;
//   * rows above `heavy_threshold` edges are cut into chunks; one CTA reduces a chunk (its 8
//     warps take interleaved 32-edge batches, then add up through shared memory in warp order)
//     into `partial`, and a warp of the second launch folds the chunks of a row in chunk order.
//     No atomics anywhere: results are bit-reproducible run to run.
//
// An Op supplies (all __device__):
//   static constexpr int  NV      floats of per-lane state
//   static constexpr bool kMax    state combines with max instead of +
//   static constexpr bool kUseW2  second per-edge scalar is used
//   int heads                     logical rows = rows * heads
//   void  row_begin(i, h, gl)                 load per-row scalars (i = LOCAL row); full warp converged
//   Edge  load_edge(e, c)                     per-edge scalars, evaluated by one lane per edge
//   void  apply(acc, e, c, w, w2, gl, valid)  fold one edge into the state; executed by the whole
//                                             warp (slots without an edge get valid = false, c = a
//                                             safe row id, w = w2 = 0)
//   void  finish(acc, i, h, gl, writer)       epilogue for the row; `writer` lanes (slot 0) store
// An Op that declares `static constexpr bool kBatchDot = true` (GAT backward, destination side: a dot product of
// the gathered row with a per-row vector, then scalar work per edge) supplies instead of apply():
//   float partial(c, gl)                      this lane's part of <row vector, gathered row c>
//   void  edge_done(acc, e, w, dot, valid)    one LANE per edge: the scalar work and the per-edge stores
// and the engine reduces the 32 dot products of a batch together (process_batches_dot).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace peagnn {

struct Edge {
  int c;
  float w;
  float w2;
};

constexpr int kCtaThreads = 256;
constexpr int kWarpsPerCta = kCtaThreads / 32;
constexpr int kSparseRowsPerWarp = 8;  // light rows per warp when the view is sparse (avg degree < 8), else 1
constexpr unsigned kFull = 0xffffffffu;

template <class Op>
__device__ __forceinline__ float combine(float a, float b) {
  return Op::kMax ? fmaxf(a, b) : a + b;
}
template <class Op>
__device__ __forceinline__ float identity() {
  return Op::kMax ? -INFINITY : 0.f;
}

template <class Op, class = void>
struct is_batch_dot : std::false_type {};
template <class Op>
struct is_batch_dot<Op, std::void_t<decltype(Op::kBatchDot)>> : std::bool_constant<Op::kBatchDot> {};

// Sum over the G lanes of a slot (xor offsets < G never leave the slot).
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ bool bit_set(const uint32_t* __restrict__ bm, int id) {
  return (__ldg(bm + (id >> 5)) >> (id & 31)) & 1u;
}

// Filtered form of process_batches: edges whose gathered node is inactive are dropped before they reach the
// slots.  Four 32-edge batches are probed per iteration (four independent index loads + bitmap probes in flight:
// the walk is latency-bound otherwise); a batch with nothing active costs one coalesced index load, one bitmap
// probe and a ballot - the common case when the filter is a training batch.  The survivors keep their order.
template <class Op, int G>
__device__ __forceinline__ void process_batches_filtered(Op& op, const int32_t* __restrict__ col,
                                                         const uint32_t* __restrict__ active, int first, int end,
                                                         int step, float* acc, int lane, int safe_row) {
  constexpr int EPW = 32 / G;
  constexpr int UN = 4;
  const int gl = lane % G;
  const int slot = lane / G;
  for (int base0 = first; base0 < end; base0 += UN * step) {
    int c[UN];
    bool on[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int e = base0 + u * step + lane;
      c[u] = (base0 + u * step < end && e < end) ? __ldg(col + e) : -1;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) on[u] = c[u] >= 0 && bit_set(active, c[u]);
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const unsigned act = __ballot_sync(kFull, on[u]);
      if (act == 0u) continue;
      const int base = base0 + u * step;
      Edge cur;
      if (on[u]) {
        cur = op.load_edge(base + lane, c[u]);
      } else {
        cur.c = safe_row; cur.w = 0.f; cur.w2 = 0.f;
      }
      const int n_act = __popc(act);
      for (int s = 0; s * EPW < n_act; ++s) {
        const int k = s * EPW + slot;
        const bool valid = k < n_act;
        const int src = (int)__fns(act, 0, valid ? k + 1 : 1);
        const int cc = __shfl_sync(kFull, cur.c, src);
        const float w = __shfl_sync(kFull, cur.w, src);
        const float w2 = Op::kUseW2 ? __shfl_sync(kFull, cur.w2, src) : 0.f;
        op.apply(acc, base + src, cc, w, w2, gl, valid);
      }
    }
  }
}

// Process the 32-edge batches first, first + step, ... < end of one row.
template <class Op, int G>
__device__ __forceinline__ void process_batches(Op& op, const int32_t* __restrict__ col, int first,
                                                int end, int step, float* acc, int lane, int safe_row) {
  constexpr int EPW = 32 / G;  // edges in flight per warp instruction
  const int gl = lane % G;
  const int slot = lane / G;
  if (first >= end) return;
  Edge cur;
  {
    const int e = first + lane;
    if (e < end) {
      cur = op.load_edge(e, __ldg(col + e));
    } else {
      cur.c = safe_row; cur.w = 0.f; cur.w2 = 0.f;
    }
  }
  for (int base = first; base < end; base += step) {
    Edge nxt;
    {
      const int e = base + step + lane;     // prefetch the next batch's scalars
      if (e < end) {
        nxt = op.load_edge(e, __ldg(col + e));
      } else {
        nxt.c = safe_row; nxt.w = 0.f; nxt.w2 = 0.f;
      }
    }
    const int cnt = min(32, end - base);
    if (cnt == 32) {
#pragma unroll
      for (int s = 0; s < G; ++s) {
        const int j = s * EPW + slot;
        const int c = __shfl_sync(kFull, cur.c, j);
        const float w = __shfl_sync(kFull, cur.w, j);
        const float w2 = Op::kUseW2 ? __shfl_sync(kFull, cur.w2, j) : 0.f;
        op.apply(acc, base + j, c, w, w2, gl, true);
      }
    } else {
      const int steps = (cnt + EPW - 1) / EPW;
      for (int s = 0; s < steps; ++s) {
        const int j = s * EPW + slot;
        const int c = __shfl_sync(kFull, cur.c, j);
        const float w = __shfl_sync(kFull, cur.w, j);
        const float w2 = Op::kUseW2 ? __shfl_sync(kFull, cur.w2, j) : 0.f;
        op.apply(acc, base + j, c, w, w2, gl, j < cnt);
      }
    }
    cur = nxt;
  }
}

// Transposed reduction inside a slot: every lane holds G partial sums v[0..G) (one per batch iteration); afterwards
// lane gl of the slot holds the slot-wide total of v[gl].  G - 1 shuffles for G totals - a plain group_sum per value
// would take G * log2(G).
template <int G>
__device__ __forceinline__ float slot_transpose_sum(float* v, int gl) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const bool up = (gl & o) != 0;
#pragma unroll
    for (int k = 0; k < o; ++k) {
      const float send = up ? v[k] : v[k + o];
      const float keep = up ? v[k + o] : v[k];
      v[k] = keep + __shfl_xor_sync(kFull, send, o);
    }
  }
  return v[0];
}

// process_batches for kBatchDot ops.  The gather loop only accumulates each lane's part of the dot product of
// every edge of the batch (no per-edge reduction, no scalar work, no store); one transposed reduction then leaves
// the total of edge `base + lane` in lane `lane`, which already holds that edge's scalars from the coalesced batch
// load - so the exp / softmax-gradient arithmetic runs once per batch on 32 edges and the per-edge results are
// stored coalesced.  acc[] holds PER-LANE sums here; fold_slots adds them up over the whole warp.
template <class Op, int G>
__device__ __forceinline__ void process_batches_dot(Op& op, const int32_t* __restrict__ col, int first,
                                                    int end, int step, float* acc, int lane, int safe_row) {
  constexpr int EPW = 32 / G;
  const int gl = lane % G;
  const int slot = lane / G;
  if (first >= end) return;
  Edge cur;
  {
    const int e = first + lane;
    if (e < end) {
      cur = op.load_edge(e, __ldg(col + e));
    } else {
      cur.c = safe_row; cur.w = 0.f; cur.w2 = 0.f;
    }
  }
  const int back = (lane % EPW) * G + lane / EPW;   // the lane that ends up with edge `lane`'s total
  for (int base = first; base < end; base += step) {
    Edge nxt;
    {
      const int e = base + step + lane;
      if (e < end) {
        nxt = op.load_edge(e, __ldg(col + e));
      } else {
        nxt.c = safe_row; nxt.w = 0.f; nxt.w2 = 0.f;
      }
    }
    const int cnt = min(32, end - base);
    float part[G];
    if (cnt == 32) {                                // branch-free: the G gathers of a full batch are issued together
#pragma unroll
      for (int s = 0; s < G; ++s) part[s] = op.partial(__shfl_sync(kFull, cur.c, s * EPW + slot), gl);
    } else {                                        // a row's tail: groups of four gathers under one warp-uniform test
      const int steps = (cnt + EPW - 1) / EPW;
#pragma unroll
      for (int s0 = 0; s0 < G; s0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) part[s0 + u] = 0.f;
        if (s0 < steps) {
#pragma unroll
          for (int u = 0; u < 4; ++u)              // lanes past the row's end hold the safe row
            part[s0 + u] = op.partial(__shfl_sync(kFull, cur.c, (s0 + u) * EPW + slot), gl);
        }
      }
    }
    float tot = slot_transpose_sum<G>(part, gl);    // lane (slot, gl) : edge gl * EPW + slot
    tot = __shfl_sync(kFull, tot, back);
    op.edge_done(acc, base + lane, cur.w, tot, lane < cnt);
    cur = nxt;
  }
}

// One row's (or one chunk's) edges through whichever walk the op and the view's filters ask for.
template <class Op, int G, bool FILT>
__device__ __forceinline__ void walk_row(Op& op, const peagnn_csr_t& g, int first, int end, int step,
                                         float* acc, int lane, int safe_row) {
  if constexpr (is_batch_dot<Op>::value) {
    process_batches_dot<Op, G>(op, g.col, first, end, step, acc, lane, safe_row);   // (no column filter for these ops)
  } else {
    if (FILT && g.active_cols)
      process_batches_filtered<Op, G>(op, g.col, g.active_cols, first, end, step, acc, lane, safe_row);
    else
      process_batches<Op, G>(op, g.col, first, end, step, acc, lane, safe_row);
  }
}

// Fold the 32/G slots of a warp: afterwards every lane holds the row total of its chunk.  (kBatchDot ops keep
// per-lane sums: the fold runs over all 32 lanes.)
template <class Op, int G>
__device__ __forceinline__ void fold_slots(float* acc) {
#pragma unroll
  for (int o = is_batch_dot<Op>::value ? 1 : G; o < 32; o <<= 1) {
#pragma unroll
    for (int v = 0; v < Op::NV; ++v) acc[v] = combine<Op>(acc[v], __shfl_xor_sync(kFull, acc[v], o));
  }
}

// Launch 1 (only when the view has heavy rows): one CTA per (chunk, head).
template <class Op, int G, bool FILT = false>
__global__ void __launch_bounds__(kCtaThreads) csr_chunk_kernel(const peagnn_csr_t g, const Op op_in) {
  __shared__ float sm[kWarpsPerCta * G * Op::NV];
  Op op = op_in;
  const int heads = op.heads;
  const int chunk = blockIdx.x / heads;
  const int h = blockIdx.x - chunk * heads;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane % G;
  const int i = g.chunk_row[chunk];
  if (FILT && g.active_rows && !bit_set(g.active_rows, i)) return;     // nobody folds this row's partials
  const int cb = g.chunk_begin[chunk], ce = g.chunk_end[chunk];
  float acc[Op::NV];
#pragma unroll
  for (int v = 0; v < Op::NV; ++v) acc[v] = identity<Op>();
  op.row_begin(i, h, gl);
  walk_row<Op, G, FILT>(op, g, cb + warp * 32, ce, kWarpsPerCta * 32, acc, lane, g.row_offset + i);
  fold_slots<Op, G>(acc);
  if (lane < G) {
#pragma unroll
    for (int v = 0; v < Op::NV; ++v) sm[(warp * Op::NV + v) * G + gl] = acc[v];
  }
  __syncthreads();
  float* dst = g.partial + (size_t)blockIdx.x * (G * Op::NV);
  for (int idx = threadIdx.x; idx < G * Op::NV; idx += kCtaThreads) {
    float r = sm[idx];
#pragma unroll
    for (int q = 1; q < kWarpsPerCta; ++q) r = combine<Op>(r, sm[q * (G * Op::NV) + idx]);
    dst[idx] = r;
  }
}

// Launch 2: the first blocks fold the heavy rows' chunk partials, the rest take kRowsPerWarp light
// rows per warp.
template <class Op, int G, int RPW, bool FILT = false>
__global__ void __launch_bounds__(kCtaThreads) csr_rows_kernel(const peagnn_csr_t g, const Op op_in) {
  Op op = op_in;
  const int heads = op.heads;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane % G;
  const bool writer = lane < G;
  const long long n_heavy_l = (long long)g.n_heavy * heads;
  const long long heavy_blocks = (n_heavy_l + kWarpsPerCta - 1) / kWarpsPerCta;
  float acc[Op::NV];
#pragma unroll
  for (int v = 0; v < Op::NV; ++v) acc[v] = identity<Op>();

  if ((long long)blockIdx.x < heavy_blocks) {
    const long long hl = (long long)blockIdx.x * kWarpsPerCta + warp;
    if (hl >= n_heavy_l) return;
    const int hi = (int)(hl / heads);
    const int h = (int)(hl - (long long)hi * heads);
    const int i = g.heavy_rows[hi];
    if (FILT && g.active_rows && !bit_set(g.active_rows, i)) return;
    op.row_begin(i, h, gl);
    const int c0 = g.heavy_chunk_ptr[hi], c1 = g.heavy_chunk_ptr[hi + 1];
    for (int ch = c0; ch < c1; ++ch) {
      const float* p = g.partial + ((size_t)ch * heads + h) * (G * Op::NV);
#pragma unroll
      for (int v = 0; v < Op::NV; ++v) acc[v] = combine<Op>(acc[v], p[v * G + gl]);
    }
    op.finish(acc, i, h, gl, writer);
    return;
  }
  // Light rows: each warp takes kRowsPerWarp consecutive logical rows.  Rows WITHOUT edges (most
  // rows of most relations: every node that is not a target of the relation) only need the
  // epilogue; they are handed to the warp's slots 32/G at a time so that several rows' loads and
  // stores are in flight together.  Rows with edges then run one at a time on the whole warp.
  constexpr int EPW = 32 / G;
  const long long total = (long long)g.nrows * heads;
  const long long lr0 = (((long long)blockIdx.x - heavy_blocks) * kWarpsPerCta + warp) * RPW;
  if (lr0 >= total) return;
  const long long my_lr = lr0 + lane;
  bool mine = lane < RPW && my_lr < total;
  if (FILT && g.active_rows && mine) mine = bit_set(g.active_rows, (int)(my_lr / heads));
  int start_l = 0, end_l = 0;
  if (mine) {
    const int i_l = (int)(my_lr / heads);
    start_l = g.rowptr[i_l];
    end_l = g.rowptr[i_l + 1];
  }
  const int deg_l = end_l - start_l;
  const bool heavy_l = g.n_heavy > 0 && deg_l > g.heavy_threshold;   // folded above
  const unsigned empty_mask = __ballot_sync(kFull, mine && deg_l == 0);
  unsigned work_mask = __ballot_sync(kFull, mine && deg_l > 0 && !heavy_l);

  const int n_empty = __popc(empty_mask);
  const int slot = lane / G;
  for (int base = 0; base < n_empty; base += EPW) {
    const int k = base + slot;
    const bool has = k < n_empty;
    const int src_lane = has ? (int)__fns(empty_mask, 0, k + 1) : (int)__fns(empty_mask, 0, 1);
    const long long lr = lr0 + src_lane;
    const int i = (int)(lr / heads);
    const int h = (int)(lr - (long long)i * heads);
    float e_acc[Op::NV];
#pragma unroll
    for (int v = 0; v < Op::NV; ++v) e_acc[v] = identity<Op>();
    op.row_begin(i, h, gl);            // slots hold different rows here; shuffles stay inside a slot
    op.finish(e_acc, i, h, gl, has);
  }
  while (work_mask) {
    const int l = __ffs(work_mask) - 1;
    work_mask &= work_mask - 1;
    const long long lr = lr0 + l;
    const int i = (int)(lr / heads);
    const int h = (int)(lr - (long long)i * heads);
    const int start = __shfl_sync(kFull, start_l, l);
    const int end = __shfl_sync(kFull, end_l, l);
#pragma unroll
    for (int v = 0; v < Op::NV; ++v) acc[v] = identity<Op>();
    op.row_begin(i, h, gl);
    walk_row<Op, G, FILT>(op, g, start, end, 32, acc, lane, g.row_offset + i);
    fold_slots<Op, G>(acc);
    op.finish(acc, i, h, gl, writer);
  }
}

template <class Op, int G, bool FILT = false>
int launch_csr(const peagnn_csr_t& g, const Op& op, cudaStream_t stream, const char* what) {
  const int heads = op.heads;
  if (g.n_heavy > 0) {
    PEAGNN_REQUIRE(g.partial != nullptr && g.heavy_rows && g.heavy_chunk_ptr && g.chunk_row &&
                       g.chunk_begin && g.chunk_end && g.n_chunks > 0,
                   "%s: heavy-row work list incomplete", what);
    csr_chunk_kernel<Op, G, FILT><<<(unsigned)(g.n_chunks * heads), kCtaThreads, 0, stream>>>(g, op);
    int rc = check_launch(what);
    if (rc) return rc;
  }
  const long long heavy_blocks = ((long long)g.n_heavy * heads + kWarpsPerCta - 1) / kWarpsPerCta;
  // sparse views (most rows have no edge) pack several rows per warp so that the edge-less rows'
  // epilogues overlap; dense views keep one row per warp for parallelism and balance
  const bool sparse = g.nnz > 0 && g.nnz < 8ll * g.nrows;
  // a row filter that marks a few percent of the rows (a training batch): one warp per 32-row bitmap word
  const bool few_rows = FILT && g.active_rows && g.sparse_filter;
  const int rpw = few_rows ? 32 : sparse ? kSparseRowsPerWarp : 1;
  const long long rows_per_block = (long long)kWarpsPerCta * rpw;
  const long long light_blocks = ((long long)g.nrows * heads + rows_per_block - 1) / rows_per_block;
  const long long blocks = heavy_blocks + light_blocks;
  if (blocks == 0) return PEAGNN_OK;
  if constexpr (FILT) {
    if (few_rows) {
      csr_rows_kernel<Op, G, 32, FILT><<<(unsigned)blocks, kCtaThreads, 0, stream>>>(g, op);
      return check_launch(what);
    }
  }
  if (sparse)
    csr_rows_kernel<Op, G, kSparseRowsPerWarp, FILT><<<(unsigned)blocks, kCtaThreads, 0, stream>>>(g, op);
  else
    csr_rows_kernel<Op, G, 1, FILT><<<(unsigned)blocks, kCtaThreads, 0, stream>>>(g, op);
  return check_launch(what);
}

// Slot geometry for a feature width: lanes per slot, float4 chunks per lane.
struct Geometry {
  int G;
  int CPL;
};
inline Geometry geometry_for(int feat) {
  const int f4 = (feat + 3) / 4;
  if (f4 <= 4) return {4, 1};
  if (f4 <= 8) return {8, 1};
  if (f4 <= 16) return {16, 1};
  if (f4 <= 32) return {32, 1};
  if (f4 <= 64) return {32, 2};
  return {32, 4};
}

}  // namespace peagnn
