// One traversal engine for every CSR-by-destination kernel in the library (GCN / SAGE sums,
// GAT row max, GAT softmax aggregate, GAT backward passes).
//
// Scheduling (B200: 148 SMs, 64 resident warps each, gathers served mostly by the 126 MB L2):
//   * a "group" of G lanes (G = feature width / 4, a power of two <= 32) owns one destination
//     row; each lane keeps one float4 of the row in registers, so a gathered neighbour row is
//     read with one 128-bit load per lane and a warp covers 32/G rows;
//   * neighbour ids (and the per-edge scalars an op needs) are fetched G*IPL at a time with one
//     coalesced load per lane and handed round with warp shuffles, so the gather loop has
//     G*IPL independent 128-bit loads in flight per group;
//   * rows above `heavy_threshold` edges are cut into chunks; one CTA reduces a chunk (its 256/G
//     groups take interleaved batches, then add up through shared memory in group order) into
//     `partial`, and a group of the second launch folds the chunks of a row in chunk order.
//     No atomics anywhere: results are bit-reproducible run to run.
//
// An Op supplies (all __device__):
//   static constexpr int  NV      floats of per-lane state
//   static constexpr bool kMax    state combines with max instead of +
//   static constexpr bool kUseW2  second per-edge scalar is used
//   int heads                     logical rows = rows * heads
//   void  row_begin(i, h, gl, gmask)          load per-row scalars (i = LOCAL row)
//   Edge  load_edge(e, c)                      per-edge scalars, evaluated by one lane per edge
//   void  apply(acc, e, c, w, w2, gl, gmask)   fold one edge into the state (all lanes of the group)
//   void  finish(acc, i, h, gl, gmask)         epilogue for the row
#pragma once
#include "common.cuh"

namespace peagnn {

struct Edge {
  int c;
  float w;
  float w2;
};

constexpr int kCtaThreads = 256;

template <class Op>
__device__ __forceinline__ float combine(float a, float b) {
  return Op::kMax ? fmaxf(a, b) : a + b;
}
template <class Op>
__device__ __forceinline__ float identity() {
  return Op::kMax ? -INFINITY : 0.f;
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (G == 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << ((lane / G) * G);
  }
}

template <class Op, int G, int IPL>
__device__ __forceinline__ void process_batches(Op& op, const int32_t* __restrict__ col, int first,
                                                int end, int step, float* acc, int gl,
                                                unsigned gmask) {
  constexpr int BATCH = G * IPL;
  for (int base = first; base < end; base += step) {
    Edge es[IPL];
#pragma unroll
    for (int q = 0; q < IPL; ++q) {
      const int e = base + q * G + gl;
      if (e < end) {
        const int c = __ldg(col + e);
        es[q] = op.load_edge(e, c);
      } else {
        es[q].c = 0;
        es[q].w = 0.f;
        es[q].w2 = 0.f;
      }
    }
    const int cnt = min(BATCH, end - base);
    if (cnt == BATCH) {
#pragma unroll
      for (int q = 0; q < IPL; ++q) {
#pragma unroll 8
        for (int k = 0; k < G; ++k) {
          const int c = __shfl_sync(gmask, es[q].c, k, G);
          const float w = __shfl_sync(gmask, es[q].w, k, G);
          const float w2 = Op::kUseW2 ? __shfl_sync(gmask, es[q].w2, k, G) : 0.f;
          op.apply(acc, base + q * G + k, c, w, w2, gl, gmask);
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < IPL; ++q) {
        const int lim = min(G, cnt - q * G);
        for (int k = 0; k < lim; ++k) {
          const int c = __shfl_sync(gmask, es[q].c, k, G);
          const float w = __shfl_sync(gmask, es[q].w, k, G);
          const float w2 = Op::kUseW2 ? __shfl_sync(gmask, es[q].w2, k, G) : 0.f;
          op.apply(acc, base + q * G + k, c, w, w2, gl, gmask);
        }
      }
    }
  }
}

// Launch 1 (only when the view has heavy rows): one CTA per (chunk, head).
template <class Op, int G, int IPL>
__global__ void __launch_bounds__(kCtaThreads) csr_chunk_kernel(const peagnn_csr_t g, const Op op_in) {
  constexpr int GPB = kCtaThreads / G;
  constexpr int BATCH = G * IPL;
  __shared__ float sm[kCtaThreads * Op::NV];
  Op op = op_in;
  const int heads = op.heads;
  const int chunk = blockIdx.x / heads;
  const int h = blockIdx.x - chunk * heads;
  const int gl = threadIdx.x % G;
  const int grp = threadIdx.x / G;
  const unsigned gmask = group_mask<G>();
  const int i = g.chunk_row[chunk];
  const int cb = g.chunk_begin[chunk], ce = g.chunk_end[chunk];
  float acc[Op::NV];
#pragma unroll
  for (int v = 0; v < Op::NV; ++v) acc[v] = identity<Op>();
  op.row_begin(i, h, gl, gmask);
  process_batches<Op, G, IPL>(op, g.col, cb + grp * BATCH, ce, GPB * BATCH, acc, gl, gmask);
#pragma unroll
  for (int v = 0; v < Op::NV; ++v) sm[(grp * Op::NV + v) * G + gl] = acc[v];
  __syncthreads();
  float* dst = g.partial + (size_t)blockIdx.x * (G * Op::NV);
  for (int idx = threadIdx.x; idx < G * Op::NV; idx += kCtaThreads) {
    float r = sm[idx];
    for (int q = 1; q < GPB; ++q) r = combine<Op>(r, sm[q * (G * Op::NV) + idx]);
    dst[idx] = r;
  }
}

// Launch 2: the first blocks fold the heavy rows' chunk partials, the rest take one light row
// per group.
template <class Op, int G, int IPL>
__global__ void __launch_bounds__(kCtaThreads) csr_rows_kernel(const peagnn_csr_t g, const Op op_in) {
  constexpr int GPB = kCtaThreads / G;
  Op op = op_in;
  const int heads = op.heads;
  const int gl = threadIdx.x % G;
  const int grp = threadIdx.x / G;
  const unsigned gmask = group_mask<G>();
  const long long n_heavy_l = (long long)g.n_heavy * heads;
  const long long heavy_blocks = (n_heavy_l + GPB - 1) / GPB;
  float acc[Op::NV];
#pragma unroll
  for (int v = 0; v < Op::NV; ++v) acc[v] = identity<Op>();

  if ((long long)blockIdx.x < heavy_blocks) {
    const long long hl = (long long)blockIdx.x * GPB + grp;
    if (hl >= n_heavy_l) return;
    const int hi = (int)(hl / heads);
    const int h = (int)(hl - (long long)hi * heads);
    const int i = g.heavy_rows[hi];
    op.row_begin(i, h, gl, gmask);
    const int c0 = g.heavy_chunk_ptr[hi], c1 = g.heavy_chunk_ptr[hi + 1];
    for (int ch = c0; ch < c1; ++ch) {
      const float* p = g.partial + ((size_t)ch * heads + h) * (G * Op::NV);
#pragma unroll
      for (int v = 0; v < Op::NV; ++v) acc[v] = combine<Op>(acc[v], p[v * G + gl]);
    }
    op.finish(acc, i, h, gl, gmask);
    return;
  }
  const long long lr = ((long long)blockIdx.x - heavy_blocks) * GPB + grp;
  if (lr >= (long long)g.nrows * heads) return;
  const int i = (int)(lr / heads);
  const int h = (int)(lr - (long long)i * heads);
  const int start = g.rowptr[i], end = g.rowptr[i + 1];
  if (g.n_heavy > 0 && end - start > g.heavy_threshold) return;  // folded above
  op.row_begin(i, h, gl, gmask);
  process_batches<Op, G, IPL>(op, g.col, start, end, G * IPL, acc, gl, gmask);
  op.finish(acc, i, h, gl, gmask);
}

template <class Op, int G, int IPL>
int launch_csr(const peagnn_csr_t& g, const Op& op, cudaStream_t stream, const char* what) {
  constexpr int GPB = kCtaThreads / G;
  const int heads = op.heads;
  if (g.n_heavy > 0) {
    PEAGNN_REQUIRE(g.partial != nullptr && g.heavy_rows && g.heavy_chunk_ptr && g.chunk_row &&
                       g.chunk_begin && g.chunk_end && g.n_chunks > 0,
                   "%s: heavy-row work list incomplete", what);
    csr_chunk_kernel<Op, G, IPL><<<(unsigned)(g.n_chunks * heads), kCtaThreads, 0, stream>>>(g, op);
    int rc = check_launch(what);
    if (rc) return rc;
  }
  const long long heavy_blocks = ((long long)g.n_heavy * heads + GPB - 1) / GPB;
  const long long light_blocks = ((long long)g.nrows * heads + GPB - 1) / GPB;
  const long long blocks = heavy_blocks + light_blocks;
  if (blocks == 0) return PEAGNN_OK;
  csr_rows_kernel<Op, G, IPL><<<(unsigned)blocks, kCtaThreads, 0, stream>>>(g, op);
  return check_launch(what);
}

// Group geometry for a feature width: lanes per row, float4 chunks per lane.
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

template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, G);
  return v;
}

}  // namespace peagnn
