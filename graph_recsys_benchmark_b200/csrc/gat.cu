// K3: GAT edge-softmax aggregation and its backward (GATConv of PyG-1.5.0 as called from
// models/peagat.py:16-21).  All four passes run on the shared CSR traversal
// engine (csr_traverse.cuh): one warp per (destination row, head).
//
//   e_ij   = leaky_relu(a_i[i] + a_j[j], slope)          j over row i's neighbours plus i itself
//   alpha  = exp(e_ij - max_i) / (sum_i exp(e - max_i) + 1e-16)
//   out_i  = sum_j alpha_ij * H[j]
// leaky_relu is monotone, so max_i = leaky_relu(a_i[i] + max_j a_j[j]): the row max is a plain
// segmented max over gathered scalars (pass 1) and the aggregate needs one pass over the rows.
#include "csr_traverse.cuh"

namespace peagnn {

// ---- pass 1: row max (G = 1: 32 edges per warp instruction) -----------------------------------
struct RowMaxOp {
  static constexpr int NV = 1;
  static constexpr bool kMax = true;
  static constexpr bool kUseW2 = false;
  int heads;
  const float* __restrict__ a_i;
  const float* __restrict__ a_j;
  float* __restrict__ rowmax;
  float slope;
  int row_offset;
  int self_loop;   // add the implicit self loop (0 when the view stores self loops as edges)
  int h_;

  __device__ __forceinline__ void row_begin(int, int h, int) { h_ = h; }
  __device__ __forceinline__ Edge load_edge(int, int c) const {
    Edge e;
    e.c = c;
    e.w = __ldg(a_j + (int64_t)c * heads + h_);
    e.w2 = 0.f;
    return e;
  }
  __device__ __forceinline__ void apply(float* acc, int, int, float w, float, int, bool valid) const {
    if (valid) acc[0] = fmaxf(acc[0], w);
  }
  __device__ __forceinline__ void finish(float* acc, int i, int h, int, bool writer) const {
    const int64_t gi = (int64_t)(row_offset + i) * heads + h;
    const float m = self_loop ? fmaxf(acc[0], __ldg(a_j + gi)) : acc[0];  // the self loop
    if (writer) rowmax[gi] = leaky(__ldg(a_i + gi) + m, slope);
  }
};

// ---- pass 2: softmax-weighted aggregate ------------------------------------------------------
template <int CPL, int G>
struct GatAggOp {
  static constexpr int NV = 4 * CPL + 1;  // + running denominator
  static constexpr bool kMax = false;
  static constexpr bool kUseW2 = false;
  int heads;
  const float* __restrict__ H;
  int64_t ldh;
  int feat, f4;
  const float* __restrict__ a_i;
  const float* __restrict__ a_j;
  const float* __restrict__ rowmax;
  float* __restrict__ denom;
  float* __restrict__ out;
  int64_t ldo;
  const float* __restrict__ bias;
  float slope;
  int row_offset, relu;
  int self_loop;
  // per-row registers
  float ai_, m_;
  int h_;

  __device__ __forceinline__ void row_begin(int i, int h, int) {
    const int64_t gi = (int64_t)(row_offset + i) * heads + h;
    ai_ = __ldg(a_i + gi);
    m_ = __ldg(rowmax + gi);
    h_ = h;
  }
  __device__ __forceinline__ Edge load_edge(int, int c) const {
    Edge e;
    e.c = c;
    e.w = expf(leaky(ai_ + __ldg(a_j + (int64_t)c * heads + h_), slope) - m_);
    e.w2 = 0.f;
    return e;
  }
  // no per-lane predicate in the gather loop (see SpmmOp::apply): lanes beyond the row width re-read its last chunk and
  // their sums are never stored; `valid` is the literal true on the full-batch path
  __device__ __forceinline__ void apply(float* acc, int, int c, float w, float, int gl, bool valid) const {
    if (!valid) { w = 0.f; c = 0; }      // idle slots: row 0 always exists in the gathered table, the view's safe row may not
    const float* xr = H + row_off(c, (unsigned)ldh) + h_ * feat;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      float4 v = ldg4(xr + 4 * min(gl + ch * G, f4 - 1));
      if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
      acc[4 * ch + 0] = fmaf(w, v.x, acc[4 * ch + 0]);
      acc[4 * ch + 1] = fmaf(w, v.y, acc[4 * ch + 1]);
      acc[4 * ch + 2] = fmaf(w, v.z, acc[4 * ch + 2]);
      acc[4 * ch + 3] = fmaf(w, v.w, acc[4 * ch + 3]);
    }
    acc[4 * CPL] += w;
  }
  __device__ __forceinline__ void finish(float* acc, int i, int h, int gl, bool writer) const {
    if (!writer) return;
    const int64_t gi = (int64_t)(row_offset + i) * heads + h;
    const float ws = self_loop ? expf(leaky(ai_ + __ldg(a_j + gi), slope) - m_) : 0.f;  // self loop, appended last
    const float den = acc[4 * CPL] + ws + 1e-16f;
    if (gl == 0 && denom) denom[gi] = den;
    const float inv = 1.f / den;
    const float* xi = H + (int64_t)(row_offset + i) * ldh + (int64_t)h * feat;
    float* o = out + (int64_t)i * ldo + (int64_t)h * feat;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      const int idx = gl + ch * G;
      if (idx < f4) {
        const float4 v = self_loop ? ldg4(xi + 4 * idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 a;
        a.x = fmaf(ws, v.x, acc[4 * ch + 0]) * inv;
        a.y = fmaf(ws, v.y, acc[4 * ch + 1]) * inv;
        a.z = fmaf(ws, v.z, acc[4 * ch + 2]) * inv;
        a.w = fmaf(ws, v.w, acc[4 * ch + 3]) * inv;
        if (bias) {
          const float4 b = ldg4(bias + (int64_t)h * feat + 4 * idx);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        st4(o + 4 * idx, a);
      }
    }
  }
};

// ---- backward, destination side ---------------------------------------------------------------
// d alpha_ij = <dout_i, H_j>;  d e_ij = alpha_ij (d alpha_ij - <dout_i, agg_i>);
// ds_ij = d e_ij * leaky'(a_i[i] + a_j[j]);  d a_i[i] = sum_j ds_ij.
// A kBatchDot op (csr_traverse.cuh): the gather loop only accumulates each lane's part of <dout_i, H_j>; the 32 dot
// products of a batch are reduced together and the exp / gradient arithmetic then runs one lane per edge, with the
// (alpha, ds) pair of 32 consecutive edges stored as one coalesced 256-byte line.
template <int CPL, int G>
struct GatBwdDstOp {
  static constexpr int NV = 1;
  static constexpr bool kMax = false;
  static constexpr bool kUseW2 = false;
  static constexpr bool kBatchDot = true;
  int heads;
  const float* __restrict__ H;
  int64_t ldh;
  int feat, f4;
  const float* __restrict__ a_i;
  const float* __restrict__ a_j;
  const float* __restrict__ rowmax;
  const float* __restrict__ denom;
  const float* __restrict__ agg;
  int64_t lda;
  const float* __restrict__ agg_bias;   // subtracted from `agg` on load when non-null
  const float* __restrict__ dout;
  int64_t ldd;
  float2* __restrict__ ads_e;           // [nnz, heads] (alpha, ds)
  float* __restrict__ alpha_self;
  float* __restrict__ ds_self;
  float* __restrict__ d_ai;
  float slope;
  int row_offset;
  int self_loop;
  // per-row registers
  float ai_, m_, inv_den_, D_;
  int h_;
  float4 g_[CPL];

  __device__ __forceinline__ void row_begin(int i, int h, int gl) {
    const int64_t gi = (int64_t)(row_offset + i) * heads + h;
    ai_ = __ldg(a_i + gi);
    m_ = __ldg(rowmax + gi);
    inv_den_ = 1.f / __ldg(denom + gi);
    h_ = h;
    const float* dp = dout + (int64_t)i * ldd + (int64_t)h * feat;
    const float* ap = agg + (int64_t)i * lda + (int64_t)h * feat;
    float d = 0.f;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      const int idx = gl + ch * G;
      g_[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < f4) {
        g_[ch] = ldg4(dp + 4 * idx);
        float4 a = ldg4(ap + 4 * idx);
        if (agg_bias) {
          const float4 b = ldg4(agg_bias + (int64_t)h * feat + 4 * idx);
          a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w;
        }
        d += g_[ch].x * a.x + g_[ch].y * a.y + g_[ch].z * a.z + g_[ch].w * a.w;
      }
    }
    D_ = group_sum<G>(d);
  }
  __device__ __forceinline__ Edge load_edge(int, int c) const {
    Edge e;
    e.c = c;
    e.w = ai_ + __ldg(a_j + (int64_t)c * heads + h_);  // raw logit
    e.w2 = 0.f;
    return e;
  }
  // this lane's part of <dout_i, H_c>.  No per-lane predicate (see SpmmOp::apply): lanes beyond the row width re-read
  // its last chunk against g_ = 0
  __device__ __forceinline__ float partial(int c, int gl) const {
    const float* xr = H + row_off(c, (unsigned)ldh) + h_ * feat;
    float d = 0.f;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      const float4 v = ldg4(xr + 4 * min(gl + ch * G, f4 - 1));
      d = fmaf(g_[ch].x, v.x, d);
      d = fmaf(g_[ch].y, v.y, d);
      d = fmaf(g_[ch].z, v.z, d);
      d = fmaf(g_[ch].w, v.w, d);
    }
    return d;
  }
  __device__ __forceinline__ float edge_grad(float s_raw, float dot, float& alpha) const {
    alpha = expf(leaky(s_raw, slope) - m_) * inv_den_;
    return alpha * (dot - D_) * (s_raw > 0.f ? 1.f : slope);
  }
  // one lane per edge
  __device__ __forceinline__ void edge_done(float* acc, int e, float s_raw, float dot, bool valid) const {
    if (!valid) return;
    float alpha;
    const float ds = edge_grad(s_raw, dot, alpha);
    ads_e[(int64_t)e * heads + h_] = make_float2(alpha, ds);
    acc[0] += ds;
  }
  __device__ __forceinline__ void finish(float* acc, int i, int h, int gl, bool writer) const {
    const int64_t node = row_offset + i;
    const int64_t gi = node * heads + h;
    if (!self_loop) {
      if (writer && gl == 0) d_ai[gi] = acc[0];
      return;
    }
    const float dot = group_sum<G>(partial((int)node, gl));      // every lane of the slot (shuffles inside)
    float alpha;
    const float ds = edge_grad(ai_ + __ldg(a_j + gi), dot, alpha);
    if (writer && gl == 0) {
      alpha_self[gi] = alpha;
      ds_self[gi] = ds;
      d_ai[gi] = acc[0] + ds;
    }
  }
};

// ---- backward, source side (transposed structure) ---------------------------------------------
template <int CPL, int G>
struct GatBwdSrcOp {
  static constexpr int NV = 4 * CPL + 1;
  static constexpr bool kMax = false;
  static constexpr bool kUseW2 = true;
  int heads;
  const int32_t* __restrict__ perm;
  const float2* __restrict__ ads_e;     // (alpha, ds) per edge, destination order: ONE 8-byte read through the permutation
  const float* __restrict__ alpha_self;
  const float* __restrict__ ds_self;
  const float* __restrict__ dout;
  int64_t ldd;
  int feat, f4;
  float* __restrict__ dH;
  int64_t ldh;
  float* __restrict__ d_aj;
  int row_offset;
  int self_loop;
  int h_;

  __device__ __forceinline__ void row_begin(int, int h, int) { h_ = h; }
  __device__ __forceinline__ Edge load_edge(int e, int c) const {
    const int64_t p = (int64_t)__ldg(perm + e) * heads + h_;
    const float2 ad = __ldg(ads_e + p);
    Edge r;
    r.c = c;
    r.w = ad.x;
    r.w2 = ad.y;
    return r;
  }
  __device__ __forceinline__ void apply(float* acc, int, int c, float w, float w2, int gl, bool valid) const {
    if (!valid) { w = 0.f; w2 = 0.f; c = 0; }
    const float* xr = dout + row_off(c, (unsigned)ldd) + h_ * feat;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      float4 v = ldg4(xr + 4 * min(gl + ch * G, f4 - 1));
      if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
      acc[4 * ch + 0] = fmaf(w, v.x, acc[4 * ch + 0]);
      acc[4 * ch + 1] = fmaf(w, v.y, acc[4 * ch + 1]);
      acc[4 * ch + 2] = fmaf(w, v.z, acc[4 * ch + 2]);
      acc[4 * ch + 3] = fmaf(w, v.w, acc[4 * ch + 3]);
    }
    acc[4 * CPL] += w2;
  }
  __device__ __forceinline__ void finish(float* acc, int i, int h, int gl, bool writer) const {
    if (!writer) return;
    const int64_t node = row_offset + i;
    const int64_t gi = node * heads + h;
    const float ws = self_loop ? __ldg(alpha_self + gi) : 0.f;
    if (gl == 0) d_aj[gi] = acc[4 * CPL] + (self_loop ? __ldg(ds_self + gi) : 0.f);
    const float* xi = dout + node * ldd + (int64_t)h * feat;
    float* o = dH + (int64_t)i * ldh + (int64_t)h * feat;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      const int idx = gl + ch * G;
      if (idx < f4) {
        const float4 v = self_loop ? ldg4(xi + 4 * idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 a;
        a.x = fmaf(ws, v.x, acc[4 * ch + 0]);
        a.y = fmaf(ws, v.y, acc[4 * ch + 1]);
        a.z = fmaf(ws, v.z, acc[4 * ch + 2]);
        a.w = fmaf(ws, v.w, acc[4 * ch + 3]);
        st4(o + 4 * idx, a);
      }
    }
  }
};

static int check_gat_common(const peagnn_csr_t* g, int feat, int heads, const char* what) {
  PEAGNN_REQUIRE(g && g->rowptr && (g->col || g->nrows == 0), "%s: null graph", what);
  PEAGNN_REQUIRE(feat > 0 && feat % 4 == 0 && feat <= 256 && heads > 0, "%s: feat=%d (multiple of 4, <= 256), heads=%d", what, feat, heads);
  return PEAGNN_OK;
}

}  // namespace peagnn

using namespace peagnn;

#define PEAGNN_GEOM_DISPATCH(feat_, CALL)             \
  do {                                                \
    const Geometry ge__ = geometry_for(feat_);        \
    if (ge__.G == 4) { CALL(1, 4); }                  \
    else if (ge__.G == 8) { CALL(1, 8); }             \
    else if (ge__.G == 16) { CALL(1, 16); }           \
    else if (ge__.CPL == 1) { CALL(1, 32); }          \
    else { CALL(2, 32); }                             \
  } while (0)

extern "C" int peagnn_gat_rowmax(const peagnn_csr_t* g, const float* a_i, const float* a_j, int32_t heads,
                                 float slope, float* rowmax, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_gat_common(g, 4, heads, "peagnn_gat_rowmax");
  if (rc) return rc;
  PEAGNN_REQUIRE(a_i && a_j && rowmax, "peagnn_gat_rowmax: null pointer");
  if (g->nrows == 0) return PEAGNN_OK;
  RowMaxOp op;
  op.heads = heads; op.a_i = a_i; op.a_j = a_j; op.rowmax = rowmax; op.slope = slope;
  op.row_offset = g->row_offset; op.h_ = 0; op.self_loop = !g->explicit_self_loops;
  if (g->active_rows) return launch_csr<RowMaxOp, 1, true>(*g, op, stream, "peagnn_gat_rowmax");
  return launch_csr<RowMaxOp, 1>(*g, op, stream, "peagnn_gat_rowmax");
}

extern "C" int peagnn_gat_aggregate(const peagnn_csr_t* g, const float* H, int64_t ldh, int32_t feat,
                                    int32_t heads, const float* a_i, const float* a_j, float slope,
                                    const float* rowmax, float* denom, float* out, int64_t ldo,
                                    const float* bias, int relu, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_gat_common(g, feat, heads, "peagnn_gat_aggregate");
  if (rc) return rc;
  PEAGNN_REQUIRE(H && a_i && a_j && rowmax && out && ldh % 4 == 0 && ldo % 4 == 0 && aligned16(H) && aligned16(out) && (!bias || aligned16(bias)),
                 "peagnn_gat_aggregate: bad pointers / alignment");
  if (g->nrows == 0) return PEAGNN_OK;
#define CALL(CPL_, G_)                                                                           \
  {                                                                                              \
    GatAggOp<CPL_, G_> op;                                                                       \
    op.heads = heads; op.H = H; op.ldh = ldh; op.feat = feat; op.f4 = feat / 4; op.a_i = a_i;    \
    op.a_j = a_j; op.rowmax = rowmax; op.denom = denom; op.out = out; op.ldo = ldo;              \
    op.bias = bias; op.slope = slope; op.row_offset = g->row_offset; op.relu = relu;             \
    op.ai_ = 0.f; op.m_ = 0.f; op.h_ = 0; op.self_loop = !g->explicit_self_loops;                \
    if (g->active_rows) return launch_csr<GatAggOp<CPL_, G_>, G_, true>(*g, op, stream, "peagnn_gat_aggregate"); \
    return launch_csr<GatAggOp<CPL_, G_>, G_>(*g, op, stream, "peagnn_gat_aggregate");           \
  }
  PEAGNN_GEOM_DISPATCH(feat, CALL);
#undef CALL
  return PEAGNN_OK;
}

extern "C" int peagnn_gat_backward_dst(const peagnn_csr_t* g, const float* H, int64_t ldh, int32_t feat,
                                       int32_t heads, const float* a_i, const float* a_j, float slope,
                                       const float* rowmax, const float* denom, const float* agg,
                                       int64_t lda, const float* agg_bias, const float* dout, int64_t ldd,
                                       float* ads_e, float* alpha_self, float* ds_self,
                                       float* d_ai, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_gat_common(g, feat, heads, "peagnn_gat_backward_dst");
  if (rc) return rc;
  PEAGNN_REQUIRE(H && a_i && a_j && rowmax && denom && agg && dout && d_ai && (g->explicit_self_loops || (alpha_self && ds_self)),
                 "peagnn_gat_backward_dst: null pointer");
  PEAGNN_REQUIRE(ldh % 4 == 0 && lda % 4 == 0 && ldd % 4 == 0 && aligned16(H) && aligned16(agg) && aligned16(dout) &&
                     (reinterpret_cast<uintptr_t>(ads_e) & 7) == 0,
                 "peagnn_gat_backward_dst: alignment");
  PEAGNN_REQUIRE(ads_e || g->nnz == 0, "peagnn_gat_backward_dst: null per-edge buffer");
  PEAGNN_REQUIRE(!g->active_cols, "peagnn_gat_backward_dst: a column filter is not supported on the destination side");
  if (g->nrows == 0) return PEAGNN_OK;
#define CALL(CPL_, G_)                                                                             \
  {                                                                                                \
    GatBwdDstOp<CPL_, G_> op;                                                                      \
    op.heads = heads; op.H = H; op.ldh = ldh; op.feat = feat; op.f4 = feat / 4; op.a_i = a_i;      \
    op.a_j = a_j; op.rowmax = rowmax; op.denom = denom; op.agg = agg; op.lda = lda;                \
    op.agg_bias = agg_bias; op.dout = dout; op.ldd = ldd; op.ads_e = reinterpret_cast<float2*>(ads_e); \
    op.alpha_self = alpha_self; op.ds_self = ds_self; op.d_ai = d_ai; op.slope = slope;            \
    op.row_offset = g->row_offset; op.ai_ = op.m_ = op.inv_den_ = op.D_ = 0.f; op.h_ = 0;          \
    op.self_loop = !g->explicit_self_loops;                                                        \
    for (int q = 0; q < CPL_; ++q) op.g_[q] = make_float4(0.f, 0.f, 0.f, 0.f);                     \
    if (g->active_rows) return launch_csr<GatBwdDstOp<CPL_, G_>, G_, true>(*g, op, stream, "peagnn_gat_backward_dst"); \
    return launch_csr<GatBwdDstOp<CPL_, G_>, G_>(*g, op, stream, "peagnn_gat_backward_dst");       \
  }
  PEAGNN_GEOM_DISPATCH(feat, CALL);
#undef CALL
  return PEAGNN_OK;
}

extern "C" int peagnn_gat_backward_src(const peagnn_csr_t* gt, const int32_t* perm, const float* ads_e,
                                       const float* alpha_self, const float* ds_self,
                                       const float* dout, int64_t ldd, int32_t feat, int32_t heads,
                                       float* dH, int64_t ldh, float* d_aj, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_gat_common(gt, feat, heads, "peagnn_gat_backward_src");
  if (rc) return rc;
  PEAGNN_REQUIRE(dout && dH && d_aj && (gt->explicit_self_loops || (alpha_self && ds_self)), "peagnn_gat_backward_src: null pointer");
  PEAGNN_REQUIRE(ldh % 4 == 0 && ldd % 4 == 0 && aligned16(dH) && aligned16(dout) && (reinterpret_cast<uintptr_t>(ads_e) & 7) == 0,
                 "peagnn_gat_backward_src: alignment");
  if (gt->nrows == 0) return PEAGNN_OK;
#define CALL(CPL_, G_)                                                                              \
  {                                                                                                 \
    GatBwdSrcOp<CPL_, G_> op;                                                                       \
    op.heads = heads; op.perm = perm; op.ads_e = reinterpret_cast<const float2*>(ads_e);            \
    op.alpha_self = alpha_self; op.ds_self = ds_self; op.dout = dout; op.ldd = ldd;                 \
    op.feat = feat; op.f4 = feat / 4; op.dH = dH; op.ldh = ldh; op.d_aj = d_aj;                     \
    op.row_offset = gt->row_offset; op.h_ = 0; op.self_loop = !gt->explicit_self_loops;             \
    if (gt->active_cols || gt->active_rows) return launch_csr<GatBwdSrcOp<CPL_, G_>, G_, true>(*gt, op, stream, "peagnn_gat_backward_src"); \
    return launch_csr<GatBwdSrcOp<CPL_, G_>, G_>(*gt, op, stream, "peagnn_gat_backward_src");       \
  }
  PEAGNN_GEOM_DISPATCH(feat, CALL);
#undef CALL
  return PEAGNN_OK;
}
