// K1 / K2: weighted CSR aggregation - the message passing of GCNConv and SAGEConv
// (reference call sites models/peagcn.py:16-21, models/peasage.py:16-21 via models/base.py:137-139;
// PyG-1.5.0 gather -> scale -> torch_scatter.scatter_add restated as a gather-side segmented sum).
#include "csr_traverse.cuh"

namespace peagnn {

template <int CPL, int G>
struct SpmmOp {
  static constexpr int NV = 4 * CPL;
  static constexpr bool kMax = false;
  static constexpr bool kUseW2 = false;
  int heads;
  const float* __restrict__ X;
  int64_t ldx;
  int f4;  // float4 chunks per row
  float* __restrict__ out;
  int64_t ldo;
  const float* __restrict__ rs;
  const float* __restrict__ cs;
  const float* __restrict__ bias;
  int row_offset;
  int self_loop, relu, accumulate;

  __device__ __forceinline__ void row_begin(int, int, int) {}

  __device__ __forceinline__ Edge load_edge(int, int c) const {
    Edge e;
    e.c = c;
    e.w = cs ? __ldg(cs + c) : 1.f;
    e.w2 = 0.f;
    return e;
  }

  __device__ __forceinline__ void apply(float* acc, int, int c, float w, float, int gl, bool valid) const {
    const float* xr = X + row_off(c, (unsigned)ldx);
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      const int idx = gl + ch * G;
      if (valid && idx < f4) {
        const float4 v = ldg4(xr + 4 * idx);
        acc[4 * ch + 0] = fmaf(w, v.x, acc[4 * ch + 0]);
        acc[4 * ch + 1] = fmaf(w, v.y, acc[4 * ch + 1]);
        acc[4 * ch + 2] = fmaf(w, v.z, acc[4 * ch + 2]);
        acc[4 * ch + 3] = fmaf(w, v.w, acc[4 * ch + 3]);
      }
    }
  }

  __device__ __forceinline__ void finish(float* acc, int i, int, int gl, bool writer) const {
    if (!writer) return;
    const int gi = row_offset + i;
    const float r = rs ? __ldg(rs + gi) : 1.f;
    const float sw = self_loop ? (cs ? __ldg(cs + gi) : 1.f) : 0.f;
    const float* xi = X + (int64_t)gi * ldx;
    float* o = out + (int64_t)i * ldo;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      const int idx = gl + ch * G;
      if (idx < f4) {
        float4 a = make_float4(acc[4 * ch], acc[4 * ch + 1], acc[4 * ch + 2], acc[4 * ch + 3]);
        if (self_loop) {
          const float4 v = ldg4(xi + 4 * idx);
          a.x = fmaf(sw, v.x, a.x);
          a.y = fmaf(sw, v.y, a.y);
          a.z = fmaf(sw, v.z, a.z);
          a.w = fmaf(sw, v.w, a.w);
        }
        a.x *= r; a.y *= r; a.z *= r; a.w *= r;
        if (bias) {
          const float4 b = ldg4(bias + 4 * idx);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        if (accumulate) {
          const float4 p = *reinterpret_cast<const float4*>(o + 4 * idx);
          a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        if (relu) {
          a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
        }
        st4(o + 4 * idx, a);
      }
    }
  }
};

template <int CPL, int G, bool FILT>
static int run_spmm(const peagnn_csr_t& g, const float* X, int64_t ldx, int feat, float* out,
                    int64_t ldo, const float* rs, const float* cs, int self_loop, const float* bias,
                    int relu, int accumulate, cudaStream_t stream) {
  SpmmOp<CPL, G> op;
  op.heads = 1;
  op.X = X; op.ldx = ldx; op.f4 = feat / 4; op.out = out; op.ldo = ldo;
  op.rs = rs; op.cs = cs; op.bias = bias; op.row_offset = g.row_offset;
  op.self_loop = self_loop && !g.explicit_self_loops; op.relu = relu; op.accumulate = accumulate;
  return launch_csr<SpmmOp<CPL, G>, G, FILT>(g, op, stream, FILT ? "peagnn_spmm_filtered" : "peagnn_spmm");
}

}  // namespace peagnn

using namespace peagnn;

extern "C" size_t peagnn_partial_floats(int32_t n_chunks, int32_t feat, int32_t heads) {
  const Geometry ge = geometry_for(feat);
  return (size_t)(n_chunks > 0 ? n_chunks : 0) * (size_t)(heads > 0 ? heads : 1) * (size_t)ge.G *
         (size_t)(4 * ge.CPL + 2);
}

template <bool FILT>
static int spmm_dispatch(const peagnn_csr_t& g, const float* X, int64_t ldx, int32_t feat, float* out, int64_t ldo,
                         const float* rs, const float* cs, int self_loop, const float* bias, int relu, int accumulate,
                         cudaStream_t stream, const char* what) {
  PEAGNN_REQUIRE(g.rowptr && (g.col || g.nrows == 0), "%s: null graph", what);
  PEAGNN_REQUIRE(feat > 0 && feat % 4 == 0 && feat <= 512, "%s: feat=%d must be a multiple of 4, <= 512", what, feat);
  PEAGNN_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0 && ldx >= feat && ldo >= feat, "%s: leading dims must be multiples of 4 and >= feat", what);
  PEAGNN_REQUIRE(aligned16(X) && aligned16(out) && (!bias || aligned16(bias)), "%s: pointers must be 16-byte aligned", what);
  if (g.nrows == 0) return PEAGNN_OK;
  const Geometry ge = geometry_for(feat);
#define PEAGNN_SPMM_CASE(CPL_, G_) \
  return run_spmm<CPL_, G_, FILT>(g, X, ldx, feat, out, ldo, rs, cs, self_loop, bias, relu, accumulate, stream)
  if (ge.G == 4) PEAGNN_SPMM_CASE(1, 4);
  if (ge.G == 8) PEAGNN_SPMM_CASE(1, 8);
  if (ge.G == 16) PEAGNN_SPMM_CASE(1, 16);
  if (ge.CPL == 1) PEAGNN_SPMM_CASE(1, 32);
  if (ge.CPL == 2) PEAGNN_SPMM_CASE(2, 32);
  PEAGNN_SPMM_CASE(4, 32);
#undef PEAGNN_SPMM_CASE
}

extern "C" int peagnn_spmm(const peagnn_csr_t* g, const float* X, int64_t ldx, int32_t feat,
                           float* out, int64_t ldo, const float* rs, const float* cs,
                           int self_loop, const float* bias, int relu, int accumulate,
                           peagnn_stream_t stream_) {
  PEAGNN_REQUIRE(g, "peagnn_spmm: null graph");
  peagnn_csr_t view = *g;
  view.active_rows = view.active_cols = nullptr;
  return spmm_dispatch<false>(view, X, ldx, feat, out, ldo, rs, cs, self_loop, bias, relu, accumulate,
                              static_cast<cudaStream_t>(stream_), "peagnn_spmm");
}

extern "C" int peagnn_spmm_filtered(const peagnn_csr_t* g, const float* X, int64_t ldx, int32_t feat,
                                    float* out, int64_t ldo, const float* rs, const float* cs, int self_loop,
                                    const float* bias, int relu, int accumulate, const uint32_t* active_rows,
                                    const uint32_t* active_cols, peagnn_stream_t stream_) {
  PEAGNN_REQUIRE(g, "peagnn_spmm_filtered: null graph");
  peagnn_csr_t view = *g;
  view.active_rows = active_rows;
  view.active_cols = active_cols;
  if (!active_rows && !active_cols)
    return spmm_dispatch<false>(view, X, ldx, feat, out, ldo, rs, cs, self_loop, bias, relu, accumulate,
                                static_cast<cudaStream_t>(stream_), "peagnn_spmm_filtered");
  return spmm_dispatch<true>(view, X, ldx, feat, out, ldo, rs, cs, self_loop, bias, relu, accumulate,
                             static_cast<cudaStream_t>(stream_), "peagnn_spmm_filtered");
}

namespace peagnn {
__global__ void mark_rows_kernel(const int64_t* __restrict__ ids, int64_t n, int mod, int rem, uint32_t* __restrict__ bitmap) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t id = ids[k];
  if (mod > 1) {
    if (id % mod != rem) return;
    id /= mod;
  }
  atomicOr(bitmap + (id >> 5), 1u << (id & 31));     // integer OR: order-independent, still deterministic
}
}  // namespace peagnn

extern "C" int peagnn_mark_rows(const int64_t* ids, int64_t n, int32_t mod, int32_t rem, uint32_t* bitmap,
                                peagnn_stream_t stream_) {
  PEAGNN_REQUIRE(ids && bitmap && n >= 0, "peagnn_mark_rows: bad arguments");
  if (n == 0) return PEAGNN_OK;
  mark_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(ids, n, mod, rem, bitmap);
  return check_launch("peagnn_mark_rows");
}
