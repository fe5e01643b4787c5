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
  // per-thread registers: this lane's chunk(s) of a row, as byte-ready pointers.  A lane beyond the row's width
  // (widths that do not fill the slot) re-reads the last chunk and its sums are never stored, so the gather loop
  // carries no per-lane predicate - one IMAD.WIDE + LDG.128 + 4 FFMA per edge and slot (the predicated form compiled
  // to 22 instructions per pair of edges, half of them register moves).
  const char* xl_[CPL];   // byte pointers: address of a gathered chunk = xl_ + c * ld_ (bytes) is ONE IMAD.WIDE.U32
  unsigned ld_;

  __device__ __forceinline__ void row_begin(int, int, int gl) {
    ld_ = (unsigned)ldx * 4u;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) xl_[ch] = reinterpret_cast<const char*>(X + 4 * min(gl + ch * G, f4 - 1));
  }

  __device__ __forceinline__ Edge load_edge(int, int c) const {
    Edge e;
    e.c = c;
    e.w = cs ? __ldg(cs + c) : 1.f;
    e.w2 = 0.f;
    return e;
  }

  // `valid` is the literal true on the full-batch path (the select below folds away); on a row's last, partial batch
  // the idle slots gather the safe row and are zeroed after the load, so a non-finite value there cannot leak in.
  __device__ __forceinline__ void apply(float* acc, int, int c, float w, float, int, bool valid) const {
    if (!valid) c = 0;      // an idle slot's "safe row" is a ROW id of this view; the gathered table may be shorter (shard transposes)
    const unsigned long long off = (unsigned long long)(unsigned)c * (unsigned long long)ld_;
#pragma unroll
    for (int ch = 0; ch < CPL; ++ch) {
      float4 v = __ldg(reinterpret_cast<const float4*>(xl_[ch] + off));
      if (!valid) v = make_float4(0.f, 0.f, 0.f, 0.f);
      acc[4 * ch + 0] = fmaf(w, v.x, acc[4 * ch + 0]);
      acc[4 * ch + 1] = fmaf(w, v.y, acc[4 * ch + 1]);
      acc[4 * ch + 2] = fmaf(w, v.z, acc[4 * ch + 2]);
      acc[4 * ch + 3] = fmaf(w, v.w, acc[4 * ch + 3]);
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

// north_star (2): the channel's first projection fused into the aggregation epilogue.  While the aggregated 64-float row
// is still in the slot's registers (16 lanes x float4), every lane takes 4 output columns of  H = relu(a W + b):
// the row's values travel inside the slot with width-16 shuffles, W ([64, 64] row-major, 16 KB, L1-resident) is read with
// 128-bit loads, fp32 FMA in k order - exact fp32, no tensor-core split needed for 8 kFLOP per row.  Up to two metapaths
// can share a first-step relation, so up to two (W, b, H) triples are served from one aggregated row.  The aggregate
// itself is still written (the backward pass needs it for d W).
struct SpmmProjOp : SpmmOp<1, 16> {
  const float* __restrict__ PW[2];
  const float* __restrict__ Pb[2];
  float* __restrict__ PH[2];
  int64_t ldh;
  int n_proj;

  __device__ __forceinline__ void finish(float* acc, int i, int, int gl, bool writer) const {
    const int gi = row_offset + i;
    const float r = rs ? __ldg(rs + gi) : 1.f;
    float4 a = make_float4(acc[0], acc[1], acc[2], acc[3]);
    if (self_loop) {
      const float sw = cs ? __ldg(cs + gi) : 1.f;
      const float4 v = ldg4(X + (int64_t)gi * ldx + 4 * gl);
      a.x = fmaf(sw, v.x, a.x); a.y = fmaf(sw, v.y, a.y); a.z = fmaf(sw, v.z, a.z); a.w = fmaf(sw, v.w, a.w);
    }
    a.x *= r; a.y *= r; a.z *= r; a.w *= r;
    if (writer) st4(out + (int64_t)i * ldo + 4 * gl, a);
    for (int q = 0; q < n_proj; ++q) {
      const float* __restrict__ W = PW[q];
      float4 h = Pb[q] ? ldg4(Pb[q] + 4 * gl) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int k4 = 0; k4 < 16; ++k4) {                 // 4 input columns per step, held by lane k4 of this slot
        const float ax = __shfl_sync(kFull, a.x, k4, 16), ay = __shfl_sync(kFull, a.y, k4, 16);
        const float az = __shfl_sync(kFull, a.z, k4, 16), aw = __shfl_sync(kFull, a.w, k4, 16);
        const float4 w0 = ldg4(W + (4 * k4 + 0) * 64 + 4 * gl), w1 = ldg4(W + (4 * k4 + 1) * 64 + 4 * gl);
        const float4 w2 = ldg4(W + (4 * k4 + 2) * 64 + 4 * gl), w3 = ldg4(W + (4 * k4 + 3) * 64 + 4 * gl);
        h.x = fmaf(ax, w0.x, h.x); h.y = fmaf(ax, w0.y, h.y); h.z = fmaf(ax, w0.z, h.z); h.w = fmaf(ax, w0.w, h.w);
        h.x = fmaf(ay, w1.x, h.x); h.y = fmaf(ay, w1.y, h.y); h.z = fmaf(ay, w1.z, h.z); h.w = fmaf(ay, w1.w, h.w);
        h.x = fmaf(az, w2.x, h.x); h.y = fmaf(az, w2.y, h.y); h.z = fmaf(az, w2.z, h.z); h.w = fmaf(az, w2.w, h.w);
        h.x = fmaf(aw, w3.x, h.x); h.y = fmaf(aw, w3.y, h.y); h.z = fmaf(aw, w3.z, h.z); h.w = fmaf(aw, w3.w, h.w);
      }
      if (relu) { h.x = fmaxf(h.x, 0.f); h.y = fmaxf(h.y, 0.f); h.z = fmaxf(h.z, 0.f); h.w = fmaxf(h.w, 0.f); }
      if (writer) st4(PH[q] + (int64_t)i * ldh + 4 * gl, h);
    }
  }
};

// Same aggregation with the GATHERED table stored in bf16 (opt-in): a 64-float row is 128 bytes,
// one 16-byte load per lane of an 8-lane slot covers 8 elements, so a warp instruction gathers FOUR edges instead of
// two and moves half the L2 bytes.  Accumulation, scales, bias and the output stay fp32.
template <int G>
struct SpmmBf16Op {
  static constexpr int NV = 8;
  static constexpr bool kMax = false;
  static constexpr bool kUseW2 = false;
  int heads;
  const uint16_t* __restrict__ X;   // bf16 bit patterns, row-major
  int64_t ldx;                      // in elements
  int f8;                           // 8-element chunks per row
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
  static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
  }
  __device__ __forceinline__ void apply(float* acc, int, int c, float w, float, int gl, bool valid) const {
    if (!valid) c = 0;
    uint4 v = __ldg(reinterpret_cast<const uint4*>(X + row_off(c, (unsigned)ldx) + 8 * min(gl, f8 - 1)));
    if (!valid) v = make_uint4(0u, 0u, 0u, 0u);       // folds away on the full-batch path (valid is the literal true)
    float f[8];
    unpack(v, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, f[k], acc[k]);
  }
  __device__ __forceinline__ void finish(float* acc, int i, int, int gl, bool writer) const {
    if (!writer || gl >= f8) return;
    const int gi = row_offset + i;
    const float r = rs ? __ldg(rs + gi) : 1.f;
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = acc[k];
    if (self_loop) {
      const float sw = cs ? __ldg(cs + gi) : 1.f;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + (int64_t)gi * ldx + 8 * gl));
      float f[8];
      unpack(v, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] = fmaf(sw, f[k], a[k]);
    }
    float* o = out + (int64_t)i * ldo + 8 * gl;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 q = make_float4(a[4 * h] * r, a[4 * h + 1] * r, a[4 * h + 2] * r, a[4 * h + 3] * r);
      if (bias) {
        const float4 b = ldg4(bias + 8 * gl + 4 * h);
        q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
      }
      if (accumulate) {
        const float4 p = *reinterpret_cast<const float4*>(o + 4 * h);
        q.x += p.x; q.y += p.y; q.z += p.z; q.w += p.w;
      }
      if (relu) { q.x = fmaxf(q.x, 0.f); q.y = fmaxf(q.y, 0.f); q.z = fmaxf(q.z, 0.f); q.w = fmaxf(q.w, 0.f); }
      st4(o + 4 * h, q);
    }
  }
};

// fp32 -> bf16 (round to nearest even), row by row; 8 elements per thread
__global__ void to_bf16_kernel(const float* __restrict__ X, int64_t ldx, int64_t n, int f8, uint16_t* __restrict__ out, int64_t ldo) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t / f8;
  const int c = (int)(t - row * f8);
  if (row >= n) return;
  const float4 a = ldg4(X + row * ldx + 8 * c), b = ldg4(X + row * ldx + 8 * c + 4);
  const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t lo = __float_as_uint(f[2 * k]), hi = __float_as_uint(f[2 * k + 1]);
    lo = (lo + 0x7fffu + ((lo >> 16) & 1u)) >> 16;            // RNE on the dropped 16 bits (finite inputs)
    hi = (hi + 0x7fffu + ((hi >> 16) & 1u)) & 0xffff0000u;
    w[k] = lo | hi;
  }
  *reinterpret_cast<uint4*>(out + row * ldo + 8 * c) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <int CPL, int G, bool FILT>
static int run_spmm(const peagnn_csr_t& g, const float* X, int64_t ldx, int feat, float* out,
                    int64_t ldo, const float* rs, const float* cs, int self_loop, const float* bias,
                    int relu, int accumulate, cudaStream_t stream) {
  SpmmOp<CPL, G> op;
  op.heads = 1;
  op.X = X; op.ldx = ldx; op.f4 = feat / 4; op.out = out; op.ldo = ldo;
  op.rs = rs; op.cs = cs; op.bias = bias; op.row_offset = g.row_offset;
  op.self_loop = self_loop && !g.explicit_self_loops; op.relu = relu; op.accumulate = accumulate;
  op.ld_ = (unsigned)ldx * 4u;
  for (int ch = 0; ch < CPL; ++ch) op.xl_[ch] = reinterpret_cast<const char*>(X);
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

extern "C" int peagnn_spmm_proj(const peagnn_csr_t* g, const float* X, int64_t ldx, float* out, int64_t ldo, const float* rs,
                                const float* cs, int self_loop, const uint32_t* active_rows, int32_t n_proj,
                                const float* W0, const float* b0, float* H0, const float* W1, const float* b1, float* H1,
                                int64_t ldh, int relu, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(g && g->rowptr && (g->col || g->nrows == 0), "peagnn_spmm_proj: null graph");
  PEAGNN_REQUIRE(X && out && ldx % 4 == 0 && ldo % 4 == 0 && ldx >= 64 && ldo >= 64 && aligned16(X) && aligned16(out),
                 "peagnn_spmm_proj: 64-wide tables with 16-byte aligned rows");
  PEAGNN_REQUIRE(n_proj >= 1 && n_proj <= 2 && W0 && H0 && aligned16(W0) && aligned16(H0) && (!b0 || aligned16(b0)) &&
                     ldh % 4 == 0 && ldh >= 64 && (n_proj == 1 || (W1 && H1 && aligned16(W1) && aligned16(H1) && (!b1 || aligned16(b1)))),
                 "peagnn_spmm_proj: one or two (W [64, 64], b, H) triples");
  if (g->nrows == 0) return PEAGNN_OK;
  peagnn_csr_t view = *g;
  view.active_rows = active_rows;
  view.active_cols = nullptr;
  SpmmProjOp op;
  op.heads = 1;
  op.X = X; op.ldx = ldx; op.f4 = 16; op.out = out; op.ldo = ldo;
  op.rs = rs; op.cs = cs; op.bias = nullptr; op.row_offset = view.row_offset;
  op.self_loop = self_loop && !view.explicit_self_loops; op.relu = relu; op.accumulate = 0;
  op.ld_ = (unsigned)ldx * 4u;
  op.xl_[0] = reinterpret_cast<const char*>(X);
  op.PW[0] = W0; op.Pb[0] = b0; op.PH[0] = H0;
  op.PW[1] = W1; op.Pb[1] = b1; op.PH[1] = H1;
  op.ldh = ldh; op.n_proj = n_proj;
  if (active_rows) return launch_csr<SpmmProjOp, 16, true>(view, op, stream, "peagnn_spmm_proj");
  return launch_csr<SpmmProjOp, 16, false>(view, op, stream, "peagnn_spmm_proj");
}

extern "C" int peagnn_to_bf16(const float* X, int64_t ldx, int64_t n, int32_t feat, uint16_t* out, int64_t ldo,
                              peagnn_stream_t stream_) {
  PEAGNN_REQUIRE(X && out && n >= 0 && feat > 0 && feat % 8 == 0 && ldx % 4 == 0 && ldo % 8 == 0 && ldx >= feat && ldo >= feat &&
                     aligned16(X) && aligned16(out),
                 "peagnn_to_bf16: feat must be a multiple of 8, rows 16-byte aligned");
  if (n == 0) return PEAGNN_OK;
  const int64_t threads = n * (feat / 8);
  to_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(X, ldx, n, feat / 8, out, ldo);
  return check_launch("peagnn_to_bf16");
}

extern "C" int peagnn_spmm_bf16(const peagnn_csr_t* g, const uint16_t* Xb, int64_t ldx, int32_t feat, float* out,
                                int64_t ldo, const float* rs, const float* cs, int self_loop, const float* bias, int relu,
                                int accumulate, const uint32_t* active_rows, const uint32_t* active_cols,
                                peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(g && g->rowptr && (g->col || g->nrows == 0), "peagnn_spmm_bf16: null graph");
  PEAGNN_REQUIRE(feat == 64, "peagnn_spmm_bf16: feat=%d (only the 64-wide first-step tables have a bf16 form)", feat);
  PEAGNN_REQUIRE(Xb && out && ldx % 8 == 0 && ldo % 4 == 0 && ldx >= feat && ldo >= feat && aligned16(Xb) && aligned16(out) &&
                     (!bias || aligned16(bias)),
                 "peagnn_spmm_bf16: pointers must be 16-byte aligned, leading dimensions multiples of 8 / 4");
  if (g->nrows == 0) return PEAGNN_OK;
  peagnn_csr_t view = *g;
  view.active_rows = active_rows;
  view.active_cols = active_cols;
  SpmmBf16Op<8> op;
  op.heads = 1;
  op.X = Xb; op.ldx = ldx; op.f8 = feat / 8; op.out = out; op.ldo = ldo;
  op.rs = rs; op.cs = cs; op.bias = bias; op.row_offset = view.row_offset;
  op.self_loop = self_loop && !view.explicit_self_loops; op.relu = relu; op.accumulate = accumulate;
  if (active_rows || active_cols) return launch_csr<SpmmBf16Op<8>, 8, true>(view, op, stream, "peagnn_spmm_bf16");
  return launch_csr<SpmmBf16Op<8>, 8, false>(view, op, stream, "peagnn_spmm_bf16");
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
