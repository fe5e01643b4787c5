// K6: pair scoring, BPR loss and the entity-aware regulariser
// (reference models/base.py:208-214 predict, :43-48 BPR, :50-76 entity-aware term).
// These kernels touch only B <= a few thousand rows: they are latency-bound, so the design goal
// is few launches and no host synchronisation, not bandwidth.
#include "common.cuh"
#include "scatter_sorted.cuh"

namespace peagnn {

extern "C" int peagnn_linear_wgrad(const float*, int64_t, const float*, int64_t, const float*, int64_t,
                                   int64_t, int32_t, int32_t, int, float*, float*, float*, size_t,
                                   peagnn_stream_t);
extern "C" size_t peagnn_wgrad_workspace_floats(int64_t, int32_t, int32_t);

// The fc1 loops are unrolled completely for the shipped repr_dim (16) and below; wider layers keep the
// inner loop unrolled only (complete unrolling of 64 x 64 x 3 FMAs takes ptxas minutes for no gain).
template <int D>
constexpr int kOuterUnroll = D <= 16 ? D : 1;

template <int D>
struct FcSmem {
  float w1[D * 2 * D];  // [D][2D] (out, in)
  float b1[D];
  float w2[D];
  float b2;
};

template <int D>
__device__ __forceinline__ void load_fc(FcSmem<D>& s, const float* w1, const float* b1, const float* w2,
                                        const float* b2) {
  for (int i = threadIdx.x; i < D * 2 * D; i += blockDim.x) s.w1[i] = __ldg(w1 + i);
  for (int i = threadIdx.x; i < D; i += blockDim.x) { s.b1[i] = __ldg(b1 + i); s.w2[i] = __ldg(w2 + i); }
  if (threadIdx.x == 0) s.b2 = __ldg(b2);
  __syncthreads();
}

template <int D>
__device__ __forceinline__ void load_row(const float* repr, int64_t ldr, int64_t id, float* r) {
#pragma unroll
  for (int c = 0; c < D / 4; ++c) {
    const float4 v = ldg4(repr + id * ldr + 4 * c);
    r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
  }
}

// hu[j] = b1[j] + sum_k W1[j][k] u[k]           (user half of fc1, shared by pos and neg)
template <int D>
__device__ __forceinline__ void user_half(const FcSmem<D>& s, const float* u, float* hu) {
#pragma unroll(kOuterUnroll<D>)
  for (int j = 0; j < D; ++j) {
    float a = s.b1[j];
#pragma unroll
    for (int k = 0; k < D; ++k) a = fmaf(s.w1[j * 2 * D + k], u[k], a);
    hu[j] = a;
  }
}
// score = b2 + sum_j w2[j] relu(hu[j] + sum_k W1[j][D+k] it[k]); optionally returns pre-activations
template <int D>
__device__ __forceinline__ float item_score(const FcSmem<D>& s, const float* hu, const float* it, float* pre) {
  float sc = s.b2;
#pragma unroll(kOuterUnroll<D>)
  for (int j = 0; j < D; ++j) {
    float a = hu[j];
#pragma unroll
    for (int k = 0; k < D; ++k) a = fmaf(s.w1[j * 2 * D + D + k], it[k], a);
    if (pre) pre[j] = a;
    sc = fmaf(s.w2[j], fmaxf(a, 0.f), sc);
  }
  return sc;
}

template <int D>
__global__ void __launch_bounds__(128) predict_kernel(const float* __restrict__ repr, int64_t ldr,
                                                      const int64_t* __restrict__ unids, const int64_t* __restrict__ inids,
                                                      int64_t B, const float* w1, const float* b1, const float* w2,
                                                      const float* b2, float* __restrict__ scores) {
  __shared__ FcSmem<D> s;
  load_fc<D>(s, w1, b1, w2, b2);
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float u[D], it[D], hu[D];
  load_row<D>(repr, ldr, unids[b], u);
  load_row<D>(repr, ldr, inids[b], it);
  user_half<D>(s, u, hu);
  scores[b] = item_score<D>(s, hu, it, nullptr);
}

// Workspace layout for the gradient path (all [2B, *], first B rows = positive branch, last B = negative):
//   cat  [2B, 2D]  fc1 inputs            dh [2B, D]  d loss / d fc1 pre-activation
//   hid  [2B, D]   relu(fc1) outputs     g4 [2B, 4]  column 0 = d loss / d score
template <int D>
__global__ void __launch_bounds__(128) bpr_kernel(const float* __restrict__ repr, int64_t ldr,
                                                  const int64_t* __restrict__ batch, int cols, int64_t B,
                                                  const float* w1, const float* b1, const float* w2, const float* b2,
                                                  float* __restrict__ loss_terms, int need_grad,
                                                  int32_t* __restrict__ gkeys, float* __restrict__ grows,
                                                  float* __restrict__ cat, float* __restrict__ dh,
                                                  float* __restrict__ hid, float* __restrict__ g4) {
  __shared__ FcSmem<D> s;
  load_fc<D>(s, w1, b1, w2, b2);
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t uid = batch[b * cols], pid = batch[b * cols + 1], nid = batch[b * cols + 2];
  float u[D], hu[D];
  load_row<D>(repr, ldr, uid, u);
  user_half<D>(s, u, hu);
  float it[D], pre_p[D], pre_n[D];
  load_row<D>(repr, ldr, pid, it);
  const float sp = item_score<D>(s, hu, it, pre_p);
  float itn[D];
  load_row<D>(repr, ldr, nid, itn);
  const float sn = item_score<D>(s, hu, itn, pre_n);
  const float z = sp - sn;
  loss_terms[b] = neg_log_sigmoid(z);
  if (!need_grad) return;

  const float g = neg_log_sigmoid_grad(z);  // d(-log sigmoid(z))/dz ; d/dsp = g, d/dsn = -g
  float du[D], dp[D], dn[D];
#pragma unroll
  for (int k = 0; k < D; ++k) du[k] = dp[k] = dn[k] = 0.f;
#pragma unroll(kOuterUnroll<D>)
  for (int j = 0; j < D; ++j) {
    const float dhp = pre_p[j] > 0.f ? g * s.w2[j] : 0.f;
    const float dhn = pre_n[j] > 0.f ? -g * s.w2[j] : 0.f;
    dh[b * D + j] = dhp;
    dh[(B + b) * D + j] = dhn;
    hid[b * D + j] = fmaxf(pre_p[j], 0.f);
    hid[(B + b) * D + j] = fmaxf(pre_n[j], 0.f);
    const float dsum = dhp + dhn;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      du[k] = fmaf(dsum, s.w1[j * 2 * D + k], du[k]);
      dp[k] = fmaf(dhp, s.w1[j * 2 * D + D + k], dp[k]);
      dn[k] = fmaf(dhn, s.w1[j * 2 * D + D + k], dn[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < D; ++k) {
    cat[b * 2 * D + k] = u[k];
    cat[b * 2 * D + D + k] = it[k];
    cat[(B + b) * 2 * D + k] = u[k];
    cat[(B + b) * 2 * D + D + k] = itn[k];
  }
  g4[b * 4] = g; g4[b * 4 + 1] = 0.f; g4[b * 4 + 2] = 0.f; g4[b * 4 + 3] = 0.f;
  g4[(B + b) * 4] = -g; g4[(B + b) * 4 + 1] = 0.f; g4[(B + b) * 4 + 2] = 0.f; g4[(B + b) * 4 + 3] = 0.f;
  // row gradients go to slots b (user), B + b (positive item), 2B + b (negative item) of a [3B, D] scratch;
  // scatter_rows_sorted() adds them into d_repr in a fixed order (no float atomics: bit-reproducible)
  gkeys[b] = (int32_t)uid; gkeys[B + b] = (int32_t)pid; gkeys[2 * B + b] = (int32_t)nid;
#pragma unroll
  for (int c = 0; c < D / 4; ++c) {
    st4(grows + (size_t)b * D + 4 * c, make_float4(du[4 * c], du[4 * c + 1], du[4 * c + 2], du[4 * c + 3]));
    st4(grows + (size_t)(B + b) * D + 4 * c, make_float4(dp[4 * c], dp[4 * c + 1], dp[4 * c + 2], dp[4 * c + 3]));
    st4(grows + (size_t)(2 * B + b) * D + 4 * c, make_float4(dn[4 * c], dn[4 * c + 1], dn[4 * c + 2], dn[4 * c + 3]));
  }
}

// Fixed-order sum of `n` floats by one CTA: out = (accumulate ? out : 0) + scale * sum.
__global__ void __launch_bounds__(1024) sum_terms_kernel(const float* __restrict__ terms, int64_t n, float scale,
                                                         int accumulate, float* __restrict__ out) {
  __shared__ float red[1024];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += terms[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + scale * red[0];
}

__global__ void copy_row0_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, float* __restrict__ zero1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
  if (i == 0 && zero1) zero1[0] = 0.f;
}

// ---- entity-aware regulariser ---------------------------------------------------------------
// One 32-lane warp per sample; lane l holds float4 chunk l of each embedding row (emb <= 128).
__global__ void __launch_bounds__(256) entity_kernel(const float* __restrict__ x, int64_t ldx, int emb,
                                                     const int64_t* __restrict__ batch, int64_t B, float coff,
                                                     float* __restrict__ terms, int need_grad,
                                                     int32_t* __restrict__ gkeys, float* __restrict__ grows) {
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  if (b >= B) return;
  const bool on = 4 * lane < emb;
  const int64_t* row = batch + b * 9;
  float total = 0.f;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int64_t anchor = side == 0 ? row[1] : row[0];
    const int64_t epos = side == 0 ? row[3] : row[6];
    const int64_t eneg = side == 0 ? row[4] : row[7];
    const float mask = (float)(side == 0 ? row[5] : row[8]);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), p = a, n = a;
    if (on) {
      a = ldg4(x + anchor * ldx + 4 * lane);
      p = ldg4(x + epos * ldx + 4 * lane);
      n = ldg4(x + eneg * ldx + 4 * lane);
    }
    const float4 dp = make_float4(a.x - p.x, a.y - p.y, a.z - p.z, a.w - p.w);
    const float4 dn = make_float4(a.x - n.x, a.y - n.y, a.z - n.z, a.w - n.w);
    float sp = dp.x * dp.x + dp.y * dp.y + dp.z * dp.z + dp.w * dp.w;
    float sn = dn.x * dn.x + dn.y * dn.y + dn.z * dn.z + dn.w * dn.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sp += __shfl_xor_sync(0xffffffffu, sp, o);
      sn += __shfl_xor_sync(0xffffffffu, sn, o);
    }
    const float z = (sp - sn) * mask;
    total += neg_log_sigmoid(z);
    if (need_grad) {
      // slots (side, b, {anchor, e+, e-}) of a [6B, emb] scratch, summed into dx in a fixed order afterwards
      const int64_t slot = ((int64_t)side * B + b) * 3;
      const bool live = mask != 0.f;
      if (lane == 0) {
        gkeys[slot] = live ? (int32_t)anchor : kScatterSkip;
        gkeys[slot + 1] = live ? (int32_t)epos : kScatterSkip;
        gkeys[slot + 2] = live ? (int32_t)eneg : kScatterSkip;
      }
      if (on && live) {
        const float gz = neg_log_sigmoid_grad(z) * mask * coff * 2.f;   // d/d(sp) = gz/2 * 2 (from the square)
        // d anchor = gz * (dp - dn) ; d epos = -gz * dp ; d eneg = +gz * dn
        st4(grows + (size_t)slot * emb + 4 * lane,
            make_float4(gz * (dp.x - dn.x), gz * (dp.y - dn.y), gz * (dp.z - dn.z), gz * (dp.w - dn.w)));
        st4(grows + (size_t)(slot + 1) * emb + 4 * lane, make_float4(-gz * dp.x, -gz * dp.y, -gz * dp.z, -gz * dp.w));
        st4(grows + (size_t)(slot + 2) * emb + 4 * lane, make_float4(gz * dn.x, gz * dn.y, gz * dn.z, gz * dn.w));
      }
    }
  }
  if (lane == 0) terms[b] = total;
}

template <int D>
static int run_predict(const float* repr, int64_t ldr, const int64_t* unids, const int64_t* inids, int64_t B,
                       const float* w1, const float* b1, const float* w2, const float* b2, float* scores,
                       cudaStream_t stream) {
  predict_kernel<D><<<(unsigned)((B + 127) / 128), 128, 0, stream>>>(repr, ldr, unids, inids, B, w1, b1, w2, b2, scores);
  return check_launch("peagnn_predict");
}

template <int D>
static int run_bpr(const float* repr, int64_t ldr, const int64_t* batch, int cols, int64_t B, const float* w1,
                   const float* b1, const float* w2, const float* b2, float* loss, int need_grad, float* d_repr,
                   int64_t lddr, float* d_w1, float* d_b1, float* d_w2, float* d_b2, float* ws, size_t ws_floats,
                   cudaStream_t stream) {
  float* terms = ws;
  size_t off = ((size_t)B + 3) / 4 * 4;
  float* cat = ws + off; off += (size_t)2 * B * 2 * D;
  float* dh = ws + off; off += (size_t)2 * B * D;
  float* hid = ws + off; off += (size_t)2 * B * D;
  float* g4 = ws + off; off += (size_t)2 * B * 4;
  float* tmp = ws + off; off += 4 * D + 4;   // [4, D] result of the fc2 weight gradient + db scratch
  float* grows = ws + off; off += (size_t)3 * B * D;
  int32_t* gkeys = reinterpret_cast<int32_t*>(ws + off); off += ((size_t)3 * B + 3) / 4 * 4;
  float* wg = ws + off;       // the weight-gradient scratch and the scatter's sort scratch take turns here
  const size_t wg_floats = ws_floats - off;
  bpr_kernel<D><<<(unsigned)((B + 127) / 128), 128, 0, stream>>>(repr, ldr, batch, cols, B, w1, b1, w2, b2, terms,
                                                               need_grad, gkeys, grows, cat, dh, hid, g4);
  int rc = check_launch("peagnn_bpr_loss");
  if (rc) return rc;
  sum_terms_kernel<<<1, 1024, 0, stream>>>(terms, B, 1.f, 0, loss);
  rc = check_launch("peagnn_bpr_loss(sum)");
  if (rc || !need_grad) return rc;
  rc = scatter_rows_sorted(gkeys, grows, D, 3 * B, d_repr, lddr, wg, stream, "peagnn_bpr_loss(scatter)");
  if (rc) return rc;
  // d fc1.weight [D, 2D] = dh^T @ cat ; d fc1.bias = colsum(dh)
  rc = peagnn_linear_wgrad(cat, 2 * D, dh, D, nullptr, 0, 2 * B, 2 * D, D, /*out_in=*/1, d_w1, d_b1, wg, wg_floats, stream);
  if (rc) return rc;
  // d fc2.weight [1, D] = row 0 of ( g4^T @ hid ) ; d fc2.bias = g - g = 0
  rc = peagnn_linear_wgrad(hid, D, g4, 4, nullptr, 0, 2 * B, D, 4, /*out_in=*/1, tmp, nullptr, wg, wg_floats, stream);
  if (rc) return rc;
  copy_row0_kernel<<<1, 128, 0, stream>>>(tmp, D, d_w2, d_b2);
  return check_launch("peagnn_bpr_loss(fc2)");
}

}  // namespace peagnn

using namespace peagnn;

static int bpr_check_dim(int D, const char* what) {
  PEAGNN_REQUIRE(D == 8 || D == 16 || D == 32 || D == 64, "%s: repr_dim %d not in {8,16,32,64}", what, D);
  return PEAGNN_OK;
}

extern "C" int peagnn_predict(const float* repr, int64_t ldr, int32_t D, const int64_t* unids,
                              const int64_t* inids, int64_t B, const float* fc1_w, const float* fc1_b,
                              const float* fc2_w, const float* fc2_b, float* scores, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = bpr_check_dim(D, "peagnn_predict");
  if (rc) return rc;
  PEAGNN_REQUIRE(repr && unids && inids && fc1_w && fc1_b && fc2_w && fc2_b && scores && ldr % 4 == 0 && aligned16(repr),
                 "peagnn_predict: bad pointers");
  if (B == 0) return PEAGNN_OK;
  switch (D) {
    case 8: return run_predict<8>(repr, ldr, unids, inids, B, fc1_w, fc1_b, fc2_w, fc2_b, scores, stream);
    case 16: return run_predict<16>(repr, ldr, unids, inids, B, fc1_w, fc1_b, fc2_w, fc2_b, scores, stream);
    case 32: return run_predict<32>(repr, ldr, unids, inids, B, fc1_w, fc1_b, fc2_w, fc2_b, scores, stream);
    default: return run_predict<64>(repr, ldr, unids, inids, B, fc1_w, fc1_b, fc2_w, fc2_b, scores, stream);
  }
}

extern "C" size_t peagnn_bpr_workspace_floats(int64_t B, int32_t D) {
  const size_t fixed = ((size_t)B + 3) / 4 * 4 + (size_t)2 * B * (2 * D + D + D + 4) + 4 * D + 4 +
                       (size_t)3 * B * D + ((size_t)3 * B + 3) / 4 * 4;
  const size_t w1 = peagnn_wgrad_workspace_floats(2 * B, 2 * D, D);
  const size_t w2 = peagnn_wgrad_workspace_floats(2 * B, D, 4);
  const size_t sc = (scatter_scratch_bytes(3 * B) + 3) / 4;
  size_t turn = w1 > w2 ? w1 : w2;
  if (sc > turn) turn = sc;
  return fixed + turn + 64;
}

extern "C" int peagnn_bpr_loss(const float* repr, int64_t ldr, int32_t D, const int64_t* batch,
                               int32_t batch_cols, int64_t B, const float* fc1_w, const float* fc1_b,
                               const float* fc2_w, const float* fc2_b, float* loss, int need_grad,
                               float* d_repr, int64_t lddr, float* d_fc1_w, float* d_fc1_b, float* d_fc2_w,
                               float* d_fc2_b, float* workspace, size_t workspace_floats,
                               peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = bpr_check_dim(D, "peagnn_bpr_loss");
  if (rc) return rc;
  PEAGNN_REQUIRE(repr && batch && fc1_w && fc1_b && fc2_w && fc2_b && loss && workspace && batch_cols >= 3 && B > 0 &&
                     ldr % 4 == 0 && aligned16(repr) && aligned16(workspace),
                 "peagnn_bpr_loss: bad arguments");
  PEAGNN_REQUIRE(!need_grad || (d_repr && d_fc1_w && d_fc1_b && d_fc2_w && d_fc2_b && lddr % 4 == 0 && aligned16(d_repr)),
                 "peagnn_bpr_loss: gradient buffers missing");
  if (workspace_floats < peagnn_bpr_workspace_floats(B, D)) {
    set_error("peagnn_bpr_loss: workspace %zu < %zu floats", workspace_floats, peagnn_bpr_workspace_floats(B, D));
    return PEAGNN_ERR_WORKSPACE;
  }
#define CASE(D_) case D_: return run_bpr<D_>(repr, ldr, batch, batch_cols, B, fc1_w, fc1_b, fc2_w, fc2_b, loss, need_grad, d_repr, lddr, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b, workspace, workspace_floats, stream)
  switch (D) { CASE(8); CASE(16); CASE(32); default: return run_bpr<64>(repr, ldr, batch, batch_cols, B, fc1_w, fc1_b, fc2_w, fc2_b, loss, need_grad, d_repr, lddr, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b, workspace, workspace_floats, stream); }
#undef CASE
}

extern "C" size_t peagnn_entity_workspace_floats(int64_t B, int32_t emb, int need_grad) {
  const size_t terms = ((size_t)B + 3) / 4 * 4;
  if (!need_grad) return terms;
  return terms + (size_t)6 * B * emb + ((size_t)6 * B + 3) / 4 * 4 + (scatter_scratch_bytes(6 * B) + 3) / 4 + 64;
}

extern "C" int peagnn_entity_reg(const float* x, int64_t ldx, int32_t emb, const int64_t* batch, int64_t B,
                                 float coff, float* loss, int need_grad, float* dx, int64_t lddx,
                                 float* workspace, size_t workspace_floats, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(x && batch && loss && workspace && emb > 0 && emb % 4 == 0 && emb <= 128 && ldx % 4 == 0 && aligned16(x) && B > 0,
                 "peagnn_entity_reg: emb must be a multiple of 4, <= 128");
  PEAGNN_REQUIRE(!need_grad || (dx && lddx % 4 == 0 && aligned16(dx)), "peagnn_entity_reg: dx missing");
  const size_t need = peagnn_entity_workspace_floats(B, emb, need_grad);
  PEAGNN_REQUIRE(workspace_floats >= need && aligned16(workspace), "peagnn_entity_reg: workspace %zu < %zu floats",
                 workspace_floats, need);
  const size_t terms_f = ((size_t)B + 3) / 4 * 4;
  float* grows = workspace + terms_f;
  int32_t* gkeys = reinterpret_cast<int32_t*>(grows + (size_t)6 * B * emb);
  float* scratch = reinterpret_cast<float*>(gkeys) + ((size_t)6 * B + 3) / 4 * 4;
  entity_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, stream>>>(x, ldx, emb, batch, B, coff, workspace, need_grad,
                                                                     gkeys, grows);
  int rc = check_launch("peagnn_entity_reg");
  if (rc) return rc;
  sum_terms_kernel<<<1, 1024, 0, stream>>>(workspace, B, coff, 1, loss);
  rc = check_launch("peagnn_entity_reg(sum)");
  if (rc || !need_grad) return rc;
  return scatter_rows_sorted(gkeys, grows, emb, 6 * B, dx, lddx, scratch, stream, "peagnn_entity_reg(scatter)");
}
