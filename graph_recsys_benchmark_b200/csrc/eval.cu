// K7: evaluation scorer + ranker (reference solvers.py:33-104 metrics(), :21-31 candidates,
// utils/rec_utils.py:7-30 hit / ndcg / auc).  One warp per user replaces the reference's
// per-user Python loop (pandas merge, two predict() calls, torch.sort, numpy metrics,
// >= 6 host<->device crossings per user).
//
// Ranking semantics: torch.sort(cat[pos, neg], descending=True) (solvers.py:88) followed by
// hit_vec = indices < n_pos.  The position of candidate c is the number of candidates that sort
// before it; ties are broken by original index (stable order, what the CPU sort the oracle uses
// does), so a positive beats a negative it ties with.
#include "common.cuh"

namespace peagnn {

constexpr int kEvalWarps = 4;
constexpr int kEvalMaxC = 1024;
constexpr int kEvalCols = 36;

template <int D>
__global__ void __launch_bounds__(kEvalWarps * 32) eval_kernel(
    const float* __restrict__ repr, int64_t ldr, const int64_t* __restrict__ users,
    const int64_t* __restrict__ cand, int64_t U, int C, int n_pos, const float* __restrict__ w1g,
    const float* __restrict__ b1g, const float* __restrict__ w2g, const float* __restrict__ b2g,
    double* __restrict__ per_user, float* __restrict__ scores_out) {
  __shared__ float w1[D * 2 * D];
  __shared__ float b1[D], w2[D];
  __shared__ float b2;
  __shared__ float sc[kEvalWarps][kEvalMaxC];
  for (int i = threadIdx.x; i < D * 2 * D; i += blockDim.x) w1[i] = __ldg(w1g + i);
  for (int i = threadIdx.x; i < D; i += blockDim.x) { b1[i] = __ldg(b1g + i); w2[i] = __ldg(w2g + i); }
  if (threadIdx.x == 0) b2 = __ldg(b2g);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t u_idx = (int64_t)blockIdx.x * kEvalWarps + warp;
  if (u_idx >= U) return;
  float* s = sc[warp];

  // scores: every lane recomputes the user half of fc1, then takes candidates lane, lane+32, ...
  float hu[D];
  {
    float u[D];
    const int64_t uid = users[u_idx];
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      const float4 v = ldg4(repr + uid * ldr + 4 * c);
      u[4 * c] = v.x; u[4 * c + 1] = v.y; u[4 * c + 2] = v.z; u[4 * c + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float a = b1[j];
#pragma unroll
      for (int k = 0; k < D; ++k) a = fmaf(w1[j * 2 * D + k], u[k], a);
      hu[j] = a;
    }
  }
  for (int c = lane; c < C; c += 32) {
    const int64_t iid = cand[u_idx * C + c];
    float it[D];
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      const float4 v = ldg4(repr + iid * ldr + 4 * q);
      it[4 * q] = v.x; it[4 * q + 1] = v.y; it[4 * q + 2] = v.z; it[4 * q + 3] = v.w;
    }
    float score = b2;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      float a = hu[j];
#pragma unroll
      for (int k = 0; k < D; ++k) a = fmaf(w1[j * 2 * D + D + k], it[k], a);
      score = fmaf(w2[j], fmaxf(a, 0.f), score);
    }
    s[c] = score;
    if (scores_out) scores_out[u_idx * C + c] = score;
  }
  __syncwarp();

  // position of every positive in the stable descending order
  int hits_at[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) hits_at[k] = 0;
  int first = C;   // smallest position among the positives
  int wins = 0;    // (pos, neg) pairs with pos > neg   (AUC numerator)
  float loss = 0.f;
  for (int p = 0; p < n_pos; ++p) {
    const float sp = s[p];
    int before = 0, win = 0;
    float l = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float v = s[c];
      before += (v > sp) || (v == sp && c < p);
      if (c >= n_pos) {
        win += sp > v;
        l += neg_log_sigmoid(sp - v);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, o);
      win += __shfl_xor_sync(0xffffffffu, win, o);
      l += __shfl_xor_sync(0xffffffffu, l, o);
    }
    wins += win;
    loss += l;
    first = min(first, before);
#pragma unroll
    for (int k = 0; k < 16; ++k) hits_at[k] += before < (k + 5);
  }
  if (lane == 0) {
    double* o = per_user + u_idx * kEvalCols;
    const double denom = log2((double)first + 2.0);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      o[k] = hits_at[k] > 0 ? 1.0 : 0.0;
      // rec_utils.ndcg: sum(hit[:K]) / log2(argmax(hit[:K]) + 2); no hit -> 0 / log2(2) = 0
      o[16 + k] = hits_at[k] > 0 ? (double)hits_at[k] / denom : 0.0;
    }
    const int n_neg = C - n_pos;
    o[32] = (double)wins / ((double)n_pos * (double)n_neg);
    o[33] = (double)loss;
    o[34] = (double)first;
    o[35] = 0.0;
  }
}

}  // namespace peagnn

using namespace peagnn;

extern "C" int peagnn_eval_rank(const float* repr, int64_t ldr, int32_t D, const int64_t* users,
                                const int64_t* cand, int64_t U, int32_t C, int32_t n_pos,
                                const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b,
                                double* per_user, float* scores_out, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(D == 8 || D == 16 || D == 32, "peagnn_eval_rank: repr_dim %d not in {8,16,32}", D);
  PEAGNN_REQUIRE(repr && users && cand && fc1_w && fc1_b && fc2_w && fc2_b && per_user && ldr % 4 == 0 && aligned16(repr),
                 "peagnn_eval_rank: bad pointers");
  PEAGNN_REQUIRE(C > 0 && C <= kEvalMaxC && n_pos > 0 && n_pos < C, "peagnn_eval_rank: need 0 < n_pos < C <= %d", kEvalMaxC);
  if (U == 0) return PEAGNN_OK;
  const unsigned blocks = (unsigned)((U + kEvalWarps - 1) / kEvalWarps);
  switch (D) {
    case 8: eval_kernel<8><<<blocks, kEvalWarps * 32, 0, stream>>>(repr, ldr, users, cand, U, C, n_pos, fc1_w, fc1_b, fc2_w, fc2_b, per_user, scores_out); break;
    case 16: eval_kernel<16><<<blocks, kEvalWarps * 32, 0, stream>>>(repr, ldr, users, cand, U, C, n_pos, fc1_w, fc1_b, fc2_w, fc2_b, per_user, scores_out); break;
    default: eval_kernel<32><<<blocks, kEvalWarps * 32, 0, stream>>>(repr, ldr, users, cand, U, C, n_pos, fc1_w, fc1_b, fc2_w, fc2_b, per_user, scores_out); break;
  }
  return check_launch("peagnn_eval_rank");
}
