// K4: node-wise projections - `x @ weight` of GCNConv, `lin` of GATConv, `lin_rel`/`lin_root` of
// SAGEConv (PyG-1.5.0; reference call sites models/pea{gcn,gat,sage}.py:16-21) and their
// gradients.  fp32 FFMA on purpose: K and M are <= 128, so every projection moves
// N*(K+M)*4 bytes for 2*N*K*M flop (<= 32 flop/B at K=M=64) - the tile stays HBM-bound on a
// B200 (its fp32 ridge is ~10 flop/B against the measured 6.5 TB/s only if the FFMA pipes are
// kept full), and fp32 keeps the 1e-5 parity bar that TF32 tensor-core inputs would break.
#include <algorithm>
#include "common.cuh"
#include "dense_v2.cuh"
#include "dense_tc.cuh"
#include "dense_umma.cuh"
#include <stdlib.h>

namespace peagnn {

constexpr int kLinThreads = 256;
constexpr int kRPT = 4;  // rows per thread

static inline int pow2_ge(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}
static inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// Y = act(gate(X) @ W + b) (+Y).  Thread (ty, tx) owns rows ty*RPT..+RPT-1 of the tile and
// columns 4*tx..4*tx+3; TX = M/4 column groups, TY = 256/TX row groups, tile = TY*RPT rows.
// Both operands are read from shared memory as float4 along k.
template <int TX>
__global__ void __launch_bounds__(kLinThreads) linear_kernel(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ mask, int64_t ldm,
    int64_t n_rows, int K, int M, const float* __restrict__ W, int w_is_out_in,
    const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ Y, int64_t ldy,
    const float* __restrict__ out_mask, int64_t ldom) {
  constexpr int TY = kLinThreads / TX;
  constexpr int BM = TY * kRPT;
  extern __shared__ __align__(16) float smem[];
  const int ldw = M;         // Ws[k][m]
  const int ldxs = K + 4;    // Xs[r][k], padded (keeps 16-byte alignment, staggers banks)
  float* Ws = smem;
  float* Xs = smem + (size_t)K * ldw;

  for (int idx = threadIdx.x; idx < K * M; idx += kLinThreads) {
    const int k = idx / M, m = idx - k * M;
    Ws[idx] = w_is_out_in ? __ldg(W + (size_t)m * K + k) : __ldg(W + idx);
  }
  const int tx = threadIdx.x % TX;
  const int ty = threadIdx.x / TX;
  const bool col_ok = 4 * tx < M;
  const int k4n = K / 4;
  const int64_t n_tiles = (n_rows + BM - 1) / BM;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * BM;
    __syncthreads();  // previous tile's readers done (and Ws visible on the first pass)
    for (int idx = threadIdx.x; idx < BM * k4n; idx += kLinThreads) {
      const int r = idx / k4n, c = idx - r * k4n;
      const int64_t row = row0 + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n_rows) {
        v = ldg4(X + row * ldx + 4 * c);
        if (mask) {
          const float4 g = ldg4(mask + row * ldm + 4 * c);
          v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f;
          v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
        }
      }
      st4(Xs + (size_t)r * ldxs + 4 * c, v);
    }
    __syncthreads();
    float acc[kRPT][4];
#pragma unroll
    for (int r = 0; r < kRPT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    if (col_ok) {
      const float* xrow = Xs + (size_t)(ty * kRPT) * ldxs;
      const float* wcol = Ws + 4 * tx;
      for (int k4 = 0; k4 < k4n; ++k4) {
        float4 w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(wcol + (size_t)(4 * k4 + j) * ldw);
#pragma unroll
        for (int r = 0; r < kRPT; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(xrow + (size_t)r * ldxs + 4 * k4);
          acc[r][0] = fmaf(a.x, w[0].x, acc[r][0]); acc[r][1] = fmaf(a.x, w[0].y, acc[r][1]);
          acc[r][2] = fmaf(a.x, w[0].z, acc[r][2]); acc[r][3] = fmaf(a.x, w[0].w, acc[r][3]);
          acc[r][0] = fmaf(a.y, w[1].x, acc[r][0]); acc[r][1] = fmaf(a.y, w[1].y, acc[r][1]);
          acc[r][2] = fmaf(a.y, w[1].z, acc[r][2]); acc[r][3] = fmaf(a.y, w[1].w, acc[r][3]);
          acc[r][0] = fmaf(a.z, w[2].x, acc[r][0]); acc[r][1] = fmaf(a.z, w[2].y, acc[r][1]);
          acc[r][2] = fmaf(a.z, w[2].z, acc[r][2]); acc[r][3] = fmaf(a.z, w[2].w, acc[r][3]);
          acc[r][0] = fmaf(a.w, w[3].x, acc[r][0]); acc[r][1] = fmaf(a.w, w[3].y, acc[r][1]);
          acc[r][2] = fmaf(a.w, w[3].z, acc[r][2]); acc[r][3] = fmaf(a.w, w[3].w, acc[r][3]);
        }
      }
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) b = ldg4(bias + 4 * tx);
#pragma unroll
      for (int r = 0; r < kRPT; ++r) {
        const int64_t row = row0 + ty * kRPT + r;
        if (row < n_rows) {
          float4 o = make_float4(acc[r][0] + b.x, acc[r][1] + b.y, acc[r][2] + b.z, acc[r][3] + b.w);
          float* yp = Y + row * ldy + 4 * tx;
          if (accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(yp);
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (out_mask) {
            const float4 g = ldg4(out_mask + row * ldom + 4 * tx);
            o.x = g.x > 0.f ? o.x : 0.f; o.y = g.y > 0.f ? o.y : 0.f; o.z = g.z > 0.f ? o.z : 0.f; o.w = g.w > 0.f ? o.w : 0.f;
          }
          st4(yp, o);
        }
      }
    }
  }
}

template <int TX>
static int launch_linear(const float* X, int64_t ldx, const float* mask, int64_t ldm, int64_t n,
                         int K, int M, const float* W, int w_is_out_in, const float* bias, int relu,
                         int accumulate, float* Y, int64_t ldy, const float* out_mask, int64_t ldom,
                         cudaStream_t stream) {
  constexpr int TY = kLinThreads / TX;
  constexpr int BM = TY * kRPT;
  const size_t smem = ((size_t)K * M + (size_t)BM * (K + 4)) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(linear_kernel<TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  PEAGNN_REQUIRE(smem <= 200 * 1024, "peagnn_linear: tile needs %zu bytes of shared memory", smem);
  const int64_t tiles = (n + BM - 1) / BM;
  const int blocks = (int)imin64(tiles, (int64_t)kNumSMs * 4);
  linear_kernel<TX><<<blocks, kLinThreads, smem, stream>>>(X, ldx, mask, ldm, n, K, M, W, w_is_out_in,
                                                           bias, relu, accumulate, Y, ldy, out_mask, ldom);
  return check_launch("peagnn_linear");
}

// ---- weight / bias gradient -----------------------------------------------------------------
// Stage 1: each CTA takes a slab of rows and accumulates the whole [K, M] product in registers
// (thread tile 4x4, several tiles per thread when K*M/16 > 256, several row groups when it is
// smaller); stage 2 adds the per-CTA partials in CTA order.
constexpr int kWgRows = 32;  // rows staged per pass

template <int TPT /*tiles per thread*/>
__global__ void __launch_bounds__(kLinThreads) wgrad_kernel(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t ldd,
    const float* __restrict__ mask, int64_t ldm, int64_t n_rows, int K, int M, int64_t rows_per_cta,
    float* __restrict__ partial /* [grid][K*M + M] */) {
  extern __shared__ __align__(16) float smem[];
  const int k4n = K / 4, m4n = M / 4;
  const int tiles = k4n * m4n;                 // power of two
  const int RG = tiles >= kLinThreads ? 1 : kLinThreads / tiles;   // row groups
  float* Xs = smem;                            // [kWgRows][K]
  float* Ds = smem + (size_t)kWgRows * K;      // [kWgRows][M]
  float* red = Ds + (size_t)kWgRows * M;       // [RG][K*M + M] for the cross-group fold (RG > 1)

  const int tile0 = threadIdx.x % tiles;       // first tile of this thread (when tiles < 256)
  const int rg = tiles >= kLinThreads ? 0 : threadIdx.x / tiles;
  float acc[TPT][16];
  float accb[TPT][4];
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[t][j] = 0.f;
    accb[t][0] = accb[t][1] = accb[t][2] = accb[t][3] = 0.f;
  }
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = imin64(n_rows, r_begin + rows_per_cta);

  for (int64_t base = r_begin; base < r_end; base += kWgRows) {
    const int nr = (int)imin64(kWgRows, r_end - base);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kWgRows * k4n; idx += kLinThreads) {
      const int r = idx / k4n, c = idx - r * k4n;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nr) v = ldg4(X + (base + r) * ldx + 4 * c);
      st4(Xs + (size_t)r * K + 4 * c, v);
    }
    for (int idx = threadIdx.x; idx < kWgRows * m4n; idx += kLinThreads) {
      const int r = idx / m4n, c = idx - r * m4n;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nr) {
        v = ldg4(dY + (base + r) * ldd + 4 * c);
        if (mask) {
          const float4 g = ldg4(mask + (base + r) * ldm + 4 * c);
          v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f;
          v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
        }
      }
      st4(Ds + (size_t)r * M + 4 * c, v);
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
      const int tile = (TPT == 1) ? tile0 : threadIdx.x + t * kLinThreads;
      const int tk = tile / m4n, tm = tile - tk * m4n;
      for (int r = rg; r < kWgRows; r += RG) {
        const float4 a = *reinterpret_cast<const float4*>(Xs + (size_t)r * K + 4 * tk);
        const float4 d = *reinterpret_cast<const float4*>(Ds + (size_t)r * M + 4 * tm);
        acc[t][0] = fmaf(a.x, d.x, acc[t][0]); acc[t][1] = fmaf(a.x, d.y, acc[t][1]);
        acc[t][2] = fmaf(a.x, d.z, acc[t][2]); acc[t][3] = fmaf(a.x, d.w, acc[t][3]);
        acc[t][4] = fmaf(a.y, d.x, acc[t][4]); acc[t][5] = fmaf(a.y, d.y, acc[t][5]);
        acc[t][6] = fmaf(a.y, d.z, acc[t][6]); acc[t][7] = fmaf(a.y, d.w, acc[t][7]);
        acc[t][8] = fmaf(a.z, d.x, acc[t][8]); acc[t][9] = fmaf(a.z, d.y, acc[t][9]);
        acc[t][10] = fmaf(a.z, d.z, acc[t][10]); acc[t][11] = fmaf(a.z, d.w, acc[t][11]);
        acc[t][12] = fmaf(a.w, d.x, acc[t][12]); acc[t][13] = fmaf(a.w, d.y, acc[t][13]);
        acc[t][14] = fmaf(a.w, d.z, acc[t][14]); acc[t][15] = fmaf(a.w, d.w, acc[t][15]);
        if (tk == 0) { accb[t][0] += d.x; accb[t][1] += d.y; accb[t][2] += d.z; accb[t][3] += d.w; }
      }
    }
  }
  // write out: [K*M] row-major [k][m], then [M] column sums
  const int KM = K * M;
  float* dst = partial + (size_t)blockIdx.x * (KM + M);
  if (RG == 1) {
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
      const int tile = (TPT == 1) ? tile0 : threadIdx.x + t * kLinThreads;
      const int tk = tile / m4n, tm = tile - tk * m4n;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dst[(size_t)(4 * tk + a) * M + 4 * tm + b] = acc[t][4 * a + b];
      if (tk == 0) {
#pragma unroll
        for (int b = 0; b < 4; ++b) dst[KM + 4 * tm + b] = accb[t][b];
      }
    }
  } else {
    __syncthreads();
    {
      const int tk = tile0 / m4n, tm = tile0 - tk * m4n;
      float* rb = red + (size_t)rg * (KM + M);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) rb[(size_t)(4 * tk + a) * M + 4 * tm + b] = acc[0][4 * a + b];
      if (tk == 0) {
#pragma unroll
        for (int b = 0; b < 4; ++b) rb[KM + 4 * tm + b] = accb[0][b];
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < KM + M; idx += kLinThreads) {
      float s = red[idx];
      for (int q = 1; q < RG; ++q) s += red[(size_t)q * (KM + M) + idx];
      dst[idx] = s;
    }
  }
}

// column sums only (K == 0)
__global__ void __launch_bounds__(kLinThreads) colsum_kernel(const float* __restrict__ dY, int64_t ldd,
                                                             const float* __restrict__ mask, int64_t ldm,
                                                             int64_t n_rows, int M, int64_t rows_per_cta,
                                                             float* __restrict__ partial /* [grid][M] */) {
  // thread -> float4 column c4 = t % (M/4), row lane t / (M/4); four rows in flight per thread
  __shared__ __align__(16) float red[kLinThreads * 4];
  const int m4 = M / 4;
  const int RL = kLinThreads / m4;
  const int c4 = threadIdx.x % m4, rl = threadIdx.x / m4;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = imin64(n_rows, r_begin + rows_per_cta);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < RL) {
    for (int64_t r0 = r_begin + rl; r0 < r_end; r0 += 4 * RL) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = r0 + (int64_t)u * RL;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < r_end) {
          v[u] = ldg4(dY + r * ldd + 4 * c4);
          if (mask) {
            const float4 g = ldg4(mask + r * ldm + 4 * c4);
            v[u].x = g.x > 0.f ? v[u].x : 0.f; v[u].y = g.y > 0.f ? v[u].y : 0.f;
            v[u].z = g.z > 0.f ? v[u].z : 0.f; v[u].w = g.w > 0.f ? v[u].w : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
  }
  st4(red + 4 * threadIdx.x, s);
  __syncthreads();
  if ((int)threadIdx.x < M) {
    const int cc4 = threadIdx.x / 4, j = threadIdx.x % 4;
    float t = 0.f;
    for (int q = 0; q < RL; ++q) t += red[4 * (q * m4 + cc4) + j];
    partial[(size_t)blockIdx.x * M + threadIdx.x] = t;
  }
}

__global__ void wgrad_finalize_kernel(const float* __restrict__ partial, int n_parts, int K, int M,
                                      int w_is_out_in, float* __restrict__ dW, float* __restrict__ db) {
  const int KM = K * M;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= KM + M) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s += partial[(size_t)p * (KM + M) + idx];
  if (idx < KM) {
    if (dW) {
      const int k = idx / M, m = idx - k * M;
      dW[w_is_out_in ? (size_t)m * K + k : (size_t)idx] = s;
    }
  } else if (db) {
    db[idx - KM] = s;
  }
}

// PEAGNN_DENSE selects the projection kernels for the hot shapes (A/B timing): "ffma" = fp32 pipe
// (dense_v2.cuh), "mma" = mma.sync 3xTF32 (dense_tc.cuh), "umma" = tcgen05 3xTF32 (dense_umma.cuh);
// default: tcgen05 where the output is 64 wide, mma.sync for the narrower outputs (measured per shape,
// profiles/r1_dense_tensor_cores.md).
static int dense_mode() {
  static const int mode = [] {
    const char* e = getenv("PEAGNN_DENSE");
    if (e && e[0] == 'f') return 0;
    if (e && e[0] == 'm') return 1;
    if (e && e[0] == 'u') return 2;
    if (e && e[0] == 't') return 5;   // "ts": the A-in-TMEM linear for every K = 64 shape (default: 64 -> 64 only)
    return 3;
  }();
  return mode;
}
static bool use_tensor_cores() { return dense_mode() != 0; }

static int wgrad_parts(int64_t n_rows) {
  const int64_t by_rows = (n_rows + 255) / 256;
  return (int)imax64(1, imin64(by_rows, (int64_t)kNumSMs * 2));
}

__global__ void relu_backward_kernel(const float* __restrict__ dy, int64_t ldd, const float* __restrict__ act,
                                     int64_t lda, int64_t n_rows, int f4, float* __restrict__ out, int64_t ldo) {
  const int64_t total = n_rows * f4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int64_t r = idx / f4;
    const int c = (int)(idx - r * f4);
    float4 d = ldg4(dy + r * ldd + 4 * c);
    const float4 a = ldg4(act + r * lda + 4 * c);
    d.x = a.x > 0.f ? d.x : 0.f; d.y = a.y > 0.f ? d.y : 0.f;
    d.z = a.z > 0.f ? d.z : 0.f; d.w = a.w > 0.f ? d.w : 0.f;
    st4(out + r * ldo + 4 * c, d);
  }
}

// ---- GAT logits ------------------------------------------------------------------------------
// one G-lane group per (node, head)
template <int G>
__global__ void __launch_bounds__(256) gat_scores_kernel(const float* __restrict__ H, int64_t ldh, int64_t n_rows,
                                                         int feat, int heads, const float* __restrict__ att_i,
                                                         const float* __restrict__ att_j, float* __restrict__ a_i,
                                                         float* __restrict__ a_j) {
  const int gl = threadIdx.x % G;
  const int64_t lr = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
  if (lr >= n_rows * heads) return;
  const unsigned lane = threadIdx.x & 31u;
  unsigned gmask = 0xffffffffu;
  if constexpr (G < 32) gmask = ((1u << G) - 1u) << ((lane / G) * G);
  const int64_t n = lr / heads;
  const int h = (int)(lr - n * heads);
  const int f4 = feat / 4;
  float si = 0.f, sj = 0.f;
  for (int c = gl; c < f4; c += G) {
    const float4 v = ldg4(H + n * ldh + (int64_t)h * feat + 4 * c);
    const float4 wi = ldg4(att_i + (int64_t)h * feat + 4 * c);
    const float4 wj = ldg4(att_j + (int64_t)h * feat + 4 * c);
    si += v.x * wi.x + v.y * wi.y + v.z * wi.z + v.w * wi.w;
    sj += v.x * wj.x + v.y * wj.y + v.z * wj.z + v.w * wj.w;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    si += __shfl_xor_sync(gmask, si, o, G);
    sj += __shfl_xor_sync(gmask, sj, o, G);
  }
  if (gl == 0) { a_i[lr] = si; a_j[lr] = sj; }
}

// dH (+)= d_ai (x) att_i + d_aj (x) att_j ; per-CTA partial sums of d_att_i / d_att_j.
// thread -> float4 column c = t % C4 (C4 = heads*feat/4, power of two <= 256), row lane t / C4.
__global__ void __launch_bounds__(256) gat_scores_bwd_kernel(
    const float* __restrict__ H, int64_t ldh, int64_t n_rows, int feat, int heads,
    const float* __restrict__ att_i, const float* __restrict__ att_j, const float* __restrict__ d_ai,
    const float* __restrict__ d_aj, float* __restrict__ dH, int64_t ldd, int accumulate,
    int64_t rows_per_cta, float* __restrict__ partial /* [grid][2*heads*feat] */) {
  extern __shared__ __align__(16) float smem[];  // [RL][2*HF]
  const int HF = heads * feat;
  const int C4 = HF / 4;
  const int RL = 256 / C4;
  const int c = threadIdx.x % C4, rl = threadIdx.x / C4;
  const int h = (4 * c) / feat;
  const float4 wi = ldg4(att_i + 4 * c), wj = ldg4(att_j + 4 * c);
  float4 gi = make_float4(0.f, 0.f, 0.f, 0.f), gj = gi;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = imin64(n_rows, r_begin + rows_per_cta);
  for (int64_t r = r_begin + rl; r < r_end; r += RL) {
    const float di = __ldg(d_ai + r * heads + h), dj = __ldg(d_aj + r * heads + h);
    const float4 v = ldg4(H + r * ldh + 4 * c);
    gi.x = fmaf(di, v.x, gi.x); gi.y = fmaf(di, v.y, gi.y); gi.z = fmaf(di, v.z, gi.z); gi.w = fmaf(di, v.w, gi.w);
    gj.x = fmaf(dj, v.x, gj.x); gj.y = fmaf(dj, v.y, gj.y); gj.z = fmaf(dj, v.z, gj.z); gj.w = fmaf(dj, v.w, gj.w);
    float4 o = make_float4(di * wi.x + dj * wj.x, di * wi.y + dj * wj.y, di * wi.z + dj * wj.z, di * wi.w + dj * wj.w);
    float* dp = dH + r * ldd + 4 * c;
    if (accumulate) {
      const float4 p = *reinterpret_cast<const float4*>(dp);
      o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
    }
    st4(dp, o);
  }
  float* mine = smem + (size_t)rl * 2 * HF;
  st4(mine + 4 * c, gi);
  st4(mine + HF + 4 * c, gj);
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * HF; idx += 256) {
    float s = smem[idx];
    for (int q = 1; q < RL; ++q) s += smem[(size_t)q * 2 * HF + idx];
    partial[(size_t)blockIdx.x * 2 * HF + idx] = s;
  }
}

__global__ void sum_parts_kernel(const float* __restrict__ partial, int n_parts, int len, float* __restrict__ out0,
                                 int len0, float* __restrict__ out1) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= len) return;
  float s = 0.f;
  for (int p = 0; p < n_parts; ++p) s += partial[(size_t)p * len + idx];
  if (idx < len0) { if (out0) out0[idx] = s; }
  else if (out1) out1[idx - len0] = s;
}

// ---- double-precision column means (evaluation summary) -------------------------------------
__global__ void __launch_bounds__(256) colmean_stage1(const double* __restrict__ A, int64_t lda, int64_t n_rows,
                                                      int cols, int64_t rows_per_cta, double* __restrict__ partial) {
  __shared__ double red[256];
  const int RL = 256 / cols;
  const int c = threadIdx.x % cols, rl = threadIdx.x / cols;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = imin64(n_rows, r_begin + rows_per_cta);
  double s = 0.0;
  if (rl < RL)
    for (int64_t r = r_begin + rl; r < r_end; r += RL) s += A[r * lda + c];
  red[threadIdx.x] = s;
  __syncthreads();
  if ((int)threadIdx.x < cols) {
    double t = 0.0;
    for (int q = 0; q < RL; ++q) t += red[q * cols + threadIdx.x];
    partial[(size_t)blockIdx.x * cols + threadIdx.x] = t;
  }
}
__global__ void colmean_stage2(const double* __restrict__ partial, int n_parts, int cols, double n_rows_d,
                               double* __restrict__ out) {
  const int c = threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int p = 0; p < n_parts; ++p) s += partial[(size_t)p * cols + c];
  out[c] = s / n_rows_d;
}

}  // namespace peagnn

using namespace peagnn;

extern "C" int peagnn_linear(const float* X, int64_t ldx, const float* mask, int64_t ldm, int64_t n,
                             int32_t K, int32_t M, const float* W, int w_is_out_in, const float* bias,
                             int relu, int accumulate, float* Y, int64_t ldy, const float* out_mask,
                             int64_t ldom, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(K > 0 && M > 0 && K % 4 == 0 && M % 4 == 0 && K <= 128 && M <= 128,
                 "peagnn_linear: K=%d, M=%d must be multiples of 4 in [4, 128]", K, M);
  PEAGNN_REQUIRE(X && W && Y && ldx % 4 == 0 && ldy % 4 == 0 && ldx >= K && ldy >= M && (!mask || (ldm % 4 == 0 && ldm >= K)),
                 "peagnn_linear: bad pointers / leading dimensions");
  PEAGNN_REQUIRE(aligned16(X) && aligned16(Y) && (!mask || aligned16(mask)) && (!bias || aligned16(bias)) &&
                     (!out_mask || (aligned16(out_mask) && ldom % 4 == 0 && ldom >= M)),
                 "peagnn_linear: pointers must be 16-byte aligned");
  if (n == 0) return PEAGNN_OK;
  // hot shapes without an input gate: 3xTF32 tensor-core kernels (dense_tc.cuh); PEAGNN_DENSE=ffma keeps
  // the fp32-pipe kernels for A/B timing
  // 64 -> 64 (26 launches of a PEAGCN step): the warp-specialised TS-form kernel - 50 us vs 55 us (SS form) at
  // 291 k rows; 64 -> 16 stays on mma.sync (29 us vs 35 us), profiles/r2_dense.md
  if (!mask && K == 64 && ((dense_mode() >= 3 && M == 64 && n >= 4096) || (dense_mode() == 5 && (M == 16 || M == 64)))) {
    if (M == 16) return launch_linear_umma_ts<64, 16>(X, ldx, n, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, stream);
    return launch_linear_umma_ts<64, 64>(X, ldx, n, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, stream);
  }
  if (!mask && (dense_mode() == 2 || (dense_mode() >= 3 && M == 64)) && (K == 16 || K == 32 || K == 64) &&
      (M == 16 || M == 32 || M == 64)) {
#define PEAGNN_LINUM(K_, N_) \
  return launch_linear_umma<K_, N_>(X, ldx, n, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, stream)
#define PEAGNN_LINUM_K(K_) \
  do { if (M == 16) PEAGNN_LINUM(K_, 16); if (M == 32) PEAGNN_LINUM(K_, 32); PEAGNN_LINUM(K_, 64); } while (0)
    if (K == 16) PEAGNN_LINUM_K(16);
    if (K == 32) PEAGNN_LINUM_K(32);
    PEAGNN_LINUM_K(64);
#undef PEAGNN_LINUM_K
#undef PEAGNN_LINUM
  }
  if (!mask && use_tensor_cores() && (K == 16 || K == 32 || K == 64) && (M == 16 || M == 32 || M == 64)) {
#define PEAGNN_LINTC(K_, MT_) \
  return launch_linear_tc<K_, MT_>(X, ldx, n, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, stream)
#define PEAGNN_LINTC_K(K_) \
  do { if (M == 16) PEAGNN_LINTC(K_, 2); if (M == 32) PEAGNN_LINTC(K_, 4); PEAGNN_LINTC(K_, 8); } while (0)
    if (K == 16) PEAGNN_LINTC_K(16);
    if (K == 32) PEAGNN_LINTC_K(32);
    PEAGNN_LINTC_K(64);
#undef PEAGNN_LINTC_K
#undef PEAGNN_LINTC
  }
  const int tx = pow2_ge(M / 4);
  if ((K == 64 || K == 16) && tx <= 16) {   // hot shapes: prefetching register-tile kernels
#define PEAGNN_LIN2(K_, TX_, RPT_) \
  return launch_linear_v2<K_, TX_, RPT_>(X, ldx, mask, ldm, n, M, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, stream)
    if (K == 64) {
      if (tx <= 4) PEAGNN_LIN2(64, 4, 2);
      if (tx == 8) PEAGNN_LIN2(64, 8, 4);
      PEAGNN_LIN2(64, 16, 8);
    } else {
      if (tx <= 4) PEAGNN_LIN2(16, 4, 2);
      if (tx == 8) PEAGNN_LIN2(16, 8, 4);
      PEAGNN_LIN2(16, 16, 8);
    }
#undef PEAGNN_LIN2
  }
#define PEAGNN_LIN_CASE(TX_) \
  return launch_linear<TX_>(X, ldx, mask, ldm, n, K, M, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, stream)
  if (tx <= 4) PEAGNN_LIN_CASE(4);
  if (tx == 8) PEAGNN_LIN_CASE(8);
  if (tx == 16) PEAGNN_LIN_CASE(16);
  PEAGNN_LIN_CASE(32);
#undef PEAGNN_LIN_CASE
}

// ---- grouped projections ---------------------------------------------------------------------------------
// CTAs are dealt to the problems in proportion to their 128-row tiles (at least one each); `budget` = the CTAs that are
// resident at once for this kernel, so that a group of small problems is one wave whose CTAs each take a few tiles.
static int fill_group(const peagnn_linear_problem_t* probs, int count, int budget, LinearGroup& g) {
  long long total = 0;
  for (int k = 0; k < count; ++k) total += (probs[k].n + 127) / 128;
  g.count = count;
  int at = 0;
  for (int k = 0; k < count; ++k) {
    g.p[k] = probs[k];
    g.block_start[k] = at;
    const long long tiles = (probs[k].n + 127) / 128;
    long long c = tiles;
    if (total > budget) c = imax64(1, imin64(tiles, tiles * budget / total));
    at += (int)(tiles == 0 ? 0 : c);
  }
  g.block_start[count] = at;
  return at;
}

template <int K, int N>
static int launch_linear_umma_ts_grouped(const peagnn_linear_problem_t* probs, int count, int w_is_out_in, int relu,
                                         int accumulate, cudaStream_t stream) {
  constexpr size_t smem = (size_t)2 * (N * K * 4) + (size_t)4 * 32 * (N + 4) * 4 + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(linear_umma_ts_kernel_grouped<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  LinearGroup g;
  const int blocks = fill_group(probs, count, kNumSMs, g);
  if (blocks == 0) return PEAGNN_OK;
  linear_umma_ts_kernel_grouped<K, N><<<blocks, kPipeThreads, smem, stream>>>(g, w_is_out_in, relu, accumulate);
  return check_launch("peagnn_linear_grouped(umma ts)");
}

template <int K, int N>
static int launch_linear_umma_grouped(const peagnn_linear_problem_t* probs, int count, int w_is_out_in, int relu,
                                      int accumulate, cudaStream_t stream) {
  constexpr size_t a_bytes = (size_t)2 * (128 * K * 4), stage_bytes = (size_t)8 * 32 * (N / 2 + 4) * 4;
  constexpr size_t smem = (a_bytes > stage_bytes ? a_bytes : stage_bytes) + (size_t)2 * (N * K * 4) + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(linear_umma_kernel_grouped<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  LinearGroup g;
  const int blocks = fill_group(probs, count, 2 * kNumSMs, g);
  if (blocks == 0) return PEAGNN_OK;
  linear_umma_kernel_grouped<K, N><<<blocks, kUmThreads, smem, stream>>>(g, w_is_out_in, relu, accumulate);
  return check_launch("peagnn_linear_grouped(umma)");
}

template <int K, int MT>
static int launch_linear_tc_grouped(const peagnn_linear_problem_t* probs, int count, int w_is_out_in, int relu,
                                    int accumulate, cudaStream_t stream) {
  constexpr int LDXS = (K % 32 == 0) ? K + 16 : K;
  constexpr size_t smem = ((size_t)2 * (K / 16) * MT * 32 * 4 + (size_t)128 * LDXS) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(linear_tc_kernel_grouped<K, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  LinearGroup g;
  const int blocks = fill_group(probs, count, 2 * kNumSMs, g);
  if (blocks == 0) return PEAGNN_OK;
  linear_tc_kernel_grouped<K, MT><<<blocks, kTcThreads, smem, stream>>>(g, w_is_out_in, relu, accumulate);
  return check_launch("peagnn_linear_grouped(tc)");
}

extern "C" int peagnn_linear_grouped(const peagnn_linear_problem_t* problems, int32_t count, int32_t K, int32_t M,
                                     int w_is_out_in, int relu, int accumulate, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(count >= 0 && (problems || count == 0), "peagnn_linear_grouped: null problem list");
  PEAGNN_REQUIRE(K > 0 && M > 0 && K % 4 == 0 && M % 4 == 0 && K <= 128 && M <= 128,
                 "peagnn_linear_grouped: K=%d, M=%d must be multiples of 4 in [4, 128]", K, M);
  for (int k = 0; k < count; ++k) {
    const peagnn_linear_problem_t& q = problems[k];
    PEAGNN_REQUIRE(q.n >= 0 && (q.n == 0 || (q.X && q.W && q.Y && q.ldx % 4 == 0 && q.ldy % 4 == 0 && q.ldx >= K && q.ldy >= M)),
                   "peagnn_linear_grouped: problem %d: bad pointers / leading dimensions", k);
    PEAGNN_REQUIRE(q.n == 0 || (aligned16(q.X) && aligned16(q.Y) && (!q.bias || aligned16(q.bias)) &&
                                (!q.out_mask || (aligned16(q.out_mask) && q.ldom % 4 == 0 && q.ldom >= M))),
                   "peagnn_linear_grouped: problem %d: pointers must be 16-byte aligned", k);
  }
  // the shapes of a PEAGNN step under the default kernel choice of peagnn_linear; anything else: one launch per problem
  const bool ts = dense_mode() >= 3 && K == 64 && M == 64;
  const bool um = dense_mode() >= 3 && K == 16 && M == 64;
  const bool tc = dense_mode() >= 3 && K == 64 && M == 16;
  if (!(ts || um || tc)) {
    for (int k = 0; k < count; ++k) {
      const peagnn_linear_problem_t& q = problems[k];
      const int rc = peagnn_linear(q.X, q.ldx, nullptr, 0, q.n, K, M, q.W, w_is_out_in, q.bias, relu, accumulate, q.Y, q.ldy,
                                   q.out_mask, q.ldom, stream_);
      if (rc) return rc;
    }
    return PEAGNN_OK;
  }
  for (int first = 0; first < count; first += PEAGNN_MAX_GROUP) {
    const int c = count - first < PEAGNN_MAX_GROUP ? count - first : PEAGNN_MAX_GROUP;
    int rc;
    if (ts) rc = launch_linear_umma_ts_grouped<64, 64>(problems + first, c, w_is_out_in, relu, accumulate, stream);
    else if (um) rc = launch_linear_umma_grouped<16, 64>(problems + first, c, w_is_out_in, relu, accumulate, stream);
    else rc = launch_linear_tc_grouped<64, 2>(problems + first, c, w_is_out_in, relu, accumulate, stream);
    if (rc) return rc;
  }
  return PEAGNN_OK;
}

extern "C" size_t peagnn_wgrad_workspace_floats(int64_t n, int32_t K, int32_t M) {
  return (size_t)wgrad_parts(n) * ((size_t)K * M + M) + 64;
}

extern "C" int peagnn_linear_wgrad(const float* X, int64_t ldx, const float* dY, int64_t ldd,
                                   const float* mask, int64_t ldm, int64_t n, int32_t K, int32_t M,
                                   int w_is_out_in, float* dW, float* db, float* workspace,
                                   size_t workspace_floats, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE((K == 0 && M >= 4 && M % 4 == 0 && M <= 256) ||
                     (K >= 4 && K <= 128 && is_pow2(K) && M >= 4 && M <= 128 && is_pow2(M)),
                 "peagnn_linear_wgrad: K=%d, M=%d must be powers of two in [4, 128] (or K = 0 with M a multiple of 4, <= 256)", K, M);
  PEAGNN_REQUIRE(dY && workspace && ldd >= M && (K == 0 || (X && ldx >= K)), "peagnn_linear_wgrad: bad pointers");
  PEAGNN_REQUIRE(aligned16(dY) && ldd % 4 == 0 && (K == 0 || (aligned16(X) && ldx % 4 == 0)) && (!mask || (aligned16(mask) && ldm % 4 == 0)),
                 "peagnn_linear_wgrad: alignment");
  if (workspace_floats < peagnn_wgrad_workspace_floats(n, K, M)) {
    set_error("peagnn_linear_wgrad: workspace %zu < %zu floats", workspace_floats, peagnn_wgrad_workspace_floats(n, K, M));
    return PEAGNN_ERR_WORKSPACE;
  }
  const int parts = wgrad_parts(n);
  const int64_t rows_per_cta = ((n + parts - 1) / parts + kWgRows - 1) / kWgRows * kWgRows;
  const int KM = K * M;
  if (n == 0) {
    if (dW && KM) cudaMemsetAsync(dW, 0, sizeof(float) * KM, stream);
    if (db) cudaMemsetAsync(db, 0, sizeof(float) * M, stream);
    return check_launch("peagnn_linear_wgrad(memset)");
  }
  if ((dense_mode() == 2 && ((K == 64 && (M == 64 || M == 32 || M == 16)) || (K == 16 && M == 64))) ||
      (dense_mode() >= 3 && K == 64 && M == 64)) {   // default: tcgen05 for the square shape, mma.sync below
    const int64_t rpc = ((n + parts - 1) / parts + kUmWgRows - 1) / kUmWgRows * kUmWgRows;
    int rc2;
    if (K == 16) rc2 = launch_wgrad_umma<16, 64>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    else if (M == 64) rc2 = launch_wgrad_umma<64, 64>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    else if (M == 32) rc2 = launch_wgrad_umma<64, 32>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    else rc2 = launch_wgrad_umma<64, 16>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    if (rc2) return rc2;
    wgrad_finalize_v2_kernel<<<(KM + M + kFinOutputs - 1) / kFinOutputs, 256, 0, stream>>>(workspace, parts, K, M, w_is_out_in, dW, db);
    return check_launch("peagnn_linear_wgrad(umma stage2)");
  }
  if (use_tensor_cores() && ((K == 64 && (M == 64 || M == 32 || M == 16)) || (K == 16 && M == 64))) {
    const int64_t rpc = ((n + parts - 1) / parts + kTcWgRows - 1) / kTcWgRows * kTcWgRows;
    int rc2;
    if (K == 16) rc2 = launch_wgrad_tc<16, 64>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    else if (M == 64) rc2 = launch_wgrad_tc<64, 64>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    else if (M == 32) rc2 = launch_wgrad_tc<64, 32>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    else rc2 = launch_wgrad_tc<64, 16>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    if (rc2) return rc2;
    wgrad_finalize_v2_kernel<<<(KM + M + kFinOutputs - 1) / kFinOutputs, 256, 0, stream>>>(workspace, parts, K, M, w_is_out_in, dW, db);
    return check_launch("peagnn_linear_wgrad(tc stage2)");
  }
  if (K == 64 && (M == 64 || M == 32 || M == 16)) {   // hot shapes
    const int64_t rpc = ((n + parts - 1) / parts + kWg2Rows - 1) / kWg2Rows * kWg2Rows;
    int rc2;
    if (M == 64) rc2 = launch_wgrad_v2<64, 64>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    else if (M == 32) rc2 = launch_wgrad_v2<64, 32>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    else rc2 = launch_wgrad_v2<64, 16>(X, ldx, dY, ldd, mask, ldm, n, parts, rpc, workspace, stream);
    if (rc2) return rc2;
    wgrad_finalize_v2_kernel<<<(KM + M + kFinOutputs - 1) / kFinOutputs, 256, 0, stream>>>(workspace, parts, K, M, w_is_out_in, dW, db);
    return check_launch("peagnn_linear_wgrad(v2 stage2)");
  }
  if (K == 0) {
    colsum_kernel<<<parts, kLinThreads, 0, stream>>>(dY, ldd, mask, ldm, n, M, rows_per_cta, workspace);
  } else {
    const int tiles = (K / 4) * (M / 4);
    const int RG = tiles >= kLinThreads ? 1 : kLinThreads / tiles;
    const size_t smem = ((size_t)kWgRows * (K + M) + (RG > 1 ? (size_t)RG * (KM + M) : 0)) * sizeof(float);
    PEAGNN_REQUIRE(smem <= 200 * 1024, "peagnn_linear_wgrad: shared memory %zu", smem);
    if (tiles <= kLinThreads) {
      static bool a1 = false;
      if (!a1) { cudaFuncSetAttribute(wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); a1 = true; }
      wgrad_kernel<1><<<parts, kLinThreads, smem, stream>>>(X, ldx, dY, ldd, mask, ldm, n, K, M, rows_per_cta, workspace);
    } else if (tiles == 2 * kLinThreads) {
      wgrad_kernel<2><<<parts, kLinThreads, smem, stream>>>(X, ldx, dY, ldd, mask, ldm, n, K, M, rows_per_cta, workspace);
    } else {
      PEAGNN_REQUIRE(tiles == 4 * kLinThreads, "peagnn_linear_wgrad: unsupported K*M");
      wgrad_kernel<4><<<parts, kLinThreads, smem, stream>>>(X, ldx, dY, ldd, mask, ldm, n, K, M, rows_per_cta, workspace);
    }
  }
  int rc = check_launch("peagnn_linear_wgrad(stage1)");
  if (rc) return rc;
  wgrad_finalize_kernel<<<(KM + M + 255) / 256, 256, 0, stream>>>(workspace, parts, K, M, w_is_out_in, dW, db);
  return check_launch("peagnn_linear_wgrad(stage2)");
}

// ---- grouped weight gradients -----------------------------------------------------------------------------
struct WgradFinal {
  int count;
  int n_parts[PEAGNN_MAX_GROUP];
  const float* partial[PEAGNN_MAX_GROUP];
  float* dW[PEAGNN_MAX_GROUP];
  float* db[PEAGNN_MAX_GROUP];
};
// wgrad_finalize_v2_kernel for a group: blockIdx.y = problem; parts folded in CTA order (deterministic)
__global__ void __launch_bounds__(256) wgrad_finalize_grouped_kernel(const __grid_constant__ WgradFinal f, int K, int M,
                                                                     int w_is_out_in) {
  __shared__ float red[256];
  const int KM = K * M;
  const int k = blockIdx.y;
  const float* __restrict__ partial = f.partial[k];
  const int n_parts = f.n_parts[k];
  const int o = threadIdx.x & (kFinOutputs - 1), s = threadIdx.x / kFinOutputs;
  constexpr int SLICES = 256 / kFinOutputs;
  const int idx = blockIdx.x * kFinOutputs + o;
  float v = 0.f;
  if (idx < KM + M) {
#pragma unroll 4
    for (int p = s; p < n_parts; p += SLICES) v += partial[(size_t)p * (KM + M) + idx];
  }
  red[threadIdx.x] = v;
  __syncthreads();
  if (s == 0 && idx < KM + M) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < SLICES; ++q) t += red[q * kFinOutputs + o];
    if (idx < KM) {
      if (f.dW[k]) {
        const int kk = idx / M, m = idx - kk * M;
        f.dW[k][w_is_out_in ? (size_t)m * K + kk : (size_t)idx] = t;
      }
    } else if (f.db[k]) {
      f.db[k][idx - KM] = t;
    }
  }
}

extern "C" size_t peagnn_wgrad_grouped_workspace_floats(int32_t count, int32_t K, int32_t M) {
  // every problem gets at least one part; the parts of a launch share a budget of 2 CTAs per SM
  const size_t per_launch = (size_t)PEAGNN_MAX_GROUP + 2 * kNumSMs;
  const size_t launches = ((size_t)(count > 0 ? count : 1) + PEAGNN_MAX_GROUP - 1) / PEAGNN_MAX_GROUP;
  const size_t single = peagnn_wgrad_workspace_floats(1 << 30, K, M);     // the per-problem fallback's need
  const size_t grouped = launches * per_launch * ((size_t)K * M + M) + 64;
  return grouped > single ? grouped : single;
}

template <int K, int M, int ROWS_UNIT, class Launch>
static int wgrad_grouped_launch(const peagnn_wgrad_problem_t* probs, int count, int w_is_out_in, float* workspace,
                                cudaStream_t stream, Launch launch, const char* what) {
  const int KM = K * M;
  WgradGroup g;
  WgradFinal f;
  long long total = 0;
  for (int k = 0; k < count; ++k) total += (probs[k].n + 255) / 256;
  const long long budget = 2 * kNumSMs;
  g.count = f.count = count;
  int at = 0;
  for (int k = 0; k < count; ++k) {
    const long long by_rows = (probs[k].n + 255) / 256;
    long long c = by_rows;
    if (total > budget) c = imax64(1, imin64(by_rows, by_rows * budget / total));
    if (by_rows == 0) c = 0;
    g.p[k] = probs[k];
    g.block_start[k] = at;
    g.rows_per_cta[k] = c ? ((probs[k].n + c - 1) / c + ROWS_UNIT - 1) / ROWS_UNIT * ROWS_UNIT : 0;
    g.partial[k] = workspace + (size_t)at * (KM + M);
    f.n_parts[k] = (int)c;
    f.partial[k] = g.partial[k];
    f.dW[k] = probs[k].dW;
    f.db[k] = probs[k].db;
    at += (int)c;
  }
  g.block_start[count] = at;
  if (at > 0) {
    launch(g, at);
    int rc = check_launch(what);
    if (rc) return rc;
  }
  // problems without rows fold zero parts: their outputs are written as zeros
  wgrad_finalize_grouped_kernel<<<dim3((KM + M + kFinOutputs - 1) / kFinOutputs, count), 256, 0, stream>>>(f, K, M, w_is_out_in);
  return check_launch("peagnn_linear_wgrad_grouped(stage2)");
}

extern "C" int peagnn_linear_wgrad_grouped(const peagnn_wgrad_problem_t* problems, int32_t count, int32_t K, int32_t M,
                                           int w_is_out_in, float* workspace, size_t workspace_floats,
                                           peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(count >= 0 && (problems || count == 0), "peagnn_linear_wgrad_grouped: null problem list");
  PEAGNN_REQUIRE(K >= 4 && K <= 128 && is_pow2(K) && M >= 4 && M <= 128 && is_pow2(M),
                 "peagnn_linear_wgrad_grouped: K=%d, M=%d must be powers of two in [4, 128]", K, M);
  PEAGNN_REQUIRE(workspace && workspace_floats >= peagnn_wgrad_grouped_workspace_floats(count, K, M),
                 "peagnn_linear_wgrad_grouped: workspace too small");
  for (int k = 0; k < count; ++k) {
    const peagnn_wgrad_problem_t& q = problems[k];
    PEAGNN_REQUIRE(q.n >= 0 && (q.n == 0 || (q.X && q.dY && q.ldx >= K && q.ldd >= M && q.ldx % 4 == 0 && q.ldd % 4 == 0 &&
                                             aligned16(q.X) && aligned16(q.dY))),
                   "peagnn_linear_wgrad_grouped: problem %d: bad pointers / leading dimensions / alignment", k);
  }
  const bool um = dense_mode() >= 3 && K == 64 && M == 64;
  const bool tc = dense_mode() >= 3 && K == 64 && M == 16;
  if (!(um || tc)) {
    for (int k = 0; k < count; ++k) {
      const peagnn_wgrad_problem_t& q = problems[k];
      if (q.n == 0) {
        if (q.dW) cudaMemsetAsync(q.dW, 0, sizeof(float) * K * M, stream);
        if (q.db) cudaMemsetAsync(q.db, 0, sizeof(float) * M, stream);
        continue;
      }
      const int rc = peagnn_linear_wgrad(q.X, q.ldx, q.dY, q.ldd, nullptr, 0, q.n, K, M, w_is_out_in, q.dW, q.db, workspace,
                                         workspace_floats, stream_);
      if (rc) return rc;
    }
    return PEAGNN_OK;
  }
  const size_t per_launch = ((size_t)PEAGNN_MAX_GROUP + 2 * kNumSMs) * ((size_t)K * M + M);
  int launch_no = 0;
  for (int first = 0; first < count; first += PEAGNN_MAX_GROUP, ++launch_no) {
    const int c = count - first < PEAGNN_MAX_GROUP ? count - first : PEAGNN_MAX_GROUP;
    float* ws = workspace + (size_t)launch_no * per_launch;
    int rc;
    if (um) {
      constexpr size_t smem = (size_t)2 * (kUmWgRows / 4) * 144 * (16 + 64 / 8) + 128;
      static bool attr_set = false;
      if (!attr_set) {
        cudaFuncSetAttribute(wgrad_umma_kernel_grouped<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
      }
      rc = wgrad_grouped_launch<64, 64, kUmWgRows>(problems + first, c, w_is_out_in, ws, stream,
          [&](const WgradGroup& g, int blocks) { wgrad_umma_kernel_grouped<64, 64><<<blocks, kUmThreads, smem, stream>>>(g); },
          "peagnn_linear_wgrad_grouped(umma)");
    } else {
      rc = wgrad_grouped_launch<64, 16, kTcWgRows>(problems + first, c, w_is_out_in, ws, stream,
          [&](const WgradGroup& g, int blocks) { wgrad_tc_kernel_grouped<64, 16><<<blocks, kTcThreads, 0, stream>>>(g); },
          "peagnn_linear_wgrad_grouped(tc)");
    }
    if (rc) return rc;
  }
  return PEAGNN_OK;
}

extern "C" int peagnn_relu_backward(const float* dy, int64_t ldd, const float* act, int64_t lda, int64_t n,
                                    int32_t feat, float* out, int64_t ldo, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(dy && act && out && feat > 0 && feat % 4 == 0 && ldd % 4 == 0 && lda % 4 == 0 && ldo % 4 == 0,
                 "peagnn_relu_backward: bad arguments");
  PEAGNN_REQUIRE(aligned16(dy) && aligned16(act) && aligned16(out), "peagnn_relu_backward: alignment");
  if (n == 0) return PEAGNN_OK;
  const int64_t total = n * (feat / 4);
  const int blocks = (int)imin64((total + 255) / 256, (int64_t)kNumSMs * 8);
  relu_backward_kernel<<<blocks, 256, 0, stream>>>(dy, ldd, act, lda, n, feat / 4, out, ldo);
  return check_launch("peagnn_relu_backward");
}

extern "C" int peagnn_gat_scores(const float* H, int64_t ldh, int64_t n, int32_t feat, int32_t heads,
                                 const float* att_i, const float* att_j, float* a_i, float* a_j,
                                 peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(H && att_i && att_j && a_i && a_j && feat > 0 && feat % 4 == 0 && heads > 0 && ldh % 4 == 0 && ldh >= (int64_t)feat * heads,
                 "peagnn_gat_scores: bad arguments");
  PEAGNN_REQUIRE(aligned16(H) && aligned16(att_i) && aligned16(att_j), "peagnn_gat_scores: alignment");
  if (n == 0) return PEAGNN_OK;
  const int g = std::min(32, pow2_ge(feat / 4));
  const int64_t threads = n * heads * g;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  switch (g) {
    case 1: gat_scores_kernel<1><<<blocks, 256, 0, stream>>>(H, ldh, n, feat, heads, att_i, att_j, a_i, a_j); break;
    case 2: gat_scores_kernel<2><<<blocks, 256, 0, stream>>>(H, ldh, n, feat, heads, att_i, att_j, a_i, a_j); break;
    case 4: gat_scores_kernel<4><<<blocks, 256, 0, stream>>>(H, ldh, n, feat, heads, att_i, att_j, a_i, a_j); break;
    case 8: gat_scores_kernel<8><<<blocks, 256, 0, stream>>>(H, ldh, n, feat, heads, att_i, att_j, a_i, a_j); break;
    case 16: gat_scores_kernel<16><<<blocks, 256, 0, stream>>>(H, ldh, n, feat, heads, att_i, att_j, a_i, a_j); break;
    default: gat_scores_kernel<32><<<blocks, 256, 0, stream>>>(H, ldh, n, feat, heads, att_i, att_j, a_i, a_j); break;
  }
  return check_launch("peagnn_gat_scores");
}

extern "C" int peagnn_gat_scores_backward(const float* H, int64_t ldh, int64_t n, int32_t feat, int32_t heads,
                                          const float* att_i, const float* att_j, const float* d_ai,
                                          const float* d_aj, float* dH, int64_t ldd, int accumulate,
                                          float* d_att_i, float* d_att_j, float* workspace,
                                          size_t workspace_floats, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int HF = heads * feat;
  PEAGNN_REQUIRE(H && att_i && att_j && d_ai && d_aj && dH && workspace && feat % 4 == 0 && HF > 0 && HF <= 1024 && is_pow2(HF / 4),
                 "peagnn_gat_scores_backward: heads*feat/4 must be a power of two <= 256");
  PEAGNN_REQUIRE(aligned16(H) && aligned16(dH) && aligned16(att_i) && aligned16(att_j) && ldh % 4 == 0 && ldd % 4 == 0,
                 "peagnn_gat_scores_backward: alignment");
  const int parts = wgrad_parts(n);
  if (workspace_floats < (size_t)parts * 2 * HF) {
    set_error("peagnn_gat_scores_backward: workspace %zu < %zu floats", workspace_floats, (size_t)parts * 2 * HF);
    return PEAGNN_ERR_WORKSPACE;
  }
  const int64_t rows_per_cta = (n + parts - 1) / parts;
  const int RL = 256 / (HF / 4);
  const size_t smem = (size_t)RL * 2 * HF * sizeof(float);
  gat_scores_bwd_kernel<<<parts, 256, smem, stream>>>(H, ldh, n, feat, heads, att_i, att_j, d_ai, d_aj, dH, ldd,
                                                      accumulate, rows_per_cta, workspace);
  int rc = check_launch("peagnn_gat_scores_backward(stage1)");
  if (rc) return rc;
  sum_parts_kernel<<<(2 * HF + 255) / 256, 256, 0, stream>>>(workspace, parts, 2 * HF, d_att_i, HF, d_att_j);
  return check_launch("peagnn_gat_scores_backward(stage2)");
}

extern "C" int peagnn_column_mean(const double* A, int64_t lda, int64_t n, int32_t cols, double* out,
                                  double* workspace, size_t workspace_doubles, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(A && out && workspace && cols > 0 && cols <= 64 && n > 0, "peagnn_column_mean: bad arguments");
  const int parts = (int)imax64(1, imin64((n + 255) / 256, (int64_t)kNumSMs));
  if (workspace_doubles < (size_t)parts * cols) {
    set_error("peagnn_column_mean: workspace %zu < %zu doubles", workspace_doubles, (size_t)parts * cols);
    return PEAGNN_ERR_WORKSPACE;
  }
  const int64_t rows_per_cta = (n + parts - 1) / parts;
  colmean_stage1<<<parts, 256, 0, stream>>>(A, lda, n, cols, rows_per_cta, workspace);
  int rc = check_launch("peagnn_column_mean(stage1)");
  if (rc) return rc;
  colmean_stage2<<<1, 64, 0, stream>>>(workspace, parts, cols, (double)n, out);
  return check_launch("peagnn_column_mean(stage2)");
}
