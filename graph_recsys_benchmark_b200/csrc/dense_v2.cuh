// Second-generation projection kernels for the hot shapes of the PEAGNN channels
// (K, M in {16, 32, 64}): fp32 FFMA register tiles fed from shared memory, with the NEXT tile's
// global loads issued into registers before the current tile is computed (software prefetch), so
// HBM latency overlaps the FFMA work even at 2 CTAs / SM.
//   linear_v2 : Y = act(gate(X) @ W + b (+ Y));  thread tile RPT rows x 4 cols, rows interleaved
//               (ty + TY*r) so the two row groups of a warp hit different banks.
//   wgrad_v2  : dW = X^T @ gate(dY), db = colsum(gate(dY));  thread tile 4 (k) x 8 (m), row groups
//               folded through shared memory in a fixed order (deterministic).
#pragma once
#include "common.cuh"

namespace peagnn {

constexpr int kV2Threads = 256;

template <int K, int TX, int RPT, bool HAS_MASK>
__global__ void __launch_bounds__(kV2Threads) linear_v2_kernel(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ mask, int64_t ldm,
    int64_t n_rows, int M, const float* __restrict__ W, int w_is_out_in,
    const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ Y, int64_t ldy,
    const float* __restrict__ out_mask, int64_t ldom) {
  constexpr int TY = kV2Threads / TX;
  constexpr int BM = TY * RPT;
  constexpr int K4 = K / 4;
  constexpr int LDXS = K + 4;
  constexpr int LDW = 4 * TX;
  constexpr int NPRE = BM * K4 / kV2Threads;   // float4 per thread per tile
  static_assert(BM * K4 % kV2Threads == 0, "tile must split evenly over the CTA");
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                  // [K][LDW]
  float* Xs = smem + K * LDW;        // [BM][LDXS]

  for (int idx = threadIdx.x; idx < K * LDW; idx += kV2Threads) {
    const int k = idx / LDW, m = idx - k * LDW;
    float w = 0.f;
    if (m < M) w = w_is_out_in ? __ldg(W + (size_t)m * K + k) : __ldg(W + (size_t)k * M + m);
    Ws[idx] = w;
  }
  const int tx = threadIdx.x % TX;
  const int ty = threadIdx.x / TX;
  const bool col_ok = 4 * tx < M;
  const int64_t n_tiles = (n_rows + BM - 1) / BM;

  float4 pre[NPRE];
  float4 prem[HAS_MASK ? NPRE : 1];
  auto fetch = [&](int64_t tile) {
    const int64_t row0 = tile * BM;
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int idx = threadIdx.x + j * kV2Threads;
      const int r = idx / K4, c = idx - r * K4;
      const int64_t row = row0 + r;
      pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (HAS_MASK) prem[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < n_rows) {
        pre[j] = ldg4(X + row * ldx + 4 * c);
        if (HAS_MASK) prem[j] = ldg4(mask + row * ldm + 4 * c);
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int idx = threadIdx.x + j * kV2Threads;
      const int r = idx / K4, c = idx - r * K4;
      float4 v = pre[j];
      if (HAS_MASK) {
        const float4 g = prem[j];
        v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f;
        v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
      }
      st4(Xs + r * LDXS + 4 * c, v);
    }
  };

  int64_t tile = blockIdx.x;
  if (tile < n_tiles) fetch(tile);
  for (; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();          // previous tile's readers are done (Ws visible on the first pass)
    stash();
    __syncthreads();
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) fetch(next);   // in flight while this tile is computed

    float acc[RPT][4];
#pragma unroll
    for (int r = 0; r < RPT; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
    const float* xbase = Xs + ty * LDXS;
    const float* wbase = Ws + 4 * tx;
#pragma unroll 4
    for (int k4 = 0; k4 < K4; ++k4) {
      float4 w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const float4*>(wbase + (4 * k4 + j) * LDW);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(xbase + (r * TY) * LDXS + 4 * k4);
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
    if (col_ok) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) b = ldg4(bias + 4 * tx);
      const int64_t row0 = tile * BM;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int64_t row = row0 + ty + r * TY;
        if (row < n_rows) {
          float4 o = make_float4(acc[r][0] + b.x, acc[r][1] + b.y, acc[r][2] + b.z, acc[r][3] + b.w);
          float* yp = Y + row * ldy + 4 * tx;
          if (accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(yp);
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (out_mask) {   // relu backward fused on the way out: keep the gradient where the activation was > 0
            const float4 g = ldg4(out_mask + row * ldom + 4 * tx);
            o.x = g.x > 0.f ? o.x : 0.f; o.y = g.y > 0.f ? o.y : 0.f; o.z = g.z > 0.f ? o.z : 0.f; o.w = g.w > 0.f ? o.w : 0.f;
          }
          st4(yp, o);
        }
      }
    }
  }
}

template <int K, int TX, int RPT>
static int launch_linear_v2(const float* X, int64_t ldx, const float* mask, int64_t ldm, int64_t n, int M,
                            const float* W, int w_is_out_in, const float* bias, int relu, int accumulate,
                            float* Y, int64_t ldy, const float* out_mask, int64_t ldom, cudaStream_t stream) {
  constexpr int TY = kV2Threads / TX;
  constexpr int BM = TY * RPT;
  const size_t smem = ((size_t)K * 4 * TX + (size_t)BM * (K + 4)) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(linear_v2_kernel<K, TX, RPT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(linear_v2_kernel<K, TX, RPT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_set = true;
  }
  const int64_t tiles = (n + BM - 1) / BM;
  const int blocks = (int)imin64(tiles, (int64_t)kNumSMs * 2);
  if (mask)
    linear_v2_kernel<K, TX, RPT, true><<<blocks, kV2Threads, smem, stream>>>(X, ldx, mask, ldm, n, M, W, w_is_out_in,
                                                                            bias, relu, accumulate, Y, ldy, out_mask, ldom);
  else
    linear_v2_kernel<K, TX, RPT, false><<<blocks, kV2Threads, smem, stream>>>(X, ldx, mask, ldm, n, M, W, w_is_out_in,
                                                                             bias, relu, accumulate, Y, ldy, out_mask, ldom);
  return check_launch("peagnn_linear(v2)");
}

// -------------------------------------------------------------------------------------------------
// dW[k][m] partials.  Thread tile: k in [4tk, 4tk+4), m in [4tm, 4tm+4) U [M/2 + 4tm, M/2 + 4tm + 4).
constexpr int kWg2Rows = 64;

template <int K, int M, bool HAS_MASK>
__global__ void __launch_bounds__(kV2Threads) wgrad_v2_kernel(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t ldd,
    const float* __restrict__ mask, int64_t ldm, int64_t n_rows, int64_t rows_per_cta,
    float* __restrict__ partial /* [grid][K*M + M] */) {
  constexpr int TK = K / 4, TM = M / 8;
  constexpr int TILES = TK * TM;
  constexpr int RG = kV2Threads / TILES;                 // row groups
  constexpr int K4 = K / 4, M4 = M / 4;
  constexpr int NX = kWg2Rows * K4 / kV2Threads;          // float4 of X per thread per pass
  constexpr int ND = kWg2Rows * M4 / kV2Threads;          // float4 of dY per thread per pass
  static_assert(TILES <= kV2Threads && kV2Threads % TILES == 0, "tile count must divide the CTA");
  static_assert(kWg2Rows * K4 % kV2Threads == 0 && kWg2Rows * M4 % kV2Threads == 0, "staging must split evenly");
  constexpr int KM = K * M;
  constexpr int STAGE = kWg2Rows * (K + M);
  constexpr int FOLD = RG * (KM + M);
  extern __shared__ __align__(16) float smem[];           // max(STAGE, FOLD) floats
  float* Xs = smem;                                        // [rows][K]
  float* Ds = smem + kWg2Rows * K;                         // [rows][M]

  const int tile = threadIdx.x % TILES;
  const int rg = threadIdx.x / TILES;
  const int tk = tile / TM, tm = tile - tk * TM;
  float acc[4][8];
  float accb[8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll
  for (int b = 0; b < 8; ++b) accb[b] = 0.f;

  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = imin64(n_rows, r_begin + rows_per_cta);

  float4 px[NX], pd[ND], pm[HAS_MASK ? ND : 1];
  auto fetch = [&](int64_t base) {
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      const int idx = threadIdx.x + j * kV2Threads;
      const int r = idx / K4, c = idx - r * K4;
      px[j] = (base + r < r_end) ? ldg4(X + (base + r) * ldx + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int idx = threadIdx.x + j * kV2Threads;
      const int r = idx / M4, c = idx - r * M4;
      const bool ok = base + r < r_end;
      pd[j] = ok ? ldg4(dY + (base + r) * ldd + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (HAS_MASK) pm[j] = ok ? ldg4(mask + (base + r) * ldm + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      const int idx = threadIdx.x + j * kV2Threads;
      st4(Xs + 4 * idx, px[j]);                            // [r][K] is contiguous in idx
    }
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int idx = threadIdx.x + j * kV2Threads;
      float4 v = pd[j];
      if (HAS_MASK) {
        const float4 g = pm[j];
        v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f;
        v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
      }
      st4(Ds + 4 * idx, v);
    }
  };

  if (r_begin < r_end) fetch(r_begin);
  for (int64_t base = r_begin; base < r_end; base += kWg2Rows) {
    __syncthreads();
    stash();
    __syncthreads();
    if (base + kWg2Rows < r_end) fetch(base + kWg2Rows);
#pragma unroll 4
    for (int r = rg; r < kWg2Rows; r += RG) {
      const float4 a = *reinterpret_cast<const float4*>(Xs + r * K + 4 * tk);
      const float4 d0 = *reinterpret_cast<const float4*>(Ds + r * M + 4 * tm);
      const float4 d1 = *reinterpret_cast<const float4*>(Ds + r * M + M / 2 + 4 * tm);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], dv[j], acc[i][j]);
      if (tk == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) accb[j] += dv[j];
      }
    }
  }
  // fold the row groups in order, then write this CTA's partial
  __syncthreads();
  float* mine = smem + (size_t)rg * (KM + M);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mine[(4 * tk + i) * M + 4 * tm + j] = acc[i][j];
      mine[(4 * tk + i) * M + M / 2 + 4 * tm + j] = acc[i][4 + j];
    }
  }
  if (tk == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mine[KM + 4 * tm + j] = accb[j];
      mine[KM + M / 2 + 4 * tm + j] = accb[4 + j];
    }
  }
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * (KM + M);
  for (int idx = threadIdx.x; idx < KM + M; idx += kV2Threads) {
    float s = smem[idx];
#pragma unroll
    for (int q = 1; q < RG; ++q) s += smem[(size_t)q * (KM + M) + idx];
    dst[idx] = s;
  }
  (void)STAGE; (void)FOLD;
}

// out[idx] = sum over parts: a CTA owns 16 outputs, 16 threads per output walk interleaved slices of the
// parts (short dependent chains - this kernel is pure latency), slices folded in a fixed order.
constexpr int kFinOutputs = 16;

__global__ void __launch_bounds__(256) wgrad_finalize_v2_kernel(const float* __restrict__ partial, int n_parts,
                                                                int K, int M, int w_is_out_in,
                                                                float* __restrict__ dW, float* __restrict__ db) {
  __shared__ float red[256];
  const int KM = K * M;
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
      if (dW) {
        const int k = idx / M, m = idx - k * M;
        dW[w_is_out_in ? (size_t)m * K + k : (size_t)idx] = t;
      }
    } else if (db) {
      db[idx - KM] = t;
    }
  }
}

template <int K, int M>
static int launch_wgrad_v2(const float* X, int64_t ldx, const float* dY, int64_t ldd, const float* mask,
                           int64_t ldm, int64_t n, int parts, int64_t rows_per_cta, float* workspace,
                           cudaStream_t stream) {
  constexpr int TILES = (K / 4) * (M / 8);
  constexpr int RG = kV2Threads / TILES;
  constexpr size_t stage = (size_t)kWg2Rows * (K + M);
  constexpr size_t fold = (size_t)RG * (K * M + M);
  const size_t smem = (stage > fold ? stage : fold) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wgrad_v2_kernel<K, M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(wgrad_v2_kernel<K, M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_set = true;
  }
  if (mask)
    wgrad_v2_kernel<K, M, true><<<parts, kV2Threads, smem, stream>>>(X, ldx, dY, ldd, mask, ldm, n, rows_per_cta, workspace);
  else
    wgrad_v2_kernel<K, M, false><<<parts, kV2Threads, smem, stream>>>(X, ldx, dY, ldd, mask, ldm, n, rows_per_cta, workspace);
  return check_launch("peagnn_linear_wgrad(v2)");
}

}  // namespace peagnn
