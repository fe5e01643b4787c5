// Tensor-core projection kernels for the hot shapes of the PEAGNN channels (K, M in {16, 32, 64}).
//
// fp32 in, fp32 out, fp32-level accuracy: every operand is split x = hi + lo with hi = tf32(x),
// lo = tf32(x - hi), and a product is issued as three m16n8k8 TF32 MMAs (lo*hi, hi*lo, hi*hi,
// small terms first) into one fp32 accumulator ("3xTF32").  The dropped lo*lo term and the tf32
// rounding of lo are both ~2^-22 relative, so results agree with an fp32 FFMA GEMM to ~1e-6;
// the parity tests keep their 1e-5 bound.  The FFMA kernels in dense_v2.cuh were bound by the fp32
// pipe at ~4x the HBM time of these shapes; with the MMAs the kernels are back on the HBM roofline.
//
//   linear_tc : Y = act(X @ W + b (+ Y)), optional relu-backward gate on the way out.
//               CTA tile 128 rows (8 warps x 16 rows x all M columns); W is split once per CTA and
//               kept in shared memory in fragment order (one LDS.128 per MMA triple); the X tile is
//               prefetched into registers one tile ahead and staged through shared memory.
//               The contraction index is permuted inside every 16-wide block (thread t owns columns
//               4t..4t+3) so an A fragment pair is one LDS.128 per row; W's fragments use the same map.
//   wgrad_tc  : dW = X^T @ gate(dY), db = colsum(gate(dY)).  The reduction runs over rows, 64 per
//               staged tile; each warp owns a 16 x (M / WN) block of dW for all rows of the CTA's
//               slab.  MMA accumulators are flushed into fp32 registers after every tile so the
//               tensor core's truncating adder never sees a long chain.  Deterministic (no atomics).
#pragma once
#include "common.cuh"

namespace peagnn {

constexpr int kTcThreads = 256;

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// c += a (16x8, row) * b (8x8, col); no volatile: the scheduler may interleave independent tiles
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// -------------------------------------------------------------------------------------------------
template <int K, int MT /* n-tiles of 8 columns: M = 8 * MT */>
__device__ __forceinline__ void linear_tc_kernel_body(
    const float* __restrict__ X, int64_t ldx, int64_t n_rows, const float* __restrict__ W, int w_is_out_in,
    const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ Y, int64_t ldy,
    const float* __restrict__ out_mask, int64_t ldom, const int block_id, const int n_blocks) {
  constexpr int M = 8 * MT;
  constexpr int BM = 128;
  constexpr int K4 = K / 4, KB = K / 16;
  constexpr int LDXS = (K % 32 == 0) ? K + 16 : K;      // row stride = 16 banks (mod 32): LDS.128 conflict-free
  constexpr int NPRE = BM * K4 / kTcThreads;            // float4 per thread per tile
  constexpr int NG = MT < 4 ? MT : 4;                   // n-tiles issued together
  static_assert(K % 16 == 0 && BM * K4 % kTcThreads == 0 && MT % NG == 0, "unsupported shape");
  extern __shared__ __align__(16) float smem[];
  uint4* Wf = reinterpret_cast<uint4*>(smem);           // [2 * KB][MT][32] : {hi b0, hi b1, lo b0, lo b1}
  float* Xs = smem + 2 * KB * MT * 32 * 4;              // [BM][LDXS]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;

  for (int idx = threadIdx.x; idx < 2 * KB * MT * 32; idx += kTcThreads) {
    const int l = idx & 31, nt = (idx >> 5) % MT, ks = (idx >> 5) / MT;
    const int k0 = 16 * (ks >> 1) + 4 * (l & 3) + 2 * (ks & 1);   // contraction slots t / t+4 of this k-step
    const int m = 8 * nt + (l >> 2);
    const float w0 = w_is_out_in ? __ldg(W + (size_t)m * K + k0) : __ldg(W + (size_t)k0 * M + m);
    const float w1 = w_is_out_in ? __ldg(W + (size_t)m * K + k0 + 1) : __ldg(W + (size_t)(k0 + 1) * M + m);
    uint4 f;
    split_tf32(w0, f.x, f.z);
    split_tf32(w1, f.y, f.w);
    Wf[idx] = f;
  }

  const int64_t n_tiles = (n_rows + BM - 1) / BM;
  float4 pre[NPRE];
  auto fetch = [&](int64_t tile) {
    const int64_t row0 = tile * BM;
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int idx = threadIdx.x + j * kTcThreads;
      const int r = idx / K4, c = idx - r * K4;
      const int64_t row = row0 + r;
      pre[j] = row < n_rows ? ldg4(X + row * ldx + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  int64_t tile = block_id;
  if (tile < n_tiles) fetch(tile);
  for (; tile < n_tiles; tile += n_blocks) {
    __syncthreads();          // previous tile's readers are done (Wf visible on the first pass)
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int idx = threadIdx.x + j * kTcThreads;
      const int r = idx / K4, c = idx - r * K4;
      st4(Xs + r * LDXS + 4 * c, pre[j]);
    }
    __syncthreads();
    const int64_t next = tile + n_blocks;
    if (next < n_tiles) fetch(next);   // in flight while this tile is computed

    float acc[MT][4];
#pragma unroll
    for (int nt = 0; nt < MT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    const float* xr0 = Xs + (warp * 16 + g) * LDXS + 4 * t;
    const float* xr1 = xr0 + 8 * LDXS;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      const float4 xa = *reinterpret_cast<const float4*>(xr0 + 16 * kb);
      const float4 xb = *reinterpret_cast<const float4*>(xr1 + 16 * kb);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        uint32_t ah[4], al[4];
        split_tf32(s ? xa.z : xa.x, ah[0], al[0]);
        split_tf32(s ? xb.z : xb.x, ah[1], al[1]);
        split_tf32(s ? xa.w : xa.y, ah[2], al[2]);
        split_tf32(s ? xb.w : xb.y, ah[3], al[3]);
        const uint4* wf = Wf + (size_t)(2 * kb + s) * MT * 32 + lane;
#pragma unroll
        for (int n0 = 0; n0 < MT; n0 += NG) {
          uint4 w[NG];
#pragma unroll
          for (int j = 0; j < NG; ++j) w[j] = wf[(n0 + j) * 32];
#pragma unroll
          for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], al, w[j].x, w[j].y);
#pragma unroll
          for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], ah, w[j].z, w[j].w);
#pragma unroll
          for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], ah, w[j].x, w[j].y);
        }
      }
    }

    // epilogue: thread holds (row g, cols 2t, 2t+1) and (row g + 8, same cols) of every n-tile
    const int64_t r0 = tile * BM + warp * 16 + g;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t row = r0 + 8 * h;
      if (row < n_rows) {
        float* yrow = Y + row * ldy + 2 * t;
        const float* grow = out_mask ? out_mask + row * ldom + 2 * t : nullptr;
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
          float2 o = make_float2(acc[nt][2 * h], acc[nt][2 * h + 1]);
          if (bias) {
            const float2 b = __ldg(reinterpret_cast<const float2*>(bias + 8 * nt + 2 * t));
            o.x += b.x; o.y += b.y;
          }
          float2* yp = reinterpret_cast<float2*>(yrow + 8 * nt);
          if (accumulate) {
            const float2 p = *yp;
            o.x += p.x; o.y += p.y;
          }
          if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
          if (grow) {   // relu backward fused on the way out
            const float2 gt = __ldg(reinterpret_cast<const float2*>(grow + 8 * nt));
            o.x = gt.x > 0.f ? o.x : 0.f; o.y = gt.y > 0.f ? o.y : 0.f;
          }
          *yp = o;
        }
      }
    }
  }
}

template <int K, int MT /* n-tiles of 8 columns: M = 8 * MT */>
__global__ void __launch_bounds__(kTcThreads, 2) linear_tc_kernel(
    const float* __restrict__ X, int64_t ldx, int64_t n_rows, const float* __restrict__ W, int w_is_out_in,
    const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ Y, int64_t ldy,
    const float* __restrict__ out_mask, int64_t ldom) {
  linear_tc_kernel_body<K, MT>(X, ldx, n_rows, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, (int)blockIdx.x, (int)gridDim.x);
}
// grouped form: CTA b works on the problem whose block range holds b, as block (b - first) of (last - first)
template <int K, int MT /* n-tiles of 8 columns: M = 8 * MT */>
__global__ void __launch_bounds__(kTcThreads, 2) linear_tc_kernel_grouped(const __grid_constant__ LinearGroup grp, int w_is_out_in, int relu,
                                                 int accumulate) {
  const int k = group_of_block(grp, (int)blockIdx.x);
  const peagnn_linear_problem_t& q = grp.p[k];
  linear_tc_kernel_body<K, MT>(q.X, q.ldx, q.n, q.W, w_is_out_in, q.bias, relu, accumulate, q.Y, q.ldy, q.out_mask, q.ldom,
      (int)blockIdx.x - grp.block_start[k], grp.block_start[k + 1] - grp.block_start[k]);
}

template <int K, int MT>
static int launch_linear_tc(const float* X, int64_t ldx, int64_t n, const float* W, int w_is_out_in,
                            const float* bias, int relu, int accumulate, float* Y, int64_t ldy,
                            const float* out_mask, int64_t ldom, cudaStream_t stream) {
  constexpr int LDXS = (K % 32 == 0) ? K + 16 : K;
  constexpr size_t smem = ((size_t)2 * (K / 16) * MT * 32 * 4 + (size_t)128 * LDXS) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(linear_tc_kernel<K, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const int64_t tiles = (n + 127) / 128;
  const int blocks = (int)imin64(tiles, (int64_t)kNumSMs * 2);
  linear_tc_kernel<K, MT><<<blocks, kTcThreads, smem, stream>>>(X, ldx, n, W, w_is_out_in, bias, relu, accumulate,
                                                                Y, ldy, out_mask, ldom);
  return check_launch("peagnn_linear(tc)");
}

// -------------------------------------------------------------------------------------------------
constexpr int kTcWgRows = 64;

template <int K, int M, bool HAS_MASK>
__device__ __forceinline__ void wgrad_tc_kernel_body(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t ldd,
    const float* __restrict__ mask, int64_t ldm, int64_t n_rows, int64_t rows_per_cta,
    float* __restrict__ partial /* [parts][K*M + M] */, const int block_id) {
  constexpr int WM = K / 16;                 // warps along dW's rows (one 16-row MMA tile each)
  constexpr int WN = 8 / WM;                 // warps along dW's columns
  constexpr int NT = (M / 8) / WN;           // 8-column tiles per warp
  constexpr int K4 = K / 4, M4 = M / 4;
  constexpr int LDX = K + 8, LDD = M + 8;    // row stride = 8 banks (mod 32): the 4 x 8 fragment reads are conflict-free
  constexpr int NX = kTcWgRows * K4 / kTcThreads;
  constexpr int ND = kTcWgRows * M4 / kTcThreads;
  static_assert(K % 16 == 0 && 8 % WM == 0 && (M / 8) % WN == 0 && NT >= 1, "unsupported shape");
  static_assert(kTcWgRows * K4 % kTcThreads == 0 && kTcWgRows * M4 % kTcThreads == 0, "staging must split evenly");
  constexpr int KM = K * M;
  __shared__ __align__(16) float Xs[kTcWgRows * LDX];
  __shared__ __align__(16) float Ds[kTcWgRows * LDD];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int mt = warp % WM, nt0 = (warp / WM) * NT;

  float acc[NT][4], master[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = master[j][i] = 0.f;
  float4 bsum[ND];
#pragma unroll
  for (int j = 0; j < ND; ++j) bsum[j] = make_float4(0.f, 0.f, 0.f, 0.f);

  const int64_t r_begin = (int64_t)block_id * rows_per_cta;
  const int64_t r_end = imin64(n_rows, r_begin + rows_per_cta);

  float4 px[NX], pd[ND];
  auto fetch = [&](int64_t base) {
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      const int idx = threadIdx.x + j * kTcThreads;
      const int r = idx / K4, c = idx - r * K4;
      px[j] = (base + r < r_end) ? ldg4(X + (base + r) * ldx + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int idx = threadIdx.x + j * kTcThreads;
      const int r = idx / M4, c = idx - r * M4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (base + r < r_end) {
        v = ldg4(dY + (base + r) * ldd + 4 * c);
        if (HAS_MASK) {
          const float4 gt = ldg4(mask + (base + r) * ldm + 4 * c);
          v.x = gt.x > 0.f ? v.x : 0.f; v.y = gt.y > 0.f ? v.y : 0.f;
          v.z = gt.z > 0.f ? v.z : 0.f; v.w = gt.w > 0.f ? v.w : 0.f;
        }
      }
      pd[j] = v;
    }
  };

  if (r_begin < r_end) fetch(r_begin);
  for (int64_t base = r_begin; base < r_end; base += kTcWgRows) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NX; ++j) {
      const int idx = threadIdx.x + j * kTcThreads;
      const int r = idx / K4, c = idx - r * K4;
      st4(Xs + r * LDX + 4 * c, px[j]);
    }
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int idx = threadIdx.x + j * kTcThreads;
      const int r = idx / M4, c = idx - r * M4;
      st4(Ds + r * LDD + 4 * c, pd[j]);
      bsum[j].x += pd[j].x; bsum[j].y += pd[j].y; bsum[j].z += pd[j].z; bsum[j].w += pd[j].w;   // column c is fixed per (thread, j)
    }
    __syncthreads();
    if (base + kTcWgRows < r_end) fetch(base + kTcWgRows);

#pragma unroll
    for (int ks = 0; ks < kTcWgRows / 8; ++ks) {
      // A = X^T: element (m = dW row, k = data row);  B = dY: element (k = data row, n = dW column)
      const float* xa = Xs + (8 * ks + t) * LDX + 16 * mt + g;
      uint32_t ah[4], al[4];
      split_tf32(xa[0], ah[0], al[0]);
      split_tf32(xa[8], ah[1], al[1]);
      split_tf32(xa[4 * LDX], ah[2], al[2]);
      split_tf32(xa[4 * LDX + 8], ah[3], al[3]);
      const float* db = Ds + (8 * ks + t) * LDD + 8 * nt0 + g;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(db[8 * j], bh0, bl0);
        split_tf32(db[8 * j + 4 * LDD], bh1, bl1);
        mma_tf32(acc[j], al, bh0, bh1);
        mma_tf32(acc[j], ah, bl0, bl1);
        mma_tf32(acc[j], ah, bh0, bh1);
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) { master[j][i] += acc[j][i]; acc[j][i] = 0.f; }
  }

  float* dst = partial + (size_t)block_id * (KM + M);
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    float* p = dst + (size_t)(16 * mt + g) * M + 8 * (nt0 + j) + 2 * t;
    *reinterpret_cast<float2*>(p) = make_float2(master[j][0], master[j][1]);
    *reinterpret_cast<float2*>(p + 8 * M) = make_float2(master[j][2], master[j][3]);
  }
  // bias gradient: per-thread column sums folded over the 64 staging rows in a fixed order
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ND; ++j) {
    const int idx = threadIdx.x + j * kTcThreads;
    const int r = idx / M4, c = idx - r * M4;
    st4(Ds + r * LDD + 4 * c, bsum[j]);
  }
  __syncthreads();
  if (threadIdx.x < M) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < kTcWgRows; ++r) s += Ds[r * LDD + threadIdx.x];
    dst[KM + threadIdx.x] = s;
  }
}

template <int K, int M, bool HAS_MASK>
__global__ void __launch_bounds__(kTcThreads, 2) wgrad_tc_kernel(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t ldd,
    const float* __restrict__ mask, int64_t ldm, int64_t n_rows, int64_t rows_per_cta,
    float* __restrict__ partial /* [grid][K*M + M] */) {
  wgrad_tc_kernel_body<K, M, HAS_MASK>(X, ldx, dY, ldd, mask, ldm, n_rows, rows_per_cta, partial, (int)blockIdx.x);
}
// grouped form (peagnn_linear_wgrad_grouped): CTA b is part (b - first) of the problem whose block range holds b
template <int K, int M>
__global__ void __launch_bounds__(kTcThreads, 2) wgrad_tc_kernel_grouped(const __grid_constant__ WgradGroup grp) {
  const int k = group_of_block(grp, (int)blockIdx.x);
  const peagnn_wgrad_problem_t& q = grp.p[k];
  wgrad_tc_kernel_body<K, M, false>(q.X, q.ldx, q.dY, q.ldd, nullptr, 0, q.n, grp.rows_per_cta[k], grp.partial[k],
                       (int)blockIdx.x - grp.block_start[k]);
}

template <int K, int M>
static int launch_wgrad_tc(const float* X, int64_t ldx, const float* dY, int64_t ldd, const float* mask,
                           int64_t ldm, int64_t n, int parts, int64_t rows_per_cta, float* workspace,
                           cudaStream_t stream) {
  if (mask)
    wgrad_tc_kernel<K, M, true><<<parts, kTcThreads, 0, stream>>>(X, ldx, dY, ldd, mask, ldm, n, rows_per_cta, workspace);
  else
    wgrad_tc_kernel<K, M, false><<<parts, kTcThreads, 0, stream>>>(X, ldx, dY, ldd, mask, ldm, n, rows_per_cta, workspace);
  return check_launch("peagnn_linear_wgrad(tc)");
}

}  // namespace peagnn
