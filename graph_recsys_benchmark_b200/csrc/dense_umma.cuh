// tcgen05 projection kernel for the hot shapes of the PEAGNN channels:  Y = act(X @ W + b (+ Y)).
//
// fp32 in / fp32 out with fp32-level accuracy on the 5th-generation tensor cores: every operand is
// split x = hi + lo (hi = tf32(x), lo = tf32(x - hi)) and the product is issued as three chains of
// tcgen05.mma.kind::tf32 (lo*hi, hi*lo, hi*hi - small terms first) into one fp32 accumulator in
// tensor memory ("3xTF32", error ~2^-22 per product; parity tests keep their 1e-5 bound).
//
// One CTA owns a 128-row tile at a time (UMMA M = 128, N = all output columns, K = 8 per instruction):
//   * W is split once per CTA into B_hi / B_lo, kept in shared memory for the CTA's lifetime;
//   * the X tile is prefetched into registers one tile ahead, split, and written as A_hi / A_lo in the
//     canonical no-swizzle K-major core-matrix layout (8 rows x 16 bytes per core matrix; consecutive
//     8-row groups 128 B apart, consecutive 16-byte K chunks (BM/8)*128 B apart);
//   * one thread issues the 3 * K/8 MMAs and commits them to an mbarrier; all 8 warps then read the
//     accumulator back (tcgen05.ld 32x32b: warp w owns TMEM lanes 32*(w%4).., half of the columns),
//     transpose it through shared memory and apply bias / accumulate / relu / relu-backward gate on
//     row-contiguous 128-bit accesses (a thread-per-row store, even 256 bits wide, measured slower).
// Two CTAs are resident per SM, so one CTA's split + epilogue overlaps the other's loads and MMAs; there
// is no warp specialisation inside a CTA.  Layout / descriptor bit fields follow the PTX ISA's tcgen05
// shared-memory and instruction descriptors.
#pragma once
#include "common.cuh"
#include "dense_tc.cuh"   // split_tf32

namespace peagnn {

constexpr int kUmThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: no swizzle, K-major; addresses / offsets in 16-byte units
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;      // K direction: next 16-byte chunk of the same rows
  d |= (uint64_t)(sbo_bytes >> 4) << 32;      // M/N direction: next group of 8 rows
  d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
  return d;                                   // base offset 0, layout type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();   // a lost MMA completion must fail the launch, not hang the device
}

// x = hi + lo with both halves on the tf32 grid; round-to-nearest (ties away, as cvt.rna) done on the
// bit pattern: add half an ulp of the 13 dropped bits, clear them.  x - hi is exact in fp32.
__device__ __forceinline__ void split_tf32_fast(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

template <int K, int N>
__device__ __forceinline__ void linear_umma_kernel_body(
    const float* __restrict__ X, int64_t ldx, int64_t n_rows, const float* __restrict__ W, int w_is_out_in,
    const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ Y, int64_t ldy,
    const float* __restrict__ out_mask, int64_t ldom, const int block_id, const int n_blocks) {
  constexpr int BM = 128;
  constexpr int KC = K / 4;                     // 16-byte chunks along K
  constexpr int KS = K / 8;                     // MMA k-steps
  constexpr int CG = KC / 4;                    // chunk groups of 4 (one warp load covers 8 rows x 4 chunks)
  constexpr int ITS = 2 * CG;                   // (8 rows x 4 chunks) blocks per warp per tile
  constexpr uint32_t A_SBO = 128, A_LBO = (BM / 8) * 128;
  constexpr uint32_t B_SBO = 128, B_LBO = (N / 8) * 128;
  constexpr int A_BYTES = BM * K * 4, B_BYTES = N * K * 4;
  constexpr uint32_t TMEM_COLS = N < 32 ? 32 : N;
  constexpr int CW = N / 2;                     // output columns per warp in the epilogue
  constexpr int SP = CW + 4;                    // staging row pitch (floats): 16-byte stores stay conflict-free
  constexpr int STAGE_BYTES = 8 * 32 * SP * 4;  // epilogue staging, overlaid on the A buffers
  constexpr int A_REGION = 2 * A_BYTES > STAGE_BYTES ? 2 * A_BYTES : STAGE_BYTES;
  static_assert(K % 16 == 0 && N % 16 == 0 && N <= 256 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "unsupported shape");
  // instruction descriptor: D = f32, A = B = tf32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  extern __shared__ __align__(128) uint8_t umma_smem[];
  uint8_t* sAh = umma_smem;
  uint8_t* sAl = sAh + A_BYTES;
  uint8_t* sBh = sAh + A_REGION;
  uint8_t* sBl = sBh + B_BYTES;
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ uint32_t tmem_base_slot;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rsub = lane & 7, csub = lane >> 3;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_done)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // B = W^T as an N x K, K-major operand: element (n, k) = W[k][n]
  for (int idx = threadIdx.x; idx < N * KC; idx += kUmThreads) {
    const int j = idx / N, n = idx - j * N;
    float w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      w[e] = w_is_out_in ? __ldg(W + (size_t)n * K + 4 * j + e) : __ldg(W + (size_t)(4 * j + e) * N + n);
    uint4 hi, lo;
    split_tf32(w[0], hi.x, lo.x); split_tf32(w[1], hi.y, lo.y);
    split_tf32(w[2], hi.z, lo.z); split_tf32(w[3], hi.w, lo.w);
    const uint32_t off = (uint32_t)(j * (N / 8) + (n >> 3)) * 128u + (uint32_t)(n & 7) * 16u;
    *reinterpret_cast<uint4*>(sBh + off) = hi;
    *reinterpret_cast<uint4*>(sBl + off) = lo;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t bar = smem_u32(&mma_done);
  const uint32_t aH = smem_u32(sAh), aL = smem_u32(sAl), bH = smem_u32(sBh), bL = smem_u32(sBl);

  const int64_t n_tiles = (n_rows + BM - 1) / BM;
  float4 pre[ITS];
  auto fetch = [&](int64_t tile) {
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
      const int b = warp * ITS + it;
      const int64_t row = tile * BM + 8 * (b / CG) + rsub;
      const int chunk = 4 * (b % CG) + csub;
      pre[it] = row < n_rows ? ldg4(X + row * ldx + 4 * chunk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  uint32_t phase = 0;
  int64_t tile = block_id;
  if (tile < n_tiles) fetch(tile);
  for (; tile < n_tiles; tile += n_blocks) {
    // split this tile into A_hi / A_lo (the previous tile's MMAs were waited for by every thread)
#pragma unroll
    for (int it = 0; it < ITS; ++it) {
      const int b = warp * ITS + it;
      const int i = b / CG, chunk = 4 * (b % CG) + csub;
      uint4 hi, lo;
      split_tf32_fast(pre[it].x, hi.x, lo.x); split_tf32_fast(pre[it].y, hi.y, lo.y);
      split_tf32_fast(pre[it].z, hi.z, lo.z); split_tf32_fast(pre[it].w, hi.w, lo.w);
      const uint32_t off = (uint32_t)(chunk * (BM / 8) + i) * 128u + (uint32_t)rsub * 16u;
      *reinterpret_cast<uint4*>(sAh + off) = hi;
      *reinterpret_cast<uint4*>(sAl + off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // orders the previous tile's tcgen05.ld
    __syncthreads();
    const int64_t next = tile + n_blocks;
    if (next < n_tiles) fetch(next);   // in flight while the MMAs and the epilogue run

    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t a0 = term == 0 ? aL : aH;
        const uint32_t b0 = term == 1 ? bL : bH;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          umma_tf32(tmem_base, umma_desc(a0 + ks * 2 * A_LBO, A_LBO, A_SBO), umma_desc(b0 + ks * 2 * B_LBO, B_LBO, B_SBO),
                    IDESC, (term | ks) != 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    if (warp == 0) mbar_wait(bar, phase);   // one warp polls; the others sleep at the barrier
    phase ^= 1;
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w reads TMEM lanes 32 * (w % 4).. (one accumulator row per thread), columns
    // [(w / 4) * CW, +CW), transposes its 32 x CW block through shared memory (the A buffers are free
    // once the MMAs have completed) and finishes row-contiguous: full 32-byte sectors on every
    // global access (a thread-per-row store would touch half sectors and make L2 fetch before write).
    const int col0 = (warp >> 2) * CW;
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0;
    float* stage = reinterpret_cast<float*>(sAh) + warp * (32 * SP);
    __syncwarp();   // lanes leave the mbarrier spin at different times; tcgen05.ld is warp-collective
#pragma unroll
    for (int u = 0; u < CW / 8; ++u) {
      float v[8];
      tmem_ld8(taddr + 8 * u, v);
      st4(stage + lane * SP + 8 * u, make_float4(v[0], v[1], v[2], v[3]));
      st4(stage + lane * SP + 8 * u + 4, make_float4(v[4], v[5], v[6], v[7]));
    }
    __syncwarp();
    constexpr int C4 = CW / 4;                  // float4 per row segment
    constexpr int RPI = 32 / C4;                // rows per warp instruction
    const int c = col0 + 4 * (lane % C4);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = ldg4(bias + c);
#pragma unroll
    for (int it = 0; it < C4; ++it) {
      const int r = it * RPI + lane / C4;
      const int64_t row = tile * BM + 32 * (warp & 3) + r;
      if (row < n_rows) {
        float4 o = *reinterpret_cast<const float4*>(stage + r * SP + 4 * (lane % C4));
        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
        float* yp = Y + row * ldy + c;
        if (accumulate) {
          const float4 p = *reinterpret_cast<const float4*>(yp);
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (out_mask) {   // relu backward fused on the way out
          const float4 g = ldg4(out_mask + row * ldom + c);
          o.x = g.x > 0.f ? o.x : 0.f; o.y = g.y > 0.f ? o.y : 0.f; o.z = g.z > 0.f ? o.z : 0.f; o.w = g.w > 0.f ? o.w : 0.f;
        }
        st4(yp, o);
      }
    }
    __syncthreads();   // staging lives in the A buffers: nobody may split the next tile into them yet
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

template <int K, int N>
__global__ void __launch_bounds__(kUmThreads, 2) linear_umma_kernel(
    const float* __restrict__ X, int64_t ldx, int64_t n_rows, const float* __restrict__ W, int w_is_out_in,
    const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ Y, int64_t ldy,
    const float* __restrict__ out_mask, int64_t ldom) {
  linear_umma_kernel_body<K, N>(X, ldx, n_rows, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, (int)blockIdx.x, (int)gridDim.x);
}
// grouped form: CTA b works on the problem whose block range holds b, as block (b - first) of (last - first)
template <int K, int N>
__global__ void __launch_bounds__(kUmThreads, 2) linear_umma_kernel_grouped(const __grid_constant__ LinearGroup grp, int w_is_out_in, int relu,
                                                 int accumulate) {
  const int k = group_of_block(grp, (int)blockIdx.x);
  const peagnn_linear_problem_t& q = grp.p[k];
  linear_umma_kernel_body<K, N>(q.X, q.ldx, q.n, q.W, w_is_out_in, q.bias, relu, accumulate, q.Y, q.ldy, q.out_mask, q.ldom,
      (int)blockIdx.x - grp.block_start[k], grp.block_start[k + 1] - grp.block_start[k]);
}

template <int K, int N>
static int launch_linear_umma(const float* X, int64_t ldx, int64_t n, const float* W, int w_is_out_in,
                              const float* bias, int relu, int accumulate, float* Y, int64_t ldy,
                              const float* out_mask, int64_t ldom, cudaStream_t stream) {
  constexpr size_t a_bytes = (size_t)2 * (128 * K * 4), stage_bytes = (size_t)8 * 32 * (N / 2 + 4) * 4;
  constexpr size_t smem = (a_bytes > stage_bytes ? a_bytes : stage_bytes) + (size_t)2 * (N * K * 4) + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(linear_umma_kernel<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const int64_t tiles = (n + 127) / 128;
  const int blocks = (int)imin64(tiles, (int64_t)kNumSMs * 2);
  linear_umma_kernel<K, N><<<blocks, kUmThreads, smem, stream>>>(X, ldx, n, W, w_is_out_in, bias, relu, accumulate,
                                                                 Y, ldy, out_mask, ldom);
  return check_launch("peagnn_linear(umma)");
}

// -------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05:  dW[K, M] = X^T @ gate(dY),  db = colsum(gate(dY)).
//
// The reduction runs over data rows, 64 per staged tile (8 MMA k-steps of 8 rows).  Operands are the
// TRANSPOSED tiles, both K-major in the MMA's sense (contiguous along the data-row index):
//   A = X^T  : UMMA M = 128 rows = the K features of X (rows K..127 are zero padding, written once),
//   B = dY^T : UMMA N = M rows.
// A thread loads a 4-row x 4-column block (four float4, rows fully coalesced), and the transposition is
// free: chunk e of the block is (r0[e], r1[e], r2[e], r3[e]) - one 16-byte store per feature.  Groups of
// 8 features are 144 B apart (not 128) so the eight stores of a quarter warp hit eight different bank quads.
// The accumulator is read back after EVERY tile and added into fp32 registers (the tensor core's adder
// never sees a long chain; same policy as wgrad_tc); per-CTA partials go to the workspace, stage 2 is
// wgrad_finalize_v2_kernel.  Deterministic: no atomics anywhere.
constexpr int kUmWgRows = 64;

template <int K, int M, bool HAS_MASK>
__device__ __forceinline__ void wgrad_umma_kernel_body(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t ldd,
    const float* __restrict__ mask, int64_t ldm, int64_t n_rows, int64_t rows_per_cta,
    float* __restrict__ partial /* [parts][K*M + M] */, const int block_id) {
  constexpr int KQ = K / 4, MQ = M / 4;                  // column quads of X / dY
  constexpr int KCH = kUmWgRows / 4;                     // 16-byte chunks along the reduction (4 data rows each)
  constexpr uint32_t SBO = 144;                          // 8-feature group pitch
  constexpr uint32_t A_LBO = 16 * SBO, B_LBO = (M / 8) * SBO;
  constexpr int A_BYTES = KCH * A_LBO, B_BYTES = KCH * B_LBO;
  constexpr uint32_t TMEM_COLS = M < 32 ? 32 : M;
  constexpr int CW = M / 2;                              // accumulator columns per warp
  constexpr int KM = K * M;
  static_assert(K % 16 == 0 && K <= 64 && M % 16 == 0 && M <= 64, "unsupported shape");
  static_assert(KQ * KCH <= kUmThreads && MQ * KCH <= kUmThreads, "one 4x4 block per thread");
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(M >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

  extern __shared__ __align__(128) uint8_t umma_smem[];
  uint8_t* sAh = umma_smem;
  uint8_t* sAl = sAh + A_BYTES;
  uint8_t* sBh = sAl + A_BYTES;
  uint8_t* sBl = sBh + B_BYTES;
  __shared__ __align__(8) uint64_t mma_done;
  __shared__ uint32_t tmem_base_slot;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_done)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * A_BYTES / 16; i += kUmThreads)    // zero padding rows (and everything else) once
    reinterpret_cast<uint4*>(sAh)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t bar = smem_u32(&mma_done);
  const uint32_t aH = smem_u32(sAh), aL = smem_u32(sAl), bH = smem_u32(sBh), bL = smem_u32(sBl);

  // this thread's 4x4 blocks: data rows 4*kc .. 4*kc+3 of the tile, columns 4*q .. 4*q+3
  const bool has_x = threadIdx.x < KQ * KCH, has_d = threadIdx.x < MQ * KCH;
  const int xq = threadIdx.x % KQ, xkc = threadIdx.x / KQ;
  const int dq = threadIdx.x % MQ, dkc = threadIdx.x / MQ;

  const int64_t r_begin = (int64_t)block_id * rows_per_cta;
  const int64_t r_end = imin64(n_rows, r_begin + rows_per_cta);

  float4 px[4], pd[4];
  auto fetch = [&](int64_t base) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      px[j] = pd[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int64_t rx = base + 4 * xkc + j, rd = base + 4 * dkc + j;
      if (has_x && rx < r_end) px[j] = ldg4(X + rx * ldx + 4 * xq);
      if (has_d && rd < r_end) {
        float4 v = ldg4(dY + rd * ldd + 4 * dq);
        if (HAS_MASK) {
          const float4 gt = ldg4(mask + rd * ldm + 4 * dq);
          v.x = gt.x > 0.f ? v.x : 0.f; v.y = gt.y > 0.f ? v.y : 0.f;
          v.z = gt.z > 0.f ? v.z : 0.f; v.w = gt.w > 0.f ? v.w : 0.f;
        }
        pd[j] = v;
      }
    }
  };
  // chunk e of a 4x4 block = column e of its four rows; feature f lives in group f / 8, slot f % 8
  auto store_block = [&](uint8_t* hi_base, uint8_t* lo_base, const float4 (&r)[4], int q, int kc, uint32_t lbo) {
    const float c[4][4] = {{r[0].x, r[1].x, r[2].x, r[3].x}, {r[0].y, r[1].y, r[2].y, r[3].y},
                           {r[0].z, r[1].z, r[2].z, r[3].z}, {r[0].w, r[1].w, r[2].w, r[3].w}};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = 4 * q + e;
      uint4 hi, lo;
      split_tf32_fast(c[e][0], hi.x, lo.x); split_tf32_fast(c[e][1], hi.y, lo.y);
      split_tf32_fast(c[e][2], hi.z, lo.z); split_tf32_fast(c[e][3], hi.w, lo.w);
      const uint32_t off = (uint32_t)kc * lbo + (uint32_t)(f >> 3) * SBO + (uint32_t)(f & 7) * 16u;
      *reinterpret_cast<uint4*>(hi_base + off) = hi;
      *reinterpret_cast<uint4*>(lo_base + off) = lo;
    }
  };

  float master[CW];
#pragma unroll
  for (int i = 0; i < CW; ++i) master[i] = 0.f;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int col0 = (warp >> 2) * CW;
  const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)col0;
  const bool reads_acc = 32 * (warp & 3) < K;            // TMEM lanes K..127 hold the zero padding

  uint32_t phase = 0;
  if (r_begin < r_end) fetch(r_begin);
  for (int64_t base = r_begin; base < r_end; base += kUmWgRows) {
    if (has_x) store_block(sAh, sAl, px, xq, xkc, A_LBO);
    if (has_d) {
      store_block(sBh, sBl, pd, dq, dkc, B_LBO);
#pragma unroll
      for (int j = 0; j < 4; ++j) { bsum.x += pd[j].x; bsum.y += pd[j].y; bsum.z += pd[j].z; bsum.w += pd[j].w; }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (base + kUmWgRows < r_end) fetch(base + kUmWgRows);

    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t a0 = term == 0 ? aL : aH;
        const uint32_t b0 = term == 1 ? bL : bH;
#pragma unroll
        for (int ks = 0; ks < kUmWgRows / 8; ++ks)
          umma_tf32(tmem_base, umma_desc(a0 + ks * 2 * A_LBO, A_LBO, SBO), umma_desc(b0 + ks * 2 * B_LBO, B_LBO, SBO),
                    IDESC, (term | ks) != 0);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    if (warp == 0) mbar_wait(bar, phase);
    phase ^= 1;
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (reads_acc) {
#pragma unroll
      for (int u = 0; u < CW / 8; ++u) {
        float v[8];
        tmem_ld8(taddr + 8 * u, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) master[8 * u + e] += v[e];
      }
    }
  }

  // this CTA's partial: thread owns feature 32 * (warp % 4) + lane, columns [col0, col0 + CW)
  float* dst = partial + (size_t)block_id * (KM + M);
  const int f = 32 * (warp & 3) + lane;
  if (reads_acc && f < K) {
#pragma unroll
    for (int u = 0; u < CW / 4; ++u)
      st4(dst + (size_t)f * M + col0 + 4 * u, make_float4(master[4 * u], master[4 * u + 1], master[4 * u + 2], master[4 * u + 3]));
  }
  // bias gradient: per-thread column sums folded over the 16 row groups in a fixed order
  __syncthreads();
  float* red = reinterpret_cast<float*>(sAh);            // [KCH][M]
  if (has_d) st4(red + dkc * M + 4 * dq, bsum);
  __syncthreads();
  if (threadIdx.x < M) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < KCH; ++r) s += red[r * M + threadIdx.x];
    dst[KM + threadIdx.x] = s;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

template <int K, int M, bool HAS_MASK>
__global__ void __launch_bounds__(kUmThreads, 2) wgrad_umma_kernel(
    const float* __restrict__ X, int64_t ldx, const float* __restrict__ dY, int64_t ldd,
    const float* __restrict__ mask, int64_t ldm, int64_t n_rows, int64_t rows_per_cta,
    float* __restrict__ partial /* [grid][K*M + M] */) {
  wgrad_umma_kernel_body<K, M, HAS_MASK>(X, ldx, dY, ldd, mask, ldm, n_rows, rows_per_cta, partial, (int)blockIdx.x);
}
// grouped form (peagnn_linear_wgrad_grouped): CTA b is part (b - first) of the problem whose block range holds b
template <int K, int M>
__global__ void __launch_bounds__(kUmThreads, 2) wgrad_umma_kernel_grouped(const __grid_constant__ WgradGroup grp) {
  const int k = group_of_block(grp, (int)blockIdx.x);
  const peagnn_wgrad_problem_t& q = grp.p[k];
  wgrad_umma_kernel_body<K, M, false>(q.X, q.ldx, q.dY, q.ldd, nullptr, 0, q.n, grp.rows_per_cta[k], grp.partial[k],
                       (int)blockIdx.x - grp.block_start[k]);
}

template <int K, int M>
static int launch_wgrad_umma(const float* X, int64_t ldx, const float* dY, int64_t ldd, const float* mask,
                             int64_t ldm, int64_t n, int parts, int64_t rows_per_cta, float* workspace,
                             cudaStream_t stream) {
  constexpr size_t smem = (size_t)2 * (kUmWgRows / 4) * 144 * (16 + M / 8) + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wgrad_umma_kernel<K, M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(wgrad_umma_kernel<K, M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  if (mask)
    wgrad_umma_kernel<K, M, true><<<parts, kUmThreads, smem, stream>>>(X, ldx, dY, ldd, mask, ldm, n, rows_per_cta, workspace);
  else
    wgrad_umma_kernel<K, M, false><<<parts, kUmThreads, smem, stream>>>(X, ldx, dY, ldd, mask, ldm, n, rows_per_cta, workspace);
  return check_launch("peagnn_linear_wgrad(umma)");
}

// =================================================================================================
// Warp-specialised TS-form linear kernel (the default for the 64 -> 64 projections; PEAGNN_DENSE=umma selects the
// SS-form kernel above instead).  One CTA per SM, 13 warps:
//   warps 0-7   producers : global -> registers (two tiles ahead) -> hi / lo split -> A operand stage
//   warps 8-11  epilogue  : TMEM accumulator (2 buffers) -> shared staging -> row-contiguous stores
//   warp  12    MMA       : one thread issues the 3 * K/8 tcgen05.mma of a tile and commits them to the
//                           stage's "empty" barrier and the accumulator's "full" barrier
// so the loads, the split, the MMAs and the epilogue of consecutive tiles overlap inside the CTA.
// (An SS-form variant of this pipeline - A_hi / A_lo staged in shared memory - measured 59 us for the 64 -> 64
// projection of 291 k rows against 55 us for the two-CTA kernel above and 50 us for the TS form below, and was
// removed: profiles/r2_dense.md.)
constexpr int kPipeProducerWarps = 8;
constexpr int kPipeThreads = 32 * (kPipeProducerWarps + 4 + 1);

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// The TS form of the GEMM.  The A operand never touches shared memory: every
// producer thread owns one row of the tile (= one TMEM lane), loads it 32 bytes at a time, splits it and
// writes A_hi / A_lo straight into tensor memory with tcgen05.st; B (the weight matrix) stays in shared
// memory (round 1's hypothesis that the SS-form TF32 MMAs are held back by the shared-memory operand path: the TS
// form is 9 % faster at 64 -> 64, parity-green on the 180 linear cases of tests/test_gpu_kernels.py).  TMEM columns: [0, 2N) two accumulators, then per stage [A_hi (K) | A_lo (K)].
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int K, int N>
__device__ __forceinline__ void linear_umma_ts_kernel_body(
    const float* __restrict__ X, int64_t ldx, int64_t n_rows, const float* __restrict__ W, int w_is_out_in,
    const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ Y, int64_t ldy,
    const float* __restrict__ out_mask, int64_t ldom, const int block_id, const int n_blocks) {
  constexpr int BM = 128;
  constexpr int KC = K / 4, KS = K / 8;
  constexpr int KPW = KS / 2;                              // k-steps per producer warp (two warps share a lane quarter)
  constexpr uint32_t B_SBO = 128, B_LBO = (N / 8) * 128;
  constexpr int B_BYTES = N * K * 4;
  constexpr uint32_t A_COL0 = 2 * N;                       // first TMEM column of the A stages
  constexpr uint32_t TMEM_NEED = 2 * N + 4 * K;
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  constexpr int SP = N + 4;
  static_assert(K % 16 == 0 && N % 16 == 0 && N <= 64 && TMEM_NEED <= 512, "unsupported shape");
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  extern __shared__ __align__(128) uint8_t umma_smem[];
  uint8_t* sBh = umma_smem;
  uint8_t* sBl = sBh + B_BYTES;
  float* sOut = reinterpret_cast<float*>(sBl + B_BYTES);   // [4 warps][32][SP]
  __shared__ __align__(8) uint64_t bars[8];                // full[2] empty[2] acc_full[2] acc_empty[2]
  __shared__ uint32_t tmem_base_slot;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars[0 + s]), 32 * kPipeProducerWarps);
      mbar_init(smem_u32(&bars[2 + s]), 1);
      mbar_init(smem_u32(&bars[4 + s]), 1);
      mbar_init(smem_u32(&bars[6 + s]), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = threadIdx.x; idx < N * KC; idx += kPipeThreads) {
    const int j = idx / N, n = idx - j * N;
    float w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      w[e] = w_is_out_in ? __ldg(W + (size_t)n * K + 4 * j + e) : __ldg(W + (size_t)(4 * j + e) * N + n);
    uint4 hi, lo;
    split_tf32(w[0], hi.x, lo.x); split_tf32(w[1], hi.y, lo.y);
    split_tf32(w[2], hi.z, lo.z); split_tf32(w[3], hi.w, lo.w);
    const uint32_t off = (uint32_t)(j * (N / 8) + (n >> 3)) * 128u + (uint32_t)(n & 7) * 16u;
    *reinterpret_cast<uint4*>(sBh + off) = hi;
    *reinterpret_cast<uint4*>(sBl + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]);
  const uint32_t afull0 = smem_u32(&bars[4]), aempty0 = smem_u32(&bars[6]);
  const int64_t n_tiles = (n_rows + BM - 1) / BM;
  const int64_t my_tiles = block_id < n_tiles ? (n_tiles - block_id + n_blocks - 1) / n_blocks : 0;

  if (warp < kPipeProducerWarps) {
    // ---------------------------------------------------------------- producers: one row (= TMEM lane) per thread
    const int q = warp & 3, h = warp >> 2;
    float4 pre0[2 * KPW], pre1[2 * KPW];
    auto fetch = [&](int64_t i, float4 (&pre)[2 * KPW]) {
      const int64_t row = (block_id + i * n_blocks) * BM + 32 * q + lane;
      const bool ok = i < my_tiles && row < n_rows;
#pragma unroll
      for (int j = 0; j < 2 * KPW; ++j)
        pre[j] = ok ? ldg4(X + row * ldx + 8 * (h * KPW) + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto produce = [&](int64_t i, float4 (&pre)[2 * KPW]) {
      const int s = (int)(i & 1);
      mbar_wait(empty0 + 8 * s, (uint32_t)(((i >> 1) & 1) ^ 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = tmem_base + ((uint32_t)(32 * q) << 16) + A_COL0 + (uint32_t)(s * 2 * K) + 8u * (uint32_t)(h * KPW);
#pragma unroll
      for (int j = 0; j < KPW; ++j) {
        uint32_t hi[8], lo[8];
        split_tf32_fast(pre[2 * j].x, hi[0], lo[0]); split_tf32_fast(pre[2 * j].y, hi[1], lo[1]);
        split_tf32_fast(pre[2 * j].z, hi[2], lo[2]); split_tf32_fast(pre[2 * j].w, hi[3], lo[3]);
        split_tf32_fast(pre[2 * j + 1].x, hi[4], lo[4]); split_tf32_fast(pre[2 * j + 1].y, hi[5], lo[5]);
        split_tf32_fast(pre[2 * j + 1].z, hi[6], lo[6]); split_tf32_fast(pre[2 * j + 1].w, hi[7], lo[7]);
        tmem_st8(a_hi + 8 * j, hi);
        tmem_st8(a_hi + K + 8 * j, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(full0 + 8 * s);
      fetch(i + 2, pre);
    };
    fetch(0, pre0);
    fetch(1, pre1);
    for (int64_t i = 0; i < my_tiles; i += 2) {
      produce(i, pre0);
      if (i + 1 < my_tiles) produce(i + 1, pre1);
    }
  } else if (warp == kPipeProducerWarps + 4) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t bH = smem_u32(sBh), bL = smem_u32(sBl);
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int s = (int)(i & 1);
        const uint32_t par = (uint32_t)((i >> 1) & 1);
        mbar_wait(full0 + 8 * s, par);
        mbar_wait(aempty0 + 8 * s, par ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = tmem_base + A_COL0 + (uint32_t)(s * 2 * K), a_lo = a_hi + K;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t a0 = term == 0 ? a_lo : a_hi;
          const uint32_t b0 = term == 1 ? bL : bH;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks)
            umma_tf32_ts(tmem_base + s * N, a0 + 8 * ks, umma_desc(b0 + ks * 2 * B_LBO, B_LBO, B_SBO), IDESC,
                         (term | ks) != 0);
        }
        umma_commit(empty0 + 8 * s);
        umma_commit(afull0 + 8 * s);
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (as in linear_umma_pipe_kernel)
    const int q = warp - kPipeProducerWarps;
    float* stage = sOut + q * (32 * SP);
    constexpr int C4 = N / 4;
    constexpr int RPI = 32 / C4;
    const int c = 4 * (lane % C4);
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = ldg4(bias + c);
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int s = (int)(i & 1);
      mbar_wait(afull0 + 8 * s, (uint32_t)((i >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __syncwarp();
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(s * N);
#pragma unroll
      for (int u = 0; u < N / 8; ++u) {
        float v[8];
        tmem_ld8(taddr + 8 * u, v);
        st4(stage + lane * SP + 8 * u, make_float4(v[0], v[1], v[2], v[3]));
        st4(stage + lane * SP + 8 * u + 4, make_float4(v[4], v[5], v[6], v[7]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(aempty0 + 8 * s);
      __syncwarp();
      const int64_t tile = block_id + i * n_blocks;
#pragma unroll
      for (int it = 0; it < 32 / RPI; ++it) {
        const int r = it * RPI + lane / C4;
        const int64_t row = tile * BM + 32 * q + r;
        if (row < n_rows) {
          float4 o = *reinterpret_cast<const float4*>(stage + r * SP + c);
          o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          float* yp = Y + row * ldy + c;
          if (accumulate) {
            const float4 p = *reinterpret_cast<const float4*>(yp);
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          if (out_mask) {
            const float4 g = ldg4(out_mask + row * ldom + c);
            o.x = g.x > 0.f ? o.x : 0.f; o.y = g.y > 0.f ? o.y : 0.f; o.z = g.z > 0.f ? o.z : 0.f; o.w = g.w > 0.f ? o.w : 0.f;
          }
          st4(yp, o);
        }
      }
      __syncwarp();
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

template <int K, int N>
__global__ void __launch_bounds__(kPipeThreads, 1) linear_umma_ts_kernel(
    const float* __restrict__ X, int64_t ldx, int64_t n_rows, const float* __restrict__ W, int w_is_out_in,
    const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ Y, int64_t ldy,
    const float* __restrict__ out_mask, int64_t ldom) {
  linear_umma_ts_kernel_body<K, N>(X, ldx, n_rows, W, w_is_out_in, bias, relu, accumulate, Y, ldy, out_mask, ldom, (int)blockIdx.x, (int)gridDim.x);
}
// grouped form: CTA b works on the problem whose block range holds b, as block (b - first) of (last - first)
template <int K, int N>
__global__ void __launch_bounds__(kPipeThreads, 1) linear_umma_ts_kernel_grouped(const __grid_constant__ LinearGroup grp, int w_is_out_in, int relu,
                                                 int accumulate) {
  const int k = group_of_block(grp, (int)blockIdx.x);
  const peagnn_linear_problem_t& q = grp.p[k];
  linear_umma_ts_kernel_body<K, N>(q.X, q.ldx, q.n, q.W, w_is_out_in, q.bias, relu, accumulate, q.Y, q.ldy, q.out_mask, q.ldom,
      (int)blockIdx.x - grp.block_start[k], grp.block_start[k + 1] - grp.block_start[k]);
}

template <int K, int N>
static int launch_linear_umma_ts(const float* X, int64_t ldx, int64_t n, const float* W, int w_is_out_in,
                                 const float* bias, int relu, int accumulate, float* Y, int64_t ldy,
                                 const float* out_mask, int64_t ldom, cudaStream_t stream) {
  constexpr size_t smem = (size_t)2 * (N * K * 4) + (size_t)4 * 32 * (N + 4) * 4 + 128;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(linear_umma_ts_kernel<K, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const int64_t tiles = (n + 127) / 128;
  const int blocks = (int)imin64(tiles, (int64_t)kNumSMs);
  linear_umma_ts_kernel<K, N><<<blocks, kPipeThreads, smem, stream>>>(X, ldx, n, W, w_is_out_in, bias, relu, accumulate,
                                                                    Y, ldy, out_mask, ldom);
  return check_launch("peagnn_linear(umma ts)");
}

}  // namespace peagnn
