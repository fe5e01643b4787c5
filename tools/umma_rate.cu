// Round-2 diagnostic (written at the end of round 1, compile-checked only): issue-to-completion cost of
// tcgen05.mma.kind::tf32 M128 x N x K8 with the A operand in shared memory (SS) or in tensor memory (TS),
// operands resident, nothing else running.  Tests the hypothesis of profiles/r1_dense_tensor_cores.md that
// the SS form is bound by the shared-memory operand path.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I graph_recsys_benchmark_b200/csrc \
//        tools/umma_rate.cu -o /tmp/umma_rate && /tmp/umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "dense_umma.cuh"

using namespace peagnn;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* cycles) {
  constexpr int K = 64, KS = K / 8;
  constexpr uint32_t A_LBO = 16 * 128, B_LBO = (N / 8) * 128;
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                      // 128 x 64 tf32, core-matrix layout (contents irrelevant: zeros)
  uint8_t* sB = smem + 128 * K * 4;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 32) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (128 * K * 4 + N * K * 4) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (TS) {                                // A (64 columns) into TMEM columns [64, 128): zeros
    uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < K; c += 8) tmem_st8(tm + ((uint32_t)(32 * warp) << 16) + 64 + c, z);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(sA), b = smem_u32(sB);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        if (TS) umma_tf32_ts(tm, tm + 64 + 8 * ks, umma_desc(b + ks * 2 * B_LBO, B_LBO, 128), IDESC, 1);
        else umma_tf32(tm, umma_desc(a + ks * 2 * A_LBO, A_LBO, 128), umma_desc(b + ks * 2 * B_LBO, B_LBO, 128), IDESC, 1);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    if (blockIdx.x == 0) cycles[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
}

template <int N, bool TS>
static void run(const char* name, int ctas) {
  long long* d;
  cudaMalloc(&d, sizeof(long long));
  const size_t smem = 128 * 64 * 4 + N * 64 * 4;
  cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int reps = 2000;
  rate_kernel<N, TS><<<ctas, 128, smem>>>(reps, d);
  cudaDeviceSynchronize();
  rate_kernel<N, TS><<<ctas, 128, smem>>>(reps, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-28s ctas=%3d  %8.1f cycles per 128x%dx8 tf32 MMA   (%s)\n", name, ctas, (double)h / (reps * 8.0), N, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int ctas : {1, 148}) {
    run<64, false>("SS  A,B in shared memory", ctas);
    run<16, false>("SS  N = 16", ctas);
    run<64, true>("TS  A in tensor memory", ctas);
    run<16, true>("TS  N = 16", ctas);
  }
  return 0;
}
