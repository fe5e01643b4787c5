// Measurement probes (no product arithmetic): what the memory system gives a kernel shaped like the
// aggregation, so that bench.py can report the aggregation against a MEASURED ceiling instead of a quoted one.
//
//   peagnn_probe_gather : a warp streams a list of int32 row ids (coalesced, like CSR column indices) and
//     sums the addressed F-float rows of a table with 128-bit loads, G = F/4 lanes per row and 32/G rows
//     in flight per instruction - the access pattern of csr_rows_kernel without row bookkeeping, edge
//     scalars, folds or epilogues.  With a table that fits the 126 MB L2 (ML-25M: 291 k x 64 fp32 = 75 MB)
//     this is the L2 random-row-gather ceiling; bytes = n_idx * (4 + 4F).
#include "common.cuh"

namespace peagnn {

template <int F>
__global__ void __launch_bounds__(256) probe_gather_kernel(const float* __restrict__ table, unsigned ld,
                                                           const int32_t* __restrict__ idx, int64_t n_idx,
                                                           float* __restrict__ out) {
  constexpr int G = F / 4;            // lanes per row
  constexpr int S = 32 / G;           // rows per warp instruction
  const int lane = threadIdx.x & 31;
  const int sub = lane / G, part = lane % G;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t base = warp * 32; base < n_idx; base += n_warps * 32) {
    const int32_t mine = base + lane < n_idx ? __ldg(idx + base + lane) : 0;
#pragma unroll
    for (int it = 0; it < G; ++it) {
      const int32_t row = __shfl_sync(0xffffffffu, mine, it * S + sub);
      const float4 v = ldg4(table + row_off(row, ld) + 4 * part);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  // one store per lane so the loads cannot be optimised away
  st4(out + ((size_t)warp * 32 + lane) * 4, acc);
}

}  // namespace peagnn

using namespace peagnn;

extern "C" size_t peagnn_probe_out_floats(void) { return (size_t)kNumSMs * 8 * 256 * 4; }

extern "C" int peagnn_probe_gather(const float* table, int64_t ld, int32_t feat, const int32_t* idx, int64_t n_idx,
                                   float* out, peagnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PEAGNN_REQUIRE(table && idx && out && n_idx > 0 && ld % 4 == 0 && aligned16(table) && aligned16(out),
                 "peagnn_probe_gather: bad arguments");
  const int blocks = kNumSMs * 8;
  switch (feat) {
    case 16: probe_gather_kernel<16><<<blocks, 256, 0, stream>>>(table, (unsigned)ld, idx, n_idx, out); break;
    case 32: probe_gather_kernel<32><<<blocks, 256, 0, stream>>>(table, (unsigned)ld, idx, n_idx, out); break;
    case 64: probe_gather_kernel<64><<<blocks, 256, 0, stream>>>(table, (unsigned)ld, idx, n_idx, out); break;
    case 128: probe_gather_kernel<128><<<blocks, 256, 0, stream>>>(table, (unsigned)ld, idx, n_idx, out); break;
    default: PEAGNN_REQUIRE(false, "peagnn_probe_gather: feat %d not in {16,32,64,128}", feat);
  }
  return check_launch("peagnn_probe_gather");
}
