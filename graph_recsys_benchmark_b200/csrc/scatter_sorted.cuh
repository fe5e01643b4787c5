// Deterministic scatter-add of M gradient rows into a node table (replaces float atomics in K6):
// stable radix sort of (node id, slot), then the first slot of every run of equal ids sums the run
// in slot order and adds it to the table with a plain read-modify-write (one writer per node).
// Two runs of the same step therefore produce bit-identical gradients.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace peagnn {

constexpr int32_t kScatterSkip = 0x7fffffff;   // slots carrying this key are dropped

static inline size_t scatter_align(size_t x) { return (x + 255) & ~size_t(255); }

static inline size_t scatter_sort_temp_bytes(int64_t M) {
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)M, 0, 31);
  return tb;
}

// bytes of scratch behind the caller's key / row arrays: sorted keys, slot ids (in, out), cub temp
static inline size_t scatter_scratch_bytes(int64_t M) {
  return 3 * scatter_align((size_t)M * 4) + scatter_align(scatter_sort_temp_bytes(M)) + 256;
}

__global__ void scatter_iota_kernel(int32_t* __restrict__ ids, int64_t M) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) ids[i] = (int32_t)i;
}

// LANES = W / 4 float4 chunks per row; thread (i, c) owns chunk c of sorted position i.
__global__ void __launch_bounds__(256) scatter_runs_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ slots,
                                                           int64_t M, const float* __restrict__ rows, int W,
                                                           float* __restrict__ table, int64_t ld) {
  const int lanes = W >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t / lanes;
  const int c = (int)(t - i * lanes);
  if (i >= M) return;
  const int32_t k = keys[i];
  if (k == kScatterSkip || (i > 0 && keys[i - 1] == k)) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t j = i; j < M && keys[j] == k; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(rows + (size_t)slots[j] * W + 4 * c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float4* dst = reinterpret_cast<float4*>(table + (size_t)k * ld + 4 * c);
  float4 cur = *dst;
  cur.x += acc.x; cur.y += acc.y; cur.z += acc.z; cur.w += acc.w;
  *dst = cur;
}

// keys [M] int32 (node id or kScatterSkip), rows [M, W] fp32 (W % 4 == 0), scratch >= scatter_scratch_bytes(M)
static inline int scatter_rows_sorted(const int32_t* keys, const float* rows, int W, int64_t M, float* table, int64_t ld,
                                      void* scratch, cudaStream_t stream, const char* what) {
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(scratch) + 255) & ~uintptr_t(255));
  const size_t seg = scatter_align((size_t)M * 4);
  int32_t* k_out = reinterpret_cast<int32_t*>(ws);
  int32_t* s_in = reinterpret_cast<int32_t*>(ws + seg);
  int32_t* s_out = reinterpret_cast<int32_t*>(ws + 2 * seg);
  void* temp = ws + 3 * seg;
  size_t temp_bytes = scatter_sort_temp_bytes(M);
  scatter_iota_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(s_in, M);
  int rc = check_launch(what);
  if (rc) return rc;
  cudaError_t ce = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, k_out, s_in, s_out, (int)M, 0, 31, stream);
  if (ce != cudaSuccess) {
    set_error("%s(sort): %s", what, cudaGetErrorString(ce));
    return PEAGNN_ERR_CUDA;
  }
  const int64_t threads = M * (W >> 2);
  scatter_runs_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(k_out, s_out, M, rows, W, table, ld);
  return check_launch(what);
}

}  // namespace peagnn
